#!/usr/bin/env python
"""bench.py -- headline benchmark of the LGCA hot path on B200 (site updates per second).

Contract:  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload periodic|karman]
For N > 1 launch under torchrun (one rank per GPU); rank 0 prints ONE JSON line.

Workload (BASELINE.json metric "FHP-III site updates/s at 1/2/4/8 B200"): config C5,
lgca-periodic FHP-III, 32768 x 32768 sites PER GPU stacked in y (weak scaling, row strips, halo ring),
all-fluid periodic torus, occupancy P = 1/7, chirality P = 1/2 generated on the device from a
counter-based hash.  One bench "step" = 100 lattice updates (k-step temporal blocking
inside) followed by the coarse-grained post-process of the snapshot and the device->host read of the
coarse fields (C5: "coarse-grained output every 100 steps").  `--workload karman` runs config C3
(FHP-III 16384 x 8192, walls + cylinder) instead; its N=1 numbers are also attached to the default line
under "karman".

value    : whole-job site updates/s with the lattice resident in HBM (CUDA events, max over ranks).
e2e      : the same metric through the reference-facing call sequence with HOST buffers:
           copy_data_to_device (pinned reference-layout state bytes) -> 100 x collide_and_propagate ->
           snapshot + post_process (coarse fields to host) -> copy_data_from_device, per step.
roofline : dominant kernel = step_wave_kernel (k fused steps per launch); algorithmic bytes per launch =
           sites * (2*NUM_DIR + mask planes)/8 * k, divided by the average launch duration measured with
           CUDA events on the engine's compute stream; peak = MEASURED_PEAKS.json hbm_gbs (burst copy).
           With k-step temporal blocking real DRAM traffic is 1/k of the algorithmic bytes, so frac may
           exceed 1; `traffic` is the ncu dram bytes per launch from profiles/.
cpu_baseline / --impl reference : the UNMODIFIED reference OMP path (oracle/_ref, built from
           /root/reference by oracle/Makefile) on all host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UPDATES_PER_STEP = 100
WORKLOADS = {
    # name: (model, dim_x, rows per GPU, bc, cg_radius, description)
    "periodic": ("FHP_III", 32768, 32768, "periodic", 16,
                 "lgca-periodic FHP-III 32768x32768 per GPU (BASELINE config C5), weak scaling in y"),
    "karman": ("FHP_III", 16384, 8192, "karman", 16,
               "lgca-karman FHP-III 16384x8192, pipe walls + cylinder, x-periodic (BASELINE config C3)"),
    # C4 is a STRONG-scaling case: the 65536 x 32768 box is split over the N GPUs (rows per GPU = 32768 / N)
    "box": ("FHP_II", 65536, 32768, "reflecting_back", 16,
            "lgca-box FHP-II 65536x32768, bounce-back frame, row strips over N GPUs (BASELINE config C4), strong scaling"),
}
CPU_SAMPLE = {"periodic": (4096, 4096), "karman": (4096, 2048), "box": (4096, 4096)}
STRONG = {"box"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_stamp():
    """Hash of the sources the fused-step kernel is compiled from: ncu traffic numbers in profiles/traffic.json are
    only attached to a bench line when they were captured from exactly this kernel."""
    import hashlib
    h = hashlib.sha256()
    for f in ("lgca_step_wave.cu", "lgca_collide.cuh", "lgca_common.cuh", "lgca_wave_pins.h"):
        h.update(open(os.path.join(ROOT, "lgca_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(key):
    """(bytes per launch, note) from profiles/traffic.json -- an ncu `--set full` capture of the same kernel build."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, "no ncu capture under profiles/"
    try:
        t = json.load(open(tp))
    except Exception as ex:
        return None, "unreadable traffic.json: %r" % ex
    ent = t.get("entries", {}).get(key)
    if ent is None:
        return None, "no ncu capture for %s" % key
    if ent.get("kernel_stamp") != kernel_source_stamp():
        return None, "stale: ncu capture %s was taken from kernel sources %s, this build is %s" % (
            key, ent.get("kernel_stamp"), kernel_source_stamp())
    return float(ent["dram_bytes_per_launch"]), "ncu dram__bytes_read.sum + dram__bytes_write.sum, capture %s" % ent.get("source", "?")


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons DURING the timed region (NVML, 20 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.err = index, [], False, None
        self.nv = self.h = self.get_reasons = None
        self.max_sm = None
        try:  # NVML is initialised BEFORE the timed region: the region may be shorter than nvmlInit()
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as ex:  # pragma: no cover
            self.err = repr(ex)

    def sample(self):
        nv, h = self.nv, self.h
        self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                             int(self.get_reasons(h))))

    def run(self):
        if self.nv is None:
            return
        try:
            while not self.stop_flag:
                self.sample()
                time.sleep(0.005)
        except Exception as ex:  # pragma: no cover
            self.err = repr(ex)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        # NVML clocks-event-reason bits
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        sm = sorted(s[0] for s in self.samples)
        seen = 0
        for s_ in self.samples:
            seen |= s_[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "power_w_max": max(s_[1] for s_ in self.samples),
                "reasons": sorted(v for b_, v in bits.items() if seen & b_), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own OMP path on the host cores
# ------------------------------------------------------------------------------------------------------
def reference_setup(workload):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_checkers as cc
    model, _, _, bc, cg, _ = WORKLOADS[workload]
    dx, dy = CPU_SAMPLE[workload]
    cores = os.cpu_count() or 1
    if cc.ref_available():
        kind = "reference"
        cc._ref_lib().lgca_ref_set_threads(1)  # deterministic construction (rand() inside omp loops)
        if workload == "periodic":
            lat = cc.Ref(model, "periodic", dx - 1, 0.2, cg, threads=1)
        else:
            lat = cc.Ref(model, "periodic", dx - 1, 0.2, cg, threads=1, dims=(dx, dy))
        lat.apply_bc(bc)
        lat.init("random")
        lat.set_threads(cores)
    else:  # the plain-C port (same algorithm, OpenMP over cells)
        kind = "port"
        lat = cc.Oracle(model, dims=(dx, dy), cg=cg)
        lat.apply_bc(bc)
        lat.init("random")
    return lat, kind, cores, (dx, dy)


def time_reference(workload, updates, warm=1):
    lat, kind, cores, (dx, dy) = reference_setup(workload)
    lat.step(warm)
    t0 = time.perf_counter()
    lat.step(updates)
    dt = time.perf_counter() - t0
    val = dx * dy * updates / dt
    sample = "%s %dx%d %s, %d updates after %d warm-up (%.1f s)" % (WORKLOADS[workload][0], dx, dy, WORKLOADS[workload][3],
                                                                      updates, warm, dt)
    return val, dt, {"value": val, "unit": "site updates/s", "cores": cores, "kind": kind, "sample": sample}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model, dx, rows, bc, cg, desc = WORKLOADS[args.workload]
    ups = 2  # lattice updates of the bounded sample per "step"
    lat, kind, cores, (sx, sy) = reference_setup(args.workload)
    for _ in range(args.warmup):
        lat.step(ups)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lat.step(ups)
    dt = time.perf_counter() - t0
    val = sx * sy * ups * args.steps / dt
    line = {
        "impl": "reference", "metric": "FHP-III site updates/s", "value": val, "unit": "site updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (byte per cell, bit per direction)",
        "data": "synthetic",
        "config": {"workload": desc, "global_lattice": [dx, rows * args.gpus],
                   "note": "reference CPU path (OMP_Lattice::collide_and_propagate) on host cores; each step = %d updates of a "
                           "bounded %dx%d sample of the workload (same model/BC/init)" % (ups, sx, sy)},
        "cpu_baseline": {"value": val, "unit": "site updates/s", "cores": cores, "kind": kind,
                         "sample": "%s %dx%d %s, %d steps x %d updates" % (model, sx, sy, bc, args.steps, ups)},
        "e2e": {"value": val, "unit": "site updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------
def build_engine(workload, rank, world, local_rank, k_fuse):
    import lgca_b200
    from lgca_b200.ring import partition_rows
    model, dx, rows, bc, cg, _ = WORKLOADS[workload]
    dim_y = rows if workload in STRONG else rows * world
    if world == 1:
        e = lgca_b200.Engine(model, dx, dim_y, cg_radius=cg, bf_dir="x" if workload == "karman" else 0, device=local_rank,
                             k_fuse=k_fuse, flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
    else:
        y0, yr = partition_rows(dim_y, world, 2 * cg)[rank]
        e = lgca_b200.Engine(model, dx, dim_y, cg_radius=cg, device=local_rank, k_fuse=k_fuse, y_begin=y0, y_rows=yr,
                             flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
    e.apply_bc_device(bc)
    e.init_random_device(seed=1)
    return e


class VtiWriter(threading.Thread):
    """Writes the coarse fields of every bench step as `mean_res_<step>_r<rank>.vti` in the background (config C5:
    "coarse-grained .vti output every 100 steps"); double-buffered so that the write of step i overlaps step i+1."""

    def __init__(self, directory, rank, cdx, cdy, origin_y):
        super().__init__(daemon=True)
        import queue
        self.q, self.dir, self.rank, self.cdx, self.cdy, self.oy = queue.Queue(maxsize=1), directory, rank, cdx, cdy, origin_y
        self.bytes = 0
        self.files = 0
        self.error = None
        self.start()

    def run(self):
        from lgca_b200.vti import write_mean_vti
        while True:
            item = self.q.get()
            if item is None:
                self.q.task_done()
                return
            step, fields = item
            try:
                path = os.path.join(self.dir, "mean_res_%d_r%d.vti" % (step, self.rank))
                write_mean_vti(path, self.cdx, self.cdy, fields["mean_density"], fields["mean_momentum"], self.oy)
                self.bytes += os.path.getsize(path)
                self.files += 1
                os.unlink(path)  # the benchmark keeps no output
            except Exception as ex:  # never leave the main thread waiting on a dead writer
                self.error = repr(ex)
            finally:
                self.q.task_done()

    def submit(self, step, fields):
        self.q.put((step, fields))  # blocks while the previous file is still being written

    def drain(self):
        self.q.join()


def bench_step(ring, e, coarse_bufs, writer, step_no):
    """One bench step with the lattice resident in HBM: 100 updates + snapshot + coarse post-process of a snapshot +
    read-back of the coarse fields + .vti write.

    Software-pipelined like the reference's viewers, which simulate and post-process concurrently
    (apps/pipe/pipe_viewer.cpp:100-184): the 100 updates of this step are enqueued first, then the host post-processes
    the PREVIOUS snapshot (device snapshot buffer insulates it) while the GPU keeps stepping, then the new state is
    snapshotted.  Every step performs exactly one of each operation."""
    ring.step(UPDATES_PER_STEP)        # asynchronous
    out = coarse_bufs[0]
    if writer is not None:
        writer.drain()                 # the previous file is done before its buffer is refilled
    e.post_process(cell=False, mean=True, exact=False, out=out)   # of the snapshot taken at the end of the previous step
    if writer is not None:
        writer.submit(step_no * UPDATES_PER_STEP, out)
    e.snapshot()                       # stream-ordered after the 100 updates (and after the post-process read)


def info_y_begin(e):
    return int(e.info().y_begin)


def multi_gpu_parity(rank, world, local_rank, device, k_fuse, native):
    """Decomposition invariance on PHYSICAL GPUs, inside the bench run: a small global lattice (FHP-III 8192 x 2048*N,
    device counter-hash init keyed on the global cell) is advanced 13*k+1 steps as N row strips through the halo ring --
    once all fluid, once with the Karman walls + cylinder, with snapshots in between (the strips' plane sets rotate) --
    and rank 0 advances the same lattice as ONE whole lattice on its GPU.  Every strip's bytes must equal the matching
    rows of the single-GPU result (blake2b digests, all-gathered).  Returns the report; the caller fails the run on a
    mismatch."""
    import hashlib
    import torch.distributed as dist
    import lgca_b200
    from lgca_b200.ring import Ring, partition_rows
    dx, rows, cg = 8192, 2048, 16
    dim_y = rows * world
    report = {"ok": True, "lattice": [dx, dim_y], "variants": []}
    for bc in ("periodic", "karman"):
        y0, yr = partition_rows(dim_y, world, 2 * cg)[rank]
        e = lgca_b200.Engine("FHP_III", dx, dim_y, cg_radius=cg, device=local_rank, k_fuse=k_fuse, y_begin=y0, y_rows=yr,
                             flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
        e.apply_bc_device(bc)
        e.init_random_device(seed=7)
        ring = Ring(e, rank, world, device=device, native=native)
        ring.start()
        k = e.steps_per_exchange()
        steps = 13 * k + 1
        ring.step(5 * k)
        e.snapshot()
        ring.step(steps - 5 * k)
        e.snapshot()
        mine = hashlib.blake2b(e.download().tobytes(), digest_size=16).hexdigest()
        particles = e.count_particles()
        e.sync()
        dist.barrier()
        e.ring_disconnect()    # every rank drops its mappings of the neighbours' memory before anyone frees it
        dist.barrier()
        e.close()
        digests = [None] * world
        dist.all_gather_object(digests, (mine, particles))
        if rank == 0:
            one = lgca_b200.Engine("FHP_III", dx, dim_y, cg_radius=cg, device=local_rank, k_fuse=k_fuse,
                                   flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
            one.apply_bc_device(bc)
            one.init_random_device(seed=7)
            one.step(steps)
            whole = one.download()
            want_particles = one.count_particles()
            one.close()
            parts = partition_rows(dim_y, world, 2 * cg)
            want = [hashlib.blake2b(whole[a * dx:(a + n) * dx].tobytes(), digest_size=16).hexdigest() for a, n in parts]
            ok = [d[0] for d in digests] == want and sum(d[1] for d in digests) == want_particles
            report["variants"].append({"bc": bc, "steps": steps, "k_fuse": k, "ok": bool(ok),
                                       "digest": hashlib.blake2b(whole.tobytes(), digest_size=16).hexdigest(),
                                       "strips_equal": [a == b for a, b in zip([d[0] for d in digests], want)]})
            report["ok"] = report["ok"] and bool(ok)
    flag = [report["ok"] if rank == 0 else None]
    dist.broadcast_object_list(flag, src=0)
    report["ok"] = bool(flag[0])
    return report


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import lgca_b200
    from lgca_b200.ring import Ring

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lgca_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL writes its debug output (the image sets NCCL_DEBUG=VERSION, and the
        # version banner is printed at WARN level too) to stdout unless told otherwise
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)

    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(rank, world, local_rank, device, args.k_fuse, native=not args.nccl_halo)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"multi_gpu_parity": parity, "error": "row strips over %d GPUs differ from the single-GPU result" % world}))
            dist.destroy_process_group()
            return 3

    model, dx, rows, bc, cg, desc = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    if strong:
        rows //= world
    sites_rank = dx * rows
    sites_total = sites_rank * world
    e = build_engine(args.workload, rank, world, local_rank, args.k_fuse)
    ring = Ring(e, rank, world, device=device, native=not args.nccl_halo)
    ring.start()
    info = e.info()
    particles0 = e.count_particles()
    coarse_bufs = [{}, {}]
    coarse = coarse_bufs[0]
    import tempfile
    vti_dir = tempfile.mkdtemp(prefix="lgca_vti_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    writer = None if args.no_vti else VtiWriter(vti_dir, rank, dx // (2 * cg), rows // (2 * cg), info_y_begin(e) // (2 * cg))

    def barrier():
        e.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ----------------------------------------------------------------
    step_no = 0
    e.snapshot()
    for _ in range(max(args.warmup, 3)):
        step_no += 1
        bench_step(ring, e, coarse_bufs, writer, step_no)
    if writer:
        writer.drain()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.ExternalStream(e.compute_stream(), device=device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = e.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        step_no += 1
        bench_step(ring, e, coarse_bufs, writer, step_no)
    if writer:
        writer.drain()  # the last file is on disk before the clock stops
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = e.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(2)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = sites_total * UPDATES_PER_STEP * args.steps / (ms * 1e-3)

    # ---- end to end through host buffers -----------------------------------------------------------------
    host_state = lgca_b200.capi.PinnedArray((sites_rank,), "uint8")
    e.download(host_state.array)
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    t_up = t_down = 0.0
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        e.upload(state=host_state.array)             # copy_data_to_device()
        t_up += time.perf_counter() - ta
        # ghost rows of the freshly uploaded strips: stream-ordered republish (the neighbours acknowledge on the
        # device that they no longer read the old rows) -- no host barrier, ranks drift freely
        ring.republish()
        ring.step(UPDATES_PER_STEP)                  # 100 x collide_and_propagate()
        e.snapshot()                                 # copy_data_to_output_buffer()
        if writer:
            writer.drain()
        e.post_process(cell=False, mean=True, exact=False, out=coarse)  # post_process() -> host coarse fields
        if writer:
            writer.submit(0, coarse)                 # IoVti::write of the coarse fields
        ta = time.perf_counter()
        e.download(host_state.array)                 # copy_data_from_device()
        t_down += time.perf_counter() - ta
    if writer:
        writer.drain()
    barrier()
    e2e_s = time.perf_counter() - t0
    # per-rank PCIe rates of the two big copies (download includes waiting for the 100 updates to finish)
    h2d_gbs = sites_rank * e2e_steps / max(t_up, 1e-9) / 1e9
    if world > 1:
        t = torch.tensor([h2d_gbs, -h2d_gbs], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        h2d_range = [float(t[0].item()), float(-t[1].item())]
    else:
        h2d_range = [h2d_gbs, h2d_gbs]
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = sites_total * UPDATES_PER_STEP * e2e_steps / e2e_s
    coarse_bytes = sum(v.nbytes for v in coarse.values())
    # what the box's host links deliver when all ranks copy at once (plain cudaMemcpy of the same pinned buffer, no
    # pack/unpack kernels): the ceiling of the e2e figure above
    link_probe = {}
    try:
        src = torch.from_numpy(host_state.array)
        dev = torch.empty(sites_rank, dtype=torch.uint8, device=device)
        for name, a, b in (("h2d", dev, src), ("d2h", src, dev)):
            barrier()
            torch.cuda.synchronize()
            tp = time.perf_counter()
            a.copy_(b, non_blocking=True)
            torch.cuda.synchronize()
            gbs = sites_rank / (time.perf_counter() - tp) / 1e9
            if world > 1:
                t = torch.tensor([gbs, -gbs], device=device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                link_probe[name + "_gbs_per_rank_min_max"] = [float(t[0].item()), float(-t[1].item())]
            else:
                link_probe[name + "_gbs_per_rank_min_max"] = [gbs, gbs]
        link_probe["note"] = "all %d ranks copy %d MiB at once between pinned host memory and HBM (cudaMemcpyAsync)" % (world, sites_rank >> 20)
        del dev
    except Exception as ex:
        link_probe = {"error": repr(ex)}
    host_state.free()

    # ---- mass conservation over the whole job (global particle count) ---------------------------------------
    particles1 = e.count_particles()
    if world > 1:
        t = torch.tensor([particles0, particles1], device=device, dtype=torch.int64)
        dist.all_reduce(t)
        particles0, particles1 = int(t[0].item()), int(t[1].item())
    conserved = particles0 == particles1

    # ---- dominant kernel: average launch duration of the fused-step kernel, CUDA events on its stream ------
    # (last: on a strip this skips the halo exchange and leaves the edge rows stale)
    k = info.k_fuse
    e.sync()
    e.timed_kernel(5)
    kms = min(e.timed_kernel(25) for _ in range(3))
    # the ring couples the ranks (every block waits for both neighbours' ghost rows), so the job runs at the pace of the
    # slowest GPU: the spread of the stand-alone kernel time over the ranks shows how much of the N-GPU loss is silicon
    kms_range = [kms, kms]
    if world > 1:
        t = torch.tensor([kms, -kms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        kms_range = [float(t[0].item()), float(-t[1].item())]
    alg_bytes = sites_rank * info.bytes_per_site_step_x8 / 8.0 * k
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    traffic, traffic_note = measured_traffic("%s_k%d" % (args.workload, k))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_note, "kernel": "step_wave_kernel<FHP_II rule, K=%d>" % k,
                "launch_ms": kms, "launch_ms_rank_min_max": kms_range, "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "dram_frac": (traffic / (kms * 1e-3) / 1e9 / peak) if traffic else None,
                "limiter": "integer pipe + issue slots (LOP3/SHF/SHFL; sm__pipe_alu and issue_active in profiles/), not DRAM: dram_frac is the share of the copy peak the kernel really moves",
                "note": "algorithmic bytes = sites*(2*NUM_DIR+masks)/8 per step x k fused steps per launch; "
                        "real DRAM traffic (`traffic`, ncu) is ~1/k of it (temporal blocking), so frac exceeds 1; "
                        "dram_frac = traffic / launch time / peak",
                "launch_timing": "CUDA events on the kernel's stream around 25 consecutive launches / 25; consecutive launches of a "
                                 "call are chained (the next launch fills the warp slots the previous one frees, ordered by "
                                 "per-chunk completion counters), so this is the effective time per launch, not an isolated one"}

    line = None
    if rank == 0:
        line = {
            "metric": "%s site updates/s" % model.replace("_", "-"), "value": value, "unit": "site updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u32 bit-planes (1 bit per site and direction)",
            "data": "synthetic",
            "config": {"workload": desc, "global_lattice": [dx, rows * world], "sites_per_gpu": sites_rank,
                       "updates_per_step": UPDATES_PER_STEP, "k_fuse": k, "parallelism": "single GPU" if world == 1 else "row strips x%d, %s" % (
                           world, "halo ring: NCCL send/recv" if args.nccl_halo else
                           "halo ring: peer stores into the neighbours' ghost rows over NVLink (CUDA IPC) + device epoch flags, "
                           "edge tiles wait in-kernel"),
                       "cache": "inputs larger than L2 (%.0f MB of bit-planes per GPU vs 126 MB L2)" % (
                           sites_rank * info.num_planes / 8 / 1e6) if sites_rank * info.num_planes / 8 > 200e6 else
                       "lattice (%.0f MB) is L2-resident" % (sites_rank * info.num_planes / 8 / 1e6),
                       "step": "100 updates + snapshot + coarse post-process + D2H of coarse fields + .vti write of them "
                               "(%s)" % ("disabled" if args.no_vti else "%d files, %.1f MB each, background thread" % (
                                   writer.files, writer.bytes / max(writer.files, 1) / 1e6))},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "site updates/s", "h2d_bytes_per_step": sites_rank * world,
                    "d2h_bytes_per_step": (sites_rank + coarse_bytes) * world, "steps": e2e_steps,
                    "h2d_gbs_per_rank_min_max": h2d_range, "host_link_probe": link_probe,
                    "path": "upload(pinned state bytes) -> [strips: stream-ordered republish of the ghost rows, no barrier] -> "
                            "100 steps -> snapshot+post_process -> download"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "particles_conserved": bool(conserved),
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
    # ---- extras on rank 0 at N=1: CPU baseline and the Karman (C3) measurement ----------------------------
    if world == 1 and line is not None:
        if not args.no_cpu_baseline:
            try:
                _, _, cb = time_reference(args.workload, updates=args.cpu_updates)
                line["cpu_baseline"] = cb
            except Exception as ex:  # the checker is optional for the measurement itself
                line["cpu_baseline"] = {"value": None, "unit": "site updates/s", "cores": os.cpu_count(), "kind": "unavailable",
                                        "sample": repr(ex)}
        if args.workload == "periodic" and not args.no_karman:
            teardown(e, world)
            line["karman"] = karman_extra(args, peak)
            line["diffusion_hpp"] = hpp_extra(args, peak)
            line["resident"] = resident_extra(args)
            if not args.no_app_tick:
                line["app_tick"] = app_tick_extra()
    if args.workload == "periodic" and not args.no_box:
        # C4 at this N (all ranks take part); the main engine goes first: 65536 x 32768 needs the memory
        teardown(e, world)
        box = box_extra(args, rank, world, local_rank, device, peak)
        if line is not None:
            line["box"] = box
    if writer:
        writer.q.put(None)
        if writer.error:
            raise SystemExit("bench.py: .vti writer failed: %s" % writer.error)
    try:
        os.rmdir(vti_dir)
    except OSError:
        pass
    if line is not None:
        print(json.dumps(line))
    teardown(e, world)
    if world > 1:
        dist.destroy_process_group()
    return 0


def teardown(e, world):
    """Close a strip: every rank drops its mappings of the neighbours' memory before anyone frees it."""
    if getattr(e, "h", None) is None:
        return
    e.sync()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        e.ring_disconnect()
        dist.barrier()
    e.close()


def box_extra(args, rank, world, local_rank, device, peak):
    """BASELINE config C4 (lgca-box FHP-II 65536 x 32768, bounce-back frame) cut into row strips over the N GPUs of the
    run: STRONG scaling, device-timed (CUDA events on every rank's compute stream, max over ranks).  Attached to the
    default line at every N, so the driver's 1/2/4/8 runs carry the C4 curve."""
    import torch
    import torch.distributed as dist
    from lgca_b200.ring import Ring
    model, dx, rows, bc, cg, desc = WORKLOADS["box"]
    e = build_engine("box", rank, world, local_rank, args.k_fuse)
    ring = Ring(e, rank, world, device=device, native=not args.nccl_halo)
    ring.start()
    k = e.steps_per_exchange() if world > 1 else e.info().k_fuse
    n0 = e.count_particles()
    stream = torch.cuda.ExternalStream(e.compute_stream(), device=device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ring.step(4 * k)
    best = None
    for _ in range(3):
        e.sync()
        if world > 1:
            dist.barrier()
        ev0.record(stream)
        ring.step(10 * k)
        ev1.record(stream)
        e.sync()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None else min(best, ms)
    n1 = e.count_particles()
    if world > 1:
        t = torch.tensor([n0, n1], device=device, dtype=torch.int64)
        dist.all_reduce(t)
        n0, n1 = int(t[0].item()), int(t[1].item())
    info = e.info()
    value = dx * rows * 10 * k / (best * 1e-3)
    alg = value * info.bytes_per_site_step_x8 / 8.0 / 1e9
    teardown(e, world)
    return {"workload": desc, "value": value, "unit": "site updates/s", "n_gpus": world, "scaling": "strong",
            "per_gpu": value / world, "k_fuse": k, "us_per_update": best * 1e3 / (10 * k), "rows_per_gpu": rows // world,
            "roofline_achieved_gbs_per_gpu": alg / world, "roofline_frac_per_gpu": alg / world / peak,
            "particles_conserved": bool(n0 == n1)}


def app_tick_extra():
    """The reference's REAL app schedule end to end (apps/karman/karman_viewer.cpp:100-184: mean velocity -> body force ->
    5 steps -> snapshot -> post-process per tick), 1000 steps: the headless C++ apps on B200_Lattice (host rand() stream,
    order-exact host mean velocity, exact body force -- the bit-exact path the parity tests check) against the same
    schedule on the unmodified reference (oracle/_ref) on the host cores (bounded number of ticks)."""
    import re
    out = {}
    bin_dir = os.path.join(ROOT, "lgca_b200", "host", "bin")
    cases = [("pipe_c1", "lgca-pipe", ["--model", "FHP_I"], ("FHP_I", "pipe", 80.0, 0.3, 10, "pipe"), 40),
             ("karman_default", "lgca-karman", [], ("FHP_III", "karman", 80.0, 0.3, 20, "karman"), 8)]
    for key, app, extra, ref_cfg, ref_ticks in cases:
        ent = {}
        exe = os.path.join(bin_dir, app)
        if os.path.exists(exe):
            # best of three runs (all wall times are kept in `wall_s_runs`): the tick loop is 0.2-0.9 s of many short device
            # calls next to this process's own CUDA context, and single runs have shown one-off stalls of the body-force calls
            m, p, walls = None, None, []
            for _ in range(3):
                p = subprocess.run([exe, "--steps", "1000", "--quiet"] + extra, capture_output=True, text=True)
                mm = re.search(r"Tick loop: (\S+) s wall for (\d+) steps in (\d+) ticks \(mean velocity (\S+) s, body force (\S+) s, "
                               r"stepping (\S+) s, snapshot \+ post-process (\S+) s\); (\S+) site updates/s", p.stdout)
                if not mm or p.returncode != 0:
                    m = None
                    break
                walls.append(float(mm.group(1)))
                if m is None or float(mm.group(1)) < float(m.group(1)):
                    m = mm
            if m and p.returncode == 0:
                ent["b200"] = {"steps": int(m.group(2)), "ticks": int(m.group(3)), "wall_s": float(m.group(1)),
                               "mean_velocity_s": float(m.group(4)), "body_force_s": float(m.group(5)), "stepping_s": float(m.group(6)),
                               "post_process_s": float(m.group(7)), "site_updates_per_s": float(m.group(8)),
                               "ms_per_tick": float(m.group(1)) / int(m.group(3)) * 1e3, "wall_s_runs": walls}
            else:
                ent["b200"] = {"error": (p.stdout + p.stderr)[-300:]}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import cpu_checkers as cc
            if cc.ref_available():
                model, case, Re, Ma, cg, bc = ref_cfg
                cores = os.cpu_count() or 1
                r = cc.Ref(model, case, Re, Ma, cg, threads=1)
                r.apply_bc(bc)
                r.init("random")
                r.set_threads(cores)
                r.snapshot(); r.post_process()
                forcing, u = r.initial_forcing(), r.u
                t0 = time.perf_counter()
                for _ in range(ref_ticks):
                    r.set_threads(1)       # the reference's racy counter needs one thread to be deterministic
                    mv = r.mean_velocity()
                    r.set_threads(cores)
                    if mv[0] < u:
                        if mv[0] > 0.9 * u:
                            forcing = r.equilibrium_forcing()
                        r.body_force(forcing)
                    r.step(5)
                    r.snapshot(); r.post_process()
                dt = time.perf_counter() - t0
                ent["reference"] = {"ticks": ref_ticks, "steps": 5 * ref_ticks, "wall_s": dt, "ms_per_tick": dt / ref_ticks * 1e3,
                                    "site_updates_per_s": r.num_cells * 5 * ref_ticks / dt, "cores": cores}
                r.close()
        except Exception as ex:
            ent["reference"] = {"error": repr(ex)}
        if "b200" in ent and "reference" in ent and "ms_per_tick" in ent["b200"] and "ms_per_tick" in ent["reference"]:
            ent["tick_speedup"] = ent["reference"]["ms_per_tick"] / ent["b200"]["ms_per_tick"]
        out[key] = ent
    return out


def resident_extra(args):
    """Lattices that fit on chip (the reference's own app sizes): device time per update of the SM-resident kernel
    (library default there) against the HBM-streaming wavefront kernel, for one viewer tick (5 steps) and a long call."""
    import lgca_b200
    from lgca_b200.capi import FLAG_NO_CELL_FIELDS, FLAG_NO_RESIDENT
    out = {}
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    for key, model, dx, dy, bc in (("pipe_c1_fhp1_1400x700", "FHP_I", 1400, 700, "pipe"), ("pipe_default_fhp3_1480x740", "FHP_III", 1480, 740, "pipe"),
                                   ("karman_default_fhp3_4400x2200", "FHP_III", 4400, 2200, "karman"),
                                   ("hpp_2048x2048", "HPP", 2048, 2048, "periodic")):
        ent = {}
        for kernel, flags in (("resident", 0), ("wave", FLAG_NO_RESIDENT)):
            e = lgca_b200.Engine(model, dx, dy, device=dev, flags=flags | FLAG_NO_CELL_FIELDS)
            e.apply_bc_device(bc)
            e.init_random_device(1)
            res = {}
            for n, reps in ((5, 100), (1000, 3)):
                e.timed_steps(n)
                best = min(sum(e.timed_steps(n) for _ in range(reps)) / reps for _ in range(3))
                res["us_per_update_call_of_%d" % n] = best * 1e3 / n
            res["site_updates_per_s"] = dx * dy / (res["us_per_update_call_of_1000"] * 1e-6)
            ent[kernel] = res
            e.close()
        ent["speedup"] = ent["wave"]["us_per_update_call_of_1000"] / ent["resident"]["us_per_update_call_of_1000"]
        out[key] = ent
    return out


def karman_extra(args, peak):
    """Config C3 on one GPU: device-resident throughput and kernel roofline (attached to the default line)."""
    e = build_engine("karman", 0, 1, int(os.environ.get("LOCAL_RANK", "0")), args.k_fuse)
    model, dx, rows, bc, cg, desc = WORKLOADS["karman"]
    info = e.info()
    k = info.k_fuse
    e.timed_steps(k * 50)
    ms = min(e.timed_steps(k * 100) for _ in range(3)) / 100
    alg = dx * rows * info.bytes_per_site_step_x8 / 8.0 * k
    out = {"workload": desc, "value": dx * rows * k / (ms * 1e-3), "unit": "site updates/s", "k_fuse": k, "launch_ms": ms,
           "roofline_achieved_gbs": alg / (ms * 1e-3) / 1e9, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak,
           "cache": "bit-planes 117 MB: partly L2-resident"}
    e.close()
    return out


def hpp_extra(args, peak):
    """Config C2 on one GPU (HPP 4096 x 4096 periodic; throughput does not depend on the particle pattern, so the
    lattice is filled by the device initialiser instead of the diffusion disc)."""
    import lgca_b200
    e = lgca_b200.Engine("HPP", 4096, 4096, device=int(os.environ.get("LOCAL_RANK", "0")), k_fuse=args.k_fuse,
                         flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
    e.apply_bc_device("periodic")
    e.init_random_device(seed=1)
    info = e.info()
    k = info.k_fuse
    e.timed_kernel(200)
    ms = min(e.timed_kernel(1000) for _ in range(3))
    alg = 4096 * 4096 * info.bytes_per_site_step_x8 / 8.0 * k
    out = {"workload": "lgca-diffusion HPP 4096x4096 periodic (BASELINE config C2)", "value": 4096 * 4096 * k / (ms * 1e-3),
           "unit": "site updates/s", "k_fuse": k, "launch_ms": ms, "roofline_achieved_gbs": alg / (ms * 1e-3) / 1e9,
           "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak, "cache": "bit-planes 8 MB: L2-resident"}
    e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="periodic", choices=sorted(WORKLOADS))
    ap.add_argument("--k-fuse", type=int, default=0)
    ap.add_argument("--cpu-updates", type=int, default=12, help="updates of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-karman", action="store_true")
    ap.add_argument("--no-vti", action="store_true", help="skip the .vti write of the coarse fields")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU decomposition-invariance check (N > 1)")
    ap.add_argument("--no-box", action="store_true", help="skip the C4 box extra")
    ap.add_argument("--no-app-tick", action="store_true", help="skip the end-to-end app schedule extra")
    ap.add_argument("--nccl-halo", action="store_true", help="move ghost rows with NCCL send/recv instead of the native peer-store ring")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
