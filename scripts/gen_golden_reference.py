#!/usr/bin/env python
"""Generates tests/golden/reference_runs.json from the UNMODIFIED reference (oracle/_ref, compiled in place from
/root/reference by oracle/Makefile).  Run in the build container (needs /root/reference):

    python scripts/gen_golden_reference.py [--only karman_default,...]

Contents (all hashes FNV-1a-64 of the reference-layout state bytes):
  * karman_default : the north-star target as written -- lgca-karman at the app's defaults
        ("karman", Re 80, Ma 0.3, cg 20, FHP-III -> 4400 x 2200; apps/karman/karman_viewer.h:83-96) on the canonical,
        serialised tick schedule of the viewer (apps/karman/karman_viewer.cpp:100-184): mean velocity -> body force ->
        5 steps -> snapshot -> post-process, for 1000 steps; hashes at steps 0/5/100/500/1000, mean velocity at tick
        starts, rand() draws consumed by the body force, particle count.
  * pipe_default_fhp3 : the same for lgca-pipe at its defaults (FHP-III 1480 x 740).
  * hpp_4096 / fhp3_32768x512 : pure stepping at BASELINE config widths (C2: HPP 4096 x 4096 diffusion disc, periodic;
        C5 width: FHP-III 32768 x 512 periodic), 13 steps, hashes at 0/1/13.
Construction, initialisation and get_mean_velocity run on ONE thread (rand() inside OpenMP loops, racy counter --
SURVEY.md 8c determinism rules); stepping and post-processing use all cores (thread-count independent).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpu_checkers as cc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_runs.json")


def canonical_schedule(model, case, Re, Ma, cg, bc, steps, marks, cores):
    r = cc.Ref(model, case, Re, Ma, cg, threads=1)
    r.apply_bc(bc)
    r.init("random")
    out = {"model": model, "ctor": [case, Re, Ma, cg], "dims": [r.dim_x, r.dim_y], "bc": bc, "pp_interval": 5,
           "particles": r.n_particles(), "initial_forcing": r.initial_forcing(), "equilibrium_forcing": r.equilibrium_forcing(),
           "u": r.u, "hashes": {"0": r.hash()}, "mv_at_tick_start": {}, "forcing_at_tick": {}}
    out["chirality_hash"] = cc.fnv1a64(r.rnd)
    out["cell_type_hash"] = cc.fnv1a64(r.cell_type)
    r.set_threads(cores)
    r.snapshot()
    r.post_process()
    forcing = r.initial_forcing()
    u = r.u
    done = 0
    t0 = time.time()
    while done < steps:
        r.set_threads(1)
        mv = r.mean_velocity()
        r.set_threads(cores)
        tick_end = done + 5
        if tick_end in marks:
            out["mv_at_tick_start"][str(tick_end)] = [float(mv[0]), float(mv[1])]
        applied = 0
        if mv[0] < u:
            if mv[0] > 0.9 * u:
                forcing = r.equilibrium_forcing()
            r.body_force(forcing)
            applied = forcing
        if tick_end in marks:
            out["forcing_at_tick"][str(tick_end)] = applied
        r.step(5)
        done += 5
        r.snapshot()
        r.post_process()
        if done in marks:
            out["hashes"][str(done)] = r.hash()
            print("  %s step %d hash %s (%.0f s)" % (case, done, out["hashes"][str(done)], time.time() - t0), flush=True)
    assert r.n_particles() == out["particles"]
    r.close()
    return out


def pure_stepping(model, ctor, dims, bc, init, marks, cores):
    """Pure stepping.  `ctor` = (case, Re, Ma, cg) builds the shape through the reference's own constructor.  Shapes the
    constructor cannot produce (`dims`) are initialised by the pinned oracle (restated glibc rand(), seed 1) and copied
    into the reference's arrays, so that a test can rebuild the identical initial data without the reference."""
    if dims is None:
        r = cc.Ref(model, *ctor, threads=1)
        r.apply_bc(bc)
        r.init(init)
        how = "reference ctor + init"
    else:
        o = cc.Oracle(model, dims=dims, cg=ctor[3])
        o.apply_bc(bc)
        o.init(init)
        r = cc.Ref(model, "periodic", 63, 0.2, ctor[3], threads=1, dims=dims)
        r.cell_type[:] = o.cell_type
        r.rnd[:] = o.rnd
        r.state[:] = o.state
        how = "oracle init (dims) copied into the reference"
    out = {"model": model, "ctor": list(ctor) if dims is None else None, "dims": [r.dim_x, r.dim_y], "bc": bc, "init": init,
           "cg": ctor[3], "initial_data": how, "particles": r.n_particles(), "chirality_hash": cc.fnv1a64(r.rnd),
           "hashes": {"0": r.hash()}}
    r.set_threads(cores)
    done = 0
    for m in marks:
        r.step(m - done)
        done = m
        out["hashes"][str(m)] = r.hash()
    assert r.n_particles() == out["particles"]
    r.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if not cc.ref_available():
        raise SystemExit("oracle/_ref/liblgca_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    cores = os.cpu_count() or 1
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    data["_comment"] = ("Known answers generated from the UNMODIFIED reference (oracle/_ref) by scripts/gen_golden_reference.py; "
                        "see that script for the exact schedules.")
    want = set(filter(None, args.only.split(",")))
    jobs = {
        "karman_default": lambda: canonical_schedule("FHP_III", "karman", 80.0, 0.3, 20, "karman", 1000, {5, 100, 500, 1000}, cores),
        "pipe_default_fhp3": lambda: canonical_schedule("FHP_III", "pipe", 80.0, 0.3, 10, "pipe", 1000, {5, 100, 500, 1000}, cores),
        "hpp_4096": lambda: pure_stepping("HPP", ("periodic", 4095.0, 0.2, 16), None, "periodic", "diffusion", [1, 13], cores),
        "fhp3_32768x512": lambda: pure_stepping("FHP_III", ("periodic", 63.0, 0.2, 16), (32768, 512), "periodic", "random", [1, 13], cores),
    }
    for name, job in jobs.items():
        if want and name not in want:
            continue
        print("generating", name, flush=True)
        t0 = time.time()
        data[name] = job()
        data[name]["generated_in_s"] = round(time.time() - t0, 1)
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
