"""Summarise ncu artefacts from gpurun_out/ into profiles/ (tracked).  Usage:
   python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r01_launches_bench_c5.md
   python scripts/summarize_ncu.py full gpurun_out/wave_k4_c5.ncu-rep profiles/r01_wave_k4_c5.md [key]"""
import csv, io, json, os, subprocess, sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) -- source: %s\n\n" % os.path.basename(src))
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ns | share | avg ns |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.0f | %.1f%% | %.0f |\n" % (k, n, t, 100 * t / tot, t / n))
    print(open(dst).read())


def full(src, dst, key=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    r = rows[2]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary -- source: %s\n\nkernel: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % (
            os.path.basename(src), r[hdr.index("Kernel Name")]))
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = r[i]
                f.write("| %s | %s | %s |\n" % (k, r[i], units[i]))
    print(open(dst).read())
    if key:
        def tobytes(k):
            i = hdr.index(k); v = float(r[i].replace(",", "")); u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        # stamped with the kernel sources the capture was taken from: bench.py attaches a number only to the same build
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
        from bench import kernel_source_stamp
        tp = os.path.join(os.path.dirname(dst), "traffic.json")
        t = json.load(open(tp)) if os.path.exists(tp) else {}
        if "entries" not in t:
            t = {"entries": {}}
        t["entries"][key] = {"dram_bytes_per_launch": tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'),
                             "kernel_stamp": kernel_source_stamp(), "source": os.path.basename(dst)}
        json.dump(t, open(tp, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
