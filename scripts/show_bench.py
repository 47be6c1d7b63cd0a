"""Pretty-print bench.py JSON lines (stdin or files): python scripts/show_bench.py gpurun_out/x.json"""
import json
import sys


def show(d):
    print("N=%d value=%.4g ms/step=%.3f e2e=%.4g conserved=%s launch_ms=%.4f frac=%.3f | %s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("particles_conserved"),
        d["roofline"]["launch_ms"], d["roofline"]["frac"], d["config"].get("parallelism")))
    print("   traffic:", d["roofline"].get("traffic"), "|", d["roofline"].get("traffic_source"))
    for k in ("karman", "diffusion_hpp"):
        if k in d:
            print("   %s: %.4g frac=%.3f" % (k, d[k]["value"], d[k]["roofline_frac"]))
    if "box" in d:
        b = d["box"]
        print("   box (C4, N=%d): %.4g total, %.4g per GPU, %.2f us/update" % (b["n_gpus"], b["value"], b["per_gpu"], b["us_per_update"]))
    if "multi_gpu_parity" in d:
        print("   multi_gpu_parity:", d["multi_gpu_parity"]["ok"], [v["digest"][:8] for v in d["multi_gpu_parity"]["variants"]])
    for k, v in d.get("resident", {}).items():
        print("   resident %-32s %.3f us/update (wave %.3f) x%.2f; tick of 5: %.3f vs %.3f" % (
            k, v["resident"]["us_per_update_call_of_1000"], v["wave"]["us_per_update_call_of_1000"], v["speedup"],
            v["resident"]["us_per_update_call_of_5"], v["wave"]["us_per_update_call_of_5"]))
    for k, v in d.get("app_tick", {}).items():
        b, r = v.get("b200", {}), v.get("reference", {})
        print("   app_tick %-16s b200 %.2f ms/tick (mv %.2f bf %.2f step %.2f pp %.2f s total)  reference %.1f ms/tick  x%.1f" % (
            k, b.get("ms_per_tick", float("nan")), b.get("mean_velocity_s", 0), b.get("body_force_s", 0), b.get("stepping_s", 0),
            b.get("post_process_s", 0), r.get("ms_per_tick", float("nan")), v.get("tick_speedup", float("nan"))))
    if "cpu_baseline" in d:
        print("   cpu:", d["cpu_baseline"])
    print("   clocks:", d.get("clocks"))


for src in ([open(a) for a in sys.argv[1:]] or [sys.stdin]):
    for line in src:
        line = line.strip()
        if line.startswith("{"):
            show(json.loads(line))
