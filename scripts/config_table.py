"""Device-resident throughput of the five BASELINE.json configurations on one GPU (development helper; the bench
contract lives in bench.py).  Prints a markdown table."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200

PEAK = 6547.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

CONFIGS = [
    ("C1 lgca-pipe FHP-I 1400x700 (app default size)", "FHP_I", 1400, 700, "pipe"),
    ("C2 lgca-diffusion HPP 4096x4096 periodic", "HPP", 4096, 4096, "periodic"),
    ("C3 lgca-karman FHP-III 16384x8192, walls + cylinder", "FHP_III", 16384, 8192, "karman"),
    ("C4 lgca-box FHP-II 65536x32768 (whole lattice on ONE GPU)", "FHP_II", 65536, 32768, "reflecting_back"),
    ("C5 lgca-periodic FHP-III 32768x32768", "FHP_III", 32768, 32768, "periodic"),
]
print("| config | k_fuse | us / update | site updates/s | alg. B/site/step | alg. GB/s | x measured HBM peak (%.0f GB/s) |" % PEAK)
print("|---|---:|---:|---:|---:|---:|---:|")
for name, model, dx, dy, bc in CONFIGS:
    e = lgca_b200.Engine(model, dx, dy, flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
    e.apply_bc_device(bc)
    e.init_random_device(1)
    i = e.info()
    k = i.k_fuse
    launches = max(4, int(2e-2 / (dx * dy * k / 8e12)))
    e.timed_kernel(max(2, launches // 4))
    ms = min(e.timed_kernel(launches) for _ in range(3))
    ups = dx * dy * k / (ms * 1e-3)
    bps = i.bytes_per_site_step_x8 / 8.0
    print("| %s | %d | %.2f | %.3g | %.3f | %.0f | %.2f |" % (name, k, ms * 1e3 / k, ups, bps, ups * bps / 1e9, ups * bps / 1e9 / PEAK), flush=True)
    e.close()
