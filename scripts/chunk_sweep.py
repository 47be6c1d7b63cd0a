"""Chunk-height sweep of the fused-step kernel (development helper): LGCA_B200_CHUNK_ROWS is read when the plan is made --
only by a tuning build: scripts/build_variant.sh ab_tuning.so -DLGCA_B200_TUNING; LGCA_B200_LIB=$PWD/ab_tuning.so python scripts/chunk_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200

CASES = [("FHP_III", 32768, 32768, "periodic"), ("FHP_III", 16384, 8192, "karman"), ("FHP_II", 65536, 32768, "reflecting_back")]
CHUNKS = (0, 1024, 700, 490, 410, 328, 246, 200, 164, 124, 100, 82)
KS = (6, 5)
TAILS = [""]
if len(sys.argv) > 4:  # tail plans "pct,rows" separated by ';', e.g. "0,0;20,32;30,24"
    TAILS = sys.argv[4].split(";")
if len(sys.argv) > 1:  # e.g. chunk_sweep.py karman 6 0,32,48,64,96
    CASES = [c for c in CASES if c[3] == sys.argv[1]]
    KS = tuple(int(x) for x in sys.argv[2].split(","))
    CHUNKS = tuple(int(x) for x in sys.argv[3].split(","))
for model, dx, dy, bc in CASES:
    for k in KS:
        for cr in CHUNKS:
            if cr > dy:
                continue
            for tail in TAILS:
                if cr:
                    os.environ["LGCA_B200_CHUNK_ROWS"] = str(cr)
                else:
                    os.environ.pop("LGCA_B200_CHUNK_ROWS", None)
                os.environ["LGCA_B200_TAIL"] = tail or "0,0"
                e = lgca_b200.Engine(model, dx, dy, k_fuse=k, flags=lgca_b200.capi.FLAG_NO_CELL_FIELDS)
                e.apply_bc_device(bc)
                e.init_random_device(1)
                launches = max(4, int(2e-2 / (dx * dy * k / 8e12)))
                e.timed_kernel(max(2, launches // 4))
                ms = min(e.timed_kernel(launches) for _ in range(3))
                print("%s %dx%d %s k=%d chunk_rows=%s tail=%s: %.2f us/update  %.3e sites/s" % (
                    model, dx, dy, bc, k, cr or "planner", tail or "-", ms * 1e3 / k, dx * dy * k / (ms * 1e-3)), flush=True)
                e.close()
