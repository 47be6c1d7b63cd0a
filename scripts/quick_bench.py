"""Quick device-side timing of the stepping kernels (development helper, not the bench contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200

def run(model, dx, dy, bc, k_fuse=0, flags=0, steps=240, reps=3):
    e = lgca_b200.Engine(model, dx, dy, k_fuse=k_fuse, flags=flags)
    e.apply_bc_device(bc)
    e.init_random_device(1)
    n0 = e.count_particles()
    e.timed_steps(steps)
    best = min(e.timed_steps(steps) for _ in range(reps))
    i = e.info()
    ups = dx * dy * steps / (best * 1e-3)
    gbs = ups * i.bytes_per_site_step_x8 / 8 / 1e9
    ok = e.count_particles() == n0
    print(f"{model} {dx}x{dy} {bc} k={i.k_fuse} flags={flags}: {best/steps*1e3:.1f} us/step  {ups/1e9:.1f} G sites/s  "
          f"alg {gbs:.0f} GB/s  conserved={ok}", flush=True)
    e.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for flags, k in ((2, 1), (0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (0, 8)):
        if which != "all" and which != ("simple" if flags else f"k{k}"):
            continue
        run("FHP_III", 16384, 8192, "karman", k, flags)
        run("HPP", 4096, 4096, "periodic", k, flags)
        run("FHP_III", 32768, 32768, "periodic", k, flags, steps=120)
        run("FHP_II", 16384, 8192, "reflecting_back", k, flags)
        run("FHP_I", 1400, 700, "pipe", k, flags, steps=240)
