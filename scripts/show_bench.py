import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        print("N=%d value=%.4g ms/step=%.3f e2e=%.4g conserved=%s launch_ms=%.4f frac=%.3f | %s" % (
            d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("particles_conserved"),
            d["roofline"]["launch_ms"], d["roofline"]["frac"], d["config"].get("parallelism")))
        if "karman" in d: print("   karman: %.4g frac=%.3f" % (d["karman"]["value"], d["karman"]["roofline_frac"]))
        if "cpu_baseline" in d: print("   cpu:", d["cpu_baseline"])
        print("   clocks:", d.get("clocks"))
