"""Diagnosis: does the app's body force slow down while a parent process holds a CUDA context (as in bench.py's app_tick extra)?"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgca_b200
from lgca_b200.capi import FLAG_NO_CELL_FIELDS, FLAG_NO_CHAIN

def run(tag):
    p = subprocess.run([os.path.join(ROOT, "scripts/dbg/lgca-pipe-trace"), "--model", "FHP_I", "--steps", "1000", "--quiet"], capture_output=True, text=True)
    tick = [l for l in p.stdout.splitlines() if l.startswith("Tick")]
    calls = [l for l in p.stderr.splitlines() if l.startswith("bf:")]
    dev = sum(float(l.split()[-2]) for l in calls)
    draw = sum(float(l.split()[-5]) for l in calls)
    slow = sorted(calls, key=lambda l: -float(l.split()[-2]))[:3]
    print(tag, tick[0][:150] if tick else p.stdout[-200:], "| device ms", round(dev, 2), "draw ms", round(draw, 2), flush=True)
    for l in slow:
        print("    ", l, flush=True)

run("no parent context")
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
run("parent: torch context only")
for flags, name in ((FLAG_NO_CHAIN, "serial launches"), (0, "chained launches")):
    e = lgca_b200.Engine("FHP_III", 32768, 8192, flags=FLAG_NO_CELL_FIELDS | flags)
    e.init_random_device(1)
    e.step(60); e.sync()
    run("parent: engine alive, %s, idle" % name)
    e.close()
    run("parent: engine closed after %s" % name)
