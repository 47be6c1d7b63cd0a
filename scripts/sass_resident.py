"""Counts the TMA-family / mbarrier / 128-bit shared instructions of the SM-resident kernel in the built library.
   python scripts/sass_resident.py > profiles/r02_resident_sass.md"""
import collections
import re
import subprocess

sass = subprocess.run(["cuobjdump", "-sass", "lgca_b200/liblgca_b200.so"], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
fn, k, counts = None, -1, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        fn = names[k].replace("lgca_b200::", "") if "resident" in m.group(1) else None
        if fn:
            counts[fn] = collections.Counter()
        continue
    if fn:
        m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?((?:UBLKCP|SYNCS|LDS\.128|STS\.128|UTMALDG|UBLKPF|LDG|STG|BAR)[\w.]*)", line)
        if m:
            counts[fn][m.group(1)] += 1
print("# SASS evidence: SM-resident kernel (csrc/lgca_step_resident.cu), `cuobjdump -sass lgca_b200/liblgca_b200.so` (sm_100a)\n")
print("`UBLKCP.S.G` = TMA-family bulk copy global -> shared (staging the lattice once per call), `UBLKCP.G.S` = shared -> global")
print("(write-back after the last step), `SYNCS.*TRANS64` = mbarrier transaction waits of those copies, `LDS.128/STS.128` = the")
print("four-words-per-thread row accesses.  Template arguments: <MODEL, HAS_NO_SLIP, HAS_SLIP, OWN> (OWN = words a thread owns statically, 0 = dynamic four-word groups).\n")
for fn, c in counts.items():
    print("* `%s`: " % fn + ", ".join("%s x%d" % kv for kv in sorted(c.items())))
