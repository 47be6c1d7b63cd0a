"""Instruction mix of the largest loop of a kernel (steady-state body).  usage: sass_loop.py <so> <substring of mangled name>"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, txt in ins:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    print(name[:120])
    print("total instructions:", len(ins), " largest loop: 0x%x..0x%x" % best)
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    cnt = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        cnt[op.split(".")[0] if not op.startswith("IMAD") else op] += 1
    print("loop body instructions:", len(body))
    alu = sum(v for k, v in cnt.items() if k in ("LOP3", "IADD3", "SHF", "ISETP", "SEL", "PRMT", "LEA", "VIADD", "IADD", "PLOP3", "VIADDMNMX", "IABS"))
    print("alu-pipe (approx):", alu)
    for k, v in cnt.most_common(30):
        print("  %-16s %d" % (k, v))
