"""lgca_wave_pins.h (resident-blocks hints of the 160 wavefront-kernel variants) is generated: the committed header must be what
scripts/gen_wave_pins.py produces from the committed ptxas logs, and the final build's log must show the outcome the header
comments promise (no variant below its pre-chain occupancy class unless it is listed as unpinned; spills bounded)."""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOGS = [os.path.join(ROOT, "profiles", n) for n in ("r03_wave_ptxas_before_chain.log", "r03_wave_ptxas_unpinned.log",
                                                    "r03_wave_ptxas_all_pinned.log")]


def _table(path):
    txt, d = open(path).read(), {}
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", txt):
        mm = re.search(r"step_wave_kernelILi(\d)ELi(\d)ELb(\d)ELb(\d)ELb(\d)", m.group(1))
        if mm:
            d[tuple(int(x) for x in mm.groups())] = (int(m.group(4)), int(m.group(3)))
    return d


def test_header_is_generated_from_the_committed_logs(tmp_path):
    work = tmp_path / "repo"
    (work / "lgca_b200" / "csrc").mkdir(parents=True)
    (work / "scripts").mkdir()
    shutil.copy(os.path.join(ROOT, "scripts", "gen_wave_pins.py"), work / "scripts")
    subprocess.check_call([sys.executable, "scripts/gen_wave_pins.py"] + LOGS, cwd=work, stdout=subprocess.DEVNULL)
    want = open(work / "lgca_b200" / "csrc" / "lgca_wave_pins.h").read()
    assert want == open(os.path.join(ROOT, "lgca_b200", "csrc", "lgca_wave_pins.h")).read()


def test_final_build_keeps_the_occupancy_classes():
    before = _table(LOGS[0])
    final = _table(os.path.join(ROOT, "profiles", "r03_wave_ptxas_final.log"))
    assert len(before) == len(final) == 160
    hdr = open(os.path.join(ROOT, "lgca_b200", "csrc", "lgca_wave_pins.h")).read()
    unpinned = {tuple(int(x) for x in m.groups()) for m in
                re.finditer(r"return  0; // model (\d) K (\d) ns (\d) sl (\d) irreg (\d)", hdr)}
    assert len(unpinned) <= 8
    for k, (regs, spill) in before.items():
        if k in unpinned:
            continue
        assert 512 // final[k][0] >= 512 // regs, (k, regs, final[k])      # warps per scheduler not lower than before
        assert final[k][1] <= spill + 32, (k, spill, final[k])              # at most eight more spilled words
    # the kernels of the headline configs carry no spill at all
    for k in ((2, 6, 0, 0, 0), (2, 6, 1, 0, 0), (0, 6, 0, 0, 0)):
        assert final[k][1] == 0
