"""The LOP3 network search tool (scripts/lop3_search.c) that produced the 5-op change-mask block of the FHP collision
network: it must still find that block, and still prove that the per-pair output block cannot be done in 3 ops.
CPU only, a few seconds."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool(tmp_path):
    exe = str(tmp_path / "lop3_search")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "scripts", "lop3_search.c")])
    return exe


def test_change_mask_block_has_a_five_op_network(tmp_path):
    exe = _tool(tmp_path)
    r = subprocess.run([exe, "T", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "feasible" in r.stdout and "NOT feasible" not in r.stdout, r.stdout
    r = subprocess.run([exe, "T", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "NOT feasible" in r.stdout, r.stdout


def test_pair_output_block_needs_four_ops(tmp_path):
    exe = _tool(tmp_path)
    r = subprocess.run([exe, "pair", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "NOT feasible" in r.stdout, r.stdout
    r = subprocess.run([exe, "pair", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
