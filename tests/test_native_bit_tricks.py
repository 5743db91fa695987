"""The word-level tricks of the reachability closure (k_lights.cu: fill_up / fill_down, what a carry adds, the
frontier tests) restated on the host and checked against one another on millions of random and structured words."""
import os
import subprocess


def test_fill_and_frontier_identities(tmp_path):
    src = os.path.join(os.path.dirname(__file__), "native", "fill_check.cpp")
    exe = str(tmp_path / "fill_check")
    subprocess.run(["g++", "-O2", "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr
