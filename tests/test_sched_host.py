"""Host-side property test of the backward sweep's schedule arithmetic (flyp_b200/csrc/sched.h, shared verbatim by the
pair kernel, the partial-sum reduction kernel and the host): compiled with g++ and run on the CPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sweep_schedule_covers_every_unit_once(tmp_path):
    exe = str(tmp_path / "sched_check")
    src = os.path.join(ROOT, "tests", "native", "sched_check.cpp")
    inc = os.path.join(ROOT, "flyp_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, src, "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.startswith("OK")
