"""-m "not gpu": host-side arithmetic the sort-free front-end rests on (tests/cpp/host_logic.cpp): add_repeat against
sequential additions, the reconstruction of the sensor voxel's sum from its few non-origin samples, and the bounding box
of a beam's samples from its extreme samples -- all bit for bit, on the CPU, with the library's own common.cuh."""
import os
import subprocess

from conftest import ROOT


def test_fused_frontend_arithmetic_claims(tmp_path):
    exe = str(tmp_path / "host_logic")
    src = os.path.join(ROOT, "tests", "cpp", "host_logic.cpp")
    inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I" + inc, src, "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.strip().endswith("0 bad"), r.stdout[-500:]
