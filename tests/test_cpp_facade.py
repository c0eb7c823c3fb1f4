"""The C++ facade (include/la3dm_b200/octomap.h: la3dm::BGKOctoMap etc. over the C ABI) -- what a maintainer of the
reference's nodes would compile against.  CPU: the header and a node-like program compile and link against the product
library.  GPU: the program's leaf walk agrees with the golden vectors of the reference."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

SRC = os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp")
LIBDIR = os.path.join(ROOT, "la3dm_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "facade_demo")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o",
                           exe, "-L" + LIBDIR, "-lla3dm_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_facade_compiles_and_links(tmp_path):
    exe = build(tmp_path)
    assert subprocess.run([exe]).returncode == 2          # usage error: no scan file, no GPU work attempted


@pytest.mark.gpu
def test_facade_matches_golden_sequence(tmp_path, scans):
    exe = build(tmp_path)
    pts, org = scans["sim_structured"]
    f = tmp_path / "scans.bin"
    with open(f, "wb") as fh:
        np.array([pts.shape[0], pts.shape[1]], np.int32).tofile(fh)
        org.astype(np.float32).tofile(fh)
        pts.astype(np.float32).tofile(fh)
    out = subprocess.run([exe, str(f)], check=True, capture_output=True, text=True).stdout.split()
    g = golden("golden_bgk_sim_structured_seq.npz")
    want = g["summaries"][-1]
    got = [float(x) for x in out]
    assert got[:4] == list(want[:4])                      # leaves, FREE, OCCUPIED, UNKNOWN
    assert abs(got[4] - want[6]) <= 1e-4 * want[6]        # sum of probabilities
    assert 0.0 < got[11] < 1.0
    assert got[12] >= got[13] > 10 and got[14] == 0     # RayCaster steps, valid steps, inconsistencies
