#!/usr/bin/env python3
"""Generates tests/golden/ref_inlet.npz from the REFERENCE's own interpolator classes (NearestNeighborInterpolator / KNNInterpolatorHD behind InletVelocityField(HD),
compiled from FX/interpolation.cpp / interpolation_hd.cpp where they lie into baseline/_ref/luw_inlet_parity_on_host by baseline/build_reference_driver.py):
the velocities they return at the open-face positions of a 23 x 19 x 13 lattice for four sample clouds (regular grid with ties and coincident samples, jittered,
sparse faces, collinear samples). Run in the build container only:  python tests/golden/make_golden_inlet.py"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "baseline", "_ref", "luw_inlet_parity_on_host")

with tempfile.TemporaryDirectory() as tmp:
    raw_path = os.path.join(tmp, "inlet.bin")
    subprocess.check_call([EXE], env=dict(os.environ, LUW_INLET_GOLDEN=raw_path), stdout=subprocess.DEVNULL)
    raw = open(raw_path, "rb").read()
off = 0


def take(dtype, count):
    global off
    a = np.frombuffer(raw, dtype, count, off).copy()
    off += a.nbytes
    return a


clouds, npos = (int(v) for v in take(np.uint32, 2))
out = {"pos": take(np.float32, 3 * npos).reshape(npos, 3)}
for k in range(clouds):
    n = int(take(np.uint32, 1)[0])
    out[f"z_threshold_{k}"] = take(np.float32, 1)
    out[f"P_{k}"] = take(np.float32, 3 * n).reshape(n, 3)
    out[f"U_{k}"] = take(np.float32, 3 * n).reshape(n, 3)
    out[f"hd_{k}"] = take(np.float32, 3 * npos).reshape(npos, 3)
    out[f"nearest_{k}"] = take(np.float32, 3 * npos).reshape(npos, 3)
assert off == len(raw)
np.savez_compressed(os.path.join(HERE, "ref_inlet.npz"), **out)
print({k: v.shape for k, v in out.items()})
