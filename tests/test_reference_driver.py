"""GPU: the reference's own, unmodified case driver (FX/setup.cpp + main.cpp + info.cpp + interpolation*.cpp + fluxcorrection.cpp, built by
baseline/build_reference_driver.py against this repo's host layer and CUDA library) runs the reference's own example project
(examples/example_ProfileResearch_noDEM, BASELINE configs[0]: deck-driven profile inflow, STL voxelisation, von Karman inlet, nudging, sponge, FP16C DDFs)
on the B200 and writes its VTK results. This is the drop-in claim end to end; numerical parity of every kernel it launches is covered by the other GPU tests."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "baseline", "_ref", "luw_reference_driver")
CASE = os.path.join(ROOT, "baseline", "_ref", "case_profile")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.isfile(DRIVER) and os.path.isdir(CASE)), reason="baseline/_ref was not built (needs /root/reference at build time)")]


def read_vtk(path):
    raw = open(path, "rb").read()
    head, _, data = raw.partition(b"LOOKUP_TABLE default\n")
    dims = tuple(int(v) for v in re.search(rb"DIMENSIONS (\d+) (\d+) (\d+)", head).groups())
    comps = int(re.search(rb"SCALARS data float (\d+)", head).group(1))
    a = np.frombuffer(data, dtype=">f4")
    assert a.size == dims[0] * dims[1] * dims[2] * comps
    return dims, a.reshape(dims[2], dims[1], dims[0], comps)


def test_reference_case_driver_runs_its_example_deck(tmp_path):
    case = str(tmp_path / "case")
    shutil.copytree(CASE, case)
    r = subprocess.run([DRIVER, os.path.join(case, "conf.luwpf")], capture_output=True, text=True, timeout=600, cwd=case)
    log = r.stdout[-6000:] + r.stderr[-2000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    open(os.path.join(out_dir, "reference_driver.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, log
    assert "Grid Resolution" in r.stdout and "253" in r.stdout
    vtks = [os.path.join(b, f) for b, _, fs in os.walk(case) for f in fs if f.endswith(".vtk")]
    assert vtks, log
    fields = [p for p in vtks if re.search(r"(^|[_/-])u[-_.]", os.path.basename(p)) or "_u" in os.path.basename(p)]
    checked = 0
    for p in vtks:
        dims, a = read_vtk(p)
        assert dims[0] == 253 and dims[1] == 250
        assert np.isfinite(a).all(), p
        if a.shape[3] == 3:  # a velocity field in SI units: the inflow profile tops out at 7.8 m/s
            speed = np.sqrt((a.astype(np.float64) ** 2).sum(axis=3))
            assert 1.0 < float(speed.max()) < 40.0, (p, float(speed.max()))
            checked += 1
    assert checked >= 1, [os.path.basename(p) for p in vtks]
