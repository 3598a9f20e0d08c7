"""CPU: SURVEY 8-f2 without a GPU. latticeurbanwind_b200/host/inlet_outlet_surface.cpp (sample search through luw_inlet_nearest / luw_inlet_knn, quadratic fit on the
host) against the reference's own InletVelocityField / InletVelocityFieldHD functors, compiled from FX/interpolation.cpp / interpolation_hd.cpp where they lie: every
velocity bit-identical over five sample clouds (ties, coincident samples, sparse planes, singular fits, no samples). The two search entry points are the repo's KERNEL
SOURCE (csrc/lbm_inlet.cuh) compiled for the host (tests/host_emulation/inlet_on_host.cpp, linked into the test executable in place of libluw_cuda.so's): this checks
the logic of the kernels and of the host code around them; the same harness runs on the device in tests/test_reference_driver.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "baseline", "_ref", "luw_inlet_parity_on_host")
SOURCES = [os.path.join(ROOT, p) for p in ("latticeurbanwind_b200/host/inlet_outlet_surface.cpp", "latticeurbanwind_b200/csrc/lbm_inlet.cuh", "baseline/inlet_parity.cpp",
                                            "tests/host_emulation/inlet_on_host.cpp", "tests/host_emulation/cuda_on_host.hpp")]


def _fresh():
    return os.path.isfile(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(s) for s in SOURCES)


def test_inlet_mapping_equals_the_reference_functors_on_host():
    if not _fresh():
        if not os.path.isdir("/root/reference"):
            pytest.skip("baseline/_ref/luw_inlet_parity_on_host is missing or stale and the reference tree is not here to rebuild it")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "baseline", "build_reference_driver.py")], env=dict(os.environ, LUW_DROPIN_SKIP_T="1"), stdout=subprocess.DEVNULL)
    for batch in ("", "777"):  # default batching (one batch here) and 777 cells per call of the K-nearest search
        r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL, env=dict(os.environ, LUW_INLET_BATCH=batch))
        assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-1000:]
        assert "20 of 20 runs identical" in r.stdout, r.stdout[-4000:]
        assert r.stdout.count("IDENTICAL") == 20
        launches = int(r.stdout.strip().rsplit("search kernels launched:", 1)[1])
        assert (launches > 50) if batch else (launches == 50), launches
