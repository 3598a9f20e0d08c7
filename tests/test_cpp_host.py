"""The C++ host layer (latticeurbanwind_b200/host: the reference's LBM / LBM_Domain / Memory<T> API over the C ABI), driven through its
command-line case driver luw_host_case."""
import os
import subprocess

import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "latticeurbanwind_b200", "lib")
DRIVER = os.path.join(LIB, "luw_host_case")
ZONES = dict(downstream_face=2, buffer_N=3, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=3, sponge_inv_tau=0.02)


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "latticeurbanwind_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "latticeurbanwind_b200", "host")], stdout=subprocess.DEVNULL)


def run_driver(tmp_path, shape, D, precision, features, arith, nu, steps, flags, rho, u, f=H.FORCE, omega=H.OMEGA, zones=ZONES, check=True, stats=0):
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as fh:
        fh.write(flags.tobytes()); fh.write(rho.tobytes()); fh.write(u.tobytes())
    args = [DRIVER, *map(str, shape), *map(str, D), str(precision), str(features), str(arith), repr(float(nu)), str(steps), str(zones["downstream_face"]),
            str(zones["buffer_N"]), repr(zones["buffer_inv_tau"]), str(zones["buffer_nudge_vertical"]), str(zones["sponge_N"]), repr(zones["sponge_inv_tau"]),
            *[repr(float(v)) for v in f], *[repr(float(v)) for v in omega], inp, out]
    env = dict(os.environ, LUW_CASE_STATS=str(stats)) if stats else None
    r = subprocess.run(args, capture_output=True, text=True, env=env)
    if check:
        assert r.returncode == 0, r.stderr
        N = int(np.prod(shape))
        raw = np.fromfile(out, np.float32)
        if stats:  # rho, u, avg_u (interleaved), avg_rho, M2_u, M2_v, M2_w
            assert raw.size == 11 * N
            return raw[:N].copy(), raw[N:4 * N].copy(), raw[4 * N:7 * N].copy(), raw[7 * N:8 * N].copy(), raw[8 * N:9 * N].copy(), raw[9 * N:10 * N].copy(), raw[10 * N:].copy()
        return raw[:N].copy(), raw[N:].copy()
    return r


def test_host_layer_builds_and_exports_the_reference_api():
    build()
    syms = subprocess.check_output(["nm", "-DC", os.path.join(LIB, "libluw_host.so")], text=True)
    for name in ("LBM::run(", "LBM::reset()", "LBM::update_fields()", "LBM::LBM(uint3", "LBM_Domain::LBM_Domain(", "lbm_settings", "lbm_kernel_literal"):
        assert name in syms, name


def test_kernel_literal_matches_the_python_twin():
    """def_w as the kernel sees it: C++ lbm_kernel_literal == cases.kernel_literal (both restate FX/utilities.hpp:2741-2750)."""
    import ctypes as C
    build()
    L = C.CDLL(os.path.join(LIB, "libluw_host.so"))
    fn = getattr(L, "_Z18lbm_kernel_literalf")
    fn.restype, fn.argtypes = C.c_float, [C.c_float]
    rng = np.random.default_rng(3)
    vals = np.concatenate([np.exp(rng.uniform(-20, 20, 400)), [1.0, 1.9999992, 0.5, 1e-7, 3.3333334e-3, 12345.678]]).astype(np.float32)
    for v in vals:
        assert np.float32(fn(float(v))) == cases.kernel_literal(v), v


def test_no_device_is_a_boxed_error_and_exit_1(tmp_path):
    """Reference error convention (FX/utilities.hpp:4370-4382): print a boxed message, exit(1). Without a GPU the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    build()
    shape = (16, 8, 8)
    flags, rho, u = cases.periodic_box(*shape)
    r = run_driver(tmp_path, shape, (1, 1, 1), 0, 0, 0, 1 / 6, 1, flags, rho, u, check=False)
    assert r.returncode == 1 and "Error" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "fp16s"])
def test_cpp_lbm_equals_oracle(oracle_lib, tmp_path, precision):
    """LBM(N, nu) + flags/rho/u through the accessors + run(steps) + read_from_device == the oracle, bit for bit (STRICT arithmetic)."""
    O = oracle_lib
    shape = (128, 20, 12)
    flags, rho, u = cases.urban(*shape, seed=5, edge=4, pitch=8)
    nu = 1e-6
    feat = H.FEATURE_SETS["luw"]
    ref = H.run_cpu(O.Oracle(), O, shape, precision, feat, flags, rho, u, 6, cases.relaxation_rate(nu), zones=ZONES)
    got_rho, got_u = run_driver(tmp_path, shape, (1, 1, 1), precision, feat, 0, nu, 6, flags, rho, u)
    assert np.array_equal(got_rho, ref[1]) and np.array_equal(got_u, ref[2])


@pytest.mark.gpu
@pytest.mark.parametrize("D,arith", [((2, 1, 1), 0), ((1, 2, 2), 0), ((2, 2, 2), 0), ((1, 2, 2), 1), ((2, 1, 1), 1)], ids=["2x1x1-strict", "1x2x2-strict", "2x2x2-strict", "1x2x2-fast", "2x1x1-fast"])
def test_cpp_decomposed_equals_single_domain(tmp_path, D, arith):
    """n_gpu = [Dx,Dy,Dz] through the C++ layer (domains share the GPUs that exist): identical fields to the single-domain run.
    STRICT arithmetic is bit-identical across all step kernels; FAST is bit-identical between the two-pass tile kernels of any tile shape (x-split blocks have
    Nx/Dx + 2 cells per row and run the tile kernel through the padded device row pitch)."""
    shape = (128, 24, 16)
    flags, rho, u = cases.urban(*shape, seed=21, edge=4, pitch=8)
    feat = H.FEATURE_SETS["luw"]
    one = run_driver(tmp_path, shape, (1, 1, 1), 1, feat, arith, 1e-6, 7, flags, rho, u)
    dec = run_driver(tmp_path, shape, D, 1, feat, arith, 1e-6, 7, flags, rho, u)
    assert np.array_equal(one[0], dec[0]) and np.array_equal(one[1], dec[1])


@pytest.mark.gpu
def test_cpp_statistics_match_the_oracle_loop(oracle_lib, tmp_path):
    """LBM_Statistics (device-side Welford over the last 4 of 7 steps, all domains, stitched into the reference's avg_u / avg_rho / M2 layout):
    (1) the 2x2x2 decomposition gives the same statistics as the single domain, bit for bit; (2) they equal the reference's host loop
    (oracle restatement of FX/setup.cpp:4441-4488) applied to the per-step fields of the same STRICT run."""
    O = oracle_lib
    shape = (128, 24, 16)
    N = int(np.prod(shape))
    flags, rho, u = cases.urban(*shape, seed=21, edge=4, pitch=8)
    feat = H.FEATURE_SETS["luw"]
    one = run_driver(tmp_path, shape, (1, 1, 1), 1, feat, 0, 1e-6, 7, flags, rho, u, stats=4)
    dec = run_driver(tmp_path, shape, (2, 2, 2), 1, feat, 0, 1e-6, 7, flags, rho, u, stats=4)
    for a, b in zip(one, dec):
        assert np.array_equal(a, b)
    orc = O.OracleStats()
    u_avg, rho_avg = np.zeros(3 * N, np.float32), np.zeros(N, np.float32)
    m2 = [np.zeros(N, np.float32) for _ in range(3)]
    for steps in (4, 5, 6, 7):  # the fields the host sees after each of the last four steps
        ref = H.run_cpu(O.Oracle(), O, shape, 1, feat, flags, rho, u, steps, cases.relaxation_rate(1e-6), zones=ZONES)
        orc.accumulate(ref[1], ref[2], u_avg, rho_avg, *m2)
    assert np.array_equal(one[2], u_avg) and np.array_equal(one[3], rho_avg)
    assert np.array_equal(one[4], m2[0]) and np.array_equal(one[5], m2[1]) and np.array_equal(one[6], m2[2])


BENCH = os.path.join(LIB, "luw_host_bench")


def _host_bench(*args, env=None):
    import json
    build()
    r = subprocess.run([BENCH, *map(str, args)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_cpp_api_end_to_end_loop():
    """bench.py's e2e.cpp_host leg in miniature: LBM built through the global accessors, one run(1) per step, per-step boundary upload and probe read-back
    through Memory<T>'s ranged transfers (host buffers)."""
    res = _host_bench("urban", 256, 128, 64, 1, 62, 20, 5)
    assert res["tiled"] == 1 and res["mlups"] > 100.0 and res["h2d_bytes_per_step"] == 12 * 256 * 128 and res["d2h_bytes_per_step"] == 16 * 256 * 128
    assert 0.01 < res["probe_mean_ux"] < 0.2 and res["host_mirror_u"] == 1 and res["host_mirror_flags"] == 1


@pytest.mark.gpu
def test_one_gigacell_domain_without_host_mirrors():
    """1024^3 = 1.07 G cells FP16S (59 GB of HBM): the reference would hold 17 B per cell on the host (FX/lbm.cpp:95-106, 18 GB here, 171 GB for the 10 G-cell
    configuration). Memory<T>'s mirrors are lazy: a case that does not index whole fields on the host never allocates them; a ranged read of a > 4 GB field lands in
    pageable memory whose untouched pages stay uncommitted."""
    import torch
    if torch.cuda.mem_get_info(0)[0] < 70 * 2 ** 30:
        pytest.skip("needs 70 GB of free device memory")
    res = _host_bench("rest", 1024, 1024, 1024, 1, 0, 3, 1)
    assert res["cells"] == 1024 ** 3 and res["tiled"] == 1 and res["mlups"] > 1000.0
    assert res["host_mirror_rho"] == 0 and res["host_mirror_flags"] == 0, res
    assert res["probe_mean_ux"] == 0.0  # a fluid at rest stays at rest
    assert res["peak_rss_mb"] < 8000, res  # no 18 GB host image: what is resident is the CUDA context (3 - 4 GB on these boxes); the u mirror exists after the ranged read, 16 KB of it committed
    assert res["device_mb"] > 50000
