"""GPU parity of the thermal D3Q7 extension through the C ABI: LUW_TEMPERATURE domains against the oracle, bit for bit (STRICT).

All of these passed on the round-1 driver's B200 (19 XPASS in GPUTEST_r01.json); the xfail markers are gone. The kernel source is additionally checked
bit for bit against the oracle by compiling it for the host (tests/test_kernel_source_on_host.py).
"""
import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from tests import helpers as H

pytestmark = pytest.mark.gpu
NAMES = ("fi", "rho", "u", "gi", "T")
PRECS = pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])


@PRECS
@pytest.mark.parametrize("fset", ["luwT", "chanT"])
def test_strict_equals_oracle(oracle_lib, precision, fset):
    O = oracle_lib
    flags, rho, u, T = H.thermal_case()
    w = cases.relaxation_rate(1e-6)
    kw = dict(update_at_end=(fset == "chanT"))
    want = H.run_cpu_thermal(O.Oracle(), O, H.THERMAL_SHAPE, precision, O.FEATURE_SETS[fset], flags, rho, u, T, H.THERMAL_STEPS, w, **kw)
    got = H.run_cuda_thermal(H.THERMAL_SHAPE, precision, H.FEATURE_SETS[fset], flags, rho, u, T, H.THERMAL_STEPS, w, 0, **kw)
    for g, r, name in zip(got, want, NAMES):
        assert np.array_equal(g, r), name


# FAST arithmetic (mul+add contraction): tolerance on T after 24 steps. NOT yet calibrated on the device: 10x what contraction does to the same kernel source compiled
# for the host with -ffp-contract=fast -mfma (max|dT| 1.3e-6 / 5.1e-4 / 1.9e-4, rel-L2 1.8e-7 / 5.9e-5 / 3.0e-5 for FP32 / FP16S / FP16C); the flow keeps the
# tolerances of tests/test_gpu_parity.py
TOL_FAST_T = {0: dict(max_abs=2e-5, rel_l2=3e-6), 1: dict(max_abs=5e-3, rel_l2=6e-4), 2: dict(max_abs=2e-3, rel_l2=3e-4)}


@PRECS
def test_fast_within_tolerance(oracle_lib, precision):
    from tests.test_gpu_parity import TOL_FAST
    O = oracle_lib
    flags, rho, u, T = H.thermal_case()
    w = cases.relaxation_rate(1e-6)
    want = H.run_cpu_thermal(O.Oracle(), O, H.THERMAL_SHAPE, precision, O.FEATURE_SETS["luwT"], flags, rho, u, T, 24, w)
    got = H.run_cuda_thermal(H.THERMAL_SHAPE, precision, H.FEATURE_SETS["luwT"], flags, rho, u, T, 24, w, 1)
    tol = TOL_FAST_T[precision]
    assert float(np.abs(got[4] - want[4]).max()) <= tol["max_abs"] and H.rel_l2(got[4], want[4]) <= tol["rel_l2"]
    assert H.rel_l2(got[2], want[2]) <= 2 * TOL_FAST[precision]["rel_l2_u"]


NO_BUOYANCY = dict(H.THERMAL, beta=0.0)  # every LUW mode builds its LBM with f = 0: the buoyancy term vanishes and the thermal step runs as two kernels


@PRECS
@pytest.mark.parametrize("arith", [0, 1], ids=["strict", "fast"])
@pytest.mark.parametrize("shape", [(128, 12, 10), (253, 9, 7)], ids=["128x12x10", "253x9x7-oddNx"])
def test_two_kernel_thermal_step_equals_oracle(oracle_lib, precision, arith, shape):
    """LUW's shipped switches (FX/defines.hpp:14-24: UPDATE_FIELDS, VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, SUBGRID, TEMPERATURE, + nudging and sponge) on a lattice the
    TMA tile kernel takes: momentum in the tiled kernel, which hands the velocity before the force half-step to k_thermal_g (csrc/lbm_kernels.cuh). STRICT: every
    DDF, rho, u, g and T value equals the oracle's; FAST: within the tolerances of the fused kernel. With a live buoyancy term the same domain must fall back to the
    fused kernel and still equal the oracle."""
    from tests.test_gpu_parity import TOL_FAST
    O = oracle_lib
    flags, rho, u, T = H.thermal_case(shape, seed=5)
    w = cases.relaxation_rate(1e-6)
    for thermal, f in ((NO_BUOYANCY, H.FORCE), (H.THERMAL, (0.0, 0.0, 0.0)), (H.THERMAL, H.FORCE)):  # beta = 0 / f = 0: two kernels; both live: fused
        if arith == 1 and thermal is H.THERMAL and f == H.FORCE:
            continue
        want = H.run_cpu_thermal(O.Oracle(), O, shape, precision, O.FEATURE_SETS["luwT"], flags, rho, u, T, 8, w, f=f, thermal=thermal)
        got = H.run_cuda_thermal(shape, precision, H.FEATURE_SETS["luwT"], flags, rho, u, T, 8, w, arith, f=f, thermal=thermal, expect_tiles=True)
        if arith == 0:
            for g, r, name in zip(got, want, NAMES):
                assert np.array_equal(g, r), (name, thermal["beta"], f)
        else:
            tol = TOL_FAST_T[precision]
            assert float(np.abs(got[4] - want[4]).max()) <= tol["max_abs"] and H.rel_l2(got[4], want[4]) <= tol["rel_l2"]
            assert H.rel_l2(got[2], want[2]) <= 2 * TOL_FAST[precision]["rel_l2_u"]


def test_two_kernel_thermal_step_without_update_fields(oracle_lib):
    """FEAT = 14 | TEMPERATURE: rho / u / T are not stored per step (update_fields on demand), the pre-force velocity still reaches k_thermal_g."""
    O = oracle_lib
    shape = (128, 12, 10)
    flags, rho, u, T = H.thermal_case(shape, seed=6)
    w = cases.relaxation_rate(1e-6)
    feat = H.FEATURE_SETS["luwnf"] | 64
    want = H.run_cpu_thermal(O.Oracle(), O, shape, 1, feat, flags, rho, u, T, 7, w, thermal=NO_BUOYANCY, update_at_end=True)
    got = H.run_cuda_thermal(shape, 1, feat, flags, rho, u, T, 7, w, 0, thermal=NO_BUOYANCY, update_at_end=True, expect_tiles=True)
    for g, r, name in zip(got, want, NAMES):
        assert np.array_equal(g, r), name


def test_wide_lattice_and_batched_steps(oracle_lib):
    """Rows wider than one thread block, padded pitch (Nx = 150 -> 160), luw_run_steps."""
    O = oracle_lib
    shape = (150, 12, 10)
    flags, rho, u, T = H.thermal_case(shape, seed=9)
    w = cases.relaxation_rate(1e-6)
    want = H.run_cpu_thermal(O.Oracle(), O, shape, O.FP16S, O.FEATURE_SETS["luwT"], flags, rho, u, T, 5, w)
    got = H.run_cuda_thermal(shape, O.FP16S, H.FEATURE_SETS["luwT"], flags, rho, u, T, 5, w, 0, batched=True)
    for g, r, name in zip(got, want, NAMES):
        assert np.array_equal(g, r), name


def test_zero_gravity_leaves_the_flow_untouched():
    """LUW runs with f = 0: the thermal domain's fi / rho / u equal the plain domain's (which takes the TMA-tiled kernel here) bit for bit."""
    shape = (128, 16, 12)
    flags, rho, u, T = H.thermal_case(shape, seed=3)
    w = cases.relaxation_rate(1e-6)
    zero = (0.0, 0.0, 0.0)
    a = H.run_cuda_thermal(shape, 1, H.FEATURE_SETS["luwT"], flags, rho, u, T, 6, w, 0, f=zero)
    b = H.run_cuda(shape, 1, H.FEATURE_SETS["luw"], flags, rho, u, 6, w, 0, f=zero)
    for x, y, name in zip(a[:3], b, NAMES):
        assert np.array_equal(x, y), name


@pytest.mark.parametrize("D", [(2, 1, 1), (1, 2, 2), (2, 2, 2)], ids=["2x1x1", "1x2x2", "2x2x2"])
def test_decomposed_equals_single_domain(D):
    """Halo exchange of gi after fi every step, T + gi at initialisation (FX/lbm.cpp LBM::initialize / do_time_step): D domains == one domain."""
    from latticeurbanwind_b200 import _cabi as A
    from latticeurbanwind_b200.lbm import LBM
    shape = (64, 24, 16)
    flags, rho, u, T = H.thermal_case(shape, seed=21)
    zones = dict(downstream_face=2, buffer_N=3, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=3, sponge_inv_tau=0.02)
    res = []
    for dd in ((1, 1, 1), D):
        lbm = LBM(shape, D=dd, nu=1e-6, precision=A.FP16S, features=H.FEATURE_SETS["luwT"], arith=A.ARITH_STRICT, f=H.FORCE, omega=H.OMEGA,
                  alpha=2.0e-3, beta=0.4, **zones)
        lbm.flags[:], lbm.rho[:], lbm.u[:], lbm.T[:] = flags, rho, u, T
        lbm.run(7)
        lbm.read_from_device()
        res.append((lbm.rho.copy(), lbm.u.copy(), lbm.T.copy()))
        lbm.close()
    for x, y, name in zip(res[0], res[1], ("rho", "u", "T")):
        assert np.array_equal(x, y), name
    assert not np.array_equal(res[0][2], T)


def test_cell_set_moves_boundary_temperatures():
    """Boundary-field upload for T (the case driver's TYPE_T cells, FX/setup.cpp:5268-5317) without moving the whole field."""
    from latticeurbanwind_b200 import _cabi as A
    from latticeurbanwind_b200.domain import CellSet, Domain
    Nx, Ny, Nz = 37, 6, 5
    rng = np.random.default_rng(1)
    with Domain(Nx, Ny, Nz, precision=1, features=A.UPDATE_FIELDS | A.TEMPERATURE, w=1.0, arith=0) as d:
        cells = np.sort(rng.choice(d.N, 50, replace=False)).astype(np.uint64)
        vals = (1.0 + 0.1 * rng.standard_normal(50)).astype(np.float32)
        cs = CellSet(d, cells)
        cs.upload(A.FIELD_T, vals)
        d.read_from_device(A.FIELD_T)
        d.finish_queue()
        want = np.ones(d.N, np.float32)
        want[cells] = vals
        assert np.array_equal(d.T, want)
        back = np.zeros(50, np.float32)
        cs.download(A.FIELD_T, back)
        d.finish_queue()  # luw_sync also drains the copy stream the read-back runs on
        cs.close()
    assert np.array_equal(back, vals)


def test_plain_domains_reject_thermal_calls():
    from latticeurbanwind_b200 import _cabi as A
    from latticeurbanwind_b200.domain import Domain
    with Domain(16, 4, 4, precision=0, features=0, w=1.0, arith=0) as d:
        with pytest.raises(A.LuwError):
            d.set_thermal(1.0)
        with pytest.raises(A.LuwError):
            d.read_gi()


# ---------------------------------------------------------------------------------------------- the C++ host layer (reference LBM API) with lbm.T
def _run_cpp(tmp_path, shape, D, precision, steps, flags, rho, u, T, alpha, beta):
    import os
    import subprocess
    from tests import test_cpp_host as CPP
    CPP.build()
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as fh:
        fh.write(flags.tobytes()); fh.write(rho.tobytes()); fh.write(u.tobytes()); fh.write(T.tobytes())
    z = CPP.ZONES
    args = [CPP.DRIVER, *map(str, shape), *map(str, D), str(precision), str(H.FEATURE_SETS["luwT"]), "0", repr(1e-6), str(steps), str(z["downstream_face"]),
            str(z["buffer_N"]), repr(z["buffer_inv_tau"]), str(z["buffer_nudge_vertical"]), str(z["sponge_N"]), repr(z["sponge_inv_tau"]),
            *[repr(float(v)) for v in H.FORCE], *[repr(float(v)) for v in H.OMEGA], inp, out]
    r = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, LUW_CASE_ALPHA=repr(alpha), LUW_CASE_BETA=repr(beta)))
    assert r.returncode == 0, r.stderr
    N = int(np.prod(shape))
    raw = np.fromfile(out, np.float32)
    assert raw.size == 5 * N
    return raw[:N].copy(), raw[N:4 * N].copy(), raw[4 * N:].copy()


def test_cpp_lbm_with_temperature_equals_oracle(oracle_lib, tmp_path):
    """LBM(N, nu, f, sigma, alpha, beta) + lbm.T through the stitched accessor + run(steps): rho, u, T equal the oracle's (STRICT), and the
    2x2x2 decomposition (communicate_T / communicate_gi) equals the single domain."""
    from tests import test_cpp_host as CPP
    O = oracle_lib
    shape = (64, 20, 12)
    flags, rho, u, T = H.thermal_case(shape, seed=5)
    alpha, beta = 2.0e-3, 0.4
    thermal = dict(w_T=cases.kernel_literal(np.float32(1.0) / (np.float32(2.0) * np.float32(alpha) + np.float32(0.5))), beta=cases.kernel_literal(beta), T_avg=1.0)
    ref = H.run_cpu_thermal(O.Oracle(), O, shape, 1, O.FEATURE_SETS["luwT"], flags, rho, u, T, 6, cases.relaxation_rate(1e-6), zones=CPP.ZONES, thermal=thermal)
    one = _run_cpp(tmp_path, shape, (1, 1, 1), 1, 6, flags, rho, u, T, alpha, beta)
    for g, r, name in zip(one, (ref[1], ref[2], ref[4]), ("rho", "u", "T")):
        assert np.array_equal(g, r), name
    dec = _run_cpp(tmp_path, shape, (2, 2, 2), 1, 6, flags, rho, u, T, alpha, beta)
    for g, r, name in zip(dec, one, ("rho", "u", "T")):
        assert np.array_equal(g, r), name


def test_device_side_mean_temperature():
    """luw_stats_accumulate on a LUW_TEMPERATURE domain also keeps avg_T (FX/setup.cpp:4481-4486: t_avg += (T - t_avg) * inv_n, products and sums rounded separately)."""
    from latticeurbanwind_b200 import _cabi as A
    from latticeurbanwind_b200.domain import Domain, Stats
    Nx, Ny, Nz = 37, 6, 5
    rng = np.random.default_rng(4)
    with Domain(Nx, Ny, Nz, precision=1, features=A.UPDATE_FIELDS | A.TEMPERATURE, w=1.0, arith=0) as d:
        st = Stats(d)
        mean = np.zeros(d.N, np.float32)
        for k in range(1, 5):
            d.T[:] = (1.0 + 0.05 * rng.standard_normal(d.N)).astype(np.float32)
            d.write_to_device(A.FIELD_T)
            st.accumulate()
            inv_n = np.float32(1.0) / np.float32(k)
            mean = (mean + ((d.T - mean).astype(np.float32) * inv_n).astype(np.float32)).astype(np.float32)
        got = st.download_T()
        st.close()
    assert np.array_equal(got, mean)
