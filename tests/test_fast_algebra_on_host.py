"""CPU: the FAST two-pass collision of csrc/lbm_vec.cuh (moments_of / fast_prepare / fast_relax_*: an algebraic regrouping of FX/kernel.cpp:1016-1113, 1686-1748)
against the as-written STRICT formulation of the same header (collide_strict2, pinned bit for bit to the oracle by the GPU tests), both compiled for the host
(tests/host_emulation/vec_on_host.cpp) and evaluated on the same seeded DDFs. A wrong coefficient or sign in the regrouping is an O(1e-3..1) relative error in the
add terms; rounding differences between the two formulations are O(1e-7). The approximate reciprocal / square root of the device build are replaced by exact ones
here, so this checks the algebra, not the device's approximations (those are covered by the FAST tolerances of tests/test_gpu_parity.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
Q = 19
CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0], np.float64)
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1], np.float64)
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1], np.float64)
W = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, np.float64)


@pytest.fixture(scope="module")
def vec_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host_emulation"), "libluw_vec_on_host.so"])
    L = C.CDLL(os.path.join(HERE, "host_emulation", "libluw_vec_on_host.so"))
    L.emu_fast_vs_strict.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_float, C.c_void_p, C.c_float] + [C.c_void_p] * 4
    L.emu_fast_vs_strict.restype = C.c_int
    L.emu_fast_equilibrium.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
    L.emu_fast_equilibrium.restype = C.c_int
    return L


def seeded_ddfs(npairs, seed):
    """Shifted DDFs (f_i - w_i) of turbulent-looking cells: equilibrium of rho ~ 1 +- 2e-2, |u| <= 0.15 plus a non-equilibrium part of a few per cent."""
    rng = np.random.default_rng(seed)
    n = 2 * npairs
    rho = 1.0 + 2e-2 * rng.standard_normal(n)
    u = 0.15 * (2.0 * rng.random((n, 3)) - 1.0)
    cu = 3.0 * (u[:, :1] * CX + u[:, 1:2] * CY + u[:, 2:3] * CZ)
    usq = (u ** 2).sum(1, keepdims=True)
    feq = W * rho[:, None] * (1.0 + cu + 0.5 * cu * cu - 1.5 * usq) - W
    f = feq + W * 3e-2 * rng.standard_normal((n, Q))
    return f.astype(np.float32)


@pytest.mark.parametrize("scale", [1.0, 32768.0], ids=["S1", "S2^15"])
@pytest.mark.parametrize("feat", [0, 4, 5, 6, 12, 14, 15])
def test_fast_two_pass_algebra_matches_the_as_written_collision(vec_lib, feat, scale):
    npairs = 4000
    f = seeded_ddfs(npairs, 1000 + feat)
    out = [np.zeros(npairs * 2 * Q, np.float32) for _ in range(2)] + [np.zeros(npairs * 8, np.float32) for _ in range(2)]
    fo = np.array([1e-5, -2e-5, 3e-5, 0.0, 5.6e-6, 4.7e-6], np.float32)
    for w in (1.9999992, 1.0, 0.6):  # LUW's nu ~ 1e-7 (relaxation at the stability limit, Smagorinsky does the work), nu = 1/6, a viscous case
        rc = vec_lib.emu_fast_vs_strict(feat, npairs, f.ctypes.data, C.c_float(w), fo.ctypes.data, C.c_float(scale), *[o.ctypes.data for o in out])
        assert rc == 0
        strict, fast, ru_s, ru_f = out
        assert np.isfinite(fast).all()
        # post-collision DDFs: |f| <= 0.05; the regrouping only changes the rounding of O(0.1 .. 1) intermediates (a few float ulps: ~1e-7 absolute),
        # a wrong coefficient or sign would be >= 1e-4
        err = float(np.abs(fast.astype(np.float64) - strict).max())
        ref = float(np.abs(strict).max())
        assert 0.01 < ref < 1.0 and err <= 3e-7, (feat, w, err, ref)
        assert float(np.abs(ru_f.astype(np.float64) - ru_s).max()) <= 4e-7, (feat, w)


@pytest.mark.parametrize("scale", [1.0, 32768.0], ids=["S1", "S2^15"])
@pytest.mark.parametrize("feat", [4, 14, 15])
def test_type_e_lanes_relax_to_the_equilibrium_of_the_boundary_fields(vec_lib, feat, scale):
    """FX/kernel.cpp:1503-1522, 1747: a TYPE_E cell takes rho / u from the boundary fields, applies the force half-step (Coriolis) and the clamp, and sets f := feq.
    The FAST two-pass path does that as the relaxation with rate 1 and without forcing term (fast_prepare, `any_e`): whatever was streamed in must be wiped."""
    npairs = 3000
    rng = np.random.default_rng(77 + feat)
    f = seeded_ddfs(npairs, 5000 + feat) * np.float32(3.0)  # what the TYPE_E cells' slots hold does not matter
    bnd = np.zeros((2 * npairs, 4), np.float32)
    bnd[:, 0] = 1.0 + 1e-2 * rng.standard_normal(2 * npairs)
    bnd[:, 1:] = 0.12 * (2.0 * rng.random((2 * npairs, 3)) - 1.0)
    fo = np.array([1e-5, -2e-5, 3e-5, 0.0, 5.6e-6, 4.7e-6], np.float32)
    out = np.zeros(npairs * 2 * Q, np.float32)
    rc = vec_lib.emu_fast_equilibrium(feat, npairs, f.ctypes.data, bnd.ctypes.data, C.c_float(1.9999992), fo.ctypes.data, C.c_float(scale), out.ctypes.data)
    assert rc == 0 and np.isfinite(out).all()
    rho, u = bnd[:, 0].astype(np.float64), bnd[:, 1:].astype(np.float64)
    if feat & 2:  # VOLUME_FORCE: F = f - 2 rho Omega x u, u += F / (2 rho), clamp to +-c
        om = fo[3:].astype(np.float64)
        F = fo[:3].astype(np.float64) - 2.0 * rho[:, None] * np.cross(np.broadcast_to(om, u.shape), u)
        u = u + F / (2.0 * rho[:, None])
    u = np.clip(u, -0.57735027, 0.57735027)
    cu = 3.0 * (u[:, :1] * CX + u[:, 1:2] * CY + u[:, 2:3] * CZ)
    feq = W * rho[:, None] * (1.0 + cu + 0.5 * cu * cu - 1.5 * (u ** 2).sum(1, keepdims=True)) - W
    err = float(np.abs(out.reshape(-1, Q).astype(np.float64) - feq).max())
    assert err <= 3e-7, (feat, scale, err)
