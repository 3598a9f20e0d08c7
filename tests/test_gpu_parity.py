"""GPU parity: CUDA path through the C ABI vs the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): flags and cell indexing bit-exact; u / rho within a stated tolerance after N steps.
 * LUW_ARITH_STRICT: the kernels evaluate the reference's expressions as written -> every DDF, rho and u value must be EQUAL to the
   oracle's (the oracle itself is pinned bit-for-bit to the reference kernel text, tests/test_oracle_vs_reference.py).
 * LUW_ARITH_FAST: contraction allowed -> tolerance stated below.
"""
import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from tests import helpers as H

pytestmark = pytest.mark.gpu

STEPS = 24
# FAST-mode tolerances after STEPS steps on the small urban case (lattice units; c_s = 0.577, |u| <= 0.17)
TOL_FAST = {0: dict(rel_l2_u=2e-5, max_abs_u=2e-6, rel_l2_rho=1e-6), 1: dict(rel_l2_u=2e-3, max_abs_u=4e-4, rel_l2_rho=1e-4), 2: dict(rel_l2_u=1e-3, max_abs_u=2e-4, rel_l2_rho=1e-4)}


@pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
@pytest.mark.parametrize("fset", ["bench", "plain", "core", "luw", "luwnf"])
def test_strict_equals_oracle(oracle_lib, precision, fset):
    O = oracle_lib
    shape = (48, 40, 32)
    flags, rho, u = H.small_urban(*shape)
    w = cases.relaxation_rate(1e-6)
    feat = H.FEATURE_SETS[fset]
    ref = H.run_cpu(O.Oracle(), O, shape, precision, feat, flags, rho, u, STEPS, w)
    got = H.run_cuda(shape, precision, feat, flags, rho, u, STEPS, w, arith=0)
    assert np.array_equal(H.decode(O, None, got[0], precision), H.decode(O, None, ref[0], precision)), "DDFs differ"
    assert np.array_equal(got[1], ref[1]), "rho differs"
    assert np.array_equal(got[2], ref[2]), "u differs"


@pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
def test_fast_within_tolerance(oracle_lib, precision):
    O = oracle_lib
    shape = (48, 40, 32)
    flags, rho, u = H.small_urban(*shape)
    w = cases.relaxation_rate(1e-6)
    feat = H.FEATURE_SETS["luw"]
    ref = H.run_cpu(O.Oracle(), O, shape, precision, feat, flags, rho, u, STEPS, w)
    got = H.run_cuda(shape, precision, feat, flags, rho, u, STEPS, w, arith=1)
    tol = TOL_FAST[precision]
    assert H.rel_l2(got[2], ref[2]) <= tol["rel_l2_u"]
    assert float(np.abs(got[2] - ref[2]).max()) <= tol["max_abs_u"]
    assert H.rel_l2(got[1], ref[1]) <= tol["rel_l2_rho"]


def test_update_fields_on_demand(oracle_lib):
    """Without UPDATE_FIELDS the host sees rho/u only after update_fields (FX/lbm.hpp:406-412)."""
    O = oracle_lib
    shape = (32, 24, 16)
    flags, rho, u = H.small_urban(*shape)
    w = cases.relaxation_rate(1e-4)
    feat = H.FEATURE_SETS["luwnf"]
    for precision in (0, 1):
        ref = H.run_cpu(O.Oracle(), O, shape, precision, feat, flags, rho, u, 7, w, update_at_end=True)
        got = H.run_cuda(shape, precision, feat, flags, rho, u, 7, w, arith=0, update_at_end=True, batched=True)
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
