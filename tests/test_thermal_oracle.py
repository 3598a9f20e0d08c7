"""CPU: thermal D3Q7 transport (FX/kernel.cpp:1306-1336, 1442-1450, 1639-1684, 1981-2000, 2337-2377) in the C oracle.

Pinned two ways, bit for bit: against fixtures written by the reference's own kernel text built with -DTEMPERATURE (tests/golden/ref_thermal.json,
tests/golden/make_golden_thermal.py; runs on any machine) and, where oracle/_ref exists (the build container), against that text directly on a
larger case with other zone / thermal constants."""
import hashlib
import json
import os

import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from oracle import oracle as O
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "ref_thermal.json")) as _f:
    HASHES = json.load(_f)
NAMES = ("fi", "rho", "u", "gi", "T")
PRECS = pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
needs_ref = pytest.mark.skipif(not O.ref_available(O.FP32, "luwT"), reason="oracle/_ref (TEMPERATURE build) not present on this machine")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_fixture_metadata():
    assert tuple(HASHES["shape"]) == H.THERMAL_SHAPE and HASHES["steps"] == H.THERMAL_STEPS


@PRECS
@pytest.mark.parametrize("fset", ["luwT", "chanT"])
def test_step_hashes(oracle_lib, precision, fset):
    flags, rho, u, T = H.thermal_case()
    r = H.run_cpu_thermal(O.Oracle(), O, H.THERMAL_SHAPE, precision, O.FEATURE_SETS[fset], flags, rho, u, T, H.THERMAL_STEPS, cases.relaxation_rate(1e-6),
                          update_at_end=(fset == "chanT"))
    want = HASHES["cases"][f"{O.PREC_NAME[precision]}_{fset}"]
    for name, arr in zip(NAMES, r):
        assert sha(arr) == want[name], name
    assert not np.array_equal(r[4], T)  # the temperature field did move


@PRECS
def test_halo_payload_hashes(oracle_lib, precision):
    for key, val in H.golden_thermal_halo(O.Oracle(), O, precision).items():
        assert sha(val) == HASHES["cases"][key], key


def test_g_eq_moments(oracle_lib):
    """D3Q7 equilibrium in DDF-shifted form: sum g_eq = T - 1, first moments = T u / 4 * ... (c_s^2 = 1/4): sum c_i g_eq_i = T u."""
    orc = O.Oracle()
    rng = np.random.default_rng(3)
    for _ in range(100):
        T, ux, uy, uz = 1.0 + 0.1 * rng.normal(), *(0.1 * rng.normal(size=3))
        g = orc.g_eq(T, ux, uy, uz).astype(np.float64)
        assert abs(g.sum() - (np.float32(T) - 1.0)) < 1e-6
        assert abs((g[1] - g[2]) - np.float32(T) * np.float32(ux)) < 1e-6
        assert abs((g[3] - g[4]) - np.float32(T) * np.float32(uy)) < 1e-6
        assert abs((g[5] - g[6]) - np.float32(T) * np.float32(uz)) < 1e-6


def test_momentum_is_untouched_without_buoyancy(oracle_lib):
    """Every LUW mode builds the LBM with zero gravity (FX/setup.cpp:4935,5720,6018): f = 0 makes the buoyancy term vanish, and fi / rho / u of the
    thermal entry points equal the plain ones bit for bit -- the property that lets the flow parity stand when TEMPERATURE is added."""
    flags, rho, u, T = H.thermal_case()
    w = cases.relaxation_rate(1e-6)
    feat = O.FEATURE_SETS["luw"]
    a = H.run_cpu_thermal(O.Oracle(), O, H.THERMAL_SHAPE, O.FP16S, feat | O.TEMPERATURE, flags, rho, u, T, 6, w, f=(0.0, 0.0, 0.0))
    b = H.run_cpu(O.Oracle(), O, H.THERMAL_SHAPE, O.FP16S, feat, flags, rho, u, 6, w, f=(0.0, 0.0, 0.0))
    for x, y, name in zip(a[:3], b, NAMES):
        assert np.array_equal(x, y), name


def test_uniform_temperature_is_a_fixed_point(oracle_lib):
    """T == const, no TYPE_T sources: g_eq is velocity-weighted but the total stays T and the field must not drift beyond rounding."""
    shape = (12, 10, 8)
    N = int(np.prod(shape))
    flags, rho = np.zeros(N, np.uint8), np.ones(N, np.float32)
    u = np.concatenate([np.full(N, 0.05, np.float32), np.zeros(2 * N, np.float32)])
    T = np.full(N, 1.25, np.float32)
    r = H.run_cpu_thermal(O.Oracle(), O, shape, O.FP32, O.UPDATE_FIELDS | O.TEMPERATURE, flags, rho, u, T, 10, 1.0, f=(0.0, 0.0, 0.0), omega=(0.0, 0.0, 0.0))
    assert np.max(np.abs(r[4] - 1.25)) < 1e-5


@needs_ref
@PRECS
@pytest.mark.parametrize("fset", ["luwT", "chanT"])
def test_against_reference_text(oracle_lib, precision, fset):
    shape = (30, 26, 18)
    flags, rho, u, T = H.thermal_case(shape, seed=5)
    zones = dict(downstream_face=3, buffer_N=5, buffer_inv_tau=0.02, buffer_nudge_vertical=0, sponge_N=6, sponge_inv_tau=0.05)
    thermal = dict(w_T=1.0 / (2.0 * 0.05 + 0.5), beta=1.5, T_avg=1.01)
    w = cases.relaxation_rate(1e-5)
    kw = dict(zones=zones, thermal=thermal, update_at_end=(fset == "chanT"))
    a = H.run_cpu_thermal(O.Oracle(), O, shape, precision, O.FEATURE_SETS[fset], flags, rho, u, T, 12, w, **kw)
    b = H.run_cpu_thermal(O.Reference(precision, fset), O, shape, precision, O.FEATURE_SETS[fset], flags, rho, u, T, 12, w, **kw)
    for x, y, name in zip(a, b, NAMES):
        assert np.array_equal(x, y), name


@needs_ref
@PRECS
def test_halo_payloads_against_reference_text(oracle_lib, precision):
    a = H.golden_thermal_halo(O.Oracle(), O, precision)
    b = H.golden_thermal_halo(O.Reference(precision, "luwT"), O, precision)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
