"""CPU, build container only (needs oracle/_ref built from /root/reference): the C oracle against the reference's own kernel text,
bit for bit, on larger and more varied cases than the committed fixtures. Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (no /root/reference on this machine)")


@pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
@pytest.mark.parametrize("fset", ["bench", "plain", "core", "luw", "luwnf"])
def test_urban_steps_bit_exact(oracle_lib, precision, fset):
    shape = (40, 36, 24)
    flags, rho, u = H.small_urban(*shape)
    w = cases.relaxation_rate(1e-6)
    a = H.run_cpu(O.Oracle(), O, shape, precision, O.FEATURE_SETS[fset], flags, rho, u, 16, w, update_at_end=(fset == "luwnf"))
    b = H.run_cpu(O.Reference(precision, fset), O, shape, precision, O.FEATURE_SETS[fset], flags, rho, u, 16, w, update_at_end=(fset == "luwnf"))
    for x, y, name in zip(a, b, ("fi", "rho", "u")):
        assert np.array_equal(x, y), name


@pytest.mark.parametrize("downstream", [0, 1, 2, 3, 4])
def test_relaxation_zone_variants(oracle_lib, downstream):
    shape = (30, 28, 20)
    flags, rho, u = H.small_urban(*shape)
    zones = dict(downstream_face=downstream, buffer_N=5, buffer_inv_tau=0.02, buffer_nudge_vertical=downstream % 2, sponge_N=1 + downstream, sponge_inv_tau=0.03)
    w = cases.relaxation_rate(1e-5)
    a = H.run_cpu(O.Oracle(), O, shape, O.FP32, O.FEATURE_SETS["luw"], flags, rho, u, 8, w, zones=zones)
    b = H.run_cpu(O.Reference(O.FP32, "luw"), O, shape, O.FP32, O.FEATURE_SETS["luw"], flags, rho, u, 8, w, zones=zones)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
def test_halo_payloads_bit_exact(oracle_lib, precision):
    a = H.golden_halo(O.Oracle(), O, precision)
    b = H.golden_halo(O.Reference(precision, "luw"), O, precision)
    for key in a:
        assert np.array_equal(a[key], b[key]), key


def test_feq_and_codecs(oracle_lib):
    orc, ref = O.Oracle(), O.Reference(O.FP16C, "luw")
    rng = np.random.default_rng(5)
    for _ in range(200):
        rho, ux, uy, uz = 1.0 + 0.05 * rng.normal(), *(0.1 * rng.normal(size=3))
        assert np.array_equal(orc.f_eq(rho, ux, uy, uz), ref.f_eq(rho, ux, uy, uz))
    for h in range(0, 65536, 7):
        assert orc.fp16c_to_float(h) == ref.fp16c_to_float(h) or (np.isnan(orc.fp16c_to_float(h)) and np.isnan(ref.fp16c_to_float(h)))
    for x in H.codec_sweep()[::5]:
        assert orc.float_to_fp16c(x) == ref.float_to_fp16c(x)
