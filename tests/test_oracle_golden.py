"""CPU: the C oracle (oracle/luw_oracle.c) against fixtures produced by the REFERENCE's own kernel text (tests/golden/make_golden.py).

The reference has no golden vectors for the LBM step; the fixtures in tests/golden/ are outputs of its stream_collide / initialize /
transfer_* / vk_inlet_apply text compiled for host threads (oracle/ref_shim). Bar: bit-exact (hash of the raw arrays).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "ref_hashes.json")) as _f:
    HASHES = json.load(_f)
TINY = np.load(os.path.join(GOLD, "ref_tiny.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_fixture_metadata():
    assert tuple(HASHES["shape"]) == H.GOLDEN_SHAPE and HASHES["steps"] == H.GOLDEN_STEPS


@pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
@pytest.mark.parametrize("fset", ["bench", "plain", "core", "luw", "luwnf"])
def test_step_hashes(oracle_lib, precision, fset):
    O = oracle_lib
    fi, rho, u = H.golden_run(O.Oracle(), O, precision, fset)
    want = HASHES["cases"][f"{O.PREC_NAME[precision]}_{fset}"]
    assert sha(fi) == want["fi"] and sha(rho) == want["rho"] and sha(u) == want["u"]


@pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])
@pytest.mark.parametrize("fset", ["bench", "luw"])
def test_tiny_arrays(oracle_lib, precision, fset):
    O = oracle_lib
    fi, rho, u = H.golden_run(O.Oracle(), O, precision, fset, shape=H.TINY_SHAPE, steps=H.TINY_STEPS)
    key = f"{O.PREC_NAME[precision]}_{fset}"
    assert np.array_equal(fi, TINY[key + "_fi"]) and np.array_equal(rho, TINY[key + "_rho"]) and np.array_equal(u, TINY[key + "_u"])


@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "fp16s"])
def test_halo_payloads(oracle_lib, precision):
    O = oracle_lib
    for key, val in H.golden_halo(O.Oracle(), O, precision).items():
        assert sha(val) == HASHES["cases"][key], key


def test_fp16c_codec(oracle_lib):
    O = oracle_lib
    orc = O.Oracle()
    dec = np.array([orc.fp16c_to_float(h) for h in range(65536)], np.float32)
    enc = np.array([orc.float_to_fp16c(x) for x in H.codec_sweep()], np.uint16)
    assert sha(dec) == HASHES["cases"]["fp16c_decode_all"]
    assert sha(enc) == HASHES["cases"]["fp16c_encode_sweep"]
    # round trip: every code survives decode -> encode (except -0 variants that the encoder canonicalises identically)
    back = np.array([orc.float_to_fp16c(x) for x in dec], np.uint16)
    assert np.array_equal(back, np.arange(65536, dtype=np.uint16))


def test_fp16s_codec_against_numpy(oracle_lib):
    """FP16S is IEEE binary16 of 2^15*f (FX/lbm.cpp:709-710): numpy's float16 is an independent implementation of the same rounding."""
    O = oracle_lib
    orc = O.Oracle()
    sweep = H.codec_sweep()
    sweep = sweep[np.abs(sweep) < 60000.0]
    enc = np.array([orc.float_to_half(x) for x in sweep], np.uint16)
    assert np.array_equal(enc, sweep.astype(np.float16).view(np.uint16))
    dec = np.array([orc.half_to_float(h) for h in range(65536)], np.float32)
    want = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    ok = np.isfinite(want)
    assert np.array_equal(dec[ok], want[ok])


def test_vk_inlet(oracle_lib):
    O = oracle_lib
    pc, pf, pd, md, M, V, N = H.vk_case()
    orc = O.Oracle().bind(O.make_params(*H.VK_SHAPE, O.FP16C, O.FEATURE_SETS["luw"]))
    for interp in (0, 1):
        u = np.zeros(3 * N, np.float32)
        orc.vk_inlet_apply(interp, 3.0, 4.0, 0.25, pc, pf, pd, md, M, V, u)
        assert np.array_equal(u, TINY[f"vk_u_interp{interp}"])
