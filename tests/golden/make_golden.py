#!/usr/bin/env python3
"""Generates tests/golden/*.json|npz from the REFERENCE's own kernel text (oracle/_ref/libluwref_*.so, built by oracle/Makefile from
/root/reference/core/cfd_core/FluidX3D/src/kernel.cpp where it lies). Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors for the LBM step, so these fixtures ARE the pin: they are outputs of the
reference's stream_collide / initialize / update_fields / transfer_* / vk_inlet_apply text executed on host threads, on the seeded
inputs of tests/helpers.py. tests/test_oracle_golden.py checks the C oracle against them on any machine (no /root/reference needed).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    O.build()
    out = {"_about": "sha256 of the raw fi / rho / u images after GOLDEN_STEPS steps of the reference kernel text; see make_golden.py",
           "shape": list(H.GOLDEN_SHAPE), "steps": H.GOLDEN_STEPS, "cases": {}}
    tiny = {}
    for precision in (O.FP32, O.FP16S, O.FP16C):
        for fset in O.FEATURE_SETS:
            if O.FEATURE_SETS[fset] & O.TEMPERATURE:
                continue  # thermal sets: make_golden_thermal.py
            ref = O.Reference(precision, fset)
            fi, rho, u = H.golden_run(ref, O, precision, fset)
            out["cases"][f"{O.PREC_NAME[precision]}_{fset}"] = {"fi": sha(fi), "rho": sha(rho), "u": sha(u)}
            if fset in ("luw", "bench"):
                fi, rho, u = H.golden_run(ref, O, precision, fset, shape=H.TINY_SHAPE, steps=H.TINY_STEPS)
                tiny[f"{O.PREC_NAME[precision]}_{fset}_fi"] = fi
                tiny[f"{O.PREC_NAME[precision]}_{fset}_rho"] = rho
                tiny[f"{O.PREC_NAME[precision]}_{fset}_u"] = u
    # halo payloads of a 2x2x2 decomposition block (domain with halos on all axes), both parities
    for precision in (O.FP32, O.FP16S):
        ref = O.Reference(precision, "luw")
        for key, val in H.golden_halo(ref, O, precision).items():
            out["cases"][key] = sha(val)
    # codecs: every FP16C code -> float, and a float sweep -> FP16C (FX/kernel.cpp:864-875)
    ref = O.Reference(O.FP16C, "luw")
    dec = np.array([ref.fp16c_to_float(h) for h in range(65536)], np.float32)
    sweep = H.codec_sweep()
    enc = np.array([ref.float_to_fp16c(x) for x in sweep], np.uint16)
    out["cases"]["fp16c_decode_all"] = sha(dec)
    out["cases"]["fp16c_encode_sweep"] = sha(enc)
    # von Karman inlet (FX/kernel.cpp:2495-2571): libm cosf on this image; stored as values (tolerance-compared on the GPU)
    pc, pf, pd, md, M, V, N = H.vk_case()
    for interp in (0, 1):
        u = np.zeros(3 * N, np.float32)
        ref.bind(O.make_params(*H.VK_SHAPE, O.FP16C, O.FEATURE_SETS["luw"]))
        ref.vk_inlet_apply(interp, 3.0, 4.0, 0.25, pc, pf, pd, md, M, V, u)
        tiny[f"vk_u_interp{interp}"] = u
    with open(os.path.join(HERE, "ref_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "ref_tiny.npz"), **tiny)
    print("wrote", len(out["cases"]), "hash entries and", len(tiny), "arrays")


if __name__ == "__main__":
    main()
