#!/usr/bin/env python3
"""Generates tests/golden/ref_thermal.json from the REFERENCE's own kernel text built with -DTEMPERATURE (oracle/_ref/libluwref_<prec>_luwT.so and
_chanT.so, see oracle/Makefile). Run in the build container only:

    python tests/golden/make_golden_thermal.py

sha256 of the raw fi / rho / u / gi / T images after THERMAL_STEPS steps of the reference's initialize + stream_collide (+ update_fields for the set
without UPDATE_FIELDS) on tests/helpers.thermal_case, and of every gi / T halo payload of a 2x2x2 block. tests/test_thermal_oracle.py checks the C
oracle against them on any machine.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from latticeurbanwind_b200 import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    O.build()
    out = {"_about": "reference kernel text with TEMPERATURE; see make_golden_thermal.py", "shape": list(H.THERMAL_SHAPE), "steps": H.THERMAL_STEPS,
           "cases": {}}
    flags, rho, u, T = H.thermal_case()
    w = cases.relaxation_rate(1e-6)
    for precision in (O.FP32, O.FP16S, O.FP16C):
        for fset in ("luwT", "chanT"):
            r = H.run_cpu_thermal(O.Reference(precision, fset), O, H.THERMAL_SHAPE, precision, O.FEATURE_SETS[fset], flags, rho, u, T, H.THERMAL_STEPS, w,
                                  update_at_end=(fset == "chanT"))
            out["cases"][f"{O.PREC_NAME[precision]}_{fset}"] = dict(zip(("fi", "rho", "u", "gi", "T"), map(sha, r)))
        for key, val in H.golden_thermal_halo(O.Reference(precision, "luwT"), O, precision).items():
            out["cases"][key] = sha(val)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_thermal.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"ref_thermal.json: {len(out['cases'])} entries")


if __name__ == "__main__":
    main()
