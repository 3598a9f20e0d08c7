#!/usr/bin/env python3
"""Generates tests/golden/ref_stats.json from the REFERENCE's own averaging loop (oracle/_ref/libluwref_stats.so, built by oracle/Makefile from the text of
FX/setup.cpp:4441-4488 where it lies). Run in the build container only:  python tests/golden/make_golden_stats.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

out = {"_about": "sha256 of u_avg (interleaved), rho_avg, M2_u, M2_v, M2_w after tests.helpers.stats_samples() through the reference text",
       "arrays": [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in H.stats_run(O.RefStats())]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stats.json"), "w"), indent=1)
print(out)
