"""GPU: the two device voxelisers on a city-sized mesh -- 1024 x 1024 x 128 lattice, 20 000 boxes (240 012 triangles) over a base slab, z-rays -- timed through
luw_voxelize_mesh (host triangles in, synchronous: bin grid build + uploads + kernel). Flags of the two runs must be equal. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from latticeurbanwind_b200 import _cabi as A
from latticeurbanwind_b200.domain import Domain
from tests import helpers as H
from tests.test_kernel_source_on_host import _city_mesh

out = {}
for shape, boxes in (((1024, 1024, 128), 20000), ((253, 250, 59), 760)):
    p0, p1, p2, pmin, pmax = _city_mesh(shape, boxes, 9)
    ntri = p0.size // 3
    bbu = H.vox_bbu(ntri, pmin, pmax)
    res = {}
    for bins in ("1", "0"):
        os.environ["LUW_VOXELIZE_BINS"] = bins
        with Domain(*shape, precision=A.FP16S, features=H.FEATURE_SETS["luw"], w=1.0, arith=A.ARITH_FAST, **H.ZONES) as d:
            ts = []
            for rep in range(2):
                d.finish_queue()
                t = time.perf_counter()
                d.voxelize_mesh(2, 1, p0, p1, p2, bbu)
                ts.append(time.perf_counter() - t)
            d.read_from_device(A.FIELD_FLAGS); d.finish_queue()
            res[bins] = (min(ts), np.array(d.flags).copy())
    assert np.array_equal(res["1"][1], res["0"][1]), "binned and all-triangles flags differ"
    out["x".join(map(str, shape))] = {"triangles": ntri, "solid_cells": int(((res["1"][1] & 3) == 1).sum()), "binned_s": round(res["1"][0], 5), "all_triangles_s": round(res["0"][0], 5)}
print(json.dumps(out))
