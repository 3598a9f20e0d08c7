"""GPU: wall time of the two inflow-sample searches (luw_inlet_nearest / luw_inlet_knn, host buffers in and out) at the size of a C3 lattice's open faces:
1024 x 1024 top + 4 x 1024 x 255 sides = 2.09 M face cells; SurfData clouds of 20 000 samples per face. Prints one JSON line."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from latticeurbanwind_b200 import _cabi as A

L = A.lib()
rng = np.random.default_rng(3)
out = {}
for ncells, npts in ((1 << 20, 20000), (1 << 18, 20000)):
    cell = rng.uniform(-512, 512, (2, ncells)).astype(np.float32)
    pts = rng.uniform(-512, 512, (npts, 2)).astype(np.float32)
    kept = np.zeros((ncells, 64), np.uint32); used = np.zeros(ncells, np.uint32); mr = np.zeros(ncells, np.float32); ex = np.zeros(ncells, np.int32)
    for rep in range(2):
        t = time.perf_counter()
        A.check(L.luw_inlet_knn(0, ncells, cell.ctypes.data, npts, pts.ctypes.data, kept.ctypes.data, used.ctypes.data, mr.ctypes.data, ex.ctypes.data))
        dt = time.perf_counter() - t
    # spot check against numpy on a few cells: the kept set is the 64 smallest r2
    for c in range(0, ncells, ncells // 7):
        s1 = pts[:, 0] - cell[0, c]; s2 = pts[:, 1] - cell[1, c]; r2 = s1 * s1 + s2 * s2
        want = np.sort(r2)[:64]; got = np.sort(r2[kept[c]])
        assert used[c] == 64 and np.array_equal(want, got) and mr[c] == want[-1], (c, want[-3:], got[-3:], mr[c])
    out[f"knn_{ncells}_cells_x_{npts}_samples_s"] = round(dt, 4)
ncells, npts = 1 << 21, 100000
cell = rng.uniform(-512, 512, (3, ncells)).astype(np.float32)
pts = rng.uniform(-512, 512, (npts, 3)).astype(np.float32)
near = np.zeros(ncells, np.uint32)
for rep in range(2):
    t = time.perf_counter()
    A.check(L.luw_inlet_nearest(0, ncells, cell.ctypes.data, npts, pts.ctypes.data, near.ctypes.data))
    dt = time.perf_counter() - t
for c in range(0, ncells, ncells // 5):
    d = cell[:, c][None, :] - pts; d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    assert near[c] == int(np.argmin(d2)), (c, near[c], int(np.argmin(d2)))
out[f"nearest_{ncells}_cells_x_{npts}_samples_s"] = round(dt, 4)
print(json.dumps(out))
