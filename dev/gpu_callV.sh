#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_V.log
python - <<'PY'
# voxeliser timing on a case-sized mesh: 9 k triangles over a 1024 x 1024 x 256 domain (example_ProfileResearch-like triangle count)
import time, numpy as np, sys
sys.path.insert(0, ".")
from latticeurbanwind_b200 import _cabi as A
from latticeurbanwind_b200.domain import Domain
from tests import helpers as H
rng = np.random.default_rng(1)
Nx, Ny, Nz = 1024, 1024, 256
tris = H._box_tris((1.0, 1.0, 1.0), (Nx - 1.0, Ny - 1.0, 2.3))
for _ in range(765):
    cx, cy = rng.uniform(20, Nx - 20), rng.uniform(20, Ny - 20); w, d, h = rng.uniform(8, 30), rng.uniform(8, 30), rng.uniform(10, 120)
    tris += H._box_tris((cx - w / 2, cy - d / 2, 2.5), (cx + w / 2, cy + d / 2, 2.5 + h))
P = np.array(tris, np.float32); p0, p1, p2 = (np.ascontiguousarray(P[:, k, :]).reshape(-1) for k in range(3))
bbu = H.vox_bbu(P.shape[0], P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0))
with Domain(Nx, Ny, Nz, precision=1, features=63, w=1.9, arith=1, buffer_N=16, buffer_inv_tau=0.01, sponge_N=20, sponge_inv_tau=0.02) as d:
    d.upload_all(); d.finish_queue()
    t0 = time.perf_counter(); d.voxelize_mesh(2, 1, p0, p1, p2, bbu); dt = time.perf_counter() - t0
    d.read_from_device(A.FIELD_FLAGS); d.finish_queue()
    print("voxelize: %d triangles, %d columns, %.1f ms, %d solid cells" % (P.shape[0], Nx * Ny, dt * 1e3, int(((d.flags & 3) == 1).sum())))
PY
