"""Development aid: time the step kernel of several tile variants on one workload, building the host case once.
usage: python dev/variant_sweep.py workload v0,v1,... [steps] [warmup]   (prints one line per variant; LUW_TILE_VARIANT is read at domain creation)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench as B
from latticeurbanwind_b200 import _cabi as A, cases
from latticeurbanwind_b200.domain import Domain

workload, variants = sys.argv[1], [v for v in sys.argv[2].split(",")]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 10
case, shape, precision, features, fset, nu, desc = B.WORKLOADS[workload]
N = int(np.prod(shape))
flags, rho, u = cases.block_case(case, shape)
zones = B.ZONES if features & (B.F_NUDGE | B.F_SPONGE) else {}
peak, _ = B.measured_peaks()
for v in variants:
    if v == "d":
        os.environ.pop("LUW_TILE_VARIANT", None)
    else:
        os.environ["LUW_TILE_VARIANT"] = v
    d = Domain(*shape, precision=precision, features=features, w=cases.relaxation_rate(nu), arith=A.ARITH_FAST, **zones)
    d.flags[:], d.rho[:], d.u[:] = flags, rho, u
    d.omega = B.OMEGA if features & B.F_VF else (0.0, 0.0, 0.0)
    d.upload_all(); d.t = 1; d.enqueue_initialize(); d.t = 0
    d.run_steps(warm); d.finish_queue()
    best = 1e30
    for rep in range(3):
        d.timer_begin(); d.run_steps(steps); ms = d.timer_end()
        best = min(best, ms / steps)
    d.read_from_device(A.FIELD_RHO, 0, 1024); d.finish_queue()
    mlups = N / best / 1e3
    print(json.dumps({"workload": workload, "variant": v, "tiles": d.uses_tiles(), "ms_per_step": round(best, 4), "mlups": round(mlups), "frac": round(mlups * B.alg_bytes(precision, features) / 1e3 / peak, 4)}), flush=True)
    d.close()
