"""Development aid: run a few steps of a 2-domain decomposition with both domains on device 0 (under `ncu --metrics gpu__time_duration.sum`)
to see what the halo kernels of each axis cost.  usage: python dev/halo_times.py <axis>"""
import sys
import numpy as np
sys.path.insert(0, ".")
from latticeurbanwind_b200 import cases
from latticeurbanwind_b200.lbm import LBM
axis = int(sys.argv[1])
D = [1, 1, 1]; D[axis] = 2
shape = [512, 512, 256]
shape[axis] = 2 * (shape[axis] - 2) if axis != 2 else 2 * (512 - 2)
if axis == 2:
    shape = [512, 256, 1020]
lbm = LBM(tuple(shape), D=tuple(D), devices=[0, 0], nu=1 / 6, precision=1, features=4, arith=1)
flags, rho, u = cases.block_case("channel", lbm.Ng)
lbm.flags[:], lbm.rho[:], lbm.u[:] = flags, rho, u
lbm.run(4)
lbm.close()
print("done", lbm.Nl)
