"""Development aid: one CUDA-vs-oracle comparison from the command line, e.g. under compute-sanitizer.
usage: python dev/run_case.py Nx Ny Nz precision fset arith steps [case]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from latticeurbanwind_b200 import cases
from oracle import oracle as O
from tests import helpers as H
Nx, Ny, Nz, prec, fset, arith, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6]), int(sys.argv[7])
case = sys.argv[8] if len(sys.argv) > 8 else "urban"
shape = (Nx, Ny, Nz)
flags, rho, u = (cases.urban(*shape, seed=77, edge=3, pitch=6) if case == "urban" else cases.CASES[case](*shape))
w = cases.relaxation_rate(1e-6 if case == "urban" else 1 / 6)
feat = H.FEATURE_SETS[fset]
zones = dict(downstream_face=2, buffer_N=3, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=2, sponge_inv_tau=0.02)
upd = not (feat & 1)
ref = H.run_cpu(O.Oracle(), O, shape, prec, feat, flags, rho, u, steps, w, zones=zones, update_at_end=upd)
got = H.run_cuda(shape, prec, feat, flags, rho, u, steps, w, arith=arith, zones=zones, update_at_end=upd)
a, b = H.decode(O, None, got[0], prec), H.decode(O, None, ref[0], prec)
bad = np.nonzero(a != b)[0]
N = Nx * Ny * Nz
print("fi mismatches:", bad.size, "of", a.size)
for k in bad[:12]:
    n = k % N
    print("  slot", k // N, "cell", (n % Nx, (n // Nx) % Ny, n // (Nx * Ny)), "got", a[k], "want", b[k])
print("rho equal:", np.array_equal(got[1], ref[1]), " u equal:", np.array_equal(got[2], ref[2]),
      " rel_l2(u):", H.rel_l2(got[2], ref[2]), " max|du|:", float(np.abs(got[2] - ref[2]).max()), " rel_l2(rho):", H.rel_l2(got[1], ref[1]))
