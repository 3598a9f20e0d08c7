"""Development aid: run a few steps of one configuration (for ncu). usage: prof_case.py Nx Ny Nz precision features arith case steps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latticeurbanwind_b200 import cases
from latticeurbanwind_b200.domain import Domain
Nx, Ny, Nz, prec, feat, arith = (int(v) for v in sys.argv[1:7])
case, steps = sys.argv[7], int(sys.argv[8])
zones = dict(downstream_face=2, buffer_N=16, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=20, sponge_inv_tau=0.02)
with Domain(Nx, Ny, Nz, precision=prec, features=feat, w=cases.relaxation_rate(1e-6 if feat & 8 else 1 / 6), arith=arith, **zones) as d:
    if case == "periodic_box":
        d.rho[:] = 1; d.u[:] = 0; d.u[:Nx * Ny * Nz] = 0.05
    else:
        flags, rho, u = cases.CASES[case](Nx, Ny, Nz)
        d.rho[:], d.u[:], d.flags[:] = rho, u, flags
    d.omega = (0, 5.6e-6, 4.7e-6)
    d.upload_all(); d.t = 1; d.enqueue_initialize(); d.t = 0
    d.run_steps(steps); d.finish_queue()
    d.timer_begin(); d.run_steps(steps); ms = d.timer_end()
    print(f"{ms/steps:.3f} ms/step, {Nx*Ny*Nz*steps/ms/1e3:.0f} MLUPs, tiles={d.uses_tiles()}")
