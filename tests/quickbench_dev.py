import os, sys, time, numpy as np
sys.path.insert(0,'.')
from latticeurbanwind_b200 import cases, _cabi as A
from latticeurbanwind_b200.domain import Domain
print(A.device_info(0).name, "variant", os.environ.get("LUW_TILE_VARIANT","0"), "notile", os.environ.get("LUW_NO_TILE","0"))
def bench(shape, prec, feat, arith, steps=30, case="periodic_box"):
    Nx,Ny,Nz=shape
    zones=dict(downstream_face=2,buffer_N=16,buffer_inv_tau=0.01,buffer_nudge_vertical=1,sponge_N=20,sponge_inv_tau=0.02)
    with Domain(Nx,Ny,Nz,precision=prec,features=feat,w=cases.relaxation_rate(1e-6 if feat&8 else 1/6),arith=arith,**zones) as d:
        if case=="periodic_box":
            d.rho[:]=1; d.u[:]=0; d.u[:Nx*Ny*Nz]=0.05
        else:
            flags,rho,u=cases.CASES[case](Nx,Ny,Nz)
            d.rho[:],d.u[:],d.flags[:]=rho,u,flags
        d.omega=(0,5.6e-6,4.7e-6)
        d.upload_all(); d.t=1; d.enqueue_initialize(); d.t=0
        d.run_steps(6); d.finish_queue()
        d.timer_begin(); d.run_steps(steps); ms=d.timer_end()
        mlups=Nx*Ny*Nz*steps/ms/1e3
        B=153 if prec==0 else 77
        print(f"{case} {shape} prec={prec} feat={feat} arith={arith} tiles={d.uses_tiles()}: {ms/steps:.3f} ms/step {mlups:.0f} MLUPs {mlups*B/1e3:.0f} GB/s alg", flush=True)
precs = [int(x) for x in os.environ.get("QB_PRECS","0,1,2").split(",")]
for prec in precs:
    for arith in (0,1):
        bench((512,512,512),prec,0,arith)
for prec in precs:
    for arith in (0,1):
        bench((512,512,256),prec,1|2|4|8|16|32,arith,case="urban")
        bench((512,512,256),prec,2|4|8|16|32,arith,case="urban")
