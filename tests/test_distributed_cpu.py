"""CPU, world_size 2 over gloo: the host logic of the one-process-per-GPU driver (domain split, neighbour routing, x->y->z exchange
order, slot parity) with the CPU oracle standing in for the device kernels. A decomposed run must reproduce the single-domain run
bit for bit (the property the reference claims for its own multi-GPU path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latticeurbanwind_b200 import _cabi_consts as K
from latticeurbanwind_b200.lbm import DistributedLBM, split, _local_index
from oracle import oracle as O
from tests import helpers as H

SHAPE, STEPS = (24, 20, 16), 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, D, precision, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flags, rho, u = H.golden_case(SHAPE)
        feat = O.FEATURE_SETS["luw"]
        zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
        orc = O.Oracle()
        orc.set_threads(2)
        lbm = DistributedLBM(SHAPE, D, routing_only=True, precision=precision, features=feat, nu=1e-6, f=H.FORCE, omega=H.OMEGA, **zones)
        shape, Ov, fl, rh, uu = H.cut_block(SHAPE, D, lbm.d, flags, rho, u)
        assert shape == lbm.Nl and Ov == lbm.O
        p = O.make_params(*shape, precision, feat, w=lbm.w, D=D, O=Ov, **zones)
        orc.bind(p)
        fi = np.zeros(19 * p.N, O.ddf_dtype(precision))
        state = {"t": 1}

        def extract(payload, axis, sp, sm):
            if payload == K.HALO_FI:
                orc.extract_fi(axis, state["t"], sp.numpy().view(fi.dtype), sm.numpy().view(fi.dtype), fi)
            else:
                orc.extract_rho_u_flags(axis, sp.numpy(), sm.numpy(), rh, uu, fl)

        def insert(payload, axis, rp, rm):
            if payload == K.HALO_FI:
                orc.insert_fi(axis, state["t"], rp.numpy().view(fi.dtype), rm.numpy().view(fi.dtype), fi)
            else:
                orc.insert_rho_u_flags(axis, rp.numpy(), rm.numpy(), rh, uu, fl)

        # LBM::initialize (FX/lbm.cpp:1221-1260)
        lbm.communicate(K.HALO_RHO_U_FLAGS, extract, insert)
        orc.initialize(fi, rh, uu, fl)
        lbm.communicate(K.HALO_RHO_U_FLAGS, extract, insert)
        lbm.communicate(K.HALO_FI, extract, insert)
        for t in range(STEPS):  # do_time_step (FX/lbm.cpp:1262-1290)
            state["t"] = t
            orc.stream_collide(fi, rh, uu, fl, t, H.FORCE, H.OMEGA)
            lbm.communicate(K.HALO_FI, extract, insert)
        np.savez(os.path.join(out, f"rank{rank}.npz"), rho=rh, u=uu, gidx=lbm.gidx, Nl=np.array(lbm.Nl))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("D", [(2, 1, 1), (1, 1, 2), (1, 2, 1)], ids=["2x1x1", "1x1x2", "1x2x1"])
def test_two_ranks_reproduce_single_domain(tmp_path, D, oracle_lib):
    precision = O.FP16S
    mp.spawn(_worker, args=(2, _free_port(), D, precision, str(tmp_path)), nprocs=2, join=True)
    flags, rho, u = H.golden_case(SHAPE)
    zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
    from latticeurbanwind_b200 import cases
    ref = H.run_cpu(O.Oracle(), O, SHAPE, precision, O.FEATURE_SETS["luw"], flags, rho, u, STEPS, cases.relaxation_rate(1e-6), zones=zones)
    N = int(np.prod(SHAPE))
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        Nl = tuple(int(v) for v in z["Nl"])
        keep = np.ones(Nl[::-1], bool)
        for a in range(3):
            if D[a] > 1:
                sl = [slice(None)] * 3
                sl[2 - a] = [0, -1]
                keep[tuple(sl)] = False
        k, g = keep.reshape(-1), z["gidx"]
        assert np.array_equal(z["rho"][k], ref[1][g[k]])
        n = g.size
        for c in range(3):
            assert np.array_equal(z["u"][c * n:(c + 1) * n][k], ref[2][c * N + g[k]])


def _worker_thermal(rank, world, port, D, precision, out):
    """The thermal D3Q7 extension through the same routing: communicate_T + communicate_gi at the end of LBM::initialize, communicate_gi after
    communicate_fi in do_time_step (FX/lbm.cpp LBM::initialize / do_time_step), oracle kernels standing in for the device."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flags, rho, u, T = H.thermal_case(SHAPE)
        feat = O.FEATURE_SETS["luwT"]
        zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
        orc = O.Oracle()
        orc.set_threads(2)
        lbm = DistributedLBM(SHAPE, D, routing_only=True, precision=precision, features=feat, nu=1e-6, f=H.FORCE, omega=H.OMEGA, **zones)
        assert lbm.thermal
        shape, Ov, fl, rh, uu = H.cut_block(SHAPE, D, lbm.d, flags, rho, u)
        Tl = H.cut_block(SHAPE, D, lbm.d, flags, T, np.zeros(3 * T.size, np.float32))[3]
        p = O.make_params(*shape, precision, feat, w=lbm.w, D=D, O=Ov, **zones)
        orc.bind(p)
        orc.set_thermal(**H.THERMAL)
        dt = O.ddf_dtype(precision)
        fi, gi = np.zeros(19 * p.N, dt), np.zeros(7 * p.N, dt)
        state = {"t": 1}

        def extract(payload, axis, sp, sm):
            if payload == K.HALO_FI:
                orc.extract_fi(axis, state["t"], sp.numpy().view(dt), sm.numpy().view(dt), fi)
            elif payload == K.HALO_GI:
                orc.extract_gi(axis, state["t"], sp.numpy().view(dt), sm.numpy().view(dt), gi)
            elif payload == K.HALO_T:
                orc.extract_T(axis, sp.numpy().view(np.float32), sm.numpy().view(np.float32), Tl)
            else:
                orc.extract_rho_u_flags(axis, sp.numpy(), sm.numpy(), rh, uu, fl)

        def insert(payload, axis, rp, rm):
            if payload == K.HALO_FI:
                orc.insert_fi(axis, state["t"], rp.numpy().view(dt), rm.numpy().view(dt), fi)
            elif payload == K.HALO_GI:
                orc.insert_gi(axis, state["t"], rp.numpy().view(dt), rm.numpy().view(dt), gi)
            elif payload == K.HALO_T:
                orc.insert_T(axis, rp.numpy().view(np.float32), rm.numpy().view(np.float32), Tl)
            else:
                orc.insert_rho_u_flags(axis, rp.numpy(), rm.numpy(), rh, uu, fl)

        lbm.communicate(K.HALO_RHO_U_FLAGS, extract, insert)
        orc.initialize_thermal(fi, rh, uu, fl, gi, Tl)
        lbm.communicate(K.HALO_RHO_U_FLAGS, extract, insert)
        lbm.communicate(K.HALO_FI, extract, insert)
        lbm.communicate(K.HALO_T, extract, insert)
        lbm.communicate(K.HALO_GI, extract, insert)
        for t in range(STEPS):
            state["t"] = t
            orc.stream_collide_thermal(fi, rh, uu, fl, t, H.FORCE, H.OMEGA, gi, Tl)
            lbm.communicate(K.HALO_FI, extract, insert)
            lbm.communicate(K.HALO_GI, extract, insert)
        np.savez(os.path.join(out, f"rank{rank}.npz"), rho=rh, u=uu, T=Tl, gidx=lbm.gidx, Nl=np.array(lbm.Nl))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("D", [(2, 1, 1), (1, 1, 2)], ids=["2x1x1", "1x1x2"])
def test_two_ranks_reproduce_single_domain_with_temperature(tmp_path, D, oracle_lib):
    precision = O.FP16S
    mp.spawn(_worker_thermal, args=(2, _free_port(), D, precision, str(tmp_path)), nprocs=2, join=True)
    flags, rho, u, T = H.thermal_case(SHAPE)
    zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
    from latticeurbanwind_b200 import cases
    ref = H.run_cpu_thermal(O.Oracle(), O, SHAPE, precision, O.FEATURE_SETS["luwT"], flags, rho, u, T, STEPS, cases.relaxation_rate(1e-6), zones=zones)
    N = int(np.prod(SHAPE))
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        Nl = tuple(int(v) for v in z["Nl"])
        keep = np.ones(Nl[::-1], bool)
        for a in range(3):
            if D[a] > 1:
                sl = [slice(None)] * 3
                sl[2 - a] = [0, -1]
                keep[tuple(sl)] = False
        k, g = keep.reshape(-1), z["gidx"]
        assert np.array_equal(z["rho"][k], ref[1][g[k]])
        assert np.array_equal(z["T"][k], ref[4][g[k]])
        n = g.size
        for c in range(3):
            assert np.array_equal(z["u"][c * n:(c + 1) * n][k], ref[2][c * N + g[k]])


def test_split_matches_reference_rules():
    Ng, Nl, doms = split((751, 742, 174), (2, 3, 1))  # FX/lbm.cpp:1057-1073: round down to multiples of D, +2 halo layers, O = d*N/D - H
    assert Ng == (750, 741, 174) and Nl == (377, 249, 174)
    assert doms[0] == ((0, 0, 0), (-1, -1, 0)) and doms[-1] == ((1, 2, 0), (374, 493, 0))
    g = _local_index(Ng, Nl, doms[0][1])
    assert g[0] == (Ng[0] - 1) + Ng[0] * (Ng[1] - 1)  # local (0,0,0) is the periodic image of global (Nx-1, Ny-1, 0)
