"""GPU, world_size 2 (two devices; on a one-GPU box both ranks share device 0 and rendezvous over gloo -- CUDA IPC works between processes on one device, NCCL does
not, so only the IPC transport runs there): the one-process-per-GPU driver (DistributedLBM: step kernel, halo extract, payload over
NVLink -- remote stores into the neighbour's IPC-mapped receive block, or NCCL send/recv -- halo insert, all ordered on one stream) reproduces the
single-domain run of the same kernels bit for bit."""
import os
import socket

import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from tests import helpers as H

pytestmark = pytest.mark.gpu
SHAPE, STEPS = (256, 24, 16), 9
ZONES = dict(downstream_face=2, buffer_N=3, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=3, sponge_inv_tau=0.02)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, D, precision, arith, transport, out):
    import torch
    import torch.distributed as dist
    from latticeurbanwind_b200.lbm import DistributedLBM
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    shared = torch.cuda.device_count() < world  # both ranks on one GPU: the contexts time-slice; every device-side wait is bounded
    dev = 0 if shared else rank
    torch.cuda.set_device(dev)
    if shared:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    try:
        flags, rho, u = cases.urban(*SHAPE, seed=21, edge=4, pitch=8)
        lbm = DistributedLBM(SHAPE, D, device=dev, nu=1e-6, precision=precision, features=H.FEATURE_SETS["luw"], arith=arith, f=H.FORCE, omega=H.OMEGA, transport=transport, **ZONES)
        shape, Ov, fl, rh, uu = H.cut_block(SHAPE, D, lbm.d, flags, rho, u)
        assert shape == lbm.Nl and Ov == lbm.O
        lbm.initialize(fl, rh, uu)
        lbm.run(STEPS)
        lbm.domain.download_all()
        np.savez(os.path.join(out, f"rank{rank}.npz"), rho=lbm.domain.rho, u=lbm.domain.u, gidx=lbm.gidx, Nl=np.array(lbm.Nl), overlapped=np.array(lbm.domain.overlapped_steps()))
        lbm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["ipc", "nccl"])
@pytest.mark.parametrize("D,arith", [((1, 2, 1), 1), ((1, 1, 2), 0), ((2, 1, 1), 0), ((1, 2, 2), 1)], ids=["1x2x1-fast", "1x1x2-strict", "2x1x1-strict", "1x2x2-fast-4ranks"])
def test_two_ranks_nccl_reproduce_single_domain(tmp_path, D, arith, transport):
    import torch
    import torch.multiprocessing as mp
    world = D[0] * D[1] * D[2]
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    if torch.cuda.device_count() < world and transport == "nccl":
        pytest.skip("NCCL needs one device per rank")
    from latticeurbanwind_b200.lbm import LBM
    precision = 1
    mp.spawn(_worker, args=(world, _free_port(), D, precision, arith, transport, str(tmp_path)), nprocs=world, join=True)
    flags, rho, u = cases.urban(*SHAPE, seed=21, edge=4, pitch=8)
    one = LBM(SHAPE, D=(1, 1, 1), nu=1e-6, precision=precision, features=H.FEATURE_SETS["luw"], arith=arith, f=H.FORCE, omega=H.OMEGA, **ZONES)
    one.flags[:], one.rho[:], one.u[:] = flags, rho, u
    one.run(STEPS)
    one.read_from_device()
    N = int(np.prod(SHAPE))
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        if transport == "ipc":  # y / z splits: every step's exchange ran on the halo stream, overlapped with the interior strips (luw_step_halo_ipc); x faces involve every strip
            assert int(z["overlapped"]) == (STEPS if D[0] == 1 else 0), int(z["overlapped"])
        Nl = tuple(int(v) for v in z["Nl"])
        keep = np.ones(Nl[::-1], bool)
        for a in range(3):
            if D[a] > 1:
                sl = [slice(None)] * 3
                sl[2 - a] = [0, -1]
                keep[tuple(sl)] = False
        k, g = keep.reshape(-1), z["gidx"]
        assert np.array_equal(z["rho"][k], one.rho[g[k]])
        n = g.size
        for c in range(3):
            assert np.array_equal(z["u"][c * n:(c + 1) * n][k], one.u[c * N + g[k]])
    one.close()
