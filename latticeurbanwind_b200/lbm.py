"""Host-side mirror of the reference's `LBM` object (FX/lbm.hpp:223-633, FX/lbm.cpp:1057-1312) over the C ABI.

    lbm = LBM((Nx, Ny, Nz), D=(Dx, Dy, Dz), nu=..., precision=..., features=...)
    lbm.flags[...] / lbm.u[...] / lbm.rho[...]      global host images, reference layout n = x + (y + z*Ny)*Nx, u = [ux | uy | uz]
    lbm.run(steps)                                   first call initialises (upload + `initialize` kernel + halo fill), then steps
    lbm.read_from_device()                           device -> global host images

Two drivers share the domain split and the step sequence of the reference (`do_time_step`: stream_collide on every domain, then
`communicate_fi` axis by axis x -> y -> z):
  * LBM             one process owns all Dx*Dy*Dz domains (any mix of devices; several domains may share one GPU, which is how
                    decomposed runs are tested on a single B200). Halo payloads move device-to-device.
  * DistributedLBM  one process per GPU (torch.distributed, NCCL): each rank owns ONE domain; halo payloads move with batched
                    isend/irecv over NVLink. This replaces the host-staged exchange of FX/lbm.cpp:1907-1935.
torch is used for device buffers, streams and the process group only; all LBM work is in the CUDA library behind the C ABI.
"""
import numpy as np

from . import _cabi as A
from . import cases
from .domain import Domain

AXES = (0, 1, 2)


def split(shape, D):
    """Domain split of FX/lbm.cpp:1057-1073: global N rounded down to multiples of D, local N/D + 2 halo layers on decomposed axes,
    offset O = d*N/D - H. Returns (global shape, local shape, list of (d, O))."""
    D = tuple(int(v) for v in D)
    Ng = tuple((int(n) // d) * d for n, d in zip(shape, D))
    H = tuple(1 if d > 1 else 0 for d in D)
    Nl = tuple(n // d + 2 * h for n, d, h in zip(Ng, D, H))
    doms = []
    for dz in range(D[2]):
        for dy in range(D[1]):
            for dx in range(D[0]):
                d = (dx, dy, dz)
                doms.append((d, tuple(d[a] * (Ng[a] // D[a]) - H[a] for a in AXES)))
    return Ng, Nl, doms


def _local_index(Ng, Nl, O):
    """Global linear indices of every local cell (periodic at the global edges), the stitching of FX/lbm.hpp:274-297."""
    idx = [np.mod(np.arange(Nl[a]) + O[a], Ng[a]) for a in AXES]
    g = idx[0][None, None, :] + Ng[0] * (idx[1][None, :, None] + Ng[1] * idx[2][:, None, None])
    return g.reshape(-1)


class _Base:
    def __init__(self, shape, D=(1, 1, 1), nu=1.0 / 6.0, precision=A.FP32, features=A.UPDATE_FIELDS, arith=A.ARITH_FAST,
                 f=(0.0, 0.0, 0.0), omega=(0.0, 0.0, 0.0), w=None, alpha=0.0, beta=0.0, T_avg=1.0, **zones):
        self.D = tuple(int(v) for v in D)
        self.Ng, self.Nl, self._split = split(shape, self.D)
        self.N = int(np.prod(self.Ng))
        self.precision, self.features, self.arith = precision, features, arith
        self.w = cases.relaxation_rate(nu) if w is None else w
        self.f, self.omega = tuple(f), tuple(omega)
        self.zones = zones
        # thermal D3Q7 extension (features & TEMPERATURE): LBM ctor arguments alpha (thermal diffusion coefficient) and beta (thermal expansion
        # coefficient), FX/lbm.cpp:1040-1047; def_w_T = 1/(2 alpha + 1/2) in float arithmetic like FX/lbm.cpp:750
        self.thermal = bool(features & A.TEMPERATURE)
        self.w_T = float(cases.kernel_literal(np.float32(1.0) / (np.float32(2.0) * np.float32(alpha) + np.float32(0.5))))
        self.beta, self.T_avg = float(cases.kernel_literal(beta)), float(cases.kernel_literal(T_avg))
        self.t = 0
        self.initialized = False

    def _make_domain(self, d, O, device):
        dom = Domain(*self.Nl, D=self.D, O=O, precision=self.precision, features=self.features, w=self.w, arith=self.arith,
                     device=device, **self.zones)
        dom.f, dom.omega = self.f, self.omega
        if self.thermal:
            dom.set_thermal(self.w_T, self.beta, self.T_avg)
        return dom

    def set_coriolis(self, ox, oy, oz):  # FX/lbm.hpp:496-498
        self.omega = (ox, oy, oz)
        for dom in self.domains:
            dom.omega = self.omega

    def get_N(self):
        return self.N

    def get_t(self):
        return self.t


class LBM(_Base):
    """All domains in one process (reference: `LBM` with `lbm_domain[d]`, FX/lbm.hpp:426)."""

    def __init__(self, shape, D=(1, 1, 1), devices=None, **kw):
        super().__init__(shape, D, **kw)
        ndev = max(A.device_count(), 1)
        self.devices = list(devices) if devices is not None else [i % ndev for i in range(len(self._split))]
        self.domains = [self._make_domain(d, O, dev) for (d, O), dev in zip(self._split, self.devices)]
        self.rho = np.ones(self.N, np.float32)
        self.u = np.zeros(3 * self.N, np.float32)
        self.flags = np.zeros(self.N, np.uint8)
        self.T = np.ones(self.N, np.float32) if self.thermal else None
        self._gidx = [_local_index(self.Ng, self.Nl, O) for _, O in self._split]

    # ---- host <-> device (Memory_Container::write_to_device / read_from_device, FX/lbm.hpp:406-423)
    def write_to_device(self):
        for dom, g in zip(self.domains, self._gidx):
            dom.rho[:] = self.rho[g]
            dom.flags[:] = self.flags[g]
            for c in range(3):
                dom.u[c * dom.N:(c + 1) * dom.N] = self.u[c * self.N + g]
            if self.thermal:
                dom.T[:] = self.T[g]
            dom.upload_all()

    def read_from_device(self):
        H = tuple(1 if d > 1 else 0 for d in self.D)
        for dom, g in zip(self.domains, self._gidx):
            dom.download_all()
            keep = np.ones(self.Nl[::-1], bool)
            if H[0]: keep[:, :, 0] = keep[:, :, -1] = False
            if H[1]: keep[:, 0, :] = keep[:, -1, :] = False
            if H[2]: keep[0, :, :] = keep[-1, :, :] = False
            k = keep.reshape(-1)
            self.rho[g[k]] = dom.rho[k]
            self.flags[g[k]] = dom.flags[k]
            for c in range(3):
                self.u[c * self.N + g[k]] = dom.u[c * dom.N:(c + 1) * dom.N][k]
            if self.thermal:
                self.T[g[k]] = dom.T[k]

    # ---- halo exchange (FX/lbm.cpp:1895-1958): extract, peer copies and insert are enqueued by the library, stream-ordered
    def _handles(self):
        import ctypes as C
        if not hasattr(self, "_harr"):
            self._harr = (C.c_void_p * len(self.domains))(*[dom._h for dom in self.domains])
        return self._harr

    def communicate(self, payload):
        t = self.domains[0].t
        for axis in AXES:
            A.check(A.lib().luw_halo_exchange(self._handles(), len(self.domains), payload, axis, t))

    # ---- LBM::initialize / do_time_step / run (FX/lbm.cpp:1221-1312)
    def initialize(self):
        self.write_to_device()
        for dom in self.domains:
            dom.t = 1
        self.communicate(A.HALO_RHO_U_FLAGS)
        for dom in self.domains:
            dom.enqueue_initialize()
        self.communicate(A.HALO_RHO_U_FLAGS)
        self.communicate(A.HALO_FI)
        if self.thermal:  # communicate_T(); communicate_gi(): FX/lbm.cpp LBM::initialize
            self.communicate(A.HALO_T)
            self.communicate(A.HALO_GI)
        for dom in self.domains:
            dom.finish_queue()
            dom.t = 0
        self.t = 0
        self.initialized = True

    def do_time_step(self):
        for dom in self.domains:
            dom.f, dom.omega = self.f, self.omega
            dom.enqueue_stream_collide()
        self.communicate(A.HALO_FI)
        if self.thermal:
            self.communicate(A.HALO_GI)
        for dom in self.domains:
            dom.increment_time_step()
        self.t += 1

    def run(self, steps=1):
        if not self.initialized:
            self.initialize()
        if len(self.domains) == 1:
            dom = self.domains[0]
            dom.f, dom.omega = self.f, self.omega
            dom.run_steps(steps)
            self.t += steps
        else:
            A.check(A.lib().luw_run_steps_multi(self._handles(), len(self.domains), self.t, steps, *map(float, self.f), *map(float, self.omega)))
            for dom in self.domains:
                dom.increment_time_step(steps)
            self.t += steps
        for dom in self.domains:
            dom.finish_queue()

    def update_fields(self):
        for dom in self.domains:
            dom.enqueue_update_fields()

    def close(self):
        for dom in self.domains:
            dom.close()


class DistributedLBM(_Base):
    """One domain per rank (torch.distributed); rank r owns domain r of the split. Works with NCCL (GPU) and, for the host logic, with
    gloo: with `routing_only=True` no device domain is created and the class only does what is host logic -- the domain split, neighbour ranks, the
    x -> y -> z exchange order on host tensors -- while the caller supplies the extract / insert callbacks (the CPU tests plug the oracle in there).
    There is no CPU fallback of the LBM itself: initialize() / run() need the device domain."""

    def __init__(self, shape, D, group=None, device=0, routing_only=False, transport="ipc", **kw):
        super().__init__(shape, D, **kw)
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world == len(self._split), "one rank per domain"
        self.d, self.O = self._split[self.rank]
        self.routing_only = bool(routing_only)
        self.device = device
        self.gidx = _local_index(self.Ng, self.Nl, self.O)
        if not self.routing_only:
            self.domain = self._make_domain(self.d, self.O, device)
            self.domains = [self.domain]
            # one explicit stream orders the step kernels, the halo kernels and the NCCL calls (torch's default stream has handle 0, which the C ABI
            # reads as "use the domain's own stream": kernels and NCCL would then run unordered on two streams)
            self._stream = torch.cuda.Stream(device=device)
            self.domain.set_stream(self._stream.cuda_stream)
            self.transport = transport
            if transport == "ipc":
                self._connect_ipc()
        else:
            self.domain, self.domains = None, []
        self._bufs = {}

    def _connect_ipc(self):
        """Exchange the CUDA IPC handles of the receive blocks (once) and map the two neighbours' blocks per decomposed axis. After this the
        step path makes no NCCL call: extract kernels store into the neighbours' memory over NVLink (luw_halo_ipc_exchange)."""
        dist = self._dist
        for axis in AXES:
            if self.D[axis] < 2:
                continue
            mine = self.domain.halo_ipc_export(axis)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=self.group)
            self.domain.halo_ipc_connect(axis, handles[self.rank_of(axis, +1)], handles[self.rank_of(axis, -1)])
        dist.barrier(group=self.group)

    def rank_of(self, axis, step):
        d = list(self.d)
        d[axis] = (d[axis] + step) % self.D[axis]
        return d[0] + self.D[0] * (d[1] + self.D[1] * d[2])

    def halo_bytes(self, payload, axis):
        A_ = (self.Nl[1] * self.Nl[2], self.Nl[2] * self.Nl[0], self.Nl[0] * self.Nl[1])[axis]
        ddf = 4 if self.precision == A.FP32 else 2
        per = {A.HALO_RHO_U_FLAGS: 17, A.HALO_FI: 5 * ddf, A.HALO_GI: ddf, A.HALO_T: 4}[payload]
        return per * A_

    def _buffers(self, payload, axis):
        key = (payload, axis)
        if key not in self._bufs:
            torch = self._torch
            dev = torch.device("cpu") if self.routing_only else torch.device("cuda", self.device)
            self._bufs[key] = tuple(torch.empty(self.halo_bytes(payload, axis), dtype=torch.uint8, device=dev) for _ in range(4))
        return self._bufs[key]

    def exchange(self, axis, sp, sm, rp, rm):
        """send_p -> (+) neighbour's recv_m, send_m -> (-) neighbour's recv_p. With 2 ranks on an axis both neighbours are the same rank."""
        dist = self._dist
        up, dn = self.rank_of(axis, +1), self.rank_of(axis, -1)
        if up == self.rank:  # a single domain on this axis never gets here; kept for completeness
            rm.copy_(sp); rp.copy_(sm)
            return
        ops = [dist.P2POp(dist.isend, sp, up, self.group, tag=2 * axis), dist.P2POp(dist.isend, sm, dn, self.group, tag=2 * axis + 1),
               dist.P2POp(dist.irecv, rm, dn, self.group, tag=2 * axis), dist.P2POp(dist.irecv, rp, up, self.group, tag=2 * axis + 1)]
        if not self.routing_only:
            with self._torch.cuda.stream(self._stream):  # NCCL enqueues on the current stream
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
        else:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def communicate(self, payload, extract, insert):
        for axis in AXES:
            if self.D[axis] < 2:
                continue
            if not self.routing_only and self.transport == "ipc":
                self.domain.halo_ipc_exchange(payload, axis)
                continue
            sp, sm, rp, rm = self._buffers(payload, axis)
            extract(payload, axis, sp, sm)
            self.exchange(axis, sp, sm, rp, rm)
            insert(payload, axis, rp, rm)

    # device-side callbacks
    def _extract(self, payload, axis, sp, sm):
        self.domain.halo_extract(payload, axis, sp.data_ptr(), sm.data_ptr())

    def _insert(self, payload, axis, rp, rm):
        self.domain.halo_insert(payload, axis, rp.data_ptr(), rm.data_ptr())

    def upload_slabs(self, slabs):
        """`slabs`: an iterable of (z0, flags, rho, u) pieces covering the local lattice plane range by plane range (z0 = first local z plane of the piece), uploaded
        as they come, so that a block of 10^9 cells never needs its 17 B per cell on the host. Follow with initialize() without images."""
        dom = self.domain
        plane, N = self.Nl[0] * self.Nl[1], dom.N
        for z0, fl, rh, uu in slabs:
            n = fl.size
            dom.upload_range(A.FIELD_FLAGS, fl, z0 * plane)
            dom.upload_range(A.FIELD_RHO, rh, z0 * plane)
            for c in range(3):
                dom.upload_range(A.FIELD_U, uu[c * n:(c + 1) * n], c * N + z0 * plane)
            dom.finish_queue()

    def initialize(self, flags=None, rho=None, u=None, T=None):
        """flags / rho / u (/ T with TEMPERATURE): LOCAL host images of this rank's domain (halo layers included); none at all after upload_slabs()."""
        if self.domain is None:
            raise RuntimeError("DistributedLBM(routing_only=True) has no device domain: the LBM step has no CPU fallback")
        dom = self.domain
        dom.f, dom.omega = self.f, self.omega
        if flags is None:
            pass  # the device fields are in place (upload_slabs)
        else:
            dom.rho[:], dom.u[:], dom.flags[:] = rho, u, flags
            if self.thermal and T is not None:
                dom.T[:] = T
            dom.upload_all()
        dom.t = 1
        self.communicate(A.HALO_RHO_U_FLAGS, self._extract, self._insert)
        dom.enqueue_initialize()
        self.communicate(A.HALO_RHO_U_FLAGS, self._extract, self._insert)
        self.communicate(A.HALO_FI, self._extract, self._insert)
        if self.thermal:
            self.communicate(A.HALO_T, self._extract, self._insert)
            self.communicate(A.HALO_GI, self._extract, self._insert)
        dom.finish_queue()
        dom.t = 0
        self.t = 0
        self.initialized = True

    def do_time_step(self):
        dom = self.domain
        if not self.routing_only and self.transport == "ipc":  # one call: step + exchanges, the y / z exchanges overlapped with the interior strips (luw_step_halo_ipc)
            dom.step_halo_ipc()
        else:
            dom.enqueue_stream_collide()
            self.communicate(A.HALO_FI, self._extract, self._insert)
            if self.thermal:
                self.communicate(A.HALO_GI, self._extract, self._insert)
        dom.increment_time_step()
        self.t += 1

    def run(self, steps=1):
        for _ in range(steps):
            self.do_time_step()

    def close(self):
        if self.domain is not None:
            self.domain.finish_queue()
            try:  # the neighbours store into this rank's IPC-mapped receive blocks: nobody frees before everybody has drained its stream
                self._dist.barrier(group=self.group)
            except Exception:
                pass
            self.domain.close()
            self.domain = None
