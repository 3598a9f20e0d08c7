"""CPU: the package's CUDA kernel SOURCE (csrc/lbm_kernels.cuh), compiled with g++ behind a thin emulation of blockIdx / threadIdx / the few
intrinsics it uses (tests/host_emulation/), run over the same launch grid on host threads, against the oracle -- bit for bit, in the device memory
layout (padded row pitch). This is how kernel logic is checked in the container without a GPU: the thermal D3Q7 kernels were written
after this round's GPU budget was spent, so this test -- not a B200 run -- is their parity evidence until tests/test_thermal_gpu.py has been observed
on the device. The plain stream_collide (GPU-verified) runs through the same harness first, which validates the harness itself.
Nothing in the package can load this library; it is test infrastructure like oracle/."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from latticeurbanwind_b200 import cases
from oracle import oracle as O
from tests import helpers as H

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emulation")
pytestmark = pytest.mark.skipif(not os.path.isfile("/usr/local/cuda/include/cuda_fp16.h"), reason="CUDA headers not installed")
PRECS = pytest.mark.parametrize("precision", [0, 1, 2], ids=["fp32", "fp16s", "fp16c"])


class StepArgs(C.Structure):
    _fields_ = [("t", C.c_uint64)] + [(n, C.c_float) for n in ("fx", "fy", "fz", "ox", "oy", "oz")]


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "libluw_kernels_on_host.so"))
    L.emu_sizeof_domain_const.restype = C.c_uint64
    L.emu_make_domain.argtypes = [C.c_void_p] + [C.c_uint32] * 6 + [C.c_int] * 4 + [C.c_uint32, C.c_float, C.c_int, C.c_uint32, C.c_float, C.c_int, C.c_uint32] + \
        [C.c_void_p] * 8 + [C.c_float] * 3
    L.emu_zone_tables.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
    for name in ("emu_initialize", "emu_initialize_thermal"):
        getattr(L, name).argtypes = [C.c_void_p]
    for name in ("emu_stream_collide", "emu_stream_collide_thermal", "emu_update_fields_thermal"):
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(StepArgs)]
    L.emu_halo_gi.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.emu_halo_T.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostDomain:
    """Device-layout arrays (rows padded to a multiple of 16 elements, like luw_domain_create) + the DomainConst the kernels take."""

    def __init__(self, L, shape, precision, features, w, zones, D=(1, 1, 1), Ov=(0, 0, 0), thermal=None):
        self.L, self.shape, self.precision = L, shape, precision
        Nx, Ny, Nz = shape
        self.Px = (Nx + 15) & ~15
        self.Nd = self.Px * Ny * Nz
        dt = O.ddf_dtype(precision)
        self.fi, self.gi = np.zeros(19 * self.Nd, dt), np.zeros(7 * self.Nd, dt)
        self.rho, self.u, self.flags, self.T = np.ones(self.Nd, np.float32), np.zeros(3 * self.Nd, np.float32), np.zeros(self.Nd, np.uint8), np.ones(self.Nd, np.float32)
        self.wbuf, self.sigma = np.zeros(zones["buffer_N"] + 1, np.float32), np.zeros(zones["sponge_N"], np.float32)
        L.emu_zone_tables(zones["buffer_N"], zones["sponge_N"], np.float32(zones["sponge_inv_tau"]), _p(self.wbuf), _p(self.sigma))
        self.c = C.create_string_buffer(int(L.emu_sizeof_domain_const()))
        th = thermal or dict(w_T=1.0, beta=0.0, T_avg=1.0)
        rc = L.emu_make_domain(self.c, Nx, Ny, Nz, *D, *Ov, precision, features, np.float32(w), zones["downstream_face"], zones["buffer_N"],
                               np.float32(zones["buffer_inv_tau"]), zones["buffer_nudge_vertical"], zones["sponge_N"], _p(self.wbuf), _p(self.sigma),
                               _p(self.fi), _p(self.rho), _p(self.u), _p(self.flags), _p(self.gi), _p(self.T),
                               np.float32(th["w_T"]), np.float32(th["beta"]), np.float32(th["T_avg"]))
        assert rc == 0

    def put(self, dev, dense, comps):
        Nx, Ny, Nz = self.shape
        dev.reshape(comps, Nz, Ny, self.Px)[..., :Nx] = dense.reshape(comps, Nz, Ny, Nx)

    def get(self, dev, comps):
        Nx, Ny, Nz = self.shape
        return np.ascontiguousarray(dev.reshape(comps, Nz, Ny, self.Px)[..., :Nx]).reshape(-1)

    def args(self, t, f, omega):
        return StepArgs(t, *map(np.float32, f), *map(np.float32, omega))


def test_harness_reproduces_the_gpu_verified_step(emu, oracle_lib):
    """Validation of the harness: the plain (GPU-verified) stream_collide through it equals the oracle."""
    shape = (20, 18, 14)
    flags, rho, u = H.small_urban(*shape)
    w = cases.relaxation_rate(1e-6)
    feat = O.FEATURE_SETS["luw"]
    for precision in (O.FP32, O.FP16S, O.FP16C):
        want = H.run_cpu(O.Oracle(), O, shape, precision, feat, flags, rho, u, 6, w)
        d = HostDomain(emu, shape, precision, feat, w, H.ZONES)
        d.put(d.rho, rho, 1); d.put(d.u, u, 3); d.put(d.flags, flags, 1)
        assert emu.emu_initialize(d.c) == 0
        for t in range(6):
            assert emu.emu_stream_collide(d.c, C.byref(d.args(t, H.FORCE, H.OMEGA))) == 0
        for got, ref, name in zip((d.get(d.fi, 19), d.get(d.rho, 1), d.get(d.u, 3)), want, ("fi", "rho", "u")):
            assert np.array_equal(got, ref), (precision, name)


@PRECS
@pytest.mark.parametrize("fset", ["luwT", "chanT"])
def test_thermal_kernels_equal_the_oracle(emu, oracle_lib, precision, fset):
    flags, rho, u, T = H.thermal_case()
    w = cases.relaxation_rate(1e-6)
    feat = O.FEATURE_SETS[fset]
    steps = H.THERMAL_STEPS
    want = H.run_cpu_thermal(O.Oracle(), O, H.THERMAL_SHAPE, precision, feat, flags, rho, u, T, steps, w, update_at_end=(fset == "chanT"))
    d = HostDomain(emu, H.THERMAL_SHAPE, precision, feat, w, H.ZONES, thermal=H.THERMAL)
    d.put(d.rho, rho, 1); d.put(d.u, u, 3); d.put(d.flags, flags, 1); d.put(d.T, T, 1)
    assert emu.emu_initialize_thermal(d.c) == 0
    for t in range(steps):
        assert emu.emu_stream_collide_thermal(d.c, C.byref(d.args(t, H.FORCE, H.OMEGA))) == 0
    if fset == "chanT":
        assert emu.emu_update_fields_thermal(d.c, C.byref(d.args(steps, H.FORCE, H.OMEGA))) == 0
    got = (d.get(d.fi, 19), d.get(d.rho, 1), d.get(d.u, 3), d.get(d.gi, 7), d.get(d.T, 1))
    for g, r, name in zip(got, want, ("fi", "rho", "u", "gi", "T")):
        assert np.array_equal(g, r), name


@PRECS
def test_thermal_halo_kernels_equal_the_oracle(emu, oracle_lib, precision):
    """Block (0,0,0) of a 2x2x2 decomposition: gi / T payloads in the reference's face order (xfast = 0) and the lattice after inserting them."""
    flags, rho, u, T = H.thermal_case(H.GOLDEN_SHAPE)
    Ncell = int(np.prod(H.GOLDEN_SHAPE))
    shape, Ov, fl, rh, ul = H.cut_block(H.GOLDEN_SHAPE, (2, 2, 2), (0, 0, 0), flags, rho, u)
    Tl = H.cut_block(H.GOLDEN_SHAPE, (2, 2, 2), (0, 0, 0), flags, T, np.zeros(3 * Ncell, np.float32))[3]
    zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
    feat = O.FEATURE_SETS["luwT"]
    w = cases.relaxation_rate(1e-6)
    want = H.golden_thermal_halo(O.Oracle(), O, precision)
    name = O.PREC_NAME[precision]
    d = HostDomain(emu, shape, precision, feat, w, zones, D=(2, 2, 2), Ov=Ov, thermal=H.THERMAL)
    d.put(d.rho, rh, 1); d.put(d.u, ul, 3); d.put(d.flags, fl, 1); d.put(d.T, Tl, 1)
    assert emu.emu_initialize_thermal(d.c) == 0
    Nx, Ny, Nz = shape
    for t in range(3):
        assert emu.emu_stream_collide_thermal(d.c, C.byref(d.args(t, H.FORCE, H.OMEGA))) == 0
        for axis in range(3):
            A = (Ny * Nz, Nz * Nx, Nx * Ny)[axis]
            bp, bm = np.zeros(A, d.gi.dtype), np.zeros(A, d.gi.dtype)
            assert emu.emu_halo_gi(d.c, axis, t & 1, 0, 0, _p(bp), _p(bm)) == 0
            assert np.array_equal(bp, want[f"thalo_{name}_t{t}_axis{axis}_p"]) and np.array_equal(bm, want[f"thalo_{name}_t{t}_axis{axis}_m"]), (t, axis)
            assert emu.emu_halo_gi(d.c, axis, t & 1, 1, 0, _p(bm), _p(bp)) == 0
        assert np.array_equal(d.get(d.gi, 7), want[f"thalo_{name}_t{t}_gi"]), t
    for axis in range(3):
        A = (Ny * Nz, Nz * Nx, Nx * Ny)[axis]
        bp, bm = np.zeros(A, np.float32), np.zeros(A, np.float32)
        assert emu.emu_halo_T(d.c, axis, 0, 0, _p(bp), _p(bm)) == 0
        assert np.array_equal(bp, want[f"thalo_{name}_T_axis{axis}_p"]) and np.array_equal(bm, want[f"thalo_{name}_T_axis{axis}_m"]), axis
        assert emu.emu_halo_T(d.c, axis, 1, 0, _p(bm), _p(bp)) == 0
    assert np.array_equal(d.get(d.T, 1), want[f"thalo_{name}_T"])
    # xfast payload order (library-internal buffers): extract + insert with swapped buffers must give the same lattice as the reference order
    g0 = d.gi.copy()
    for order in (0, 1):
        d.gi[:] = g0
        for axis in range(3):
            A = (Ny * Nz, Nz * Nx, Nx * Ny)[axis]
            bp, bm = np.zeros(A, d.gi.dtype), np.zeros(A, d.gi.dtype)
            emu.emu_halo_gi(d.c, axis, 1, 0, order, _p(bp), _p(bm))
            emu.emu_halo_gi(d.c, axis, 1, 1, order, _p(bm), _p(bp))
        if order == 0:
            ref = d.gi.copy()
    assert np.array_equal(d.gi, ref)


def test_thermal_kernels_on_a_wide_padded_lattice(emu, oracle_lib):
    """Rows wider than one thread block with a padded pitch (Nx = 150 -> Px = 160): the x >= Nx guard and the pitched neighbour / reference-cell indices."""
    shape = (150, 12, 10)
    flags, rho, u, T = H.thermal_case(shape, seed=9)
    w = cases.relaxation_rate(1e-6)
    feat = O.FEATURE_SETS["luwT"]
    want = H.run_cpu_thermal(O.Oracle(), O, shape, O.FP16S, feat, flags, rho, u, T, 5, w)
    d = HostDomain(emu, shape, O.FP16S, feat, w, H.ZONES, thermal=H.THERMAL)
    assert d.Px == 160
    d.put(d.rho, rho, 1); d.put(d.u, u, 3); d.put(d.flags, flags, 1); d.put(d.T, T, 1)
    assert emu.emu_initialize_thermal(d.c) == 0
    for t in range(5):
        assert emu.emu_stream_collide_thermal(d.c, C.byref(d.args(t, H.FORCE, H.OMEGA))) == 0
    got = (d.get(d.fi, 19), d.get(d.rho, 1), d.get(d.u, 3), d.get(d.gi, 7), d.get(d.T, 1))
    for g, r, name in zip(got, want, ("fi", "rho", "u", "gi", "T")):
        assert np.array_equal(g, r), name
    pad = d.T.reshape(shape[2], shape[1], d.Px)[..., shape[0]:]
    assert np.all(pad == 1.0)  # the padding columns are never written


@pytest.mark.parametrize("Ny,Nz,Dy,Dz,TY,TZ", [(1024, 256, 8, 1, 4, 1), (512, 512, 2, 4, 2, 1), (14, 10, 2, 2, 4, 1), (13, 9, 2, 2, 4, 1), (40, 24, 1, 2, 2, 1), (40, 24, 2, 1, 2, 1),
                                               (9, 7, 2, 2, 4, 1), (33, 18, 2, 2, 1, 1), (6, 4, 2, 2, 4, 1), (20, 16, 1, 1, 4, 1)])
def test_boundary_first_strip_order(emu, Ny, Nz, Dy, Dz, TY, TZ):
    """luw_step_halo_ipc overlaps the halo exchange with the interior of the step: the step kernel hands out the strips that hold layers 0, 1, N-2, N-1 of the decomposed
    y / z axes first (csrc/lbm_common.cuh strip_of) and counts them. The order must be a permutation of all strips, the first so_nb of it must be exactly the strips that
    touch those layers, and no interior strip may touch them -- whatever the tile shape, partial tiles included."""
    Ty, Tz = -(-Ny // TY), -(-Nz // TZ)
    out, order = np.zeros(5, np.uint32), np.zeros(Ty * Tz, np.uint32)
    emu.emu_strip_order.argtypes = [C.c_uint32] * 6 + [C.c_void_p, C.c_void_p]
    ok = emu.emu_strip_order(Ny, Nz, Dy, Dz, TY, TZ, out.ctypes.data, order.ctypes.data)
    nb = int(out[4])
    assert sorted(order.tolist()) == list(range(Ty * Tz)), "not a permutation"

    def touches(strip):
        ty, tz = strip % Ty, strip // Ty
        rows = set(range(ty * TY, min((ty + 1) * TY, Ny)))
        planes = set(range(tz * TZ, min((tz + 1) * TZ, Nz)))
        return (Dy > 1 and bool(rows & {0, 1, Ny - 2, Ny - 1})) or (Dz > 1 and bool(planes & {0, 1, Nz - 2, Nz - 1}))
    boundary = {s for s in range(Ty * Tz) if touches(s)}
    if not ok:  # nothing decomposed, or no interior left: natural order, nothing signalled
        assert nb == 0 and order.tolist() == list(range(Ty * Tz))
        return
    assert set(order[:nb].tolist()) == boundary, (sorted(set(order[:nb].tolist()) ^ boundary))
    assert not (set(order[nb:].tolist()) & boundary)


# ---------------------------------------------------------------------------------------------- binned mesh voxeliser (csrc/vox_bins.h + k_voxelize_mesh_binned)
def _city_mesh(shape, boxes, seed):
    """Many small closed boxes at fractional positions over a base slab: most triangles reach one or two bins only."""
    rng = np.random.default_rng(seed)
    Nx, Ny, Nz = shape
    tris = H._box_tris((1.0, 1.0, 1.0), (Nx - 1.0, Ny - 1.0, 2.3))
    for _ in range(boxes):
        cx, cy = rng.uniform(4, Nx - 4), rng.uniform(4, Ny - 4)
        w, d, h = rng.uniform(1.2, 5.5), rng.uniform(1.2, 5.5), rng.uniform(2.0, Nz - 5.0)
        tris += H._box_tris((cx - w / 2, cy - d / 2, 1.7), (cx + w / 2, cy + d / 2, 1.7 + h))
    P = np.array(tris, np.float32)
    p0, p1, p2 = (np.ascontiguousarray(P[:, k, :]).reshape(-1) for k in range(3))
    return p0, p1, p2, P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0)


def _voxelize_both(emu, oracle_lib, shape, mesh, direction, D=(1, 1, 1), Ov=(0, 0, 0), preset=False):
    p0, p1, p2, pmin, pmax = mesh
    ntri = p0.size // 3
    bbu = H.vox_bbu(ntri, pmin, pmax)
    N = int(np.prod(shape))
    flags, u = np.zeros(N, np.uint8), np.zeros(3 * N, np.float32)
    if preset:
        rng = np.random.default_rng(1)
        flags[:] = rng.choice(np.array([0, 1, 2, 0x40, 0x81], np.uint8), N)
        u[rng.integers(0, 3 * N, 500)] = 0.01
    p = oracle_lib.make_params(*shape, oracle_lib.FP16S, oracle_lib.FEATURE_SETS["luw"], D=D, O=Ov, **H.ZONES)
    want = flags.copy()
    oracle_lib.Oracle().bind(p).voxelize_mesh(direction, u.copy(), want, 1, p0, p1, p2, bbu)
    d = HostDomain(emu, shape, 1, H.FEATURE_SETS["luw"], 1.0, H.ZONES, D=D, Ov=Ov)
    d.put(d.flags, flags, 1); d.put(d.u, u, 3)
    stats = np.zeros(3, np.uint64)
    emu.emu_voxelize_binned.argtypes = [C.c_void_p, C.c_uint32, C.c_uint8] + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p]
    assert emu.emu_voxelize_binned(d.c, direction, 1, _p(p0), _p(p1), _p(p2), ntri, _p(bbu), _p(stats)) == 0
    return d.get(d.flags, 1), want, stats, ntri


@pytest.mark.parametrize("direction", [2, 0, 1], ids=["z-rays", "x-rays", "y-rays"])
@pytest.mark.parametrize("preset", [False, True], ids=["empty", "preset"])
def test_binned_voxelizer_source_equals_the_oracle(emu, oracle_lib, direction, preset):
    """Each block walks only the triangles whose padded projected box reaches its 32 x 4 columns: the flags must be those of the all-triangles voxeliser (the oracle,
    pinned to the reference's kernel text in test_voxelize.py), on the mesh with grazing rays and on lattices with pre-set flags."""
    got, want, stats, ntri = _voxelize_both(emu, oracle_lib, H.VOX_SHAPE, H.vox_mesh(), direction, preset=preset)
    assert np.array_equal(got, want)
    assert 500 < int(((got & 3) == 1).sum()) < got.size // 2


def test_binned_voxelizer_in_a_decomposed_block(emu, oracle_lib):
    Nx, Ny, Nz = H.VOX_SHAPE
    shape, D, Ov = (Nx // 2 + 2, Ny, Nz // 2 + 2), (2, 1, 2), (Nx // 2 - 1, 0, Nz // 2 - 1)
    got, want, _, _ = _voxelize_both(emu, oracle_lib, shape, H.vox_mesh(), 2, D=D, Ov=Ov)
    assert np.array_equal(got, want) and ((got & 3) == 1).sum() > 50


@pytest.mark.parametrize("direction", [2, 1], ids=["z-rays", "y-rays"])
def test_binned_voxelizer_on_a_city_of_small_boxes(emu, oracle_lib, direction):
    """400 boxes on a 210 x 150 x 30 lattice (odd extents: partial bins on both axes): identical flags, and the lists are a small fraction of columns x triangles."""
    shape = (210, 150, 30)
    got, want, stats, ntri = _voxelize_both(emu, oracle_lib, shape, _city_mesh(shape, 400, 5), direction)
    assert np.array_equal(got, want)
    assert int(((got & 3) == 1).sum()) > 20000
    bins, entries, longest = (int(v) for v in stats)
    assert entries * 128 < 0.2 * ntri * bins * 128 and longest < ntri  # work in ray/triangle tests: bins x 128 x list length vs columns x all triangles


STL = os.path.join(os.path.dirname(HERE), os.pardir, "baseline", "_ref", "case_profile", "proj_temp", "CaseE_PF.stl")


@pytest.mark.skipif(not os.path.isfile(STL), reason="baseline/_ref/case_profile (the reference's example project, staged by baseline/build_reference_driver.py) is not here")
def test_binned_voxelizer_on_the_reference_example_mesh(emu, oracle_lib):
    """The example project's building mesh (9 210 triangles) scaled onto the example's 253 x 250 x 59 lattice: identical flags with ~70 times fewer ray / triangle tests
    (about 1 % of its triangles are ill-conditioned in projection and are tested by every column, csrc/vox_bins.h "Conditioning")."""
    raw = open(STL, "rb").read()
    n = int(np.frombuffer(raw[80:84], np.uint32)[0])
    rec = np.frombuffer(raw[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    V = rec["v"].astype(np.float32)
    lo, hi = V.reshape(-1, 3).min(0), V.reshape(-1, 3).max(0)
    shape = (253, 250, 59)
    scale = np.float32(0.8 * min(shape[0] / (hi[0] - lo[0]), shape[1] / (hi[1] - lo[1])))
    V = ((V - lo) * scale + np.array([shape[0] * 0.1, shape[1] * 0.1, 0.7], np.float32)).astype(np.float32)
    mesh = tuple(np.ascontiguousarray(V[:, k, :]).reshape(-1) for k in range(3)) + (V.reshape(-1, 3).min(0), V.reshape(-1, 3).max(0))
    got, want, stats, ntri = _voxelize_both(emu, oracle_lib, shape, mesh, 2)
    assert np.array_equal(got, want) and int(((got & 3) == 1).sum()) > 10000
    assert int(stats[1]) * 128 * 30 < ntri * shape[0] * shape[1]


def _rotated_prisms(shape, count, seed):
    """Extruded buildings with oblique walls (footprints rotated by arbitrary angles, one of them by exactly 45 degrees through lattice points) and a tilted roof each:
    vertical triangles whose projection along z is a line segment (zero projected area), the case in which a barycentric ray test is at its most fragile."""
    rng = np.random.default_rng(seed)
    Nx, Ny, Nz = shape
    tris = H._box_tris((1.0, 1.0, 1.0), (Nx - 1.0, Ny - 1.0, 2.3))
    for k in range(count):
        cx, cy = rng.uniform(10, Nx - 10), rng.uniform(10, Ny - 10)
        w, d, h = rng.uniform(2.5, 7.5), rng.uniform(2.5, 7.5), rng.uniform(3.0, Nz - 6.0)
        ang = np.pi / 4 if k == 0 else rng.uniform(0, np.pi)
        if k == 0:
            cx, cy, w, d = 20.5, 20.5, 8.0 * np.sqrt(2.0), 4.0 * np.sqrt(2.0)  # corners on lattice points, walls along the diagonals through cell centres
        c, s = np.cos(ang), np.sin(ang)
        corners = [(cx + c * a - s * b, cy + s * a + c * b) for a, b in ((-w / 2, -d / 2), (w / 2, -d / 2), (w / 2, d / 2), (-w / 2, d / 2))]
        z0, top = 1.7, [1.7 + h, 1.7 + h + 0.8, 1.7 + h + 1.1, 1.7 + h + 0.3]
        lo = [np.array([x, y, z0], np.float32) for x, y in corners]
        hi = [np.array([x, y, z], np.float32) for (x, y), z in zip(corners, top)]
        tris += [(lo[0], lo[2], lo[1]), (lo[0], lo[3], lo[2]), (hi[0], hi[1], hi[2]), (hi[0], hi[2], hi[3])]
        for i in range(4):
            j = (i + 1) % 4
            tris += [(lo[i], lo[j], hi[j]), (lo[i], hi[j], hi[i])]
    P = np.array(tris, np.float32)
    p0, p1, p2 = (np.ascontiguousarray(P[:, k, :]).reshape(-1) for k in range(3))
    return p0, p1, p2, P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0)


@pytest.mark.parametrize("direction", [2, 0], ids=["z-rays", "x-rays"])
def test_binned_voxelizer_with_oblique_vertical_walls(emu, oracle_lib, direction):
    shape = (96, 80, 28)
    got, want, stats, ntri = _voxelize_both(emu, oracle_lib, shape, _rotated_prisms(shape, 40, 17), direction)
    assert np.array_equal(got, want), f"{int((got != want).sum())} cells differ"
    assert int(((got & 3) == 1).sum()) > 5000


def test_binned_voxelizer_with_degenerate_slivers(emu, oracle_lib):
    """1 500 vertical slivers whose three corners project onto one line up to rounding (half of them through lattice points): the reference's barycentric test answers
    with rounding noise for rays near such a line, several sliver lengths away -- and those answers are its flags. The bin grid hands such triangles to every column
    (csrc/vox_bins.h "Conditioning"; without that rule this mesh differs from the all-triangles result in ~100 cells), so the flags stay identical."""
    shape = (96, 80, 28)
    rng = np.random.default_rng(3)
    tris = H._box_tris((1.0, 1.0, 1.0), (95.0, 79.0, 2.3))
    for k in range(1500):
        if k % 2 == 0:
            ax, ay, dx, dy = rng.integers(5, 60) + 0.0, rng.integers(5, 50) + 0.0, rng.integers(1, 4) * 1.0, rng.integers(1, 4) * 1.0
        else:
            ax, ay, dx, dy = rng.uniform(5, 60), rng.uniform(5, 50), rng.uniform(-3, 3), rng.uniform(-3, 3)
        t1, t2 = rng.uniform(0.5, 6), rng.uniform(0.5, 6)
        tris.append((np.array([ax, ay, rng.uniform(2, 10)], np.float32), np.array([ax + t1 * dx, ay + t1 * dy, rng.uniform(2, 10)], np.float32),
                     np.array([ax + t2 * dx, ay + t2 * dy, rng.uniform(10, 20)], np.float32)))
    P = np.array(tris, np.float32)
    mesh = tuple(np.ascontiguousarray(P[:, k, :]).reshape(-1) for k in range(3)) + (P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0))
    for direction in (2, 0):
        got, want, stats, ntri = _voxelize_both(emu, oracle_lib, shape, mesh, direction)
        assert np.array_equal(got, want), f"direction {direction}: {int((got != want).sum())} cells differ"


@pytest.mark.parametrize("D", [(2, 2, 1), (2, 1, 2)], ids=["2x2x1", "2x1x2"])
def test_binned_voxelizer_in_every_block_of_a_decomposition(emu, oracle_lib, D):
    """Blocks with halo layers start one cell outside the lattice (offset -1): the bin grid follows the block's offset like the ray origins do."""
    from latticeurbanwind_b200.lbm import split
    Ng, Nl, doms = split(H.VOX_SHAPE, D)
    for d, Ov in doms:
        got, want, _, _ = _voxelize_both(emu, oracle_lib, tuple(Nl), H.vox_mesh(), 2, D=D, Ov=tuple(Ov))
        assert np.array_equal(got, want), (d, Ov)


@pytest.mark.parametrize("trial", [0, 1, 2, 3])
def test_binned_voxelizer_on_triangle_soup(emu, oracle_lib, trial):
    """800 unrelated triangles from a third of a cell to several lattice widths in size, some horizontal, some vertical, some with coinciding or lattice-point corners,
    over lattices with and without pre-set flags, three ray directions: the flags of the all-triangles voxeliser, cell for cell."""
    shape = (70, 45, 33)
    rng = np.random.default_rng(100 + trial)
    n = 800
    scale = rng.choice([0.3, 2.0, 8.0, 40.0], n)[:, None, None]
    centre = rng.uniform(-5, 75, (n, 1, 3)) * np.array([1, 0.7, 0.5])
    P = (centre + rng.normal(0, 1, (n, 3, 3)) * scale).astype(np.float32)
    P[:50, :, 2] = P[:50, :1, 2]
    P[50:100, 1, :2] = P[50:100, 0, :2]
    P[100:120, 2] = P[100:120, 1]
    P[120:140] = np.round(P[120:140])
    mesh = tuple(np.ascontiguousarray(P[:, k, :]).reshape(-1) for k in range(3)) + (P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0))
    for direction in (2, 0, 1):
        got, want, _, _ = _voxelize_both(emu, oracle_lib, shape, mesh, direction, preset=bool(trial % 2))
        assert np.array_equal(got, want), f"direction {direction}: {int((got != want).sum())} cells differ"
