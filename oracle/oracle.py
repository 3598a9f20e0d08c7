"""TEST INFRASTRUCTURE: ctypes front-ends for the CPU oracle (oracle/libluw_oracle.so) and for the reference's own kernel
text compiled for host threads (oracle/_ref/libluwref_<prec>_<set>.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Both front-ends expose the same methods so that a test can run the same script against either:
    initialize(fi, rho, u, flags); stream_collide(fi, rho, u, flags, t, f, omega); update_fields(...);
    extract_fi / insert_fi / extract_rho_u_flags / insert_rho_u_flags; vk_inlet_apply(...)
Arrays are numpy, SoA exactly like the reference buffers: fi[19*N] (float32 or uint16), u[3*N], rho[N], flags[N] uint8,
n = x + (y + z*Ny)*Nx.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FP32, FP16S, FP16C = 0, 1, 2
PREC_NAME = {FP32: "fp32", FP16S: "fp16s", FP16C: "fp16c"}
UPDATE_FIELDS, VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, SUBGRID, BUFFER_NUDGING, TOP_SPONGE = 1, 2, 4, 8, 16, 32
TEMPERATURE = 64  # thermal D3Q7: selects the *_thermal entry points (the C oracle takes gi / T per call and ignores the bit)
FEATURE_SETS = {  # must match oracle/Makefile
    "bench": 0,
    "chan": EQUILIBRIUM_BOUNDARIES,
    "plain": UPDATE_FIELDS | EQUILIBRIUM_BOUNDARIES,
    "core": UPDATE_FIELDS | VOLUME_FORCE | EQUILIBRIUM_BOUNDARIES | SUBGRID,
    "luw": UPDATE_FIELDS | VOLUME_FORCE | EQUILIBRIUM_BOUNDARIES | SUBGRID | BUFFER_NUDGING | TOP_SPONGE,
    "luwnf": VOLUME_FORCE | EQUILIBRIUM_BOUNDARIES | SUBGRID | BUFFER_NUDGING | TOP_SPONGE,
    "luwT": UPDATE_FIELDS | VOLUME_FORCE | EQUILIBRIUM_BOUNDARIES | SUBGRID | BUFFER_NUDGING | TOP_SPONGE | TEMPERATURE,
    "chanT": VOLUME_FORCE | EQUILIBRIUM_BOUNDARIES | TEMPERATURE,
}
TYPE_S, TYPE_E, TYPE_T = 0x01, 0x02, 0x04


class Thermal(C.Structure):
    """luwo_thermal -- def_w_T, def_beta, def_T_avg (FX/lbm.cpp:750-752)."""
    _fields_ = [("w_T", C.c_float), ("beta", C.c_float), ("T_avg", C.c_float)]


class Params(C.Structure):
    """luwo_params (oracle/luw_oracle.h) -- per-domain constants of FX/lbm.cpp:612-783."""
    _fields_ = [("Nx", C.c_uint32), ("Ny", C.c_uint32), ("Nz", C.c_uint32),
                ("Dx", C.c_uint32), ("Dy", C.c_uint32), ("Dz", C.c_uint32),
                ("Ox", C.c_int32), ("Oy", C.c_int32), ("Oz", C.c_int32),
                ("precision", C.c_uint32), ("features", C.c_uint32), ("w", C.c_float),
                ("downstream_face", C.c_int32), ("buffer_N", C.c_uint32), ("buffer_inv_tau", C.c_float),
                ("buffer_nudge_vertical", C.c_int32), ("sponge_N", C.c_uint32), ("sponge_inv_tau", C.c_float)]

    @property
    def N(self):
        return int(self.Nx) * int(self.Ny) * int(self.Nz)


class _RefParams(C.Structure):
    _fields_ = [("Nx", C.c_uint32), ("Ny", C.c_uint32), ("Nz", C.c_uint32),
                ("Dx", C.c_uint32), ("Dy", C.c_uint32), ("Dz", C.c_uint32),
                ("Ox", C.c_int32), ("Oy", C.c_int32), ("Oz", C.c_int32), ("w", C.c_float),
                ("downstream_face", C.c_int32), ("buffer_N", C.c_uint32), ("buffer_inv_tau", C.c_float),
                ("buffer_nudge_vertical", C.c_int32), ("sponge_N", C.c_uint32), ("sponge_inv_tau", C.c_float)]


def make_params(Nx, Ny, Nz, precision=FP32, features=0, w=1.0, D=(1, 1, 1), O=(0, 0, 0), downstream_face=0,
                buffer_N=1, buffer_inv_tau=0.0, buffer_nudge_vertical=0, sponge_N=1, sponge_inv_tau=0.0):
    return Params(Nx, Ny, Nz, D[0], D[1], D[2], O[0], O[1], O[2], precision, features, np.float32(w),
                  downstream_face, buffer_N, np.float32(buffer_inv_tau), buffer_nudge_vertical, sponge_N, np.float32(sponge_inv_tau))


def build(force=False):
    """Compile oracle/libluw_oracle.so (and oracle/_ref when the reference tree is present). Building the checker is not using it."""
    if force or not os.path.isfile(os.path.join(HERE, "libluw_oracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "libluw_oracle.so")) < os.path.getmtime(os.path.join(HERE, "luw_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "-j8", "libluw_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", HERE, "-j8", "ref"], stdout=subprocess.DEVNULL)


def ddf_dtype(precision):
    return np.float32 if precision == FP32 else np.uint16


def _p(a, ty=C.c_void_p):
    return a.ctypes.data_as(ty)


class Oracle:
    """Our C restatement."""
    kind = "port"

    def __init__(self):
        path = os.path.join(HERE, "libluw_oracle.so")
        if not os.path.isfile(path):
            build()
        L = self.lib = C.CDLL(path)
        L.luwo_half_to_float.restype = C.c_float; L.luwo_half_to_float.argtypes = [C.c_uint16]
        L.luwo_float_to_half_rte.restype = C.c_uint16; L.luwo_float_to_half_rte.argtypes = [C.c_float]
        L.luwo_fp16c_to_float.restype = C.c_float; L.luwo_fp16c_to_float.argtypes = [C.c_uint16]
        L.luwo_float_to_fp16c.restype = C.c_uint16; L.luwo_float_to_fp16c.argtypes = [C.c_float]
        L.luwo_calculate_f_eq.argtypes = [C.c_float] * 4 + [C.c_void_p]
        PP = C.POINTER(Params)
        f6 = [C.c_float] * 6
        L.luwo_initialize.argtypes = [PP] + [C.c_void_p] * 4
        L.luwo_stream_collide.argtypes = [PP] + [C.c_void_p] * 4 + [C.c_uint64] + f6
        L.luwo_update_fields.argtypes = [PP] + [C.c_void_p] * 4 + [C.c_uint64] + f6
        L.luwo_transfer_extract_fi.argtypes = [PP, C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        L.luwo_transfer_insert_fi.argtypes = [PP, C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        L.luwo_transfer_extract_rho_u_flags.argtypes = [PP, C.c_uint32] + [C.c_void_p] * 5
        L.luwo_transfer_insert_rho_u_flags.argtypes = [PP, C.c_uint32] + [C.c_void_p] * 5
        L.luwo_vk_inlet_apply.argtypes = [C.c_uint64, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_void_p] * 5
        L.luwo_set_threads.argtypes = [C.c_int]
        L.luwo_get_threads.restype = C.c_int
        PT = C.POINTER(Thermal)
        L.luwo_calculate_g_eq.argtypes = [C.c_float] * 4 + [C.c_void_p]
        L.luwo_initialize_thermal.argtypes = [PP] + [C.c_void_p] * 6
        L.luwo_stream_collide_thermal.argtypes = [PP, PT] + [C.c_void_p] * 4 + [C.c_uint64] + f6 + [C.c_void_p] * 2
        L.luwo_update_fields_thermal.argtypes = [PP, PT] + [C.c_void_p] * 4 + [C.c_uint64] + f6 + [C.c_void_p] * 2
        L.luwo_transfer_extract_gi.argtypes = [PP, C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        L.luwo_transfer_insert_gi.argtypes = [PP, C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        L.luwo_transfer_extract_T.argtypes = [PP, C.c_uint32] + [C.c_void_p] * 3
        L.luwo_transfer_insert_T.argtypes = [PP, C.c_uint32] + [C.c_void_p] * 3
        self.th = Thermal(1.0, 0.0, 1.0)

    def set_threads(self, n):
        self.lib.luwo_set_threads(int(n))

    def threads(self):
        return int(self.lib.luwo_get_threads())

    def bind(self, params):
        self.p = params
        return self

    # codecs
    def half_to_float(self, h): return self.lib.luwo_half_to_float(int(h))
    def float_to_half(self, f): return self.lib.luwo_float_to_half_rte(float(f))
    def fp16c_to_float(self, h): return self.lib.luwo_fp16c_to_float(int(h))
    def float_to_fp16c(self, f): return self.lib.luwo_float_to_fp16c(float(f))

    def f_eq(self, rho, ux, uy, uz):
        out = np.zeros(19, np.float32)
        self.lib.luwo_calculate_f_eq(float(rho), float(ux), float(uy), float(uz), _p(out))
        return out

    def initialize(self, fi, rho, u, flags):
        self.lib.luwo_initialize(C.byref(self.p), _p(fi), _p(rho), _p(u), _p(flags))

    def stream_collide(self, fi, rho, u, flags, t, f=(0, 0, 0), omega=(0, 0, 0)):
        self.lib.luwo_stream_collide(C.byref(self.p), _p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega))

    def update_fields(self, fi, rho, u, flags, t, f=(0, 0, 0), omega=(0, 0, 0)):
        self.lib.luwo_update_fields(C.byref(self.p), _p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega))

    def extract_fi(self, direction, t, bp, bm, fi):
        self.lib.luwo_transfer_extract_fi(C.byref(self.p), direction, int(t), _p(bp), _p(bm), _p(fi))

    def insert_fi(self, direction, t, bp, bm, fi):
        self.lib.luwo_transfer_insert_fi(C.byref(self.p), direction, int(t), _p(bp), _p(bm), _p(fi))

    def extract_rho_u_flags(self, direction, bp, bm, rho, u, flags):
        self.lib.luwo_transfer_extract_rho_u_flags(C.byref(self.p), direction, _p(bp), _p(bm), _p(rho), _p(u), _p(flags))

    def insert_rho_u_flags(self, direction, bp, bm, rho, u, flags):
        self.lib.luwo_transfer_insert_rho_u_flags(C.byref(self.p), direction, _p(bp), _p(bm), _p(rho), _p(u), _p(flags))

    # thermal D3Q7 (gi: 7*N DDFs in the storage type of fi, T: N floats)
    def set_thermal(self, w_T, beta=0.0, T_avg=1.0):
        self.th = Thermal(np.float32(w_T), np.float32(beta), np.float32(T_avg))

    def g_eq(self, T, ux, uy, uz):
        out = np.zeros(7, np.float32)
        self.lib.luwo_calculate_g_eq(float(T), float(ux), float(uy), float(uz), _p(out))
        return out

    def initialize_thermal(self, fi, rho, u, flags, gi, T):
        self.lib.luwo_initialize_thermal(C.byref(self.p), _p(fi), _p(rho), _p(u), _p(flags), _p(gi), _p(T))

    def stream_collide_thermal(self, fi, rho, u, flags, t, f, omega, gi, T):
        self.lib.luwo_stream_collide_thermal(C.byref(self.p), C.byref(self.th), _p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega), _p(gi), _p(T))

    def update_fields_thermal(self, fi, rho, u, flags, t, f, omega, gi, T):
        self.lib.luwo_update_fields_thermal(C.byref(self.p), C.byref(self.th), _p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega), _p(gi), _p(T))

    def extract_gi(self, direction, t, bp, bm, gi):
        self.lib.luwo_transfer_extract_gi(C.byref(self.p), direction, int(t), _p(bp), _p(bm), _p(gi))

    def insert_gi(self, direction, t, bp, bm, gi):
        self.lib.luwo_transfer_insert_gi(C.byref(self.p), direction, int(t), _p(bp), _p(bm), _p(gi))

    def extract_T(self, direction, bp, bm, T):
        self.lib.luwo_transfer_extract_T(C.byref(self.p), direction, _p(bp), _p(bm), _p(T))

    def insert_T(self, direction, bp, bm, T):
        self.lib.luwo_transfer_insert_T(C.byref(self.p), direction, _p(bp), _p(bm), _p(T))

    def voxelize_mesh(self, direction, u, flags, flag, p0, p1, p2, bbu):
        self.lib.luwo_voxelize_mesh.restype = None
        self.lib.luwo_voxelize_mesh(C.byref(self.p), C.c_uint32(direction), _p(u), _p(flags), C.c_uint8(flag), _p(p0), _p(p1), _p(p2), _p(bbu))

    def vk_inlet_apply(self, use_interp, t0, t1, alpha, point_cell, point_face, point_data, mode_data, mode_count, mode_stride, u):
        P = point_cell.shape[0]
        self.lib.luwo_vk_inlet_apply(self.p.N, int(use_interp), float(t0), float(t1), float(alpha), P, int(mode_count), int(mode_stride),
                                     _p(point_cell), _p(point_face), _p(point_data), _p(mode_data), _p(u))


def stats_accumulate(lib_fn, count, rho, u, u_avg, rho_avg, m2_u, m2_v, m2_w):
    lib_fn(C.c_uint64(rho.size), count, _p(rho), _p(u), _p(u_avg), _p(rho_avg), _p(m2_u), _p(m2_v), _p(m2_w))


class RefStats:
    """The reference's own averaging loop (FX/setup.cpp:4441-4488), text extracted by ref_shim/make_ref_stats.py and built into oracle/_ref/libluwref_stats.so."""
    PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libluwref_stats.so")

    @classmethod
    def available(cls):
        return os.path.isfile(cls.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)
        self.lib.luwref_stats_accumulate.restype = None
        self.count = C.c_uint64(0)

    def accumulate(self, rho, u, u_avg, rho_avg, m2_u, m2_v, m2_w):
        stats_accumulate(self.lib.luwref_stats_accumulate, C.byref(self.count), rho, u, u_avg, rho_avg, m2_u, m2_v, m2_w)


class OracleStats:
    """C restatement (luwo_stats_accumulate)."""

    def __init__(self):
        self.lib = Oracle().lib
        self.lib.luwo_stats_accumulate.restype = None
        self.count = 0

    def accumulate(self, rho, u, u_avg, rho_avg, m2_u, m2_v, m2_w):
        self.count += 1
        stats_accumulate(self.lib.luwo_stats_accumulate, C.c_uint64(self.count), rho, u, u_avg, rho_avg, m2_u, m2_v, m2_w)


def ref_available(precision=FP32, feature_set="luw"):
    return os.path.isfile(os.path.join(HERE, "_ref", f"libluwref_{PREC_NAME[precision]}_{feature_set}.so"))


class Reference:
    """The reference's own kernel text (FX/kernel.cpp) compiled for host threads through oracle/ref_shim."""
    kind = "reference"

    def __init__(self, precision, feature_set):
        path = os.path.join(HERE, "_ref", f"libluwref_{PREC_NAME[precision]}_{feature_set}.so")
        L = self.lib = C.CDLL(path)
        self.precision, self.feature_set = precision, feature_set
        L.luwref_features.restype = C.c_uint32
        L.luwref_sizeof_fpxx.restype = C.c_uint32
        assert L.luwref_features() == FEATURE_SETS[feature_set]
        assert L.luwref_sizeof_fpxx() == (4 if precision == FP32 else 2)
        f6 = [C.c_float] * 6
        L.luwref_set_params.argtypes = [C.POINTER(_RefParams)]
        if FEATURE_SETS[feature_set] & TEMPERATURE:  # built with -DTEMPERATURE: the kernels carry gi / T as trailing arguments
            L.luwref_set_thermal.argtypes = [C.c_float] * 3
            L.luwref_initialize_thermal.argtypes = [C.c_void_p] * 6
            L.luwref_stream_collide_thermal.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + f6 + [C.c_void_p] * 2
            L.luwref_update_fields_thermal.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + f6 + [C.c_void_p] * 2
            for name in ("luwref_transfer_extract_gi", "luwref_transfer_insert_gi", "luwref_transfer_extract_T", "luwref_transfer_insert_T"):
                getattr(L, name).argtypes = [C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        else:
            L.luwref_initialize.argtypes = [C.c_void_p] * 4
            L.luwref_stream_collide.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + f6
            L.luwref_update_fields.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + f6
        L.luwref_transfer_extract_fi.argtypes = [C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        L.luwref_transfer_insert_fi.argtypes = [C.c_uint32, C.c_uint64] + [C.c_void_p] * 3
        L.luwref_transfer_extract_rho_u_flags.argtypes = [C.c_uint32, C.c_uint64] + [C.c_void_p] * 5
        L.luwref_transfer_insert_rho_u_flags.argtypes = [C.c_uint32, C.c_uint64] + [C.c_void_p] * 5
        L.luwref_vk_inlet_apply.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_void_p] * 5
        L.luwref_half_to_float_custom.restype = C.c_float; L.luwref_half_to_float_custom.argtypes = [C.c_uint16]
        L.luwref_float_to_half_custom.restype = C.c_uint16; L.luwref_float_to_half_custom.argtypes = [C.c_float]
        L.luwref_calculate_f_eq.argtypes = [C.c_float] * 4 + [C.c_void_p]
        L.luwref_voxelize_mesh.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint8] + [C.c_void_p] * 4

    def bind(self, params):
        assert params.precision == self.precision and params.features == FEATURE_SETS[self.feature_set]
        assert params.N <= 0xFFFFFFFF  # the shim fixes uxx=uint, like the reference does for N<=2^32-1 (FX/lbm.cpp:631)
        self.p = params
        rp = _RefParams(params.Nx, params.Ny, params.Nz, params.Dx, params.Dy, params.Dz, params.Ox, params.Oy, params.Oz, params.w,
                        params.downstream_face, params.buffer_N, params.buffer_inv_tau, params.buffer_nudge_vertical,
                        params.sponge_N, params.sponge_inv_tau)
        self.lib.luwref_set_params(C.byref(rp))
        return self

    def fp16c_to_float(self, h): return self.lib.luwref_half_to_float_custom(int(h))
    def float_to_fp16c(self, f): return self.lib.luwref_float_to_half_custom(float(f))

    def f_eq(self, rho, ux, uy, uz):
        out = np.zeros(19, np.float32)
        self.lib.luwref_calculate_f_eq(float(rho), float(ux), float(uy), float(uz), _p(out))
        return out

    def initialize(self, fi, rho, u, flags):
        self.lib.luwref_initialize(_p(fi), _p(rho), _p(u), _p(flags))

    def stream_collide(self, fi, rho, u, flags, t, f=(0, 0, 0), omega=(0, 0, 0)):
        self.lib.luwref_stream_collide(_p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega))

    def update_fields(self, fi, rho, u, flags, t, f=(0, 0, 0), omega=(0, 0, 0)):
        self.lib.luwref_update_fields(_p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega))

    def extract_fi(self, direction, t, bp, bm, fi):
        self.lib.luwref_transfer_extract_fi(direction, int(t), _p(bp), _p(bm), _p(fi))

    def insert_fi(self, direction, t, bp, bm, fi):
        self.lib.luwref_transfer_insert_fi(direction, int(t), _p(bp), _p(bm), _p(fi))

    def extract_rho_u_flags(self, direction, bp, bm, rho, u, flags):
        self.lib.luwref_transfer_extract_rho_u_flags(direction, 0, _p(bp), _p(bm), _p(rho), _p(u), _p(flags))

    def insert_rho_u_flags(self, direction, bp, bm, rho, u, flags):
        self.lib.luwref_transfer_insert_rho_u_flags(direction, 0, _p(bp), _p(bm), _p(rho), _p(u), _p(flags))

    def set_thermal(self, w_T, beta=0.0, T_avg=1.0):
        self.lib.luwref_set_thermal(float(np.float32(w_T)), float(np.float32(beta)), float(np.float32(T_avg)))

    def initialize_thermal(self, fi, rho, u, flags, gi, T):
        self.lib.luwref_initialize_thermal(_p(fi), _p(rho), _p(u), _p(flags), _p(gi), _p(T))

    def stream_collide_thermal(self, fi, rho, u, flags, t, f, omega, gi, T):
        self.lib.luwref_stream_collide_thermal(_p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega), _p(gi), _p(T))

    def update_fields_thermal(self, fi, rho, u, flags, t, f, omega, gi, T):
        self.lib.luwref_update_fields_thermal(_p(fi), _p(rho), _p(u), _p(flags), int(t), *map(float, f), *map(float, omega), _p(gi), _p(T))

    def extract_gi(self, direction, t, bp, bm, gi):
        self.lib.luwref_transfer_extract_gi(direction, int(t), _p(bp), _p(bm), _p(gi))

    def insert_gi(self, direction, t, bp, bm, gi):
        self.lib.luwref_transfer_insert_gi(direction, int(t), _p(bp), _p(bm), _p(gi))

    def extract_T(self, direction, bp, bm, T):
        self.lib.luwref_transfer_extract_T(direction, 0, _p(bp), _p(bm), _p(T))

    def insert_T(self, direction, bp, bm, T):
        self.lib.luwref_transfer_insert_T(direction, 0, _p(bp), _p(bm), _p(T))

    def voxelize_mesh(self, direction, u, flags, flag, p0, p1, p2, bbu, t=1):
        fi = np.zeros(19 * self.p.N, ddf_dtype(self.precision))  # only touched for moving geometry
        self.lib.luwref_voxelize_mesh(direction, _p(fi), _p(u), _p(flags), int(t), int(flag), _p(p0), _p(p1), _p(p2), _p(bbu))

    def vk_inlet_apply(self, use_interp, t0, t1, alpha, point_cell, point_face, point_data, mode_data, mode_count, mode_stride, u):
        P = point_cell.shape[0]
        self.lib.luwref_vk_inlet_apply(int(use_interp), float(t0), float(t1), float(alpha), P, int(mode_count), int(mode_stride),
                                       _p(point_cell), _p(point_face), _p(point_data), _p(mode_data), _p(u))
