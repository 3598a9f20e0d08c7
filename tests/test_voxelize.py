"""Mesh voxelisation (kernel voxelize_mesh FX/kernel.cpp:2381-2471): C oracle vs the reference kernel text (CPU), CUDA kernel vs
the oracle (GPU). Bar: flags bit-exact."""
import numpy as np
import pytest

from tests import helpers as H

TYPE_S = 1


def _run(engine_voxelize, direction, D=(1, 1, 1), Ov=(0, 0, 0), shape=H.VOX_SHAPE, preset=False):
    p0, p1, p2, pmin, pmax = H.vox_mesh()
    N = int(np.prod(shape))
    flags = np.zeros(N, np.uint8)
    u = np.zeros(3 * N, np.float32)
    if preset:  # cells that are already solid / equilibrium: solid ones outside the mesh with u == 0 are released, others keep their bits
        rng = np.random.default_rng(1)
        flags[:] = rng.choice(np.array([0, 1, 2, 0x40, 0x81], np.uint8), N)
        u[rng.integers(0, 3 * N, 500)] = 0.01
    engine_voxelize(direction, u, flags, TYPE_S, p0, p1, p2, H.vox_bbu(p0.size // 3, pmin, pmax))
    return flags


@pytest.mark.parametrize("direction", [2, 0, 1], ids=["z-rays", "x-rays", "y-rays"])
@pytest.mark.parametrize("preset", [False, True], ids=["empty", "preset"])
def test_oracle_voxelizer_equals_the_reference_text(oracle_lib, direction, preset):
    O = oracle_lib
    if not O.ref_available(O.FP16S, "luw"):
        pytest.skip("oracle/_ref is built only where /root/reference exists")
    p = O.make_params(*H.VOX_SHAPE, O.FP16S, O.FEATURE_SETS["luw"], **H.ZONES)
    got = _run(O.Oracle().bind(p).voxelize_mesh, direction, preset=preset)
    want = _run(O.Reference(O.FP16S, "luw").bind(p).voxelize_mesh, direction, preset=preset)
    assert np.array_equal(got, want)
    solid = int(((got & 3) == 1).sum())
    assert 500 < solid < got.size // 2


def test_oracle_voxelizer_in_a_decomposed_block(oracle_lib):
    """Block (1,0,1) of a 2x1x2 split: the ray origin and the column range follow the domain offset (def_Ox.., FX/kernel.cpp:2391-2393,2433)."""
    O = oracle_lib
    if not O.ref_available(O.FP16S, "luw"):
        pytest.skip("oracle/_ref is built only where /root/reference exists")
    Nx, Ny, Nz = H.VOX_SHAPE
    shape, D, Ov = (Nx // 2 + 2, Ny, Nz // 2 + 2), (2, 1, 2), (Nx // 2 - 1, 0, Nz // 2 - 1)
    p = O.make_params(*shape, O.FP16S, O.FEATURE_SETS["luw"], D=D, O=Ov, **H.ZONES)
    got = _run(O.Oracle().bind(p).voxelize_mesh, 2, shape=shape)
    want = _run(O.Reference(O.FP16S, "luw").bind(p).voxelize_mesh, 2, shape=shape)
    assert np.array_equal(got, want) and ((got & 3) == 1).sum() > 50


@pytest.mark.gpu
@pytest.mark.parametrize("direction", [2, 0, 1], ids=["z-rays", "x-rays", "y-rays"])
@pytest.mark.parametrize("preset", [False, True], ids=["empty", "preset"])
@pytest.mark.parametrize("bins", ["1", "0"], ids=["binned", "all-triangles"])
def test_cuda_voxelizer_equals_oracle(oracle_lib, direction, preset, bins, monkeypatch):
    """Both device voxelisers: the bin-grid kernel luw_voxelize_mesh takes by default (csrc/vox_bins.h) and the all-triangles kernel behind LUW_VOXELIZE_BINS=0."""
    monkeypatch.setenv("LUW_VOXELIZE_BINS", bins)
    from latticeurbanwind_b200 import _cabi as A
    from latticeurbanwind_b200.domain import Domain
    O = oracle_lib
    p = O.make_params(*H.VOX_SHAPE, O.FP16S, O.FEATURE_SETS["luw"], **H.ZONES)
    want = _run(O.Oracle().bind(p).voxelize_mesh, direction, preset=preset)

    def cuda(direction, u, flags, flag, p0, p1, p2, bbu):
        with Domain(*H.VOX_SHAPE, precision=A.FP16S, features=H.FEATURE_SETS["luw"], w=1.0, arith=A.ARITH_FAST, **H.ZONES) as d:
            d.flags[:], d.u[:] = flags, u
            d.write_to_device(A.FIELD_FLAGS); d.write_to_device(A.FIELD_U)
            d.voxelize_mesh(direction, flag, p0, p1, p2, bbu)
            d.read_from_device(A.FIELD_FLAGS); d.finish_queue()
            flags[:] = d.flags
    got = _run(cuda, direction, preset=preset)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------- multi-domain driver: per-domain triangle culling (FX/lbm.cpp:1457-1495)
def _cull(p0, p1, p2, direction, Ov, shape):
    import ctypes as C
    import os
    from tests import test_cpp_host as CPP
    CPP.build()
    L = C.CDLL(os.path.join(CPP.LIB, "libluw_host.so"))
    fn = L.luw_host_cull_triangles
    fn.restype = C.c_uint32
    fn.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_uint32] + [C.c_int] * 3 + [C.c_uint32] * 3 + [C.c_void_p]
    n = p0.size // 3
    ids = np.zeros(n, np.uint32)
    k = fn(p0.ctypes.data, p1.ctypes.data, p2.ctypes.data, n, direction, *Ov, *shape, ids.ctypes.data)
    return ids[:k]


@pytest.mark.parametrize("D", [(2, 2, 1), (2, 1, 2), (4, 2, 1)], ids=["2x2x1", "2x1x2", "4x2x1"])
def test_culled_triangle_subsets_give_the_flags_of_the_whole_mesh(oracle_lib, D):
    """LBM::voxelize_triangles_on_device hands every domain only the triangles whose projected bounding box overlaps it (the reference's culling). For every
    domain with a non-empty subset the voxeliser must produce the flags of the whole mesh; the subsets are smaller than the mesh and keep its order."""
    O = oracle_lib
    from latticeurbanwind_b200.lbm import split
    p0, p1, p2, pmin, pmax = H.vox_mesh()
    ntri = p0.size // 3
    Ng, Nl, doms = split(H.VOX_SHAPE, D)
    smaller = 0
    for d, Ov in doms:
        ids = _cull(p0, p1, p2, 2, Ov, Nl)
        assert np.all(np.diff(ids.astype(np.int64)) > 0) and ids.size <= ntri
        if ids.size == 0:
            continue  # the reference skips the pass for such a domain (FX/lbm.cpp:499)
        smaller += ids.size < ntri
        p = O.make_params(*Nl, O.FP16S, O.FEATURE_SETS["luw"], D=D, O=Ov, **H.ZONES)
        orc = O.Oracle().bind(p)
        N = int(np.prod(Nl))
        out = []
        for sel in (np.arange(ntri), ids):
            q0, q1, q2 = (np.ascontiguousarray(a.reshape(-1, 3)[sel]).reshape(-1) for a in (p0, p1, p2))
            flags, u = np.zeros(N, np.uint8), np.zeros(3 * N, np.float32)
            orc.voxelize_mesh(2, u, flags, TYPE_S, q0, q1, q2, H.vox_bbu(sel.size, pmin, pmax))  # bounding box of the WHOLE mesh either way (FX/lbm.cpp:497)
            out.append(flags)
        assert np.array_equal(out[0], out[1]), d
    assert smaller >= 1


@pytest.mark.gpu
def test_cpp_lbm_voxelises_decomposed_like_single_domain(oracle_lib, tmp_path):
    """LBM::voxelize_triangles_on_device through the C++ host layer (luw_host_case, LUW_CASE_TRIANGLES): the 2x2x1 and 2x1x2 decompositions (culled subsets per
    domain) produce the flags of the single domain, which are the oracle voxeliser's; the flow that follows is identical too."""
    import os
    import subprocess
    from latticeurbanwind_b200 import cases
    from tests import test_cpp_host as CPP
    O = oracle_lib
    CPP.build()
    shape = H.VOX_SHAPE
    N = int(np.prod(shape))
    p0, p1, p2, pmin, pmax = H.vox_mesh()
    tri = str(tmp_path / "tri.bin")
    with open(tri, "wb") as fh:
        fh.write(np.uint32(p0.size // 3).tobytes()); fh.write(pmin.astype(np.float32).tobytes()); fh.write(pmax.astype(np.float32).tobytes())
        fh.write(p0.tobytes()); fh.write(p1.tobytes()); fh.write(p2.tobytes())
    flags, rho, u = cases.periodic_box(*shape)
    flags = np.zeros(N, np.uint8)
    feat = H.FEATURE_SETS["core"]
    res = {}
    for D in ((1, 1, 1), (2, 2, 1), (2, 1, 2)):
        inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        with open(inp, "wb") as fh:
            fh.write(flags.tobytes()); fh.write(rho.tobytes()); fh.write(u.tobytes())
        z = CPP.ZONES
        args = [CPP.DRIVER, *map(str, shape), *map(str, D), "1", str(feat), "0", repr(1e-3), "4", str(z["downstream_face"]), str(z["buffer_N"]), repr(z["buffer_inv_tau"]),
                str(z["buffer_nudge_vertical"]), str(z["sponge_N"]), repr(z["sponge_inv_tau"]), *[repr(float(v)) for v in H.FORCE], *[repr(float(v)) for v in H.OMEGA], inp, out]
        r = subprocess.run(args, capture_output=True, text=True, env=dict(os.environ, LUW_CASE_TRIANGLES=tri))
        assert r.returncode == 0, r.stderr
        raw = np.fromfile(out, np.uint8)
        assert raw.size == 16 * N + N
        res[D] = (raw[:16 * N].view(np.float32).copy(), raw[16 * N:].copy())
    p = O.make_params(*shape, O.FP16S, O.FEATURE_SETS["core"])
    want, uu = np.zeros(N, np.uint8), np.zeros(3 * N, np.float32)
    O.Oracle().bind(p).voxelize_mesh(2, uu, want, TYPE_S, p0, p1, p2, H.vox_bbu(p0.size // 3, pmin, pmax))
    assert np.array_equal(res[(1, 1, 1)][1], want)
    for D in ((2, 2, 1), (2, 1, 2)):
        assert np.array_equal(res[D][1], want), D
        assert np.array_equal(res[D][0], res[(1, 1, 1)][0]), D
