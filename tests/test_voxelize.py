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
def test_cuda_voxelizer_equals_oracle(oracle_lib, direction, preset):
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
