"""GPU: the reference's own, unmodified case driver (FX/setup.cpp + main.cpp + info.cpp + interpolation*.cpp + fluxcorrection.cpp, built by
baseline/build_reference_driver.py against this repo's host layer and CUDA library) runs the reference's own example project
(examples/example_ProfileResearch_noDEM, BASELINE configs[0]: deck-driven profile inflow, STL voxelisation, von Karman inlet, nudging, sponge, FP16C DDFs)
on the B200 and writes its VTK results. This is the drop-in claim end to end; numerical parity of every kernel it launches is covered by the other GPU tests."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "baseline", "_ref", "luw_reference_driver")
CASE = os.path.join(ROOT, "baseline", "_ref", "case_profile")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.isfile(DRIVER) and os.path.isdir(CASE)), reason="baseline/_ref was not built (needs /root/reference at build time)")]


def read_vtk(path):
    raw = open(path, "rb").read()
    head, _, data = raw.partition(b"LOOKUP_TABLE default\n")
    dims = tuple(int(v) for v in re.search(rb"DIMENSIONS (\d+) (\d+) (\d+)", head).groups())
    comps = int(re.search(rb"SCALARS data float (\d+)", head).group(1))
    a = np.frombuffer(data, dtype=">f4")
    assert a.size == dims[0] * dims[1] * dims[2] * comps
    return dims, a.reshape(dims[2], dims[1], dims[0], comps)


def test_reference_case_driver_runs_its_example_deck(tmp_path):
    case = str(tmp_path / "case")
    shutil.copytree(CASE, case)
    r = subprocess.run([DRIVER, os.path.join(case, "conf.luwpf")], capture_output=True, text=True, timeout=600, cwd=case)
    log = r.stdout[-6000:] + r.stderr[-2000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    open(os.path.join(out_dir, "reference_driver.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, log
    assert "Grid Resolution" in r.stdout and "253" in r.stdout
    vtks = [os.path.join(b, f) for b, _, fs in os.walk(case) for f in fs if f.endswith(".vtk")]
    assert vtks, log
    fields = [p for p in vtks if re.search(r"(^|[_/-])u[-_.]", os.path.basename(p)) or "_u" in os.path.basename(p)]
    checked = 0
    for p in vtks:
        dims, a = read_vtk(p)
        assert dims[0] == 253 and dims[1] == 250
        assert np.isfinite(a).all(), p
        if a.shape[3] == 3:  # a velocity field in SI units: the inflow profile tops out at 7.8 m/s
            speed = np.sqrt((a.astype(np.float64) ** 2).sum(axis=3))
            assert 1.0 < float(speed.max()) < 40.0, (p, float(speed.max()))
            checked += 1
    assert checked >= 1, [os.path.basename(p) for p in vtks]


def test_deck_driven_path_equals_oracle(tmp_path, oracle_lib):
    """Parity of the deck-driven path: the same binary, STRICT arithmetic, von Karman inlet off (its per-step update is covered by test_vk_inlet_matches_oracle),
    20 steps. The host layer's LUW_DUMP_DIR hook writes what the case driver handed it -- the triangles of its voxelisation call, the host images at initialize(),
    the kernel constants -- and rho / u after step 20. (a) the flags after voxelising CaseE_PF.stl equal the oracle voxeliser's on the same triangles, bit for bit;
    (b) rho / u after 20 steps equal the oracle's from the same initial images, bit for bit; (c) the 253-wide lattice ran on the TMA tile kernel."""
    O = oracle_lib
    case, dump = str(tmp_path / "case"), str(tmp_path / "dump")
    shutil.copytree(CASE, case)
    os.makedirs(dump)
    deck = open(os.path.join(case, "conf.luwpf")).read()
    deck = re.sub(r"(?m)^run_nstep\s*=.*$", "run_nstep = 20", deck) + "\nvk_inlet_enable = false\n"
    open(os.path.join(case, "conf.luwpf"), "w").write(deck)
    env = dict(os.environ, LUW_ARITH="0", LUW_DUMP_DIR=dump, LUW_DUMP_STEP="20", LUW_VERBOSE="1")
    r = subprocess.run([DRIVER, os.path.join(case, "conf.luwpf")], capture_output=True, text=True, timeout=600, cwd=case, env=env, stdin=subprocess.DEVNULL)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "tile kernel" in r.stderr, "the example deck's 253 x 250 x 59 lattice did not reach the TMA tile kernel:\n" + r.stderr[-2000:]
    prm = dict(line.split() for line in open(os.path.join(dump, "params.txt")))
    shape = (int(prm["Nx"]), int(prm["Ny"]), int(prm["Nz"]))
    N = shape[0] * shape[1] * shape[2]
    precision, features = int(prm["precision"]), int(prm["features"])
    assert shape[:2] == (253, 250) and int(prm["arith"]) == 0
    zones = dict(downstream_face=int(prm["downstream_face"]), buffer_N=int(prm["buffer_N"]), buffer_inv_tau=float(prm["buffer_inv_tau"]),
                 buffer_nudge_vertical=int(prm["buffer_nudge_vertical"]), sponge_N=int(prm["sponge_N"]), sponge_inv_tau=float(prm["sponge_inv_tau"]))
    p = O.make_params(*shape, precision, features, w=float(prm["w"]), **zones)
    rd = lambda name, dt: np.fromfile(os.path.join(dump, name + ".bin"), dt)
    # (a) voxelisation
    ntri, direction, flag, _ = (int(v) for v in rd("vox_head", np.uint32))
    tri = rd("vox_tri", np.float32)
    bbu, p0, p1, p2 = tri[:16].copy(), tri[16:16 + 3 * ntri].copy(), tri[16 + 3 * ntri:16 + 6 * ntri].copy(), tri[16 + 6 * ntri:16 + 9 * ntri].copy()
    assert ntri > 1000 and tri.size == 16 + 9 * ntri
    flags, uu = rd("vox_flags_before", np.uint8), rd("vox_u_before", np.float32)
    O.Oracle().bind(p).voxelize_mesh(direction, uu, flags, flag, p0, p1, p2, bbu)
    got = rd("vox_flags_after", np.uint8)
    assert np.array_equal(got, flags), f"voxelised flags differ from the oracle in {int((got != flags).sum())} cells"
    assert int(((got & 3) == 1).sum()) > 10000
    # (b) 20 steps from the driver's initial images
    f = (float(prm["fx"]), float(prm["fy"]), float(prm["fz"]))
    omega = (float(prm["omega_x"]), float(prm["omega_y"]), float(prm["omega_z"]))
    ref = H.run_cpu(O.Oracle(), O, shape, precision, features, rd("init_flags", np.uint8), rd("init_rho", np.float32), rd("init_u", np.float32), 20,
                    float(prm["w"]), f=f, omega=omega, zones=zones, update_at_end=not (features & 1))
    rho, u = rd("step_rho", np.float32), rd("step_u", np.float32)
    assert rho.size == N and u.size == 3 * N
    assert np.array_equal(rho, ref[1]), f"rho differs in {int((rho != ref[1]).sum())} cells"
    assert np.array_equal(u, ref[2]), f"u differs in {int((u != ref[2]).sum())} values"


DRIVER_T = os.path.join(ROOT, "baseline", "_ref", "luw_reference_driver_T")


@pytest.mark.skipif(not os.path.isfile(DRIVER_T), reason="baseline/_ref/luw_reference_driver_T was not built")
def test_driver_with_temperature_on_runs_the_shipped_switches(tmp_path):
    """LUW ships FP16C + UPDATE_FIELDS + TEMPERATURE (FX/defines.hpp:14-24). The same unmodified case driver built with TEMPERATURE left on creates LUW_TEMPERATURE
    domains (gi, T, alpha / beta through the LBM constructor) and runs the two-kernel thermal step (TMA-tiled momentum kernel, FEAT = 15 | TEMPERATURE, + k_thermal_g).
    Every LUW mode builds its LBM with f = 0, so the temperature cannot act on the flow: the u / rho VTK files must equal the flow-only binary's byte for byte."""
    outs = {}
    for name, exe in (("flow", DRIVER), ("thermal", DRIVER_T)):
        case = str(tmp_path / name)
        shutil.copytree(CASE, case)
        r = subprocess.run([exe, os.path.join(case, "conf.luwpf")], capture_output=True, text=True, timeout=600, cwd=case, stdin=subprocess.DEVNULL, env=dict(os.environ, LUW_VERBOSE="1"))
        assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
        assert "tile kernel" in r.stderr, r.stderr[-2000:]
        if name == "thermal":
            assert "feat=79" in r.stderr, "the TEMPERATURE build did not run the thermal momentum kernel (FEAT = 15 | 64):\n" + r.stderr[-2000:]
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            open(os.path.join(ROOT, "gpurun_out", "reference_driver_T.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
        outs[name] = {os.path.basename(p): open(p, "rb").read() for p in (os.path.join(b, f) for b, _, fs in os.walk(case) for f in fs if f.endswith(".vtk"))}
    assert outs["flow"] and sorted(outs["flow"]) == sorted(k for k in outs["thermal"] if k in outs["flow"])
    for k, v in outs["flow"].items():
        assert outs["thermal"][k] == v, f"{k} differs between the TEMPERATURE-on and the flow-only driver"


FLUX_PARITY = os.path.join(ROOT, "baseline", "_ref", "luw_flux_parity")


@pytest.mark.skipif(not os.path.isfile(FLUX_PARITY), reason="baseline/_ref/luw_flux_parity was not built")
def test_surface_flux_correction_equals_the_reference_function():
    """SURVEY 8-f2, first piece: apply_flux_correction in O(surface) host time (latticeurbanwind_b200/host/fluxcorrection_surface.cpp, linked into the drop-in driver in place of
    FX/fluxcorrection.cpp) against the reference's own, unmodified function compiled under another name: three lattices x five downstream settings x with / without the
    inflow re-evaluation -> flags, u and the reported sums bit-identical (baseline/flux_parity.cpp)."""
    r = subprocess.run([FLUX_PARITY], capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert "30 of 30 runs identical" in r.stdout, r.stdout[-3000:]


INLET_PARITY = os.path.join(ROOT, "baseline", "_ref", "luw_inlet_parity")


@pytest.mark.skipif(not os.path.isfile(INLET_PARITY), reason="baseline/_ref/luw_inlet_parity was not built")
def test_surface_inlet_outlet_mapping_equals_the_reference_functions():
    """apply_inlet_outlet / apply_inlet_outlet_hd with the face cells enumerated directly and the sample searches on the GPU (csrc/lbm_inlet.cuh through
    luw_inlet_nearest / luw_inlet_knn; latticeurbanwind_b200/host/inlet_outlet_surface.cpp, linked into the drop-in driver under the reference's names) against the
    reference's own, unmodified functions compiled under other names: 20 position-level runs (5 sample clouds x threshold x nearest / K = 64 fit) and 60 lattice-level
    runs (3 lattices x 5 downstream settings x open / closed x nearest / HD, with and without the side z cap) -> every flag and velocity bit-identical
    (baseline/inlet_parity.cpp)."""
    r = subprocess.run([INLET_PARITY], capture_output=True, text=True, timeout=900, stdin=subprocess.DEVNULL)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-1000:]
    assert "80 of 80 runs identical" in r.stdout, r.stdout[-4000:]
    launches = int(r.stdout.strip().rsplit("search kernels launched:", 1)[1])
    assert launches >= 80, "the sample searches did not run on the device"


CASE_NWP = os.path.join(ROOT, "baseline", "_ref", "case_nwp")


@pytest.mark.skipif(not os.path.isdir(CASE_NWP), reason="baseline/_ref/case_nwp was not staged")
@pytest.mark.parametrize("deck,marker", [("conf.luw", "K-nearest search on the GPU"), ("conf_nearest.luw", "sample search on the GPU")], ids=["high-order", "nearest"])
def test_wrf_style_deck_maps_its_boundary_like_the_reference_functions(tmp_path, deck, marker):
    """SURVEY 8-f2 in place: a .luw deck (boundary values from proj_temp/SurfData_<datetime>.csv: 5 906 synthetic samples on the five open faces, the example's building
    mesh) through the reference's unmodified case driver, once with this repo's apply_inlet_outlet(_hd) (face cells enumerated, sample searches on the GPU) and once with
    LUW_INLET_AB=reference, which makes the same binary call the reference's own functions instead. The host images the driver hands to LBM::initialize -- flags, u, rho
    after boundary mapping AND flux correction -- must be byte-identical, and both runs must finish their 20 steps."""
    images = {}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for mode in ("ours", "reference"):
        case, dump = str(tmp_path / mode), str(tmp_path / (mode + "_dump"))
        shutil.copytree(CASE_NWP, case)
        os.makedirs(dump)
        env = dict(os.environ, LUW_DUMP_DIR=dump, LUW_DUMP_STEP="20", LUW_VERBOSE="1")
        if mode == "reference":
            env["LUW_INLET_AB"] = "reference"
        r = subprocess.run([DRIVER, os.path.join(case, deck)], capture_output=True, text=True, timeout=900, cwd=case, env=env, stdin=subprocess.DEVNULL)
        open(os.path.join(ROOT, "gpurun_out", f"reference_driver_nwp_{deck.split('.')[0]}_{mode}.log"), "w").write(r.stdout + "\n---- stderr ----\n" + r.stderr)
        assert r.returncode == 0, r.stdout[-5000:] + r.stderr[-2000:]
        assert (marker in r.stdout) == (mode == "ours") and ("LUW_INLET_AB=reference" in r.stdout) == (mode == "reference"), r.stdout[-5000:]
        images[mode] = {name: open(os.path.join(dump, name + ".bin"), "rb").read() for name in ("init_flags", "init_u", "init_rho", "step_u")}
    flags = np.frombuffer(images["ours"]["init_flags"], np.uint8)
    u = np.frombuffer(images["ours"]["init_u"], np.float32)
    assert int(((flags & 3) == 2).sum()) > 50000 and float(np.abs(u).max()) > 0.01, "the boundary was not mapped"
    for name in ("init_flags", "init_u", "init_rho"):
        assert images["ours"][name] == images["reference"][name], f"{name} differs between this repo's boundary mapping and the reference's"
    assert images["ours"]["step_u"] == images["reference"]["step_u"]
