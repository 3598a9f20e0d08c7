"""Shared test scripts: run the same LBM scenario on the CPU oracle / the reference kernel text and on the CUDA path."""
import numpy as np

from latticeurbanwind_b200 import cases

FEATURE_SETS = {"bench": 0, "chan": 4, "plain": 1 | 4, "core": 1 | 2 | 4 | 8, "luw": 1 | 2 | 4 | 8 | 16 | 32, "luwnf": 2 | 4 | 8 | 16 | 32,
                "luwT": 1 | 2 | 4 | 8 | 16 | 32 | 64, "chanT": 2 | 4 | 64}
FORCE = (1e-6, 0.0, -2e-6)
OMEGA = (0.0, 5.6e-6, 4.7e-6)
ZONES = dict(downstream_face=2, buffer_N=6, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=8, sponge_inv_tau=0.02)


def small_urban(Nx=48, Ny=40, Nz=32, seed=1234):
    return cases.urban(Nx, Ny, Nz, seed=seed, edge=6, pitch=12)


def run_cpu(engine, O, shape, precision, features, flags, rho, u, steps, w, f=FORCE, omega=OMEGA, zones=ZONES, update_at_end=False,
            D=(1, 1, 1), Ov=(0, 0, 0)):
    """engine: oracle.Oracle() or oracle.Reference(...). Returns (fi, rho, u) after `steps` stream_collide calls."""
    Nx, Ny, Nz = shape
    p = O.make_params(Nx, Ny, Nz, precision, features, w=w, D=D, O=Ov, **zones)
    fi = np.zeros(19 * p.N, O.ddf_dtype(precision))
    flags, rho, u = flags.copy(), rho.copy(), u.copy()
    engine.bind(p)
    engine.initialize(fi, rho, u, flags)
    for t in range(steps):
        engine.stream_collide(fi, rho, u, flags, t, f, omega)
    if update_at_end:
        engine.update_fields(fi, rho, u, flags, steps, f, omega)
    return fi, rho, u


def run_cuda(shape, precision, features, flags, rho, u, steps, w, arith, f=FORCE, omega=OMEGA, zones=ZONES, update_at_end=False, batched=False,
             D=(1, 1, 1), O=(0, 0, 0), expect_tiles=None):
    from latticeurbanwind_b200.domain import Domain
    Nx, Ny, Nz = shape
    with Domain(Nx, Ny, Nz, D=D, O=O, precision=precision, features=features, w=w, arith=arith, **zones) as d:
        if expect_tiles is not None:
            assert d.uses_tiles() == expect_tiles, "unexpected stream_collide implementation"
        d.rho[:], d.u[:], d.flags[:] = rho, u, flags
        d.f, d.omega = f, omega
        d.upload_all()
        d.t = 1
        d.enqueue_initialize()
        d.t = 0
        if batched:
            d.run_steps(steps)
        else:
            for _ in range(steps):
                d.enqueue_stream_collide()
                d.increment_time_step()
        if update_at_end:
            d.enqueue_update_fields()
        d.download_all()
        return d.read_fi(), d.rho.copy(), d.u.copy()


def decode(O, engine_or_none, fi, precision):
    """DDF image -> float32 values (so that +0/-0 encodings compare equal)."""
    if precision == O.FP32:
        return fi
    orc = O.Oracle()
    table = np.array([(orc.half_to_float(h) if precision == O.FP16S else orc.fp16c_to_float(h)) for h in range(65536)], np.float32)
    return table[fi]


def report(name, **metrics):
    """Measured parity errors: printed (pytest -s / -rA) and appended to gpurun_out/parity_measured.jsonl so that the margin against every written
    tolerance is on record (profiles/ keeps the copy of the last GPU run)."""
    import json
    import os
    line = json.dumps({"test": name, **{k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in metrics.items()}})
    print("[parity]", line)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_measured.jsonl"), "a") as fh:
            fh.write(line + "\n")
    except OSError:
        pass


def errors(got, ref):
    """(rel-L2(u), max-abs(u), rel-L2(rho)) of (fi, rho, u) triples."""
    return rel_l2(got[2], ref[2]), float(np.abs(got[2] - ref[2]).max()), rel_l2(got[1], ref[1])


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


# ---------------------------------------------------------------------------------------------------------------- golden fixtures
GOLDEN_SHAPE, GOLDEN_STEPS = (24, 20, 16), 10
TINY_SHAPE, TINY_STEPS = (12, 10, 8), 6
HALO_SHAPE = (14, 12, 10)  # local lattice of one block of a 2x2x2 decomposition (halo layer on every axis)
VK_SHAPE = (16, 12, 10)


def golden_case(shape):
    return cases.urban(*shape, seed=4321, edge=3, pitch=6)


def golden_run(engine, O, precision, fset, shape=GOLDEN_SHAPE, steps=GOLDEN_STEPS):
    flags, rho, u = golden_case(shape)
    zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
    return run_cpu(engine, O, shape, precision, O.FEATURE_SETS[fset], flags, rho, u, steps, cases.relaxation_rate(1e-6), zones=zones)


def halo_params(O, precision, features=None):
    zones = dict(downstream_face=2, buffer_N=4, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=5, sponge_inv_tau=0.02)
    feat = O.FEATURE_SETS["luw"] if features is None else features
    return O.make_params(*HALO_SHAPE, precision, feat, w=cases.relaxation_rate(1e-6), D=(2, 2, 2), O=(-1, -1, -1), **zones)


def cut_block(shape_global, D, d, flags, rho, u):
    """Host images of block d=(dx,dy,dz) of a D-decomposition incl. its halo layers, periodic at the global edges
    (the index stitching of FX/lbm.hpp:274-297 / FX/lbm.cpp:1057-1073). Returns (local shape, offset O, flags, rho, u)."""
    Ng = np.array(shape_global)
    D = np.array(D)
    H_ = (D > 1).astype(int)
    Nl = Ng // D + 2 * H_
    Ov = np.array(d) * (Ng // D) - H_
    idx = [np.mod(np.arange(Nl[a]) + Ov[a], Ng[a]) for a in range(3)]
    take = lambda a: a.reshape(Ng[2], Ng[1], Ng[0])[np.ix_(idx[2], idx[1], idx[0])].reshape(-1)
    Ncell = int(np.prod(Ng))
    ul = np.concatenate([take(u[c * Ncell:(c + 1) * Ncell]) for c in range(3)])
    return tuple(int(v) for v in Nl), tuple(int(v) for v in Ov), take(flags).copy(), take(rho).copy(), ul


def golden_halo(engine, O, precision):
    """Block (0,0,0) of a 2x2x2 decomposition of the GOLDEN_SHAPE case: 3 steps, and every halo payload the step path moves, at both
    slot parities (exchanged with itself, which is what a periodic 1-block-per-axis neighbourhood would do)."""
    flags, rho, u = golden_case(GOLDEN_SHAPE)
    shape, Ov, flags, rho, u = cut_block(GOLDEN_SHAPE, (2, 2, 2), (0, 0, 0), flags, rho, u)
    assert shape == HALO_SHAPE and Ov == (-1, -1, -1)
    p = halo_params(O, precision)
    fi = np.zeros(19 * p.N, O.ddf_dtype(precision))
    engine.bind(p)
    engine.initialize(fi, rho, u, flags)
    out = {}
    name = O.PREC_NAME[precision]
    for t in range(3):
        engine.stream_collide(fi, rho, u, flags, t, FORCE, OMEGA)
        for axis in range(3):
            A = (p.Ny * p.Nz, p.Nz * p.Nx, p.Nx * p.Ny)[axis]
            bp, bm = np.zeros(5 * A, fi.dtype), np.zeros(5 * A, fi.dtype)
            engine.extract_fi(axis, t, bp, bm, fi)
            out[f"halo_{name}_t{t}_axis{axis}_p"], out[f"halo_{name}_t{t}_axis{axis}_m"] = bp.copy(), bm.copy()
            engine.insert_fi(axis, t, bm, bp, fi)
        out[f"halo_{name}_t{t}_fi"] = fi.copy()
    for axis in range(3):
        A = (p.Ny * p.Nz, p.Nz * p.Nx, p.Nx * p.Ny)[axis]
        bp, bm = np.zeros(17 * A, np.uint8), np.zeros(17 * A, np.uint8)
        engine.extract_rho_u_flags(axis, bp, bm, rho, u, flags)
        out[f"halo_{name}_ruf_axis{axis}_p"], out[f"halo_{name}_ruf_axis{axis}_m"] = bp.copy(), bm.copy()
        engine.insert_rho_u_flags(axis, bm, bp, rho, u, flags)
    out[f"halo_{name}_ruf_rho"], out[f"halo_{name}_ruf_u"], out[f"halo_{name}_ruf_flags"] = rho.copy(), u.copy(), flags.copy()
    return out


def codec_sweep():
    """Floats covering the FP16C range (+-2), its subnormals, ties, and out-of-range values."""
    rng = np.random.default_rng(99)
    parts = [np.float32(s) * np.exp2(rng.uniform(-30, 2, 20000)).astype(np.float32) for s in (1.0, -1.0)]
    bits = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    bits = bits[np.isfinite(bits)]
    edge = np.array([0.0, -0.0, 1.0, -1.0, 1.9990234, 2.0, 3.0, 6.1035156e-5, 3.0517578e-5, 2.9802322e-8, 5.9604645e-8, 1e-9], np.float32)
    return np.concatenate(parts + [bits, edge]).astype(np.float32)


def vk_case(seed=7):
    """Packed inlet buffers like VonKarmanInletUpdater's (FX/setup.cpp:886-1116): west-face points of a VK_SHAPE lattice, 2 faces x 24 modes."""
    Nx, Ny, Nz = VK_SHAPE
    N = Nx * Ny * Nz
    rng = np.random.default_rng(seed)
    ys, zs = np.meshgrid(np.arange(Ny), np.arange(1, Nz), indexing="ij")
    cells_w = (0 + (ys + zs * Ny) * Nx).reshape(-1)
    cells_s = (np.arange(1, Nx)[:, None] + (0 + np.arange(1, Nz)[None, :] * Ny) * Nx).reshape(-1)
    pc = np.concatenate([cells_w, cells_s]).astype(np.uint64)
    P = pc.size
    pf = np.concatenate([np.zeros(cells_w.size, np.uint8), np.full(cells_s.size, 2, np.uint8)])
    pf[::17] = 7  # "no synthesis" marker: base velocity only
    pd = np.zeros(7 * P, np.float32)
    pd[0 * P:1 * P] = (pc % Nx).astype(np.float32)
    pd[1 * P:2 * P] = ((pc // Nx) % Ny).astype(np.float32)
    pd[2 * P:3 * P] = (pc // (Nx * Ny)).astype(np.float32)
    pd[3 * P:4 * P] = 0.08 + 0.01 * rng.random(P, dtype=np.float32)
    pd[4 * P:5 * P] = 0.01 * (rng.random(P, dtype=np.float32) - 0.5)
    pd[6 * P:7 * P] = 0.004 * rng.random(P, dtype=np.float32)
    pd[6 * P + 3:7 * P:11] = 0.0  # sigma == 0 -> base velocity only
    M, faces = 24, 5
    V = faces * M
    md = np.zeros(10 * V, np.float32)
    md[0:3 * V] = rng.normal(0, 0.3, 3 * V).astype(np.float32)
    md[3 * V:4 * V] = rng.normal(0, 0.05, V).astype(np.float32)
    md[4 * V:7 * V] = rng.normal(0, 1.0, 3 * V).astype(np.float32)
    md[7 * V:10 * V] = rng.uniform(0, 2 * np.pi, 3 * V).astype(np.float32)
    return pc, pf, pd, md, M, V, N


# ---------------------------------------------------------------------------------------------- running statistics (FX/setup.cpp:4441-4488)
STATS_N, STATS_SAMPLES = 5000, 6


def stats_samples():
    """Seeded sequence of (rho, u) samples, magnitudes like a turbulent LBM field (u ~ 0.05 +- 0.02, rho ~ 1 +- 1e-3)."""
    rng = np.random.default_rng(2024)
    out = []
    for _ in range(STATS_SAMPLES):
        rho = (1.0 + 1e-3 * rng.standard_normal(STATS_N)).astype(np.float32)
        u = (np.array([0.05, 0.0, 0.01], np.float32)[:, None] + 0.02 * rng.standard_normal((3, STATS_N)).astype(np.float32)).reshape(-1).astype(np.float32)
        out.append((rho, u))
    return out


def stats_run(engine):
    """engine: oracle.OracleStats() / oracle.RefStats(). Returns (u_avg interleaved [3n+c], rho_avg, m2_u, m2_v, m2_w) after all samples."""
    N = STATS_N
    u_avg, rho_avg = np.zeros(3 * N, np.float32), np.zeros(N, np.float32)
    m2 = [np.zeros(N, np.float32) for _ in range(3)]
    for rho, u in stats_samples():
        engine.accumulate(rho, u, u_avg, rho_avg, *m2)
    return u_avg, rho_avg, m2[0], m2[1], m2[2]


# ---------------------------------------------------------------------------------------------- mesh voxelisation (FX/kernel.cpp:2381-2471)
VOX_SHAPE = (40, 36, 28)


def _box_tris(lo, hi):
    x0, y0, z0 = lo; x1, y1, z1 = hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float32)
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3)]
    tris = []
    for a, b, c, d in quads:
        tris += [(v[a], v[b], v[c]), (v[a], v[c], v[d])]
    return tris


def vox_mesh(seed=3):
    """A small 'city' like luwvox produces: a base slab, extruded boxes at fractional positions, one tilted tetrahedron (rays graze its edges), in lattice
    coordinates of a VOX_SHAPE domain (mesh already translated like FX/setup.cpp:4084-4087). Returns p0, p1, p2 (3 floats per triangle) and pmin, pmax."""
    rng = np.random.default_rng(seed)
    Nx, Ny, Nz = VOX_SHAPE
    tris = _box_tris((1.0, 1.0, 1.0), (Nx - 1.0, Ny - 1.0, 2.3))
    for _ in range(7):
        cx, cy = rng.uniform(6, Nx - 6), rng.uniform(6, Ny - 6)
        w, d, h = rng.uniform(2.2, 6.5), rng.uniform(2.2, 6.5), rng.uniform(4.0, Nz - 6.0)
        tris += _box_tris((cx - w / 2, cy - d / 2, 1.7), (cx + w / 2, cy + d / 2, 1.7 + h))
    t = np.array([[20.2, 8.1, 2.0], [27.9, 9.3, 2.0], [23.5, 15.7, 2.0], [24.1, 11.2, 17.4]], np.float32)
    for a, b, c in [(0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)]:
        tris.append((t[a], t[b], t[c]))
    P = np.array(tris, np.float32)  # [T, 3 vertices, 3]
    p0, p1, p2 = (np.ascontiguousarray(P[:, k, :]).reshape(-1) for k in range(3))
    return p0, p1, p2, P.reshape(-1, 3).min(0), P.reshape(-1, 3).max(0)


def vox_bbu(ntri, pmin, pmax):
    """bounding_box_and_velocity[16] of LBM_Domain::voxelize_mesh_on_device (FX/lbm.cpp:497,529-549): triangle count as float bits, bbox -+ 2 cells, resting geometry."""
    bbu = np.zeros(16, np.float32)
    bbu[:1].view(np.uint32)[0] = ntri
    bbu[1:4] = pmin - np.float32(2.0)
    bbu[4:7] = pmax + np.float32(2.0)
    return bbu


# ---------------------------------------------------------------------------------------------- thermal D3Q7 (FX/kernel.cpp:1306-1336,1639-1684)
THERMAL = dict(w_T=1.0 / (2.0 * 2.0e-3 + 0.5), beta=0.4, T_avg=1.0)  # def_w_T = 1/(2 alpha + 1/2), FX/lbm.cpp:750
THERMAL_SHAPE, THERMAL_STEPS = (20, 18, 14), 8


def thermal_case(shape=THERMAL_SHAPE, seed=77):
    """The urban case plus a temperature field: every TYPE_E cell also carries TYPE_T (what the case driver does for a WRF deck with a T column,
    FX/setup.cpp:5268-5317 -- and what makes the T sponge's read of the top row independent of the work-item order), a few interior fluid cells are
    heated TYPE_T sources, the rest starts from a stratified profile with seeded noise."""
    Nx, Ny, Nz = shape
    flags, rho, u = cases.urban(Nx, Ny, Nz, seed=seed, edge=3, pitch=6)
    flags = flags.copy()
    flags[(flags & 0x03) == 0x02] |= 0x04
    rng = np.random.default_rng(seed)
    z = (np.arange(Nx * Ny * Nz) // (Nx * Ny)).astype(np.float32)
    T = (np.float32(1.0) + np.float32(0.02) * z / np.float32(Nz) + np.float32(1e-3) * rng.standard_normal(Nx * Ny * Nz).astype(np.float32)).astype(np.float32)
    fluid = np.flatnonzero(flags == 0)
    hot = fluid[rng.choice(fluid.size, size=max(4, fluid.size // 200), replace=False)]
    flags[hot] |= 0x04
    T[hot] = np.float32(1.05)
    return flags, rho, u, T


def run_cpu_thermal(engine, O, shape, precision, features, flags, rho, u, T, steps, w, f=FORCE, omega=OMEGA, zones=ZONES, thermal=THERMAL,
                    update_at_end=False, D=(1, 1, 1), Ov=(0, 0, 0)):
    """Like run_cpu with the TEMPERATURE extension: returns (fi, rho, u, gi, T)."""
    Nx, Ny, Nz = shape
    p = O.make_params(Nx, Ny, Nz, precision, features, w=w, D=D, O=Ov, **zones)
    fi, gi = np.zeros(19 * p.N, O.ddf_dtype(precision)), np.zeros(7 * p.N, O.ddf_dtype(precision))
    flags, rho, u, T = flags.copy(), rho.copy(), u.copy(), T.copy()
    engine.bind(p)
    engine.set_thermal(**thermal)
    engine.initialize_thermal(fi, rho, u, flags, gi, T)
    for t in range(steps):
        engine.stream_collide_thermal(fi, rho, u, flags, t, f, omega, gi, T)
    if update_at_end:
        engine.update_fields_thermal(fi, rho, u, flags, steps, f, omega, gi, T)
    return fi, rho, u, gi, T


def golden_thermal_halo(engine, O, precision):
    """Block (0,0,0) of a 2x2x2 decomposition of the thermal case: 3 steps with the gi / T halo payloads at both slot parities (exchanged with itself)."""
    flags, rho, u, T = thermal_case(GOLDEN_SHAPE)
    Ncell = int(np.prod(GOLDEN_SHAPE))
    shape, Ov, fl, rh, ul = cut_block(GOLDEN_SHAPE, (2, 2, 2), (0, 0, 0), flags, rho, u)
    Tl = cut_block(GOLDEN_SHAPE, (2, 2, 2), (0, 0, 0), flags, T, np.zeros(3 * Ncell, np.float32))[3]
    p = halo_params(O, precision, O.FEATURE_SETS["luwT"])
    fi, gi = np.zeros(19 * p.N, O.ddf_dtype(precision)), np.zeros(7 * p.N, O.ddf_dtype(precision))
    engine.bind(p)
    engine.set_thermal(**THERMAL)
    engine.initialize_thermal(fi, rh, ul, fl, gi, Tl)
    out, name = {}, O.PREC_NAME[precision]
    for t in range(3):
        engine.stream_collide_thermal(fi, rh, ul, fl, t, FORCE, OMEGA, gi, Tl)
        for axis in range(3):
            A = (p.Ny * p.Nz, p.Nz * p.Nx, p.Nx * p.Ny)[axis]
            bp, bm = np.zeros(A, gi.dtype), np.zeros(A, gi.dtype)
            engine.extract_gi(axis, t, bp, bm, gi)
            out[f"thalo_{name}_t{t}_axis{axis}_p"], out[f"thalo_{name}_t{t}_axis{axis}_m"] = bp.copy(), bm.copy()
            engine.insert_gi(axis, t, bm, bp, gi)
        out[f"thalo_{name}_t{t}_gi"] = gi.copy()
    for axis in range(3):
        A = (p.Ny * p.Nz, p.Nz * p.Nx, p.Nx * p.Ny)[axis]
        bp, bm = np.zeros(A, np.float32), np.zeros(A, np.float32)
        engine.extract_T(axis, bp, bm, Tl)
        out[f"thalo_{name}_T_axis{axis}_p"], out[f"thalo_{name}_T_axis{axis}_m"] = bp.copy(), bm.copy()
        engine.insert_T(axis, bm, bp, Tl)
    out[f"thalo_{name}_T"] = Tl.copy()
    return out


def run_cuda_thermal(shape, precision, features, flags, rho, u, T, steps, w, arith, f=FORCE, omega=OMEGA, zones=ZONES, thermal=THERMAL, update_at_end=False,
                     batched=False, D=(1, 1, 1), O=(0, 0, 0), expect_tiles=None):
    """run_cpu_thermal's script on the CUDA path (Domain over the C ABI): returns (fi, rho, u, gi, T). A tiled thermal domain runs the two-kernel step (TMA-tiled
    momentum kernel + k_thermal_g) while the buoyancy term vanishes (f = 0 or beta = 0) and the fused one-cell-per-thread kernel otherwise."""
    from latticeurbanwind_b200.domain import Domain
    Nx, Ny, Nz = shape
    with Domain(Nx, Ny, Nz, D=D, O=O, precision=precision, features=features, w=w, arith=arith, **zones) as d:
        assert d.thermal
        if expect_tiles is not None:
            assert d.uses_tiles() == expect_tiles
        d.set_thermal(**thermal)
        d.rho[:], d.u[:], d.flags[:], d.T[:] = rho, u, flags, T
        d.f, d.omega = f, omega
        d.upload_all()
        d.t = 1
        d.enqueue_initialize()
        d.t = 0
        if batched:
            d.run_steps(steps)
        else:
            for _ in range(steps):
                d.enqueue_stream_collide()
                d.increment_time_step()
        if update_at_end:
            d.enqueue_update_fields()
        d.download_all()
        return d.read_fi(), d.rho.copy(), d.u.copy(), d.read_gi(), d.T.copy()
