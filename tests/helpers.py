"""Shared test scripts: run the same LBM scenario on the CPU oracle / the reference kernel text and on the CUDA path."""
import numpy as np

from latticeurbanwind_b200 import cases

FEATURE_SETS = {"bench": 0, "plain": 1 | 4, "core": 1 | 2 | 4 | 8, "luw": 1 | 2 | 4 | 8 | 16 | 32, "luwnf": 2 | 4 | 8 | 16 | 32}
FORCE = (1e-6, 0.0, -2e-6)
OMEGA = (0.0, 5.6e-6, 4.7e-6)
ZONES = dict(downstream_face=2, buffer_N=6, buffer_inv_tau=0.01, buffer_nudge_vertical=1, sponge_N=8, sponge_inv_tau=0.02)


def small_urban(Nx=48, Ny=40, Nz=32, seed=1234):
    return cases.urban(Nx, Ny, Nz, seed=seed, edge=6, pitch=12)


def run_cpu(engine, O, shape, precision, features, flags, rho, u, steps, w, f=FORCE, omega=OMEGA, zones=ZONES, update_at_end=False):
    """engine: oracle.Oracle() or oracle.Reference(...). Returns (fi, rho, u) after `steps` stream_collide calls."""
    Nx, Ny, Nz = shape
    p = O.make_params(Nx, Ny, Nz, precision, features, w=w, **zones)
    fi = np.zeros(19 * p.N, O.ddf_dtype(precision))
    flags, rho, u = flags.copy(), rho.copy(), u.copy()
    engine.bind(p)
    engine.initialize(fi, rho, u, flags)
    for t in range(steps):
        engine.stream_collide(fi, rho, u, flags, t, f, omega)
    if update_at_end:
        engine.update_fields(fi, rho, u, flags, steps, f, omega)
    return fi, rho, u


def run_cuda(shape, precision, features, flags, rho, u, steps, w, arith, f=FORCE, omega=OMEGA, zones=ZONES, update_at_end=False, batched=False):
    from latticeurbanwind_b200.domain import Domain
    Nx, Ny, Nz = shape
    with Domain(Nx, Ny, Nz, precision=precision, features=features, w=w, arith=arith, **zones) as d:
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


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))
