"""Synthetic, seeded case builders for the LBM step (host side, numpy only).

They do for the synthetic configurations of BASELINE.json what the reference's case driver does for a deck
(FX/setup.cpp:4931-6153): write `flags`, `rho`, `u` host images in the reference's SoA layout
(n = x + (y + z*Ny)*Nx; u = [ux | uy | uz]) -- TYPE_S ground / buildings, TYPE_E on the open faces with an inflow profile.

  * periodic_box   -- all fluid, fully periodic: the upstream BENCHMARK protocol (FX/setup_examples.cpp:5-37) with a perturbation
  * channel        -- C2 "empty channel": TYPE_E on x=0 / x=Nx-1, TYPE_S on the y and z walls
  * urban          -- C3 "staggered cube array": ground TYPE_S, cubes, TYPE_E on the 4 sides + top, log-law inflow along +x
"""
import numpy as np

TYPE_S, TYPE_E = np.uint8(0x01), np.uint8(0x02)


def _grid(Nx, Ny, Nz):
    z, y, x = np.meshgrid(np.arange(Nz), np.arange(Ny), np.arange(Nx), indexing="ij")
    return x, y, z


def _perturb(u, seed, amp):
    if amp > 0.0:
        rng = np.random.default_rng(seed)
        u += (amp * (2.0 * rng.random(u.shape, dtype=np.float32) - 1.0)).astype(np.float32)
    return u


def periodic_box(Nx, Ny, Nz, seed=1234, u0=(0.05, 0.0, 0.0), amp=1e-3):
    N = Nx * Ny * Nz
    flags = np.zeros(N, np.uint8)
    rho = np.ones(N, np.float32)
    u = np.empty(3 * N, np.float32)
    for c in range(3):
        u[c * N:(c + 1) * N] = np.float32(u0[c])
    return flags, rho, _perturb(u, seed, amp)


def channel(Nx, Ny, Nz, seed=1234, u0=(0.05, 0.0, 0.0), amp=1e-3):
    flags, rho, u = periodic_box(Nx, Ny, Nz, seed, u0, amp)
    x, y, z = _grid(Nx, Ny, Nz)
    f = flags.reshape(Nz, Ny, Nx)
    f[(x == 0) | (x == Nx - 1)] = TYPE_E
    f[(y == 0) | (y == Ny - 1) | (z == 0) | (z == Nz - 1)] = TYPE_S
    return flags, rho, u


def log_law(z, Nz, u_ref=0.1, z0=0.05):
    """u(z) = u_ref * ln((z+0.5)/z0) / ln((Nz-0.5)/z0), clipped to [0, u_ref]."""
    prof = u_ref * np.log((z + 0.5) / z0) / np.log((Nz - 0.5) / z0)
    return np.clip(prof, 0.0, u_ref).astype(np.float32)


def cube_heights(ix, iy):
    h = (ix.astype(np.uint64) * np.uint64(73856093)) ^ (iy.astype(np.uint64) * np.uint64(19349663))
    return (16 + 8 * (h % np.uint64(5))).astype(np.int64)


def urban(Nx, Ny, Nz, seed=1234, edge=16, pitch=32, footprint=0.75, u_ref=0.1, amp=1e-3):
    """Staggered cube array over the central `footprint` fraction of the xy plane; heights 16+8*(hash%5)."""
    N = Nx * Ny * Nz
    x, y, z = _grid(Nx, Ny, Nz)
    flags = np.zeros((Nz, Ny, Nx), np.uint8)
    x0, x1 = int(Nx * (1 - footprint) / 2), int(Nx * (1 + footprint) / 2)
    y0, y1 = int(Ny * (1 - footprint) / 2), int(Ny * (1 + footprint) / 2)
    iy = (y - y0) // pitch
    xs = x - x0 - (iy % 2) * (pitch // 2)  # alternate rows shifted by half a pitch
    ix = xs // pitch
    inside = (x >= x0) & (x < x1) & (y >= y0) & (y < y1) & (xs >= 0) & (xs % pitch < edge) & ((y - y0) % pitch < edge)
    h = cube_heights(np.maximum(ix, 0), np.maximum(iy, 0))
    flags[inside & (z <= np.minimum(h, Nz - 3))] = TYPE_S
    flags[z == 0] = TYPE_S
    side = (x == 0) | (x == Nx - 1) | (y == 0) | (y == Ny - 1) | (z == Nz - 1)
    flags[side & (z > 0)] = TYPE_E
    rho = np.ones(N, np.float32)
    u = np.zeros(3 * N, np.float32)
    u[:N] = log_law(z.astype(np.float32), Nz, u_ref).reshape(-1)
    u = _perturb(u, seed, amp)
    fl = flags.reshape(-1)
    for c in range(3):  # boundary cells carry the clean profile; solids are at rest
        uc = u[c * N:(c + 1) * N]
        if c == 0:
            uc[fl == TYPE_E] = log_law(z.astype(np.float32), Nz, u_ref).reshape(-1)[fl == TYPE_E]
        else:
            uc[fl == TYPE_E] = 0.0
        uc[fl == TYPE_S] = 0.0
    return fl, rho, u


CASES = {"periodic_box": periodic_box, "channel": channel, "urban": urban}


def block_case(name, Ng, O=(0, 0, 0), Nl=None, seed=1234, amp=1e-3, u_ref=0.1, edge=16, pitch=32, footprint=0.75, out=None):
    """The same three cases, built directly for ONE block of a decomposed lattice (local shape Nl incl. halo layers at global offset O,
    periodic at the global edges) with broadcasting instead of full coordinate grids, so that 512^3 ... 1024x1024x256 blocks are cheap.
    Geometry (flags, base profile) is a function of the GLOBAL coordinate, so neighbouring blocks agree on their shared layers; the
    seeded perturbation of interior fluid cells is per block (seed + block offset). `out` = (flags, rho, u) arrays to fill (e.g. pinned)."""
    Nl = tuple(Ng) if Nl is None else tuple(Nl)
    Nx, Ny, Nz = Nl
    N = Nx * Ny * Nz
    xg, yg, zg = (np.mod(np.arange(Nl[a]) + O[a], Ng[a]) for a in range(3))
    flags, rho, u = out if out is not None else (np.empty(N, np.uint8), np.empty(N, np.float32), np.empty(3 * N, np.float32))
    f3 = flags.reshape(Nz, Ny, Nx)
    f3[:] = 0
    rho[:] = 1.0
    u3 = u.reshape(3, Nz, Ny, Nx)
    X, Y, Z = xg[None, None, :], yg[None, :, None], zg[:, None, None]
    if name == "periodic_box":
        base = np.full(Nz, 0.05, np.float32)
    elif name == "channel":
        f3[np.broadcast_to((X == 0) | (X == Ng[0] - 1), f3.shape)] = TYPE_E
        f3[np.broadcast_to((Y == 0) | (Y == Ng[1] - 1) | (Z == 0) | (Z == Ng[2] - 1), f3.shape)] = TYPE_S
        base = np.full(Nz, 0.05, np.float32)
    elif name == "urban":
        x0, x1 = int(Ng[0] * (1 - footprint) / 2), int(Ng[0] * (1 + footprint) / 2)
        y0, y1 = int(Ng[1] * (1 - footprint) / 2), int(Ng[1] * (1 + footprint) / 2)
        x2, y2 = xg[None, :], yg[:, None]
        iy = (y2 - y0) // pitch
        xs = x2 - x0 - (iy % 2) * (pitch // 2)
        ix = xs // pitch
        inside = (x2 >= x0) & (x2 < x1) & (y2 >= y0) & (y2 < y1) & (xs >= 0) & (xs % pitch < edge) & ((y2 - y0) % pitch < edge)
        h = np.where(inside, np.minimum(cube_heights(np.maximum(ix, 0), np.maximum(iy, 0)), Ng[2] - 3), -1)
        side2 = (x2 == 0) | (x2 == Ng[0] - 1) | (y2 == 0) | (y2 == Ng[1] - 1)
        for k, z in enumerate(zg):
            plane = f3[k]
            if z == 0:
                plane[:] = TYPE_S
                continue
            plane[h >= z] = TYPE_S
            if z == Ng[2] - 1:
                plane[:] = TYPE_E
            else:
                plane[np.broadcast_to(side2, plane.shape)] = TYPE_E
        base = log_law(zg.astype(np.float32), Ng[2], u_ref)
    else:
        raise KeyError(name)
    rng = np.random.default_rng([seed, O[0] & 0xFFFF, O[1] & 0xFFFF, O[2] & 0xFFFF])
    for c in range(3):
        uc = u3[c]
        if amp > 0.0:
            rng.random(out=uc.reshape(-1), dtype=np.float32)
            uc *= np.float32(2.0 * amp)
            uc -= np.float32(amp)
        else:
            uc[:] = 0.0
        if c == 0:
            uc += base[:, None, None]
        if name != "periodic_box":  # boundary cells carry the clean profile; solids are at rest
            for k in range(Nz):
                pl, fl = uc[k], f3[k]
                pl[fl == TYPE_E] = base[k] if c == 0 else 0.0
                pl[fl == TYPE_S] = 0.0
    return flags, rho, u


def kernel_literal(x):
    """The float the reference's OpenCL JIT sees for a per-case constant: the host value printed by to_string(float) with
    8 decimals (FX/utilities.hpp:2603-2634, 2741-2750: float32 arithmetic, scientific form outside [1,10)) and parsed again."""
    f32 = np.float32
    x = f32(x)
    sign = ""
    if x < 0:
        sign, x = "-", f32(-x)
    if not np.isfinite(x):
        return f32(x)
    e = 0
    if x >= f32(10.0):
        for lim, mul, de in ((1e32, 1e-32, 32), (1e16, 1e-16, 16), (1e8, 1e-8, 8), (1e4, 1e-4, 4), (1e2, 1e-2, 2), (1e1, 1e-1, 1)):
            if x >= f32(lim):
                x = f32(x * f32(mul)); e += de
    if f32(0.0) < x <= f32(1.0):
        for lim, mul, de in ((1e-31, 1e32, 32), (1e-15, 1e16, 16), (1e-7, 1e8, 8), (1e-3, 1e4, 4), (1e-1, 1e2, 2), (1e0, 1e1, 1)):
            if x < f32(lim):
                x = f32(x * f32(mul)); e -= de
    integral = int(x)
    rem = f32(f32(x - f32(integral)) * f32(1e8))
    dec = int(rem)
    if f32(rem - f32(dec)) >= f32(0.5):
        dec += 1
        if dec >= 100000000:
            dec, integral = 0, integral + 1
            if integral >= 10:
                integral, e = 1, e + 1
    text = f"{sign}{integral}.{dec:08d}" + (f"E{e}" if e != 0 else "")
    return f32(float(text))


def relaxation_rate(nu):
    """def_w = 1/tau = 1/(3 nu + 1/2) as the kernel sees it (FX/lbm.hpp:146, FX/lbm.cpp:663)."""
    return kernel_literal(np.float32(1.0) / np.float32(np.float32(3.0) * np.float32(nu) + np.float32(0.5)))
