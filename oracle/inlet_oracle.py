"""TEST INFRASTRUCTURE: CPU restatement of the reference's inflow interpolators, in numpy float32 / Python float (IEEE double) arithmetic with the
reference's operation order. Only tests/ may import this; the product path never does.

  nearest_eval     NearestNeighborInterpolator::eval + InletVelocityField::operator()   FX/interpolation.cpp:53-64
  knn_select       the plane choice and the K = 64 selection loop of KNNInterpolatorHD::eval  FX/interpolation_hd.cpp:184-296
  knn_fit          its weighted quadratic fit / Gaussian-weighted mean                  FX/interpolation_hd.cpp:298-410, solve_6x6_3rhs :57-152
  knn_hd_eval      KNNInterpolatorHD::eval + InletVelocityFieldHD::operator()           FX/interpolation_hd.cpp:184-421

Pinned: tests/golden/ref_inlet.npz holds the velocities the REFERENCE's own classes (compiled from the sources where they lie, baseline/inlet_parity.cpp under
LUW_INLET_GOLDEN) return for four sample clouds; tests/test_inlet_oracle.py requires this restatement to reproduce them. exp() is math.exp, i.e. the C library's, like the
reference's std::exp: on the machine the fixture was made on the match is bit for bit; another libm build may differ in the last bit of a weight.
"""
import math

import numpy as np

K = 64
F32 = np.float32


def nearest_eval(P, U, pos, z_threshold):
    """u = 0 below the threshold, else the velocity of the FIRST sample at the smallest squared distance dot(pos - P[i], pos - P[i])."""
    out = np.zeros((len(pos), 3), F32)
    for c, p in enumerate(pos):
        if p[2] < F32(z_threshold) or len(P) == 0:
            continue
        d = p[None, :] - P
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        out[c] = U[int(np.argmin(d2))]  # first minimum, like `if (d2 < best)`
    return out


def _bounds(P):
    lo, hi = P.min(axis=0), P.max(axis=0)
    ext = hi - lo
    tol = F32(1e-5) * F32(max(ext[0], ext[1], ext[2])) + F32(1e-6)
    return lo, hi, tol


def plane_of(P, p):
    """Face plane of the sample cloud's bounding box that the position is closest to (first minimum of |x-xmin|, |x-xmax|, |y-ymin|, |y-ymax|, |z-zmax|)."""
    lo, hi, _ = _bounds(P)
    d = [abs(p[0] - lo[0]), abs(p[0] - hi[0]), abs(p[1] - lo[1]), abs(p[1] - hi[1]), abs(p[2] - hi[2])]
    plane, dmin = 0, d[0]
    for f in range(1, 5):
        if d[f] < dmin:
            dmin, plane = d[f], f
    return plane


def on_plane(P, plane):
    """Indices of the samples on that plane (within plane_tol), in sample order, and their in-plane coordinates (a, b): (y, z), (y, z), (x, z), (x, z), (x, y)."""
    lo, hi, tol = _bounds(P)
    ref = [lo[0], hi[0], lo[1], hi[1], hi[2]][plane]
    comp = [0, 0, 1, 1, 2][plane]
    idx = np.flatnonzero(np.abs(P[:, comp] - ref) <= tol)
    ab = [(1, 2), (1, 2), (0, 2), (0, 2), (0, 1)][plane]
    return idx, np.ascontiguousarray(P[idx][:, ab])


def knn_select(q, ca, cb):
    """The selection loop over the on-plane samples q[j] = (a, b) for a cell at (ca, cb): returns (exact, kept, max_r2_kept): exact = index of the first sample with
    r2 <= 1e-16 or -1; kept = the sample indices in the reference's SLOT order; max_r2_kept as the reference leaves it."""
    s1 = q[:, 0] - F32(ca)
    s2 = q[:, 1] - F32(cb)
    r2 = s1 * s1 + s2 * s2  # float32, every operation rounded
    best_r2, best_i = [], []
    max_r2_kept = F32(0.0)
    for i in range(len(q)):
        v = r2[i]
        if v <= F32(1e-16):
            return i, best_i, max_r2_kept
        if len(best_r2) < K:
            best_r2.append(v); best_i.append(i)
            if v > max_r2_kept:
                max_r2_kept = v
        else:
            worst_k, worst = 0, best_r2[0]
            for k in range(1, K):
                if best_r2[k] > worst:
                    worst, worst_k = best_r2[k], k
            if v < worst:
                best_r2[worst_k], best_i[worst_k] = v, i
                max_r2_kept = best_r2[0]
                for k in range(1, K):
                    if best_r2[k] > max_r2_kept:
                        max_r2_kept = best_r2[k]
    return -1, best_i, max_r2_kept


def _solve6(A, b):
    """Gaussian elimination with partial pivoting (first largest |a[i][k]|), three right-hand sides, pivots below 1e-18 give up."""
    a = [row[:] for row in A]
    r = [col[:] for col in b]
    for k in range(6):
        pivot, largest = k, abs(a[k][k])
        for i in range(k + 1, 6):
            if abs(a[i][k]) > largest:
                largest, pivot = abs(a[i][k]), i
        if largest < 1e-18:
            return None
        if pivot != k:
            a[k], a[pivot] = a[pivot], a[k]
            for c in range(3):
                r[c][k], r[c][pivot] = r[c][pivot], r[c][k]
        inv = 1.0 / a[k][k]
        for i in range(k + 1, 6):
            fct = a[i][k] * inv
            if fct == 0.0:
                continue
            for j in range(k, 6):
                a[i][j] -= fct * a[k][j]
            for c in range(3):
                r[c][i] -= fct * r[c][k]
    x = [[0.0] * 6 for _ in range(3)]
    for i in range(5, -1, -1):
        s = [r[0][i], r[1][i], r[2][i]]
        for j in range(i + 1, 6):
            for c in range(3):
                s[c] -= a[i][j] * x[c][j]
        if abs(a[i][i]) < 1e-18:
            return None
        inv = 1.0 / a[i][i]
        for c in range(3):
            x[c][i] = s[c] * inv
    return x


def knn_fit(q, U_on_plane, kept, ca, cb, max_r2_kept):
    """Velocity from the kept samples (slot order): weighted quadratic fit with >= 6 samples and a regular system, else the Gaussian-weighted mean."""
    used = len(kept)
    if used == 0:
        return np.zeros(3, F32)
    R2 = float(max(F32(max_r2_kept), F32(1e-12)))
    sigma2 = 0.25 * R2
    rows = []
    for j in kept:
        q1 = float(F32(q[j, 0] - F32(ca)))
        q2 = float(F32(q[j, 1] - F32(cb)))
        w = math.exp(-(q1 * q1 + q2 * q2) / (2.0 * sigma2))
        rows.append((q1, q2, w, [float(v) for v in U_on_plane[j]]))
    if used >= 6:
        A = [[0.0] * 6 for _ in range(6)]
        b = [[0.0] * 6 for _ in range(3)]
        for q1, q2, w, u in rows:
            phi = [1.0, q1, q2, q1 * q1, q1 * q2, q2 * q2]
            for i in range(6):
                wi = w * phi[i]
                for j in range(6):
                    A[i][j] += wi * phi[j]
            for i in range(6):
                wphi = w * phi[i]
                for c in range(3):
                    b[c][i] += wphi * u[c]
        x = _solve6(A, b)
        if x is not None:
            return np.array([x[0][0], x[1][0], x[2][0]], np.float64).astype(F32)
    acc, wsum = [0.0, 0.0, 0.0], 0.0
    for q1, q2, w, u in rows:
        for c in range(3):
            acc[c] += w * u[c]
        wsum += w
    if wsum <= 0.0:
        return np.zeros(3, F32)
    inv = 1.0 / wsum
    return np.array([acc[0] * inv, acc[1] * inv, acc[2] * inv], np.float64).astype(F32)


def knn_hd_eval(P, U, pos, z_base):
    out = np.zeros((len(pos), 3), F32)
    if len(P) == 0:
        return out
    planes = {}
    for c, p in enumerate(pos):
        if p[2] < F32(z_base):
            continue
        plane = plane_of(P, p)
        if plane not in planes:
            planes[plane] = on_plane(P, plane)
        idx, q = planes[plane]
        ab = [(1, 2), (1, 2), (0, 2), (0, 2), (0, 1)][plane]
        ca, cb = p[ab[0]], p[ab[1]]
        exact, kept, max_r2 = knn_select(q, ca, cb)
        out[c] = U[idx[exact]] if exact >= 0 else knn_fit(q, U[idx], kept, ca, cb, max_r2)
    return out
