"""numpy fp32 prototype of the divide-and-conquer tridiagonal eigensolver (the algorithm eigh_dc.cu
implements): leaf solves, rank-one merges with deflation, bisection secular solver in a shifted origin,
Gu-Eisenstat z recomputation.  Run: python tests/proto/dc_proto.py [n]"""
import sys
import numpy as np

f32 = np.float32
EPS = f32(np.finfo(np.float32).eps / 2)   # LAPACK slamch('E') (relative machine eps, 5.96e-8)


def secular_roots(d, z, rho):
    """Roots of 1 + rho * sum z_i^2/(d_i - lam) = 0 for sorted d (k,), all z != 0, rho > 0.
    Returns (orig index array, mu) with lam_j = d[orig_j] + mu_j; all arithmetic fp32."""
    k = len(d)
    z2 = (z * z).astype(f32)
    orig = np.zeros(k, dtype=np.int64)
    mu = np.zeros(k, dtype=f32)
    znorm2 = f32(z2.sum(dtype=f32))
    for j in range(k):
        if j < k - 1:
            gap = f32(d[j + 1] - d[j])
            mid = f32(gap / f32(2))
            # f at the midpoint, poles shifted to d_j
            del_ = (d - d[j]).astype(f32)
            fm = f32(1) + rho * f32(np.sum(z2 / (del_ - mid), dtype=f32))
            if fm > 0:      # root in the left half: origin d_j, mu in (0, gap/2]
                o = j; lo = f32(0); hi = mid
            else:
                o = j + 1; lo = f32(-mid); hi = f32(0)
        else:
            o = j; lo = f32(0); hi = f32(rho * znorm2)
            gap = hi
        del_ = (d - d[o]).astype(f32)

        def g(m):
            return f32(1) + rho * f32(np.sum(z2 / (del_ - m), dtype=f32))
        # bracketing bisection with geometric steps while the bracket spans decades
        for it in range(200):
            if lo == 0 or hi == 0 or (lo > 0) != (hi > 0):
                # one end is the pole (0): geometric mean impossible -> arithmetic until both non-zero... use tiny
                a = lo if lo != 0 else (hi * f32(1e-30) if False else f32(0))
            alo, ahi = abs(lo), abs(hi)
            small, big = (alo, ahi) if alo < ahi else (ahi, alo)
            if small == 0:
                m = big * f32(2 ** -20) if it < 6 else big * f32(0.5)   # walk towards the pole quickly at first
                m = f32(m if (hi > 0 or lo > 0) else -m)
                if lo < 0 or hi <= 0 and lo < 0:
                    m = -abs(m)
                else:
                    m = abs(m)
            elif big > f32(4) * small:
                m = f32(np.sqrt(f32(small) * f32(big)))
                m = m if lo >= 0 else -m
            else:
                m = f32((lo + hi) * f32(0.5))
            if m <= lo or m >= hi:
                break
            if g(m) > 0:       # g increasing in mu: root is left of m
                hi = m
            else:
                lo = m
        orig[j] = o
        mu[j] = f32((lo + hi) * f32(0.5))
    return orig, mu


def merge(d1, Q1, d2, Q2, rho0):
    """Eigen-decomposition of blockdiag(T1', T2') + |rho0| u u^T given the (sorted) decompositions of the
    modified halves.  Returns sorted eigenvalues and eigenvectors (columns)."""
    n1, n2 = len(d1), len(d2)
    m = n1 + n2
    s = f32(1) if rho0 >= 0 else f32(-1)
    z = np.concatenate([Q1[-1, :], s * Q2[0, :]]).astype(f32) * f32(1 / np.sqrt(2))
    rho = f32(2 * abs(rho0))
    d = np.concatenate([d1, d2]).astype(f32)
    Q = np.zeros((m, m), dtype=f32)
    Q[:n1, :n1] = Q1
    Q[n1:, n1:] = Q2
    # sort
    perm = np.argsort(d, kind='stable')
    d = d[perm]; z = z[perm]; Q = Q[:, perm]
    tol = f32(8) * EPS * max(np.abs(d).max(), np.abs(z).max())
    if rho * np.abs(z).max() <= tol:
        return d, Q, 0
    # deflation scan
    nd = []            # non-deflated positions
    defl = []          # deflated positions
    pj = -1
    for j in range(m):
        if rho * abs(z[j]) <= tol:
            defl.append(j)
            continue
        if pj < 0:
            pj = j
            continue
        s_ = z[pj]; c_ = z[j]
        tau = f32(np.hypot(c_, s_))
        t = f32(d[j] - d[pj])
        c_ = f32(c_ / tau); s_ = f32(-s_ / tau)
        if abs(t * c_ * s_) <= tol:
            z[j] = tau; z[pj] = 0
            qp, qj = Q[:, pj].copy(), Q[:, j].copy()
            Q[:, pj] = c_ * qp + s_ * qj
            Q[:, j] = -s_ * qp + c_ * qj
            t2 = f32(d[pj] * c_ * c_ + d[j] * s_ * s_)
            d[j] = f32(d[pj] * s_ * s_ + d[j] * c_ * c_)
            d[pj] = t2
            defl.append(pj)
            pj = j
        else:
            nd.append(pj)
            pj = j
    if pj >= 0:
        nd.append(pj)
    k = len(nd)
    dl = d[nd].copy(); w = z[nd].copy()
    # NOTE: after a rotation d[pj] may be slightly out of order among the deflated; final sort handles it
    orig, mu = secular_roots(dl, w, rho)
    # delta[i][j] = dl_i - lam_j, accurately
    delta = (dl[:, None] - dl[orig][None, :]).astype(f32) - mu[None, :]
    # Gu-Eisenstat: zhat_i^2 = prod_j (lam_j - dl_i) / prod_{j != i} (dl_j - dl_i)   (signs: all ratios positive)
    zh = np.zeros(k, dtype=f32)
    for i in range(k):
        p = f32(-delta[i, i])            # lam_i - dl_i  (>0)
        for j in range(k):
            if j != i:
                p = f32(p * f32(delta[i, j] / f32(dl[i] - dl[j])))
        zh[i] = f32(np.sqrt(abs(p) / rho)) * (f32(1) if w[i] >= 0 else f32(-1))
    X = (zh[:, None] / delta).astype(f32)
    X /= np.sqrt((X * X).sum(axis=0, dtype=f32), dtype=f32)[None, :]
    lam = (dl[orig] + mu).astype(f32)
    Qn = (Q[:, nd].astype(f32) @ X).astype(f32)
    allv = np.concatenate([lam, d[defl]])
    allQ = np.concatenate([Qn, Q[:, defl]], axis=1)
    o = np.argsort(allv, kind='stable')
    return allv[o], allQ[:, o], k


def dc(d, e, leaf=32):
    n = len(d)
    if n <= leaf:
        T = np.diag(d.astype(np.float64)) + np.diag(e.astype(np.float64), 1) + np.diag(e.astype(np.float64), -1)
        w, V = np.linalg.eigh(T)
        return w.astype(f32), V.astype(f32), [0, 0]
    h = n // 2
    d = d.copy()
    r = e[h - 1]
    d[h - 1] = f32(d[h - 1] - abs(r)); d[h] = f32(d[h] - abs(r))
    w1, V1, s1 = dc(d[:h], e[:h - 1], leaf)
    w2, V2, s2 = dc(d[h:], e[h:], leaf)
    w, V, k = merge(w1, V1, w2, V2, r)
    return w, V, [s1[0] + s2[0] + k, s1[1] + s2[1] + n]


if __name__ == '__main__':
    import scipy.linalg as sl
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rng = np.random.default_rng(0)
    kinds = {}
    # K-FAC-like
    m = max(8, n // 3)
    F = np.eye(n)
    scale = np.logspace(0, -2, n)[None, :]
    mix = rng.standard_normal((n, n)) / np.sqrt(n)
    for _ in range(3):
        x = np.maximum(rng.standard_normal((m, n)) @ mix + 0.3, 0) * scale
        x[:, -1] = 1
        F = 0.95 * F + 0.05 * x.T @ x / m
    kinds['kfac'] = F
    Qr, _ = np.linalg.qr(rng.standard_normal((n, n)))
    kinds['geo'] = (Qr * np.logspace(0, -7, n)) @ Qr.T
    kinds['cluster'] = (Qr * np.concatenate([np.ones(n // 2), np.full(n - n // 2, 1e-4)])) @ Qr.T
    kinds['ident'] = 0.73 * np.eye(n)
    for name, F in kinds.items():
        F32 = F.astype(f32)
        # tridiagonalise in fp32 (LAPACK ssytrd)
        c, dd, ee, tau, info = sl.lapack.ssytrd(F32, lower=1)
        w, V, stats = dc(dd.astype(f32), ee.astype(f32))
        T = np.diag(dd.astype(np.float64)) + np.diag(ee.astype(np.float64), 1) + np.diag(ee.astype(np.float64), -1)
        wr = np.linalg.eigvalsh(T)
        V64 = V.astype(np.float64)
        orth = np.abs(V64.T @ V64 - np.eye(n)).max()
        res = np.linalg.norm(T @ V64 - V64 * w) / np.linalg.norm(T)
        print(f'{name:8s} n={n} nondeflated/total={stats[0]}/{stats[1]} orth={orth:.2e} resid={res:.2e} '
              f'eig err={np.abs(np.sort(w) - wr).max() / np.abs(wr).max():.2e}')
