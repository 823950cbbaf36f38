"""Numpy prototype in float32 (not a pytest, no GPU): feasibility of the Cholesky-preconditioned solver
(shifted Cholesky of T = V0^T F V0, one-sided block Jacobi on L with sorted pairs, Rayleigh-quotient
eigenvalues) against the current G = F V scheme, on a K-FAC-like sequence and on hard spectra.
    python tests/proto_cholesky_jacobi_fp32.py 256"""
import numpy as np, sys, os
f32 = np.float32
rng = np.random.default_rng(0)

def tour(r, k, nb):
    m = nb - 1
    if k == 0: a, b = r % m, m
    else: a, b = (r + k) % m, (r - k + m) % m
    return min(a, b), max(a, b)

def cyc_jacobi(M, tol_in, max_full=2):
    N = M.shape[0]; h = N // 2; M = M.copy(); W = np.eye(N, dtype=M.dtype)
    one = M.dtype.type(1)
    for sw in range(max_full):
        big = False; anyrot = False
        for st in range(N - 1):
            pq = [tour(st, k, N) for k in range(h)]
            P = np.array([p for p, q in pq]); Q = np.array([q for p, q in pq])
            apq = M[P, Q]; app = M[P, P]; aqq = M[Q, Q]
            do = np.abs(apq) > tol_in * np.maximum(np.abs(app), np.abs(aqq))
            if not do.any(): continue
            tau = np.where(do, (aqq - app) / np.where(do, 2 * apq, one), 0).astype(M.dtype)
            t = np.where(do, np.sign(tau + (tau == 0)) / (np.abs(tau) + np.sqrt(one + tau * tau)), 0).astype(M.dtype)
            c = (one / np.sqrt(one + t * t)).astype(M.dtype); s = (t * c).astype(M.dtype)
            anyrot = True; big |= bool(np.any(np.abs(s) >= 2e-3))
            Mp, Mq = M[:, P].copy(), M[:, Q].copy(); M[:, P] = c * Mp - s * Mq; M[:, Q] = s * Mp + c * Mq
            Mp, Mq = M[P, :].copy(), M[Q, :].copy(); M[P, :] = c[:, None] * Mp - s[:, None] * Mq; M[Q, :] = s[:, None] * Mp + c[:, None] * Mq
            Wp, Wq = W[:, P].copy(), W[:, Q].copy(); W[:, P] = c * Wp - s * Wq; W[:, Q] = s * Wp + c * Wq
        if not anyrot or not big: break
    return W[:, np.argsort(-np.diag(M))]

def onesided(X, Vacc=None, b=32, tol=3e-6, conv_tol=2e-5, max_sweeps=30):
    n = X.shape[1]; nb = n // b; X = X.copy(); hist = []
    for sw in range(max_sweeps):
        off = 0.0
        for r in range(nb - 1):
            for k in range(nb // 2):
                I, J = tour(r, k, nb)
                idx = np.r_[I * b:(I + 1) * b, J * b:(J + 1) * b]
                Y = X[:, idx]; M = Y.T @ Y
                d = np.abs(np.diag(M)); o = np.abs(M) / np.maximum.outer(d, d).clip(1e-30); np.fill_diagonal(o, 0)
                mx = o.max(); off = max(off, float(mx))
                if mx < tol: continue
                W = cyc_jacobi(M, min(tol / 8, 1e-6))
                X[:, idx] = Y @ W
                if Vacc is not None: Vacc[:, idx] = Vacc[:, idx] @ W
        hist.append(off)
        if off < conv_tol: break
        if len(hist) >= 4 and off > 0.9 * hist[-2] and off < 3e-5: break   # stagnation at the rounding floor
    return X, hist

def ferr(F, Q, lam, damp):
    F = F.astype(np.float64); Q = Q.astype(np.float64); lam = lam.astype(np.float64)
    w, U = np.linalg.eigh(F); d = damp * w.max()
    ref = (U / (np.clip(w, 0, None) + d)) @ U.T; got = (Q / (np.clip(lam, 0, None) + d)) @ Q.T
    return np.linalg.norm(got - ref) / np.linalg.norm(ref)

def solve_fv(F, V):
    Vacc = V.copy(); G, hist = onesided(F @ V, Vacc)
    nv = np.linalg.norm(Vacc, axis=0)
    return Vacc / nv, np.linalg.norm(G, axis=0) / nv, hist

def solve_chol(F, V, shift=2e-6):
    T = V.T @ (F @ V); T = (T + T.T) * f32(0.5)
    p = np.argsort(-np.diag(T)); T = T[np.ix_(p, p)]; Vp = V[:, p]
    delta = f32(shift) * np.max(np.diag(T))
    L = np.linalg.cholesky(T + delta * np.eye(T.shape[0], dtype=T.dtype)).astype(T.dtype)
    X, hist = onesided(L)
    Q = Vp @ (X / np.linalg.norm(X, axis=0))
    Q = Q / np.linalg.norm(Q, axis=0)
    lam = np.einsum('ij,ij->j', Q, F @ Q)          # Rayleigh quotients
    return Q, lam, hist

def kfac_seq(n, m, steps):
    scale = np.logspace(0, -2, n)[None, :]
    mix = rng.standard_normal((n, n)) / np.sqrt(n)
    F = np.eye(n)
    for _ in range(steps):
        x = np.maximum(rng.standard_normal((m, n)) @ mix + 0.3, 0) * scale
        x[:, -1] = 1
        F = 0.95 * F + 0.05 * x.T @ x / m
        yield F

def hard(n, kind):
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    if kind == 'geo': lam = np.logspace(0, -7, n)
    elif kind == 'lowrank': lam = np.r_[np.linspace(5, 0.5, 8), np.full(n - 8, 0.95)]
    else: lam = np.r_[np.ones(n // 2), np.full(n - n // 2, 1e-4)]
    return (Q * lam) @ Q.T

n = int(sys.argv[1])
print('--- K-FAC-like sequence, fp32')
for name, solver in (('FV', solve_fv), ('chol', solve_chol)):
    V = np.eye(n, dtype=f32); out = []
    for t, F in enumerate(kfac_seq(n, n * 5 // 16, 4)):
        F = F.astype(f32)
        Q, lam, hist = solver(F, V); V = Q
        orth = np.abs(Q.T.astype(np.float64) @ Q.astype(np.float64) - np.eye(n)).max()
        out.append(f't{t}: sw={len(hist)} err1e-3={ferr(F, Q, lam, 1e-3):.1e} err1e-5={ferr(F, Q, lam, 1e-5):.1e} orth={orth:.0e}')
    print(f'{name:5s}', ' | '.join(out), flush=True)
    rng = np.random.default_rng(0)
print('--- hard spectra, cold, fp32')
for kind in ('geo', 'lowrank', 'cluster'):
    F = hard(n, kind).astype(f32)
    for name, solver in (('FV', solve_fv), ('chol', solve_chol)):
        Q, lam, hist = solver(F, np.eye(n, dtype=f32))
        print(f'{kind:8s} {name:5s} sw={len(hist)} err1e-3={ferr(F, Q, lam, 1e-3):.1e} err1e-5={ferr(F, Q, lam, 1e-5):.1e}', flush=True)
