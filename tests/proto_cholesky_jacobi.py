"""Numpy prototype (not a pytest, no GPU): the current scheme (one-sided block Jacobi on G = F V with V
accumulated) against Veselic-Hari (T = V0^T F V0 sorted by diagonal, L = chol(T), one-sided block Jacobi on
the columns of L, eigenvectors = V0 * normalised columns, no V accumulation).
    python tests/proto_cholesky_jacobi.py 512 160"""
import numpy as np, sys, os, time
rng = np.random.default_rng(0)
SORT = int(os.environ.get('SORT', '1'))

def tour(r, k, nb):
    m = nb - 1
    if k == 0: a, b = r % m, m
    else: a, b = (r + k) % m, (r - k + m) % m
    return min(a, b), max(a, b)

def cyc_jacobi(M, tol_in, max_full=2):
    N = M.shape[0]; h = N // 2; M = M.copy(); W = np.eye(N)
    for sw in range(max_full):
        big = False; anyrot = False
        for st in range(N - 1):
            pq = [tour(st, k, N) for k in range(h)]
            P = np.array([p for p, q in pq]); Q = np.array([q for p, q in pq])
            apq = M[P, Q]; app = M[P, P]; aqq = M[Q, Q]
            do = np.abs(apq) > tol_in * np.maximum(np.abs(app), np.abs(aqq))
            if not do.any(): continue
            tau = np.where(do, (aqq - app) / np.where(do, 2 * apq, 1.0), 0.0)
            t = np.where(do, np.sign(tau + (tau == 0)) / (np.abs(tau) + np.sqrt(1 + tau * tau)), 0.0)
            c = 1 / np.sqrt(1 + t * t); s = t * c
            anyrot = True; big |= bool(np.any(np.abs(s) >= 2e-3))
            Mp, Mq = M[:, P].copy(), M[:, Q].copy(); M[:, P] = c * Mp - s * Mq; M[:, Q] = s * Mp + c * Mq
            Mp, Mq = M[P, :].copy(), M[Q, :].copy(); M[P, :] = c[:, None] * Mp - s[:, None] * Mq; M[Q, :] = s[:, None] * Mp + c[:, None] * Mq
            Wp, Wq = W[:, P].copy(), W[:, Q].copy(); W[:, P] = c * Wp - s * Wq; W[:, Q] = s * Wp + c * Wq
        if not anyrot or not big: break
    if SORT:
        W = W[:, np.argsort(-np.einsum('ij,ik,kj->j', W, M * 0 + 0, W) if False else -np.diag(M))]
    return W

def onesided(X, Vacc=None, b=32, tol=3e-6, conv_tol=2e-5, max_sweeps=30):
    n = X.shape[1]; nb = n // b; X = X.copy(); hist = []
    for sw in range(max_sweeps):
        off = 0.0
        for r in range(nb - 1):
            for k in range(nb // 2):
                I, J = tour(r, k, nb)
                idx = np.r_[I * b:(I + 1) * b, J * b:(J + 1) * b]
                Y = X[:, idx]; M = Y.T @ Y
                d = np.abs(np.diag(M)); o = np.abs(M) / np.maximum.outer(d, d).clip(1e-300); np.fill_diagonal(o, 0)
                mx = o.max(); off = max(off, mx)
                if mx < tol: continue
                W = cyc_jacobi(M, min(tol / 8, 1e-6))
                X[:, idx] = Y @ W
                if Vacc is not None: Vacc[:, idx] = Vacc[:, idx] @ W
        hist.append(off)
        if off < conv_tol: break
    return X, hist

def factors(n, m, steps):
    scale = np.logspace(0, -2, n)[None, :]
    mix = rng.standard_normal((n, n)) / np.sqrt(n)
    F = np.eye(n)
    for _ in range(steps):
        x = np.maximum(rng.standard_normal((m, n)) @ mix + 0.3, 0) * scale
        x[:, -1] = 1
        F = 0.95 * F + 0.05 * x.T @ x / m
        yield F

def ferr(F, Q, lam):
    w, U = np.linalg.eigh(F); d = 1e-3 * w.max()
    ref = (U / (w + d)) @ U.T; got = (Q / (lam + d)) @ Q.T
    return np.linalg.norm(got - ref) / np.linalg.norm(ref)

n, m = int(sys.argv[1]), int(sys.argv[2])
Fs = list(factors(n, m, 4))
for algo in ('FV', 'chol'):
    V = np.eye(n); out = []
    for t, F in enumerate(Fs):
        if algo == 'FV':
            Vacc = V.copy(); G, hist = onesided(F @ V, Vacc)
            lam = np.linalg.norm(G, axis=0) / np.linalg.norm(Vacc, axis=0); Q = Vacc / np.linalg.norm(Vacc, axis=0)
        else:
            T = V.T @ F @ V; T = (T + T.T) / 2
            # sort the start basis by decreasing diagonal (no pivoting needed afterwards)
            p = np.argsort(-np.diag(T)); T = T[np.ix_(p, p)]; Vp = V[:, p]
            L = np.linalg.cholesky(T)
            X, hist = onesided(L)
            lam = np.sum(X * X, axis=0); Q = Vp @ (X / np.sqrt(lam))
        V = Q
        out.append(f't{t}: sweeps={len(hist)} [' + ' '.join(f'{h:.0e}' for h in hist) + f'] err={ferr(F, Q, lam):.1e}')
    print(f'{algo:5s}', ' | '.join(out), flush=True)
