"""Numpy prototype (not a pytest, no GPU): block one-sided Jacobi on a K-FAC-like factor sequence;
env SORT=1/-1 writes the pair columns sorted by their new norm (descending/ascending), BIG / MAXFULL /
MODES tune the inner sweeps.   SORT=1 MODES=full2 python tests/proto_sorted_pairs.py 512 160"""
import numpy as np, sys, time, os
BIG = float(os.environ.get('BIG', '2e-3'))
MAXFULL = int(os.environ.get('MAXFULL', '2'))
SORT = int(os.environ.get('SORT', '0'))
rng = np.random.default_rng(0)

def tour(r, k, nb):
    m = nb - 1
    if k == 0: a, b = r % m, m
    else: a, b = (r + k) % m, (r - k + m) % m
    return min(a, b), max(a, b)

def rot_params(M, P, Q, tol_in, thr_fn):
    apq = M[P, Q]; app = M[P, P]; aqq = M[Q, Q]
    thr = thr_fn(app, aqq)
    do = np.abs(apq) > tol_in * thr
    tau = np.where(do, (aqq - app) / np.where(do, 2 * apq, 1.0), 0.0)
    t = np.where(do, np.sign(tau + (tau == 0)) / (np.abs(tau) + np.sqrt(1 + tau * tau)), 0.0)
    c = 1 / np.sqrt(1 + t * t); s = t * c
    return c, s

def apply_rots(M, W, P, Q, c, s):
    # columns
    Mp, Mq = M[:, P].copy(), M[:, Q].copy()
    M[:, P] = c * Mp - s * Mq; M[:, Q] = s * Mp + c * Mq
    Mp, Mq = M[P, :].copy(), M[Q, :].copy()
    M[P, :] = c[:, None] * Mp - s[:, None] * Mq; M[Q, :] = s[:, None] * Mp + c[:, None] * Mq
    Wp, Wq = W[:, P].copy(), W[:, Q].copy()
    W[:, P] = c * Wp - s * Wq; W[:, Q] = s * Wp + c * Wq

def inner(M, mode, tol_in, max_full=None):
    max_full = MAXFULL if max_full is None else max_full
    """returns W, number of parallel steps executed"""
    N = M.shape[0]; h = N // 2
    M = M.copy(); W = np.eye(N)
    thr_fn = lambda a, b: np.maximum(np.abs(a), np.abs(b))
    steps = 0
    def full_sweep():
        nonlocal steps
        big = False; any_rot = False
        for st in range(N - 1):
            pq = [tour(st, k, N) for k in range(h)]
            P = np.array([p for p, q in pq]); Q = np.array([q for p, q in pq])
            c, s = rot_params(M, P, Q, tol_in, thr_fn)
            if not np.any(s != 0): continue
            steps += 1; any_rot = True
            big |= bool(np.any(np.abs(s) >= BIG))
            apply_rots(M, W, P, Q, c, s)
        return any_rot, big
    def bip_sweep():
        nonlocal steps
        big = False; any_rot = False
        for st in range(h):
            P = np.arange(h); Q = h + (P + st) % h
            c, s = rot_params(M, P, Q, tol_in, thr_fn)
            if not np.any(s != 0): continue
            steps += 1; any_rot = True
            big |= bool(np.any(np.abs(s) >= BIG))
            apply_rots(M, W, P, Q, c, s)
        return any_rot, big
    if mode == 'full2':
        for _ in range(max_full):
            a, b = full_sweep()
            if not a or not b: break
    elif mode == 'bip+full':
        a, b = bip_sweep()
        if a and b:
            for _ in range(max_full):
                a, b = full_sweep()
                if not a or not b: break
    elif mode == 'bip':
        bip_sweep()
    elif mode == 'bip2':
        a, b = bip_sweep()
        if a and b: bip_sweep()
    elif mode == 'bip+full1':
        a, b = bip_sweep()
        if a and b: full_sweep()
    return W, steps

def block_jacobi(F, V0, mode, b=32, tol=3e-6, conv_tol=2e-5, max_sweeps=30):
    n = F.shape[0]; nb = n // b
    V = V0.copy(); G = F @ V
    tot_steps = 0; hist = []
    for sw in range(max_sweeps):
        sweep_off = 0.0
        for r in range(nb - 1):
            for k in range(nb // 2):
                I, J = tour(r, k, nb)
                idx = np.r_[I * b:(I + 1) * b, J * b:(J + 1) * b]
                X = G[:, idx]; M = X.T @ X
                d = np.abs(np.diag(M)); off = np.abs(M) / np.maximum.outer(d, d).clip(1e-300); np.fill_diagonal(off, 0)
                mx = off.max(); sweep_off = max(sweep_off, mx)
                if mx < tol: continue
                W, st = inner(M, mode, min(tol / 8, 1e-6))
                if SORT:
                    dnew = np.einsum('ij,ik,kj->j', W, M, W)
                    W = W[:, np.argsort(-dnew if SORT > 0 else dnew)]
                tot_steps += st
                G[:, idx] = X @ W; V[:, idx] = V[:, idx] @ W
        hist.append(sweep_off)
        if sweep_off < conv_tol: break
    return G, V, hist, tot_steps

def factors(n, m, steps):
    scale = np.logspace(0, -2, n)[None, :]
    mix = rng.standard_normal((n, n)) / np.sqrt(n)
    F = np.eye(n)
    for _ in range(steps):
        x = np.maximum(rng.standard_normal((m, n)) @ mix + 0.3, 0) * scale
        x[:, -1] = 1
        F = 0.95 * F + 0.05 * x.T @ x / m
        yield F

n, m = int(sys.argv[1]), int(sys.argv[2])
Fs = list(factors(n, m, 4))
for mode in os.environ.get('MODES', 'full2,bip+full,bip+full1,bip2,bip').split(','):
    V = np.eye(n); out = []
    for t, F in enumerate(Fs):
        t0 = time.time()
        G, V, hist, steps = block_jacobi(F, V, mode)
        lam = np.linalg.norm(G, axis=0) / np.linalg.norm(V, axis=0); Q = V / np.linalg.norm(V, axis=0)
        w, U = np.linalg.eigh(F); sc = w.max(); damp = 1e-3 * sc
        ref = (U / (w + damp)) @ U.T; got = (Q / (lam + damp)) @ Q.T
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        out.append(f't{t}: sweeps={len(hist)} steps={steps} err={err:.1e}')
    print(f'{mode:10s}', ' | '.join(out), flush=True)
