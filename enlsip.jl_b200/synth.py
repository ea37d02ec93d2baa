"""Synthetic inputs for the configurations of BASELINE.json.

Host-side data generation only (numpy); nothing here is on the solve path.  Oracle and engine
read the same arrays, so parity is checked on identical bits.
"""
from __future__ import annotations

import numpy as np

HS65_X0 = np.array([-5.0, 5.0, 0.0])
HS65_LOW = np.array([-4.5, -4.5, -5.0])
HS65_UPP = np.array([4.5, 4.5, 5.0])

GP_N, GP_M = 6, 128
GP_T = 10.0 * np.arange(GP_M) / 127.0
GP_LOW = np.array([0.1, 0.05, 0.0, 0.1, 0.05, 0.0])
GP_UPP = np.array([2.2, 10.0, 10.0, 5.0, 10.0, 10.0])
GP_TRUTH = np.array([2.0, 1.0, 3.0, 1.5, 0.5, 7.0])


def gen_hs65_batch(B, seed=65, start=0):
    """C2: x0_b = (-5, 5, 0) + U(-1,1)^3.  ``start`` selects a contiguous shard [start, start+B)
    of the global stream (rank r of G uses start = r*B)."""
    rng = np.random.default_rng(seed)
    if start:
        rng.bit_generator.advance(0)  # streams are regenerated from the origin; shards slice below
    u = rng.uniform(-1.0, 1.0, size=(start + B, 3))[start:]
    return np.ascontiguousarray(HS65_X0[None, :] + u)


def gen_gauss_peaks_batch(B, seed=128, start=0, chunk=1 << 18):
    """C3: returns (y [B,128], S [B], x0 [B,6], truth [B,6]).

    Per-chunk child generators (SeedSequence.spawn) make any shard reproducible without generating
    the preceding problems: chunk c covers problems [c*chunk, (c+1)*chunk).
    """
    y = np.empty((B, GP_M))
    S = np.empty(B)
    x0 = np.empty((B, GP_N))
    truth = np.empty((B, GP_N))
    first, last = start // chunk, (start + B - 1) // chunk
    children = np.random.SeedSequence(seed).spawn(last + 1)
    for c in range(first, last + 1):
        rng = np.random.default_rng(children[c])
        u = rng.uniform(-1.0, 1.0, size=(chunk, GP_N))
        u2 = rng.uniform(-1.0, 1.0, size=(chunk, GP_N))
        lo, hi = max(start, c * chunk), min(start + B, (c + 1) * chunk)
        sl = slice(lo - c * chunk, hi - c * chunk)
        xs = GP_TRUTH[None, :] * (1.0 + 0.2 * u[sl])
        d1 = GP_T[None, :] - xs[:, 2:3]
        d2 = GP_T[None, :] - xs[:, 5:6]
        g = xs[:, 0:1] * np.exp(-xs[:, 1:2] * d1 * d1) + xs[:, 3:4] * np.exp(-xs[:, 4:5] * d2 * d2)
        eps = rng.standard_normal(size=(chunk, GP_M))[sl]
        out = slice(lo - start, hi - start)
        y[out] = g + 0.01 * eps
        S[out] = xs[:, 0] / np.sqrt(xs[:, 1]) + xs[:, 3] / np.sqrt(xs[:, 4])
        span = GP_UPP - GP_LOW
        x0[out] = np.clip(xs * (1.0 + 0.1 * u2[sl]), GP_LOW + 1e-3 * span, GP_UPP - 1e-3 * span)
        truth[out] = xs
    return y, S, x0, truth


def gen_single_index(m, n, nb, seed=4, start=0, rows=None, ineq=False, chunk=1 << 16):
    """C4 / C5: W_ij ~ N(0,1)/sqrt(n), y = tanh(W x*) + 0.01 eps, x* ~ U(-1,1)^n,
    block constraints on groups of 4 parameters (rho from x*), x0 = x* (1 + 0.05 u).

    Returns dict(W [rows,n], y [rows], rho [nb], x0 [n], truth [n]).  ``start``/``rows`` select the
    row shard [start, start+rows) of the global m-row problem (row chunks have their own child
    generators, so a shard is reproducible without generating its predecessors); x*, x0, rho are
    global and identical on every shard.
    """
    rows = m - start if rows is None else rows
    root = np.random.SeedSequence(seed)
    g = np.random.default_rng(root.spawn(1)[0])
    truth = g.uniform(-1.0, 1.0, size=n)
    x0 = truth * (1.0 + 0.05 * g.uniform(-1.0, 1.0, size=n))
    blocks = (truth[: 4 * nb] ** 2).reshape(nb, 4).sum(axis=1)
    if ineq:   # half the blocks active at the solution (0.9), the rest slack (1.5)
        rho = np.where(np.arange(nb) % 2 == 0, 0.9, 1.5) * blocks
    else:
        rho = blocks.copy()
    W = np.empty((rows, n))
    y = np.empty(rows)
    first, last = start // chunk, (start + rows - 1) // chunk
    children = np.random.SeedSequence([seed, 1]).spawn(last + 1)
    for c in range(first, last + 1):
        rng = np.random.default_rng(children[c])
        lo, hi = max(start, c * chunk), min(start + rows, (c + 1) * chunk)
        Wc = rng.standard_normal(size=(chunk, n)) / np.sqrt(n)
        ec = rng.standard_normal(size=chunk)
        sl = slice(lo - c * chunk, hi - c * chunk)
        out = slice(lo - start, hi - start)
        W[out] = Wc[sl]
        y[out] = (np.tanh(Wc @ truth) + 0.01 * ec)[sl]      # whole-chunk product: a shard sees the same bits as the full run
    return dict(W=W, y=y, rho=rho, x0=x0, truth=truth)
