"""Known-answer matrices for the pivoted QR (`qr(M, ColumnNorm())` = LAPACK dgeqp3) restatements.

    python tests/golden/make_qrcp_kat.py      ->  tests/golden/qrcp_kat.npz

tol3z_*: 4 x 3 matrices on which dlaqp2's partial-norm recompute rule decides the pivot order, built so that every
IEEE-754 implementation of dlaqp2 computes the same bits whatever its summation order or FMA contraction:
    column 0 = 3 e_0                  (pivot of step 0; nothing below the diagonal, so dlarfg gives tau = 0, H = I)
    column 1 = e_0 + s e_1,  s = k 2^-20   (1 + s^2 is exact in binary64; after step 0 its true remaining norm is s)
    column 2 = g e_2,        g = s (1 + 2e-9)
After step 0 the downdate ratio of column 1 is temp2 = s^2 / (1 + s^2), inside (sqrt(2^-53), sqrt(2^-52)] =
(1.0537e-8, 1.4901e-8].  LAPACK (tol3z = sqrt(dlamch('Epsilon')) = sqrt(2^-53)) keeps the DOWNDATED norm, which
cancellation makes 4e-9 .. 7e-9 (relative) larger than s -- with or without an FMA in 1 - tq^2 -- hence larger than g:
pivots [0, 1, 2].  A restatement with tol3z = sqrt(eps) = 1.4901e-8 recomputes the norm (exactly s < g): pivots
[0, 2, 1].  Each matrix is checked against scipy.linalg.lapack.dgeqp3 before it is stored.
"""
import os

import numpy as np
from scipy.linalg import lapack

HERE = os.path.dirname(os.path.abspath(__file__))


def dlaqp2_pivots(A, tol3z):
    f = np.array(A, dtype=float, order="F")
    rows, cols = f.shape
    k = min(rows, cols)
    p = np.arange(cols)
    vn1 = np.array([np.linalg.norm(f[:, j]) for j in range(cols)])
    vn2 = vn1.copy()
    for i in range(k):
        pvt = i + int(np.argmax(vn1[i:]))
        if pvt != i:
            f[:, [pvt, i]] = f[:, [i, pvt]]
            p[[pvt, i]] = p[[i, pvt]]
            vn1[pvt] = vn1[i]
            vn2[pvt] = vn2[i]
        tau = 0.0
        if i < rows - 1:
            alpha = f[i, i]
            xn = np.linalg.norm(f[i + 1:, i])
            if xn != 0:
                beta = -np.copysign(np.hypot(alpha, xn), alpha)
                tau = (beta - alpha) / beta
                f[i + 1:, i] *= 1.0 / (alpha - beta)
                f[i, i] = beta
        if i < cols - 1 and tau != 0:
            v = np.concatenate([[1.0], f[i + 1:, i]])
            w = tau * (v @ f[i:, i + 1:])
            f[i:, i + 1:] -= np.outer(v, w)
        for j in range(i + 1, cols):
            if vn1[j] != 0:
                tq = abs(f[i, j]) / vn1[j]
                temp = max(1 - tq * tq, 0.0)
                if temp * (vn1[j] / vn2[j]) ** 2 <= tol3z:
                    vn1[j] = vn2[j] = np.linalg.norm(f[i + 1:, j]) if i < rows - 1 else 0.0
                else:
                    vn1[j] *= np.sqrt(temp)
    return p


if __name__ == "__main__":
    out = {}
    for i, k in enumerate((109, 112, 118, 126)):
        sv = k * 2.0 ** -20
        A = np.zeros((4, 3))
        A[0, 0] = 3.0
        A[0, 1] = 1.0
        A[1, 1] = sv
        A[2, 2] = sv * (1.0 + 2e-9)
        pa = dlaqp2_pivots(A, 1.0536712127723509e-08)
        pb = dlaqp2_pivots(A, 1.4901161193847656e-08)
        _, jp, _, _, _ = lapack.dgeqp3(np.asfortranarray(A))
        assert np.array_equal(pa, [0, 1, 2]) and np.array_equal(pb, [0, 2, 1]) and np.array_equal(jp - 1, pa), (k, pa, pb, jp)
        out["tol3z_%d" % i] = A
        out["tol3z_%d_pivots" % i] = pa
        out["tol3z_%d_wrong" % i] = pb
    np.savez_compressed(os.path.join(HERE, "qrcp_kat.npz"), **out)
    print("stored", sorted(out))
