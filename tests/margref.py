"""High-precision reference of the Schur complement of MarginalizationInfo::marginalize
(factor/marginalization_factor.cpp:263-281): A' = Arr - Arm Amm^+ Amr, b' = br - Arm Amm^+ bm, where Amm^+ inverts the
eigenvalues of the symmetrised Amm that are > eps = 1e-8 and zeroes the rest - evaluated with mpmath at 40 digits, so
that the FP64 implementations (CPU oracle: Householder + QL like Eigen; GPU: cyclic Jacobi) can be measured against the
exact value of the SAME rule.  TEST INFRASTRUCTURE."""
import mpmath as mp
import numpy as np


def exact_schur(A, b, m, eps=1e-8, dps=40):
    """-> (A' [n x n], b' [n], eigenvalues of Amm) as float64 arrays, computed at `dps` digits"""
    old = mp.mp.dps
    mp.mp.dps = dps
    try:
        pos = A.shape[0]
        n = pos - m
        Am = mp.matrix(A.tolist())
        bm = mp.matrix(b.tolist())
        Amm = mp.matrix(m, m)
        for i in range(m):
            for j in range(m):
                Amm[i, j] = (Am[i, j] + Am[j, i]) / 2
        E, Q = mp.eigsy(Amm)
        inv = mp.matrix(m, m)
        for k in range(m):
            if E[k] > eps:
                for i in range(m):
                    vi = Q[i, k] / E[k]
                    for j in range(m):
                        inv[i, j] += vi * Q[j, k]
        Arm = Am[m:pos, 0:m]
        T = Arm * inv
        Ap = Am[m:pos, m:pos] - T * Am[0:m, m:pos]
        bp = bm[m:pos, 0] - T * bm[0:m, 0]
        return (np.array([[float(Ap[i, j]) for j in range(n)] for i in range(n)]), np.array([float(bp[i]) for i in range(n)]),
                np.array([float(E[k]) for k in range(m)]))
    finally:
        mp.mp.dps = old


def conditioning(A, b, m, rel=2.0 ** -52, seed=0, eps=1e-8, dps=40):
    """Change of the EXACT A' under a symmetric relative perturbation of the input of one FP64 ulp: what no FP64
    implementation of this rule can stay under, whatever its eigensolver.  -> max |dA'| / max |A'|"""
    rng = np.random.default_rng(seed)
    N = rng.uniform(-1.0, 1.0, A.shape)
    N = (N + N.T) / 2
    A2 = A * (1.0 + rel * N)
    Ap, _, _ = exact_schur(A, b, m, eps, dps)
    Ap2, _, _ = exact_schur(A2, b, m, eps, dps)
    return float(np.abs(Ap2 - Ap).max() / np.abs(Ap).max())
