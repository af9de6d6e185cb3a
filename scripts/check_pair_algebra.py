"""Host check of the thread-pair algebra of volume_action_pair (csrc/volume_pair.cuh): an element's sum-factorised stiffness / mass
action split between two threads that work in MIRRORED frames of the second index (thread 1: column j -> NB-1-j, quadrature column
ty -> NQ-1-ty), so that both use the same table coefficients; the off-diagonal metric term changes sign in the mirrored frame.
Compares the pair algorithm (numpy, thread by thread, with the same exchange steps as the kernel) with the direct formulas of
source/StiffnessMatrix.cpp:132-182 / source/MassMatrix.cpp:170-205 for every (n_basis, n_quad) the library instantiates."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import setup_np as S


def direct(P, D, U, G, stiff):
    if stiff:
        Pu, Du = P @ U, D @ U                      # [tx][j]
        Dx, Dy = Du @ P.T, Pu @ D.T                # [tx][ty]
        F0 = G[0] * Dx + G[1] * Dy
        F1 = G[1] * Dx + G[2] * Dy
        a0, a1 = F0 @ P, F1 @ D                    # [tx][q]
        return D.T @ a0 + P.T @ a1                 # [ii][q]
    Pu = P @ U
    return P.T @ ((G[0] * (Pu @ P.T)) @ P)


def pair(P, D, U, G, stiff):
    NQ, NB = P.shape
    JA, TA = (NB + 1) // 2, (NQ + 1) // 2
    out = [np.zeros((NB, JA)) for _ in range(2)]
    col = lambda h, jj: NB - 1 - jj if h else jj
    qcol = lambda h, tt: NQ - 1 - tt if h else tt
    mid = (1.0, 0.0)
    for tx in range(NQ):
        pu, du = [None, None], [None, None]
        for h in range(2):
            pu[h] = np.array([P[tx] @ U[:, col(h, jj)] for jj in range(JA)])
            du[h] = np.array([D[tx] @ U[:, col(h, jj)] for jj in range(JA)])
            if NB % 2:  # the middle column belongs to thread 0
                pu[h][JA - 1] *= mid[h]
                du[h][JA - 1] *= mid[h]
        part = [None, None]
        for h in range(2):
            o = 1 - h
            a0o, a0n, a1o, a1n = (np.zeros(JA) for _ in range(4))
            for tt in range(TA):
                ty = qcol(h, tt)
                dead = NQ % 2 and h == 1 and tt == TA - 1  # middle quadrature column belongs to thread 0 (metric stored as 0)
                A, B, C = (0.0, 0.0, 0.0) if dead else ((G[0][tx, ty], (-1.0 if h else 1.0) * G[1][tx, ty], G[2][tx, ty]) if stiff else (G[0][tx, ty], 0, 0))
                if stiff:
                    Dx = sum(P[tt, jj] * du[h][jj] + P[tt, NB - 1 - jj] * du[o][jj] for jj in range(JA))
                    Dy = sum(D[tt, jj] * pu[h][jj] + D[tt, NB - 1 - jj] * pu[o][jj] for jj in range(JA))
                    F0, F1 = A * Dx + B * Dy, B * Dx + C * Dy
                    for q in range(JA):
                        a0o[q] += P[tt, q] * F0
                        a0n[q] += P[tt, NB - 1 - q] * F0
                        a1o[q] += D[tt, q] * F1
                        a1n[q] += D[tt, NB - 1 - q] * F1
                else:
                    ppu = sum(P[tt, jj] * pu[h][jj] + P[tt, NB - 1 - jj] * pu[o][jj] for jj in range(JA))
                    val = A * ppu
                    for q in range(JA):
                        a0o[q] += P[tt, q] * val
                        a0n[q] += P[tt, NB - 1 - q] * val
            part[h] = (a0o, a0n, a1o, a1n)
        for h in range(2):
            o = 1 - h
            a0 = part[h][0] + part[o][1]
            a1 = part[h][2] + part[o][3]
            for q in range(JA):
                out[h][:, q] += (D[tx] * a0[q] + P[tx] * a1[q]) if stiff else P[tx] * a0[q]
    Y = np.zeros((NB, NB))
    for h in range(2):
        for jj in range(JA):
            if NB % 2 and h == 1 and jj == JA - 1:
                continue
            Y[:, col(h, jj)] = out[h][:, jj]
    return Y


def main():
    rng = np.random.default_rng(0)
    worst = 0.0
    for nb in range(2, 10):
        b = S.Basis(nb)
        for nq, stiff in [(nb + 1, True), (nb + 2, True), (nb + 1, False), (1 + 3 * nb // 2 + 1, False)]:
            xq, _ = S.gauss_legendre(nq)
            P, D = b.eval(xq), b.deriv(xq)
            U = rng.uniform(-1, 1, (nb, nb))
            G = rng.uniform(0.5, 1.5, (3, nq, nq))
            G[1] -= 1.0
            ref, got = direct(P, D, U, G, stiff), pair(P, D, U, G, stiff)
            err = np.abs(ref - got).max() / np.abs(ref).max()
            worst = max(worst, err)
            assert err < 1e-12, (nb, nq, stiff, err)
    print("pair algebra ok, worst relative difference %.2e" % worst)
    return worst


if __name__ == "__main__":
    main()
