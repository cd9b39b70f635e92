"""CPU model of the block mathematics of the assembly kernels (csrc/assemble.cu header; k_assemble_v4 / v5): the sparse
Jacobian closed form on the COMPRESSED pair arrays, with the pair index d(a, g) and the sign of a - g exactly as
k_assemble_v5 looks them up, against the oracle's dense-Jacobian block (oracle/assemble.py::kernel_block, which follows
train.py:165-227).  Pins the sign conventions and the kept-column-atom enumeration of phase A.  No GPU."""

import numpy as np
import pytest

from oracle import assemble as oassemble
from oracle import desc as odesc


def pidx(a, g):
    return a * (a - 1) // 2 + g if a > g else g * (g - 1) // 2 + a


def atom_perm_from_desc_perm(tp, N):
    """P with tp[d(a, b)] = d(P a, P b) (csrc/assemble.cu atom_perm_from_desc_perm)."""
    pairs = [(a, b) for a in range(N) for b in range(a)]
    inv = {pidx(a, b): (a, b) for a, b in pairs}
    P = np.zeros(N, dtype=int)
    for a in range(N):
        others = [o for o in range(N) if o != a][:2]
        s0 = set(inv[int(tp[pidx(a, others[0])])])
        s1 = set(inv[int(tp[pidx(a, others[1])])])
        (P[a],) = tuple(s0 & s1)
    return P


def block_v5(x_i, g_i, x_j, g_j, tril_perms, sig, kept_atoms):
    """K_ij restricted to the columns of `kept_atoms`, the way k_assemble_v5 computes it: per permutation the u rows
    (all row atoms; they also deliver |delta|^2), the v rows and the diagonal sums Dg of the kept column atoms only, then
    the 3x3 sub-blocks (row atom a, kept column atom b)."""
    N = odesc.n_atoms_from_dim(x_i.shape[0])
    out = np.zeros((3 * N, 3 * N))
    inv_div = 1.0 / (3.0 * sig**4)
    for tp in tril_perms:
        P = atom_perm_from_desc_perm(tp, N)
        Pi = np.argsort(P)
        u = np.zeros((N, 3))
        q2 = 0.0
        for a in range(N):  # u rows
            for g in range(N):
                if g == a:
                    continue
                di, dj = pidx(a, g), pidx(P[a], P[g])
                assert dj == tp[di]  # the descriptor permutation is induced by the atom permutation
                d = x_i[di] - x_j[dj]
                q2 += d * d
                u[a] += g_i[di] * (d if a > g else -d)
        u = -u
        nrm = np.sqrt(5.0) * np.sqrt(0.5 * q2)  # every pair twice
        base = np.exp(-nrm / sig) * inv_div * 5.0
        c1, c2 = base * 5.0, (sig**2 + sig * nrm) * base
        for b in kept_atoms:
            pib = Pi[b]
            v = np.zeros(3)
            for g in range(N):  # v row of the kept column atom b
                if g == b:
                    continue
                dj, di = pidx(b, g), pidx(pib, Pi[g])
                d = x_i[di] - x_j[dj]
                v += g_j[dj] * (d if b > g else -d)
            v = -v
            a_diag = pib  # Dg rows: a = P^-1 b
            Dg = np.zeros((3, 3))
            for g in range(N):
                if g == a_diag:
                    continue
                pg = P[g]
                di, dj = pidx(a_diag, g), pidx(b, pg)
                sgn = 1.0 if (a_diag > g) == (b > pg) else -1.0
                Dg += sgn * np.outer(g_i[di], g_j[dj])
            for a in range(N):  # phase B
                blk = c1 * np.outer(u[a], v)
                if b != P[a]:
                    sg = 1.0 if (a > pib) == (P[a] > b) else -1.0
                    blk += c2 * sg * np.outer(g_i[pidx(a, pib)], g_j[pidx(P[a], b)])
                else:
                    blk -= c2 * Dg
                out[3 * a : 3 * a + 3, 3 * b : 3 * b + 3] += blk
    return out


@pytest.mark.parametrize('N,n_rot', [(6, 1), (9, 2)])
def test_block_closed_form_on_compressed_pair_arrays(N, n_rot):
    from sgdml_b200 import synth

    perms = synth.rotor_swap_group(N, n_rot, 1 if N > 6 else 0)
    R = synth.geometries(N, 2, 3).reshape(2, -1)
    x, g = odesc.from_R(R)
    tril_perms = odesc.tril_perms_from_lin(odesc.tril_perms_lin(perms), len(perms))
    sig = 12.0
    ref = oassemble.kernel_block(x[0], g[0], x[1], g[1], tril_perms, sig)
    kept = [0, 2, N - 1]
    got = block_v5(x[0], g[0], x[1], g[1], tril_perms, sig, kept)
    cols = np.concatenate([np.arange(3 * b, 3 * b + 3) for b in kept])
    assert np.abs(got[:, cols] - ref[:, cols]).max() < 1e-12 * np.abs(ref).max()
    other = np.setdiff1d(np.arange(3 * N), cols)
    assert not got[:, other].any()  # columns of atoms without a kept column are never touched
