"""Oracle (test infrastructure): Matern-5/2 Hessian-kernel matrix assembly.

NumPy restatement of sgdml/train.py:97-232 (_assemble_kernel_mat_wkr, force-force
blocks) with its column selection logic, and train.py:234-300 (energy-constraint rows and columns) and the column
selection logic of train.py:1260-1407 (_assemble_kernel_mat).

Reference sign convention: this returns K as the reference's _assemble_kernel_mat
does; the analytic solver negates it (analytic.py:65).
"""

import multiprocessing as mp

import numpy as np

from . import desc as odesc


def _column_block(j, R_desc, R_d_desc, tril_perms, sig):
    """All row blocks of block-column j: returns (M*3N, 3N).

    One call == one _assemble_kernel_mat_wkr(j, ...) task (train.py:97-232), with the
    reference's Python loop over i (train.py:194) carried by NumPy broadcasting.
    """
    M, D = R_desc.shape
    S = tril_perms.shape[0]
    N = odesc.n_atoms_from_dim(D)
    dim_i = 3 * N

    # train.py:165-177: permuted copies of x_j and of the dense Jacobian J_j (rows permuted)
    xj_perms = R_desc[j][tril_perms]  # (S, D): [p, d] = x_j[tril_perm_p[d]]
    Jj = odesc.d_desc_from_comp(R_d_desc[j])[0]  # (D, 3N)
    Jj_perms = Jj[tril_perms, :]  # (S, D, 3N)

    mat52_base_div = 3 * sig**4  # train.py:179
    sqrt5 = np.sqrt(5.0)
    sig_pow2 = sig**2

    diff = R_desc[:, None, :] - xj_perms[None, :, :]  # (M, S, D)  train.py:199
    norm = sqrt5 * np.sqrt(np.sum(diff * diff, axis=2))  # (M, S)  train.py:201
    base = np.exp(-norm / sig) / mat52_base_div * 5  # train.py:202

    # train.py:209-214: sum_p (5*base_p*diff_p) (x) (diff_p . J_j^(p))
    inner = np.einsum('ipd,pdk->ipk', diff, Jj_perms)  # (M, S, 3N)
    W = np.einsum('ipd,ipk->idk', diff * (base * 5)[:, :, None], inner)  # (M, D, 3N)
    # train.py:216-220: minus sum_p J_j^(p) * (sig^2 + sig*norm_p) * base_p
    W -= np.einsum('pdk,ip->idk', Jj_perms, (sig_pow2 + sig * norm) * base)

    # train.py:223-227: K_ij = J_i^T W
    Ji = odesc.d_desc_from_comp(R_d_desc)  # (M, D, 3N)
    blk = np.einsum('idr,idk->irk', Ji, W)  # (M, 3N, 3N)
    return blk.reshape(M * dim_i, dim_i)


def _column_block_star(args):
    return _column_block(*args)


def assemble(R_desc, R_d_desc, tril_perms_lin, sig, col_idxs=None, n_procs=1):
    """K (3NM, n_cols), reference sign.

    col_idxs: None (all columns), a slice on block boundaries, or a sorted unique
    integer array of K column indices (train.py:1357-1407).
    """
    R_desc = np.ascontiguousarray(R_desc, dtype=np.float64)
    R_d_desc = np.ascontiguousarray(R_d_desc, dtype=np.float64)
    M, D = R_desc.shape
    S = len(tril_perms_lin) // D
    tril_perms = odesc.tril_perms_from_lin(tril_perms_lin, S)
    N = odesc.n_atoms_from_dim(D)
    dim_i = 3 * N
    n = M * dim_i

    if col_idxs is None:
        col_idxs = np.arange(n)
    elif isinstance(col_idxs, slice):
        col_idxs = np.arange(n)[col_idxs]
    col_idxs = np.asarray(col_idxs, dtype=np.int64)
    assert np.array_equal(col_idxs, np.unique(col_idxs))  # train.py:1341-1345

    m_idxs = col_idxs // dim_i
    pts = np.unique(m_idxs)
    K = np.empty((n, len(col_idxs)))

    jobs = [(int(j), R_desc, R_d_desc, tril_perms, sig) for j in pts]
    if n_procs > 1 and len(jobs) > 1:
        with mp.get_context('fork').Pool(n_procs) as pool:
            blocks = pool.imap(_column_block_star, jobs, chunksize=1)
            blocks = list(blocks)
    else:
        blocks = map(_column_block_star, jobs)

    pos = 0
    for j, blk in zip(pts, blocks):
        keep = col_idxs[m_idxs == j] - j * dim_i  # train.py:1393-1407 keep_idxs_3n
        K[:, pos : pos + len(keep)] = blk[:, keep]
        pos += len(keep)
    return K


def assemble_E_cstr(R_desc, R_d_desc, tril_perms_lin, sig):
    """Kernel matrix with energy constraints, (3NM + M) x (3NM + M), reference sign (train.py:234-300, 1325-1335,
    1370): the force-force part as :func:`assemble`, then for every pair of training points (i, j)
      K[n + i, blk_j] = K_fe(i, j) = -sum_p c_p delta_p^T J_j^(p),  c_p = 5 (n_p + sig) e^{-n_p/sig} / (3 sig^3),
                                      delta_p = x_i - x_j[perm_p]                       (train.py:237-248),
      K[blk_i, n + j] = the same with the roles of i and j exchanged                     (train.py:266-294),
      K[n + i, n + j] = -sum_p (1 + (n_p/sig)(1 + n_p/(3 sig))) e^{-n_p/sig},  delta = x_j - x_i[perm_p] (train.py:296-300)."""
    R_desc = np.ascontiguousarray(R_desc, dtype=np.float64)
    M, D = R_desc.shape
    S = len(tril_perms_lin) // D
    tril_perms = odesc.tril_perms_from_lin(tril_perms_lin, S)
    N = odesc.n_atoms_from_dim(D)
    dim_i = 3 * N
    n = M * dim_i
    K = np.zeros((n + M, n + M))
    K[:n, :n] = assemble(R_desc, R_d_desc, tril_perms_lin, sig)
    sqrt5 = np.sqrt(5.0)
    J = odesc.d_desc_from_comp(R_d_desc)  # (M, D, 3N)
    for j in range(M):
        xj_perms = R_desc[j][tril_perms]  # (S, D)
        Jj_perms = J[j][tril_perms, :]  # (S, D, 3N)
        for i in range(M):
            diff = R_desc[i][None, :] - xj_perms  # train.py:199
            norm = sqrt5 * np.linalg.norm(diff, axis=1)
            K_fe = 5 * diff / (3 * sig**3) * (norm[:, None] + sig) * np.exp(-norm / sig)[:, None]  # train.py:237-243
            K[n + i, j * dim_i : (j + 1) * dim_i] = -np.einsum('pd,pdk->k', K_fe, Jj_perms)  # train.py:245-248
            # column n + j (train.py:266-300): descriptor of j against the permuted copies of i
            xi_perms = R_desc[i][tril_perms]
            Ji_perms = J[i][tril_perms, :]
            diff2 = R_desc[j][None, :] - xi_perms
            norm2 = sqrt5 * np.linalg.norm(diff2, axis=1)
            K_fe2 = 5 * diff2 / (3 * sig**3) * (norm2[:, None] + sig) * np.exp(-norm2 / sig)[:, None]
            K[i * dim_i : (i + 1) * dim_i, n + j] = -np.einsum('pd,pdk->k', K_fe2, Ji_perms)
            K[n + i, n + j] = -(1 + (norm2 / sig) * (1 + norm2 / (3 * sig))).dot(np.exp(-norm2 / sig))
    return K


def kernel_block(x_i, g_i, x_j, g_j, tril_perms, sig):
    """Single (3N, 3N) block K_ij via the closed form of SURVEY.md section 8 row a-K
    (sparse J, no dense Jacobians); slow, used to cross-check :func:`assemble`."""
    D = x_i.shape[0]
    N = odesc.n_atoms_from_dim(D)
    a, b = odesc.tril_pairs(N)
    Ji = odesc.d_desc_from_comp(g_i)[0]
    Jj = odesc.d_desc_from_comp(g_j)[0]
    out = np.zeros((3 * N, 3 * N))
    for tp in tril_perms:
        delta = x_i - x_j[tp]
        nrm = np.sqrt(5.0) * np.linalg.norm(delta)
        e = np.exp(-nrm / sig)
        Jjp = Jj[tp, :]
        c1 = 25.0 / (3 * sig**4) * e
        c2 = 5.0 / (3 * sig**4) * (sig**2 + sig * nrm) * e
        out += Ji.T @ (c1 * np.outer(delta, delta @ Jjp) - c2 * Jjp)
    return out
