"""Oracle (test infrastructure): inverse-distance descriptor, its Jacobian and the
atom-perm -> descriptor-perm map.  NumPy restatement of sgdml/utils/desc.py.

Pair order everywhere is ``np.tril_indices(N, -1)``: d <-> (a, b) with a > b, listed
(1,0),(2,0),(2,1),(3,0)... (desc.py:109-110, 198).
"""

import numpy as np


def tril_pairs(n_atoms):
    """Rows a and columns b (a > b) of the lower-triangle pair list (desc.py:266)."""
    a, b = np.tril_indices(n_atoms, k=-1)
    return a, b


def n_atoms_from_dim(dim_d):
    """Invert D = N(N-1)/2 (train.py:146)."""
    return int((1 + np.sqrt(8 * dim_d + 1)) / 2)


def pbc_diff(diffs, lat_and_inv):
    """Minimum-image convention (desc.py:44-77): d -= lat @ round(lat_inv @ d); lat holds the lattice vectors
    as COLUMNS; np.around rounds half to even."""
    lat, lat_inv = lat_and_inv
    c = np.einsum('ij,...j->...i', np.asarray(lat_inv, dtype=np.float64), diffs)
    return diffs - np.einsum('ij,...j->...i', np.asarray(lat, dtype=np.float64), np.around(c))


def from_R(R, lat_and_inv=None):
    """Descriptor and compressed Jacobian for M geometries.

    Follows desc.py:80-110 (_pdist), 139-163 (_r_to_desc), 166-205 (_r_to_d_desc),
    288-365 (Desc.from_R): x_d = 1/|r_a - r_b|, g_d = (r_a - r_b)/|r_a - r_b|^3; with a lattice the pair
    differences are clamped to the super cell first (desc.py:100-108, 200-201).

    R : (M, 3N) or (M, N, 3) float64.  Returns R_desc (M, D), R_d_desc (M, D, 3).
    """
    R = np.asarray(R, dtype=np.float64)
    M = R.shape[0]
    r = R.reshape(M, -1, 3)
    a, b = tril_pairs(r.shape[1])
    pdiff = r[:, a, :] - r[:, b, :]  # desc.py:193-198
    if lat_and_inv is not None:
        pdiff = pbc_diff(pdiff, lat_and_inv)
    # SciPy's pdist (desc.py:103) computes sqrt of the sum of squared differences
    dist = np.sqrt(np.sum(pdiff * pdiff, axis=-1))
    R_desc = 1.0 / dist  # desc.py:163
    R_d_desc = pdiff / (dist**3)[..., None]  # desc.py:203
    return R_desc, R_d_desc


def perm_to_tril_perm(perm):
    """Atom permutation (N,) -> descriptor permutation (D,) (desc.py:509-539).

    tril_perm[d(a, b)] = d(perm[a], perm[b]); integer, must be bit-exact.
    """
    perm = np.asarray(perm)
    n = len(perm)
    a, b = tril_pairs(n)
    pair_idx = np.zeros((n, n), dtype=np.int64)
    pair_idx[a, b] = np.arange(len(a))
    pair_idx[b, a] = np.arange(len(a))
    return pair_idx[perm[a], perm[b]].astype(int)


def tril_perms_lin(perms):
    """(S, N) atom perms -> (S*D,) linearised descriptor perms (train.py:897-904).

    tril_perms_lin[d*S + p] = tril_perm_p[d] + p*D  (Fortran flatten of (S, D)).
    """
    perms = np.asarray(perms)
    S, N = perms.shape
    D = N * (N - 1) // 2
    tril_perms = np.array([perm_to_tril_perm(p) for p in perms])
    return (tril_perms + np.arange(S)[:, None] * D).flatten('F')


def tril_perms_from_lin(lin, n_perms):
    """Inverse of :func:`tril_perms_lin`: (S*D,) -> (S, D) plain descriptor perms."""
    lin = np.asarray(lin)
    D = lin.size // n_perms
    tp = lin.reshape(D, n_perms).T  # [p, d] = lin[d*S + p]
    return tp - np.arange(n_perms)[:, None] * D


def d_desc_from_comp(R_d_desc):
    """Compressed (M, D, 3) -> dense (M, D, 3N) Jacobian (desc.py:422-471):
    J[d, 3b:3b+3] = +g_d, J[d, 3a:3a+3] = -g_d."""
    R_d_desc = np.asarray(R_d_desc)
    if R_d_desc.ndim == 2:
        R_d_desc = R_d_desc[None]
    M, D, _ = R_d_desc.shape
    N = n_atoms_from_dim(D)
    a, b = tril_pairs(N)
    out = np.zeros((M, D, N, 3))
    dr = np.arange(D)
    out[:, dr, b, :] = R_d_desc
    out[:, dr, a, :] = -R_d_desc
    return out.reshape(M, D, 3 * N)


def d_desc_dot_vec(R_d_desc, vecs):
    """(J v)_d = g_d . (v_b - v_a) (desc.py:368-385).  vecs (M, 3N) -> (M, D)."""
    R_d_desc = np.asarray(R_d_desc)
    if R_d_desc.ndim == 2:
        R_d_desc = R_d_desc[None]
    vecs = np.asarray(vecs)
    if vecs.ndim == 1:
        vecs = vecs[None]
    N = n_atoms_from_dim(R_d_desc.shape[1])
    a, b = tril_pairs(N)
    v = vecs.reshape(vecs.shape[0], N, 3)
    return np.einsum('...ij,...ij->...i', R_d_desc, v[:, b, :] - v[:, a, :])


def vec_dot_d_desc(R_d_desc, vecs):
    """(J^T w): atom b gets +g_d w_d, atom a gets -g_d w_d (desc.py:388-408).

    R_d_desc (M, D, 3) or (D, 3); vecs (M, D) or (D,) -> (M, 3N).
    """
    R_d_desc = np.asarray(R_d_desc)
    if R_d_desc.ndim == 2:
        R_d_desc = R_d_desc[None]
    vecs = np.asarray(vecs)
    if vecs.ndim == 1:
        vecs = vecs[None]
    D = R_d_desc.shape[1]
    N = n_atoms_from_dim(D)
    a, b = tril_pairs(N)
    n = max(R_d_desc.shape[0], vecs.shape[0])
    full = np.zeros((n, N, N, 3))
    full[:, a, b, :] = R_d_desc * vecs[..., None]
    full[:, b, a, :] = -full[:, a, b, :]
    return full.sum(axis=1).reshape(n, -1)
