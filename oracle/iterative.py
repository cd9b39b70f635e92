"""Oracle (test infrastructure): Nystroem-preconditioned CG of sgdml/solvers/iterative.py.

NumPy/SciPy restatement of _nystroem_cholesky_factor (iterative.py:208-351), the preconditioner
P v = (B^T B v - v)/lam (iterative.py:136-138), the kernel operator (iterative.py:183-204) and the
solve (iterative.py:737-752, fixed inducing columns, no restarts).  The triangular solves and
Cholesky factorisations are SciPy/LAPACK calls in the reference too.
"""

import numpy as np
import scipy as sp
import scipy.linalg
import scipy.sparse.linalg

from . import assemble as oassemble
from . import predict as opredict


def cho_factor_stable(M, pre_reg=False, eps_mag_max=1):
    """iterative.py:414-471 (returns (factor, lower) or None)."""
    eps = np.finfo(float).eps
    eps_mag = int(np.floor(np.log10(eps)))
    if pre_reg:
        M[np.diag_indices_from(M)] += eps
        eps_mag += 1
    for reg in 10.0 ** np.arange(eps_mag, eps_mag_max + 1):
        try:
            return sp.linalg.cho_factor(M, overwrite_a=False, check_finite=False)
        except np.linalg.LinAlgError:
            M[np.diag_indices_from(M)] += reg
    return None


def nystroem_factor(R_desc, R_d_desc, tril_perms_lin, sig, lam, col_idxs, force_qr=False):
    """B = L_inv_K_mn (m, n), iterative.py:208-351.  force_qr: take the QR branch of
    iterative.py:312-322 even if the inner Cholesky factorisation would succeed (tests)."""
    col_idxs = np.asarray(col_idxs)
    K_nm = oassemble.assemble(R_desc, R_d_desc, tril_perms_lin, sig, col_idxs=col_idxs)  # iterative.py:237-247
    K_mm = -K_nm[col_idxs, :]  # iterative.py:253
    L_mm, lower = cho_factor_stable(K_mm, pre_reg=True)  # iterative.py:267
    K_nm = sp.linalg.solve_triangular(L_mm, K_nm.T, lower=lower, trans='T', check_finite=False).T  # :278-287
    inner = K_nm.T.dot(K_nm)  # iterative.py:293
    inner[np.diag_indices_from(inner)] += lam
    res = None if force_qr else cho_factor_stable(inner, eps_mag_max=-14)  # iterative.py:305
    if res is not None:
        L, lower = res
    else:  # iterative.py:312-322: R factor of the stacked ((n + m) x m) matrix [K_nm; sqrt(lam) I]
        m = K_nm.shape[1]
        L = np.linalg.qr(np.vstack([K_nm, np.sqrt(lam) * np.eye(m)]), mode='r')
        lower = False
    K_nm = sp.linalg.solve_triangular(L, K_nm.T, lower=lower, trans='T', check_finite=False).T  # :337-347
    return K_nm.T


def precon(B, lam):
    def P(v):  # iterative.py:136-138
        return (B.T.dot(B.dot(v)) - v) / lam

    return P


def kernel_op(model_like, R_desc, R_d_desc, lam):
    """iterative.py:183-204 on the oracle predictor (std = 1, c = 0)."""
    m = dict(model_like)
    m['std'], m['c'] = 1.0, 0.0
    p = opredict.Predictor(m)
    p.set_R_desc(R_desc)
    p.set_R_d_desc(R_d_desc)

    def K(v):
        p.set_alphas(v)
        return p.predict()[1].ravel() - lam * v

    return K


def solve(model_like, R_desc, R_d_desc, tril_perms_lin, sig, lam, y, inducing_pts_idxs, tol=1e-4):
    """alphas via scipy.sparse.linalg.cg(-K_op, y, M=P_op, rtol=tol) (iterative.py:740-752)."""
    n = y.size
    B = nystroem_factor(R_desc, R_d_desc, tril_perms_lin, sig, lam, inducing_pts_idxs)
    P = precon(B, lam)
    K = kernel_op(model_like, R_desc, R_d_desc, lam)
    A_op = sp.sparse.linalg.LinearOperator((n, n), matvec=lambda v: -K(v))
    P_op = sp.sparse.linalg.LinearOperator((n, n), matvec=P)
    iters = [0]

    def cb(xk):
        iters[0] += 1

    x, info = sp.sparse.linalg.cg(A_op, y, M=P_op, rtol=tol, atol=0, maxiter=10 * n, callback=cb)
    return -x, info, iters[0], B
