"""Oracle (test infrastructure): closed-form solve of the sGDML linear system.

Restatement of sgdml/solvers/analytic.py:49-151.  The dense factorisation itself is
third-party arithmetic in the reference too (SciPy -> LAPACK dpotrf/dpotrs,
analytic.py:94-99; SciPy unpinned by setup.py:53, 1.18.1 in this image), so the oracle
calls the same routine.
"""

import numpy as np
import scipy as sp
import scipy.linalg


def analytic_solve(K_ref_sign, y, lam):
    """alphas = -(-K + lam I)^-1 y  (analytic.py:65, 82, 94-99).

    K_ref_sign : K exactly as _assemble_kernel_mat returns it (not yet negated).
    Raises np.linalg.LinAlgError if -K + lam I is not positive definite (the reference
    then falls back to LU, analytic.py:101-114; see :func:`lu_solve`).
    """
    A = -np.asarray(K_ref_sign)  # analytic.py:65
    A[np.diag_indices_from(A)] += lam  # analytic.py:82
    L, lower = sp.linalg.cho_factor(A, overwrite_a=True, check_finite=False)
    return -sp.linalg.cho_solve((L, lower), y, check_finite=False)


def lu_solve(K_ref_sign, y, lam):
    """LU fallback (analytic.py:101-114)."""
    A = -np.asarray(K_ref_sign)
    A[np.diag_indices_from(A)] += lam
    return -sp.linalg.solve(A, y, check_finite=False)
