"""Independent check of a trained model through the K.v identity (reference iterative.py:183-204):
(K - lam I) alphas is evaluated with the PREDICTOR kernels, which share no code with the assembly and Cholesky
kernels that produced alphas, and compared with the labels y the analytic solver was given (analytic.py:65-99)."""

import numpy as np


def residual_report(model, task, predictor=None):
    """-> dict(residual_rel, backward_err, force_rel_max): ||(K - lam I) alphas - y|| / ||y|| with K.alphas from
    GDMLPredict.kmatvec_train (raw sums), y = F_train / std."""
    import sgdml_b200
    from sgdml_b200.desc import Desc

    M, N = task['R_train'].shape[:2]
    if predictor is None:
        predictor = sgdml_b200.GDMLPredict(model)
        _, R_d_desc = Desc(N).from_R(np.ascontiguousarray(task['R_train'], dtype=np.float64).reshape(M, -1))
        predictor.set_R_d_desc(R_d_desc if M > 1 else R_d_desc[None])
    alphas = np.ascontiguousarray(model['alphas_F'], dtype=np.float64)
    predictor.set_alphas(alphas)
    Kv = predictor.kmatvec_train().ravel()
    std = float(model['std'])
    y = np.asarray(task['F_train'], dtype=np.float64).ravel() / std
    lam = float(model['lam'])
    r = Kv - lam * alphas - y
    ny = float(np.linalg.norm(y))
    F_pred = Kv.reshape(M, -1) * std  # = GDMLPredict.predict on the training points (forces)
    F_ref = np.asarray(task['F_train'], dtype=np.float64).reshape(M, -1)
    return {
        'residual_rel': float(np.linalg.norm(r) / ny),
        'residual_max_rel': float(np.max(np.abs(r)) / np.max(np.abs(y))),
        # normwise backward error of the solve: ||r|| / (||Kv|| + lam ||alphas|| + ||y||)
        'backward_err': float(np.linalg.norm(r) / (np.linalg.norm(Kv) + lam * np.linalg.norm(alphas) + ny)),
        'force_rel_max_train': float(np.max(np.abs(F_pred - F_ref)) / np.max(np.abs(F_ref))),
        'alphas_norm': float(np.linalg.norm(alphas)),
    }
