"""Oracle (test infrastructure): GDMLTrain.train for the analytic solver.

Restatement of sgdml/train.py:836-1088 (train), 727-832 (create_model) and
1090-1258 (_recov_int_const; the dataset self-diagnostics warnings are omitted, only
the integration constant train.py:1258 is computed).
"""

import numpy as np

from . import assemble as oassemble
from . import desc as odesc
from . import predict as opredict
from . import solve as osolve


def labels(task):
    """y = F_train.ravel()/std (train.py:939-947), use_E_cstr off."""
    y = np.asarray(task['F_train'], dtype=np.float64).ravel().copy()
    y_std = np.std(y)
    y /= y_std
    return y, y_std


def create_model(task, solver, R_desc, R_d_desc, tril_perms_lin, std, alphas_F):
    """train.py:727-832: hot-path keys of the model dict (metadata copied if present)."""
    N = odesc.n_atoms_from_dim(R_d_desc.shape[1])
    R_d_desc_alpha = odesc.d_desc_dot_vec(R_d_desc, alphas_F.reshape(-1, 3 * N))  # train.py:791
    model = {
        'type': 'm',
        'solver_name': solver,
        'z': task['z'],
        'R_desc': R_desc.T,  # train.py:807 (transposed!)
        'R_d_desc_alpha': R_d_desc_alpha,
        'c': 0.0,
        'std': std,
        'sig': task['sig'],
        'lam': task['lam'],
        'alphas_F': alphas_F,
        'perms': task['perms'],
        'tril_perms_lin': tril_perms_lin,
        'use_E': task['use_E'],
    }
    for k in ('dataset_name', 'dataset_theory', 'idxs_train', 'md5_train', 'idxs_valid', 'md5_valid'):
        if k in task:
            model[k] = task[k]
    return model


def recov_int_const(model, task, R_desc, R_d_desc):
    """c = mean(E_ref - E_pred(train)) (train.py:1136-1147, 1258)."""
    p = opredict.Predictor(model)
    p.set_R_desc(R_desc)
    p.set_R_d_desc(R_d_desc)
    E_pred, _ = p.predict()
    E_ref = np.squeeze(task['E_train'])
    return np.sum(E_ref - E_pred) / E_ref.shape[0]


def train(task, n_procs=1, return_K=False):
    """Analytic-solver training (train.py:836-1088 with use_analytic_solver=True)."""
    n_train, n_atoms = task['R_train'].shape[:2]
    tril_perms_lin = odesc.tril_perms_lin(task['perms'])  # train.py:897-904
    R = np.asarray(task['R_train'], dtype=np.float64).reshape(n_train, -1)
    R_desc, R_d_desc = odesc.from_R(R)  # train.py:926-935
    y, y_std = labels(task)
    K = oassemble.assemble(R_desc, R_d_desc, tril_perms_lin, task['sig'], n_procs=n_procs)
    alphas = osolve.analytic_solve(K.copy() if return_K else K, y, task['lam'])
    model = create_model(task, 'analytic', R_desc, R_d_desc, tril_perms_lin, y_std, alphas)
    if model['use_E']:
        model['c'] = recov_int_const(model, task, R_desc, R_d_desc)  # train.py:1074-1079
    if return_K:
        return model, K
    return model
