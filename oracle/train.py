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
    """y = F_train.ravel()/std (train.py:939-947); with energy constraints the centred, sign-flipped energies
    are appended before the standardisation.  Returns (y, y_std, E_train_mean or None)."""
    y = np.asarray(task['F_train'], dtype=np.float64).ravel().copy()
    E_train_mean = None
    if task['use_E'] and task.get('use_E_cstr', False):
        E_train = np.asarray(task['E_train'], dtype=np.float64).ravel().copy()
        E_train_mean = np.mean(E_train)
        y = np.hstack((y, -E_train + E_train_mean))
    y_std = np.std(y)
    y /= y_std
    return y, y_std, E_train_mean


def create_model(task, solver, R_desc, R_d_desc, tril_perms_lin, std, alphas_F, alphas_E=None):
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
    if task['use_E'] and task.get('use_E_cstr', False):
        model['alphas_E'] = alphas_E  # train.py:822-823
    if 'lattice' in task:
        model['lattice'] = task['lattice']  # train.py:826-827
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
    lat_and_inv = None
    if 'lattice' in task:  # train.py:906-911
        lat = np.asarray(task['lattice'], dtype=np.float64)
        lat_and_inv = (lat, np.linalg.inv(lat))
    R_desc, R_d_desc = odesc.from_R(R, lat_and_inv)  # train.py:926-935
    y, y_std, E_train_mean = labels(task)
    use_E_cstr = E_train_mean is not None
    if use_E_cstr:
        K = oassemble.assemble_E_cstr(R_desc, R_d_desc, tril_perms_lin, task['sig'])
    else:
        K = oassemble.assemble(R_desc, R_d_desc, tril_perms_lin, task['sig'], n_procs=n_procs)
    alphas = osolve.analytic_solve(K.copy() if return_K else K, y, task['lam'])
    alphas_F, alphas_E = (alphas[:-n_train], alphas[-n_train:]) if use_E_cstr else (alphas, None)  # train.py:1052-1056
    model = create_model(task, 'analytic', R_desc, R_d_desc, tril_perms_lin, y_std, alphas_F, alphas_E)
    if model['use_E']:
        # train.py:1071-1086: with energy constraints c is the mean training energy
        model['c'] = recov_int_const(model, task, R_desc, R_d_desc) if E_train_mean is None else E_train_mean
    if return_K:
        return model, K
    return model
