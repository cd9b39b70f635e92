#!/usr/bin/env python
"""Runs the REFERENCE's command-line functions (sgdml/cli.py: create -> train -> test) on a synthetic dataset,
optionally with the engine installed behind them (sgdml_b200.integration.install_into_reference), and prints one
JSON line with the written model's keys, errors and predictions of the model file by the UNMODIFIED reference
GDMLPredict.  Run in its own process: the reference's CLI calls os._exit on errors.

    python tools/reference_cli_dropin.py --engine b200 --workdir /tmp/x     # engine behind the reference CLI
    python tools/reference_cli_dropin.py --engine reference --workdir /tmp/y # the reference alone (CPU)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))


def make_dataset(path, n_atoms, n_points, seed=0):
    from sgdml.utils import io  # the reference

    from sgdml_b200 import synth

    perms = synth.rotor_swap_group(n_atoms, 1, 1)
    R = synth.geometries(n_atoms, n_points, seed)
    E, F = synth.toy_pes(R)
    ds = {
        'type': 'd',
        'code_version': 'synthetic',
        'name': 'synthetic_%d' % n_atoms,
        'theory': 'toy_inverse_distance',
        'z': np.arange(1, n_atoms + 1) % 9 + 1,
        'R': R,
        'E': E,
        'F': F,
        'perms': perms,
        'r_unit': 'Ang',
        'e_unit': 'kcal/mol',
    }
    ds['md5'] = io.dataset_md5(ds)
    np.savez_compressed(path, **ds)
    return ds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--engine', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workdir', required=True)
    ap.add_argument('--n-atoms', type=int, default=9)
    ap.add_argument('--n-train', type=int, default=40)
    ap.add_argument('--n-valid', type=int, default=20)
    ap.add_argument('--n-test', type=int, default=30)
    ap.add_argument('--sig', type=int, default=20)
    a = ap.parse_args()
    os.makedirs(a.workdir, exist_ok=True)

    import sgdml  # the reference (baseline/_ref)
    from sgdml import cli
    from sgdml.utils import io

    assert os.path.realpath(os.path.dirname(sgdml.__file__)).startswith(os.path.realpath(os.path.join(ROOT, 'baseline', '_ref')))
    if a.engine == 'b200':
        from sgdml_b200.integration import install_into_reference

        install_into_reference(sgdml)
    ds_path = os.path.join(a.workdir, 'dataset.npz')
    make_dataset(ds_path, a.n_atoms, a.n_train + a.n_valid + a.n_test + 10)
    _, dataset = io.is_file_type(ds_path, 'dataset')
    np.random.seed(0)
    task_dir = os.path.join(a.workdir, 'tasks')
    use_torch = False  # the engine needs no flag once installed; the reference then runs its NumPy path
    cli.create((ds_path, dataset), None, a.n_train, a.n_valid, [a.sig], False, True, False, True, task_dir=task_dir, command='create')
    task_files = sorted(f for f in os.listdir(task_dir) if f.startswith('task'))
    cli.train((task_dir, task_files), (ds_path, dataset), False, True, None, 1, use_torch, command='train')
    model_files = sorted(f for f in os.listdir(task_dir) if f.startswith('model'))
    cli.test((task_dir, model_files), (ds_path, dataset), a.n_test, True, None, 1, use_torch, command='test')
    # the written model file, read back by the UNMODIFIED reference predictor on the CPU
    from sgdml.predict import GDMLPredict as RefPredict

    model_path = os.path.join(task_dir, model_files[0])
    with np.load(model_path, allow_pickle=True) as f:
        model = {k: f[k] for k in f.files}
    Rq = dataset['R'][-8:].reshape(8, -1)
    E, F = RefPredict(model, max_processes=1, use_torch=False).predict(Rq)
    out = {
        'engine': a.engine,
        'model_file': os.path.basename(model_path),
        'keys': sorted(model.keys()),
        'shapes': {k: list(np.asarray(model[k]).shape) for k in ('R_desc', 'R_d_desc_alpha', 'alphas_F', 'perms', 'tril_perms_lin')},
        'dtypes': {k: str(np.asarray(model[k]).dtype) for k in ('R_desc', 'R_d_desc_alpha', 'alphas_F', 'perms', 'tril_perms_lin', 'sig', 'c', 'std')},
        'solver_name': str(model['solver_name']),
        'f_err': {k: float(v) for k, v in model['f_err'].item().items()},
        'e_err': {k: float(v) for k, v in model['e_err'].item().items()},
        'n_test': int(model['n_test']),
        'E_ref_predict': E.tolist(),
        'F_ref_predict_first': F[0].tolist(),
        'alphas_F_head': np.asarray(model['alphas_F'])[:6].tolist(),
        'c': float(model['c']),
        'std': float(model['std']),
    }
    print('DROPIN_JSON ' + json.dumps(out))
    sys.stdout.flush()
    os._exit(0)


if __name__ == '__main__':
    main()
