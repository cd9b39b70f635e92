"""CPU: host-side logic, the C-ABI library surface and loud failure without a GPU."""

import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import oracle.desc as odesc


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'sgdml_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sgdml_b200_[a-z0-9_A-Z]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from sgdml_b200 import _lib

    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(handle, s), 'missing export ' + s
    # and the ctypes table binds every one of them
    assert set(syms) == set(_lib.SIGNATURES)
    assert _lib.lib().sgdml_b200_abi_version() == 1


def test_tril_perms_lin_host_integer_bit_exact(golden):
    """a-P0 is host integer code in the library: runs (and must be bit-exact) without a GPU."""
    from sgdml_b200.desc import Desc, tril_perms_lin

    out = tril_perms_lin(golden['perms'])
    assert out.dtype == np.int64 and np.array_equal(out, golden['tril_perms_lin'])
    assert np.array_equal(Desc.perm(golden['perms'][-1]), odesc.perm_to_tril_perm(golden['perms'][-1]))


def test_tril_perms_lin_rejects_non_permutations():
    from sgdml_b200 import _lib
    from sgdml_b200.desc import tril_perms_lin

    with pytest.raises(_lib.EngineError):
        tril_perms_lin(np.array([[0, 1, 1]]))


@pytest.mark.skipif(os.environ.get('SGDML_B200_EXPECT_GPU') == '1', reason='GPU box')
def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip('a GPU is visible')
    import sgdml_b200
    from sgdml_b200 import _lib
    from sgdml_b200.desc import Desc

    with pytest.raises(_lib.EngineError, match='no CPU fallback'):
        sgdml_b200.GDMLTrain()
    with pytest.raises(_lib.EngineError):
        Desc(3).from_R(np.zeros((2, 9)))
    rc = _lib.lib().sgdml_b200_potrf(np.eye(4).ctypes.data, 4, 4, None)
    assert rc == -1002  # SGDML_B200_ERR_NO_DEVICE


def test_synth_generators_are_deterministic():
    from sgdml_b200 import synth

    a = synth.geometries(9, 5, 0)
    b = synth.geometries(9, 5, 0)
    assert np.array_equal(a, b) and a.shape == (5, 9, 3)
    g = synth.rotor_swap_group(9, 1, 1)
    assert g.shape == (6, 9) and np.array_equal(g[0], np.arange(9))
    assert len({tuple(p) for p in g}) == 6
    g5 = synth.rotor_swap_group(42, 5, 0)
    assert g5.shape == (243, 42)
    E, F = synth.toy_pes(a)
    h = 1e-6
    ap = a.copy()
    ap[0, 2, 1] += h
    Ep, _ = synth.toy_pes(ap)
    assert abs(-(Ep[0] - E[0]) / h - F[0, 2, 1]) < 1e-4


@pytest.mark.parametrize('force_port', [False, True])
def test_bench_reference_arm_runs_on_cpu(force_port):
    """`bench.py --impl reference` prints one JSON line with the contract's keys (tiny sample): the
    unmodified reference from baseline/_ref when it is installed, the oracle port otherwise."""
    import json
    import subprocess
    import sys

    env = dict(os.environ)
    if force_port:
        env['SGDML_B200_NO_REFERENCE'] = '1'
    out = subprocess.check_output(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'ethanol', '--n-train', '20',
         '--steps', '1', '--warmup', '0', '--ref-batch', '8'],
        text=True,
        env=env,
    )
    line = json.loads(out.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0
    for k in ('metric', 'unit', 'cpu_baseline', 'e2e', 'config', 'higher_is_better'):
        assert k in line
    have_ref = os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'sgdml', 'predict.py'))
    assert line['cpu_baseline']['kind'] == ('reference' if have_ref and not force_port else 'port')
