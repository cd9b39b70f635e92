"""Bring-up harness for the tcgen05 FP64-via-INT8 GEMM (csrc/ozaki.cu): against NumPy FP64 and against an
exact NumPy model of the slicing.  GPU only."""

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['128', '64'], autouse=True)
def unit_width(request, monkeypatch):
    """Both pipeline-unit widths of the kernel: 128-byte swizzle (9-slot ring) first -- the layout every library
    GEMM uses, so failures there point at the kernel logic rather than at the 64-byte swizzle descriptors."""
    monkeypatch.setenv('SGDML_B200_OZAKI_BK', request.param)
    return request.param


def _run(m, n, k, S, tri=False, alpha=1.0, seed=0, scale_rows=False):
    import torch

    from sgdml_b200 import _lib

    _lib.require_gpu()
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, k))
    B = A if tri else rng.standard_normal((n, k))
    if scale_rows:  # rows of very different magnitude: the per-row exponents matter
        A = A * np.exp2(rng.integers(-20, 20, size=(m, 1)).astype(np.float64))
        if not tri:
            B = B * np.exp2(rng.integers(-20, 20, size=(n, 1)).astype(np.float64))
        else:
            B = A
    C0 = rng.standard_normal((m, n))
    Ad = torch.from_numpy(A).cuda()
    Bd = Ad if tri else torch.from_numpy(B).cuda()
    Cd = torch.from_numpy(C0).cuda()
    _lib.check(
        _lib.lib().sgdml_b200_ozaki_gemm_nt(
            m, n, k, float(alpha), Ad.data_ptr(), k, Bd.data_ptr(), k, Cd.data_ptr(), n, S, 1 if tri else 0, _lib.current_stream()
        ),
        'ozaki_gemm_nt',
    )
    torch.cuda.synchronize()
    return A, B, C0, Cd.cpu().numpy()


def _scale(A, B):
    return np.abs(A) @ np.abs(B).T  # componentwise error bound of a dot product


@pytest.mark.parametrize('m,n,k', [(128, 64, 128), (128, 64, 256), (256, 128, 128), (300, 200, 130), (129, 65, 1000), (64, 8, 40)])
def test_ozaki_gemm_matches_fp64(m, n, k):
    A, B, C0, C = _run(m, n, k, 7)
    ref = C0 + A @ B.T
    assert np.max(np.abs(C - ref) / (_scale(A, B) + 1e-300)) < 1e-12


def test_ozaki_gemm_row_scaling_and_alpha():
    A, B, C0, C = _run(200, 136, 384, 7, alpha=-1.0, scale_rows=True, seed=3)
    ref = C0 - A @ B.T
    assert np.max(np.abs(C - ref) / (_scale(A, B) + np.abs(C0) + 1e-300)) < 1e-12


def test_ozaki_gemm_tri():
    A, B, C0, C = _run(384, 384, 256, 7, tri=True, alpha=-1.0, seed=5)
    ref = C0 - A @ A.T
    il = np.tril_indices(384)
    assert np.max(np.abs(C[il] - ref[il]) / (_scale(A, A)[il] + 1e-300)) < 1e-12


@pytest.mark.parametrize('S', [4, 5, 6, 7])
def test_ozaki_gemm_slice_count(S):
    """The error falls by 2^-7 per slice (tools/ozaki_study.py: 7e-8, 1e-10, 4e-12, 2e-14 for S = 4..7)."""
    A, B, C0, C = _run(256, 192, 512, S, seed=7)
    err = rel_err(C - C0, A @ B.T)
    assert err < 4.0 * 2.0 ** (-7 * S + 4)


def test_potrf_with_int8_trailing_updates(monkeypatch):
    """Cholesky with the trailing updates on the tcgen05 int8 path (SGDML_B200_OZAKI_SLICES=7) against the
    FP64 DMMA factorisation of the same matrix."""
    import torch

    from sgdml_b200 import _lib

    _lib.require_gpu()
    n = 3000
    rng = np.random.default_rng(1)
    G = rng.standard_normal((n, n // 4))
    A = G @ G.T + 1e-3 * np.eye(n)  # condition ~1e6
    outs = {}
    for mode in ('0', '7'):
        monkeypatch.setenv('SGDML_B200_OZAKI_SLICES', mode)
        Ad = torch.from_numpy(A.copy()).cuda()
        _lib.check(_lib.lib().sgdml_b200_potrf(Ad.data_ptr(), n, n, _lib.current_stream()), 'potrf')
        torch.cuda.synchronize()
        outs[mode] = np.tril(Ad.cpu().numpy())
    L0, L7 = outs['0'], outs['7']
    assert rel_err(L7 @ L7.T, A) < 1e-12
    assert rel_err(L7, L0) < 1e-9


def test_large_descriptor_predictor_on_int8_path(monkeypatch):
    """GEMM-composed predictor (D > 256) with its four contractions on the tcgen05 int8 path, 4 and 5
    slices: forces against the oracle (tools/ozaki_study.py predict: 8.8e-9 / 6.5e-11)."""
    import sgdml_b200
    from oracle import predict as opredict
    from sgdml_b200 import synth

    N, M = 30, 40
    perms = synth.rotor_swap_group(N, 1, 1)
    model = synth.random_model(N, M, perms, 30, seed=2)
    Rq = synth.geometries(N, 9, 1).reshape(9, -1)
    E_ref, F_ref = opredict.Predictor(model).predict(Rq)
    for S, tol in ((4, 1e-6), (5, 1e-8), (7, 1e-11)):
        monkeypatch.setenv('SGDML_B200_OZAKI_PREDICT_SLICES', str(S))
        E, F = sgdml_b200.GDMLPredict(model).predict(Rq)
        assert rel_err(F, F_ref) < tol and rel_err(E, E_ref) < tol
