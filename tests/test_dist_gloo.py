"""CPU, world_size 2, gloo: the N>1 host logic (query sharding, K.v row sharding + all-gather,
coefficient broadcast, prediction sharded over training points + all-reduce, row-sharded Nystroem
preconditioner) with the ORACLE / NumPy as the injected compute functions."""

import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_model, load_golden


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class NumpyNystroemOps(object):
    """NumPy/SciPy stand-in for the engine side of dist.nystroem_factor_steps (CPU tensors)."""

    def __init__(self, g):
        self.g = g

    def assemble_rows(self, lo, hi, cols):
        import torch
        from oracle import assemble as oassemble

        g = self.g
        dim_i = 3 * int(g['n_atoms'])
        K_nm = oassemble.assemble(g['R_desc'], g['R_d_desc'], g['tril_perms_lin'], int(g['sig']), col_idxs=cols)
        return torch.from_numpy(np.ascontiguousarray(K_nm[lo * dim_i : hi * dim_i]))

    def new_square(self, m):
        import torch

        return torch.zeros((m, m), dtype=torch.float64)

    def put_neg_rows(self, A, pos, X, local_rows, m):
        A.numpy()[pos, :m] = -X.numpy()[local_rows, :m]

    force_qr = False  # tests: pretend the inner Cholesky failed (iterative.py:312-322)

    def cho_factor_stable(self, A, **kw):
        from oracle import iterative as oiter

        if self.force_qr and 'eps_mag_max' in kw:
            return False
        res = oiter.cho_factor_stable(A.numpy(), **kw)
        if res is None:
            return False
        A.numpy()[:] = np.tril(res[0]) if res[1] else np.triu(res[0]).T
        return True

    def trsm_right_lt(self, A, X, m):
        import scipy.linalg

        X.numpy()[:, :m] = scipy.linalg.solve_triangular(A.numpy(), X.numpy()[:, :m].T, lower=True).T

    def gram(self, X, m, A):
        A.numpy()[:] = X.numpy()[:, :m].T @ X.numpy()[:, :m]

    def add_diag(self, A, m, value):
        A.numpy()[np.diag_indices(m)] += value

    def scaled_identity(self, m, value):
        import torch

        return torch.from_numpy(np.eye(m) * value)

    def add_gram(self, Y, m, A):
        A.numpy()[:] += Y.numpy().T @ Y.numpy()

    def trace(self, A, m):
        return float(np.trace(A.numpy()))

    def potrf(self, A):
        try:
            A.numpy()[:] = np.linalg.cholesky(A.numpy())
        except np.linalg.LinAlgError:
            return False
        return True

    def row_sqnorms(self, X, m):
        return np.einsum('ij,ij->i', X.numpy()[:, :m], X.numpy()[:, :m])

    def project(self, X, m, v_loc):
        import torch

        return torch.from_numpy(X.numpy()[:, :m].T @ v_loc)

    def expand(self, X, m, lam, t, v_loc):
        return (X.numpy()[:, :m] @ t.numpy() - v_loc) / lam


def golden_inducing_cols(g):
    n = g['K'].shape[0]
    return np.sort(np.random.default_rng(7).choice(n, 3 * 3 * int(g['n_atoms']), replace=False))


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import predict as opredict
    from sgdml_b200 import dist as sdist

    g = load_golden(case)
    model = golden_model(g)
    p = opredict.Predictor(model)

    # 1. query batch sharded, gathered everywhere
    E, F = sdist.predict_sharded(lambda R: p.predict(R), g['R_query'])
    # 2. K.v with row sharding + one all-gather
    m1 = dict(model)
    m1['std'], m1['c'] = 1.0, 0.0
    pk = opredict.Predictor(m1)
    pk.set_R_desc(g['R_desc'])
    pk.set_R_d_desc(g['R_d_desc'])
    pk.set_alphas(g['v'])

    def rows(lo, hi):
        sub = opredict.Predictor(m1)
        sub.R_d_desc_alpha_perms = pk.R_d_desc_alpha_perms
        sub.set_R_desc(g['R_desc'][lo:hi])
        sub.set_R_d_desc(g['R_d_desc'][lo:hi])
        return sub.predict()[1]

    Kv = sdist.kmatvec_sharded(rows, g['R_desc'].shape[0])
    # 3. coefficients from rank 0
    a0 = g['alphas_F'] if rank == 0 else np.zeros_like(g['alphas_F'])
    a, c, std = sdist.broadcast_coefficients(a0, float(g['c']) if rank == 0 else 0.0, float(g['std']) if rank == 0 else 0.0)
    # 4. sum over training points sharded, one all-reduce
    tp = sdist.TrainPointShardedPredictor(model, opredict.Predictor)
    E_tp, F_tp = tp.predict(g['R_query'])
    # 5. row-sharded Nystroem preconditioner: (m x m) all-reduces, m-vector all-reduce + all-gather per P.v
    ops = NumpyNystroemOps(g)
    n_train, dim_i, lam = g['R_desc'].shape[0], 3 * int(g['n_atoms']), float(g['lam'])
    cols = golden_inducing_cols(g)
    X, lo, hi = sdist.run_steps(sdist.nystroem_factor_steps(ops, rank, world, n_train, dim_i, cols, lam), n_train)
    lev = sdist.run_steps(sdist.lev_scores_steps(ops, X, len(cols), dim_i), n_train)
    Pv = sdist.run_steps(sdist.precon_apply_steps(ops, X, len(cols), lam, g['v'], lo, hi, dim_i), n_train)
    # 6. the exchange function of the device-resident PCG (sgdml_b200_pcg) on a workspace tensor: all-reduce of
    #    an m-vector, in-place all-gather of a replicated n-vector with equal (8 points) and ragged (7 points) shards
    import torch

    exch = {}
    for n_pts in (8, 7):
        di = 3
        ws = torch.zeros(5 + n_pts * di + 4, dtype=torch.float64)
        ws[1:4] = torch.tensor([1.0, 2.0, 3.0]) * (rank + 1)
        sdist.exchange_on_workspace(ws, 1, 3, 0, n_pts, di)
        lo_, hi_ = sdist.shard_bounds(n_pts, world, rank)
        full = np.arange(n_pts * di, dtype=np.float64) + 100.0
        ws[5 + lo_ * di : 5 + hi_ * di] = torch.from_numpy(full[lo_ * di : hi_ * di])  # only the owned rows are valid
        sdist.exchange_on_workspace(ws, 5, n_pts * di, 1, n_pts, di)
        exch['ws%d' % n_pts] = ws.numpy().copy()
    np.savez(
        os.path.join(out_dir, 'r%d.npz' % rank), E=E, F=F, Kv=Kv, a=a, c=c, std=std, E_tp=E_tp, F_tp=F_tp, lev=lev, Pv=Pv,
        **exch
    )
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_range():
    from sgdml_b200.dist import shard_bounds

    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo(tmp_path):
    import torch.multiprocessing as mp

    case = 'n9_m16_s6'
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    g = load_golden(case)
    for r in range(2):
        with np.load(tmp_path / ('r%d.npz' % r)) as f:
            assert np.max(np.abs(f['F'] - g['F_query'])) < 1e-10 * np.max(np.abs(g['F_query']))
            assert np.max(np.abs(f['E'] - g['E_query'])) < 1e-10 * np.max(np.abs(g['E_query']))
            assert np.max(np.abs(f['Kv'].ravel() - g['Kv'])) < 1e-10 * np.max(np.abs(g['Kv']))
            assert np.array_equal(f['a'], g['alphas_F'])
            assert float(f['c']) == float(g['c']) and float(f['std']) == float(g['std'])
            assert np.max(np.abs(f['F_tp'] - g['F_query'])) < 1e-10 * np.max(np.abs(g['F_query']))
            assert np.max(np.abs(f['E_tp'] - g['E_query'])) < 1e-10 * np.max(np.abs(g['E_query']))
            for n_pts in (8, 7):
                ws = f['ws%d' % n_pts]
                assert np.array_equal(ws[1:4], np.array([3.0, 6.0, 9.0])) and ws[0] == 0 and ws[4] == 0
                assert np.array_equal(ws[5 : 5 + n_pts * 3], np.arange(n_pts * 3) + 100.0) and not ws[5 + n_pts * 3 :].any()
            lev_ref, Pv_ref = _nystroem_reference(g)
            assert np.max(np.abs(f['lev'] - lev_ref)) < 1e-6 * np.max(np.abs(lev_ref))
            assert np.max(np.abs(f['Pv'] - Pv_ref)) < 1e-6 * np.max(np.abs(Pv_ref))
    with np.load(tmp_path / 'r0.npz') as f0, np.load(tmp_path / 'r1.npz') as f1:
        assert np.array_equal(f0['Pv'], f1['Pv']) and np.array_equal(f0['lev'], f1['lev'])  # ranks stay in lockstep


def _nystroem_reference(g):
    from oracle import iterative as oiter

    B = oiter.nystroem_factor(
        g['R_desc'], g['R_d_desc'], g['tril_perms_lin'], int(g['sig']), float(g['lam']), golden_inducing_cols(g)
    )
    return np.einsum('ij,ij->j', B, B), oiter.precon(B, float(g['lam']))(g['v'])


def test_virtual_ranks_match_unsharded():
    """dist.run_steps_virtual (the single-process driver the single-GPU tests use) on 1, 2 and 3 virtual ranks."""
    from sgdml_b200 import dist as sdist

    g = load_golden('n9_m16_s6')
    n_train, dim_i, lam = g['R_desc'].shape[0], 3 * int(g['n_atoms']), float(g['lam'])
    cols = golden_inducing_cols(g)
    lev_ref, Pv_ref = _nystroem_reference(g)
    for world in (1, 2, 3):
        ops = [NumpyNystroemOps(g) for _ in range(world)]
        facs = sdist.run_steps_virtual(
            [sdist.nystroem_factor_steps(ops[r], r, world, n_train, dim_i, cols, lam) for r in range(world)]
        )
        levs = sdist.run_steps_virtual([sdist.lev_scores_steps(ops[r], facs[r][0], len(cols), dim_i) for r in range(world)])
        Pvs = sdist.run_steps_virtual(
            [sdist.precon_apply_steps(ops[r], facs[r][0], len(cols), lam, g['v'], facs[r][1], facs[r][2], dim_i) for r in range(world)]
        )
        for r in range(world):
            assert np.max(np.abs(levs[r] - lev_ref)) < 1e-6 * np.max(np.abs(lev_ref))
            assert np.max(np.abs(Pvs[r] - Pv_ref)) < 1e-6 * np.max(np.abs(Pv_ref))


def test_qr_fallback_matches_cholesky_path():
    """iterative.py:312-322 (inner matrix not positive definite): the oracle follows the reference (R of a
    Householder QR of [K_nm; sqrt(lam) I]); the engine's steps use shifted CholeskyQR3.  Forced on a
    well-conditioned case, both must reproduce the Cholesky path's preconditioner."""
    from oracle import iterative as oiter
    from sgdml_b200 import dist as sdist

    g = load_golden('n9_m16_s6')
    n_train, dim_i, lam = g['R_desc'].shape[0], 3 * int(g['n_atoms']), float(g['lam'])
    cols = golden_inducing_cols(g)
    lev_ref, Pv_ref = _nystroem_reference(g)
    B = oiter.nystroem_factor(g['R_desc'], g['R_d_desc'], g['tril_perms_lin'], int(g['sig']), lam, cols, force_qr=True)
    assert np.max(np.abs(oiter.precon(B, lam)(g['v']) - Pv_ref)) < 1e-6 * np.max(np.abs(Pv_ref))
    for world in (1, 2):
        ops = [NumpyNystroemOps(g) for _ in range(world)]
        for o in ops:
            o.force_qr = True
        facs = sdist.run_steps_virtual(
            [sdist.nystroem_factor_steps(ops[r], r, world, n_train, dim_i, cols, lam) for r in range(world)]
        )
        Pvs = sdist.run_steps_virtual(
            [sdist.precon_apply_steps(ops[r], facs[r][0], len(cols), lam, g['v'], facs[r][1], facs[r][2], dim_i) for r in range(world)]
        )
        levs = sdist.run_steps_virtual([sdist.lev_scores_steps(ops[r], facs[r][0], len(cols), dim_i) for r in range(world)])
        assert np.max(np.abs(Pvs[0] - Pv_ref)) < 1e-6 * np.max(np.abs(Pv_ref))
        assert np.max(np.abs(levs[0] - lev_ref)) < 1e-6 * np.max(np.abs(lev_ref))


def test_model_shards_add_up():
    from oracle import predict as opredict
    from sgdml_b200 import dist as sdist

    g = load_golden('n9_m16_s6')
    model = golden_model(g)
    M = model['R_desc'].shape[1]
    E = np.zeros(len(g['R_query']))
    F = np.zeros_like(g['F_query'])
    for lo, hi in [(0, 5), (5, 6), (6, M)]:
        e, f = opredict.Predictor(sdist.model_shard(model, lo, hi)).predict(g['R_query'])
        E += e
        F += f
    std, c = float(model['std']), float(model['c'])
    assert np.max(np.abs(F * std - g['F_query'])) < 1e-10 * np.max(np.abs(g['F_query']))
    assert np.max(np.abs(E * std + c - g['E_query'])) < 1e-10 * np.max(np.abs(g['E_query']))
