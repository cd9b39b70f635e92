"""CPU, world_size 2, gloo: the N>1 host logic (query sharding, K.v row sharding + all-gather,
coefficient broadcast) with the ORACLE as the injected compute function."""

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
    np.savez(os.path.join(out_dir, 'r%d.npz' % rank), E=E, F=F, Kv=Kv, a=a, c=c, std=std)
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
