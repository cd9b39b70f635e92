"""Multi-GPU plumbing for the two paths that shard: one process per
GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

* prediction over a large batch: queries are split across ranks, the model is replicated,
  there is NO data-path collective (outputs are gathered only if the caller wants them on
  every rank) -- what the reference does with ``torch.nn.DataParallel`` (predict.py:375-378);
* the K.v operator of the iterative solver: each rank evaluates a contiguous range of
  training points, then ONE all-gather of 3N*M/G doubles per iteration (iterative.py:183-204
  is the single-process operator);
* model hand-over after training on rank 0: ONE broadcast of alphas_F (+ c, std).

The compute callables are injected, so the same code runs against the CUDA engine (NCCL) and,
in the CPU tests, against the oracle (gloo).
"""

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n, world, rank):
    """Contiguous balanced shard [lo, hi) of range(n): the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device_for_backend(group=None):
    import torch

    dist = _dist()
    backend = dist.get_backend(group)
    return torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')


def all_gather_rows(local, n_total, group=None):
    """Gathers row-sharded arrays (shards as produced by shard_bounds) into the full array on
    every rank.  One all_gather; shards are padded to the largest shard."""
    import torch

    dist = _dist()
    rank, world = world_info(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    dev = _device_for_backend(group)
    tail = local.shape[1:]
    max_rows = shard_bounds(n_total, world, 0)[1]
    buf = torch.zeros((max_rows,) + tail, dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        parts.append(out[r][: hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)


def predict_sharded(predict_fn, R, gather=True, group=None):
    """predict_fn(R_shard) -> (E, F).  Splits the query batch across ranks; with gather=True every
    rank receives the full (E, F), otherwise each rank keeps (lo, hi, E_local, F_local)."""
    rank, world = world_info(group)
    R = np.asarray(R, dtype=np.float64)
    n = R.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    E, F = predict_fn(R[lo:hi])
    if not gather:
        return lo, hi, E, F
    return all_gather_rows(E, n, group), all_gather_rows(F, n, group)


def kmatvec_sharded(rows_fn, n_train, group=None):
    """rows_fn(m_lo, m_hi) -> (m_hi - m_lo, 3N) raw force sums for training points [m_lo, m_hi)
    (GDMLPredict.kmatvec_train).  Returns the full (n_train, 3N) product on every rank after one
    all-gather."""
    rank, world = world_info(group)
    lo, hi = shard_bounds(n_train, world, rank)
    return all_gather_rows(rows_fn(lo, hi), n_train, group)


def broadcast_coefficients(alphas_F, c, std, src=0, group=None):
    """One broadcast of [alphas_F, c, std] from `src`; returns (alphas_F, c, std) on every rank.
    On ranks != src, `alphas_F` only needs the right length."""
    import torch

    dist = _dist()
    rank, world = world_info(group)
    a = np.ascontiguousarray(alphas_F, dtype=np.float64)
    if world == 1:
        return a, float(c), float(std)
    dev = _device_for_backend(group)
    buf = torch.empty(a.size + 2, dtype=torch.float64, device=dev)
    if rank == src:
        buf[: a.size] = torch.from_numpy(a).to(dev)
        buf[a.size] = float(c)
        buf[a.size + 1] = float(std)
    dist.broadcast(buf, src=src, group=group)
    host = buf.cpu().numpy()
    return host[: a.size].copy(), float(host[a.size]), float(host[a.size + 1])
