"""Times the K assembly (column subset, as the Nystroem set-up uses it) for large molecules."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit('/', 2)[0])
import torch  # noqa: E402

import sgdml_b200  # noqa: E402
from sgdml_b200 import synth  # noqa: E402
from sgdml_b200.desc import Desc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n-atoms', type=int, default=100)
ap.add_argument('--n-train', type=int, default=256)
ap.add_argument('--col-points', type=int, default=8)
ap.add_argument('--rotors', type=int, default=2)
ap.add_argument('--swaps', type=int, default=0)
ap.add_argument('--sig', type=float, default=50)
a = ap.parse_args()

N, M = a.n_atoms, a.n_train
perms = synth.rotor_swap_group(N, a.rotors, a.swaps)
S = len(perms)
R = synth.geometries(N, M, 0).reshape(M, -1)
d = Desc(N)
x, g = d.from_R(R)
lin = sgdml_b200.desc.tril_perms_lin(perms)
t = sgdml_b200.GDMLTrain()
cols = np.arange(a.col_points * 3 * N, dtype=np.int64)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K, nc = t._assemble_kernel_mat_device(x, g, lin, a.sig, col_idxs=cols)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
blocks = M * a.col_points
print('N=%d M=%d S=%d col_points=%d: %.3f s, %.1f us/block, %.2f GB written (%.1f GB/s)' % (
    N, M, S, a.col_points, dt, dt / blocks * 1e6, K.numel() * 8 / 1e9, K.numel() * 8 / 1e9 / dt))
