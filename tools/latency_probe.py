"""Latency of small-batch predictions (host NumPy in/out, the MD use case)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgdml_b200
from sgdml_b200 import synth
for wl in ('ethanol', 'aspirin'):
    cfg = synth.CONFIGS[wl]
    perms = synth.rotor_swap_group(cfg['n_atoms'], cfg['n_rotors'], cfg['n_swaps'])
    model = synth.random_model(cfg['n_atoms'], cfg['n_train'], perms, cfg['sig'])
    p = sgdml_b200.GDMLPredict(model)
    for B in (1, 10, 100, 1000):
        R = synth.geometries(cfg['n_atoms'], B, 1).reshape(B, -1)
        for _ in range(5): p.predict(R)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 200
        for _ in range(n): p.predict(R)
        dt = (time.perf_counter() - t0) / n
        print('%s B=%d: %.1f us per call  (%.3g predictions/s)' % (wl, B, dt * 1e6, B / dt), flush=True)
