"""Per-call timing of GDMLPredict.predict (device-resident inputs): GPU time (events) and host time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgdml_b200
from sgdml_b200 import synth, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else 'ethanol'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
cfg = synth.CONFIGS[wl]
perms = synth.rotor_swap_group(cfg['n_atoms'], cfg['n_rotors'], cfg['n_swaps'])
model = synth.random_model(cfg['n_atoms'], cfg['n_train'], perms, cfg['sig'])
p = sgdml_b200.GDMLPredict(model)
R = torch.from_numpy(synth.geometries(cfg['n_atoms'], B, 1).reshape(B, -1)).cuda()
for _ in range(3): p.predict(R)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
host = []
evs[0].record()
for i in range(10):
    t0 = time.perf_counter(); p.predict(R); host.append((time.perf_counter() - t0) * 1e3); evs[i + 1].record()
torch.cuda.synchronize()
print(wl, 'B', B, 'gpu ms per call', ['%.2f' % evs[i].elapsed_time(evs[i + 1]) for i in range(10)])
print('host ms per call', ['%.2f' % h for h in host])
out = (torch.empty(B, dtype=torch.float64, device='cuda'), torch.empty((B, R.shape[1]), dtype=torch.float64, device='cuda'))
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10): p.predict(R, out=out)
torch.cuda.synchronize(); print('with out=: ms per call %.3f' % ((time.perf_counter() - t0) * 100))
