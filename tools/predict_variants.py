"""Device-resident throughput of the fused predictor for the main-kernel variants (0 = default, 2 = no split over k,
3 = 2 + one barrier per tile; see include/sgdml_b200.h) and B = 1 latency: CUDA-graph replay with zero-copy host
buffers, graph replay with copy nodes, plain launches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgdml_b200
from sgdml_b200 import synth, _lib
L = _lib.lib()
peak = __import__('ctypes').c_double()
L.sgdml_b200_fp64_peak_tflops(__import__('ctypes').byref(peak))
shapes = [('aspirin', None), ('ethanol', None)]
# one synthetic molecule per remaining tile configuration: DP = 72 (N = 12), 112 (N = 15), 160 (N = 18)
shapes += [('N%d' % N_, dict(n_atoms=N_, n_train=1000, n_rotors=1, n_swaps=1, sig=20)) for N_ in (12, 15, 18)]
if os.environ.get('VARIANT_SHAPES'):
    shapes = [s_ for s_ in shapes if s_[0] in os.environ['VARIANT_SHAPES'].split(',')]
B = 65536
for wl, cfg in shapes:
    cfg = cfg or synth.CONFIGS[wl]
    N, M = cfg['n_atoms'], cfg['n_train']
    perms = synth.rotor_swap_group(N, cfg['n_rotors'], cfg['n_swaps'])
    S, D = len(perms), N * (N - 1) // 2
    model = synth.random_model(N, M, perms, cfg['sig'])
    p = sgdml_b200.GDMLPredict(model)
    R = torch.from_numpy(synth.geometries(N, B, 1).reshape(B, -1)).cuda()
    ref = None
    for variant in ((4, 5, 0, 4, 5, 0) if N <= 9 else (4, 2, 3, 0, 4, 2, 3, 0)):
        L.sgdml_b200_set_predict_variant(variant)
        for _ in range(3): E, F = p.predict(R)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): E, F = p.predict(R)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None: ref = F.clone()
        dev = float((F - ref).abs().max() / ref.abs().max())
        print('%s B=%d variant %d: %.3f ms/call, %.3e pred/s, %.1f TF/s algorithmic (%.2f of %.1f), max rel dev vs variant 4: %.1e' % (
            wl, B, variant, ms, B / ms * 1e3, 9.0 * M * S * D * B / ms * 1e-9, 9.0 * M * S * D * B / ms * 1e-9 / peak.value, peak.value, dev), flush=True)
    L.sgdml_b200_set_predict_variant(0)
    R1 = synth.geometries(N, 1, 1).reshape(1, -1)
    for name, env in (('graph replay, zero-copy', {'SGDML_B200_GRAPH': '1', 'SGDML_B200_GRAPH_ZEROCOPY': '1'}),
                      ('graph replay, copy nodes', {'SGDML_B200_GRAPH': '1', 'SGDML_B200_GRAPH_ZEROCOPY': '0'}),
                      ('plain launches', {'SGDML_B200_GRAPH': '0'})):
        os.environ.update(env)
        p1 = sgdml_b200.GDMLPredict(model)  # a fresh handle: the graph cache is per model
        for _ in range(20): p1.predict(R1)
        t0 = time.perf_counter()
        for _ in range(1000): p1.predict(R1)
        print('%s B=1 host NumPy in/out, %s: %.1f us per call' % (wl, name, (time.perf_counter() - t0) / 1000 * 1e6), flush=True)
    os.environ.pop('SGDML_B200_GRAPH', None)
    os.environ.pop('SGDML_B200_GRAPH_ZEROCOPY', None)
