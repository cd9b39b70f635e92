"""Burst vs sustained FP64 (DMMA) peak with clocks sampled during the sustained run, and clocks during potrf."""
import ctypes, json, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgdml_b200 import _lib
import torch
L = _lib.lib()
def sample(tag, fn):
    rows = []
    p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown', '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
    t = threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True); t.start()
    time.sleep(0.3); r = fn(); time.sleep(0.2); p.terminate()
    print(tag, r, 'clock/power samples:', rows[::3][:14], flush=True)
v = ctypes.c_double()
L.sgdml_b200_fp64_peak_tflops(ctypes.byref(v)); print('burst TF', v.value)
def sus():
    L.sgdml_b200_fp64_peak_tflops_sustained(3.0, ctypes.byref(v)); return v.value
sample('sustained 3s TF', sus)
import numpy as np
n = 32768
A = torch.randn(n, n, dtype=torch.float64, device='cuda'); A = A @ A.T; A += n * torch.eye(n, dtype=torch.float64, device='cuda')
torch.cuda.synchronize()
def fac():
    t0 = time.perf_counter(); rc = L.sgdml_b200_potrf(A.data_ptr(), n, n, None); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return {'rc': rc, 's': dt, 'TF': n ** 3 / 3 / dt * 1e-12}
sample('potrf n=32768', fac)
