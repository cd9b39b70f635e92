"""Timing experiments on the int8-slice GEMM: which part of a tile costs what (SGDML_B200_OZAKI_DBG flags)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgdml_b200 import _lib
L = _lib.lib()
S = int(os.environ.get('OZ_S', '7'))
m = n = int(os.environ.get('OZ_N', '8192'))
for k in (128, 1024):
    A = torch.randn(m, k, dtype=torch.float64, device='cuda')
    B = torch.randn(n, k, dtype=torch.float64, device='cuda')
    C = torch.zeros(m, n, dtype=torch.float64, device='cuda')
    for rep in range(2):
        L.sgdml_b200_profile_reset(); L.sgdml_b200_profile_enable(1)
        rc = L.sgdml_b200_ozaki_gemm_nt(m, n, k, 1.0, A.data_ptr(), k, B.data_ptr(), k, C.data_ptr(), n, S, 0, None)
        torch.cuda.synchronize(); L.sgdml_b200_profile_enable(0)
        assert rc == 0, _lib.last_error()
        ms = ctypes.c_double(); sc = ctypes.c_int64(); ln = ctypes.c_int64()
        L.sgdml_b200_profile_get(3, ctypes.byref(ms), ctypes.byref(sc), ctypes.byref(ln))
    tiles = (m // 128) * (n // 64)
    print('dbg=%s bk=%s S=%d m=n=%d k=%d: %.3f ms, %.2f us per tile-wave (148 SMs), %.1f TF/s eq' % (
        os.environ.get('SGDML_B200_OZAKI_DBG', '0'), os.environ.get('SGDML_B200_OZAKI_BK', '64'), S, m, k, ms.value,
        ms.value * 1e3 / (tiles / 148.0), 2.0 * m * n * k / ms.value * 1e-9), flush=True)
