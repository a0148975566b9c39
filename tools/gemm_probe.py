"""Runs one GEMM shape through sealdec_debug_gemm (for ncu captures / quick timing)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seal_b200._lib import lib, check
mode, M, N, K = (int(x) for x in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)).astype(np.float32); W = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
b = np.zeros(N, dtype=np.float32); out = np.empty((M, N), dtype=np.float32); us = C.c_double(0)
check(lib.sealdec_debug_gemm(mode, M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data, out.ctypes.data, 0, iters, C.byref(us)))
print(f"mode {mode} {M}x{N}x{K}: {us.value:.1f} us  {2.0*M*N*K/us.value/1e6:.1f} TFLOP/s")
