"""rank-128 update C -= A B^T on an 8192 x 8192 C (the Cholesky trailing-update shape) for ncu.  usage: k128_probe.py [K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussianprocessfundamentals_b200 import engine as eng  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = 8192
a = torch.randn(K, n, dtype=torch.float64, device="cuda")   # column-major n x K
b = torch.randn(K, n, dtype=torch.float64, device="cuda")
c = torch.zeros(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.gemm(4, 0, a, n, b, n, c, n, n, n, K, -1.0, 1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.gemm(4, 0, a, n, b, n, c, n, n, n, K, -1.0, 1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("K", K, "ms", ms, "TF", 2 * n * n * K / ms / 1e9)
