"""Microbenchmark of cb200_colsum on the bias-gradient shapes of the b512 step (D-step batch 1536)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrad_b200 import kernels as K
shapes = [(1536 * 1024, 64), (1536 * 256, 64), (1536 * 256, 128), (1536 * 64, 128), (1536 * 64, 256), (1536 * 16, 256),
          (1536 * 16, 512), (512 * 1024, 64), (512 * 256, 128), (512 * 64, 256), (1536, 1024), (1536, 128)]
tot = 0.0
for M, N in shapes:
    x = torch.randn(M, N, device="cuda")
    for _ in range(3): K.colsum(x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): K.colsum(x)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    tot += us
    print("colsum M=%8d N=%4d  %8.1f us  %7.1f GB/s" % (M, N, us, M * N * 4 / us / 1e3))
print("total %.1f us" % tot)
