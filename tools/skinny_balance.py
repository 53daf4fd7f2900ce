"""Does the skinny weight stream need every SM loaded equally? Same kernel (16-row tiles, no epilogue), N swept so the grid is
0.65 ... 4 CTAs per SM slot pair; GB/s over back-to-back PDL launches on rotating weights.
    python tools/skinny_balance.py"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import _lib
from microbench import timeit, dev, st, PEAK

M = 8
for K in (8192, 3072):
    for ctas in (148, 192, 222, 296, 384, 444, 512, 592, 888, 1184):
        N = ctas * 16
        copies = max(2, int(400e6 // (N * K * 2)) + 1)
        W = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(copies)]
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)

        def fn(i):
            _lib.call('p3_gemm_skinny', x.data_ptr(), K, None, 1e-5, W[i % copies].data_ptr(), out.data_ptr(), N, None, M, N, K, 0,
                      None, 0, None, None, 0, st())
        us = timeit(fn, n=120, warm=20)
        gbs = N * K * 2 / us / 1e3
        print(f'K={K:5d} CTAs={ctas:5d} ({ctas / 148:4.2f}/SM)  {us:7.2f} us  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f}%', flush=True)
        del W
