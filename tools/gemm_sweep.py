import os, sys, torch
sys.path.insert(0, '/root/repo')
import phi3_b200
from phi3_b200 import _lib
sys.path.insert(0, '/root/repo/tools')
from microbench import timeit, dev, st
for N, K in [(9216, 3072), (3072, 8192)]:
    w = torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02
    for M in [17, 64, 128, 129, 256, 512, 781, 1024, 2048, 4096]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda i: _lib.call('p3_gemm', x.data_ptr(), K, w.data_ptr(), K, None, out.data_ptr(), N, None, None, M, N, K, 0, 0, st())
        us = timeit(fn, n=30, warm=5)
        print(f'N={N} K={K} M={M:5d} {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TF/s  W-stream {N*K*2/us/1e3:6.0f} GB/s')
