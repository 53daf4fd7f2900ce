"""Small-M tensor-core GEMM: with / without split-K (p3_gemm_fused, rotating weights so every launch streams from HBM).
    python tools/gemm_smallm.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import _lib
dev = torch.device('cuda:0')
st = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.zeros(32 << 20, dtype=torch.uint8, device=dev)
NW = 24                                     # rotate over 24 weight copies (> L2)
for N, K, epi in [(9216, 3072, 0), (3072, 3072, 3), (16384, 3072, 4), (3072, 8192, 3)]:
    Ws = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(NW)]
    for M in [16, 80, 128, 320, 781]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        No = N // 2 if epi == 4 else N
        out = torch.zeros(M, No, device=dev, dtype=torch.bfloat16)
        res = {}
        for split in (0, 1):
            args = []
            for w in Ws:
                a = _lib.GemmArgs()
                a.X, a.ldx, a.W, a.ldw, a.out, a.ldo = x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), No
                a.M, a.N, a.K, a.epi, a.impl = M, N, K, epi, 0
                if epi == 3:
                    a.resid = out.data_ptr()
                if split:
                    a.splitk_ws, a.splitk_ws_bytes = ws.data_ptr(), ws.numel()
                args.append(a)
            for a in args[:4]:
                _lib.call_struct('p3_gemm_fused', a, st())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(3):
                for a in args:
                    _lib.call_struct('p3_gemm_fused', a, st())
            e1.record()
            torch.cuda.synchronize()
            res[split] = 1e3 * e0.elapsed_time(e1) / (3 * NW)
        print(f'N={N:5d} K={K} M={M:4d} epi={epi}: no split {res[0]:7.1f} us ({N*K*2/res[0]/1e3:5.0f} GB/s)   split-K {res[1]:7.1f} us ({N*K*2/res[1]/1e3:5.0f} GB/s)')
