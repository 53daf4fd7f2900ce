"""Kernel micro-benchmarks for tuning (not the bench contract): each decode-path kernel timed with
CUDA events over back-to-back launches, rotating over enough weight / KV copies to defeat the L2.
    python tools/microbench.py [skinny] [attn] [gemm]
Env knobs: P3_SK_DEPTH, P3_PDL."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import _lib
from phi3_b200.model import interleave_gate_up, pick_splits

dev = torch.device('cuda:0')
st = lambda: torch.cuda.current_stream().cuda_stream
PEAK = 6462.7


def timeit(fn, n=60, warm=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3      # us


def bench_skinny():
    M, H, I = 8, 3072, 8192
    shapes = [('qkv+norm', 9216, H, 0, True), ('o+resid', H, H, 3, False), ('gate_up+norm+swiglu', 2 * I, H, 4, True),
              ('down+resid', H, I, 3, False), ('lm_head+norm', 32064, H, 5, True)]
    for name, N, K, epi, norm in shapes:
        copies = max(2, int(300e6 // (N * K * 2)) + 1)
        W = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(copies)]
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        nw = torch.ones(K, device=dev, dtype=torch.bfloat16)
        No = N // 2 if epi == 4 else N
        out = torch.zeros(M, No, device=dev, dtype=torch.float32 if epi == 5 else torch.bfloat16)

        def fn(i):
            _lib.call('p3_gemm_skinny', x.data_ptr(), K, nw.data_ptr() if norm else None, 1e-5, W[i % copies].data_ptr(),
                      out.data_ptr(), No, out.data_ptr() if epi == 3 else None, M, N, K, epi, None, 0, None, None, 0, st())
        us = timeit(fn)
        gbs = N * K * 2 / us / 1e3
        print(f'skinny {name:22s} N={N:6d} K={K:5d}  {us:7.2f} us  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f}% of measured HBM peak')


def bench_skinny_w4():
    """quantize_model decode GEMMs: 4-bit g64 codes (0.5625 B per weight incl. scale/bias) vs the bf16 stream"""
    from phi3_b200 import quant
    M, H, I = 8, 3072, 8192
    shapes = [('qkv+norm', 9216, H, 0, True), ('qkv', 9216, H, 0, False), ('o+resid', H, H, 3, False), ('gate_up+norm+swiglu', 2 * I, H, 4, True),
              ('down+resid', H, I, 3, False), ('lm_head+norm', 32064, H, 5, True)]
    for name, N, K, epi, norm in shapes:
        copies = max(2, int(300e6 // (N * K // 2)) // 4 + 1)
        Q = [quant.W4(torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02) for _ in range(copies)]
        for q in Q:
            q.deq = None
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        nw = torch.ones(K, device=dev, dtype=torch.bfloat16)
        No = N // 2 if epi == 4 else N
        out = torch.zeros(M, No, device=dev, dtype=torch.float32 if epi == 5 else torch.bfloat16)

        def fn(i):
            q = Q[i % copies]
            _lib.call('p3_gemm_skinny_w4', x.data_ptr(), K, nw.data_ptr() if norm else None, 1e-5, q.codes.data_ptr(), q.meta.data_ptr(),
                      out.data_ptr(), No, out.data_ptr() if epi == 3 else None, M, N, K, epi, None, 0, None, None, 0, st())
        us = timeit(fn)
        nbytes = N * K // 2 + N * (K // 64) * 4
        gbs = nbytes / us / 1e3
        print(f'skinny_w4 {name:22s} N={N:6d} K={K:5d}  {us:7.2f} us  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f}% of measured HBM peak')


def bench_attn():
    H, D = 32, 96
    for B, S in [(8, 2176), (8, 512), (1, 2176), (16, 400), (4, 8192)]:
        pps = (S + 64) // 64 + 1
        copies = max(2, int(400e6 // (B * pps * 2 * H * 64 * D * 2)) + 1)
        pools = [torch.randn(B * pps, 2, H, 64, D, device=dev).to(torch.bfloat16) for _ in range(copies)]
        bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
        qkv = torch.randn(B, 3 * H * D, device=dev).to(torch.bfloat16)
        out = torch.zeros(B, H * D, device=dev, dtype=torch.bfloat16)
        kv0 = torch.zeros(B, dtype=torch.int32, device=dev)
        for ns in sorted({1, pick_splits(B * H, (S + 63) // 64), 4, 8, 16}):
            ws = torch.zeros(_lib.lib().p3_attention_decode_workspace(B, 1, H, D, ns) // 4, device=dev)
            p = qkv.data_ptr()

            def fn(i):
                _lib.call('p3_attention_decode', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D,
                          out.data_ptr(), H * D, B, 1, H, H, D, D ** -0.5, S, kv0.data_ptr(), pools[i % copies].data_ptr(),
                          bt.data_ptr(), pps, 1, ns, ws.data_ptr(), None, None, 0, st())
            us = timeit(fn)
            gbs = B * S * 2 * H * D * 2 / us / 1e3
            print(f'attn_decode B={B:2d} S={S:5d} splits={ns:2d} ({B * H * ns:5d} CTAs)  {us:7.2f} us  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f}%')


def bench_attn_q4():
    H, D = 32, 96
    for B, S in [(8, 2176), (16, 448)]:
        pps = (S + 64) // 64 + 1
        n_quant = (S // 64) * 64
        pool = torch.randn(B * pps, 2, H, 64, D, device=dev).to(torch.bfloat16)
        qc = torch.zeros(B * pps, 2, H, 64, D // 2, dtype=torch.uint8, device=dev)
        qm = torch.zeros(B * pps, 2, H, 64, D // 32, 2, dtype=torch.bfloat16, device=dev)
        bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
        _lib.call('p3_kv_quantize_q4g32', pool.data_ptr(), qc.data_ptr(), qm.data_ptr(), bt.data_ptr(), pps, B, S, H, D, st())
        qkv = torch.randn(B, 3 * H * D, device=dev).to(torch.bfloat16)
        out = torch.zeros(B, H * D, device=dev, dtype=torch.bfloat16)
        kv0 = torch.zeros(B, dtype=torch.int32, device=dev)
        for ns in sorted({1, 2, 4} if not os.environ.get('P3_Q4_SPLITS') else {int(os.environ['P3_Q4_SPLITS'])}):
            ws = torch.zeros(_lib.lib().p3_attention_decode_workspace(B, 1, H, D, ns) // 4, device=dev)
            p = qkv.data_ptr()

            def fn(i):
                _lib.call('p3_attention_decode_q4', p, p + H * D * 2, p + 2 * H * D * 2, 3 * H * D, 3 * H * D, 3 * H * D,
                          out.data_ptr(), H * D, B, 1, H, H, D, D ** -0.5, S, n_quant, kv0.data_ptr(), pool.data_ptr(),
                          qc.data_ptr(), qm.data_ptr(), bt.data_ptr(), pps, 1, ns, ws.data_ptr(), None, None, 0, st())
            us = timeit(fn)
            q4_bytes = B * n_quant * 2 * H * (D // 2 + (D // 32) * 4) + B * (S - n_quant) * 2 * H * D * 2
            print(f'attn_decode_q4 B={B:2d} S={S:5d} splits={ns:2d}  {us:7.2f} us  {q4_bytes / us / 1e3:7.0f} GB/s of q4 bytes '
                  f'({100 * q4_bytes / us / 1e3 / PEAK:5.1f}%)  = {B * S * 2 * H * D * 2 / us / 1e3:7.0f} GB/s bf16-equivalent')


def bench_gemm():
    for M, N, K, epi in [(16384, 9216, 3072, 0), (16384, 3072, 3072, 3), (16384, 16384, 3072, 4), (16384, 3072, 8192, 3),
                         (23080, 3072, 1024, 0), (23080, 4096, 1024, 1), (23080, 1024, 4096, 6), (2885, 3072, 1024, 0),
                         (781, 9216, 3072, 0), (781, 16384, 3072, 4)]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02
        No = N // 2 if epi == 4 else N
        out = torch.zeros(M, No, device=dev, dtype=torch.float32 if epi == 6 else torch.bfloat16)
        bias = torch.zeros(N, device=dev, dtype=torch.bfloat16) if epi in (1, 6) else None

        def fn(i):
            _lib.call('p3_gemm', x.data_ptr(), K, w.data_ptr(), K, None if bias is None else bias.data_ptr(), out.data_ptr(), No,
                      out.data_ptr() if epi in (3, 6) else None, None, M, N, K, epi, 0, st())
        us = timeit(fn, n=20, warm=3)
        tf = 2 * M * N * K / us / 1e6
        print(f'gemm_tc M={M:6d} N={N:6d} K={K:5d} epi={epi}  {us:9.1f} us  {tf:7.1f} TFLOP/s  {100 * tf / 1670.7:5.1f}% of measured bf16 peak')


if __name__ == '__main__':
    which = sys.argv[1:] or ['skinny', 'attn', 'gemm']
    if 'skinny' in which:
        bench_skinny()
    if 'attn' in which:
        bench_attn()
    if 'w4' in which:
        bench_skinny_w4()
    if 'gemm' in which:
        bench_gemm()
    if 'q4' in which:
        bench_attn_q4()
