"""Same-box A/B of the decode-path and attention kernels between two builds of the library (ctypes, one process; only entry
points whose C signature is unchanged): python tools/kernel_ab.py build/r1/libr1.so phi-3-vision-mlx_b200/libphi3b200.so"""
import ctypes as C
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import _lib as L0

dev = torch.device('cuda:0')
st = lambda: torch.cuda.current_stream().cuda_stream


def load(path):
    L = C.CDLL(path)
    for name in ('p3_gemm_skinny', 'p3_gemm_skinny_qkv_rope', 'p3_attention_decode', 'p3_attention_prefill', 'p3_rmsnorm', 'p3_layernorm'):
        fn = getattr(L, name)
        fn.argtypes, fn.restype = L0._SIGS[name], C.c_int
    L.p3_attention_decode_workspace.argtypes = [C.c_int] * 5
    L.p3_attention_decode_workspace.restype = C.c_int64
    return L


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
    return a.elapsed_time(b) / n * 1e3


def main():
    libs = [(p.split('/')[-1][:14], load(p)) for p in sys.argv[1:]]
    M, H, I = 8, 3072, 8192
    rows = []
    for name, N, K, epi, norm in [('skinny qkv+norm', 9216, H, 0, True), ('skinny o+resid', H, H, 3, False), ('skinny gate_up+norm+swiglu', 2 * I, H, 4, True),
                                  ('skinny down+resid', H, I, 3, False), ('skinny lm_head+norm', 32064, H, 5, True)]:
        copies = max(2, int(300e6 // (N * K * 2)) + 1)
        W = [torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02 for _ in range(copies)]
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        nw = torch.ones(K, device=dev, dtype=torch.bfloat16)
        No = N // 2 if epi == 4 else N
        out = torch.zeros(M, No, device=dev, dtype=torch.float32 if epi == 5 else torch.bfloat16)
        res = []
        for rep in range(2):
            for tag, L in libs:
                def fn(i):
                    assert L.p3_gemm_skinny(x.data_ptr(), K, nw.data_ptr() if norm else None, 1e-5, W[i % copies].data_ptr(), out.data_ptr(), No,
                                            out.data_ptr() if epi == 3 else None, M, N, K, epi, None, 0, None, None, 0, st()) == 0
                res.append(f'{tag} {timeit(fn):7.2f}')
        rows.append(f'{name:28s} ' + ' | '.join(res))
        del W
    # decode attention, bench shape
    Hh, D = 32, 96
    for B, S in [(8, 2176), (16, 448)]:
        pps = (S + 64) // 64 + 1
        copies = max(2, int(400e6 // (B * pps * 2 * Hh * 64 * D * 2)) + 1)
        pools = [torch.randn(B * pps, 2, Hh, 64, D, device=dev).to(torch.bfloat16) for _ in range(copies)]
        bt = torch.arange(B * pps, dtype=torch.int32, device=dev).reshape(B, pps)
        qkv = torch.randn(B, 3 * Hh * D, device=dev).to(torch.bfloat16)
        out = torch.zeros(B, Hh * D, device=dev, dtype=torch.bfloat16)
        kv0 = torch.zeros(B, dtype=torch.int32, device=dev)
        p = qkv.data_ptr()
        res = []
        for rep in range(2):
            for tag, L in libs:
                def fn(i):
                    assert L.p3_attention_decode(p, p + Hh * D * 2, p + 2 * Hh * D * 2, 3 * Hh * D, 3 * Hh * D, 3 * Hh * D, out.data_ptr(), Hh * D, B, 1, Hh, Hh, D,
                                                 D ** -0.5, S, kv0.data_ptr(), pools[i % copies].data_ptr(), bt.data_ptr(), pps, 1, 1, None, None, None, 0, st()) == 0
                res.append(f'{tag} {timeit(fn):7.2f}')
        rows.append(f'attn_decode B={B} S={S}        ' + ' | '.join(res))
        del pools
    # prefill attention (tcgen05), bench shape and ViT shape
    for B, Hh, D, Lq, causal in [(8, 32, 96, 2048, 1), (40, 16, 64, 577, 0), (1, 32, 96, 787, 1)]:
        qkv = torch.randn(B * Lq, 3 * Hh * D, device=dev).to(torch.bfloat16)
        out = torch.zeros(B * Lq, Hh * D, device=dev, dtype=torch.bfloat16)
        kv0 = torch.zeros(B, dtype=torch.int32, device=dev)
        p = qkv.data_ptr()
        res = []
        for rep in range(2):
            for tag, L in libs:
                def fn(i):
                    assert L.p3_attention_prefill(p, p + Hh * D * 2, p + 2 * Hh * D * 2, 3 * Hh * D, 3 * Hh * D, 3 * Hh * D, out.data_ptr(), Hh * D, B, Lq, Hh, Hh, D,
                                                  D ** -0.5, causal, 0, kv0.data_ptr(), None, None, 0, 1, st()) == 0
                res.append(f'{tag} {timeit(fn, n=20, warm=3):7.2f}')
        rows.append(f'attn_prefill B={B} L={Lq} d={D}   ' + ' | '.join(res))
    # norms
    x = torch.randn(16384, 3072, device=dev).to(torch.bfloat16); g = torch.ones(3072, device=dev, dtype=torch.bfloat16); y = torch.empty_like(x)
    res = []
    for rep in range(2):
        for tag, L in libs:
            res.append(f'{tag} {timeit(lambda i: L.p3_rmsnorm(x.data_ptr(), g.data_ptr(), y.data_ptr(), 16384, 3072, 1e-5, st())):7.2f}')
    rows.append('rmsnorm 16384x3072           ' + ' | '.join(res))
    print('\n'.join(rows))


if __name__ == '__main__':
    main()
