"""A/B of p3_gemm between two builds of the library (ctypes, same process, same box):
    python tools/gemm_ab.py build/r1/libr1.so phi-3-vision-mlx_b200/libphi3b200.so"""
import ctypes as C
import sys
import torch

dev = torch.device('cuda:0')
SHAPES = [(23080, 4096, 1024, 1), (23080, 3072, 1024, 0), (23080, 1024, 4096, 6), (2885, 4096, 1024, 1), (2885, 1024, 4096, 6),
          (2885, 3072, 1024, 0), (2885, 1024, 1024, 6), (16384, 9216, 3072, 0), (16384, 3072, 8192, 3), (787, 9216, 3072, 0),
          (787, 3072, 3072, 3), (787, 16384, 3072, 4), (787, 3072, 8192, 3)]
if len(sys.argv) > 3 and sys.argv[3] == 'small':
    SHAPES = [s for s in SHAPES if s[0] < 4000]
    del sys.argv[3]
_p, _l, _i = C.c_void_p, C.c_int64, C.c_int32


def main():
    libs = []
    for path in sys.argv[1:]:
        L = C.CDLL(path)
        L.p3_gemm.argtypes = [_p, _l, _p, _l, _p, _p, _l, _p, _p, _l, _i, _i, _i, _i, _p]
        L.p3_gemm.restype = C.c_int
        libs.append((path, L))
    st = torch.cuda.current_stream().cuda_stream
    for M, N, K, epi in SHAPES:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02
        out = torch.zeros(M, N // 2 if epi == 4 else N, device=dev, dtype=torch.float32 if epi == 6 else torch.bfloat16)
        bias = torch.zeros(N, device=dev, dtype=torch.bfloat16) if epi in (1, 6) else None
        res = out if epi in (3, 6) else None
        line = f'M={M:6d} N={N:5d} K={K:5d} epi={epi}'
        for rep in range(2):
            for path, L in libs:
                def fn():
                    rc = L.p3_gemm(x.data_ptr(), K, w.data_ptr(), K, None if bias is None else bias.data_ptr(), out.data_ptr(), out.shape[1],
                                   None if res is None else res.data_ptr(), None, M, N, K, epi, 0, st)
                    assert rc == 0
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(30):
                    fn()
                b.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(b) / 30 * 1e3
                line += f' | {path.split("/")[-1][:12]} {us:8.1f} us {2 * M * N * K / us / 1e6:7.0f} TF/s'
        print(line)


if __name__ == '__main__':
    main()
