"""Decode attention alone at the bench shape (B=8, 32 heads, past 2176) for several split-KV factors; 32 distinct layer pools
(6.8 GB) so every launch streams cold pages. python tools/attn_splits.py [--past 2176] [--B 8]"""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument('--past', type=int, default=2176)
ap.add_argument('--B', type=int, default=8)
a = ap.parse_args()
dev = torch.device('cuda:0')
B, H, D, past, NL = a.B, 32, 96, a.past, 32
pages = (past + 64) // 64 + 1
pool = torch.randn(NL, B * pages, 2, H, 64, D, device=dev, dtype=torch.bfloat16)
bt = torch.arange(B * pages, dtype=torch.int32, device=dev).reshape(B, pages)
kvs = torch.zeros(B, dtype=torch.int32, device=dev)
qkv = torch.randn(B, 3 * H * D, device=dev).to(torch.bfloat16)
out = torch.empty(B, H * D, dtype=torch.bfloat16, device=dev)
st = torch.cuda.current_stream().cuda_stream
qp, hb = qkv.data_ptr(), H * D * 2
bytes_per = B * past * 2 * H * D * 2
for ns in (1, 2, 3, 4, 5, 6, 8):
    ws = torch.zeros(max(1, _lib.lib().p3_attention_decode_workspace(B, 1, H, D, ns) // 4), dtype=torch.float32, device=dev)
    def run():
        for li in range(NL):
            _lib.call('p3_attention_decode', qp, qp + hb, qp + 2 * hb, 3 * H * D, 3 * H * D, 3 * H * D, out.data_ptr(), H * D, B, 1, H, H, D,
                      D ** -0.5, past, kvs.data_ptr(), pool[li].data_ptr(), bt.data_ptr(), pages, 1, ns, ws.data_ptr(), None, None, 0, st)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        run()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (4 * NL)
    print(f'n_splits={ns}: {us:6.2f} us  {bytes_per / us / 1e3:7.1f} GB/s  ({B * H * ns} CTAs x {past / 64 / ns:.1f} pages)')
