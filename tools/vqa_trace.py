"""Per-launch device times of the single-image VQA prefill (BASELINE configs[1]): every C-ABI call is bracketed by a CUDA
event pair on the launching stream while the GPU is parked behind a spin kernel (so the host is ahead and an interval is
the kernel's own time incl. its launch gap). Prints a table aggregated by (entry, shape).
    python tools/vqa_trace.py [--crops 4] [--side 672]"""
import argparse
import collections
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights, _lib
from phi3_b200.model import Phi3B200
from phi3_b200.processor import Phi3VImageProcessor, hd_geometry
from phi3_b200.api import _row_stats
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--crops', type=int, default=4)
    ap.add_argument('--side', type=int, default=672)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = configs.PHI35_VISION
    model = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
    ip = Phi3VImageProcessor(num_crops=a.crops, device=dev)
    geo = hd_geometry(a.side, a.side, a.crops)
    imgs, ids = bench.make_inputs(1, 1, geo['num_img_tokens'] + 30, geo['num_img_tokens'])
    ids = ids.to(dev)
    img = imgs[0].to(dev)
    pos = torch.nonzero(ids.cpu() < 0)
    sizes = torch.tensor([[geo['H'], geo['W']]])

    def run():
        pv = ip([img])['pixel_values']
        lg, c = model(ids, pixel_values=pv, image_sizes=sizes, positions=pos, max_tokens=128, logits_rows='last')
        tok = _row_stats(model, lg[:, -1, :])['argmax']
        c.release()
        return tok

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    run()
    b0.record()
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print(f'whole call: {a0.elapsed_time(b0):.3f} ms on the device, host issue time {host_ms:.3f} ms, launches {_lib.launches}')

    rec = []
    orig_call, orig_struct = _lib.call, _lib.call_struct

    def key_of(name, args):
        if name == 'p3_gemm':
            return f'p3_gemm M={args[9]} N={args[10]} K={args[11]} epi={args[12]}'
        return name

    def call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *args)
        e1.record()
        rec.append((key_of(name, args), e0, e1))

    def call_struct(name, st, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_struct(name, st, stream)
        e1.record()
        k = f'{name} M={st.M} N={st.N} K={st.K} epi={st.epi}' if hasattr(st, 'epi') else name
        rec.append((k, e0, e1))

    import phi3_b200.model as M, phi3_b200.api as A, phi3_b200.processor as P
    mods = [m for m in (_lib, M, A, P)]
    for m in mods:                                    # the modules bind `call` by name at import
        if hasattr(m, 'call'):
            m.call = call
        if hasattr(m, 'call_struct'):
            m.call_struct = call_struct
    torch.cuda.synchronize()
    torch.cuda._sleep(int(40e6))                      # park ~20 ms so the host runs ahead of the device
    e_first, e_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_first.record()
    run()
    e_last.record()
    torch.cuda.synchronize()
    for m in mods:
        if hasattr(m, 'call'):
            m.call = orig_call
        if hasattr(m, 'call_struct'):
            m.call_struct = orig_struct
    agg = collections.OrderedDict()
    for k, e0, e1 in rec:
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1) * 1e3
    tot = sum(v[1] for v in agg.values())
    print(f'traced: {len(rec)} launches, sum of intervals {tot / 1e3:.3f} ms, first-to-last {e_first.elapsed_time(e_last):.3f} ms (torch ops in between are the difference)')
    print('| entry | launches | total us | share | avg us |\n|---|---|---|---|---|')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| {k} | {n} | {us:.0f} | {100 * us / tot:.1f}% | {us / n:.1f} |')


if __name__ == '__main__':
    main()
