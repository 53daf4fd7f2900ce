"""Small driver for ncu captures (never used for bench numbers):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --decode-steps 2
One pass of the bench workload per GPU (8 x 2048-token image+text prompts) with eager decode steps."""
import argparse
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights
from phi3_b200.model import Phi3B200
from phi3_b200.processor import Phi3VImageProcessor, hd_geometry
from phi3_b200.api import _row_stats
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--decode-steps', type=int, default=2)
    ap.add_argument('--no-vision', action='store_true')
    ap.add_argument('--batch', type=int, default=bench.B_PER_GPU)
    ap.add_argument('--ctx', type=int, default=bench.CTX)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = configs.PHI35_MINI if a.no_vision else configs.PHI35_VISION
    model = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
    geo = hd_geometry(bench.IMG, bench.IMG, 4)
    imgs, ids = bench.make_inputs(1000, a.batch, a.ctx, geo['num_img_tokens'])
    ids = ids.to(dev)
    kw = {}
    if not a.no_vision:
        ip = Phi3VImageProcessor(num_crops=4, device=dev)
        kw = dict(pixel_values=ip([imgs[i].to(dev) for i in range(a.batch)])['pixel_values'],
                  image_sizes=torch.tensor([[geo['H'], geo['W']]] * a.batch), positions=torch.nonzero(ids.cpu() < 0))
    else:
        ids = ids.clamp(min=3)
    logits, cache = model(ids, max_tokens=bench.NEW, logits_rows='last', **kw)
    tok = _row_stats(model, logits[:, -1, :])['argmax']
    hist = model.greedy_decode(tok, cache, a.decode_steps, use_graph=False)
    torch.cuda.synchronize()
    print('done', hist.shape)


if __name__ == '__main__':
    main()
