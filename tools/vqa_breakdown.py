"""Where the single-image VQA prefill time goes (BASELINE configs[1]): wall vs GPU-busy, per phase."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phi3_b200  # noqa
from phi3_b200 import configs, weights, _lib
from phi3_b200.model import Phi3B200
from phi3_b200.processor import Phi3VImageProcessor, hd_geometry
from phi3_b200.api import _row_stats
import bench
dev = torch.device('cuda:0')
cfg = configs.PHI35_VISION
model = Phi3B200(cfg, weights.random_weights(cfg, seed=0, device=dev), device=dev)
ip = Phi3VImageProcessor(num_crops=4, device=dev)
geo = hd_geometry(672, 672, 4)
imgs, ids = bench.make_inputs(1, 1, geo['num_img_tokens'] + 30, geo['num_img_tokens'])
ids = ids.to(dev); img = imgs[0].to(dev)
pos = torch.nonzero(ids.cpu() < 0); sizes = torch.tensor([[geo['H'], geo['W']]])
def run():
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = _lib.launches
    pv = ip([img])['pixel_values']
    t['hd_cpu'] = time.perf_counter() - t0; torch.cuda.synchronize(); t['hd'] = time.perf_counter() - t0
    t1 = time.perf_counter()
    h = torch.empty((ids.numel(), model.H), dtype=torch.bfloat16, device=dev)
    _lib.call('p3_embed_gather', model.embed.data_ptr(), ids.to(torch.int32).reshape(-1).data_ptr(), h.data_ptr(), ids.numel(), model.H, model.V, None, torch.cuda.current_stream().cuda_stream)
    h = model._vision_embed(h, ids.shape[1], pv, sizes, pos)
    t['vit_cpu'] = time.perf_counter() - t1; torch.cuda.synchronize(); t['vit'] = time.perf_counter() - t1
    t2 = time.perf_counter()
    lg, c = model(ids, pixel_values=None, max_tokens=128, logits_rows='last')
    tok = _row_stats(model, lg[:, -1, :])['argmax']
    t['llm_cpu'] = time.perf_counter() - t2; torch.cuda.synchronize(); t['llm'] = time.perf_counter() - t2
    t['launches'] = _lib.launches - n0
    return t
for i in range(4):
    t = run()
print({k: (round(v * 1e3, 2) if isinstance(v, float) else v) for k, v in t.items()})
