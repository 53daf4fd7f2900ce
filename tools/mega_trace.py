"""Phase timeline of the persistent decode-layer kernel from its own SM-clock stamps (run on the GPU box):
    python tools/mega_trace.py [--B 8]
Prints, per phase of a 'mid' launch (o_proj, gate_up, down, next qkv), medians / maxima over the CTAs of: barrier wait,
X-fragment load, streaming time, and the cycles warps 0 / 7 spent waiting for weight data."""
import argparse
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['P3_MEGA'] = '1'
import phi3_b200  # noqa
from phi3_b200 import configs, weights
from phi3_b200.model import Phi3B200, DecodeSession


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=8)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = configs.PHI35_MINI
    w = weights.random_weights(cfg, seed=0, device=dev)
    m = Phi3B200(cfg, w, device=dev)
    del w
    ids = torch.randint(3, 32000, (a.B, 64))
    lg, c = m(ids, max_tokens=8, logits_rows='last')
    ses = DecodeSession(m, lg[:, -1].argmax(-1), c, 4, use_graph=False)
    ms = ses.mega_ses
    st = torch.cuda.current_stream().cuda_stream
    n = ms.mega.n_ctas
    dbg = torch.zeros(n, 4, 8, dtype=torch.int64, device=dev)
    for li in range(1, 20):
        ms.launch(li, st)
    torch.cuda.synchronize()
    ms.args[20].dbg = dbg.data_ptr()
    for li in range(18, 21):
        ms.launch(li, st)
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    ghz = 1.0e-3 * torch.cuda.clock_rate() if hasattr(torch.cuda, 'clock_rate') else 1.9
    out = {'clock_ghz_assumed': ghz, 'phases': []}
    t0 = d[:, 0, 0]
    for p, name in enumerate(['o_proj', 'gate_up', 'down', 'qkv']):
        s = d[:, p]
        row = {'phase': name,
               'start_us_med': ((s[:, 0] - t0) / ghz / 1e3).median().item(),
               'barrier_wait_us_med': ((s[:, 1] - s[:, 0]) / ghz / 1e3).median().item(),
               'barrier_wait_us_max': ((s[:, 1] - s[:, 0]) / ghz / 1e3).max().item(),
               'barrier_wait_us_min': ((s[:, 1] - s[:, 0]) / ghz / 1e3).min().item(),
               'xload_us_med': ((s[:, 2] - s[:, 1]) / ghz / 1e3)[s[:, 2] > 0].median().item(),
               'stream_us_med': ((s[:, 3] - s[:, 2]) / ghz / 1e3)[s[:, 3] > 0].median().item(),
               'stream_us_max': ((s[:, 3] - s[:, 2]) / ghz / 1e3)[s[:, 3] > 0].max().item(),
               'data_wait_us_w0_med': (s[:, 4] / ghz / 1e3)[s[:, 3] > 0].median().item(),
               'data_wait_us_w7_med': (s[:, 5] / ghz / 1e3)[s[:, 3] > 0].median().item()}
        out['phases'].append({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()})
    out['total_us_med'] = round(((d[:, 3, 3] - t0) / ghz / 1e3).median().item(), 2)
    # per-CTA detail grouped by the number of tiles the CTA owns in the phase
    out['by_tiles'] = []
    for p, name in enumerate(['o_proj', 'gate_up', 'down', 'qkv']):
        off = ms.mega.sched['mid'][p][0].cpu()
        cnt = (off[1:] - off[:-1])
        s = d[:, p]
        for k in sorted(set(cnt.tolist())):
            sel = cnt == k
            out['by_tiles'].append({'phase': name, 'tiles': k, 'ctas': int(sel.sum()),
                                    'arrive_us': round(((s[sel, 0] - t0[sel]) / ghz / 1e3).mean().item(), 2),
                                    'released_us': round(((s[sel, 1] - t0[sel]) / ghz / 1e3).mean().item(), 2),
                                    'x_ready_us': round(((s[sel, 2] - t0[sel]) / ghz / 1e3).mean().item(), 2),
                                    'done_us': round(((s[sel, 3] - t0[sel]) / ghz / 1e3).mean().item(), 2),
                                    'stream_us': round(((s[sel, 3] - s[sel, 2]) / ghz / 1e3).mean().item(), 2),
                                    'data_wait_w0_us': round((s[sel, 4] / ghz / 1e3).mean().item(), 2),
                                    'item_sync_wait_w0_us': round((s[sel, 6] / ghz / 1e3).mean().item(), 2),
                                    'epilogue_w0_us': round((s[sel, 7] / ghz / 1e3).mean().item(), 2)})
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
