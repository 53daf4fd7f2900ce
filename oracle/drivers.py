"""CPU ORACLE (test infrastructure only) — restatement of the reference decode drivers on the
oracle model: _generate (phi_3_vision_mlx.py:376-409), _choose_from (pv:466-487),
_constrain + _get_beam (pv:500-619). Dense log-softmax over the full vocabulary, exactly as the
reference does it; token-id level only (tokenizer decode is host glue). Parity unpinned by the
reference's own tests — these follow the cited lines statement by statement.
"""
import torch

ID_EOS = 32007


class LogitStopper:
    """pv:79-104, statement by statement (B = 1 early-stop heuristic on the EOS log-probability)."""

    def __init__(self, max_tokens, early_stop):
        self.step = 0
        self.early_stop = early_stop if isinstance(early_stop, int) and (early_stop < max_tokens) else False
        self.log_prob_sum = 0.0
        self.best_eos_sofar = -float('inf')
        self.log_prob_sum_at_best_eos = 0.0

    def __call__(self, logits):
        if not self.early_stop:
            return False
        if logits.shape[0] > 1:
            self.early_stop = False
            return False
        log_prob = torch.log_softmax(logits[:, -1, :], -1)
        log_prob_best = log_prob.max(-1).values.item()
        log_prob_eos = log_prob[:, ID_EOS].item()
        if log_prob_eos > self.best_eos_sofar:
            since = self.log_prob_sum - self.log_prob_sum_at_best_eos
            if (since < self.best_eos_sofar) and (self.step > self.early_stop):
                return True
            self.best_eos_sofar = log_prob_eos
            self.log_prob_sum_at_best_eos = self.log_prob_sum
        self.log_prob_sum += log_prob_best
        self.step += 1
        return False


def generate_ids(model, inp, max_tokens, early_stop=False):
    """pv:385-398 with both stoppers: returns tokens [B, n <= max_tokens] (rows not truncated, H12). The loop ends when every
    row has emitted EOS at least once (TokenStopper pv:106-117: checked only on steps whose token batch contains an EOS) or the
    LogitStopper fires (pv:395)."""
    logits, cache = model(**inp, max_tokens=max_tokens)
    token = logits[:, -1, :].argmax(-1)[:, None]
    out = [token]
    eos_rows = torch.ones(token.shape[0])
    logit_stopper = LogitStopper(max_tokens, early_stop)
    for _ in range(max_tokens - 1):
        logits, cache = model(input_ids=token, cache=cache)
        token = logits[:, -1, :].argmax(-1)[:, None]
        out.append(token)
        if logit_stopper(logits):                                         # pv:395
            break
        if (token == ID_EOS).any():                                       # TokenStopper pv:112-117
            eos_rows = eos_rows * (token.squeeze(1) != ID_EOS)
            if eos_rows.sum() < 1:
                break
    return torch.cat(out, 1)


def truncate_rows(tokens):
    """Streamer.end pv:73: every row is cut after its first EOS."""
    return [(r[:r.index(ID_EOS) + 1] if ID_EOS in r else r) for r in tokens.tolist()]


def choose_ids(model, inp, option_ids):
    """pv:475-477: argmax over log_softmax(last logits)[:, options]."""
    logits, _ = model(**inp, max_tokens=0)
    lp = torch.log_softmax(logits[:, -1, :], -1)
    return lp[:, option_ids].argmax(-1)


def constrain_ids(model, inp, ids_c, max_new, use_beam=False, n_beam=3, log_norm=False):
    """One tuple constraint of pv:537-606 at token-id level. Returns (synth_sofar [B, *], score_sofar [B])."""
    import math

    def log_mean(x):
        return x.sum(-1) / (math.log(x.shape[-1]) if log_norm else x.shape[-1])

    idc = torch.tensor(ids_c)

    def get_beam(lp, cache, beam_idx):                                     # pv:505-517
        token = lp[:, beam_idx, :].argmax(-1)
        arg_beam = lp[:, beam_idx, :].topk(n_beam, dim=-1).indices          # unordered top-n in the reference
        beam = torch.cat([arg_beam.reshape(-1)[:, None], idc.repeat(arg_beam.numel(), 1)], -1)
        bl, _ = model(input_ids=beam, cache=cache, n_beam=n_beam, advance_offset=0)
        bl = torch.log_softmax(bl, -1)
        first = lp[torch.arange(arg_beam.shape[0])[:, None], beam_idx, arg_beam].reshape(-1)[:, None]
        rest = bl[torch.arange(bl.shape[0])[:, None], torch.arange(beam.shape[1] - 1)[None, :], beam[:, 1:]]
        bs = torch.cat([first, rest], 1)
        k = bs.mean(1).reshape(-1, n_beam).argmax(-1)
        ar = torch.arange(k.shape[0])
        return token, arg_beam[ar, k], bs.reshape(lp.shape[0], n_beam, -1)[ar, k]

    B = torch.as_tensor(inp['input_ids']).shape[0]
    C = len(ids_c)
    pad = torch.full((B, 1), ID_EOS)
    logits, cache = model(**inp, max_tokens=max_new + C + 10)
    lp = torch.log_softmax(logits, -1)
    s0 = lp[:, -1, ids_c[0]]
    tiled = idc.repeat(B, 1)
    lr, _ = model(input_ids=tiled, cache=cache, advance_offset=0)
    lr = torch.log_softmax(lr, -1)
    s1 = lr[torch.arange(B)[:, None], torch.arange(C - 1)[None, :], tiled[:, 1:]]
    running = lp[:, -1, :].max(-1).values[:, None]
    pre_score = log_mean(torch.cat([s0[:, None], s1], 1))
    pre_synth = torch.cat([tiled, pad], 1)
    if use_beam and max_new > 0:
        token, bt, bsc = get_beam(lp, cache, -1)
        post_score = log_mean(bsc)
        post_synth = torch.cat([bt[:, None], tiled], 1)
        win = pre_score > post_score
        score_sofar = torch.where(win, pre_score, post_score)
        synth_sofar = torch.where(win[:, None], pre_synth, post_synth)
    else:
        token = lp[:, -1, :].argmax(-1)
        score_sofar, synth_sofar = pre_score, pre_synth
    token = token[:, None]
    tokens = []
    alive = torch.ones(B)
    for _ in range(max_new):
        tokens.append(token)
        tp = torch.cat([token, tiled], 1)
        logits, cache = model(input_ids=tp, cache=cache, advance_offset=1)
        lp = torch.log_softmax(logits, -1)
        g = lp[torch.arange(B)[:, None], torch.arange(C)[None, :], tp[:, 1:]]
        pre_score = log_mean(torch.cat([running, g], 1))
        pre_synth = torch.cat(tokens + [tiled, pad], 1)
        if use_beam:
            token, bt, bsc = get_beam(lp, cache, 0)
            post_score = log_mean(torch.cat([running, bsc], 1))
            post_synth = torch.cat(tokens + [bt[:, None], tiled], 1)
            win = pre_score > post_score
            score = torch.where(win, pre_score, post_score)
            synth = torch.where(win[:, None], pre_synth, post_synth)
        else:
            token = lp[:, 0, :].argmax(-1)
            score, synth = pre_score, pre_synth
        synth_sofar = torch.cat([synth_sofar, pad], 1)
        toks = torch.cat(tokens, 1)
        if toks.shape[1] >= C:                                              # _already pv:495-498
            alive = alive * (~(toks[:, -C:] == idc).all(1)).float()
        upd = (score > score_sofar) & (alive > 0)
        synth_sofar = torch.where(upd[:, None], synth, synth_sofar)
        score_sofar = torch.where(upd, score, score_sofar)
        running = torch.cat([running, lp[torch.arange(B), 0, token][:, None]], 1)
        alive = alive * (token != ID_EOS).float()
        if alive.sum() < 1:
            break
        token = token[:, None]
    return synth_sofar, score_sofar
