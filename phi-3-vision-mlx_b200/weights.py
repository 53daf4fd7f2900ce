"""Seeded random-init checkpoints with the reference's parameter names.

Names follow the module tree of /root/reference/phi.py (Phi3ForCausalLM phi:599-617,
Phi3ImageEmbedding phi:374-391, ClipModel phi:208-221), i.e. the HF safetensors keys the
reference loads at /root/reference/phi_3_vision_mlx.py:265,371-374 (patch_embedding weight
already in the reference's post-transpose [O,kh,kw,I] layout).
All tensors are bf16 (HF checkpoints are bf16); the CPU oracle and the CUDA path are fed
the same dict. There is no network in this environment, so benchmarks and parity tests
use these instead of the published checkpoints (stated in bench.py's `data` field).
"""
import torch
from .configs import CLIP_VIT_L14_336


def _names(cfg, clip_cfg, vision):
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    hd = H // cfg.num_attention_heads
    qkv = cfg.num_attention_heads * hd + 2 * cfg.num_key_value_heads * hd
    out = [('model.embed_tokens.weight', (V, H), 'emb')]
    for i in range(cfg.num_hidden_layers):
        p = f'model.layers.{i}.'
        out += [(p + 'input_layernorm.weight', (H,), 'norm'), (p + 'self_attn.qkv_proj.weight', (qkv, H), 'lin'),
                (p + 'self_attn.o_proj.weight', (H, cfg.num_attention_heads * hd), 'res'),
                (p + 'post_attention_layernorm.weight', (H,), 'norm'),
                (p + 'mlp.gate_up_proj.weight', (2 * I, H), 'lin'), (p + 'mlp.down_proj.weight', (H, I), 'res')]
    out += [('model.norm.weight', (H,), 'norm'), ('lm_head.weight', (V, H), 'lin')]
    if vision:
        c = clip_cfg
        D, F = c.hidden_size, c.intermediate_size
        P = 'model.vision_embed_tokens.img_processor.vision_model.'
        npos = (c.image_size // c.patch_size) ** 2 + 1
        out += [(P + 'embeddings.class_embedding', (D,), 'b'),
                (P + 'embeddings.patch_embedding.weight', (D, c.patch_size, c.patch_size, c.num_channels), 'lin'),
                (P + 'embeddings.position_embedding.weight', (npos, D), 'b'),
                (P + 'pre_layrnorm.weight', (D,), 'norm'), (P + 'pre_layrnorm.bias', (D,), 'b')]
        for j in range(c.num_hidden_layers):
            L = P + f'encoder.layers.{j}.'
            for n in ('q_proj', 'k_proj', 'v_proj', 'out_proj'):
                out += [(L + f'self_attn.{n}.weight', (D, D), 'cres' if n == 'out_proj' else 'lin'),
                        (L + f'self_attn.{n}.bias', (D,), 'b')]
            out += [(L + 'layer_norm1.weight', (D,), 'norm'), (L + 'layer_norm1.bias', (D,), 'b'),
                    (L + 'layer_norm2.weight', (D,), 'norm'), (L + 'layer_norm2.bias', (D,), 'b'),
                    (L + 'mlp.fc1.weight', (F, D), 'lin'), (L + 'mlp.fc1.bias', (F,), 'b'),
                    (L + 'mlp.fc2.weight', (D, F), 'cres'), (L + 'mlp.fc2.bias', (D,), 'b')]
        out += [(P + 'post_layernorm.weight', (D,), 'norm'), (P + 'post_layernorm.bias', (D,), 'b')]
        E = cfg.img_processor['image_dim_out'] * 4
        Vp = 'model.vision_embed_tokens.'
        out += [(Vp + 'glb_GN', (1, 1, E), 'gn'), (Vp + 'sub_GN', (1, 1, 1, E), 'gn'),
                (Vp + 'img_projection.0.weight', (H, E), 'lin'), (Vp + 'img_projection.0.bias', (H,), 'b'),
                (Vp + 'img_projection.2.weight', (H, H), 'lin'), (Vp + 'img_projection.2.bias', (H,), 'b')]
    return out


PEAK_ALPHA = 0.01


def random_weights(cfg, seed=0, device='cpu', clip_cfg=None, vision=None, init='gpt2'):
    """N(0,sigma) bf16 tensors, GPT-2-style: linear 0.02, residual-writing projections (o_proj,
    down_proj, CLIP out_proj/fc2) 0.02/sqrt(2*n_layers), embedding 1.0, norm gains 1+0.1*N,
    biases/pos 0.1*N, GN separators N(0,1). With the residual scaling the residual stream stays
    O(1) over 32 layers and the random network is not chaotic (a plain 0.02 init amplifies a 1e-3
    perturbation ~25x by layer 32, which says nothing about kernel accuracy); logits have O(1)
    spread (SURVEY.md §7 hard part 1).

    init='peaked' (the parity checkpoint): identical tensors, except that lm_head is TIED to the embedding through a
    fixed seeded permutation, lm_head[t] = PEAK_ALPHA * embed[perm[t]]. The residual stream of this init keeps a cosine
    of ~0.68 with the input embedding through all 32 layers, so every position has ONE logit of ~20 over a N(0, 0.6)
    background: the top-1 / top-2 margin (~17) is far above the bf16 noise of any correct implementation (~2e-2 of the
    peak), which makes "greedy ids agree on >= 99 % of positions" a property of the kernels instead of a coin flip at
    the noise floor. Greedy generation walks the permutation (x -> perm^-1(x)), so the rollout is not a repeated
    token. The flat 'gpt2' init (median margin 0.18) stays available to measure the noise floor itself."""
    assert init in ('gpt2', 'peaked')
    if vision is None:
        vision = 'V' in cfg.architectures[0]
    clip_cfg = clip_cfg or CLIP_VIT_L14_336
    g = torch.Generator(device=device).manual_seed(seed)
    w = {}
    for name, shape, kind in _names(cfg, clip_cfg, vision):
        t = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if kind == 'lin':
            t.mul_(0.02)
        elif kind == 'res':
            t.mul_(0.02 / (2 * cfg.num_hidden_layers) ** 0.5)
        elif kind == 'cres':
            t.mul_(0.02 / (2 * clip_cfg.num_hidden_layers) ** 0.5)
        elif kind == 'norm':
            t.mul_(0.1).add_(1.0)
        elif kind == 'b':
            t.mul_(0.1)
        w[name] = t.to(torch.bfloat16)
    if init == 'peaked':
        perm = torch.randperm(cfg.vocab_size, generator=torch.Generator().manual_seed(seed + 12345))
        emb = w['model.embed_tokens.weight'].to(torch.float32)
        w['lm_head.weight'] = (PEAK_ALPHA * emb[perm.to(emb.device)]).to(torch.bfloat16)
    return w
