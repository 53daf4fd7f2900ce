"""Phi-3.5 (mini / vision) on B200 — host-side mirror of the reference model call protocol.

`Phi3B200.__call__` keeps the signature of Phi3ForCausalLM.__call__ (/root/reference/phi.py:606):
    model(input_ids, pixel_values=None, image_sizes=None, positions=None, cache=None, pids=None,
          mask=None, max_tokens=0, advance_offset=None, n_beam=1) -> (logits[B,L,V], cache)
All arithmetic runs in the sm_100a kernels of libphi3b200.so through the C ABI (`_lib.call`);
PyTorch only owns device memory and streams. There is no CPU path.

HBM layout
  weights        bf16, nn.Linear layout [N,K] (K-major for both TMA operands); gate_up_proj rows
                 interleaved per 256 rows as [128 gate | 128 up] so SwiGLU fuses into the epilogue
  hidden state   bf16 [B*L, H] (the reference keeps bf16 activations between modules)
  KV cache       paged: pool[layer][page][K|V][head][64 tokens][head_dim] bf16 (zero-initialised),
                 block_table int32 [B, pages]; left-padded rows keep the reference's absolute
                 indices and carry kv_start[b] = pad length (replaces Mask4D, phi:550-563)
  quantised KV   full prompt pages as 4-bit codes + bf16 (scale,bias) per 32 values (phi:528-540)
"""
import math
import torch
from . import _lib
from ._lib import call, ptr, PAGE
from .quant import W4, fake_quant
from .configs import CLIP_VIT_L14_336


def _stream():
    return torch.cuda.current_stream().cuda_stream


import contextlib


@contextlib.contextmanager
def capture_graph(graph):
    """torch.cuda.graph(graph) with the Python cycle collector out of the way: a dead reference cycle may own CUDA graphs or
    device memory of an earlier model / cache, and destroying one of those WHILE this stream is capturing invalidates the
    capture (cudaErrorStreamCaptureInvalidated). Collect before, keep the collector off during the capture."""
    import gc
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was:
            gc.enable()


def pick_splits(n_bh, tiles, slots=2 * 148, max_splits=32):
    """Split-KV factor for decode attention from a cost model fitted to tools/microbench.py on B200:
        T(s) = 2 us + 2.5 us * waves + bytes / min(5.95 TB/s, concurrent_CTAs * 34 GB/s)
    (waves = ceil(n_bh*s / 296 resident CTAs); one CTA alone sustains ~34 GB/s from its 72 KB in
    flight). Long streams per CTA beat full waves: 256 CTAs x 34 pages reach 91 % of the measured
    HBM peak, 2048 CTAs x 4 pages only 62 %."""
    best, best_t = 1, float('inf')
    total_mb = n_bh * tiles * 24.576e-3
    for s in range(1, max_splits + 1):
        if s > 1 and tiles / s < 2:
            break
        ctas = n_bh * s
        waves = -(-ctas // slots)
        bw = min(5950.0, min(ctas, slots) * 34.0)                 # GB/s
        t = 2.0 + 2.5 * waves + total_mb / bw * 1e3               # us
        if t < best_t * 0.98:
            best, best_t = s, t
    return best


def pack_rows16(w):
    """[N, K] -> the same elements in the stream order of the 16-row skinny tiles (csrc/gemm_skinny.cu, SkParams.packed):
    [N/16 tiles][K/64 chunks][row half][k half][lane = 4 * (row % 8) + 16-byte quad][8 elements]; a tile is one contiguous block."""
    N, K = w.shape
    assert N % 16 == 0 and K % 64 == 0
    return w.view(N // 16, 2, 8, K // 64, 2, 4, 8).permute(0, 3, 1, 4, 2, 5, 6).contiguous().view(N, K)


def interleave_gate_up(w):
    """[2I, K] (gate rows then up rows, phi:470) -> per 256-row block [128 gate | 128 up]."""
    I = w.shape[0] // 2
    assert I % 128 == 0, 'intermediate_size must be a multiple of 128'
    g, u = w[:I].reshape(I // 128, 128, -1), w[I:].reshape(I // 128, 128, -1)
    return torch.stack([g, u], dim=1).reshape(2 * I, -1).contiguous()


class _KVSlab:
    """Device buffers behind one KV cache shape (B, L, max_tokens): page pool, block table, rope
    tables, decode-loop state and the captured decode-step CUDA graph. Slabs are recycled through
    `Phi3B200._slabs` when the cache object that borrowed them is dropped, so repeated generate()
    calls of the same shape neither re-allocate (7 GB at 8 x 2304 tokens) nor re-capture."""

    def __init__(self, cfg, B, L, max_tokens, dev):
        nl, nkv = cfg.num_hidden_layers, cfg.num_key_value_heads
        hd = cfg.hidden_size // cfg.num_attention_heads
        self.pages_per_seq = (L + max_tokens + PAGE - 1) // PAGE
        n_pages = B * self.pages_per_seq
        # zero-initialised: masked slots must hold finite values (0 * NaN would poison P.V); recycled
        # slabs only ever contain finite bf16 values written by the kernels
        self.pool = torch.zeros((nl, n_pages, 2, nkv, PAGE, hd), dtype=torch.bfloat16, device=dev)
        self.block_table = torch.arange(n_pages, dtype=torch.int32, device=dev).reshape(B, self.pages_per_seq)
        self.kv_start = torch.zeros(B, dtype=torch.int32, device=dev)
        self.cos = self.sin = None
        self.qcodes = self.qmeta = None
        self.session = None                  # DecodeSession state + graph, keyed by max_steps


class KVCacheB200:
    """Opaque cache handle returned to the decode drivers (the reference returns a list of
    per-layer KVCache objects, phi:581; drivers only pass it back). Indexing returns the handle
    itself so `cache[0].offset` works (no self-referencing list: the slab must be released by
    reference counting the moment the caller drops the cache)."""

    def __getitem__(self, i):
        if not -self.n_layers <= i < self.n_layers:          # a finite sequence, like the reference's list of per-layer caches
            raise IndexError(i)
        return self

    def __iter__(self):
        return iter([self] * self.n_layers)

    def __len__(self):
        return self.n_layers

    def __init__(self, model, B, L, max_tokens, quantized):
        self.n_layers = model.cfg.num_hidden_layers
        self.B, self.S_max, self.max_tokens = B, L + max_tokens, max_tokens
        self.offset = 0
        self.quantized, self.n_quant = quantized, 0
        self._key = (B, L, max_tokens, bool(quantized))
        self._slabs = model._slabs
        free = self._slabs.get(self._key)
        self.slab = free.pop() if free else _KVSlab(model.cfg, B, L, max_tokens, model.dev)
        sl = self.slab
        self.pages_per_seq, self.pool, self.block_table, self.kv_start = sl.pages_per_seq, sl.pool, sl.block_table, sl.kv_start
        self.cos = self.sin = None
        self.tab_bstride = 0
        self.qcodes, self.qmeta = sl.qcodes, sl.qmeta

    def set_tables(self, cos, sin, tbs, kvs):
        """Copy per-call tables into the slab's static buffers (addresses stay fixed for the graph)."""
        sl = self.slab
        if sl.cos is None or sl.cos.shape != cos.shape:
            sl.cos, sl.sin, sl.session = cos.clone(), sin.clone(), None
        else:
            sl.cos.copy_(cos); sl.sin.copy_(sin)
        sl.kv_start.copy_(kvs)
        self.cos, self.sin, self.tab_bstride = sl.cos, sl.sin, tbs

    def release(self):
        """Hand the slab (page pool, tables, captured decode graph) back to the model for the next cache of this shape.
        The drivers call this when a generation is finished; `__del__` is only the safety net for callers that drop the
        handle without releasing it. The handle must not be used afterwards."""
        slab, self.slab = self.slab, None
        if slab is None:
            return
        if self._key not in self._slabs and len(self._slabs) >= 6:          # bound the recycled memory
            self._slabs.pop(next(iter(self._slabs)))
        lst = self._slabs.setdefault(self._key, [])
        if len(lst) < 2:
            slab.qcodes, slab.qmeta = self.qcodes, self.qmeta
            lst.append(slab)
        self.pool = self.block_table = self.kv_start = self.cos = self.sin = self.qcodes = self.qmeta = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Phi3B200:
    _LAYER_KEYS = {'self_attn.qkv_proj': 'qkv', 'self_attn.o_proj': 'o', 'mlp.gate_up_proj': 'gu', 'mlp.down_proj': 'down'}

    def __init__(self, cfg, weights, device='cuda', clip_cfg=None, gemm_impl=0, quantize_model=False, lora=None):
        """`lora`: None or {module name ('model.layers.3.self_attn.qkv_proj'): (lora_a [in, r], lora_b [r, out], scale*alpha/r)}
        — LoRALinear (phi:84-133) folded into the decode / prefill weight copies on the device; swap with set_adapter()."""
        if not torch.cuda.is_available():
            raise RuntimeError('Phi3B200 needs a CUDA device (sm_100a); there is no CPU fallback')
        _lib.lib()
        self.cfg, self.dev = cfg, torch.device(device)
        self.clip_cfg = clip_cfg or CLIP_VIT_L14_336
        self.gemm_impl = gemm_impl
        self.H = cfg.hidden_size
        self.n_heads, self.n_kv = cfg.num_attention_heads, cfg.num_key_value_heads
        if self.n_kv != self.n_heads:
            # the reference graph has no repeat_kv (phi.py:454 multiplies q[B,H,..] with k[B,Hkv,..] directly), so
            # grouped-query checkpoints are outside its behaviour; the kernels take n_kv but the host graph does not
            raise NotImplementedError('num_key_value_heads != num_attention_heads is not supported by the reference graph')
        self.hd = self.H // self.n_heads
        self.qkv_dim = (self.n_heads + 2 * self.n_kv) * self.hd
        self.I = cfg.intermediate_size
        self.V = cfg.vocab_size
        self.eps = float(cfg.rms_norm_eps)
        self.scale = self.hd ** -0.5
        self.use_quantized_cache = bool(getattr(cfg, 'use_quantized_cache', False))
        d = lambda t: t.to(self.dev, torch.bfloat16).contiguous()
        w = weights
        # quantize_model (pv:264,291-305: nn.quantize(model, 64, 4) over every Linear / Embedding): the matrices are
        # replaced by their 4-bit images. `lin` keeps bf16(dequantised) for the tensor-core GEMMs and registers the
        # packed 4-bit stream of the decode-time skinny GEMM under the tensor's address; `emb` is dequantise-only.
        self.quantize_model = bool(quantize_model)
        self._w4 = {}

        self._lora = dict(lora or {})
        self._raw = {}                # pristine checkpoint tensors of every matrix that ever carried an adapter (for hot swaps)

        def lin(t, row_perm=None, name=None):
            t = d(t)
            ad = self._lora.get(name)
            if ad is not None:
                # LoRALinear over a Linear / QuantizedLinear (phi:94-95,129-133): base output + scale * (x A) B. Folded:
                # W' = bf16(base + scale (A B)^T) with base = the 4-bit image when the model is quantised. Such a matrix streams
                # as bf16 at decode (it is no longer a 4-bit code).
                self._raw[name] = t
                base = fake_quant(t) if self.quantize_model else t
                t = self._fold(base, ad)
                return t if row_perm is None else row_perm(t)
            if not self.quantize_model:
                return t if row_perm is None else row_perm(t)
            q = W4(t, pack=True, row_perm=row_perm)
            self._w4[q.deq.data_ptr()] = (q.codes, q.meta)
            return q.deq
        emb = (lambda t: fake_quant(d(t))) if self.quantize_model else d
        self._lin_prefill, self._emb = ((lambda t: fake_quant(d(t))) if self.quantize_model else d), emb
        self.embed = emb(w['model.embed_tokens.weight'])
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f'model.layers.{i}.'
            self.layers.append(dict(
                ln1=d(w[p + 'input_layernorm.weight']), qkv=lin(w[p + 'self_attn.qkv_proj.weight'], None, p + 'self_attn.qkv_proj'),
                o=lin(w[p + 'self_attn.o_proj.weight'], None, p + 'self_attn.o_proj'),
                ln2=d(w[p + 'post_attention_layernorm.weight']),
                gu=lin(w[p + 'mlp.gate_up_proj.weight'], interleave_gate_up, p + 'mlp.gate_up_proj'),
                down=lin(w[p + 'mlp.down_proj.weight'], None, p + 'mlp.down_proj')))
        self.norm = d(w['model.norm.weight'])
        self.lm_head = lin(w['lm_head.weight'])
        # fused prefill (north_star: SuRoPE + KV write in the QKV epilogue, RMSNorm fused with the adjacent GEMM): second copies of
        # qkv_proj / gate_up_proj with the RMSNorm gain folded in (W' = bf16(W * g), the per-row rsqrt is applied to the
        # accumulators) and, for qkv, q/k rows permuted per head so a rotary pair meets in one 32-column accumulator chunk
        import os as _os1
        self.pf_fused = _os1.environ.get('P3_PF_FUSED', '1') != '0' and self.hd % 32 == 0
        if self.pf_fused:
            self._rope_perm = self._rope_row_perm()
            for lw in self.layers:
                lw['qkv_pf'], lw['gu_pf'] = self._prefill_copy(lw, 'qkv'), self._prefill_copy(lw, 'gu')
                lw['plans'] = {k: _lib.WeightPlan(lw[k]) for k in ('qkv_pf', 'o', 'gu_pf', 'down')}
        # decode stream of o_proj / down_proj in tile order (pack_rows16): second copies of the bf16 matrices that stream as 16-row
        # tiles (2.2 GB at Phi-3.5 sizes). P3_SK_PACK=0: read the row-major matrices.
        self._pk = {}
        if _os1.environ.get('P3_SK_PACK', '1') != '0':
            for lw in self.layers:
                for key in ('o', 'down'):
                    t = lw[key]
                    if t.data_ptr() not in self._w4 and t.shape[0] % 16 == 0 and t.shape[0] < 148 * 64 and t.shape[1] % 64 == 0:
                        lw[key + '_pk'] = self._pk[t.data_ptr()] = pack_rows16(t)
        self._raw_src = w             # names only looked up for adapter targets (set_adapter); the dict stays the caller's
        # persistent decode-layer kernel (mega.py / decode_mega.cu): a second, stream-order copy of the decoder weights
        # (7.4 GB at Phi-3.5 sizes; HBM has 180 GB). bf16 weights only. Opt-in (P3_MEGA=1) while it is slower than the chain
        # of per-matrix skinny kernels it replaces (profiles/r02_decode_mega.md).
        self.mega = None
        import os as _os0
        if not self.quantize_model and _os0.environ.get('P3_MEGA', '0') == '1':
            from .mega import MegaDecoder
            try:
                self.mega = MegaDecoder(self, w)
            except ValueError:
                self.mega = None                      # shapes the kernel does not tile (K not a multiple of 128, ...)
        self.vision = None
        if any(k.startswith('model.vision_embed_tokens') for k in w):
            self._load_vision(w, d)
        self._slabs = {}              # recycled KV slabs, keyed by (B, L, max_tokens, quantized)
        self._row_maps = {}
        # split-K workspace of p3_gemm_fused (arrival counters + fp32 partial tiles); P3_SPLITK=0 turns split-K off
        self._splitk_ws = (torch.zeros(32 << 20, dtype=torch.uint8, device=self.dev)
                           if _os1.environ.get('P3_SPLITK', '1') != '0' else None)
        self.force_long_rope = None   # parallel.py: LongRoPE switch decided from the global batch (H7)
        self.prefill_chunk = 8192     # long prompts are prefilled in chunks against the paged cache (config 4: 128K)
        import os as _os
        self.pf_chain = _os.environ.get('P3_PF_CHAIN', '1') != '0'
        self._skip = set(filter(None, _os.environ.get('P3_SKIP', '').split(',')))   # timing ablation only (tools/ablate.py)
        # decode: RMSNorm as its own tiny PDL kernel instead of fused into every 32-row CTA of the following skinny GEMM.
        # Neutral for the bf16 weight stream (2.797 vs 2.788 ms/token on the bench), -8 % for the instruction-bound 4-bit one.
        self._opf_in_qkv = _os.environ.get('P3_OPF', 'qkv') != 'attn'
        self.prenorm = _os.environ.get('P3_PRENORM', '1' if self.quantize_model else '0') != '0'
        # decode: RMSNorm split between producer and consumer (p3_gemm_skinny_x): the residual epilogue also writes
        # bf16(h * gain_of_next_norm), the consumer applies rstd to its reduced accumulators -> no norm kernel, no per-warp
        # re-normalisation inside the K loop, no statistics reduction in front of it. Built for the 4-bit stream (7 -> 5 launches
        # per layer); on bf16 weights it replaces the fused-norm prologue and measured -3.8 % (8 x 2048), -5.3 % (16 x 448), -5.9 %
        # (4 x 128) per decode step (profiles/r02_split_norm_bf16_ab.log). P3_XG=0 restores the fused / separate norm.
        self.xg = _os.environ.get('P3_XG', '1') != '0'
        self.profile = None      # bench.py: list of (kind, ev0, ev1, algorithmic_bytes) when instrumenting

    # ------------------------------------------------------------------ vision weights
    def _load_vision(self, w, d):
        cc = self.clip_cfg
        P = 'model.vision_embed_tokens.img_processor.vision_model.'
        D = cc.hidden_size
        kp = 640                                         # 588 padded to a multiple of the 64-wide K block
        pw = torch.zeros((D, kp), dtype=torch.bfloat16, device=self.dev)
        pw[:, :588] = d(w[P + 'embeddings.patch_embedding.weight']).reshape(D, -1)     # [O,kh,kw,I] (pv:374)
        ql, qe = self._lin_prefill, self._emb           # quantize_model: Linear / Embedding weights -> 4-bit images (conv stays)
        v = dict(kpad=kp, patch_w=pw, cls=d(w[P + 'embeddings.class_embedding']),
                 pos=qe(w[P + 'embeddings.position_embedding.weight']),
                 pre_w=d(w[P + 'pre_layrnorm.weight']), pre_b=d(w[P + 'pre_layrnorm.bias']), layers=[])
        for j in range(cc.num_hidden_layers - 1):        # the reference runs layers[:-1] (phi:219)
            L = P + f'encoder.layers.{j}.'
            v['layers'].append(dict(
                ln1w=d(w[L + 'layer_norm1.weight']), ln1b=d(w[L + 'layer_norm1.bias']),
                qkv_w=torch.cat([ql(w[L + f'self_attn.{n}_proj.weight']) for n in 'qkv'], 0).contiguous(),
                qkv_b=torch.cat([d(w[L + f'self_attn.{n}_proj.bias']) for n in 'qkv'], 0).contiguous(),
                out_w=ql(w[L + 'self_attn.out_proj.weight']), out_b=d(w[L + 'self_attn.out_proj.bias']),
                ln2w=d(w[L + 'layer_norm2.weight']), ln2b=d(w[L + 'layer_norm2.bias']),
                fc1_w=ql(w[L + 'mlp.fc1.weight']), fc1_b=d(w[L + 'mlp.fc1.bias']),
                fc2_w=ql(w[L + 'mlp.fc2.weight']), fc2_b=d(w[L + 'mlp.fc2.bias'])))
        Vp = 'model.vision_embed_tokens.'
        v.update(glb_GN=d(w[Vp + 'glb_GN']).reshape(-1), sub_GN=d(w[Vp + 'sub_GN']).reshape(-1),
                 p0_w=ql(w[Vp + 'img_projection.0.weight']), p0_b=d(w[Vp + 'img_projection.0.bias']),
                 p2_w=ql(w[Vp + 'img_projection.2.weight']), p2_b=d(w[Vp + 'img_projection.2.bias']))
        self.vision = v

    # ------------------------------------------------------------------ LoRA adapters (phi:84-133, pv:234-245, 266-271)
    def _fold(self, base, ad):
        a, b, sc = ad
        delta = (a.to(self.dev, torch.float32) @ b.to(self.dev, torch.float32)).T                  # [out, in]
        return (base.to(torch.float32) + float(sc) * delta).to(torch.bfloat16).contiguous()

    def _prefill_copy(self, lw, key):
        """prefill copy of qkv / gate_up: RMSNorm gain folded in (+ rope row permutation for qkv), see __init__"""
        if key == 'qkv':
            return (lw['qkv'].to(torch.float32)[self._rope_perm] * lw['ln1'].to(torch.float32)[None, :]).to(torch.bfloat16).contiguous()
        return (lw['gu'].to(torch.float32) * lw['ln2'].to(torch.float32)[None, :]).to(torch.bfloat16).contiguous()

    def set_adapter(self, lora=None):
        """Swap the LoRA adapter WITHOUT reloading the model (`lora` as in __init__, None = base weights). Every matrix that
        carries the old or the new adapter is rebuilt from its pristine checkpoint tensor and written IN PLACE into the device
        copies the kernels read (decode stream, prefill copy, stream-order copy of the persistent kernel), so weight TMA plans,
        KV slabs and captured decode graphs stay valid. Result is identical to constructing the model with `lora`."""
        lora = dict(lora or {})
        names = set(self._lora) | set(lora)
        for name in sorted(names):
            li = int(name.split('.')[2])
            key = self._LAYER_KEYS.get(name.split('.', 3)[3])
            if key is None:
                raise KeyError(f'adapter target {name} is not a decoder-layer projection')
            lw = self.layers[li]
            if name not in self._raw:
                if self.quantize_model and lw[key].data_ptr() in self._w4:
                    raise NotImplementedError('adding an adapter to a matrix that was loaded as a 4-bit stream needs a reload '
                                              '(load with use_adapter=True so that the targeted matrices stream as bf16)')
                self._raw[name] = self._raw_src[name + '.weight'].to(self.dev, torch.bfloat16).contiguous()
            base = fake_quant(self._raw[name]) if self.quantize_model else self._raw[name]
            t = self._fold(base, lora[name]) if name in lora else base
            if key == 'gu':
                t = interleave_gate_up(t)
            lw[key].copy_(t)
            if key + '_pk' in lw:
                lw[key + '_pk'].copy_(pack_rows16(lw[key]))
            if self.pf_fused and key in ('qkv', 'gu'):
                lw[key + '_pf'].copy_(self._prefill_copy(lw, key))
            if self.mega is not None:
                src = t if key != 'gu' else (self._fold(base, lora[name]) if name in lora else base)   # checkpoint row order
                self.mega.repack(li, key, src)
        self._lora = lora

    def _rope_row_perm(self):
        """row order of the prefill copy of qkv_proj: per q/k head [16j..16j+15 | half+16j..half+16j+15] for j = 0..hd/32-1
        (P3_EPI_ROPE_KV, include/phi3_b200.h); v rows unchanged"""
        hd, half = self.hd, self.hd // 2
        idx = []
        for h in range(self.n_heads + self.n_kv):
            for j in range(hd // 32):
                idx += [h * hd + 16 * j + i for i in range(16)] + [h * hd + half + 16 * j + i for i in range(16)]
        idx += list(range((self.n_heads + self.n_kv) * hd, self.qkv_dim))
        return torch.tensor(idx, dtype=torch.int64, device=self.dev)

    def gemm_fused(self, x, w, out, epi, plan=None, resid=None, ss_in=None, ss_out=None, rope=None):
        """p3_gemm_fused: tensor-core GEMM with the RMSNorm row scale / sum-of-squares / rope + KV-write epilogues"""
        a = _lib.GemmArgs()
        a.X, a.ldx, a.W, a.ldw = ptr(x), x.stride(0), ptr(w), w.stride(0)
        a.out, a.ldo, a.resid = ptr(out), out.stride(0), ptr(resid)
        a.M, a.N, a.K, a.epi, a.impl = x.shape[0], w.shape[0], x.shape[1], epi, self.gemm_impl
        if ss_in is not None:
            a.ss_in, a.n_ss_in, a.eps = ptr(ss_in), ss_in.shape[1], self.eps
        a.ss_out = ptr(ss_out)
        a.w_plan = None if plan is None else plan.addr
        if self._splitk_ws is not None:
            a.splitk_ws, a.splitk_ws_bytes = ptr(self._splitk_ws), self._splitk_ws.numel()
        if rope is not None:
            for k, v in rope.items():
                setattr(a, k, v)
        _lib.call_struct('p3_gemm_fused', a, _stream())
        return out

    # ------------------------------------------------------------------ thin kernel wrappers
    def gemm(self, x, w, out, epi=_lib.EPI_NONE, bias=None, resid=None, row_map=None, N=None):
        M, K = x.shape
        N = w.shape[0] if N is None else N
        call('p3_gemm', ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(out), out.stride(0), ptr(resid),
             ptr(row_map), M, N, K, epi, self.gemm_impl, _stream())
        return out

    L2_PF_CAP = int(__import__('os').environ.get('P3_PF_CAP_MB', '16')) << 20   # bytes of the next kernel's weights parked in L2

    def skinny(self, x, w, out, epi=_lib.EPI_NONE, norm_w=None, resid=None, ss_in=None, ss_out=None, nxt=None,
               xg_gain=None, xg_out=None, rs_epi=False):
        M, K = x.shape
        if xg_out is not None or rs_epi:                    # producer / consumer halves of the split RMSNorm
            ev = self._ev()
            nxt, pf = self._prefetch_target(nxt, M)
            q = self._codes(w, M)
            a = _lib.SkinnyArgs()
            a.op, a.X, a.ldx, a.eps = 0, ptr(x), x.stride(0), self.eps
            pk = self._pk.get(w.data_ptr()) if q is None else None
            if q is not None:
                a.Wq, a.Wmeta = ptr(q[0]), ptr(q[1])
            elif pk is not None:
                a.W, a.packed = ptr(pk), 1
            else:
                a.W = ptr(w)
            a.out, a.ldo, a.resid = ptr(out), out.stride(0), ptr(resid)
            a.M, a.N, a.K, a.epi = M, w.shape[0], K, epi
            a.ss_in, a.n_ss_in, a.ss_out = ptr(ss_in), (0 if ss_in is None else ss_in.shape[0]), ptr(ss_out)
            a.l2_prefetch, a.l2_prefetch_bytes = ptr(nxt), pf
            a.xg_gain, a.xg_out, a.ldxg, a.rs_epi = ptr(xg_gain), ptr(xg_out), (0 if xg_out is None else xg_out.stride(0)), int(rs_epi)
            _lib.call_struct('p3_gemm_skinny_x', a, _stream())
            self._ev(ev, 'skinny', (q[0].numel() + q[1].numel() * 2) if q is not None else w.shape[0] * K * 2)
            return out
        if norm_w is not None and self.prenorm:
            x, norm_w, ss_in = self._prenorm(x, norm_w), None, None
        ev = self._ev()
        n_ss = 0 if ss_in is None else ss_in.shape[0]
        nxt, pf = self._prefetch_target(nxt, M)
        q = self._codes(w, M)
        if q is not None:                                   # quantize_model: stream the 4-bit codes instead of bf16
            call('p3_gemm_skinny_w4', ptr(x), x.stride(0), ptr(norm_w), self.eps, ptr(q[0]), ptr(q[1]), ptr(out), out.stride(0),
                 ptr(resid), M, w.shape[0], K, epi, ptr(ss_in), n_ss, ptr(ss_out), ptr(nxt), pf, _stream())
            self._ev(ev, 'skinny', q[0].numel() + q[1].numel() * 2)
            return out
        call('p3_gemm_skinny', ptr(x), x.stride(0), ptr(norm_w), self.eps, ptr(w), ptr(out), out.stride(0),
             ptr(resid), M, w.shape[0], K, epi, ptr(ss_in), n_ss, ptr(ss_out), ptr(nxt), pf, _stream())
        self._ev(ev, 'skinny', w.shape[0] * K * 2)
        return out

    def _prenorm(self, x, norm_w):
        """decode (<= 16 rows): x_hat = bf16(x * rsqrt(mean(x^2) + eps) * w) by a one-CTA-per-row PDL kernel"""
        xn = torch.empty_like(x)
        call('p3_rmsnorm', ptr(x), ptr(norm_w), ptr(xn), x.shape[0], x.shape[1], self.eps, _stream())
        return xn

    # Rows up to which a quantised matrix is streamed as 4-bit codes. The 4-bit kernel is instruction-bound (integer unpack + an extra
    # MMA per group for the affine term); with two 8-token groups per warp (9..16 rows) it is slower than the bf16 stream over the
    # dequantised image that the prefill GEMMs use anyway (B = 16 x 448: 3.23 vs 2.73 ms per step), so those steps take the bf16 path.
    W4_MAX_ROWS = int(__import__('os').environ.get('P3_W4_MAX_ROWS', '8'))

    def _codes(self, w, M):
        """(codes, meta) of a quantised matrix when an M-row stream should read the 4-bit image, else None"""
        return self._w4.get(w.data_ptr()) if M <= self.W4_MAX_ROWS else None

    def _prefetch_target(self, nxt, M=1):
        """(tensor, bytes) of the next kernel's weight stream to park in L2: the 4-bit codes when the matrix is quantised"""
        if nxt is None:
            return None, 0
        q = self._codes(nxt, M)
        if q is not None:
            return q[0], min(q[0].numel(), self.L2_PF_CAP)
        if self.xg and nxt.data_ptr() in self._pk:          # the decode stream reads the tile-order copy
            nxt = self._pk[nxt.data_ptr()]
        return nxt, min(nxt.numel() * 2, self.L2_PF_CAP)

    def _ev(self, start=None, kind=None, nbytes=0):
        if self.profile is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if start is not None:
            self.profile.append((kind, start, e, nbytes))
        return e

    def linear(self, x, w, out, epi=_lib.EPI_NONE, norm_w=None, resid=None, ss_in=None, ss_out=None, nxt=None, **xg):
        """Route by token count: <=16 rows is a weight stream (skinny), else tensor-core GEMM.
        `nxt`: weights of the kernel that follows (decode only): pulled into L2 across the kernel boundary.
        `xg` (xg_gain / xg_out / rs_epi): the split RMSNorm of the 4-bit decode path, see skinny()."""
        if x.shape[0] <= 16:
            return self.skinny(x, w, out, epi, norm_w, resid, ss_in, ss_out, nxt, **xg)
        if norm_w is not None:
            xn = torch.empty_like(x)
            call('p3_rmsnorm', ptr(x), ptr(norm_w), ptr(xn), x.shape[0], x.shape[1], self.eps, _stream())
            x = xn
        return self.gemm(x, w, out, epi, resid=resid)

    def _splits(self, cache, B, tiles, L=1):
        n_bh = B * self.n_heads
        if cache is not None and cache.quantized:
            # the 4-bit kernel is bound by per-warp instruction latency, not by bytes: fill the resident slots
            # (4 CTAs/SM for L <= 8 query rows, 3 for the 16-row variant) in one wave
            slots = (4 if L <= 8 else 3) * 148
            return max(1, min((slots + n_bh // 2) // n_bh, max(1, tiles // 4)))
        return pick_splits(n_bh, tiles)

    # ------------------------------------------------------------------ rope table (phi:487-507)
    def _rope_table(self, L_all, pids):
        cfg = self.cfg
        sf = math.sqrt(1 + math.log(cfg.max_position_embeddings / cfg.original_max_position_embeddings)
                       / math.log(cfg.original_max_position_embeddings))
        use_long = L_all > cfg.original_max_position_embeddings if self.force_long_rope is None else self.force_long_rope
        fac = cfg.rope_scaling['long_factor'] if use_long else cfg.rope_scaling['short_factor']   # static switch (H7)
        # inv_freq: 48 values, fp32 on the host exactly as phi:498; the [Bt, L_all, 48] table itself is built by p3_rope_table on
        # the device (a 128K-token table is 2 x 25 MB — round 1 built it with CPU torch and copied it over PCIe per call)
        inv_freq = (1.0 / (torch.tensor(fac, dtype=torch.float32)
                           * cfg.rope_theta ** (torch.arange(0, self.hd, 2, dtype=torch.float32) / self.hd))).to(self.dev)
        half = self.hd // 2
        if pids is None:
            Bt, pid_dev, Lp, stride = 1, None, 0, 0
        else:
            pid_dev = torch.as_tensor(pids).to(self.dev, torch.int32).contiguous()
            Bt, Lp = pid_dev.shape
            stride = pid_dev.stride(0)
        cos = torch.empty((Bt, L_all, half), dtype=torch.float32, device=self.dev)
        sin = torch.empty((Bt, L_all, half), dtype=torch.float32, device=self.dev)
        call('p3_rope_table', ptr(pid_dev), stride, Lp, ptr(inv_freq), ptr(cos), ptr(sin), Bt, L_all, half, float(sf), _stream())
        return cos, sin, (0 if cos.shape[0] == 1 else cos.shape[1] * cos.shape[2])

    # ------------------------------------------------------------------ vision tower (phi:135-226, 393-416)
    def _vision_embed(self, h, L, pixel_values, image_sizes, positions):
        v, cc = self.vision, self.clip_cfg
        D, nh = cc.hidden_size, cc.num_attention_heads
        st = _stream()
        sizes = (torch.as_tensor(image_sizes).cpu() // 336).tolist()
        positions = torch.as_tensor(positions).cpu().tolist()
        pv = torch.as_tensor(pixel_values).to(self.dev, torch.float32)
        # only crops 0..hc*wc of each image reach the output (phi:405-406); zero-pad crops are skipped
        if all(hw[0] * hw[1] + 1 == pv.shape[1] for hw in sizes) and pv.is_contiguous():
            px = pv.reshape(-1, *pv.shape[2:])                  # every crop slot is used: a view, no copy kernel
        else:
            px = torch.cat([pv[b, :hw[0] * hw[1] + 1] for b, hw in enumerate(sizes)], 0).contiguous()
        x = self.clip_features(px)
        # token assembly + projector + splice (phi:400-415)
        crop0, idx = 0, 0
        for b, (hc, wc) in enumerate(sizes):
            cnt = (hc * wc + 1) * 144 + 1 + (hc + 1) * 12
            feats = x[crop0 * 577:(crop0 + hc * wc + 1) * 577]
            tok = torch.empty((cnt, 4 * D), dtype=torch.bfloat16, device=self.dev)
            call('p3_gn_assemble', ptr(feats), ptr(v['sub_GN']), ptr(v['glb_GN']), ptr(tok), hc, wc, D, st)
            p1 = torch.empty((cnt, self.H), dtype=torch.bfloat16, device=self.dev)
            self.gemm(tok, v['p0_w'], p1, _lib.EPI_GELU, bias=v['p0_b'])
            r, c = positions[idx]
            row_map = self._row_maps.get((cnt, r * L + c))     # destination rows of the splice: built once per (length, offset)
            if row_map is None:
                if len(self._row_maps) > 256:
                    self._row_maps.clear()
                row_map = self._row_maps[(cnt, r * L + c)] = (torch.arange(cnt, dtype=torch.int32, device=self.dev) + (r * L + c)).contiguous()
            self.gemm(p1, v['p2_w'], h, _lib.EPI_NONE, bias=v['p2_b'], row_map=row_map)
            crop0 += hc * wc + 1
            idx += cnt
        return h

    def clip_features(self, px):
        """ClipVModel (phi:208-226): px fp32 [N,3,336,336] on the device -> fp32 residual stream [N*577, 1024] after the 23
        encoder layers the reference runs (row 577*n is the CLS token, which phi:221 drops)."""
        v, cc = self.vision, self.clip_cfg
        D, nh = cc.hidden_size, cc.num_attention_heads
        st = _stream()
        N = px.shape[0]
        T = N * 577
        A = torch.empty((N * 576, v['kpad']), dtype=torch.bfloat16, device=self.dev)
        call('p3_patch_im2col', ptr(px), ptr(A), N, v['kpad'], st)
        patches = torch.empty((N * 576, D), dtype=torch.float32, device=self.dev)
        self.gemm(A, v['patch_w'], patches, _lib.EPI_F32)
        x = torch.empty((T, D), dtype=torch.float32, device=self.dev)
        call('p3_clip_embed', ptr(patches), ptr(v['cls']), ptr(v['pos']), ptr(x), N, D, st)
        call('p3_layernorm', ptr(x), ptr(v['pre_w']), ptr(v['pre_b']), ptr(x), T, D, 1e-5, 1, st)
        xn = torch.empty((T, D), dtype=torch.bfloat16, device=self.dev)
        qkv = torch.empty((T, 3 * D), dtype=torch.bfloat16, device=self.dev)
        att = torch.empty((T, D), dtype=torch.bfloat16, device=self.dev)
        mid = torch.empty((T, cc.intermediate_size), dtype=torch.bfloat16, device=self.dev)
        hd = D // nh
        for lw in v['layers']:
            call('p3_layernorm', ptr(x), ptr(lw['ln1w']), ptr(lw['ln1b']), ptr(xn), T, D, cc.layer_norm_eps, 0, st)
            self.gemm(xn, lw['qkv_w'], qkv, _lib.EPI_NONE, bias=lw['qkv_b'])
            call('p3_attention_prefill', ptr(qkv), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, 3 * D, 3 * D, 3 * D,
                 ptr(att), D, N, 577, nh, nh, hd, hd ** -0.5, 0, 0, None, None, None, 0, 1, st)
            self.gemm(att, lw['out_w'], x, _lib.EPI_RESIDUAL_F32, bias=lw['out_b'], resid=x)
            call('p3_layernorm', ptr(x), ptr(lw['ln2w']), ptr(lw['ln2b']), ptr(xn), T, D, cc.layer_norm_eps, 0, st)
            self.gemm(xn, lw['fc1_w'], mid, _lib.EPI_QGELU, bias=lw['fc1_b'])
            self.gemm(mid, lw['fc2_w'], x, _lib.EPI_RESIDUAL_F32, bias=lw['fc2_b'], resid=x)
        return x

    # ------------------------------------------------------------------ one decoder pass
    def decode_scratch(self, B):
        """activation buffers of one decode step for B rows, allocated once per decode session (outside the captured graph).
        The sum-of-squares partials need no initialisation: every slot a consumer reads is written by its producer."""
        dev, n_part = self.dev, (self.H + 15) // 16
        e = lambda *s, dt=torch.bfloat16: torch.empty(s, dtype=dt, device=dev)
        return dict(h=e(B, self.H), qkv=e(B, self.qkv_dim), att=e(B, self.n_heads * self.hd), act=e(B, self.I),
                    ssA=e(n_part, 16, dt=torch.float32), ssB=e(n_part, 16, dt=torch.float32), logits=e(B, self.V, dt=torch.float32),
                    hgA=e(B, self.H) if self.xg else None, hgB=e(B, self.H) if self.xg else None)

    def _forward_tokens(self, ids_dev, B, L, cache, n_beam, write_cache, past, logits_rows, past_dev=None,
                        n_splits=None, h=None, ws=None, scratch=None):
        """ids_dev int32 [B*L] on device. Returns fp32 logits [B, R, V] (R = L or 1).
        `scratch`: preallocated buffers of a decode session (decode_scratch), L == 1 only."""
        st = _stream()
        T, H = B * L, self.H
        dev = self.dev
        # decode (T <= 16): per-token sum-of-squares partials travel with the residual stream so the
        # RMSNorm fused into the next skinny GEMM never re-reads X: ss[c][16] written by the producer
        ssA = ssB = None
        if scratch is not None:
            ssA, ssB = scratch['ssA'], scratch['ssB']
        elif T <= 16:
            n_part = (H + 15) // 16
            ssA = torch.empty((n_part, 16), dtype=torch.float32, device=dev)
            ssB = torch.empty((n_part, 16), dtype=torch.float32, device=dev)
        ss_cur = None
        fused = T > 16 and self.pf_fused and self.gemm_impl == 0
        ss0 = torch.empty((T, 1), dtype=torch.float32, device=dev) if fused else None
        xg = self.xg and T <= 16 and h is None             # split RMSNorm (4-bit decode): gain-scaled copies of the residual stream
        hgA = hgB = None
        if xg:
            hgA = scratch['hgA'] if scratch is not None else torch.empty((T, H), dtype=torch.bfloat16, device=dev)
            hgB = scratch['hgB'] if scratch is not None else torch.empty((T, H), dtype=torch.bfloat16, device=dev)
        if h is None:
            h = scratch['h'] if scratch is not None else torch.empty((T, H), dtype=torch.bfloat16, device=dev)
            if xg:
                call('p3_embed_gather_xg', ptr(self.embed), ptr(ids_dev), ptr(h), T, H, self.V, ptr(ssA), ptr(self.layers[0]['ln1']),
                     ptr(hgA), st)
            else:
                call('p3_embed_gather', ptr(self.embed), ptr(ids_dev), ptr(h), T, H, self.V, ptr(ss0 if fused else ssA), st)
            ss_cur = None if ssA is None else ssA[:1]
        elif fused:
            call('p3_row_sumsq', ptr(h), h.stride(0), ptr(ss0), T, H, st)
        if scratch is not None:
            qkv, att, act = scratch['qkv'], scratch['att'], scratch['act']
        else:
            qkv = torch.empty((T, self.qkv_dim), dtype=torch.bfloat16, device=dev)
            att = torch.empty((T, self.n_heads * self.hd), dtype=torch.bfloat16, device=dev)
            act = torch.empty((T, self.I), dtype=torch.bfloat16, device=dev)
        use_decode_attn = L <= 16 and cache is not None
        if use_decode_attn:
            if n_splits is None:
                n_splits = self._splits(cache, B, max(1, (past + PAGE - 1) // PAGE), L)
            if n_splits > 1 and ws is None:
                ws = torch.zeros(_lib.lib().p3_attention_decode_workspace(B, L, self.n_heads, self.hd, n_splits) // 4,
                                 dtype=torch.float32, device=dev)        # zero: holds the split-arrival counters
        qoff, koff, voff = 0, self.n_heads * self.hd * 2, (self.n_heads + self.n_kv) * self.hd * 2
        if cache is not None:
            cosT, sinT, tbs = cache.cos, cache.sin, cache.tab_bstride
            bt, bts, kvs = cache.block_table, cache.block_table.stride(0), cache.kv_start
        else:
            cosT, sinT, tbs, bt, bts, kvs = self._nc_cos, self._nc_sin, self._nc_tbs, None, 0, self._nc_kvs
        if fused:
            # per-row sum-of-squares partials travel with the residual stream: [T,1] for the input rows (embed_gather or one
            # pass over the spliced h), then [T, H/32] written by every residual epilogue; the normed GEMMs turn them into the
            # RMSNorm row scale
            ssA = torch.empty((T, H // 32), dtype=torch.float32, device=dev)
            ssB = torch.empty((T, H // 32), dtype=torch.float32, device=dev)
            ss_cur = ss0
        for li, lw in enumerate(self.layers):
            pool = cache.pool[li] if cache is not None else None
            wc = 1 if (write_cache and cache is not None) else 0
            if fused:
                rope = dict(cosT=ptr(cosT), sinT=ptr(sinT), tab_bstride=tbs, L=L, n_heads=self.n_heads, n_kv=self.n_kv, hd=self.hd,
                            past=past, row_div=n_beam, write_cache=wc, bt_stride=bts, past_dev=ptr(past_dev), pool=ptr(pool),
                            block_table=ptr(bt))
                self.gemm_fused(h, lw['qkv_pf'], qkv, _lib.EPI_ROPE_KV, lw['plans']['qkv_pf'], ss_in=ss_cur, rope=rope)
                qp = qkv.data_ptr()
                if use_decode_attn and cache.quantized and cache.n_quant > 0:       # <= 16 new tokens per row (constrain / beam steps)
                    call('p3_attention_decode_q4', qp + qoff, qp + koff, qp + voff, self.qkv_dim, self.qkv_dim,
                         self.qkv_dim, ptr(att), att.stride(0), B, L, self.n_heads, self.n_kv, self.hd, self.scale,
                         past, cache.n_quant, ptr(kvs), ptr(pool), ptr(cache.qcodes[li]), ptr(cache.qmeta[li]), ptr(bt),
                         bts, n_beam, n_splits, ptr(ws), ptr(past_dev), None, 0, st)
                elif use_decode_attn:
                    call('p3_attention_decode', qp + qoff, qp + koff, qp + voff, self.qkv_dim, self.qkv_dim,
                         self.qkv_dim, ptr(att), att.stride(0), B, L, self.n_heads, self.n_kv, self.hd, self.scale,
                         past, ptr(kvs), ptr(pool), ptr(bt), bts, n_beam, n_splits, ptr(ws), ptr(past_dev), None, 0, st)
                else:
                    call('p3_attention_prefill', qp + qoff, qp + koff, qp + voff, self.qkv_dim, self.qkv_dim, self.qkv_dim,
                         ptr(att), att.stride(0), B, L, self.n_heads, self.n_kv, self.hd, self.scale, 1, past, ptr(kvs),
                         ptr(pool), ptr(bt), bts, n_beam, st)
                self.gemm_fused(att, lw['o'], h, _lib.EPI_RESIDUAL, lw['plans']['o'], resid=h, ss_out=ssB)
                self.gemm_fused(h, lw['gu_pf'], act, _lib.EPI_SWIGLU, lw['plans']['gu_pf'], ss_in=ssB)
                self.gemm_fused(act, lw['down'], h, _lib.EPI_RESIDUAL, lw['plans']['down'], resid=h, ss_out=ssA)
                ss_cur = ssA
                continue
            if 'qkv' in self._skip and T <= 16:
                pass
            elif T <= 16:                                   # decode: qkv_proj + rope + KV write in one launch
                if xg:
                    hq, nq, sq = hgA, None, ss_cur
                else:
                    hq, nq, sq = (self._prenorm(h, lw['ln1']), None, None) if self.prenorm else (h, lw['ln1'], ss_cur)
                # the qkv kernel parks the whole o_proj stream (18.9 MB bf16) in L2 before it waits on its predecessor; it
                # survives the evict-first KV stream of the attention kernel, which then carries no prefetch duty (A/B on the
                # bench: 2.851 -> 2.814 ms/token, attention in-step roofline 0.776 -> 0.816; P3_OPF=attn restores the old split)
                qo_pf, qo_bytes = (None, 0)
                if self._opf_in_qkv:
                    qo_pf, _ = self._prefetch_target(lw['o'], T)
                    qo_bytes = qo_pf.numel() * qo_pf.element_size()
                ev = self._ev()
                q4 = self._codes(lw['qkv'], T)
                if xg:
                    a = _lib.SkinnyArgs()
                    a.op, a.X, a.ldx, a.eps, a.out = 1, ptr(hq), hq.stride(0), self.eps, ptr(qkv)
                    if q4 is not None:
                        a.Wq, a.Wmeta = ptr(q4[0]), ptr(q4[1])
                    else:
                        a.W = ptr(lw['qkv'])
                    a.K, a.ss_in, a.n_ss_in, a.rs_epi = H, ptr(sq), sq.shape[0], 1
                    a.cosT, a.sinT, a.tab_bstride = ptr(cosT), ptr(sinT), tbs
                    a.B, a.L, a.n_heads, a.n_kv, a.hd, a.past = B, L, self.n_heads, self.n_kv, self.hd, past
                    a.past_dev, a.row_div, a.pool, a.block_table, a.bt_stride, a.write_cache = ptr(past_dev), n_beam, ptr(pool), ptr(bt), bts, wc
                    a.l2_prefetch, a.l2_prefetch_bytes = ptr(qo_pf), qo_bytes
                    _lib.call_struct('p3_gemm_skinny_x', a, st)
                    self._ev(ev, 'skinny', (q4[0].numel() + q4[1].numel() * 2) if q4 is not None else self.qkv_dim * H * 2)
                elif q4 is not None:
                    call('p3_gemm_skinny_qkv_rope_w4', ptr(hq), hq.stride(0), ptr(nq), self.eps, ptr(q4[0]), ptr(q4[1]),
                         ptr(qkv), ptr(sq), 0 if sq is None else sq.shape[0], ptr(cosT), ptr(sinT), tbs, B, L,
                         self.n_heads, self.n_kv, self.hd, H, past, ptr(past_dev), n_beam, ptr(pool), ptr(bt), bts, wc, ptr(qo_pf), qo_bytes, st)
                    self._ev(ev, 'skinny', q4[0].numel() + q4[1].numel() * 2)
                else:
                    call('p3_gemm_skinny_qkv_rope', ptr(hq), hq.stride(0), ptr(nq), self.eps, ptr(lw['qkv']), ptr(qkv),
                         ptr(sq), 0 if sq is None else sq.shape[0], ptr(cosT), ptr(sinT), tbs, B, L, self.n_heads,
                         self.n_kv, self.hd, H, past, ptr(past_dev), n_beam, ptr(pool), ptr(bt), bts, wc, ptr(qo_pf), qo_bytes, st)
                    self._ev(ev, 'skinny', self.qkv_dim * H * 2)
            else:
                self.linear(h, lw['qkv'], qkv, _lib.EPI_NONE, norm_w=lw['ln1'], ss_in=ss_cur)
                call('p3_rope_kvwrite', ptr(qkv), ptr(cosT), ptr(sinT), tbs, B, L, self.n_heads, self.n_kv, self.hd, past,
                     n_beam, ptr(pool), ptr(bt), bts, wc, ptr(past_dev), st)
            qp = qkv.data_ptr()
            ev = self._ev() if use_decode_attn else None
            o_pf, pf_bytes = self._prefetch_target(lw['o']) if T <= 16 else (None, 0)   # o_proj weights ride into L2 behind the KV stream
            if T <= 16:
                pf_bytes = o_pf.numel() * o_pf.element_size()
                if self._opf_in_qkv:
                    o_pf, pf_bytes = None, 0
            if use_decode_attn and 'attn' in self._skip:
                pass
            elif use_decode_attn:
                if cache.quantized and cache.n_quant > 0:
                    call('p3_attention_decode_q4', qp + qoff, qp + koff, qp + voff, self.qkv_dim, self.qkv_dim,
                         self.qkv_dim, ptr(att), att.stride(0), B, L, self.n_heads, self.n_kv, self.hd, self.scale,
                         past, cache.n_quant, ptr(kvs), ptr(pool), ptr(cache.qcodes[li]), ptr(cache.qmeta[li]), ptr(bt),
                         bts, n_beam, n_splits, ptr(ws), ptr(past_dev), ptr(o_pf), pf_bytes, st)
                else:
                    call('p3_attention_decode', qp + qoff, qp + koff, qp + voff, self.qkv_dim, self.qkv_dim,
                         self.qkv_dim, ptr(att), att.stride(0), B, L, self.n_heads, self.n_kv, self.hd, self.scale,
                         past, ptr(kvs), ptr(pool), ptr(bt), bts, n_beam, n_splits, ptr(ws), ptr(past_dev), ptr(o_pf),
                         pf_bytes, st)
            else:
                call('p3_attention_prefill', qp + qoff, qp + koff, qp + voff, self.qkv_dim, self.qkv_dim, self.qkv_dim,
                     ptr(att), att.stride(0), B, L, self.n_heads, self.n_kv, self.hd, self.scale, 1, past, ptr(kvs),
                     ptr(pool), ptr(bt), bts, n_beam, st)
            if ev is not None:
                self._ev(ev, 'attn', B * past * 2 * self.n_kv * self.hd * 2)
            # decode: every kernel parks (part of) its successor's weights in L2 so HBM never idles at a boundary
            nxt_qkv = self.layers[li + 1]['qkv'] if li + 1 < len(self.layers) else self.lm_head
            if xg:
                g_next = self.layers[li + 1]['ln1'] if li + 1 < len(self.layers) else self.norm
                if 'o' not in self._skip:
                    self.linear(att, lw['o'], h, _lib.EPI_RESIDUAL, resid=h, ss_out=ssB, nxt=lw['gu'] if self.pf_chain else None,
                                xg_gain=lw['ln2'], xg_out=hgB)
                if 'gu' not in self._skip:
                    self.linear(hgB, lw['gu'], act, _lib.EPI_SWIGLU, ss_in=ssB, nxt=lw['down'] if self.pf_chain else None, rs_epi=True)
                if 'down' not in self._skip:
                    self.linear(act, lw['down'], h, _lib.EPI_RESIDUAL, resid=h, ss_out=ssA, nxt=nxt_qkv if self.pf_chain else None,
                                xg_gain=g_next, xg_out=hgA)
                ss_cur = ssA
                continue
            if not ('o' in self._skip and T <= 16):
                self.linear(att, lw['o'], h, _lib.EPI_RESIDUAL, resid=h, ss_out=ssB, nxt=lw['gu'] if self.pf_chain else None)
            if not ('gu' in self._skip and T <= 16):
                self.linear(h, lw['gu'], act, _lib.EPI_SWIGLU, norm_w=lw['ln2'], ss_in=ssB, nxt=lw['down'] if self.pf_chain else None)
            if not ('down' in self._skip and T <= 16):
                self.linear(act, lw['down'], h, _lib.EPI_RESIDUAL, resid=h, ss_out=ssA, nxt=nxt_qkv if self.pf_chain else None)
            ss_cur = ssA
        if logits_rows == 'none':
            return None
        if logits_rows == 'last':
            hl = h.view(B, L, H)[:, -1, :]
            R = 1
        else:
            hl, R = h, L
        hl = hl.reshape(B * R, H) if hl.is_contiguous() else hl.contiguous().reshape(B * R, H)
        logits = scratch['logits'] if scratch is not None else torch.empty((B * R, self.V), dtype=torch.float32, device=dev)
        if xg and R == L:
            self.linear(hgA, self.lm_head, logits, _lib.EPI_F32, ss_in=ss_cur, rs_epi=True)
        else:
            self.linear(hl, self.lm_head, logits, _lib.EPI_F32, norm_w=self.norm, ss_in=ss_cur if (R == L) else None)
        return logits.view(B, R, self.V)

    # ------------------------------------------------------------------ reference call protocol (phi:606, 576-592)
    def __call__(self, input_ids, pixel_values=None, image_sizes=None, positions=None, cache=None, pids=None,
                 mask=None, max_tokens=0, advance_offset=None, n_beam=1, logits_rows='all', past_dev=None, n_splits=None):
        """`past_dev` / `n_splits` (extensions, used by the graph-captured constrain step): the cache offset is read from a device
        int32 at run time so that a captured call can be replayed as the cache grows (<= 16 new tokens per row)."""
        ids = torch.as_tensor(input_ids)
        if ids.dim() == 1:
            ids = ids[None]
        B, L = ids.shape
        ids_dev = ids.to(self.dev, torch.int32).contiguous().reshape(-1)
        h = None
        if pixel_values is not None and self.vision is not None:
            h = torch.empty((B * L, self.H), dtype=torch.bfloat16, device=self.dev)
            call('p3_embed_gather', ptr(self.embed), ptr(ids_dev), ptr(h), B * L, self.H, self.V, None, _stream())
            h = self._vision_embed(h, L, pixel_values, image_sizes, positions)
        if cache is None:
            L_all = L + max_tokens
            cos, sin, tbs = self._rope_table(L_all, pids)
            kvs = torch.zeros(B, dtype=torch.int32, device=self.dev)
            if mask is not None:
                m = torch.as_tensor(mask).cpu()
                kvs = (m.shape[1] - m.sum(dim=1)).to(self.dev, torch.int32)      # left-pad length per row
            if max_tokens < 1:                                                   # KVCache passthrough (phi:521-522)
                self._nc_cos, self._nc_sin, self._nc_tbs, self._nc_kvs = cos, sin, tbs, kvs
                logits = self._forward_tokens(ids_dev, B, L, None, 1, False, 0, logits_rows, h=h)
                return logits, None
            cache = KVCacheB200(self, B, L, max_tokens, self.use_quantized_cache)
            cache.set_tables(cos, sin, tbs, kvs)
        if n_beam > 1 and cache.quantized and not getattr(self.cfg, 'allow_beam_with_quantized_cache', False):
            raise NotImplementedError('Beam Search is not yet compatible with Quantized Cache')   # phi:524-525
        past = cache.offset
        if past + L > cache.S_max:
            raise ValueError(f'KV cache overflow: {past}+{L} > {cache.S_max}')
        write = n_beam == 1
        first_fill = write and past == 0
        if write and logits_rows == 'last' and L > self.prefill_chunk:
            # chunked prefill: keys [0, past+c0) come from the paged pool, the chunk's own keys from its
            # qkv buffer; activations stay O(chunk) so a 128K prompt fits (the reference materialises
            # L x L scores and an L x L mask and cannot run it)
            ids2 = ids_dev.view(B, L)
            h2 = None if h is None else h.view(B, L, self.H)
            logits = None
            for c0 in range(0, L, self.prefill_chunk):
                c1 = min(L, c0 + self.prefill_chunk)
                logits = self._forward_tokens(ids2[:, c0:c1].contiguous().reshape(-1), B, c1 - c0, cache, 1, True, past + c0,
                                              'last' if c1 == L else 'none',
                                              h=None if h2 is None else h2[:, c0:c1].contiguous().reshape(-1, self.H))
        else:
            logits = self._forward_tokens(ids_dev, B, L, cache, n_beam, write, past, logits_rows, h=h, past_dev=past_dev,
                                          n_splits=n_splits)
        if write:
            cache.offset = past + L                                               # phi:544-547
        if first_fill and cache.quantized:
            self._quantize_prompt(cache, L)
        if advance_offset is not None:
            cache.offset = past + advance_offset                                  # phi:589-591
        return logits, cache

    def _quantize_prompt(self, cache, n_tokens):
        """mx.quantize of the prompt K,V (phi:531-533): later steps read the 4-bit image."""
        nl, n_pages = cache.pool.shape[0], cache.pool.shape[1]
        if cache.qcodes is None:
            cache.qcodes = torch.zeros((nl, n_pages, 2, self.n_kv, PAGE, self.hd // 2), dtype=torch.uint8, device=self.dev)
            cache.qmeta = torch.zeros((nl, n_pages, 2, self.n_kv, PAGE, self.hd // 32, 2), dtype=torch.bfloat16,
                                      device=self.dev)
        for li in range(nl):
            call('p3_kv_quantize_q4g32', ptr(cache.pool[li]), ptr(cache.qcodes[li]), ptr(cache.qmeta[li]),
                 ptr(cache.block_table), cache.block_table.stride(0), cache.B, n_tokens, self.n_kv, self.hd, _stream())
        cache.n_quant = (n_tokens // PAGE) * PAGE

    # ------------------------------------------------------------------ device-resident greedy loop (pv:390-398)
    def decode_session(self, first_token, cache, max_steps, use_graph=True, sampler=None):
        return DecodeSession(self, first_token, cache, max_steps, use_graph, sampler)

    def greedy_decode(self, first_token, cache, n_steps, use_graph=True, eos_check_every=0):
        """n_steps decode steps from `first_token` [B] with no per-token host sync.
        Returns int32 history [B, steps_run+1] (first_token then every generated token)."""
        ses = DecodeSession(self, first_token, cache, n_steps, use_graph)
        for i in range(n_steps):
            ses.step()
            if eos_check_every and (i + 1) % eos_check_every == 0 and ses.all_eos():
                break
        return ses.finish()


class DecodeSession:
    """One CUDA graph = one greedy decode step (embed -> 32 layers -> lm_head -> argmax -> bookkeeping).
    `past`, the step counter, the current token and the token history live on the device, so
    replaying the graph advances generation with zero host synchronisation (the reference syncs
    at mx.eval and `eos_id in token` every token, pv:393,397,113). The state buffers and the graph
    live in the cache's slab and are reused by later sessions of the same shape."""

    def __init__(self, model, first_token, cache, max_steps, use_graph=True, sampler=None):
        self.m, self.cache, self.max_steps = model, cache, max_steps
        # sampler = None (greedy argmax, the reference's only mode) or (top_p, temperature, u_table [steps+1, B])
        self.sampler = sampler
        if sampler is not None:
            cache.slab.session = None          # a sampling graph is not interchangeable with the greedy one
        B = cache.B
        dev = model.dev
        self.B = B
        first_token = first_token.to(dev, torch.int32).reshape(B).contiguous()
        self.steps_run = 0
        self.graph = None
        self.launches_per_step = 0
        st = cache.slab.session if use_graph else None
        if st is not None and st['max_steps'] == max_steps and st['offset0'] == cache.offset:
            # recycled slab: same buffers, same graph -> just reset the device-side state
            self.hist, self.tok, self.step_dev, self.past_dev, self.eos = st['hist'], st['tok'], st['step'], st['past'], st['eos']
            self.n_splits, self.graph, self.launches_per_step = st['n_splits'], st['graph'], st['lps']
            self.mega_ses, self.ws, self.scratch = st['mega'], st['ws'], st['scratch']
            self._reset(first_token)
            return
        self.mega_ses = self.ws = self.scratch = None
        self.hist = torch.zeros((B, max_steps + 1), dtype=torch.int32, device=dev)
        self.tok = first_token.clone()
        self.step_dev = torch.ones(1, dtype=torch.int32, device=dev)
        self.past_dev = torch.full((1,), cache.offset, dtype=torch.int32, device=dev)
        self.eos = torch.zeros(B, dtype=torch.int32, device=dev)
        self._reset(first_token)
        if max_steps <= 0:
            return
        if cache.offset + max_steps > cache.S_max:
            raise ValueError('KV cache overflow: decode session longer than the cache was sized for')
        tiles = (cache.offset + max_steps + PAGE - 1) // PAGE
        self.n_splits = model._splits(cache, B, tiles)
        # decode scratch lives outside the captured step: split-KV partials + arrival counters (the kernel leaves the
        # counters at zero), and the static activation buffers / argument structs of the persistent layer kernel
        self.ws = None
        if self.n_splits > 1:
            self.ws = torch.zeros(_lib.lib().p3_attention_decode_workspace(B, 1, model.n_heads, model.hd, self.n_splits) // 4,
                                  dtype=torch.float32, device=dev)
        self.scratch = None
        if model.mega is not None and B <= 8:
            self.mega_ses = model.mega.session(cache, B)
            self.mega_ses.bind(self.past_dev)
        elif B <= 16:
            self.scratch = model.decode_scratch(B)
        if use_graph and B <= 16:
            cur = torch.cuda.current_stream()
            s = torch.cuda.Stream()
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                self._one_step()                              # warm-up: smem attributes, allocator pools
            cur.wait_stream(s)
            torch.cuda.synchronize()
            self._reset(first_token)                          # the warm-up step really executed: roll back
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launches
            with capture_graph(self.graph):
                self._one_step()
            self.launches_per_step = _lib.launches - n0
            if sampler is None:
                cache.slab.session = dict(max_steps=max_steps, offset0=cache.offset, hist=self.hist, tok=self.tok,
                                          step=self.step_dev, past=self.past_dev, eos=self.eos, n_splits=self.n_splits,
                                          graph=self.graph, lps=self.launches_per_step, mega=self.mega_ses, ws=self.ws,
                                          scratch=self.scratch)

    def _reset(self, first_token):
        self.hist[:, 0] = first_token
        self.tok.copy_(first_token)
        self.step_dev.fill_(1)
        self.past_dev.fill_(self.cache.offset)
        self.eos.copy_((first_token == 32007).to(torch.int32))

    def _mega_step(self):
        """embed -> [qkv] -> 32 x (decode attention -> [o_proj, gate_up, down, next qkv | lm_head]): 2 launches per layer"""
        m, ms, c, B = self.m, self.mega_ses, self.cache, self.B
        st = _stream()
        past = c.offset + self.steps_run
        call('p3_embed_gather', ptr(m.embed), ptr(self.tok), ptr(ms.h), B, m.H, m.V, ptr(ms.ss0), st)
        ev = m._ev()
        ms.launch(0, st)
        m._ev(ev, 'mega', ms.mega.shapes['qkv'][1] * ms.mega.shapes['qkv'][2] * 2)
        qp, hb = ms.qkv.data_ptr(), m.n_heads * m.hd * 2
        koff, voff = hb, (m.n_heads + m.n_kv) * m.hd * 2
        bt, bts = c.block_table, c.block_table.stride(0)
        nl = len(m.layers)
        per_layer = sum(N * K for k, (_, N, K) in ms.mega.shapes.items() if k != 'lm') * 2
        for li in range(nl):
            ev = m._ev()
            if c.quantized and c.n_quant > 0:
                call('p3_attention_decode_q4', qp, qp + koff, qp + voff, m.qkv_dim, m.qkv_dim, m.qkv_dim, ptr(ms.att),
                     ms.att.stride(0), B, 1, m.n_heads, m.n_kv, m.hd, m.scale, past, c.n_quant, ptr(c.kv_start), ptr(c.pool[li]),
                     ptr(c.qcodes[li]), ptr(c.qmeta[li]), ptr(bt), bts, 1, self.n_splits, ptr(self.ws), ptr(self.past_dev),
                     None, 0, st)
            else:
                call('p3_attention_decode', qp, qp + koff, qp + voff, m.qkv_dim, m.qkv_dim, m.qkv_dim, ptr(ms.att),
                     ms.att.stride(0), B, 1, m.n_heads, m.n_kv, m.hd, m.scale, past, ptr(c.kv_start), ptr(c.pool[li]), ptr(bt),
                     bts, 1, self.n_splits, ptr(self.ws), ptr(self.past_dev), None, 0, st)
            m._ev(ev, 'attn', B * past * 2 * m.n_kv * m.hd * 2)
            ev = m._ev()
            ms.launch(li + 1, st)
            qkv_b = ms.mega.shapes['qkv'][1] * ms.mega.shapes['qkv'][2] * 2
            m._ev(ev, 'mega', per_layer if li + 1 < nl else per_layer - qkv_b + m.V * m.H * 2)
        return ms.logits.view(B, 1, m.V)

    def _one_step(self):
        m = self.m
        if self.mega_ses is not None:
            logits = self._mega_step()
        else:
            logits = m._forward_tokens(self.tok, self.B, 1, self.cache, 1, True, self.cache.offset + self.steps_run, 'last',
                                       past_dev=self.past_dev, n_splits=self.n_splits, ws=self.ws, scratch=self.scratch)
        if self.sampler is None:
            call('p3_row_stats', ptr(logits), self.B, m.V, m.V, ptr(self.tok), None, None, 0, None, None, 0, None, None,
                 _stream())
        else:
            top_p, temp, u = self.sampler
            call('p3_top_p_sample', ptr(logits), self.B, m.V, m.V, float(top_p), float(temp), ptr(u), ptr(self.tok), None,
                 ptr(self.step_dev), u.stride(0), _stream())
        call('p3_decode_advance', ptr(self.tok), ptr(self.hist), self.hist.stride(0), self.B, ptr(self.step_dev),
             ptr(self.past_dev), ptr(self.eos), _stream())

    def step(self):
        if self.steps_run >= self.max_steps:
            raise ValueError('decode session exhausted')
        if self.graph is not None:
            self.graph.replay()
            _lib.launches += self.launches_per_step
        else:
            self._one_step()
        self.steps_run += 1

    def all_eos(self):
        return bool(self.eos.all().item())                    # host sync (off the per-token path)

    def last_token(self):
        return self.tok

    def finish(self):
        self.cache.offset = self.cache.offset + self.steps_run
        return self.hist[:, :self.steps_run + 1].clone()      # the slab (and its hist buffer) is recycled
