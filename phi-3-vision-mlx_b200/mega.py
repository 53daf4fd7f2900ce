"""Host side of the persistent decode-layer kernel (csrc/decode_mega.cu, C ABI p3_decode_mega / p3_mega_pack).

For <= 8 decode rows the per-layer weight stream
    o_proj(+residual) -> RMSNorm -> gate_up_proj(+SwiGLU) -> down_proj(+residual) -> RMSNorm -> qkv_proj(+SuRoPE, +KV write)
(/root/reference/phi.py:437-438, 442-453, 460, 465-471, 478-485; the last launch ends in lm_head, phi.py:604-608) runs as ONE
launch of #SM persistent CTAs. This module owns
  * the stream-order copy of the decoder weights (made once at load by p3_mega_pack),
  * the tile -> CTA schedule (cumulative-byte balanced at every phase boundary),
  * the argument structs of the 33 launches of a decode step.
`pack_reference` / `emulate_phase` restate the packing and the kernel's consumption order with torch on the CPU; they exist so
that the index maps can be tested without a GPU (tests/test_host_cpu.py) and are not used by the product path.
"""
import ctypes as C
import heapq
import torch
from . import _lib
from ._lib import ptr

RESID, SWIGLU, QKV_ROPE, F32 = range(4)
WARPS, KBLOCK, XF, MAX_PART, MAX_PHASES = 8, 4096, 32, 8, 4


class MegaPhase(C.Structure):
    _fields_ = [('wp', C.c_void_p), ('kind', C.c_int32), ('N', C.c_int32), ('K', C.c_int32), ('max_tiles_per_cta', C.c_int32),
                ('x', C.c_void_p), ('ldx', C.c_int64), ('norm_w', C.c_void_p), ('ss_in', C.c_void_p), ('n_ss_in', C.c_int32),
                ('_pad', C.c_int32), ('ss_out', C.c_void_p), ('out', C.c_void_p), ('ldo', C.c_int64),
                ('cta_off', C.c_void_p), ('tile_ids', C.c_void_p)]


class MegaArgs(C.Structure):
    _fields_ = [('ph', MegaPhase * MAX_PHASES), ('n_phases', C.c_int32), ('M', C.c_int32), ('eps', C.c_float),
                ('n_ctas', C.c_int32), ('sync', C.c_void_p), ('cosT', C.c_void_p), ('sinT', C.c_void_p),
                ('tab_bstride', C.c_int64), ('n_heads', C.c_int32), ('n_kv', C.c_int32), ('hd', C.c_int32),
                ('past', C.c_int32), ('past_dev', C.c_void_p), ('pool', C.c_void_p), ('block_table', C.c_void_p),
                ('bt_stride', C.c_int32), ('_pad2', C.c_int32), ('dbg', C.c_void_p)]


def mt_of(kind):
    return 1 if kind == RESID else 2


def dims(kind, N, K):
    """(MT, tiles, K-blocks, 16-wide k blocks per warp and K-block) or None when the kernel does not cover the shape"""
    MT = mt_of(kind)
    if N <= 0 or K <= 0 or N % (16 * MT):
        return None
    n_kblk = -(-K // KBLOCK)
    if K % (n_kblk * 128) or K // (n_kblk * 128) > XF:
        return None
    return MT, N // (16 * MT), n_kblk, K // (n_kblk * 128)


def tile_rows(kind, N, n_heads=0, n_kv=0, hd=0):
    """int64 [T, MT, 16]: the W rows of every tile (mirror of mg_tile_row in decode_mega.cu)"""
    MT = mt_of(kind)
    T = N // (16 * MT)
    ti = torch.arange(T)[:, None]
    mt = torch.arange(MT)[None, :]
    if kind == RESID:
        base = 16 * ti + 0 * mt
    elif kind == SWIGLU:
        base = mt * (N // 2) + 16 * ti
    elif kind == QKV_ROPE:
        gpr, n_rope = hd // 32, (n_heads + n_kv) * (hd // 32)
        rope = (ti // gpr) * hd + 16 * (ti % gpr) + mt * (hd // 2)
        plain = (n_heads + n_kv) * hd + 32 * (ti - n_rope) + 16 * mt
        base = torch.where(ti < n_rope, rope, plain)
    else:
        base = 32 * ti + 16 * mt
    return base[:, :, None] + torch.arange(16)[None, None, :]


def k_offsets(nkb_w):
    """int64 [nkb_w, 4 (t), 4]: offsets inside a warp's K slice of the 4 consecutive elements lane t feeds to k block j
    (mirror of mg_koff): blocks are paired so one 16-byte X load serves two MMAs; an unpaired last block is plain."""
    j = torch.arange(nkb_w)[:, None, None]
    t = torch.arange(4)[None, :, None]
    i = torch.arange(4)[None, None, :]
    paired = (j | 1) < nkb_w
    return torch.where(paired, 32 * (j >> 1) + 8 * t + 4 * (j & 1), 16 * j + 4 * t) + i


def pack_reference(W, kind, n_heads=0, n_kv=0, hd=0):
    """CPU restatement of p3_mega_pack: [kb][tile][warp][j][mt][lane = g*4+t][a0 a1 a2 a3 pairs] (flat, same dtype)"""
    N, K = W.shape
    MT, T, n_kblk, nkb_w = dims(kind, N, K)
    rows = tile_rows(kind, N, n_heads, n_kv, hd)                         # [T, MT, 16]
    kk = (torch.arange(n_kblk * WARPS)[:, None, None, None] * (nkb_w * 16) + k_offsets(nkb_w)[None]).reshape(-1)
    Wt = W[rows.reshape(-1)][:, kk].reshape(T, MT, 2, 8, n_kblk, WARPS, nkb_w, 4, 2, 2)   # [T, mt, hi, g, kb, warp, j, t, p, e]
    return Wt.permute(4, 0, 5, 6, 1, 3, 7, 8, 2, 9).contiguous().reshape(-1)       # [kb, T, warp, j, mt, g, t, p, hi, e]


def build_schedule(phases, n_ctas):
    """phases: list of (kind, N, K). Greedy: every tile goes to the CTA with the least cumulative bytes so far (ties: lowest
    index), phase after phase, so the byte counts of all CTAs agree to within one tile at EVERY phase boundary — what keeps
    all rings streaming into the grid barriers. Returns [(cta_off int32 [n_ctas+1], tile_ids int32 [T], max_tiles_per_cta)]."""
    load = [(0, c) for c in range(n_ctas)]
    heapq.heapify(load)
    out = []
    for kind, N, K in phases:
        MT, T, n_kblk, nkb_w = dims(kind, N, K)
        cost = MT * 16 * K * 2
        mine = [[] for _ in range(n_ctas)]
        for ti in range(T):
            l, c = heapq.heappop(load)
            mine[c].append(ti)
            heapq.heappush(load, (l + cost, c))
        off = [0]
        for m in mine:
            off.append(off[-1] + len(m))
        ids = [ti for m in mine for ti in m]
        out.append((torch.tensor(off, dtype=torch.int32), torch.tensor(ids, dtype=torch.int32), max(len(m) for m in mine)))
    return out


def emulate_phase(packed, kind, N, K, x, cta_off, tile_ids, n_heads=0, n_kv=0, hd=0):
    """CPU model of how the kernel consumes `packed`: for every CTA, K-block, tile, warp, k block j and m-tile it takes the
    next 512-byte fragment, pairs lane (g,t)'s a0..a3 with the X values the kernel loads for that lane (x[n][kbase+16j+4t..+3])
    and accumulates D[row][n]. Returns fp32 [M, N] = x @ W^T in W-row order (before any epilogue)."""
    MT, T, n_kblk, nkb_w = dims(kind, N, K)
    M = x.shape[0]
    rows = tile_rows(kind, N, n_heads, n_kv, hd)
    y = torch.zeros(M, N, dtype=torch.float32)
    seg = nkb_w * MT * 256                                              # elements per (kb, tile, warp) segment
    xf = x.to(torch.float32)
    n_ctas = cta_off.numel() - 1
    for c in range(n_ctas):
        for kb in range(n_kblk):
            for li in range(int(cta_off[c]), int(cta_off[c + 1])):
                ti = int(tile_ids[li])
                for w in range(WARPS):
                    s0 = ((kb * T + ti) * WARPS + w) * seg
                    fr = packed[s0:s0 + seg].to(torch.float32).reshape(nkb_w, MT, 8, 4, 2, 2, 2)    # [j, mt, g, t, p, hi, e]
                    kbase = (kb * WARPS + w) * nkb_w * 16
                    xs = xf[:, kbase + k_offsets(nkb_w).reshape(-1)].reshape(M, nkb_w, 4, 2, 2)   # [n, j, t, p, e]: the lane's X loads
                    d = torch.einsum('jmgtphe,njtpe->mhgn', fr, xs)                                # [mt, hi, g, n]
                    for mt in range(MT):
                        r = rows[ti, mt]                                                            # 16 rows: hi*8 + g
                        y[:, r] += d[mt].reshape(16, M).T
    return y


class MegaDecoder:
    """Stream-order weights + schedules of one model. `session()` binds them to a KV cache's static buffers."""

    def __init__(self, model, raw_weights):
        cfg = model.cfg
        self.m = model
        L = _lib.lib()
        self.n_ctas = L.p3_decode_mega_ctas()
        if self.n_ctas <= 0:
            raise RuntimeError('p3_decode_mega_ctas failed')
        H, I, V, nl = model.H, model.I, model.V, cfg.num_hidden_layers
        self.shapes = dict(qkv=(QKV_ROPE, model.qkv_dim, H), o=(RESID, H, model.n_heads * model.hd), gu=(SWIGLU, 2 * I, H),
                           down=(RESID, H, I), lm=(F32, V, H))
        for k, (kind, N, K) in self.shapes.items():
            if dims(kind, N, K) is None or (kind == QKV_ROPE and model.hd % 32):
                raise ValueError(f'decode_mega does not cover {k}: N={N} K={K}')
        st = torch.cuda.current_stream().cuda_stream
        dev = model.dev

        def pack(w, key):
            kind, N, K = self.shapes[key]
            w = w.to(dev, torch.bfloat16).contiguous()
            assert tuple(w.shape) == (N, K), (key, tuple(w.shape), (N, K))
            out = torch.empty(N * K, dtype=torch.bfloat16, device=dev)
            _lib.call('p3_mega_pack', ptr(w), ptr(out), kind, N, K, model.n_heads, model.n_kv, model.hd, st)
            return out
        self._pack = pack
        self.layers = []
        for i in range(nl):
            p = f'model.layers.{i}.'
            self.layers.append(dict(qkv=pack(raw_weights[p + 'self_attn.qkv_proj.weight'], 'qkv'),
                                    o=pack(raw_weights[p + 'self_attn.o_proj.weight'], 'o'),
                                    gu=pack(raw_weights[p + 'mlp.gate_up_proj.weight'], 'gu'),       # checkpoint order: gate | up
                                    down=pack(raw_weights[p + 'mlp.down_proj.weight'], 'down')))
        self.lm = pack(raw_weights['lm_head.weight'], 'lm')
        torch.cuda.synchronize()
        # launch kinds: first = [qkv], mid = [o, gu, down, qkv], last = [o, gu, down, lm]
        self.sched = {}
        for name, keys in (('first', ['qkv']), ('mid', ['o', 'gu', 'down', 'qkv']), ('last', ['o', 'gu', 'down', 'lm'])):
            sc = build_schedule([self.shapes[k] for k in keys], self.n_ctas)
            for (kind, N, K), (off, ids, mx) in zip([self.shapes[k] for k in keys], sc):
                if K > KBLOCK and mx > MAX_PART:
                    raise ValueError('decode_mega: too many tiles per CTA for a multi-K-block phase')
            self.sched[name] = [(off.to(dev), ids.to(dev), mx) for off, ids, mx in sc]
        self.sync = torch.zeros(2, dtype=torch.int32, device=dev)
        self.stream_bytes_per_step = 2 * (nl * sum(N * K for k, (_, N, K) in self.shapes.items() if k != 'lm') + V * H)

    def repack(self, li, key, w):
        """rewrite the stream-order copy of one matrix in place (Phi3B200.set_adapter)"""
        self.layers[li][key].copy_(self._pack(w, key))

    def session(self, cache, B):
        return MegaSession(self, cache, B)


class MegaSession:
    """Static activation buffers + the 33 argument structs of one decode step for a (cache, batch)."""

    def __init__(self, mega, cache, B):
        m = mega.m
        self.mega, self.B = mega, B
        dev = m.dev
        H, I, V = m.H, m.I, m.V
        z = lambda *s, dt=torch.bfloat16: torch.zeros(s, dtype=dt, device=dev)
        self.h, self.qkv, self.att, self.act = z(B, H), z(B, m.qkv_dim), z(B, m.n_heads * m.hd), z(B, I)
        self.logits = z(B, V, dt=torch.float32)
        n_part = max(1, H // 16)
        self.ss0, self.ssA, self.ssB = z(1, 16, dt=torch.float32), z(n_part, 16, dt=torch.float32), z(n_part, 16, dt=torch.float32)
        nl = len(mega.layers)
        self.args = []
        lay = m.layers
        for li in range(nl + 1):
            a = MegaArgs()
            a.M, a.eps, a.n_ctas, a.sync = B, m.eps, mega.n_ctas, ptr(mega.sync)
            a.cosT, a.sinT, a.tab_bstride = ptr(cache.cos), ptr(cache.sin), cache.tab_bstride
            a.n_heads, a.n_kv, a.hd, a.past, a.past_dev = m.n_heads, m.n_kv, m.hd, 0, None
            a.block_table, a.bt_stride = ptr(cache.block_table), cache.block_table.stride(0)
            a.pool = ptr(cache.pool[li]) if li < nl else None
            phases = []
            if li > 0:
                pw = mega.layers[li - 1]
                phases += [self._phase('o', pw['o'], self.att, None, None, 0, self.ssB, self.h),
                           self._phase('gu', pw['gu'], self.h, lay[li - 1]['ln2'], self.ssB, n_part, None, self.act),
                           self._phase('down', pw['down'], self.act, None, None, 0, self.ssA, self.h)]
            ss_in, n_ss = (self.ss0, 1) if li == 0 else (self.ssA, n_part)
            if li < nl:
                phases.append(self._phase('qkv', mega.layers[li]['qkv'], self.h, lay[li]['ln1'], ss_in, n_ss, None, self.qkv))
            else:
                phases.append(self._phase('lm', mega.lm, self.h, m.norm, ss_in, n_ss, None, self.logits))
            sched = mega.sched['first' if li == 0 else ('mid' if li < nl else 'last')]
            for i, (ph, (off, ids, mx)) in enumerate(zip(phases, sched)):
                ph.cta_off, ph.tile_ids, ph.max_tiles_per_cta = ptr(off), ptr(ids), mx
                a.ph[i] = ph
            a.n_phases = len(phases)
            self.args.append(a)

    def _phase(self, key, wp, x, norm_w, ss_in, n_ss, ss_out, out):
        kind, N, K = self.mega.shapes[key]
        ph = MegaPhase()
        ph.wp, ph.kind, ph.N, ph.K = ptr(wp), kind, N, K
        ph.x, ph.ldx, ph.norm_w = ptr(x), x.stride(0), ptr(norm_w)
        ph.ss_in, ph.n_ss_in, ph.ss_out = ptr(ss_in), n_ss, ptr(ss_out)
        ph.out, ph.ldo = ptr(out), out.stride(0)
        return ph

    def bind(self, past_dev):
        for a in self.args:
            a.past_dev = ptr(past_dev)

    def launch(self, li, stream):
        _lib.call_struct('p3_decode_mega', self.args[li], stream)
