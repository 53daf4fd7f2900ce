"""Processors: tokenise / left-pad (Phi3FProcessor, /root/reference/phi.py:228-250), image-token
splice (Phi3VProcessor._merge, phi.py:263-281) and the HD transform (Phi3VImageProcessor,
phi.py:283-372) — the latter on the GPU through p3_hd_* kernels, with the PIL resampling
coefficient tables built on the host exactly as Pillow's Resample.c precompute_coeffs does.

The reference obtains its tokenizer from the downloaded checkpoint (phi.py:230); there are no
tokenizer files offline, so the tokenizer is injected. Any object with the HF surface used by
the reference works: __call__(text|list).input_ids, encode(text, add_special_tokens=...),
decode(ids), batch_decode(list_of_ids).
"""
import math
import re
import numpy as np
import torch
from ._lib import call, ptr

IMAGE_MEAN = np.array([0.48145466, 0.4578275, 0.40821073])       # phi.py:286
IMAGE_STD = np.array([0.26862954, 0.26130258, 0.27577711])       # phi.py:287


class ByteTokenizer:
    """Deterministic stand-in tokenizer (BOS=1, byte b -> id b+3, specials for the chat tags).
    Only for tests / synthetic benchmarks where the Phi-3 sentencepiece files are unavailable."""
    SPECIAL = {'<|user|>': 32010, '<|assistant|>': 32001, '<|end|>': 32007, '<|system|>': 32006}

    class _Enc:
        def __init__(self, ids):
            self.input_ids = ids

    def _enc(self, text, bos=True):
        ids = [1] if bos else []
        pat = '(' + '|'.join(re.escape(k) for k in self.SPECIAL) + ')'
        for part in re.split(pat, text):
            if part in self.SPECIAL:
                ids.append(self.SPECIAL[part])
            else:
                ids.extend(b + 3 for b in part.encode('utf-8'))
        return ids

    def __call__(self, texts):
        if isinstance(texts, str):
            return self._Enc(self._enc(texts))
        return self._Enc([self._enc(t) for t in texts])

    def encode(self, text, add_special_tokens=True):
        return self._enc(text, bos=True)      # HF LlamaTokenizer.encode keeps a leading piece; callers drop [0]

    def decode(self, ids):
        inv = {v: k for k, v in self.SPECIAL.items()}
        out, buf = [], bytearray()
        for i in ids:
            i = int(i)
            if i in inv:
                out.append(buf.decode('utf-8', 'replace')); buf = bytearray(); out.append(inv[i])
            elif 3 <= i < 259:
                buf.append(i - 3)
        out.append(buf.decode('utf-8', 'replace'))
        return ''.join(out)

    def batch_decode(self, seqs):
        return [self.decode(s) for s in seqs]


class Phi3FProcessor:
    """phi.py:228-250."""

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer

    def _tokenize(self, texts):
        if isinstance(texts, str):
            return {'input_ids': torch.tensor(self.tokenizer(texts).input_ids, dtype=torch.int64)[None]}
        ids = self.tokenizer(texts).input_ids
        n = max(len(s) for s in ids)
        pids = [[1] * (n - len(s)) + list(range(len(s))) for s in ids]          # pad position id is 1 (H8)
        mask = [[0] * (n - len(s)) + [1] * len(s) for s in ids]
        ids = [[0] * (n - len(s)) + list(s) for s in ids]                         # left pad with id 0
        t = lambda x: torch.tensor(x, dtype=torch.int64)
        return {'input_ids': t(ids), 'pids': t(pids), 'mask': t(mask)}

    def __call__(self, texts, images=None):
        if images is not None:
            print('WARNING: You are using phi3_mini_128k. Use phi3_v for VLM tasks.')
        return self._tokenize(texts)


class Phi3VProcessor(Phi3FProcessor):
    """phi.py:252-281."""

    def __init__(self, tokenizer, num_crops=16, device='cuda'):
        super().__init__(tokenizer)
        self.img_processor = Phi3VImageProcessor(num_crops=num_crops, device=device)

    def __call__(self, texts, images=None):
        if images is None:
            return self._tokenize(texts)
        return self._merge(self.img_processor(images), texts)

    def _merge(self, images, texts):
        pattern = r"<\|image_\d+\|>"
        chunks = self.tokenizer(re.split(pattern, texts)).input_ids              # each chunk gets its own BOS
        n_tok = images['num_img_tokens']
        tags = re.findall(pattern, texts)
        iids = [int(s.split("|")[1].split("_")[-1]) for s in tags]
        pads = [[-i] * n_tok[i - 1] for i in iids]
        if len(chunks) > len(pads):
            pads = pads + [[]]
        ids = []
        for c, p in zip(chunks, pads):
            ids.extend(c)
            ids.extend(p)
        ids = np.array(ids, dtype=np.int64)[None]
        return {'input_ids': torch.from_numpy(ids), 'pixel_values': images['pixel_values'],
                'image_sizes': torch.tensor(images['image_sizes'], dtype=torch.int64),
                'positions': torch.from_numpy(np.argwhere(ids < 0))}


# ---------------------------------------------------------------------------------------------
def pil_bilinear_coeffs(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle)
    filter over the full box [0, in_size). Returns (bounds int32 [out,2], kk int32 [out,ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = w.sum(axis=1, keepdims=True)
    k = np.where(ww != 0, w / np.where(ww == 0, 1.0, ww), w)
    kk = np.where(k < 0, -0.5 + k * (1 << 22), 0.5 + k * (1 << 22))
    kk = np.trunc(kk).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk, ksize


def interp336_tables(in_size):
    """Weights/indices of the reference's 2-tap 'interpolate_336' (phi.py:333-359)."""
    def cubic(x):
        a = np.abs(x); a2 = a ** 2; a3 = a ** 3
        return ((1.5 * a3 - 2.5 * a2 + 1) * (a <= 1) + (-0.5 * a3 + 2.5 * a2 - 4 * a + 2) * ((a > 1) & (a <= 2)))
    scale = 336 / in_size
    out_c = np.linspace(0, in_size - 1, 336)
    in_c = out_c / scale
    left = np.floor(in_c - 0.5).astype(np.int32)
    right = left + 1
    left = np.clip(left, 0, in_size - 1)
    right = np.clip(right, 0, in_size - 1)
    wgt = np.zeros((336, 2), dtype=np.float32)
    wgt[:, 0] = cubic(in_c - left)
    wgt[:, 1] = cubic(right - in_c)
    s = wgt[:, 0] + wgt[:, 1]                                   # float32 sum, as weights[i].sum() (taps 2,3 are 0)
    nz = s != 0
    wgt[nz] = wgt[nz] / s[nz, None]
    idx = np.stack([left, right], axis=1).astype(np.int32)
    return idx, wgt


def hd_geometry(w, h, num_crops):
    """Pure index math of HD_transform (phi.py:291-306): returns dict with the resize target,
    padding and final [H, W]. Bit-exact contract."""
    trans = w < h
    if trans:
        w, h = h, w
    scale = int(np.sqrt(num_crops * w / h))
    new_w, new_h = int(scale * 336), int(scale * 336 * h / w)
    diff = int(np.ceil(new_h / 336) * 336) - new_h
    top = int(diff / 2)
    padded_h = new_h + diff
    H, W = (new_w, padded_h) if trans else (padded_h, new_w)
    n_tok = int((H // 336 * W // 336 + 1) * 144 + 1 + (H // 336 + 1) * 12)
    return dict(trans=trans, in_w=w, in_h=h, new_w=new_w, new_h=new_h, top=top, padded_h=padded_h, H=H, W=W,
                num_img_tokens=n_tok)


class Phi3VImageProcessor:
    """phi.py:283-372 on the GPU. `num_crops` is hard-coded 16 in the reference (phi.py:285);
    it is a parameter here (default 16) because BASELINE config 2 names num_crops=4."""

    def __init__(self, num_crops=16, device='cuda'):
        self.num_crops = num_crops
        self.device = torch.device(device)
        self.lut = torch.from_numpy(((np.arange(256)[:, None] / 255.0 - IMAGE_MEAN) / IMAGE_STD)).contiguous()
        self._lut_dev = None
        self._tables = {}            # device-resident coefficient / tap tables, keyed by geometry

    def _one(self, img):
        dev = self.device
        st = torch.cuda.current_stream().cuda_stream
        if hasattr(img, 'convert'):
            img = np.asarray(img.convert('RGB'))
        if isinstance(img, torch.Tensor):                     # uint8 HWC, host (pinned) or already on the device
            arr = img.to(dev, torch.uint8, non_blocking=True).contiguous()
        else:
            arr = torch.from_numpy(np.ascontiguousarray(img, dtype=np.uint8)).to(dev)
        h0, w0 = arr.shape[:2]
        g = hd_geometry(w0, h0, self.num_crops)
        # logical (possibly transposed) source view: element strides
        sy, sx = (3, w0 * 3) if g['trans'] else (w0 * 3, 3)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        def coeffs(n_in, n_out):
            key = ('pil', n_in, n_out)
            if key not in self._tables:
                b_, k_, ks_ = pil_bilinear_coeffs(n_in, n_out)
                self._tables[key] = (t(b_), t(k_), ks_)
            return self._tables[key]

        def taps(n):
            key = ('i336', n)
            if key not in self._tables:
                i_, w_ = interp336_tables(n)
                self._tables[key] = (t(i_), t(w_))
            return self._tables[key]
        bh_d, kh_d, ksh = coeffs(g['in_w'], g['new_w'])
        bv_d, kv_d, ksv = coeffs(g['in_h'], g['new_h'])
        tmp = torch.empty((g['in_h'], g['new_w'], 3), dtype=torch.uint8, device=dev)
        call('p3_hd_resize_h', ptr(arr), sy, sx, g['in_w'], g['in_h'], ptr(tmp), g['new_w'], ptr(bh_d), ptr(kh_d), ksh, st)
        H, W = g['H'], g['W']
        out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
        call('p3_hd_resize_v_pad', ptr(tmp), g['new_w'], g['in_h'], g['new_h'], ptr(bv_d), ptr(kv_d), ksv, g['top'],
             g['padded_h'], 1 if g['trans'] else 0, ptr(out), st)
        (hi_d, hw_d), (wi_d, ww_d) = taps(H), taps(W)
        if self._lut_dev is None:
            self._lut_dev = self.lut.to(dev)
        n = (H // 336) * (W // 336) + 1
        pv = torch.empty((n, 3, 336, 336), dtype=torch.float32, device=dev)
        call('p3_hd_tile_crops', ptr(out), H, W, ptr(self._lut_dev), ptr(pv), ptr(hi_d), ptr(hw_d), ptr(wi_d), ptr(ww_d), st)
        return pv, [H, W], g['num_img_tokens'], out

    def __call__(self, images, max_crops=None):
        if not isinstance(images, (list, tuple)):
            images = [images]
        max_crops = max_crops or (self.num_crops + 1)            # reference: 17 = 16 + 1 (phi.py:311)
        pvs, shapes, ntok = [], [], []
        for img in images:
            pv, shp, nt, _ = self._one(img)
            if pv.shape[0] < max_crops:                          # pad_to_max_num_crops_tensor (phi.py:311-316)
                pad = torch.zeros((max_crops - pv.shape[0], 3, 336, 336), dtype=pv.dtype, device=pv.device)
                pv = torch.cat([pv, pad], 0)
            pvs.append(pv); shapes.append(shp); ntok.append(nt)
        return {'pixel_values': torch.stack(pvs, 0), 'image_sizes': shapes, 'num_img_tokens': ntok}
