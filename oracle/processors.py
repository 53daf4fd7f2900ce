"""CPU ORACLE (test infrastructure only) — numpy/PIL restatement of the reference processors:
Phi3FProcessor._tokenize (phi.py:233-245), Phi3VProcessor._merge (phi.py:263-281) and
Phi3VImageProcessor incl. HD_transform / pad_to_336 / interpolate_336 (phi.py:283-372).
Pinned against the reference's own code executed with MLX stubbed: tests/golden/processor_golden.json
(made by tests/golden/make_golden.py). PIL is the same third-party resampler the reference calls.
"""
import re
import numpy as np
from PIL import Image, ImageOps

MEAN = np.array([0.48145466, 0.4578275, 0.40821073])
STD = np.array([0.26862954, 0.26130258, 0.27577711])


def tokenize(tokenizer, texts):
    """phi.py:233-245: str -> ids[1,L]; list -> left pad (id 0), pids (pad slots = 1), mask."""
    if isinstance(texts, str):
        return {'input_ids': np.array(tokenizer(texts).input_ids)[None]}
    ids = tokenizer(texts).input_ids
    n = max(len(s) for s in ids)
    return {'input_ids': np.array([[0] * (n - len(s)) + list(s) for s in ids]),
            'pids': np.array([[1] * (n - len(s)) + list(range(len(s))) for s in ids]),
            'mask': np.array([[0] * (n - len(s)) + [1] * len(s) for s in ids])}


def merge(tokenizer, image_inputs, text):
    """phi.py:263-281."""
    pat = r"<\|image_\d+\|>"
    chunks = tokenizer(re.split(pat, text)).input_ids
    ntok = image_inputs['num_img_tokens']
    iids = [int(s.split("|")[1].split("_")[-1]) for s in re.findall(pat, text)]
    pads = [[-i] * ntok[i - 1] for i in iids]
    if len(chunks) > len(pads):
        pads = pads + [[]]
    ids = []
    for c, p in zip(chunks, pads):
        ids.extend(c)
        ids.extend(p)
    ids = np.array(ids)[None]
    return {'input_ids': ids, 'pixel_values': image_inputs['pixel_values'],
            'image_sizes': np.array(image_inputs['image_sizes']), 'positions': np.argwhere(ids < 0)}


def hd_transform_u8(img, num_crops=16):
    """phi.py:290-308: returns the padded uint8 HWC image (before normalisation)."""
    img = img.convert('RGB')
    w, h = img.size
    trans = w < h
    if trans:
        img = img.transpose(Image.TRANSPOSE)
        w, h = img.size
    scale = int(np.sqrt(num_crops * w / h))
    img = img.resize([int(scale * 336), int(scale * 336 * h / w)], Image.BILINEAR)
    _, hh = img.size
    diff = int(np.ceil(hh / 336) * 336) - hh
    top = int(diff / 2)
    img = ImageOps.expand(img, border=(0, top, 0, diff - top), fill=(255, 255, 255))
    if trans:
        img = img.transpose(Image.TRANSPOSE)
    return np.array(img)


def interp336_weights(in_size):
    """get_weights_and_indices of phi.py:333-359 for out_size 336 (2 live taps; taps 2,3 are 0)."""
    def cubic(x):
        a = np.abs(x)
        return ((1.5 * a ** 3 - 2.5 * a ** 2 + 1) * (a <= 1)
                + (-0.5 * a ** 3 + 2.5 * a ** 2 - 4 * a + 2) * ((a > 1) & (a <= 2)))
    scale = 336 / in_size
    oc = np.linspace(0, in_size - 1, 336)
    ic = oc / scale
    left = np.floor(ic - 0.5).astype(np.int32)
    right = left + 1
    left, right = np.clip(left, 0, in_size - 1), np.clip(right, 0, in_size - 1)
    w = np.zeros((336, 4), dtype=np.float32)
    idx = np.zeros((336, 4), dtype=np.int32)
    for i in range(336):
        idx[i, 0], idx[i, 1] = left[i], right[i]
        w[i, 0] = cubic(ic[i] - left[i])
        w[i, 1] = cubic(right[i] - ic[i])
        s = w[i].sum()
        if s != 0:
            w[i] /= s
    return w, idx


def interpolate_336(x):
    """phi.py:360-372, vectorised: out = sum_{a,b} (wh[a]*ww[b] in fp32) * in[hi[a], wi[b]] in float64."""
    N, C, H, W = x.shape
    hw, hi = interp336_weights(H)
    ww, wi = interp336_weights(W)
    out = np.zeros((N, C, 336, 336), dtype=x.dtype)
    for a in range(4):
        for b in range(4):
            wp = (hw[:, a][:, None] * ww[:, b][None, :])                  # float32 product
            out += wp[None, None].astype(np.float64) * x[:, :, hi[:, a]][:, :, :, wi[:, b]]
    return out


def image_processor(images, num_crops=16, max_crops=None):
    """phi.py:289-329. max_crops defaults to the reference's 17 when num_crops == 16."""
    max_crops = max_crops or num_crops + 1
    hd = [((hd_transform_u8(im, num_crops) / 255.0 - MEAN) / STD).transpose(2, 0, 1) for im in images]
    shapes = [[im.shape[1], im.shape[2]] for im in hd]
    ntok = [int((h // 336 * w // 336 + 1) * 144 + 1 + (h // 336 + 1) * 12) for h, w in shapes]
    glb = [interpolate_336(im[None]) for im in hd]
    crops = [im.reshape(1, 3, h // 336, 336, w // 336, 336).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 3, 336, 336)
             for im, (h, w) in zip(hd, shapes)]
    crops = [np.concatenate([g, c], 0) for g, c in zip(glb, crops)]
    out = []
    for c in crops:
        if c.shape[0] < max_crops:
            c = np.concatenate([c, np.zeros((max_crops - c.shape[0], 3, 336, 336))], 0)
        out.append(c)
    return {'pixel_values': np.stack(out, 0), 'image_sizes': shapes, 'num_img_tokens': ntok}
