"""Public inference API with the reference's signatures:
load / generate / choose / constrain (/root/reference/phi_3_vision_mlx.py:1279, 1324, 1376, 1425)
and the decode drivers behind them (_generate pv:376-409, _choose_from pv:466-487,
_constrain/_get_beam pv:500-619, Streamer/LogitStopper/TokenStopper pv:45-117,
_apply_chat_template pv:341-357). Control flow lives here in Python; every tensor op is a
kernel of libphi3b200.so.

Differences a caller can see (all opt-in or forced by the offline environment):
  * `load()` accepts tokenizer=, weights= (dict or safetensors dir), cfg=, random_init=,
    num_crops=, device= through **kwargs — there are no checkpoint / tokenizer files offline;
  * quantize_model=True quantises the given (bf16) weights at load time — 4-bit, group 64, every Linear / Embedding, as
    `nn.quantize(model, 64, 4)` does for the reference's pre-quantised checkpoint (pv:264,291-305) — instead of reading a
    `quantized_model.safetensors`;
  * use_adapter=True takes the LoRA adapter from `adapter=` ({'config': adapter_config dict, 'weights': {name: tensor}}) or
    `adapter_path=` (adapter_config.json + adapters.safetensors, the files train_lora writes, pv:1005-1013) and folds it into
    the device weight copies: W' = bf16(W + scale*alpha/rank * (A B)^T) — inference with zero extra kernels instead of
    LoRALinear's two extra matmuls per call (phi:129-133). `set_adapter(model, adapter)` swaps or removes it IN PLACE (the
    reference reloads the model). With quantize_model=True the adapted matrices are base = the 4-bit image + the low-rank
    update, i.e. LoRALinear over QuantizedLinear (phi:94-95), and stream as bf16 at decode; all other matrices stay 4-bit;
  * constrain(..., n_beam=3) exposes the beam width the reference hard-codes (pv:505);
  * images may be PIL images or uint8 HWC arrays (no URL fetching offline);
  * generate_batch(prompts, images_per_prompt, ...) is an extension: the reference raises on images + a prompt list
    (pv:377-378), so its only way to run N image prompts is N sequential batch-1 calls. generate_batch runs them as ONE
    left-padded batch with per-prompt batch-1 semantics (each row keeps its own positions 0..L_i-1, pad keys masked,
    image tokens spliced per row), which is what BASELINE config 3 (64 image+text prompts) needs;
  * load(model_path=<dir>) / load(weights=<dir>) read that directory's config.json like _get_cfg (pv:258,359-363): LM
    hyper-parameters and rope_scaling come from the checkpoint, kwargs override; `sanitized` / `quantized` keys are honoured
    (pv:262-264,371-374). The frozen PHI35_* configs are only used for random_init and weight dicts.
"""
import os
import time
import glob
import torch
from . import _lib
from ._lib import call, ptr
from .configs import PHI35_MINI, PHI35_VISION, ID_EOS, with_overrides
from .model import Phi3B200, _stream, capture_graph
from .processor import Phi3FProcessor, Phi3VProcessor, ByteTokenizer
from .weights import random_weights

PATH_ORIGINAL_PHI3_VISION = 'models/phi3_v'             # pv:38-41
PATH_ORIGINAL_PHI3_BLIND = 'models/phi3_mini_128k'
PATH_ADAPTERS = 'adapters'                              # pv:37


class Tic:                                              # phi.py:16-24
    def __init__(self):
        self.last = time.perf_counter()

    def __call__(self):
        now = time.perf_counter()
        d, self.last = now - self.last, now
        return d


# ------------------------------------------------------------------------------------- load
def _read_config(path):
    """_get_cfg (pv:359-369): config.json -> SimpleNamespace. None when the directory has no config.json."""
    import json
    from types import SimpleNamespace
    f = os.path.join(path, 'config.json')
    if not os.path.exists(f):
        return None
    try:
        d = json.load(open(f))
    except json.JSONDecodeError:
        raise ValueError(f'Invalid JSON in configuration file: {f}')
    d.setdefault('use_quantized_cache', False)
    if 'num_key_value_heads' not in d:
        d['num_key_value_heads'] = d['num_attention_heads']
    return SimpleNamespace(**d)


def _load_tokenizer(path):
    """The reference takes AutoTokenizer.from_pretrained(model_path) (phi:230). Offline, a local tokenizer.json is enough:
    load it through `tokenizers` directly when transformers cannot build the full tokenizer class."""
    try:
        from transformers import AutoTokenizer
        return AutoTokenizer.from_pretrained(path, local_files_only=True)
    except Exception:
        from transformers import PreTrainedTokenizerFast
        return PreTrainedTokenizerFast(tokenizer_file=os.path.join(path, 'tokenizer.json'))


def dequantize_mlx(wq, scales, biases, group_size=64, bits=4):
    """mx.dequantize layout (what `quantized_model.safetensors` holds after pv:291-305): `wq` uint32 [N, K*bits/32], element i
    of a row in bits [bits*(i%(32/bits)), +bits) of word i/(32/bits); scales / biases [N, K/group_size]: w = q*scale + bias."""
    per = 32 // bits
    wq = wq.to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(per, dtype=torch.int64, device=wq.device) * bits
    q = ((wq[..., None] >> sh) & ((1 << bits) - 1)).reshape(wq.shape[0], -1).to(torch.float32)      # [N, K]
    N, K = q.shape
    w = q.reshape(N, K // group_size, group_size) * scales.to(torch.float32)[..., None] + biases.to(torch.float32)[..., None]
    return w.reshape(N, K).to(torch.bfloat16)


def _read_safetensors(path, sanitized=False, quantized=None):
    """_get_wt (pv:371-374): every *.safetensors shard of the directory. Unsanitized (HF) checkpoints carry the patch conv as
    [O,I,kh,kw] and are transposed to the reference layout [O,kh,kw,I]; `sanitized` ones already are (pv:276-289). A
    `quantized` checkpoint (pv:291-305) stores weight/scales/biases triples; they are expanded to the bf16 image here and
    re-quantised by Phi3B200(quantize_model=True) with the same group size (the codes are reproduced up to the bf16 rounding
    of the image)."""
    from safetensors.torch import load_file
    w = {}
    for f in sorted(glob.glob(os.path.join(path, '*.safetensors'))):
        w.update(load_file(f))
    if not w:
        raise FileNotFoundError(f'no *.safetensors under {path}')
    key = 'model.vision_embed_tokens.img_processor.vision_model.embeddings.patch_embedding.weight'
    if key in w and not sanitized and w[key].dim() == 4 and w[key].shape[1] == 3 and w[key].shape[-1] != 3:
        w[key] = w[key].permute(0, 2, 3, 1).contiguous()   # HF [O,I,kh,kw] -> reference layout [O,kh,kw,I] (pv:374)
    if quantized:
        gs, bits = int(quantized.get('group_size', 64)), int(quantized.get('bits', 4))
        for k in [k for k in w if k.endswith('.scales')]:
            base = k[:-len('.scales')]
            w[base + '.weight'] = dequantize_mlx(w[base + '.weight'], w.pop(k), w.pop(base + '.biases'), gs, bits)
    return w


def sanitize(from_path, to_path):
    """_sanitize (pv:276-289): re-save a checkpoint directory in the reference's own parameter layout (patch conv already
    [O,kh,kw,I]) with `sanitized: true` in config.json, *.json side files copied."""
    import json
    import shutil
    from safetensors.torch import save_file
    cfg = _read_config(from_path)
    if cfg is None:
        raise FileNotFoundError(f'Configuration file not found: {from_path}/config.json')
    w = _read_safetensors(from_path, sanitized=bool(getattr(cfg, 'sanitized', False)), quantized=getattr(cfg, 'quantized', None))
    os.makedirs(to_path, exist_ok=True)
    for f in glob.glob(os.path.join(from_path, '*.json')):
        shutil.copy(f, to_path)
    d = {k: v for k, v in vars(cfg).items() if k != 'quantized'}
    d['sanitized'] = True
    json.dump(d, open(os.path.join(to_path, 'config.json'), 'w'), indent=4)
    save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(to_path, 'sanitized_model.safetensors'))


def load(blind_model=False, quantize_model=False, quantize_cache=False, use_adapter=False, **kwargs):
    """pv:1279-1322. Returns (model, processor)."""
    adapter = kwargs.pop('adapter', None)
    adapter_path = kwargs.pop('adapter_path', None)
    device = kwargs.pop('device', 'cuda')
    cfg = kwargs.pop('cfg', None)
    tokenizer = kwargs.pop('tokenizer', None)
    weights = kwargs.pop('weights', None)
    clip_cfg = kwargs.pop('clip_cfg', None)
    num_crops = kwargs.pop('num_crops', 16)
    default_path = PATH_ORIGINAL_PHI3_BLIND if blind_model else PATH_ORIGINAL_PHI3_VISION
    if quantize_model and os.path.isdir(default_path + '_Q') and 'model_path' not in kwargs and weights is None:
        default_path += '_Q'                                               # pv:1306-1315: the pre-quantised checkpoint directory
    ckpt_dir = None
    if weights is None:
        path = kwargs.pop('model_path', default_path)
        if kwargs.pop('random_init', False):
            cfg = cfg or (PHI35_MINI if blind_model else PHI35_VISION)
            weights = random_weights(cfg, seed=kwargs.pop('seed', 0), device=device, clip_cfg=clip_cfg)
        elif os.path.isdir(path):
            ckpt_dir = path
        else:
            raise FileNotFoundError(f'{path} not found and no network to fetch it (reference: _setup, pv:247-255); '
                                    'pass weights=..., or random_init=True')
    elif isinstance(weights, str):
        ckpt_dir = weights
    if ckpt_dir is not None:
        file_cfg = _read_config(ckpt_dir)
        if cfg is None and file_cfg is not None:
            cfg = file_cfg
        weights = _read_safetensors(ckpt_dir, sanitized=bool(getattr(cfg, 'sanitized', False)),
                                    quantized=getattr(cfg, 'quantized', None))
        if getattr(cfg, 'quantized', None):
            quantize_model = True                                          # pv:264: nn.quantize(model, group_size, bits) before load_weights
    cfg = cfg or (PHI35_MINI if blind_model else PHI35_VISION)
    cfg = with_overrides(cfg, use_quantized_cache=quantize_cache)          # pv:1322 -> phi.py:512,572
    if tokenizer is None:
        path = ckpt_dir or default_path
        if os.path.isdir(path) and (os.path.exists(os.path.join(path, 'tokenizer.json'))
                                    or os.path.exists(os.path.join(path, 'tokenizer.model'))):
            tokenizer = _load_tokenizer(path)
        else:
            tokenizer = ByteTokenizer()
    for k, v in kwargs.items():                                            # remaining kwargs override cfg (pv:359-363)
        setattr(cfg, k, v)
    lora = None
    if use_adapter:
        if adapter is None:
            adapter_path = adapter_path or f'{PATH_ADAPTERS}/{os.path.basename(PATH_ORIGINAL_PHI3_BLIND if blind_model else PATH_ORIGINAL_PHI3_VISION)}'
            adapter = _read_adapter(adapter_path)                          # pv:266-271, _get_adapter_path pv:462
        lora = lora_modules(adapter['config'], adapter['weights'], cfg.num_hidden_layers)
    model = Phi3B200(cfg, weights, device=device, clip_cfg=clip_cfg, quantize_model=quantize_model, lora=lora)
    if ckpt_dir is not None and getattr(cfg, 'architectures', None):       # pv:260: processor class follows the checkpoint's arch
        blind_model = not cfg.architectures[0].startswith('Phi3V')
    processor = Phi3FProcessor(tokenizer) if blind_model else Phi3VProcessor(tokenizer, num_crops=num_crops, device=device)
    return model, processor


def _read_adapter(path):
    import json
    if not os.path.isdir(path):
        raise FileNotFoundError(f'LoRA adapter directory {path} not found (pass adapter= or adapter_path=)')
    from safetensors.torch import load_file
    return {'config': json.load(open(os.path.join(path, 'adapter_config.json'))),
            'weights': load_file(os.path.join(path, 'adapters.safetensors'))}


def lora_modules(lora_cfg, lora_weights, n_layers):
    """_linear_to_lora_layers (pv:234-245): which Linear modules carry a LoRALinear and with what scale (phi:121: scale * alpha
    / rank). Returns {module name: (lora_a [in, r], lora_b [r, out], scale)} — the `lora` argument of Phi3B200 / set_adapter."""
    layers = lora_cfg['lora_layers']
    if isinstance(layers, int):
        layers = list(range(n_layers))[-layers:]
    elif not isinstance(layers, list):
        raise ValueError('Invalid type for lora_layers. Expected int (number of layers) or list (layer indices or names).')
    lp = lora_cfg['lora_parameters']
    sc = float(lp['scale']) * (float(lp['alpha']) / float(lp['rank']))
    out = {}
    for i in layers:
        for t in lora_cfg['lora_targets']:
            key = f'model.layers.{i}.{t}'
            a, b = lora_weights.get(key + '.lora_a'), lora_weights.get(key + '.lora_b')
            if a is None or b is None:
                raise KeyError(f'adapter has no {key}.lora_a / .lora_b')
            out[key] = (a, b, sc)
    return out


def set_adapter(model, adapter=None, adapter_path=None):
    """Swap (or remove, adapter=None) the LoRA adapter of a loaded model without reloading it: the reference has to call load()
    again (pv:266-271). `adapter` = {'config': adapter_config dict, 'weights': {name: tensor}} or a directory via adapter_path."""
    if adapter is None and adapter_path is not None:
        adapter = _read_adapter(adapter_path)
    lora = None if adapter is None else lora_modules(adapter['config'], adapter['weights'], model.cfg.num_hidden_layers)
    model.set_adapter(lora)
    return model


def merge_lora(weights, lora_cfg, lora_weights, n_layers):
    """_linear_to_lora_layers (pv:234-245) + LoRALinear.__call__ (phi:129-133), folded: for every targeted Linear of the
    selected decoder layers W <- bf16(W + scale * alpha / rank * (lora_a @ lora_b)^T); lora_a [in, r], lora_b [r, out]."""
    layers = lora_cfg['lora_layers']
    if isinstance(layers, int):
        layers = list(range(n_layers))[-layers:]
    elif not isinstance(layers, list):
        raise ValueError('Invalid type for lora_layers. Expected int (number of layers) or list (layer indices or names).')
    lp = lora_cfg['lora_parameters']
    scale = float(lp['scale']) * (float(lp['alpha']) / float(lp['rank']))
    out = dict(weights)
    for i in layers:
        for t in lora_cfg['lora_targets']:
            key = f'model.layers.{i}.{t}'
            a, b = lora_weights.get(key + '.lora_a'), lora_weights.get(key + '.lora_b')
            if a is None or b is None:
                raise KeyError(f'adapter has no {key}.lora_a / .lora_b')
            w = weights[key + '.weight']
            delta = (a.to(w.device, torch.float32) @ b.to(w.device, torch.float32)).T      # [out, in]
            out[key + '.weight'] = (w.to(torch.float32) + scale * delta).to(torch.bfloat16)
    return out


# ------------------------------------------------------------------------------------- helpers
def _apply_chat_template(prompt, images, verbose, apply_chat_template=True):
    """pv:341-357 (image loading from URL/file is host glue; PIL images / arrays pass through)."""
    if apply_chat_template is False:
        if verbose:
            print(f'*** Prompt ***\n{prompt}\n*** Images ***\n{images}\n*** Output ***')
        return prompt, images
    if images is not None:
        images = list(images) if isinstance(images, (list, tuple)) else [images]
        images = [_load_image(i) for i in images]
        img_prompt = '\n'.join([f'<|image_{i+1}|>' for i in range(len(images))]) + '\n'
    else:
        img_prompt = ''
    prompt = [prompt] if isinstance(prompt, str) else prompt
    prompt = [f"<|user|>\n{img_prompt}{i.strip()}<|end|>\n<|assistant|>\n" for i in prompt]
    if verbose:
        prompt_str = "\n".join(map(str.strip, prompt)).strip()
        images_str = "\n".join(f'<image {getattr(i, "size", getattr(i, "shape", ""))}>' for i in images) if images else "None"
        print(f'*** Prompt ***\n{prompt_str}\n*** Images ***\n{images_str}\n*** Output ***')
    prompt = prompt[0] if len(prompt) == 1 else prompt
    return prompt, images


def _load_image(x):
    if isinstance(x, str):
        from PIL import Image
        if x.startswith('http://') or x.startswith('https://'):
            raise ValueError('no network in this environment: pass a local path, PIL image or uint8 array')
        return Image.open(x)
    return x


def _preprocess(s):                                                        # pv:489-493
    for i in ['<|system|>', '<|user|>', '<|end|>']:
        s = s.replace(f'{i} ', f'{i}\n').replace(f'{i}\n\n', f'{i}\n')
    return s.replace('<|end|><|assistant|>', '<|end|>\n<|assistant|>')


def _row_stats(model, logits2d, n_top=0, gather=None):
    """One p3_row_stats launch over fp32 logits [R,V]. Returns dict of device tensors."""
    R, V = logits2d.shape
    dev = logits2d.device
    out = dict(argmax=torch.empty(R, dtype=torch.int32, device=dev), max=torch.empty(R, dtype=torch.float32, device=dev),
               lse=torch.empty(R, dtype=torch.float32, device=dev))
    if n_top:
        out['top_ids'] = torch.empty((R, n_top), dtype=torch.int32, device=dev)
        out['top_lp'] = torch.empty((R, n_top), dtype=torch.float32, device=dev)
    ng = 0
    if gather is not None:
        gather = gather.to(dev, torch.int32).contiguous()
        ng = gather.shape[1]
        out['gather_lp'] = torch.empty((R, ng), dtype=torch.float32, device=dev)
    call('p3_row_stats', ptr(logits2d), R, logits2d.stride(0), V, ptr(out['argmax']), ptr(out['max']), ptr(out['lse']),
         n_top, ptr(out.get('top_ids')), ptr(out.get('top_lp')), ng, ptr(gather), ptr(out.get('gather_lp')), _stream())
    return out


# ------------------------------------------------------------------------------------- generate
class Streamer:
    """pv:45-77, fed from the device-side token history."""

    def __init__(self, processor, stream, mute):
        self.tokenizer = processor.tokenizer
        self.mute = mute
        self.stream = stream and (not mute)
        self.list_tokens = []
        self.idx_sofar = 0

    def stream_token(self, tok):
        self.list_tokens.append(tok)
        txt = self.tokenizer.decode(self.list_tokens)
        idx_split = txt.rfind(' ', self.idx_sofar)
        if idx_split > 0:
            print(txt[self.idx_sofar:idx_split], end='', flush=True)
            self.idx_sofar = idx_split

    def end(self, hist):
        rows = hist.tolist()
        if self.stream:
            txt = self.tokenizer.decode(self.list_tokens)
            print(txt[self.idx_sofar:], '\n', flush=True)
            return txt, len(self.list_tokens)
        list_txt = self.tokenizer.batch_decode([(r[:r.index(ID_EOS) + 1] if ID_EOS in r else r) for r in rows])
        if not self.mute:
            for i, gen in enumerate(list_txt):
                print(f'\n< Generated text for prompt #{i} >\n{gen}')
        return list_txt, hist.numel()


class LogitStopper:
    """pv:79-104 (B=1 early-stop heuristic on the EOS log-prob)."""

    def __init__(self, max_tokens, early_stop):
        self.step = 0
        # pv:82: `early_stop if isinstance(early_stop, int) and (early_stop < max_tokens) else False` — bool is an int there:
        # True enables the heuristic with threshold 1, False (== 0) leaves it off
        self.early_stop = early_stop if isinstance(early_stop, int) and (early_stop < max_tokens) else False
        self.log_prob_sum = 0.0
        self.best_eos_sofar = -float('inf')
        self.log_prob_sum_at_best_eos = 0.0

    def __call__(self, log_prob_best, log_prob_eos):
        if not self.early_stop:
            return False
        if log_prob_eos > self.best_eos_sofar:
            since = self.log_prob_sum - self.log_prob_sum_at_best_eos
            if since < self.best_eos_sofar and self.step > self.early_stop:
                return True
            self.best_eos_sofar = log_prob_eos
            self.log_prob_sum_at_best_eos = self.log_prob_sum
        self.log_prob_sum += log_prob_best
        self.step += 1
        return False


def _batch_inputs(processor, prompts, images):
    """N image+text (or text-only) prompts -> ONE left-padded model input with per-prompt batch-1 semantics:
    every prompt goes through the processor alone, exactly as the reference's batch-1 VLM path does (pv:381, phi:263-281), then
    rows are left-padded like Phi3FProcessor._tokenize (phi:236-245: ids 0, pids 1, mask 0) and the image-token positions are
    shifted by each row's pad. Row b keeps positions 0..L_b-1, so its logits equal the batch-1 call's."""
    per = []
    for i, p in enumerate(prompts):
        im = None if images is None else images[i]
        if im is not None and not isinstance(im, (list, tuple)):
            im = [im]
        per.append(processor(p, im) if im else processor(p))
    B, L = len(per), max(d['input_ids'].shape[1] for d in per)
    ids = torch.zeros((B, L), dtype=torch.int64)
    pids = torch.ones((B, L), dtype=torch.int64)
    mask = torch.zeros((B, L), dtype=torch.int64)
    pvs, sizes, pos = [], [], []
    for b, d in enumerate(per):
        row = d['input_ids'][0]
        l = row.shape[0]
        ids[b, L - l:], pids[b, L - l:], mask[b, L - l:] = row, torch.arange(l), 1
        if 'pixel_values' in d:
            pvs.append(d['pixel_values'])
            sizes.append(torch.as_tensor(d['image_sizes']))
            pp = torch.as_tensor(d['positions']).clone()
            pp[:, 0], pp[:, 1] = b, pp[:, 1] + (L - l)
            pos.append(pp)
    out = {'input_ids': ids, 'pids': pids, 'mask': mask}
    if pvs:
        n_crops = max(p.shape[1] for p in pvs)                               # crop axis is zero-padded (phi:311-316)
        pvs = [p if p.shape[1] == n_crops else torch.cat([p, p.new_zeros((p.shape[0], n_crops - p.shape[1]) + p.shape[2:])], 1)
               for p in pvs]
        out.update(pixel_values=torch.cat(pvs, 0), image_sizes=torch.cat(sizes, 0), positions=torch.cat(pos, 0))
    return out


def _generate(model, processor, prompt, images=None, max_tokens=512, verbose=True, return_tps=False, early_stop=False,
              stream=True, mute=False, return_tokens=False, eos_check_every=16, top_p=None, temperature=1.0, seed=0,
              dict_input=None):
    """pv:376-409. The loop body is one CUDA-graph replay per token; EOS for all rows
    (TokenStopper, pv:106-117) is polled every `eos_check_every` steps instead of twice per token —
    rows are truncated at their first EOS afterwards exactly as the reference does (pv:73).
    `dict_input`: a prebuilt model input (generate_batch) instead of processor(prompt, images)."""
    if dict_input is None and images is not None and isinstance(prompt, list):
        raise ValueError('Images cannot be provided when prompt is a list')
    streamer = Streamer(processor, stream, mute)
    if dict_input is None:
        dict_input = processor(prompt, images)
    B = dict_input['input_ids'].shape[0]
    if B > 1:
        streamer.stream = False                                            # pv:53-56: batches are never streamed
    logit_stopper = LogitStopper(max_tokens, early_stop if B == 1 else False)
    per_token_sync = (streamer.stream and B == 1) or bool(logit_stopper.early_stop)
    tic = Tic()
    logits, cache = model(**dict_input, max_tokens=max_tokens, logits_rows='last')
    sampler = None
    if top_p is not None:
        # extension (the reference is greedy only): nucleus sampling with a seeded uniform table [steps, B]
        g = torch.Generator(device=model.dev).manual_seed(int(seed))
        u = torch.rand((max_tokens + 1, B), generator=g, device=model.dev)
        token = torch.empty(B, dtype=torch.int32, device=model.dev)
        call('p3_top_p_sample', ptr(logits[:, -1, :]), B, logits.stride(0), model.V, float(top_p), float(temperature),
             ptr(u), ptr(token), None, None, 0, _stream())
        sampler = (top_p, temperature, u)
    else:
        token = _row_stats(model, logits[:, -1, :])['argmax']
    if per_token_sync:
        if streamer.stream:
            streamer.stream_token(int(token[0].item()))
    torch.cuda.synchronize()
    prompt_time = tic()
    ses = model.decode_session(token, cache, max_tokens - 1, use_graph=not bool(logit_stopper.early_stop), sampler=sampler)
    for i in range(max_tokens - 1):
        if logit_stopper.early_stop:
            # un-graphed step so the EOS log-prob of this step can be read (pv:395)
            lg = model._forward_tokens(ses.tok, B, 1, cache, 1, True, cache.offset + i, 'last', past_dev=ses.past_dev,
                                       n_splits=ses.n_splits)
            s2 = _row_stats(model, lg[:, -1, :], gather=torch.full((B, 1), ID_EOS))
            ses.tok.copy_(s2['argmax'])
            call('p3_decode_advance', ptr(ses.tok), ptr(ses.hist), ses.hist.stride(0), B, ptr(ses.step_dev),
                 ptr(ses.past_dev), ptr(ses.eos), _stream())
            ses.steps_run += 1
            if streamer.stream:
                streamer.stream_token(int(ses.tok[0].item()))
            if logit_stopper(float((s2['max'] - s2['lse'])[0].item()), float(s2['gather_lp'][0, 0].item())):
                break
            if ses.all_eos():
                break
            continue
        ses.step()
        if per_token_sync:
            t = int(ses.tok[0].item())
            streamer.stream_token(t)
            if t == ID_EOS:
                break
        elif eos_check_every and (i + 1) % eos_check_every == 0 and ses.all_eos():
            break
    hist = ses.finish()
    torch.cuda.synchronize()
    cache.release()                                                        # slab + captured graph go back to the model
    result, gen_len = streamer.end(hist)
    gen_time = tic()
    prompt_len = dict_input['input_ids'].numel()
    prompt_tps = prompt_len / prompt_time
    gen_tps = (gen_len - 1) / gen_time if gen_time > 0 else float('inf')
    if verbose:
        print(f"\nPrompt: {prompt_tps:.2f} tokens-per-sec ({prompt_len} tokens / {prompt_time:.1f} sec)")
        print(f"Generate: {gen_tps:.2f} tokens-per-sec ({gen_len} tokens / {gen_time:.1f} sec)")
    if return_tokens:
        return hist
    if return_tps:
        return prompt_tps, gen_tps
    return result


def generate(prompt, images=None, preload=None, blind_model=False, quantize_model=False, quantize_cache=False,
             use_adapter=False, max_tokens=512, verbose=True, return_tps=False, early_stop=False, stream=True,
             apply_chat_template=True, enable_api=False):
    """pv:1324-1374."""
    if enable_api:
        raise NotImplementedError('enable_api (remote tool calls) is outside the hot-path scope')
    if preload is None:
        preload = load(blind_model=blind_model, quantize_model=quantize_model, quantize_cache=quantize_cache,
                       use_adapter=use_adapter)
    prompt, images = _apply_chat_template(prompt, images, verbose, apply_chat_template)
    return _generate(*preload, prompt, images, max_tokens, verbose, return_tps, early_stop, stream)


def generate_batch(prompts, images=None, preload=None, blind_model=False, quantize_model=False, quantize_cache=False,
                   use_adapter=False, max_tokens=512, verbose=False, return_tps=False, apply_chat_template=True,
                   return_tokens=False):
    """Extension of generate() (pv:1324) for BASELINE config 3: a LIST of prompts, each with its own image(s).
    `images`: None, or a list with one entry per prompt (None | image | list of images). Returns a list of strings
    (or (prompt_tps, gen_tps) / the token history). Per-prompt results equal separate batch-1 generate() calls."""
    if isinstance(prompts, str):
        raise ValueError('generate_batch takes a list of prompts; use generate() for a single one')
    if images is not None and len(images) != len(prompts):
        raise ValueError('images must hold one entry (None, an image or a list of images) per prompt')
    if preload is None:
        preload = load(blind_model=blind_model, quantize_model=quantize_model, quantize_cache=quantize_cache,
                       use_adapter=use_adapter)
    model, processor = preload
    texts, imgs = [], []
    for i, p in enumerate(prompts):
        im = None if images is None else images[i]
        t, im = _apply_chat_template(p, im, False, apply_chat_template)
        texts.append(t)
        imgs.append(im)
    # static LongRoPE switch (phi:492, H7) is a per-prompt decision in the batch-1 reference: rows must agree on it
    lens = _prompt_lengths(processor, texts, imgs)
    orig = model.cfg.original_max_position_embeddings
    if len({(l + max_tokens) > orig for l in lens}) > 1:
        raise ValueError('prompts on both sides of the LongRoPE switch (prompt + max_tokens vs '
                         f'{orig}) cannot share a batch: split them into two calls')
    prev = model.force_long_rope
    if prev is None:
        model.force_long_rope = (max(lens) + max_tokens) > orig
    try:
        dict_input = _batch_inputs(processor, texts, imgs if any(i is not None for i in imgs) else None)
        return _generate(model, processor, texts, None, max_tokens, verbose, return_tps, False, False, mute=not verbose,
                         return_tokens=return_tokens, dict_input=dict_input)
    finally:
        model.force_long_rope = prev


def _prompt_lengths(processor, texts, imgs):
    """token count of every prompt (text chunks + image tokens, phi:263-277) without running the image transform"""
    from .processor import hd_geometry
    import re
    out = []
    for t, im in zip(texts, imgs):
        if not im:
            out.append(len(processor.tokenizer(t).input_ids))
            continue
        n = sum(len(c) for c in processor.tokenizer(re.split(r"<\|image_\d+\|>", t)).input_ids)
        for x in im:
            if hasattr(x, 'size') and not hasattr(x, 'shape'):
                w, h = x.size
            else:
                h, w = x.shape[:2]
            n += hd_geometry(w, h, processor.img_processor.num_crops)['num_img_tokens']
        out.append(n)
    return out


# ------------------------------------------------------------------------------------- choose
def _choose_from(model, processor, prompt, choices='ABCDE', mute=False):
    """pv:466-487: last-position log-softmax restricted to the option tokens, argmax."""
    was_str = isinstance(prompt, str)
    options = processor([f' {i}' for i in choices])['input_ids'][:, -1]
    dict_input = processor(prompt)
    logits, _ = model(**dict_input, max_tokens=0, logits_rows='last')
    B = logits.shape[0]
    st = _row_stats(model, logits[:, -1, :], gather=options[None].expand(B, -1))
    indices = torch.argmax(st['gather_lp'], dim=-1).tolist()
    output = [choices[i] for i in indices]
    if not mute:
        if was_str:
            print(output[0])
        else:
            for i, o in enumerate(output):
                print(f'\n< Chosen option for prompt #{i} >\n{o}')
    return output[0] if was_str else output


def choose(prompt, choices='ABCDE', images=None, preload=None, blind_model=False, quantize_model=False,
           quantize_cache=False, use_adapter=False, verbose=True, apply_chat_template=True):
    """pv:1376-1423."""
    if preload is None:
        preload = load(blind_model=blind_model, quantize_model=quantize_model, quantize_cache=quantize_cache,
                       use_adapter=use_adapter)
    prompt, images = _apply_chat_template(prompt, images, verbose, apply_chat_template)
    return _choose_from(*preload, prompt, choices)


# ------------------------------------------------------------------------------------- constrain
class _ConstrainStep:
    """One constrained-decoding step (pv:567-597) as CUDA graphs: the shapes of a step are fixed for a constraint —
    [B, 1+C] tokens for the committed forward (advance_offset=1), [B*n_beam, 1+C] for the read-only shared-prefix beam forward
    (n_beam, advance_offset=0) — and the cache offset travels in a device int32 (`past_dev`), so the ~330 kernel launches of a
    step replay without any Python between them. Eager first call (warm-up) is rolled back: it only writes the KV of the
    position it is about to write again."""

    def __init__(self, model, cache, B, C, n_beam, idc, use_beam, max_new):
        self.m, self.cache, self.B, self.C, self.nb, self.use_beam = model, cache, B, C, n_beam, use_beam
        dev = model.dev
        self.tp = torch.zeros((B, 1 + C), dtype=torch.int64, device=dev)
        self.tp[:, 1:] = idc
        self.g = torch.full((B, 1 + C), -1, dtype=torch.int32, device=dev)
        self.g[:, :C] = idc.to(torch.int32)
        self.past_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        tiles = (cache.offset + max_new + 1 + C + 63) // 64
        self.ns = model._splits(cache, B, tiles)
        self.ns_beam = model._splits(cache, B * n_beam, tiles)
        if use_beam:
            self.seq = torch.zeros((B * n_beam, 1 + C), dtype=torch.int64, device=dev)
            self.seq[:, 1:] = idc
            self.gb = torch.full((B * n_beam, 1 + C), -1, dtype=torch.int32, device=dev)
            self.gb[:, :C] = idc.to(torch.int32)
        self.graphs = None

    def _main(self):
        lg, _ = self.m(self.tp, cache=self.cache, advance_offset=1, past_dev=self.past_dev, n_splits=self.ns)
        self.cache.offset -= 1                                                # the caller advances the offset per replay
        return lg, _row_stats(self.m, lg.reshape(-1, self.m.V), gather=self.g.reshape(-1, 1))

    def _beam(self, lg):
        st = _row_stats(self.m, lg[:, 0, :], n_top=self.nb)
        self.seq[:, 0] = st['top_ids'].reshape(-1)
        lb, _ = self.m(self.seq, cache=self.cache, n_beam=self.nb, advance_offset=0, past_dev=self.past_dev, n_splits=self.ns_beam)
        return st, _row_stats(self.m, lb.reshape(-1, self.m.V), gather=self.gb.reshape(-1, 1))

    def capture(self):
        off = self.cache.offset
        self.past_dev.fill_(off)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                                       # warm-up outside capture (smem attributes, allocator)
            lg, _ = self._main()
            if self.use_beam:
                self.past_dev.fill_(off + 1)
                self._beam(lg)
                self.past_dev.fill_(off)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with capture_graph(g1):
            self.lg, self.stp = self._main()
        g2 = None
        if self.use_beam:
            g2 = torch.cuda.CUDAGraph()
            with capture_graph(g2):
                self.st_top, self.s2 = self._beam(self.lg)
        self.graphs = (g1, g2)
        assert self.cache.offset == off

    def run_main(self, token):
        """forward of cat[token, constraint] at the current offset; commits `token` (offset + 1)"""
        self.tp[:, 0] = token
        self.past_dev.fill_(self.cache.offset)
        self.graphs[0].replay()
        self.cache.offset += 1
        return self.lg, self.stp

    def run_beam(self):
        """top-n_beam candidates of the first position + their read-only forwards against the shared prefix"""
        self.past_dev.fill_(self.cache.offset)
        self.graphs[1].replay()
        return self.st_top, self.s2


def _constrain(model, processor, prompt, constraints, return_full_text=False, mute=False, use_beam=False, verbose=True,
               log_norm=False, n_beam=3, return_ids=False, alive_check_every=8, sync_timing=False, use_graph=True):
    """pv:500-619. Scores are means of log-probs over [tokens so far + constraint]; only the
    gathered log-probs are ever materialised (p3_row_stats) instead of full-vocab log-softmax."""
    import math

    def mean_lp(x):                                                        # _log_mean pv:501-504
        return x.sum(-1) / (math.log(x.shape[-1]) if log_norm else x.shape[-1])

    was_str = isinstance(prompt, str)
    prompt = [prompt] if was_str else list(prompt)
    tic = Tic()
    prompt_time = constrain_time = 0.0
    prompt = [_preprocess(s) for s in prompt]
    len_ps = [len(p) for p in prompt]
    B = len(prompt)
    V = model.V
    ids_trace = []

    dev = model.dev

    def beam_step(row_logits, cache, idc):
        """_get_beam pv:505-517 on logits rows [B,V] of the position that predicts the beam token (all on the device)."""
        C = idc.shape[0]
        st = _row_stats(model, row_logits, n_top=n_beam)
        cand, cand_lp = st['top_ids'], st['top_lp']                          # [B,nb]
        seq = torch.cat([cand.reshape(-1, 1).long(), idc.repeat(B * n_beam, 1)], 1)              # [B*nb,1+C]
        lg, _ = model(seq, cache=cache, n_beam=n_beam, advance_offset=0)
        g = torch.full((B * n_beam, 1 + C), -1, dtype=torch.int32, device=dev)
        g[:, :C] = seq[:, 1:].to(torch.int32)
        s2 = _row_stats(model, lg.reshape(-1, V), gather=g.reshape(-1, 1))
        rest = s2['gather_lp'].reshape(B * n_beam, 1 + C)[:, :C]
        bs = torch.cat([cand_lp.reshape(-1, 1), rest], 1)                    # [B*nb, 1+C]
        k = torch.argmax(bs.mean(1).reshape(B, n_beam), dim=-1)
        ar = torch.arange(B, device=dev)
        return st['argmax'].long(), cand.long()[ar, k], bs.reshape(B, n_beam, -1)[ar, k]

    for constraint in constraints:
        if isinstance(constraint, str):                                      # pv:531-536
            out = _choose_from(model, processor, prompt, constraint, True)
            prompt = [' '.join([p, o]) for p, o in zip(prompt, out)]
            output = prompt
            continue
        max_new, text = constraint
        ids_c = list(processor.tokenizer.encode(text, add_special_tokens=False)[1:])   # pv:538
        C = len(ids_c)
        idc = torch.tensor(ids_c, dtype=torch.int64, device=dev)
        dict_input = processor(prompt)
        S = dict_input['input_ids'].shape[1]
        logits, cache = model(**dict_input, max_tokens=max_new + C + 10, logits_rows='last')
        last = logits[:, -1, :]
        st = _row_stats(model, last, gather=torch.full((B, 1), ids_c[0], dtype=torch.int32, device=dev))
        s0 = st['gather_lp']                                                 # [B,1]
        running = (st['max'] - st['lse'])[:, None]                           # pv:548
        tiled = idc.repeat(B, 1)
        lr, _ = model(tiled, cache=cache, advance_offset=0)                  # peek, pv:545
        g = torch.full((B, C), -1, dtype=torch.int32, device=dev)
        g[:, :C - 1] = tiled[:, 1:].to(torch.int32)
        s1 = _row_stats(model, lr.reshape(-1, V), gather=g.reshape(-1, 1))['gather_lp'].reshape(B, C)[:, :C - 1]
        eos_col = torch.full((B, 1), ID_EOS, dtype=torch.int64, device=dev)
        pre_score = mean_lp(torch.cat([s0, s1], 1))
        pre_synth = torch.cat([tiled, eos_col], 1)
        if use_beam and max_new > 0:                                         # pv:551-557
            token, beam_tok, beam_sc = beam_step(last, cache, idc)
            post_score = mean_lp(beam_sc)
            post_synth = torch.cat([beam_tok[:, None], tiled], 1)
            win = pre_score > post_score
            score_sofar = torch.where(win, pre_score, post_score)
            synth_sofar = torch.where(win[:, None], pre_synth, post_synth)
        else:
            token = st['argmax'].long()
            score_sofar, synth_sofar = pre_score, pre_synth
        tokens = []
        alive = torch.ones(B, device=dev)
        step = None
        if use_graph and max_new > 2 and 1 + C <= 16:
            step = _ConstrainStep(model, cache, B, C, n_beam, idc, use_beam, max_new)
            step.capture()
        if sync_timing:
            torch.cuda.synchronize()
        prompt_time += tic()
        # pv:567-597. Every tensor of the loop lives on the device and nothing is read back per step (the reference syncs at
        # mx.eval pv:572,597 and in `_already`); the "no row alive" exit (pv:594-595) is polled every `alive_check_every` steps —
        # extra steps after it only append EOS columns behind the cut point of pv:601.
        for i in range(max_new):
            tokens.append(token[:, None])
            if step is not None:
                lg, stp = step.run_main(token)
            else:
                tp = torch.cat([token[:, None], tiled], 1)                   # [B,1+C]
                lg, cache = model(tp, cache=cache, advance_offset=1)
                g = torch.full((B, 1 + C), -1, dtype=torch.int32, device=dev)
                g[:, :C] = tp[:, 1:].to(torch.int32)
                stp = _row_stats(model, lg.reshape(-1, V), gather=g.reshape(-1, 1))
            glp = stp['gather_lp'].reshape(B, 1 + C)[:, :C]
            pre_score = mean_lp(torch.cat([running, glp], 1))
            pre_synth = torch.cat(tokens + [tiled, eos_col], 1)
            first_max_lp = (stp['max'] - stp['lse']).reshape(B, 1 + C)[:, 0]
            if use_beam:
                if step is not None:
                    sb, s2 = step.run_beam()
                    rest = s2['gather_lp'].reshape(B * n_beam, 1 + C)[:, :C]
                    bs = torch.cat([sb['top_lp'].reshape(-1, 1), rest], 1)
                    kk = torch.argmax(bs.mean(1).reshape(B, n_beam), dim=-1)
                    ar = torch.arange(B, device=dev)
                    token, beam_tok, beam_sc = sb['argmax'].long(), sb['top_ids'].long()[ar, kk], bs.reshape(B, n_beam, -1)[ar, kk]
                else:
                    token, beam_tok, beam_sc = beam_step(lg[:, 0, :], cache, idc)
                post_score = mean_lp(torch.cat([running, beam_sc], 1))
                post_synth = torch.cat(tokens + [beam_tok[:, None], tiled], 1)
                win = pre_score > post_score
                score = torch.where(win, pre_score, post_score)
                synth = torch.where(win[:, None], pre_synth, post_synth)
            else:
                token = stp['argmax'].reshape(B, 1 + C)[:, 0].long()
                score, synth = pre_score, pre_synth
            synth_sofar = torch.cat([synth_sofar, eos_col], 1)
            toks = torch.cat(tokens, 1)
            if toks.shape[1] >= C:                                           # _already pv:495-498
                alive = alive * (~(toks[:, -C:] == idc).all(1)).float()
            upd = (score > score_sofar) & (alive > 0)
            synth_sofar = torch.where(upd[:, None], synth, synth_sofar)
            score_sofar = torch.where(upd, score, score_sofar)
            running = torch.cat([running, first_max_lp[:, None]], 1)         # lp of the greedy token (pv:592)
            alive = alive * (token != ID_EOS).float()
            if (i + 1) % alive_check_every == 0 and float(alive.sum().item()) < 1:
                break
        constrain_time += tic()
        out_ids = torch.cat([dict_input['input_ids'].to(dev), synth_sofar], 1).tolist()        # the one read-back per constraint
        out_ids = [(r[:r.index(ID_EOS, S)] if ID_EOS in r[S:] else r) for r in out_ids]
        out_ids = [[t for t in r if t not in (0, 1)] for r in out_ids]
        ids_trace.append(out_ids)
        output = [_preprocess(s) for s in processor.tokenizer.batch_decode(out_ids)]
        prompt = output
    if return_ids:
        return ids_trace
    if not return_full_text:
        output = [o[l:] for o, l in zip(output, len_ps)]
    if not mute:
        if was_str:
            print(output[0])
        else:
            for i, o in enumerate(output):
                print(f'\n< Constrained text for prompt #{i} >\n{o}')
    if verbose:
        print(f'Prompt: {prompt_time:.2f} sec\nConstrain: {constrain_time:.2f} sec')
    return output[0] if was_str else output


def constrain(prompt, constraints=[(0, '\nThe'), (100, ' The correct answer is'), 'ABCDE'], images=None, preload=None,
              blind_model=False, quantize_model=False, quantize_cache=False, use_adapter=False, verbose=True,
              apply_chat_template=True, use_beam=False, n_beam=3):
    """pv:1425-1487 (+ n_beam, an extension: the reference hard-codes 3, pv:505)."""
    if preload is None:
        preload = load(blind_model=blind_model, quantize_model=quantize_model, quantize_cache=quantize_cache,
                       use_adapter=use_adapter)
    prompt, images = _apply_chat_template(prompt, images, verbose, apply_chat_template)
    return _constrain(*preload, prompt, constraints, use_beam=use_beam, verbose=verbose, n_beam=n_beam)
