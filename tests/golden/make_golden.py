"""Generates tests/golden/processor_golden.json + processor_golden.npz by executing the
REFERENCE's own processors (/root/reference/phi.py, unmodified) with mlx / matplotlib stubbed
(recipe: SURVEY.md A.8). Run in the build container only; the GPU box has no /root/reference.
    python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_phi():
    import transformers  # noqa: F401  (must be imported before the stubs)
    from transformers import AutoTokenizer  # noqa: F401  (force the lazy import before mlx is stubbed)

    class Anything:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return self
        def __getattr__(self, n): return Anything()
        def __mro_entries__(self, bases): return (object,)

    for name in ['mlx', 'mlx.core', 'mlx.nn', 'mlx.utils', 'mlx.optimizers', 'matplotlib', 'matplotlib.pyplot']:
        m = types.ModuleType(name)
        m.__getattr__ = lambda n: Anything()
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location('ref_phi', '/root/reference/phi.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.mx.array = np.array
    return mod


class FakeTok:
    """BOS=1 then ord(ch)+50 per character; same fake used by tests."""
    class E:
        pass

    def __call__(self, texts):
        e = self.E()
        enc = lambda t: [1] + [ord(c) + 50 for c in t]
        e.input_ids = enc(texts) if isinstance(texts, str) else [enc(t) for t in texts]
        return e


def main():
    from PIL import Image
    phi = load_reference_phi()
    rng = np.random.RandomState(0)
    gold = {'geometry': [], 'tokenize': {}, 'merge': {}}
    arrays = {}
    sizes = [(336, 336), (672, 672), (1344, 1344), (1000, 700), (700, 1000), (640, 480), (1920, 1080), (300, 200),
             (1600, 1000), (1344, 336), (336, 1344), (800, 500), (37, 1200)]
    real_interp = phi.Phi3VImageProcessor.interpolate_336
    for n_crops in (16, 4):
        for (w, h) in sizes:
            ip = phi.Phi3VImageProcessor()
            ip.num_crops = n_crops
            img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
            phi.Phi3VImageProcessor.interpolate_336 = staticmethod(lambda x: np.zeros((1, 3, 336, 336)))
            try:
                out = ip([img])
                gold['geometry'].append({'w': w, 'h': h, 'num_crops': n_crops, 'image_sizes': out['image_sizes'][0],
                                         'num_img_tokens': out['num_img_tokens'][0],
                                         'pv_shape': list(out['pixel_values'].shape)})
            except Exception as ex:  # the reference itself fails (e.g. scale 0)
                gold['geometry'].append({'w': w, 'h': h, 'num_crops': n_crops, 'error': type(ex).__name__})
    phi.Phi3VImageProcessor.interpolate_336 = staticmethod(real_interp)
    # full pixel-level goldens (with the real interpolate_336 loop: ~8 s each)
    for tag, (w, h), n_crops in [('a', (500, 350), 4), ('b', (300, 420), 4), ('c', (640, 480), 16)]:
        ip = phi.Phi3VImageProcessor()
        ip.num_crops = n_crops
        arr = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        out = ip([Image.fromarray(arr)])
        pv = out['pixel_values'][0]
        n = out['image_sizes'][0][0] // 336 * out['image_sizes'][0][1] // 336 + 1
        arrays[f'img_{tag}'] = arr
        # fixtures stay small: a 1/35 pixel lattice of every used crop + exact per-crop float64 sums
        arrays[f'pv_{tag}'] = pv[:n, :, ::7, ::5].astype(np.float32)
        arrays[f'pvsum_{tag}'] = pv[:n].sum(axis=(1, 2, 3))
        arrays[f'pvpad_{tag}'] = np.array([np.abs(pv[n:]).sum()])
        gold[f'pixel_{tag}'] = {'num_crops': n_crops, 'image_sizes': out['image_sizes'][0], 'n_used': int(n),
                                'num_img_tokens': out['num_img_tokens'][0]}
    # interpolate_336 weight tables
    for in_size in (336, 672, 1008, 1344, 1680):
        x = np.zeros((1, 1, in_size, in_size))
        # recover the tables by probing with delta images is expensive; re-run the nested helper instead
        src = phi.Phi3VImageProcessor.interpolate_336
        import inspect
        code = inspect.getsource(src)
        ns = {'np': np}
        body = code.split('def get_weights_and_indices')[1].split('N, C, H, W = input.shape')[0]
        exec('def get_weights_and_indices' + '\n'.join(l[8:] if l.startswith('        ') else l for l in body.split('\n')), ns)
        wts, idx = ns['get_weights_and_indices'](336 / in_size, 336, in_size)
        arrays[f'i336_w_{in_size}'] = wts
        arrays[f'i336_i_{in_size}'] = idx
    # tokenizer / merge layouts
    fp = phi.Phi3FProcessor.__new__(phi.Phi3FProcessor)
    fp.tokenizer, fp.return_mx = FakeTok(), True
    o = fp._tokenize(['abc', 'a', 'hello'])
    gold['tokenize'] = {k: np.asarray(v).tolist() for k, v in o.items()}
    vp = phi.Phi3VProcessor.__new__(phi.Phi3VProcessor)
    vp.tokenizer, vp.return_mx = FakeTok(), True
    text = "<|user|>\n<|image_1|>\n<|image_2|>\nhi"
    o = vp._merge({'pixel_values': np.zeros((2, 1)), 'image_sizes': [[336, 336], [336, 336]], 'num_img_tokens': [5, 3]}, text)
    gold['merge'] = {'text': text, 'input_ids': np.asarray(o['input_ids']).tolist(),
                     'positions': np.asarray(o['positions']).tolist()}
    json.dump(gold, open(os.path.join(HERE, 'processor_golden.json'), 'w'), indent=1)
    np.savez_compressed(os.path.join(HERE, 'processor_golden.npz'), **arrays)
    print('wrote goldens:', len(gold['geometry']), 'geometry rows;', list(arrays))


if __name__ == '__main__':
    main()
