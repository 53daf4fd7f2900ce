"""phi3_b200 — B200-native inference hot path of Phi-3-Vision-MLX (vision encode, prefill,
batched decode) behind the reference's Python API. Import as `import phi3_b200`.

Submodules are imported lazily so that CPU-only tooling (configs, weights, processor index
math) works without the CUDA extension; anything that computes raises if it is missing.
"""
from . import configs  # noqa: F401

__all__ = ['load', 'generate', 'choose', 'constrain', 'generate_batch', 'sanitize', 'configs']


def __getattr__(name):
    if name in ('load', 'generate', 'choose', 'constrain', 'generate_batch', 'sanitize', '_generate', '_choose_from', '_constrain'):
        from . import api
        return getattr(api, name)
    if name in ('api', 'model', 'processor', 'weights', '_lib', 'parallel', 'mega', 'quant', 'server'):
        import importlib
        return importlib.import_module(f'{__name__}.{name}')
    raise AttributeError(name)
