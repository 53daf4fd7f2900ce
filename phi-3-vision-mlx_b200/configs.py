"""Frozen hyper-parameters for the two checkpoints the reference serves.

The reference reads these at run time from the downloaded HF ``config.json``
(/root/reference/phi_3_vision_mlx.py:258,359-363); none is in its tree, so the
public Phi-3.5-mini-instruct / Phi-3.5-vision-instruct values are frozen here.
LongRoPE factor vectors: /root/reference/assets/su_rope_explained.ipynb:309-310.
CLIP constants: /root/reference/phi.py:375-384.
"""
from types import SimpleNamespace
import copy

SHORT_FACTOR = [1.05, 1.05, 1.05, 1.1, 1.1, 1.1, 1.2500000000000002, 1.2500000000000002, 1.4000000000000004,
                1.4500000000000004, 1.5500000000000005, 1.8500000000000008, 1.9000000000000008, 2.000000000000001,
                2.000000000000001, 2.000000000000001, 2.000000000000001, 2.000000000000001, 2.000000000000001,
                2.000000000000001, 2.000000000000001, 2.000000000000001, 2.000000000000001, 2.000000000000001,
                2.000000000000001, 2.000000000000001, 2.000000000000001, 2.000000000000001, 2.000000000000001,
                2.000000000000001, 2.000000000000001, 2.000000000000001, 2.1000000000000005, 2.1000000000000005, 2.2,
                2.3499999999999996, 2.3499999999999996, 2.3499999999999996, 2.3499999999999996, 2.3999999999999995,
                2.3999999999999995, 2.6499999999999986, 2.6999999999999984, 2.8999999999999977, 2.9499999999999975,
                3.049999999999997, 3.049999999999997, 3.049999999999997]
LONG_FACTOR = [1.0299999713897705, 1.0499999523162842, 1.0499999523162842, 1.0799999237060547, 1.2299998998641968,
               1.2299998998641968, 1.2999999523162842, 1.4499999284744263, 1.5999999046325684, 1.6499998569488525,
               1.8999998569488525, 2.859999895095825, 3.68999981880188, 5.419999599456787, 5.489999771118164,
               5.489999771118164, 9.09000015258789, 11.579999923706055, 15.65999984741211, 15.769999504089355,
               15.789999961853027, 18.360000610351562, 21.989999771118164, 23.079999923706055, 30.009998321533203,
               32.35000228881836, 32.590003967285156, 35.56000518798828, 39.95000457763672, 53.840003967285156,
               56.20000457763672, 57.95000457763672, 59.29000473022461, 59.77000427246094, 59.920005798339844,
               61.190006256103516, 61.96000671386719, 62.50000762939453, 63.3700065612793, 63.48000717163086,
               63.48000717163086, 63.66000747680664, 63.850006103515625, 64.08000946044922, 64.760009765625,
               64.80001068115234, 64.81001281738281, 64.81001281738281]

ID_EOS = 32007   # /root/reference/phi_3_vision_mlx.py:42
ID_ASS = 32001   # /root/reference/phi_3_vision_mlx.py:43

CLIP_VIT_L14_336 = SimpleNamespace(hidden_size=1024, image_size=336, intermediate_size=4096, layer_norm_eps=1e-5,
                                   num_attention_heads=16, num_channels=3, num_hidden_layers=24, patch_size=14)

PHI35_MINI = SimpleNamespace(
    architectures=["Phi3ForCausalLM"], hidden_size=3072, num_hidden_layers=32, num_attention_heads=32,
    num_key_value_heads=32, intermediate_size=8192, vocab_size=32064, rms_norm_eps=1e-5, rope_theta=10000.0,
    max_position_embeddings=131072, original_max_position_embeddings=4096,
    rope_scaling={"type": "su", "short_factor": SHORT_FACTOR, "long_factor": LONG_FACTOR},
    use_quantized_cache=False)

PHI35_VISION = copy.deepcopy(PHI35_MINI)
PHI35_VISION.architectures = ["Phi3VForCausalLM"]
PHI35_VISION.img_processor = {"image_dim_out": 1024, "num_img_tokens": 144}


def tiny(vision=False, layers=2, **kw):
    """Small same-shaped-head config (head_dim stays 96) for quick parity runs."""
    c = copy.deepcopy(PHI35_VISION if vision else PHI35_MINI)
    c.hidden_size, c.num_attention_heads, c.num_key_value_heads = 384, 4, 4
    c.intermediate_size, c.num_hidden_layers, c.vocab_size = 1024, layers, 32064
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def tiny_clip(layers=3):
    c = copy.deepcopy(CLIP_VIT_L14_336)
    c.num_hidden_layers = layers
    return c


def with_overrides(cfg, **kw):
    c = copy.deepcopy(cfg)
    for k, v in kw.items():
        setattr(c, k, v)
    return c
