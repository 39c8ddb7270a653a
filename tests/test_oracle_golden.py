"""CPU: pin the oracle against golden vectors produced by executing the reference's own sources
(oracle/make_golden.py) -- `btsbot/architectures.py` + timm shim, `btsbot/alert_utils.py`."""
import numpy as np
import pytest
import torch

from btsbot_b200 import synth
from oracle import convnext_oracle as O
from oracle import preprocess_oracle as P
from cases import MODEL_CASES, case_state_dict


@pytest.mark.parametrize("case", list(MODEL_CASES))
def test_oracle_matches_reference_logits(case, golden_logits, golden_batch):
    img, meta = golden_batch
    cfg, sd = case_state_dict(case, golden_logits)
    out = O.forward(synth.to_torch(sd), cfg, torch.from_numpy(img), torch.from_numpy(meta)).numpy()
    ref = golden_logits[case]
    assert out.shape == ref.shape == (img.shape[0], 1)
    # two independent fp32 CPU implementations (torchvision modules vs functional restatement)
    assert np.abs(out - ref).max() < 5e-5
    # H2: the calibrated logits straddle 0 with a spread far above the tolerances used on the GPU
    assert (ref > 0).mean() == pytest.approx(0.5, abs=0.05) and ref.std() > 0.1


@pytest.mark.parametrize("case", list(MODEL_CASES))
def test_reference_owned_state_dict_keys(case, golden_logits):
    """Keys the reference's own modules create (metadata branch, heads, head surgery) must exist in ours."""
    import btsbot_b200 as b
    cfg, sd = case_state_dict(case, golden_logits)
    model = getattr(b, cfg["model_name"])(cfg)
    ours = set(model.state_dict().keys())
    for k in golden_logits[case + "_keys"]:
        assert str(k) in ours, k
    assert ours == set(sd.keys())
    model.load_state_dict(synth.to_torch(sd), strict=True)


@pytest.mark.parametrize("s", [63, 49, 32, 31])
def test_crop_oracle_matches_reference(s, golden_pre):
    t = synth.make_triplets(2, start=5000, dtype=np.float64) * 37.5
    assert np.array_equal(P.crop_triplets(t.copy(), s), golden_pre[f"crop{s}_f64"])
    got32 = P.crop_triplets(t.astype(np.float32), s)
    assert np.array_equal(got32, golden_pre[f"crop{s}_f32"])
    assert (63 - s) // 2 == {63: 0, 49: 7, 32: 15, 31: 16}[s]


def test_tail_oracle_matches_reference(golden_pre):
    from oracle.make_golden import adversarial_stamps
    drops = []
    for i, stamps in enumerate(adversarial_stamps()):
        trip, drop = P.triplet_tail(stamps, normalize=True)
        ref = golden_pre[f"tail{i}"]
        assert np.array_equal(np.isnan(trip), np.isnan(ref))
        assert np.array_equal(np.nan_to_num(trip), np.nan_to_num(ref)), i
        assert bool(golden_pre[f"tail{i}_drop"]) == drop
        drops.append(drop)
    # pad value is float32(1e-9) stored in float64, applied after normalisation
    assert golden_pre["tail1"][62, 62, 0] == np.float64(np.float32(1e-9))
    assert drops == [False, False, False, False, False, True]


def test_synth_is_sharding_independent():
    a = synth.make_triplets(300, start=100)
    b = np.concatenate([synth.make_triplets(100, start=100), synth.make_triplets(200, start=200)])
    assert np.array_equal(a, b)
    m = synth.make_metadata(600, start=7)
    assert np.array_equal(m[250:300], synth.make_metadata(50, start=257))
    n = np.sqrt((a.astype(np.float64) ** 2).sum(axis=(1, 2)))
    assert np.allclose(n, 1.0, atol=1e-5)


@pytest.mark.parametrize("kind", ["convnext_nano.d1h_in1k", "convnext_pico.d1_in1k"])
def test_trunk_oracle_matches_hf_transformers_convnext(kind):
    """Third independent implementation of the ConvNeXt trunk arithmetic (after the torchvision one behind the goldens):
    Hugging Face `transformers.ConvNextModel`, a port of the original FAIR code timm's model also descends from.  Same
    weights through a key remap -> same [B, C3, 1, 1] feature map as the functional restatement.  (timm itself is not
    installable offline; this does not pin timm, it pins the block semantics: stem conv4/s4 + LN2d,
    dw7x7 -> LN -> fc1 -> exact GELU -> fc2 -> layer scale -> + shortcut, LN2d + conv2/s2 downsampling, eps 1e-6.)"""
    transformers = pytest.importorskip("transformers")
    cfg = synth.canonical_config("mm_ConvNeXt", kind)
    sd = synth.to_torch(synth.make_state_dict(cfg, seed=7))
    arch = O.arch_of(kind)
    hc = transformers.ConvNextConfig(num_channels=3, patch_size=4, num_stages=4, hidden_sizes=list(arch["dims"]),
                                     depths=list(arch["depths"]), hidden_act="gelu", layer_scale_init_value=1.0,
                                     drop_path_rate=0.0, image_size=63)
    hf = transformers.ConvNextModel(hc).eval()
    p = "convnext_backbone."
    remap = {"embeddings.patch_embeddings.": p + "stem.0.", "embeddings.layernorm.": p + "stem.1."}
    new = {}
    for k in hf.state_dict():
        if k.startswith("layernorm."):                     # HF's pooler LayerNorm: not part of forward_features
            new[k] = hf.state_dict()[k]
            continue
        src = None
        for a, b in remap.items():
            if k.startswith(a):
                src = sd[b + k[len(a):]]
        if src is None:
            _, _, i, kind_, *rest = k.split(".")            # encoder.stages.{i}.(downsampling_layer|layers).…
            if kind_ == "downsampling_layer":
                src = sd[f"{p}stages.{i}.downsample.{rest[0]}.{rest[1]}"]
            else:
                j, name = rest[0], ".".join(rest[1:])
                q = f"{p}stages.{i}.blocks.{j}."
                src = {"layer_scale_parameter": lambda: sd[q + "gamma"],
                       "dwconv.weight": lambda: sd[q + "conv_dw.weight"], "dwconv.bias": lambda: sd[q + "conv_dw.bias"],
                       "layernorm.weight": lambda: sd[q + "norm.weight"], "layernorm.bias": lambda: sd[q + "norm.bias"],
                       "pwconv1.weight": lambda: sd[q + "mlp.fc1.weight"].flatten(1),      # Conv2d 1x1 -> Linear
                       "pwconv1.bias": lambda: sd[q + "mlp.fc1.bias"],
                       "pwconv2.weight": lambda: sd[q + "mlp.fc2.weight"].flatten(1),
                       "pwconv2.bias": lambda: sd[q + "mlp.fc2.bias"]}[name]()
        assert tuple(src.shape) == tuple(hf.state_dict()[k].shape), k
        new[k] = src
    hf.load_state_dict(new, strict=True)
    n = 6
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(n, start=70).transpose(0, 3, 1, 2)))
    with torch.no_grad():
        ref = hf(pixel_values=img).last_hidden_state
        got = O.trunk_features(sd, p, img, arch)
    assert got.shape == ref.shape == (n, arch["dims"][-1], 1, 1)
    rel = ((got - ref).abs().max() / ref.abs().max()).item()
    assert rel < 2e-5, rel
