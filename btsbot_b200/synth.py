"""Synthetic ZTF-alert inputs and perturbed random-init weights.

The reference ships no data generator; shapes, dtypes and statistics follow the
fixture it does ship (`btsbot/example_data/usage_triplets.npy` 39x63x63x3 float64 HWC,
`usage_candidates.csv`) and the column order of `btsbot/inference_example.py:53-58`.

Everything is index-keyed (numpy Philox keyed by ``(seed, chunk)`` with a fixed chunk
of 256 alerts), so alert ``i`` has the same content no matter how a job is sharded
over GPUs.  ``seed=2`` is the reference's convention
(`btsbot/train_configs/prod_config.json:67`).

Weights are generated with numpy (stream-stable across versions) and shared between
the oracle and the device path through a ``state_dict`` -- never through a seed.  At
plain random init timm's layer-scale ``gamma = 1e-6`` hides every block, so
:func:`make_state_dict` perturbs ``gamma``, the LN/BN affines and the BN running
statistics.
"""
from __future__ import annotations

import numpy as np

CHUNK = 256
IMG = 63

#: column order of `btsbot/inference_example.py:53-58`
METADATA_COLS = [
    "sgscore1", "distpsnr1", "sgscore2", "distpsnr2", "fwhm", "magpsf",
    "sigmapsf", "chipsf", "ra", "dec", "diffmaglim", "ndethist", "nmtchps",
    "age", "days_since_peak", "days_to_peak", "peakmag_so_far", "new_drb",
    "ncovhist", "nnotdet", "chinr", "sharpnr", "scorr", "sky", "maxmag_so_far",
]

#: (mean, std, min, max) measured on the 39 rows of usage_candidates.csv
METADATA_MOMENTS = np.array([
    (0.260872, 0.245833, 0.0558333, 0.555616),
    (1.50073, 1.2145, 0.00878181, 2.65519),
    (0.668644, 0.126124, 0.5, 0.800554),
    (15.2404, 3.09111, 12.5353, 19.0343),
    (3.29916, 0.952506, 1.19734, 5.85),
    (17.8588, 1.0881, 15.6103, 19.9016),
    (0.0761297, 0.0455793, 0.0275967, 0.19425),
    (25.9566, 62.0336, 1.10579, 328.73),
    (124.603, 23.0134, 97.0111, 143.799),
    (82.3717, 1.92047, 80.0691, 83.9735),
    (19.8765, 0.377509, 19.1472, 20.4643),
    (14.2308, 8.44123, 1, 29),
    (4.23077, 1.47564, 3, 6),
    (16.3563, 10.8495, 0, 42.0488),
    (12.9697, 10.4395, 0, 38.0494),
    (3.38658, 0.802938, 0, 3.99938),
    (16.9888, 1.1442, 15.6103, 18.2745),
    (0.99877, 0.00391913, 0.983447, 1),
    (646.744, 68.5673, 573, 742),
    (632.513, 68.0007, 572, 714),
    (3.23046, 2.00686, 0.775, 5.177),
    (0.126692, 0.281406, -0.298, 0.392),
    (33.8914, 29.5137, 5.86059, 131.85),
    (0.130772, 0.430285, -0.878342, 1.04186),
    (17.9236, 1.08688, 15.7521, 19.9016),
], dtype=np.float64)


def _rng(seed: int, stream: int, chunk: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[(seed << 8) | stream, chunk]))


def _chunks(start: int, n: int):
    first, last = start // CHUNK, (start + n - 1) // CHUNK if n > 0 else -1
    for c in range(first, last + 1):
        lo = max(start, c * CHUNK) - c * CHUNK
        hi = min(start + n, (c + 1) * CHUNK) - c * CHUNK
        yield c, lo, hi


def make_triplets(n: int, start: int = 0, seed: int = 2, dtype=np.float32) -> np.ndarray:
    """``[n,63,63,3]`` HWC science/reference/difference cutouts, each L2-normalised.

    sci/ref = 1 + 0.07 N(0,1) + PSF, diff = N(0,1) + PSF * U(-1,8); statistics match the
    shipped example data (sci/ref mean ~ 1/63, diff mean ~ 0).
    """
    out = np.empty((n, IMG, IMG, 3), dtype=dtype)
    yy, xx = np.mgrid[0:IMG, 0:IMG].astype(np.float64)
    r2 = (yy - 31.0) ** 2 + (xx - 31.0) ** 2
    pos = 0
    for c, lo, hi in _chunks(start, n):
        g = _rng(seed, 1, c)
        noise = g.standard_normal((CHUNK, IMG, IMG, 3))
        sigma = g.uniform(1.0, 2.5, (CHUNK, 1, 1))
        amp = g.uniform(0.0, 6.0, (CHUNK, 1, 1))
        dscale = g.uniform(-1.0, 8.0, (CHUNK, 1, 1))
        psf = amp * np.exp(-r2[None] / (2.0 * sigma ** 2))
        t = np.empty((CHUNK, IMG, IMG, 3))
        t[..., 0] = 1.0 + 0.07 * noise[..., 0] + psf
        t[..., 1] = 1.0 + 0.07 * noise[..., 1] + psf
        t[..., 2] = noise[..., 2] + psf * dscale
        t /= np.sqrt((t ** 2).sum(axis=(1, 2), keepdims=True))
        k = hi - lo
        out[pos:pos + k] = t[lo:hi].astype(dtype)
        pos += k
    return out


def make_metadata(n: int, start: int = 0, seed: int = 2) -> np.ndarray:
    """``[n,25]`` float32, columns in :data:`METADATA_COLS` order, un-normalised on purpose
    (that is what ``BatchNorm1d`` sees in the reference)."""
    out = np.empty((n, len(METADATA_COLS)), dtype=np.float32)
    mu, sd, lo_, hi_ = (METADATA_MOMENTS[:, i] for i in range(4))
    pos = 0
    for c, lo, hi in _chunks(start, n):
        g = _rng(seed, 2, c)
        v = np.clip(mu + sd * g.standard_normal((CHUNK, len(METADATA_COLS))), lo_, hi_)
        k = hi - lo
        out[pos:pos + k] = v[lo:hi].astype(np.float32)
        pos += k
    return out


def make_labels(n: int, start: int = 0, seed: int = 2) -> np.ndarray:
    """Bernoulli(0.5) labels, int64 (reference labels are ``torch.long``, train.py:135)."""
    out = np.empty((n,), dtype=np.int64)
    pos = 0
    for c, lo, hi in _chunks(start, n):
        v = (_rng(seed, 3, c).random(CHUNK) < 0.5).astype(np.int64)
        k = hi - lo
        out[pos:pos + k] = v[lo:hi]
        pos += k
    return out


# ---------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------

CONVNEXT_KINDS = {
    # timm convnext_nano* / convnext_pico* 
    "convnext_nano": dict(dims=(80, 160, 320, 640), depths=(2, 2, 8, 2)),
    "convnext_pico": dict(dims=(64, 128, 256, 512), depths=(2, 2, 6, 2)),
}


def convnext_arch(model_kind: str) -> dict:
    """Map a timm model name (``convnext_nano.d1h_in1k``, ``hf_hub:mwalmsley/zoobot-encoder-convnext_pico`` ...)
    to dims/depths."""
    k = model_kind.lower()
    for name, arch in CONVNEXT_KINDS.items():
        if name in k:
            return dict(arch, name=name)
    raise ValueError(f"unsupported ConvNeXt kind for the B200 path: {model_kind!r} "
                     f"(supported: {sorted(CONVNEXT_KINDS)})")


MAXVIT_KINDS = {
    # timm maxvit_tiny_rw_224
    "maxvit_tiny_rw": dict(embed_dim=(64, 128, 256, 512), depths=(2, 2, 5, 2), stem_width=(32, 64), dim_head=32,
                           window=7, expand=4, se_div=16, img=224),
}


def maxvit_arch(model_kind: str) -> dict:
    """Map a timm model name (``maxvit_tiny_rw_224.sw_in1k``, ``hf_hub:mwalmsley/baseline-encoder-regression-maxvit_tiny``
    -- the galaxyzoo-pretrained encoder of to_HF.py:167-168, a maxvit_tiny_rw_224 -- ...) to its MaxViT configuration."""
    k = model_kind.lower()
    for name, arch in MAXVIT_KINDS.items():
        if name in k or (name == "maxvit_tiny_rw" and k.rstrip("/").endswith("maxvit_tiny")):
            return dict(arch, name=name)
    raise ValueError(f"unsupported MaxViT kind for the B200 path: {model_kind!r} (supported: {sorted(MAXVIT_KINDS)})")


class _W:
    """numpy weight factory; ``perturb=False`` reproduces timm/torch default init statistics."""

    def __init__(self, seed: int, perturb: bool):
        self.g = np.random.Generator(np.random.Philox(key=[seed, 0xB75B07]))
        self.perturb = perturb

    def dense(self, *shape, fan_in: int):
        std = 1.0 / np.sqrt(fan_in) if self.perturb else 0.02
        w = self.g.standard_normal(shape) * std
        if not self.perturb:
            w = np.clip(w, -2.0, 2.0)
        return w.astype(np.float32)

    def bias(self, n):
        return (self.g.standard_normal(n) * 0.1).astype(np.float32) if self.perturb \
            else np.zeros(n, np.float32)

    def scale(self, n, base=1.0):
        return (base * self.g.uniform(0.5, 1.5, n)).astype(np.float32) if self.perturb \
            else np.full(n, base, np.float32)


def _convnext_trunk_sd(w: _W, prefix: str, arch: dict, sd: dict, gamma_base: float):
    dims, depths = arch["dims"], arch["depths"]
    sd[f"{prefix}stem.0.weight"] = w.dense(dims[0], 3, 4, 4, fan_in=48)
    sd[f"{prefix}stem.0.bias"] = w.bias(dims[0])
    sd[f"{prefix}stem.1.weight"] = w.scale(dims[0])
    sd[f"{prefix}stem.1.bias"] = w.bias(dims[0])
    for i, (c, d) in enumerate(zip(dims, depths)):
        if i > 0:
            cin = dims[i - 1]
            sd[f"{prefix}stages.{i}.downsample.0.weight"] = w.scale(cin)
            sd[f"{prefix}stages.{i}.downsample.0.bias"] = w.bias(cin)
            sd[f"{prefix}stages.{i}.downsample.1.weight"] = w.dense(c, cin, 2, 2, fan_in=4 * cin)
            sd[f"{prefix}stages.{i}.downsample.1.bias"] = w.bias(c)
        for j in range(d):
            p = f"{prefix}stages.{i}.blocks.{j}."
            sd[p + "gamma"] = w.scale(c, gamma_base) if w.perturb else np.full(c, 1e-6, np.float32)
            sd[p + "conv_dw.weight"] = w.dense(c, 1, 7, 7, fan_in=49)
            sd[p + "conv_dw.bias"] = w.bias(c)
            sd[p + "norm.weight"] = w.scale(c)
            sd[p + "norm.bias"] = w.bias(c)
            sd[p + "mlp.fc1.weight"] = w.dense(4 * c, c, 1, 1, fan_in=c)
            sd[p + "mlp.fc1.bias"] = w.bias(4 * c)
            sd[p + "mlp.fc2.weight"] = w.dense(c, 4 * c, 1, 1, fan_in=4 * c)
            sd[p + "mlp.fc2.bias"] = w.bias(c)


def _bn_sd(w: _W, prefix: str, n: int, sd: dict, data_std: float | None = None):
    sd[prefix + "weight"] = w.scale(n)
    sd[prefix + "bias"] = w.bias(n)
    if w.perturb and data_std is not None:
        # running statistics at the scale of the layer's actual input (a trained net's would be): L2-normalised
        # cutouts have pixel values ~1/63, so a unit-variance BatchNorm would drown the image in its own bias
        sd[prefix + "running_mean"] = (data_std * w.g.standard_normal(n)).astype(np.float32)
        sd[prefix + "running_var"] = (data_std ** 2 * w.g.uniform(0.7, 1.3, n)).astype(np.float32)
    elif w.perturb and n == len(METADATA_COLS):
        # running stats near the column moments so normalised metadata is O(1) and logits straddle 0
        mu, sdv = METADATA_MOMENTS[:, 0], METADATA_MOMENTS[:, 1]
        sd[prefix + "running_mean"] = (mu + 0.1 * sdv * w.g.standard_normal(n)).astype(np.float32)
        sd[prefix + "running_var"] = ((sdv ** 2) * w.g.uniform(0.7, 1.3, n)).astype(np.float32)
    elif w.perturb:
        sd[prefix + "running_mean"] = (0.1 * w.g.standard_normal(n)).astype(np.float32)
        sd[prefix + "running_var"] = w.g.uniform(0.7, 1.3, n).astype(np.float32)
    else:
        sd[prefix + "running_mean"] = np.zeros(n, np.float32)
        sd[prefix + "running_var"] = np.ones(n, np.float32)
    sd[prefix + "num_batches_tracked"] = np.array(0 if not w.perturb else 17, dtype=np.int64)


def _maxvit_trunk_sd(w: _W, prefix: str, arch: dict, sd: dict, branch_gain: float):
    """timm MaxxVit ('maxvit_tiny_rw_224') parameter tree.  The net has no layer scale, so the last linear map of each
    residual branch (conv3_1x1, attn.proj, mlp.fc2) is damped by ``branch_gain`` to keep the residual stream O(1)."""
    sw = arch["stem_width"]
    sd[f"{prefix}stem.conv1.weight"] = w.dense(sw[0], 3, 3, 3, fan_in=27)
    _bn_sd(w, f"{prefix}stem.norm1.", sw[0], sd, data_std=0.01)
    sd[f"{prefix}stem.conv2.weight"] = w.dense(sw[1], sw[0], 3, 3, fan_in=9 * sw[0])
    cin = sw[1]
    g = branch_gain if w.perturb else 1.0
    for i, (c, d) in enumerate(zip(arch["embed_dim"], arch["depths"])):
        heads = c // arch["dim_head"]
        for j in range(d):
            p = f"{prefix}stages.{i}.blocks.{j}."
            mid, rd = arch["expand"] * cin, arch["expand"] * cin // arch["se_div"]
            q = p + "conv."
            if j == 0 and cin != c:
                sd[q + "shortcut.expand.weight"] = w.dense(c, cin, 1, 1, fan_in=cin)
            _bn_sd(w, q + "pre_norm.", cin, sd)
            sd[q + "conv1_1x1.weight"] = w.dense(mid, cin, 1, 1, fan_in=cin)
            _bn_sd(w, q + "norm1.", mid, sd)
            sd[q + "conv2_kxk.weight"] = w.dense(mid, 1, 3, 3, fan_in=9)
            _bn_sd(w, q + "norm2.", mid, sd)
            sd[q + "se.fc1.weight"] = w.dense(rd, mid, 1, 1, fan_in=mid)
            sd[q + "se.fc1.bias"] = w.bias(rd)
            sd[q + "se.fc2.weight"] = w.dense(mid, rd, 1, 1, fan_in=rd)
            sd[q + "se.fc2.bias"] = w.bias(mid)
            sd[q + "conv3_1x1.weight"] = w.dense(c, mid, 1, 1, fan_in=mid) * np.float32(g)
            cin = c
            for part in ("attn_block.", "attn_grid."):
                q = p + part
                sd[q + "norm1.weight"], sd[q + "norm1.bias"] = w.scale(c), w.bias(c)
                _linear_sd(w, q + "attn.qkv.", 3 * c, c, sd)
                n_rel = (2 * arch["window"] - 1) ** 2
                sd[q + "attn.rel_pos.relative_position_bias_table"] = \
                    (w.g.standard_normal((n_rel, heads)) * (0.5 if w.perturb else 0.02)).astype(np.float32)
                _linear_sd(w, q + "attn.proj.", c, c, sd)
                sd[q + "attn.proj.weight"] *= np.float32(g)
                sd[q + "norm2.weight"], sd[q + "norm2.bias"] = w.scale(c), w.bias(c)
                _linear_sd(w, q + "mlp.fc1.", 4 * c, c, sd)
                _linear_sd(w, q + "mlp.fc2.", c, 4 * c, sd)
                sd[q + "mlp.fc2.weight"] *= np.float32(g)
    sd[f"{prefix}norm.weight"] = w.scale(arch["embed_dim"][-1])
    sd[f"{prefix}norm.bias"] = w.bias(arch["embed_dim"][-1])


def _linear_sd(w: _W, prefix: str, nout: int, nin: int, sd: dict):
    sd[prefix + "weight"] = w.dense(nout, nin, fan_in=nin)
    sd[prefix + "bias"] = w.bias(nout)


def make_state_dict(config: dict, seed: int = 2, perturb: bool = True, gamma_base: float = 0.5) -> dict:
    """State dict (numpy arrays, timm/reference key names) for ``config['model_name']`` in
    {mm_ConvNeXt, ConvNeXt, um_nn, frozen_fusion, mm_MaxViT, MaxViT}.  Keys follow `btsbot/architectures.py:104-171,277-372`
    and timm's ConvNeXt."""
    w = _W(seed, perturb)
    sd: dict = {}
    name = config["model_name"]
    nmeta = len(config.get("metadata_cols", []))
    if name == "mm_ConvNeXt":
        arch = convnext_arch(config.get("model_kind", "convnext_nano.d1h_in1k"))
        _convnext_trunk_sd(w, "convnext_backbone.", arch, sd, gamma_base)
        feat = arch["dims"][-1]
        if "LS" in config["train_data_version"]:
            sd["convnext_backbone.head.1.weight"] = w.scale(feat)
            sd["convnext_backbone.head.1.bias"] = w.bias(feat)
        _bn_sd(w, "metadata_branch.0.", nmeta, sd)
        _linear_sd(w, "metadata_branch.1.", config["meta_fc1_neurons"], nmeta, sd)
        _linear_sd(w, "metadata_branch.4.", config["meta_fc2_neurons"], config["meta_fc1_neurons"], sd)
        _linear_sd(w, "combined_head.0.", config["comb_fc1_neurons"], feat + config["meta_fc2_neurons"], sd)
        _linear_sd(w, "combined_head.2.", config["comb_fc2_neurons"], config["comb_fc1_neurons"], sd)
        _linear_sd(w, "combined_head.5.", 1, config["comb_fc2_neurons"], sd)
    elif name == "ConvNeXt":
        arch = convnext_arch(config.get("model_kind", "convnext_nano.d1h_in1k"))
        _convnext_trunk_sd(w, "convnext.", arch, sd, gamma_base)
        feat = arch["dims"][-1]
        sd["convnext.head.1.weight"] = w.scale(feat)
        sd["convnext.head.1.bias"] = w.bias(feat)
        _linear_sd(w, "convnext.head.3.", config["fc1_neurons"], feat, sd)
        _linear_sd(w, "convnext.head.5.", config["fc2_neurons"], config["fc1_neurons"], sd)
        _linear_sd(w, "convnext.head.8.", 1, config["fc2_neurons"], sd)
    elif name == "mm_MaxViT":
        arch = maxvit_arch(config.get("model_kind", "maxvit_tiny_rw_224.sw_in1k"))
        _maxvit_trunk_sd(w, "maxvit_backbone.", arch, sd, gamma_base)
        feat = arch["embed_dim"][-1]
        _bn_sd(w, "metadata_branch.0.", nmeta, sd)
        _linear_sd(w, "metadata_branch.1.", config["meta_fc1_neurons"], nmeta, sd)
        _linear_sd(w, "metadata_branch.4.", config["meta_fc2_neurons"], config["meta_fc1_neurons"], sd)
        _linear_sd(w, "combined_head.0.", config["comb_fc1_neurons"], feat + config["meta_fc2_neurons"], sd)
        _linear_sd(w, "combined_head.2.", config["comb_fc2_neurons"], config["comb_fc1_neurons"], sd)
        _linear_sd(w, "combined_head.5.", 1, config["comb_fc2_neurons"], sd)
    elif name == "MaxViT":
        arch = maxvit_arch(config.get("model_kind", "maxvit_tiny_rw_224.sw_in1k"))
        _maxvit_trunk_sd(w, "maxvit.", arch, sd, gamma_base)
        feat = arch["embed_dim"][-1]
        _linear_sd(w, "maxvit.head.1.", config["fc1_neurons"], feat, sd)
        _linear_sd(w, "maxvit.head.3.", config["fc2_neurons"], config["fc1_neurons"], sd)
        _linear_sd(w, "maxvit.head.6.", 1, config["fc2_neurons"], sd)
    elif name == "um_nn":
        _bn_sd(w, "network.0.", nmeta, sd)
        _linear_sd(w, "network.1.", config["meta_fc1_neurons"], nmeta, sd)
        _linear_sd(w, "network.4.", config["meta_fc2_neurons"], config["meta_fc1_neurons"], sd)
        _linear_sd(w, "network.6.", 1, config["meta_fc2_neurons"], sd)
    elif name == "frozen_fusion":
        img = make_state_dict(config["image_model_config"], seed + 101, perturb, gamma_base)
        met = make_state_dict(config["meta_model_config"], seed + 202, perturb, gamma_base)
        iname = config["image_model_config"]["model_name"]
        if iname not in ("ConvNeXt", "MaxViT") or config["meta_model_config"]["model_name"] != "um_nn":
            raise ValueError("synthetic frozen_fusion weights: ConvNeXt or MaxViT image branch + um_nn meta branch only")
        for k, v in img.items():          # head cut to [pool, LN2d, flatten] (architectures.py:309-313) / [pool] (:304-308)
            if not k.startswith(("convnext.head.3.", "convnext.head.5.", "convnext.head.8.", "maxvit.head.")):
                sd["image_branch." + k] = v
        for k, v in met.items():          # network[:-2] (architectures.py:299-303)
            if not k.startswith("network.6."):
                sd["meta_branch." + k] = v
        if iname == "ConvNeXt":
            feat = convnext_arch(config["image_model_config"].get("model_kind", "convnext_nano.d1h_in1k"))["dims"][-1]
        else:
            feat = maxvit_arch(config["image_model_config"].get("model_kind", "maxvit_tiny_rw_224.sw_in1k"))["embed_dim"][-1]
        comb_in = feat + config["meta_model_config"]["meta_fc2_neurons"]
        _linear_sd(w, "combined_head.0.", config["comb_fc1_neurons"], comb_in, sd)
        _linear_sd(w, "combined_head.2.", config["comb_fc2_neurons"], config["comb_fc1_neurons"], sd)
        _linear_sd(w, "combined_head.5.", 1, config["comb_fc2_neurons"], sd)
    else:
        raise ValueError(f"make_state_dict: unsupported model_name {name!r}")
    return sd


FINAL_LAYER = {"mm_ConvNeXt": "combined_head.5.", "frozen_fusion": "combined_head.5.",
               "ConvNeXt": "convnext.head.8.", "um_nn": "network.6.",
               "mm_MaxViT": "combined_head.5.", "MaxViT": "maxvit.head.6."}


def apply_calibration(sd: dict, config: dict, scale: float, shift: float) -> dict:
    """Re-centre/re-scale the final ``Linear(.,1)`` so logits become ``scale * (logit - shift)``.

    Random weights give logits that all share one sign, which makes
    "same labels at the 0.5 threshold" vacuous; parity tests calibrate with constants stored next to
    the golden vectors so logits straddle 0 with a spread far above the tolerance."""
    p = FINAL_LAYER[config["model_name"]]
    sd = dict(sd)
    sd[p + "weight"] = (sd[p + "weight"].astype(np.float64) * scale).astype(np.float32)
    sd[p + "bias"] = ((sd[p + "bias"].astype(np.float64) - shift) * scale).astype(np.float32)
    return sd


def to_torch(sd: dict):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def canonical_config(model_name: str = "mm_ConvNeXt", model_kind: str = "convnext_nano.d1h_in1k") -> dict:
    """Canonical bench/test config: the model keys
    `btsbot/architectures.py:107-163` reads, with `prod_config.json:49-51` sizes."""
    base = dict(
        model_name=model_name, model_kind=model_kind, pretrained=False,
        train_data_version="v12", metadata_cols=list(METADATA_COLS),
        meta_fc1_neurons=128, meta_fc2_neurons=128, meta_dropout=0.25,
        comb_fc1_neurons=128, comb_fc2_neurons=8, comb_dropout=0.2,
        fc1_neurons=128, fc2_neurons=8, dropout=0.2,
        batch_size=64, random_seed=2,
    )
    if model_name == "frozen_fusion":
        # image branch by model kind: ConvNeXt (BTSbot-convnext-*-metadata) or MaxViT (BTSbot-maxvit-tiny-*-metadata)
        base.update(
            image_model_dir="", meta_model_dir="", skip_load_state=True,
            image_model_config=canonical_config("MaxViT" if "maxvit" in model_kind.lower() else "ConvNeXt", model_kind),
            meta_model_config=canonical_config("um_nn", model_kind),
        )
    return base
