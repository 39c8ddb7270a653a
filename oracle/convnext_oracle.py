"""CPU fp32 restatement of the reference forward (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional, state-dict driven (timm / reference key names), plain ``torch.nn.functional``
on CPU tensors.  Follows:

* `btsbot/architectures.py:104-122`  ``ConvNeXt``      (pool -> LN2d -> flatten -> MLP head)
* `btsbot/architectures.py:125-171`  ``mm_ConvNeXt``   (trunk -> Flatten | pool+LN+flatten; BN1d-MLP; cat; head)
* `btsbot/architectures.py:277-293`  ``um_nn``
* `btsbot/architectures.py:296-372`  ``frozen_fusion`` (ConvNeXt head[0:3]; um_nn.network[:-2]; ReLU head)
* timm ``ConvNeXt`` (nano/pico, ``conv_mlp=True``), SURVEY.md Appendix A.1: patch stem 4x4/s4 + LN2d(eps 1e-6);
  stage = [LN2d + conv 2x2/s2] + blocks; block = dw7x7(p3) -> LN2d -> 1x1(4C) -> GELU(erf) -> 1x1(C) -> *gamma -> +x.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LN_EPS = 1e-6
BN_EPS = 1e-5

ARCHS = {
    "convnext_nano": dict(dims=(80, 160, 320, 640), depths=(2, 2, 8, 2)),
    "convnext_pico": dict(dims=(64, 128, 256, 512), depths=(2, 2, 6, 2)),
}


def arch_of(model_kind: str) -> dict:
    for k, v in ARCHS.items():
        if k in model_kind.lower():
            return v
    raise ValueError(model_kind)


def layernorm2d(x, w, b):
    # timm LayerNorm2d: permute -> F.layer_norm over C (biased variance) -> permute
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, LN_EPS).permute(0, 3, 1, 2)


def trunk_features(sd: dict, p: str, x: torch.Tensor, arch: dict, capture: dict | None = None) -> torch.Tensor:
    """timm ``forward_features``: [B,3,H,W] -> [B,C3,h,w] (NCHW)."""
    dims, depths = arch["dims"], arch["depths"]
    x = F.conv2d(x, sd[p + "stem.0.weight"], sd[p + "stem.0.bias"], stride=4)
    x = layernorm2d(x, sd[p + "stem.1.weight"], sd[p + "stem.1.bias"])
    if capture is not None:
        capture["stem"] = x
    for i, (c, d) in enumerate(zip(dims, depths)):
        if i > 0:
            q = f"{p}stages.{i}.downsample."
            x = layernorm2d(x, sd[q + "0.weight"], sd[q + "0.bias"])
            x = F.conv2d(x, sd[q + "1.weight"], sd[q + "1.bias"], stride=2)
            if capture is not None:
                capture[f"down{i}"] = x
        for j in range(d):
            q = f"{p}stages.{i}.blocks.{j}."
            s = x
            y = F.conv2d(x, sd[q + "conv_dw.weight"], sd[q + "conv_dw.bias"], padding=3, groups=c)
            y = layernorm2d(y, sd[q + "norm.weight"], sd[q + "norm.bias"])
            if capture is not None:
                capture[f"s{i}b{j}.dwln"] = y
            y = F.conv2d(y, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])
            y = F.gelu(y)
            y = F.conv2d(y, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
            x = y * sd[q + "gamma"].view(1, -1, 1, 1) + s
            if capture is not None:
                capture[f"s{i}b{j}"] = x
    return x


def pool_norm_flatten(sd, p, x):
    x = x.mean(dim=(2, 3), keepdim=True)            # SelectAdaptivePool2d('avg')
    x = layernorm2d(x, sd[p + "weight"], sd[p + "bias"])
    return x.flatten(1)


_TRAIN = False     # set by forward_train(): BatchNorm1d uses batch statistics (dropout must be 0 for parity)


def bn1d_eval(sd, p, x):
    if _TRAIN:
        return F.batch_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.1, BN_EPS)
    return (x - sd[p + "running_mean"]) / torch.sqrt(sd[p + "running_var"] + BN_EPS) * sd[p + "weight"] + sd[p + "bias"]


def lin(sd, p, x):
    return F.linear(x, sd[p + "weight"], sd[p + "bias"])


def metadata_branch(sd, p, m, act):
    # BN1d -> Linear -> act -> Dropout(eval: identity) -> Linear [-> act]
    m = bn1d_eval(sd, p + "0.", m)
    m = act(lin(sd, p + "1.", m))
    return lin(sd, p + "4.", m)


def head3(sd, p, idx, x, act):
    # Linear -> act -> Linear -> act -> Dropout -> Linear(.,1)
    x = act(lin(sd, f"{p}{idx[0]}.", x))
    x = act(lin(sd, f"{p}{idx[1]}.", x))
    return lin(sd, f"{p}{idx[2]}.", x)


def forward(sd: dict, config: dict, image_input=None, metadata_input=None, capture: dict | None = None):
    """Eval-mode logits ``[B,1]`` for ``config['model_name']`` (reference forward kwargs)."""
    with torch.no_grad():
        return _forward(sd, config, image_input, metadata_input, capture)


def forward_train(sd: dict, config: dict, image_input=None, metadata_input=None):
    """Training-mode forward WITH autograd (BatchNorm1d batch statistics; dropout layers are identity, so parity
    tests set the dropout probabilities to 0).  ``sd`` tensors that require grad receive gradients."""
    global _TRAIN
    _TRAIN = True
    try:
        return _forward(sd, config, image_input, metadata_input, None)
    finally:
        _TRAIN = False


def _forward(sd, config, image_input, metadata_input, capture):
    name = config["model_name"]
    if name == "mm_ConvNeXt":
        arch = arch_of(config.get("model_kind", "convnext_nano.d1h_in1k"))
        f = trunk_features(sd, "convnext_backbone.", image_input, arch, capture)
        if "LS" in config["train_data_version"]:
            f = pool_norm_flatten(sd, "convnext_backbone.head.1.", f)
        else:
            f = f.flatten(1)
        m = F.gelu(metadata_branch(sd, "metadata_branch.", metadata_input, F.gelu))
        if capture is not None:
            capture["features"], capture["meta"] = f, m
        return head3(sd, "combined_head.", (0, 2, 5), torch.cat((f, m), dim=1), F.gelu)
    if name == "ConvNeXt":
        arch = arch_of(config.get("model_kind", "convnext_nano.d1h_in1k"))
        f = trunk_features(sd, "convnext.", image_input, arch, capture)
        f = pool_norm_flatten(sd, "convnext.head.1.", f)
        if capture is not None:
            capture["features"] = f
        return head3(sd, "convnext.head.", (3, 5, 8), f, F.gelu)
    if name == "um_nn":
        m = F.relu(metadata_branch(sd, "network.", metadata_input, F.relu))
        return lin(sd, "network.6.", m)
    if name == "frozen_fusion":
        icfg, mcfg = config["image_model_config"], config["meta_model_config"]
        arch = arch_of(icfg.get("model_kind", "convnext_nano.d1h_in1k"))
        f = trunk_features(sd, "image_branch.convnext.", image_input, arch, capture)
        f = pool_norm_flatten(sd, "image_branch.convnext.head.1.", f)
        m = metadata_branch(sd, "meta_branch.network.", metadata_input, F.relu)   # pre-activation embedding
        if capture is not None:
            capture["features"], capture["meta"] = f, m
        return head3(sd, "combined_head.", (0, 2, 5), torch.cat((f, m), dim=1), F.relu)
    raise ValueError(name)
