"""CPU: the MaxViT restatement (oracle/maxvit_oracle.py).  timm is not installable offline and the reference holds no
tests for this path, so the restatement is checked against (a) goldens produced by the reference's own
`architectures.py` MaxViT / mm_MaxViT wrappers executed verbatim on an independently written module twin
(oracle/timm_shim.ShimMaxViT: nn modules, torchvision partition ops, SDPA), (b) torchvision's MaxViT helpers for the
relative-position index and the window / grid partitions, (c) timm's published parameter count."""
import os

import numpy as np
import pytest
import torch

import btsbot_b200 as btsbot
from btsbot_b200 import synth
from oracle import maxvit_oracle as MO
from conftest import GOLDEN

CASES = {"mm_maxvit": "mm_MaxViT", "img_maxvit": "MaxViT", "ff_maxvit": "frozen_fusion"}
KIND = "maxvit_tiny_rw_224.sw_in1k"


@pytest.fixture(scope="module")
def golden_mv():
    return np.load(os.path.join(GOLDEN, "maxvit_logits.npz"))


def maxvit_batch(golden_mv, example_inputs):
    sel = golden_mv["example_idx"]
    nsyn = int(golden_mv["nsyn"])
    trip = np.concatenate([example_inputs["triplets"][sel], synth.make_triplets(nsyn, start=2000)])
    meta = np.concatenate([example_inputs["metadata"][sel], synth.make_metadata(nsyn, start=2000)])
    return np.ascontiguousarray(trip.transpose(0, 3, 1, 2)), meta


def maxvit_case(case, golden_mv, gain=None):
    cfg = synth.canonical_config(CASES[case], KIND)
    scale, shift = golden_mv[case + "_cal"]
    sd = synth.apply_calibration(synth.make_state_dict(cfg, seed=2), cfg, float(scale if gain is None else gain), float(shift))
    return cfg, sd


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_wrappers(case, golden_mv, example_inputs):
    img, meta = maxvit_batch(golden_mv, example_inputs)
    cfg, sd = maxvit_case(case, golden_mv)
    out = MO.forward(synth.to_torch(sd), cfg, torch.from_numpy(img), torch.from_numpy(meta)).numpy()
    ref = golden_mv[case]
    assert out.shape == ref.shape == (img.shape[0], 1)
    assert np.abs(out - ref).max() < 5e-5
    assert 0.3 <= (ref > 0).mean() <= 0.7


@pytest.mark.parametrize("case", list(CASES))
def test_state_dict_keys(case, golden_mv):
    cfg, sd = maxvit_case(case, golden_mv)
    model = getattr(btsbot, cfg["model_name"])(cfg)
    ours = set(model.state_dict().keys())
    assert ours == set(sd.keys()) == {str(k) for k in golden_mv[case + "_keys"]}
    model.load_state_dict(synth.to_torch(sd), strict=True)


def test_parameter_count_matches_timm_published():
    cfg = synth.canonical_config("mm_MaxViT", KIND)
    sd = synth.make_state_dict(cfg)
    trunk = sum(v.size for k, v in sd.items()
                if k.startswith("maxvit_backbone.") and "running_" not in k and "num_batches" not in k)
    # timm publishes 29.06 M for maxvit_tiny_rw_224 including the 1000-class head (512*1000 + 1000)
    assert trunk + 513_000 == 29_057_312


def test_rel_pos_index_and_partitions_match_torchvision():
    from torchvision.models import maxvit as tv
    assert torch.equal(MO.rel_pos_index(7), tv._get_relative_position_index(7, 7))
    x = torch.randn(2, 28, 28, 5)
    xc = x.permute(0, 3, 1, 2)
    wp, swap = tv.WindowPartition(), tv.SwapAxes(-2, -3)
    win = wp(xc, 7).reshape(-1, 7, 7, 5)                       # [B, nWin, 49, C]
    assert torch.equal(MO.window_partition(x, 7), win)
    grid = swap(wp(xc, 28 // 7)).reshape(-1, 7, 7, 5)
    assert torch.equal(MO.grid_partition(x, 7), grid)
    assert torch.equal(MO.window_reverse(MO.window_partition(x, 7), 7, 28, 28), x)
    assert torch.equal(MO.grid_reverse(MO.grid_partition(x, 7), 7, 28, 28), x)


def test_resize_only_when_needed():
    arch = MO.arch_of(KIND)
    x = torch.randn(1, 3, 224, 224)
    assert MO.resize(x, arch) is x
    assert MO.resize(torch.randn(1, 3, 63, 63), arch).shape == (1, 3, 224, 224)
