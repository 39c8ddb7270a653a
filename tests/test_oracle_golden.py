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
