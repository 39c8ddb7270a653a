"""GPU: the reference-named entry points on a synthetic data directory: run_training ->
checkpoints + report, run_val, the fused gather+augmentation loader, and (with >= 2 GPUs) DDP == large-batch."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch

from btsbot_b200 import synth
from cases import case_config

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_split(root, split, ver, n_max, n, start):
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    trip = synth.make_triplets(n, start=start, dtype=np.float64)
    meta = synth.make_metadata(n, start=start)
    # learnable labels: brightness of the difference-image centre + one metadata column
    score = trip[:, 29:34, 29:34, 2].mean(axis=(1, 2)) * 40 + (meta[:, 5] - 17.86) * 0.5
    lab = (score > np.median(score)).astype(int)
    df = pd.DataFrame(meta, columns=synth.METADATA_COLS)
    df["label"] = lab
    df.to_csv(os.path.join(root, "data", f"{split}_cand_{ver}_N{n_max}.csv"), index=False)
    np.save(os.path.join(root, "data", f"{split}_triplets_{ver}_N{n_max}.npy"), trip)
    return lab


def test_augment_gather_matches_torchvision(cuda_dev):
    import torchvision.transforms.v2.functional as TF
    from btsbot_b200.utils import GpuBatchLoader
    n, s = 37, 63
    imgs = torch.randn(n, 3, s, s)
    labels = torch.arange(n)
    loader = GpuBatchLoader(imgs, None, labels, batch_size=16, shuffle=True, drop_last=False, h_flip=True, v_flip=True,
                            rot=True, device=cuda_dev, seed=7)
    # replay the loader's RNG stream on the host and apply the reference transforms (train.py:186-199 order)
    g = torch.Generator().manual_seed(7)
    perm = torch.randperm(n, generator=g)
    seen = 0
    for bi, (x, lab) in enumerate(loader):
        idx = perm[bi * 16:(bi + 1) * 16]
        b = len(idx)
        hf, vf = torch.rand(b, generator=g) < 0.5, torch.rand(b, generator=g) < 0.5
        k = torch.randint(0, 4, (b,), generator=g)
        assert torch.equal(lab.cpu(), labels[idx])
        for j in range(b):
            ref = imgs[idx[j]]
            if hf[j]:
                ref = TF.horizontal_flip(ref)
            if vf[j]:
                ref = TF.vertical_flip(ref)
            ref = TF.rotate(ref, int(k[j]) * 90)
            assert torch.equal(x[j].cpu(), ref), (bi, j)
        seen += b
    assert seen == n and len(loader) == 3


def test_run_training_and_run_val_on_synthetic_fixture(cuda_dev, tmp_path, monkeypatch):
    from btsbot_b200 import train, val
    ver, n_max = "v12", 100
    _write_split(tmp_path, "train", ver, n_max, 640, 0)
    lab_val = _write_split(tmp_path, "val", ver, n_max, 200, 5000)
    monkeypatch.chdir(tmp_path)
    cfg = dict(case_config("mm_pico"), train_data_version=ver, N_max=n_max, epochs=3, batch_size=64,
               learning_rate=2e-3, beta_1=0.9, beta_2=0.999, patience=5, random_seed=2, testing=True,
               data_aug_h_flip=True, data_aug_v_flip=True, data_aug_rot=True)
    hist = train.run_training(cfg)
    assert len(hist["train_loss"]) == 3 and all(np.isfinite(hist["train_loss"])) and all(np.isfinite(hist["val_loss"]))
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    mdir = os.path.join("models", f"mm_ConvNeXt_{ver}_N{n_max}_cuda", "testing")
    for f in ("latest_model.pth", "best_model.pth", "report.json"):
        assert os.path.isfile(os.path.join(mdir, f)), f
    rep = json.load(open(os.path.join(mdir, "report.json")))
    assert rep["train_config"]["model_name"] == "mm_ConvNeXt" and "Training history" in rep
    # the same run with the step replayed from a CUDA graph (mixed precision): finite, learning
    hist_g = train.run_training(dict(cfg, cuda_graph=True, precision="bf16", epochs=2))
    assert len(hist_g["train_loss"]) == 2 and all(np.isfinite(hist_g["train_loss"])) and all(np.isfinite(hist_g["val_loss"]))
    assert hist_g["train_loss"][-1] < hist_g["train_loss"][0] + 0.02
    loss, acc, raw_preds, labels = val.run_val(cfg, mdir, "best_model.pth", torch.tensor([1.0]), True, True)
    assert raw_preds.shape == (200,) and labels.shape == (200,) and np.array_equal(labels, lab_val.astype(np.float32))
    assert 0.0 <= acc <= 1.0 and np.isfinite(loss) and raw_preds.min() >= 0 and raw_preds.max() <= 1
    print(f"[train] 3 epochs on 640 synthetic alerts: loss {hist['train_loss']}, val_loss {hist['val_loss']}, val acc {acc:.3f}")
    # the checkpoint round-trips through the reference's key layout
    sd = torch.load(os.path.join(mdir, "best_model.pth"), map_location="cpu")
    assert "convnext_backbone.stages.3.blocks.1.mlp.fc2.weight" in sd and "metadata_branch.0.running_mean" in sd


def test_run_training_on_LS_sized_cutouts(cuda_dev, tmp_path, monkeypatch):
    """'LS' data versions carry larger cutouts (the pool + LayerNorm head of mm_ConvNeXt, architectures.py:136-141):
    load_split must pass them through like the reference's astype + transpose (train.py:139-155), and train / val run."""
    from btsbot_b200 import train, val
    ver, n_max, s = "v12LS", 100, 79
    rng = np.random.default_rng(5)
    for split, n in (("train", 128), ("val", 64)):
        os.makedirs(os.path.join(tmp_path, "data"), exist_ok=True)
        trip = rng.standard_normal((n, s, s, 3)) * 0.02
        meta = synth.make_metadata(n, start=0 if split == "train" else 900)
        lab = (trip[:, s // 2 - 2:s // 2 + 3, s // 2 - 2:s // 2 + 3, 2].mean(axis=(1, 2)) > 0).astype(int)
        df = pd.DataFrame(meta, columns=synth.METADATA_COLS)
        df["label"] = lab
        df.to_csv(os.path.join(tmp_path, "data", f"{split}_cand_{ver}_N{n_max}.csv"), index=False)
        np.save(os.path.join(tmp_path, "data", f"{split}_triplets_{ver}_N{n_max}.npy"), trip)
    monkeypatch.chdir(tmp_path)
    cfg = dict(case_config("mm_pico"), train_data_version=ver, N_max=n_max, epochs=1, batch_size=32,
               learning_rate=1e-3, beta_1=0.9, beta_2=0.999, patience=5, random_seed=2, testing=True)
    _, images, _, _ = val.load_split(cfg, "val", True, True)
    assert tuple(images.shape) == (64, 3, s, s) and images.dtype == torch.float32 and images.is_cuda
    ref = np.load(os.path.join(tmp_path, "data", f"val_triplets_{ver}_N{n_max}.npy")).astype(np.float32).transpose(0, 3, 1, 2)
    assert np.array_equal(images.cpu().numpy(), ref)                       # exact: cast + layout only
    hist = train.run_training(cfg)
    assert len(hist["train_loss"]) == 1 and np.isfinite(hist["train_loss"][0]) and np.isfinite(hist["val_loss"][0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_two_gpus_equals_large_batch():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "ddp_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DDP_CHECK_OK" in r.stdout
