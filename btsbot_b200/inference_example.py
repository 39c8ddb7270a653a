#!/usr/bin/env python3
"""`btsbot/inference_example.py` on the B200 kernels: score the shipped example alerts with a model directory.

    python -m btsbot_b200.inference_example --architecture convnext --multi_modal --pretrain randinit \
        [--data-dir btsbot/example_data]

Model files are looked up in ``models/BTSbot-<arch>-<pretrain>[-metadata]/`` exactly like the reference
(`from_HF.py:37-40`); offline there is no download."""
import argparse
import os

import numpy as np
import pandas as pd
import torch

from . import alert_utils, load_HF_model
from .synth import METADATA_COLS


def parse_args():
    p = argparse.ArgumentParser(description="Use a BTSbot model directory on the B200 kernels")
    p.add_argument("--architecture", type=str, required=True, choices=["convnext", "maxvit"])
    p.add_argument("--pretrain", type=str, default="galaxyzoo", choices=["imagenet", "galaxyzoo", "randinit"])
    p.add_argument("--multi_modal", action="store_true")
    p.add_argument("--data-dir", type=str, default="example_data")
    a = p.parse_args()
    return a.architecture, a.multi_modal, a.pretrain, a.data_dir


def run_inference(model, multi_modal, data_dir="example_data"):
    cand = pd.read_csv(os.path.join(data_dir, "usage_candidates.csv"), index_col=None)
    labels = cand["label"].values
    dev = next(model.parameters()).device
    model = model.eval()
    triplets = np.load(os.path.join(data_dir, "usage_triplets.npy"), mmap_mode="r")[:64]     # one batch of <= 64
    with torch.no_grad():
        x = alert_utils.triplets_to_model_input(np.asarray(triplets))            # astype(float32) + NHWC->NCHW (K1)
        if multi_modal:
            meta = torch.tensor(cand[METADATA_COLS].values[:64].astype(np.float32)).to(dev)
            logits = model(image_input=x, metadata_input=meta)
        else:
            logits = model(input_data=x)
        raw_preds = torch.sigmoid(logits).round().squeeze().cpu().numpy().astype(int)
    print(raw_preds)
    print(labels[:64])
    return raw_preds


if __name__ == "__main__":
    architecture, multi_modal, pretrain, data_dir = parse_args()
    run_inference(load_HF_model(architecture, multi_modal, pretrain), multi_modal, data_dir)
