"""Dataset plumbing with the reference's names (`btsbot/utils.py:12-48,51-67`)."""
from __future__ import annotations

import json
import os
from datetime import datetime

import numpy as np
import torch
from torch.utils.data import Dataset as TorchDataset


class FlexibleDataset(TorchDataset):
    """``(image, meta, label)`` / ``(image, label)`` / ``(meta, label)`` tuples, utils.py:12-41."""

    def __init__(self, images=None, metadata=None, labels=None, transform=None):
        self.images, self.metadata, self.labels, self.transform = images, metadata, labels, transform
        self.need_triplets = images is not None
        self.need_metadata = metadata is not None

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, idx):
        label = self.labels[idx]
        image = meta = None
        if self.need_triplets:
            image = self.images[idx]
            if self.transform:
                image = self.transform(image)
        if self.need_metadata:
            meta = self.metadata[idx]
        if self.need_triplets and self.need_metadata:
            return image, meta, label
        if self.images is not None:
            return image, label
        if self.metadata is not None:
            return meta, label


class RandomRightAngleRotation(object):
    """Rotate by a random multiple of 90 degrees (utils.py:44-48).  A right-angle rotation is an exact index
    permutation, so ``torch.rot90`` reproduces ``transforms.functional.rotate`` bit for bit; the draw uses
    ``np.random.choice`` like the reference so seeded runs pick the same angles."""

    def __call__(self, img):
        degrees = np.random.choice([0, 90, 180, 270])
        return torch.rot90(img, int(degrees) // 90, dims=(-2, -1))


def make_report(config, report_path, run_data, val_summ):
    """Training report as JSON -- same signature and keys as `btsbot/utils.py:51-67`."""
    history = {k: np.array(v).tolist() for k, v in run_data.items() if k != "run_name"}
    report = {
        "Run time stamp": datetime.now().strftime("%Y%m%d_%H%M%S"),
        "Run name": run_data["run_name"],
        "Training history": history,
        "train_config": dict(config),
        "val_summary": dict(val_summ),
    }
    with open(os.path.join(report_path), "w") as f:
        json.dump(report, f, indent=4)
    return report
