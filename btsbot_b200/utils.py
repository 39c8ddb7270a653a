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


class GpuBatchLoader:
    """Device-resident replacement for ``DataLoader(FlexibleDataset(..., transform=Compose([...])))`` in the training
    loop (train.py:178-209).  The whole split lives in HBM; a batch is ONE fused gather+augmentation kernel
    (``btsb_augment_gather_f32``: flips and right-angle rotations are index permutations, so results are bit-exact
    w.r.t. the torchvision transforms) instead of per-sample Python in 6 worker processes.

    Yields the same tuples as :class:`FlexibleDataset` batches -- ``(images, metadata, labels)``, ``(images, labels)``
    or ``(metadata, labels)`` -- already on the device.  ``shuffle``/``drop_last`` follow the reference's DataLoader
    settings; under ``torch.distributed`` every rank iterates its own contiguous slice of the (shared-seed) permutation.
    """

    def __init__(self, images=None, metadata=None, labels=None, batch_size=64, shuffle=False, drop_last=False,
                 h_flip=False, v_flip=False, rot=False, device="cuda", seed=0, rank=0, world_size=1):
        dev = torch.device(device)
        self.images = None if images is None else images.to(dev, torch.float32).contiguous()
        self.metadata = None if metadata is None else metadata.to(dev, torch.float32).contiguous()
        self.labels = labels.to(dev)
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.h_flip, self.v_flip, self.rot = h_flip, v_flip, rot
        self.rank, self.world = rank, world_size
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        n = len(self.labels)
        self.per_rank = n // world_size if world_size > 1 else n
        self.dev = dev

    def __len__(self):
        if self.drop_last:
            return self.per_rank // self.batch_size
        return (self.per_rank + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        import ctypes as C
        from . import _lib as L
        n = len(self.labels)
        perm = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        perm = perm[self.rank * self.per_rank:(self.rank + 1) * self.per_rank].to(self.dev)
        for i in range(len(self)):
            idx = perm[i * self.batch_size:(i + 1) * self.batch_size].contiguous()
            b = idx.numel()
            out = []
            if self.images is not None:
                flags = None
                if self.h_flip or self.v_flip or self.rot:
                    hf = (torch.rand(b, generator=self.gen) < 0.5) if self.h_flip else torch.zeros(b, dtype=torch.bool)
                    vf = (torch.rand(b, generator=self.gen) < 0.5) if self.v_flip else torch.zeros(b, dtype=torch.bool)
                    k = torch.randint(0, 4, (b,), generator=self.gen) if self.rot else torch.zeros(b, dtype=torch.long)
                    flags = (hf.long() | (vf.long() << 1) | (k << 2)).to(torch.uint8).to(self.dev)
                s = self.images.shape[-1]
                x = torch.empty((b, 3, s, s), device=self.dev, dtype=torch.float32)
                L.check(L.lib().btsb_augment_gather_f32(
                    C.c_void_p(self.images.data_ptr()), C.c_void_p(idx.data_ptr()),
                    C.c_void_p(flags.data_ptr()) if flags is not None else None, b, s, C.c_void_p(x.data_ptr()),
                    L.stream_ptr()), "augment_gather")
                out.append(x)
            if self.metadata is not None:
                out.append(self.metadata[idx])
            out.append(self.labels[idx])
            yield tuple(out)
