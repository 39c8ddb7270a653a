"""Validation loop with the reference's entry point (`btsbot/val.py:31-170`): rebuild the model, load the
checkpoint, score a split, return ``(loss, accuracy, raw_preds, labels)``.  The forward, the sigmoid/threshold
epilogue and the BCE loss run on the B200 kernels; plotting (`diagnostic_fig`, val.py:173-682) is outside the hot
path and not provided."""
import os
import os.path as path
import sys

import numpy as np
import pandas as pd
import torch

from . import architectures
from ._autograd import BCEWithLogitsLoss
from .utils import GpuBatchLoader
from . import ops

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def data_dir(dataset_version: str) -> str:
    """Same lookup as val.py:39-46 (the author's scratch directory when it exists, else the working directory)."""
    if sys.platform != "darwin" and os.path.exists(f"/scratch/nrc5378/BTSbot_training_{dataset_version}/"):
        return f"/scratch/nrc5378/BTSbot_training_{dataset_version}/"
    return ""


def load_split(config, split, need_triplets, need_metadata, drop_nan_triplets=False):
    """CSV + npy loading shared by train/val (train.py:133-171, val.py:82-101).  Returns torch CPU tensors; the
    float64 HWC triplets are cast + transposed on the GPU by kernel K1."""
    from . import alert_utils
    ver, n_str = config["train_data_version"], f"_N{config.get('N_max', 100)}"
    base = data_dir(ver)
    cand = pd.read_csv(f"{base}data/{split}_cand_{ver}{n_str}.csv", index_col=None)
    images = None
    if need_triplets:
        fpath = f"{base}data/{split}_triplets_{ver}{n_str}.npy"
        if not path.exists(fpath):
            print(f"Triplets file not found for {split}: {fpath}")
            exit(1)
        trip = np.load(fpath)
        if drop_nan_triplets and np.any(np.isnan(trip)):
            bad = np.isnan(trip).any(axis=(1, 2, 3))
            trip = trip[~bad]
            cand = cand.loc[~bad].reset_index(drop=True)
            print(f"**** Null in triplets ****\nRemoved {int(bad.sum())} alert(s) from triplets and cand/labels.")
        if tuple(trip.shape[1:]) == (63, 63, 3):
            chunks = [alert_utils.triplets_to_model_input(trip[i:i + 16384]) for i in range(0, len(trip), 16384)]
        else:
            # "LS" data versions carry larger legacy-survey cutouts (the reason mm_ConvNeXt has its pool+norm head,
            # architectures.py:136-141); K1 is specialised for the 63x63 ZTF stamp, so any other size takes the
            # reference's own astype(float32) + transpose(0,3,1,2) (train.py:139-155) as two device copies per chunk
            if trip.ndim != 4 or trip.shape[3] != 3:
                raise ValueError(f"expected triplets of shape [N,H,W,3], got {trip.shape}")
            chunks = [torch.from_numpy(np.ascontiguousarray(trip[i:i + 4096])).to(device).to(torch.float32)
                      .permute(0, 3, 1, 2).contiguous() for i in range(0, len(trip), 4096)]
        images = torch.cat(chunks) if chunks else torch.empty((0, 3) + tuple(trip.shape[1:3]), device=device)
    labels = torch.tensor(cand["label"].values, dtype=torch.long)
    metadata = None
    if need_metadata:
        cols = config.get("metadata_cols")
        if cols is None:
            print("metadata_cols not found in config")
            exit(1)
        vals = cand[cols].values.astype(np.float32)
        if np.isnan(vals).any():
            if split == "train":
                raise ValueError("NaNs found in metadata columns")
            print(f"NaNs found in {split} metadata columns")
        metadata = torch.tensor(vals)
    return cand, images, metadata, labels


def score_loader(model, loader, need_triplets, need_metadata):
    """no_grad forward over a loader (val.py:128-157).  Returns (logits [N,1], labels [N,1]) on the device."""
    all_logits, all_labels = [], []
    with torch.no_grad():
        for items in loader:
            if need_triplets and need_metadata:
                images, meta, labels = items
                logits = model(image_input=images.to(device, non_blocking=True),
                               metadata_input=meta.to(device, non_blocking=True))
            elif need_triplets:
                images, labels = items
                logits = model(input_data=images.to(device, non_blocking=True))
            else:
                meta, labels = items
                logits = model(input_data=meta.to(device, non_blocking=True))
            all_logits.append(logits.detach())
            all_labels.append(labels.unsqueeze(1).to(device, non_blocking=True).float())
    return torch.cat(all_logits, dim=0), torch.cat(all_labels, dim=0)


def run_val(config, model_dir, model_filename, bts_weight, need_triplets, need_metadata, split="val"):
    try:
        model_type = getattr(architectures, config["model_name"])
    except AttributeError:
        print(f"Could not find model of name {config['model_name']}")
        exit(0)
    model = model_type(config).to(device)
    model.load_state_dict(torch.load(path.join(model_dir, model_filename), map_location="cpu"))
    model.eval()
    loss_fn = BCEWithLogitsLoss(pos_weight=bts_weight)

    _, images, metadata, labels = load_split(config, split, need_triplets, need_metadata)
    loader = GpuBatchLoader(images, metadata, labels, batch_size=config["batch_size"], shuffle=False, device=device)
    logits, labels_d = score_loader(model, loader, need_triplets, need_metadata)

    overall_loss = loss_fn(logits, labels_d).item()
    scores, _ = ops.score(logits)                                    # sigmoid on device (val.py:153)
    all_raw_preds = scores.squeeze().cpu().numpy()
    all_labels_np = labels_d.squeeze().cpu().numpy()
    overall_accuracy = np.sum((all_raw_preds > 0.5) == all_labels_np) / len(all_labels_np)
    return overall_loss, overall_accuracy, all_raw_preds, all_labels_np
