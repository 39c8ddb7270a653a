"""Training entry points with the reference's names and config keys (`btsbot/train.py:57-60,75-566`).

    python -m btsbot_b200.train config.json            # 1 GPU
    torchrun --nproc-per-node 8 -m btsbot_b200.train config.json   # one process per GPU, NCCL gradient all-reduce

What changed underneath: the forward/backward/optimizer run on the hand-written kernels (fp32); the per-sample Python
``DataLoader`` pipeline is replaced by a device-resident loader whose batch gather + flips/rotations are one kernel;
``nn.DataParallel`` (train.py:238-240) is replaced by ``parallel.DistributedDataParallel`` (bucketed NCCL all-reduce
overlapped with the backward).  Not provided: wandb logging/sweeps, the matplotlib diagnostic figure and embedding
dumps (outside the hot path; the ``testing`` config key is accepted and ignored).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import architectures, val
from ._autograd import BCEWithLogitsLoss, FusedAdamW
from .parallel import DistributedDataParallel
from .utils import GpuBatchLoader, make_report

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

IMAGE_ONLY_MODELS = ["ConvNeXt", "MaxViT", "um_cnn"]
METADATA_ONLY_MODELS = ["um_nn"]
MULTIMODAL_MODELS = ["mm_ConvNeXt", "mm_MaxViT", "mm_cnn", "frozen_fusion"]


def classic_train(config_path):
    with open(config_path, "r") as f:
        config = json.load(f)
    run_training(config)


def _dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    return rank, (dist.get_world_size() if dist.is_initialized() else 1)


def perf_to_stdout(epoch, epochs, t0, batch, num_batches, loss, acc, flush_stdout=True):
    msg = (f"Epoch {epoch + 1}/{epochs} [{batch}/{num_batches}] {time.time() - t0:6.1f}s "
           f"loss {loss:.4f} acc {acc:.4f}")
    print("\r" + msg, end="" if flush_stdout else "\n", flush=True)


def train_epoch(dataloader, epoch, epochs, optimizer, loss_fn, model, need_triplets, need_metadata, verbose=True):
    """One epoch (train.py:481-566): zero_grad -> forward -> BCE -> backward -> step per batch; epoch loss/accuracy
    over all batches.  The per-step ``.item()`` host syncs of the reference are deferred to the end of the epoch."""
    t0 = time.time()
    num_batches = len(dataloader)
    all_logits, all_labels = [], []
    # opt-in (config "cuda_graph": true -> FusedAdamW(capturable=True), single process): batch 0 of the epoch runs eagerly,
    # then the step is captured once (the epoch's learning rate is baked into the graph) and replayed for the rest
    graphed = getattr(optimizer, "capturable", False) and not dist.is_initialized() and need_triplets and need_metadata
    stepper = None
    for i, items in enumerate(dataloader):
        if graphed and i >= 1:
            from ._autograd import GraphedTrainStep
            images, meta, labels = items
            labels = labels.unsqueeze(1).to(device, non_blocking=True).float()
            batch = (images.to(device, non_blocking=True), meta.to(device, non_blocking=True), labels)
            if stepper is None:                       # this batch is trained by the (eager) warm-up step of the capture
                stepper = GraphedTrainStep(model, optimizer, loss_fn, example=batch, warmup=1)
                all_logits.append(stepper.warmup_logits)
            else:
                stepper(*batch)
                all_logits.append(stepper.last_logits.clone())
            all_labels.append(labels)
            continue
        model.zero_grad()
        if need_triplets and need_metadata:
            images, meta, labels = items
            logits = model(image_input=images.to(device, non_blocking=True), metadata_input=meta.to(device, non_blocking=True))
        elif need_triplets:
            images, labels = items
            logits = model(input_data=images.to(device, non_blocking=True))
        else:
            meta, labels = items
            logits = model(input_data=meta.to(device, non_blocking=True))
        labels = labels.unsqueeze(1).to(device, non_blocking=True).float()
        loss = loss_fn(logits, labels)
        loss.backward()
        optimizer.step()
        all_logits.append(logits.detach())
        all_labels.append(labels)
        if verbose and (i + 1) % 50 == 0:
            acc = ((logits.detach() > 0) == (labels > 0.5)).float().mean()
            perf_to_stdout(epoch, epochs, t0, i + 1, num_batches, loss.item(), acc.item())
    logits, labels = torch.cat(all_logits, dim=0), torch.cat(all_labels, dim=0)
    with torch.no_grad():
        epoch_loss = loss_fn(logits, labels).item()
    epoch_accuracy = ((logits > 0) == (labels > 0.5)).float().mean().item()      # sigmoid(x) > 0.5  <=>  x > 0
    if verbose:
        perf_to_stdout(epoch, epochs, t0, num_batches, num_batches, epoch_loss, epoch_accuracy, flush_stdout=False)
    return epoch_loss, epoch_accuracy


def run_training(config, run_name: str = "", sweeping: bool = False):
    rank, world = _dist_setup()
    model_name = config["model_name"]
    epochs, batch_size = config["epochs"], config["batch_size"]
    learning_rate = float(config["learning_rate"])
    warmup_epochs = config.get("warmup_epochs", 0)
    beta1, beta2, patience = config["beta_1"], config["beta_2"], config["patience"]
    random_state = config["random_seed"]
    np.random.seed(random_state)
    torch.manual_seed(random_state)
    torch.cuda.manual_seed_all(random_state)

    need_triplets = model_name in IMAGE_ONLY_MODELS or model_name in MULTIMODAL_MODELS
    need_metadata = model_name in METADATA_ONLY_MODELS or model_name in MULTIMODAL_MODELS
    if not need_triplets and not need_metadata:
        raise ValueError(f"{model_name} not categorized as image-only/metadata-only/multimodal.")
    if need_metadata and config.get("metadata_cols", None) is None:
        raise ValueError("Metadata columns not found in config.")

    cand, images, metadata, labels = val.load_split(config, "train", need_triplets, need_metadata, drop_nan_triplets=True)
    num_bts, num_notbts = int((labels == 1).sum()), int((labels == 0).sum())
    if rank == 0:
        print(f"num_notbts: {num_notbts}\nnum_bts: {num_bts}")
    # One optimizer step still consumes `batch_size` alerts: the reference's DataParallel splits each batch over the GPUs
    # (train.py:238-240), so every rank takes batch_size // world of it and the averaged gradient is the gradient of the
    # whole batch -- same effective batch, learning rate and steps per epoch as the single-GPU run.
    if batch_size % world != 0:
        raise ValueError(f"batch_size {batch_size} is not divisible by the {world} ranks")
    dataloader = GpuBatchLoader(
        images, metadata, labels, batch_size=batch_size // world, shuffle=True, drop_last=True, device=device,
        h_flip=need_triplets and bool(config.get("data_aug_h_flip", True)),
        v_flip=need_triplets and bool(config.get("data_aug_v_flip", True)),
        rot=need_triplets and bool(config.get("data_aug_rot", True)),
        seed=random_state, rank=rank, world_size=world)

    bts_weight = torch.FloatTensor([num_notbts / num_bts]).to(device)
    loss_fn = BCEWithLogitsLoss(pos_weight=bts_weight)
    try:
        model_type = getattr(architectures, model_name)
    except AttributeError:
        raise ValueError(f"Could not find model of name {model_name}")
    model = model_type(config).to(device)
    if model_name == "frozen_fusion":
        print("Freezing image and metadata branches")
        for p in model.image_branch.parameters():
            p.requires_grad = False
        for p in model.meta_branch.parameters():
            p.requires_grad = False
    if world > 1 or config.get("cuda_graph", False):
        # the wrapper owns one flat gradient buffer (static addresses: what a replayed CUDA graph needs); with a single
        # process its all-reduce is a no-op
        model = DistributedDataParallel(model)

    optimizer = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=learning_rate, betas=(beta1, beta2),
                           capturable=bool(config.get("cuda_graph", False)) and world == 1)
    scheduler = torch.optim.lr_scheduler.SequentialLR(
        optimizer,
        schedulers=[torch.optim.lr_scheduler.LinearLR(optimizer, start_factor=0.01, total_iters=warmup_epochs),
                    torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=max(1, epochs - warmup_epochs),
                                                               eta_min=learning_rate * 0.01)],
        milestones=[warmup_epochs])

    run_name = "testing" if config.get("testing", False) or not run_name else run_name
    model_dir = os.path.join("models", f"{model_name}_{config['train_data_version']}_N{config.get('N_max', 100)}_"
                                       f"{device.type}", run_name)
    if rank == 0:
        os.makedirs(model_dir, exist_ok=True)
    # key names of the reference's run_data (train.py:293-299); "lr" is extra
    history = {"run_name": run_name, "train_loss": [], "train_accuracy": [], "val_loss": [], "val_accuracy": [], "lr": []}
    best_val_loss, epochs_since_improvement = float("inf"), 0
    net = model.module if isinstance(model, DistributedDataParallel) else model

    for epoch in range(epochs):
        model.train()
        loss, acc = train_epoch(dataloader, epoch, epochs, optimizer, loss_fn, model, need_triplets, need_metadata,
                                verbose=(rank == 0))
        stop = torch.zeros(1, device=device)
        if rank == 0:
            torch.save(net.state_dict(), os.path.join(model_dir, "latest_model.pth"))
            val_loss, val_acc, _, _ = val.run_val(config, model_dir, "latest_model.pth", bts_weight, need_triplets,
                                                  need_metadata)
            print(f"  val_loss {val_loss:.4f} val_acc {val_acc:.4f}")
            # train.py:334-336: the bar is the minimum over ALL previous epochs' val losses (including improvements of
            # less than 0.5 % that did not save a model), not the loss of the last saved model
            prev_best_val_loss = min([float("inf")] + history["val_loss"])
            history["train_loss"].append(loss); history["train_accuracy"].append(acc)
            history["val_loss"].append(val_loss); history["val_accuracy"].append(val_acc)
            history["lr"].append(optimizer.param_groups[0]["lr"])
            if 1.005 * val_loss < prev_best_val_loss:
                best_val_loss, epochs_since_improvement = val_loss, 0
                torch.save(net.state_dict(), os.path.join(model_dir, "best_model.pth"))
            else:
                epochs_since_improvement += 1
                if epochs_since_improvement >= patience:
                    print(f"Early stopping after {epoch + 1} epochs")
                    stop += 1
        scheduler.step()
        if world > 1:
            dist.broadcast(stop, src=0)
        if stop.item() > 0:
            break

    if rank == 0:
        summ = {"best_val_loss": best_val_loss}
        if config.get("use_test_split", False):
            tl, ta, _, _ = val.run_val(config, model_dir, "best_model.pth", bts_weight, need_triplets, need_metadata, "test")
            summ.update(test_loss=tl, test_accuracy=ta)
        make_report(config, os.path.join(model_dir, "report.json"), history, summ)
    return history


if __name__ == "__main__":
    if sys.argv[1] == "sweep":
        raise SystemExit("wandb sweeps are outside the B200 hot path; run `python -m btsbot_b200.train config.json`")
    classic_train(sys.argv[1])
    if dist.is_initialized():
        dist.destroy_process_group()
