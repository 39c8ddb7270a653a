"""torchrun script (2+ ranks, NCCL): 3 DDP training steps with per-rank batch b must reproduce 3 single-process steps
on the concatenated batch (image-only ConvNeXt: no BatchNorm, dropout 0 -> mean of per-rank means == global mean).
Also checks that gradient buckets were launched in backward order, i.e. overlapped with the remaining backward."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import btsbot_b200 as btsbot  # noqa: E402
from btsbot_b200 import synth  # noqa: E402
from btsbot_b200._autograd import BCEWithLogitsLoss, FusedAdamW  # noqa: E402
from btsbot_b200.parallel import DistributedDataParallel, shard_range  # noqa: E402
from cases import case_config  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = dict(case_config("img_pico"), dropout=0.0)
    sd = synth.to_torch(synth.make_state_dict(cfg, seed=21))
    per, steps = 8, 3
    n = per * world
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(n, start=70).transpose(0, 3, 1, 2))).to(dev)
    lab = torch.from_numpy(synth.make_labels(n, start=70)).float().unsqueeze(1).to(dev)
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([1.3]))

    def train(model, x, y):
        opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
        model.train()
        for _ in range(steps):
            model.zero_grad()
            loss_fn(model(input_data=x), y).backward()
            opt.step()
        return model

    ddp = btsbot.ConvNeXt(cfg)
    ddp.load_state_dict(sd, strict=True)
    ddp = DistributedDataParallel(ddp.to(dev), bucket_mb=1.0)
    lo, hi = shard_range(n, rank, world)
    train(ddp, img[lo:hi].contiguous(), lab[lo:hi].contiguous())
    order, early = ddp.sink.last_order, ddp.sink.last_early
    nb = len(ddp.sink.bounds)
    # every bucket is sent exactly once, (almost) all of them from inside the backward, roughly front to back
    assert sorted(order) == list(range(nb)) and nb > 4, order
    assert early >= nb - 1, (early, nb)
    assert max(abs(pos - b) for pos, b in enumerate(order)) <= 6, order

    ref = btsbot.ConvNeXt(cfg)
    ref.load_state_dict(sd, strict=True)
    ref = train(ref.to(dev), img, lab)
    worst = 0.0
    init = {k: v.to(dev) for k, v in sd.items()}
    for (k, a), (_, b) in zip(ddp.module.named_parameters(), ref.named_parameters()):
        upd = (b - init[k]).norm().item()
        worst = max(worst, (a - b).norm().item() / max(upd, 1e-12))
    # replicas stay in sync
    flat = torch.cat([p.detach().reshape(-1) for p in ddp.module.parameters()])
    other = flat.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(flat, other), "replicas diverged"
    if rank == 0:
        print(f"world {world}: worst |DDP - large batch|_F / |update|_F = {worst:.2e}; "
              f"{early}/{nb} buckets all-reduced from inside the backward")
        assert worst < 2e-2, worst
        print("DDP_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
