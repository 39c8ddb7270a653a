"""GPU: training step (forward with saved intermediates, BCE-with-logits, hand-written backward, fused AdamW)
against torch autograd on the CPU oracle.  Dropout probabilities are 0 for exact parity (the reference's dropout
masks come from torch's Philox stream; ours from a counter hash -- statistical, not bitwise, equivalence)."""
import numpy as np
import pytest
import torch

import btsbot_b200 as btsbot
from btsbot_b200 import synth
from btsbot_b200._autograd import BCEWithLogitsLoss, FusedAdamW
from cases import case_config

pytestmark = pytest.mark.gpu


def _nodrop(cfg):
    cfg = dict(cfg, meta_dropout=0.0, comb_dropout=0.0, dropout=0.0)
    for k in ("image_model_config", "meta_model_config"):
        if k in cfg:
            cfg[k] = _nodrop(cfg[k])
    return cfg


def _batch(n, start=300):
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(n, start=start).transpose(0, 3, 1, 2)))
    meta = torch.from_numpy(synth.make_metadata(n, start=start))
    lab = torch.from_numpy(synth.make_labels(n, start=start)).float().unsqueeze(1)
    return img, meta, lab


def _oracle_step(cfg, sd_np, img, meta, lab, pw, trainable=None):
    from oracle import convnext_oracle as O
    sd = {k: v.clone() for k, v in synth.to_torch(sd_np).items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and (trainable is None or k.startswith(trainable)):
            v.requires_grad_(True)
    logits = O.forward_train(sd, cfg, img, meta)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, lab, pos_weight=torch.tensor([pw]))
    loss.backward()
    return sd, logits.detach(), loss.detach()


def _call(model, cfg, img, meta):
    if cfg["model_name"] in ("mm_ConvNeXt", "frozen_fusion"):
        return model(image_input=img, metadata_input=meta)
    if cfg["model_name"] == "um_nn":
        return model(input_data=meta)
    return model(input_data=img)


@pytest.mark.parametrize("case", ["mm_pico", "mm_nano_LS", "img_pico", "um_nn", "ff_pico"])
def test_gradients_match_autograd(cuda_dev, case):
    cfg = _nodrop(case_config(case))
    sd_np = synth.make_state_dict(cfg, seed=11)
    B, pw = 6, 1.7
    img, meta, lab = _batch(B)
    trainable = "combined_head." if case == "ff_pico" else None
    ref_sd, ref_logits, ref_loss = _oracle_step(cfg, sd_np, img, meta, lab, pw, trainable)

    model = getattr(btsbot, cfg["model_name"])(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(cuda_dev).train()
    if case == "ff_pico":                                   # train.py:224-231
        for p in model.image_branch.parameters():
            p.requires_grad = False
        for p in model.meta_branch.parameters():
            p.requires_grad = False
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([pw])).to(cuda_dev)
    model.zero_grad()
    logits = _call(model, cfg, img.to(cuda_dev), meta.to(cuda_dev))
    loss = loss_fn(logits, lab.to(cuda_dev))
    loss.backward()
    torch.cuda.synchronize()
    assert (logits.detach().cpu() - ref_logits).abs().max() < 2e-4
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    worst = ("", 0.0)
    checked = 0
    for name, p in model.named_parameters():
        ref = ref_sd[name].grad
        if not p.requires_grad:
            assert p.grad is None
            continue
        assert p.grad is not None, name
        assert ref is not None, name
        scale = max(ref.abs().max().item(), 1e-6)
        err = (p.grad.cpu() - ref).abs().max().item() / scale
        checked += 1
        if err > worst[1]:
            worst = (name, err)
        assert err < 2e-3, (name, err, scale)
    print(f"[parity] {case} training step: loss {loss.item():.6f} (oracle {ref_loss.item():.6f}); "
          f"{checked} gradients, worst relative error {worst[1]:.2e} at {worst[0]}")
    # BatchNorm1d running statistics follow torch (momentum 0.1, unbiased variance)
    bn = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    if bn:
        m = meta.double()
        key = [k for k in sd_np if k.endswith("0.running_mean")][0]
        exp_mean = 0.9 * torch.from_numpy(sd_np[key]).double() + 0.1 * m.mean(0)
        assert (bn[0].running_mean.cpu().double() - exp_mean).abs().max() < 1e-3 * exp_mean.abs().max()
        assert int(bn[0].num_batches_tracked) == int(sd_np[key.replace("running_mean", "num_batches_tracked")]) + 1


def test_adamw_step_matches_torch(cuda_dev):
    cfg = _nodrop(case_config("mm_pico"))
    sd_np = synth.make_state_dict(cfg, seed=12)
    img, meta, lab = _batch(8, start=900)
    model = btsbot.mm_ConvNeXt(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(cuda_dev).train()
    opt = FusedAdamW(model.parameters(), lr=3e-3, betas=(0.9, 0.99))
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([2.0]))
    ref_sd, _, _ = _oracle_step(cfg, sd_np, img, meta, lab, 2.0)
    names = [k for k, v in ref_sd.items() if v.requires_grad]
    ref_opt = torch.optim.AdamW([ref_sd[k] for k in names], lr=3e-3, betas=(0.9, 0.99))
    losses = []
    for it in range(2):
        model.zero_grad()
        loss = loss_fn(_call(model, cfg, img.to(cuda_dev), meta.to(cuda_dev)), lab.to(cuda_dev))
        loss.backward()
        opt.step()
        losses.append(loss.item())
        ref_opt.step()
        if it == 0:                                          # second oracle step on the updated weights
            from oracle import convnext_oracle as O
            for k in names:
                ref_sd[k].grad = None
            lg = O.forward_train(ref_sd, cfg, img, meta)
            torch.nn.functional.binary_cross_entropy_with_logits(lg, lab, pos_weight=torch.tensor([2.0])).backward()
    torch.cuda.synchronize()
    # Adam's first steps are sign-like (|m|/sqrt(v) ~ 1): an element whose gradient is ~0 can legitimately move by
    # up to lr in either direction, so compare updates in the Frobenius norm and bound outliers by 2*lr per step.
    worst_fro, worst_abs = 0.0, 0.0
    got = dict(model.named_parameters())
    init = synth.to_torch(sd_np)
    for k in names:
        ref = ref_sd[k].detach()
        mine = got[k].detach().cpu()
        upd = (ref - init[k]).norm().item()
        worst_fro = max(worst_fro, (mine - ref).norm().item() / max(upd, 1e-12))
        worst_abs = max(worst_abs, (mine - ref).abs().max().item())
    print(f"[parity] two AdamW steps: losses {losses}, worst |update error|_F / |update|_F = {worst_fro:.2e}, "
          f"worst element error {worst_abs:.2e}")
    assert worst_fro < 2e-2 and worst_abs <= 4.1 * 3e-3
    assert losses[1] < losses[0]


def test_dropout_statistics_and_eval_mode(cuda_dev):
    from btsbot_b200 import _autograd as A
    x = torch.ones(1 << 20, device=cuda_dev)
    y, mask = A.dropout(x, 0.25, seed=123)
    keep = mask.float().mean().item()
    assert abs(keep - 0.75) < 5e-3
    assert torch.allclose(y[mask.bool()], torch.full((1,), 1 / 0.75, device=cuda_dev))
    y2, _ = A.dropout(x * 3, 0.25, mask=mask)
    assert torch.equal(y2 != 0, mask.bool())
    # eval mode / no_grad never touches the training path
    cfg = case_config("mm_pico")
    model = btsbot.mm_ConvNeXt(cfg).to(cuda_dev)
    img, meta, _ = _batch(4)
    model.train()
    with torch.no_grad():
        a = model(image_input=img.to(cuda_dev), metadata_input=meta.to(cuda_dev))
    model.eval()
    b = model(image_input=img.to(cuda_dev), metadata_input=meta.to(cuda_dev))
    assert torch.equal(a, b) and not a.requires_grad
