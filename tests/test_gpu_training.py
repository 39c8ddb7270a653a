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


# ---- mixed-precision (bf16 tensor-core GEMM) training step -----------------------------------------------------------
def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("M,N", [(1350, 80), (6, 512), (257, 320), (64, 64)])
def test_cast_dual_ops(cuda_dev, M, N):
    from btsbot_b200 import _autograd as A
    g = torch.Generator().manual_seed(M * 7 + N)
    x = (torch.randn(M, N, generator=g) * 1.5).to(cuda_dev)
    d = torch.randn(M, N, generator=g).to(cuda_dev)
    vec = torch.randn(N, generator=g).to(cuda_dev)
    gelu = torch.nn.functional.gelu
    xr = x.clone().requires_grad_(True)
    gelu(xr).backward(d)
    cases = [(0, None, None, x), (0, None, vec, x * vec), (1, None, None, gelu(x)), (2, d, None, xr.grad)]
    for op, x2, cv, want in cases:
        rm, t, cs = A.cast_dual(x, op=op, x2=x2, colvec=cv, want_colsum=True)
        torch.cuda.synchronize()
        ld = t.shape[1]
        assert ld % 8 == 0 and ld >= M and rm.shape == (M, N) and t.shape[0] == N
        # values: bf16 rounding of the fp32 result (erf / exp differ from torch's in the last fp32 bits -> 1 bf16 ulp)
        assert (rm.float() - want).abs().max() <= 2 ** -7 * want.abs().max() + 1e-6, op
        assert torch.equal(t[:, :M].t().contiguous(), rm), op       # the transposed copy holds the same bits
        assert (cs - want.sum(0)).abs().max() <= 1e-4 * max(1.0, want.abs().sum(0).max().item()), op
    # bf16 inputs (the 4C-wide tensors are written by the bf16-output GEMM epilogue)
    xb, db = x.bfloat16(), d.bfloat16()
    xr = xb.float().clone().requires_grad_(True)
    gelu(xr).backward(db.float())
    for op, x2, want in [(0, None, xb.float()), (1, None, gelu(xb.float())), (2, db, xr.grad)]:
        rm, t, cs = A.cast_dual(xb, op=op, x2=x2, want_colsum=True)
        torch.cuda.synchronize()
        assert (rm.float() - want).abs().max() <= 2 ** -7 * want.abs().max() + 1e-6, ("bf16 in", op)
        assert torch.equal(t[:, :M].t().contiguous(), rm), ("bf16 in", op)
        assert (cs - want.sum(0)).abs().max() <= 1e-4 * max(1.0, want.abs().sum(0).max().item()), ("bf16 in", op)


@pytest.mark.parametrize("M,N,K", [(1350, 320, 80), (300, 80, 320), (6, 2048, 512), (128, 640, 1280), (5000, 160, 640),
                                   (1000, 80, 48), (4097, 1280, 320)])
def test_tc_training_gemms_match_fp32_on_bf16_operands(cuda_dev, M, N, K):
    """forward/dgrad GEMM (bf16 operands -> fp32) and the split-K wgrad GEMM against torch fp32 matmul on the same
    bf16-rounded operands: only the fp32 summation order differs."""
    from btsbot_b200 import _autograd as A
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    dy = torch.randn(M, N, generator=g).to(cuda_dev)
    a16, a16t, _ = A.cast_dual(a)
    w16, w16t = A._w16(w)
    out = A.tc_gemm(a16, w16, bias)
    ref = _bf(a) @ _bf(w).t() + bias
    assert (out - ref).abs().max() <= 2e-4 * ref.abs().max(), "forward"
    out_nb = A.tc_gemm(a16, w16)
    assert (out_nb - (ref - bias)).abs().max() <= 2e-4 * ref.abs().max(), "forward without bias"
    dy16, dy16t, _ = A.cast_dual(dy)
    dx = A.tc_gemm(dy16, w16t)                       # [M,N] @ [N,K]
    ref_dx = _bf(dy) @ _bf(w)
    assert (dx - ref_dx).abs().max() <= 2e-4 * ref_dx.abs().max(), "dgrad"
    dw = A.tc_wgrad(dy16t, a16t, M)                  # [N,K] = dy^T a
    ref_dw = _bf(dy).t() @ _bf(a)
    assert (dw - ref_dw).abs().max() <= 2e-4 * ref_dw.abs().max() + 1e-5, "wgrad"
    dw2 = A.tc_wgrad_mn(dy16, a16)                   # the same from the row-major copies (MN-major UMMA operands)
    assert (dw2 - ref_dw).abs().max() <= 2e-4 * ref_dw.abs().max() + 1e-5, "wgrad (MN-major)"
    torch.cuda.synchronize()


@pytest.mark.parametrize("case", ["mm_pico", "mm_nano_LS", "img_pico"])
def test_bf16_training_step_close_to_fp32_autograd(cuda_dev, case):
    """precision="bf16": GEMMs on tcgen05 with bf16 operands.  Logits within the north-star bf16 bar (2e-2) of the fp32
    oracle; every gradient tensor within 6 % of its max-norm and > 0.995 cosine of torch autograd in fp32."""
    cfg = _nodrop(case_config(case))
    sd_np = synth.make_state_dict(cfg, seed=11)
    B, pw = 6, 1.7
    img, meta, lab = _batch(B)
    ref_sd, ref_logits, ref_loss = _oracle_step(cfg, sd_np, img, meta, lab, pw)
    model = getattr(btsbot, cfg["model_name"])(dict(cfg, precision="bf16"))
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(cuda_dev).train()
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([pw])).to(cuda_dev)
    model.zero_grad()
    from btsbot_b200 import _lib
    prof = _lib.KernelProfiler()
    _lib.profiler = prof
    try:
        logits = _call(model, cfg, img.to(cuda_dev), meta.to(cuda_dev))
        loss = loss_fn(logits, lab.to(cuda_dev))
        loss.backward()
    finally:
        _lib.profiler = None
    names = {r[0] for r in prof.records}
    assert {"t_gemm_tc", "t_gemm_tc16", "t_wgrad_tc", "t_cast_dual"} <= names, names      # the tensor-core path really ran
    torch.cuda.synchronize()
    assert (logits.detach().cpu() - ref_logits).abs().max() < 2e-2
    worst, worst_cos = ("", 0.0), ("", 1.0)
    for name, p in model.named_parameters():
        ref = ref_sd[name].grad
        assert p.grad is not None and ref is not None, name
        g = p.grad.cpu()
        assert torch.isfinite(g).all(), name
        scale = max(ref.abs().max().item(), 1e-6)
        err = (g - ref).abs().max().item() / scale
        if err > worst[1]:
            worst = (name, err)
        if ref.numel() >= 64 and ref.norm() > 1e-6:
            cos = float((g.flatten() @ ref.flatten()) / (g.norm() * ref.norm() + 1e-30))
            if cos < worst_cos[1]:
                worst_cos = (name, cos)
    print(f"[parity] {case} bf16 training step: loss {loss.item():.5f} (fp32 oracle {ref_loss.item():.5f}); worst "
          f"gradient error {worst[1]:.2e} of max-norm at {worst[0]}; worst cosine {worst_cos[1]:.5f} at {worst_cos[0]}")
    assert worst[1] < 6e-2, worst
    assert worst_cos[1] > 0.995, worst_cos


def test_bf16_training_loss_decreases(cuda_dev):
    cfg = _nodrop(case_config("mm_pico"))
    sd_np = synth.make_state_dict(cfg, seed=12)
    img, meta, lab = _batch(32, start=900)
    model = btsbot.mm_ConvNeXt(dict(cfg, precision="bf16"))
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(cuda_dev).train()
    opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.99))
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([1.0]))
    losses = []
    for _ in range(6):
        model.zero_grad()
        loss = loss_fn(model(image_input=img.to(cuda_dev), metadata_input=meta.to(cuda_dev)), lab.to(cuda_dev))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_graphed_train_step_matches_eager(cuda_dev):
    """CUDA-graph replay of the whole step == the same number of eager steps (dropout 0; split-K wgrad reductions are
    unordered, hence a tolerance), and the device-side counters move: AdamW bias correction is not frozen at capture."""
    from btsbot_b200._autograd import GraphedTrainStep
    cfg = dict(_nodrop(case_config("mm_pico")), precision="bf16")
    sd_np = synth.make_state_dict(cfg, seed=13)
    img, meta, lab = (t.to(cuda_dev) for t in _batch(16, start=500))
    img2, meta2, lab2 = (t.to(cuda_dev) for t in _batch(16, start=700))
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([1.0]))

    def make(capturable):
        m = btsbot.mm_ConvNeXt(cfg)
        m.load_state_dict(synth.to_torch(sd_np), strict=True)
        m = m.to(cuda_dev).train()
        return m, FusedAdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.99), capturable=capturable)

    ref, ropt = make(False)
    batches = [(img, meta, lab), (img, meta, lab), (img2, meta2, lab2), (img, meta, lab), (img2, meta2, lab2)]
    ref_losses = []
    for b in batches:
        ref.zero_grad()
        l = loss_fn(ref(image_input=b[0], metadata_input=b[1]), b[2])
        l.backward()
        ropt.step()
        ref_losses.append(l.item())
    g, gopt = make(True)
    stepper = GraphedTrainStep(g, gopt, loss_fn, example=batches[0], warmup=2)     # the 2 warm-up steps use batches[0]
    assert stepper.graph is not None and stepper.kernels_per_step > 100
    got_losses = [stepper(*b).item() for b in batches[2:]]
    torch.cuda.synchronize()
    # the capture pass itself is not executed: warm-up (2) + 3 replays = 5 forwards and 5 optimizer steps, like the reference
    assert int(gopt._step_dev.item()) == 5 and int(g._graph_counter.item()) == 5
    worst = 0.0
    for (n, a), (_, b) in zip(ref.named_parameters(), g.named_parameters()):
        upd = (a.detach() - synth.to_torch(sd_np)[n].to(cuda_dev)).norm().item()
        worst = max(worst, (a.detach() - b.detach()).norm().item() / max(upd, 1e-12))
    print(f"[parity] graphed vs eager training: losses {got_losses} vs {ref_losses[2:]}, worst |dp|_F / |update|_F = {worst:.2e}")
    # Adam's first steps are sign-like (|m| / sqrt(v) ~ 1), so the last-bit noise of the unordered split-K reductions is
    # amplified to a few percent of the update norm between ANY two runs, graphed or not (1.7e-2 ... 6.2e-2 observed)
    assert worst < 0.2
    assert all(abs(x - y) < 5e-3 for x, y in zip(got_losses, ref_losses[2:]))


@pytest.mark.parametrize("graphed", [False, True])
def test_eval_after_optimizer_step_sees_the_new_weights(cuda_dev, graphed):
    """FusedAdamW (and CUDA-graph replays of it) write parameters through raw pointers, which autograd's version counters
    do not see: the packed inference copy must still be rebuilt.  Image-only ConvNeXt has no BatchNorm buffer whose
    `num_batches_tracked` bump could mask a stale cache."""
    from btsbot_b200._autograd import GraphedTrainStep
    cfg = dict(_nodrop(case_config("img_pico")), precision="bf16")
    sd_np = synth.make_state_dict(cfg, seed=5)
    img, meta, lab = (t.to(cuda_dev) for t in _batch(16, start=900))
    model = btsbot.ConvNeXt(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(cuda_dev)
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([1.0]))
    opt = FusedAdamW(model.parameters(), lr=5e-3, betas=(0.9, 0.99), capturable=graphed)

    def score():
        model.eval()
        with torch.no_grad():
            out = model(input_data=img).clone()
        model.train()
        return out
    before = score()
    if graphed:
        stepper = GraphedTrainStep(model, opt, loss_fn, example=(img, None, lab), warmup=1)
        mid = score()
        stepper(img, None, lab)                                  # a replay: no Python-side optimizer call at all
        torch.cuda.synchronize()
        after = score()
        assert not torch.equal(mid, after)
    else:
        model.zero_grad()
        loss_fn(model(input_data=img), lab).backward()
        opt.step()
        after = score()
    assert not torch.equal(before, after)
    fresh = btsbot.ConvNeXt(cfg)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
    fresh = fresh.to(cuda_dev).eval()
    with torch.no_grad():
        assert torch.equal(fresh(input_data=img), after)
