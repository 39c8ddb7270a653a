"""GPU: MaxViT path (BASELINE config 4) -- every kernel between the GEMMs against plain torch on the CPU, then the
whole models against the CPU oracle and the goldens made by the reference's own wrappers (fp32 1e-4, bf16 2e-2)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import btsbot_b200 as btsbot
from btsbot_b200 import _lib as L, synth
from oracle import maxvit_oracle as MO
from test_oracle_maxvit import CASES, KIND, golden_mv, maxvit_batch, maxvit_case  # noqa: F401

pytestmark = pytest.mark.gpu

DT = {"fp32": (L.F32, torch.float32, 2e-5), "bf16": (L.BF16, torch.bfloat16, 1.2e-2)}


def p(t):
    return C.c_void_p(t.data_ptr())


def rows_to_nchw(rows, B, H, W):
    return rows.float().cpu().view(B, H, W, -1).permute(0, 3, 1, 2)


def nchw_to_rows(x, dt, dev):
    B, Cc, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, Cc).contiguous().to(dt).to(dev)


def relerr(got, ref):
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("hin", [63, 224, 30])
def test_stem1_resize_conv_bn_silu(cuda_dev, precision, hin):
    code, dt, tol = DT[precision]
    g = torch.Generator().manual_seed(1)
    B, C1, S = 3, 32, 224
    x = torch.randn(B, 3, hin, hin, generator=g) * 0.02 + 0.016
    w = torch.randn(C1, 3, 3, 3, generator=g) / 27 ** 0.5
    scale, shift = torch.rand(C1, generator=g) * 50 + 50, torch.randn(C1, generator=g) * 0.3
    xr = F.interpolate(x, size=(S, S), mode="bilinear", align_corners=False) if hin != S else x
    ref = F.silu(F.conv2d(xr, w, None, stride=2, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    wk = (w * scale.view(-1, 1, 1, 1)).reshape(C1, 27).t().contiguous().to(cuda_dev)
    out = torch.empty((B * 112 * 112, C1), device=cuda_dev, dtype=dt)
    xd, sd_ = x.to(cuda_dev), shift.to(cuda_dev)        # keep device copies alive across the asynchronous launch
    L.check(L.lib().btsb_maxvit_stem1_fwd(p(xd), B, hin, hin, S, p(wk), p(sd_), C1, p(out), code, L.stream_ptr()), "stem1")
    torch.cuda.synchronize()
    assert relerr(rows_to_nchw(out, B, 112, 112), ref) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_im2col3_and_avgpool2(cuda_dev, precision):
    code, dt, _ = DT[precision]
    g = torch.Generator().manual_seed(2)
    B, Cc, H, W = 2, 32, 14, 10
    x = torch.randn(B, Cc, H, W, generator=g).to(dt).float()
    rows = nchw_to_rows(x, dt, cuda_dev)
    col = torch.empty((B * H * W, 9 * Cc), device=cuda_dev, dtype=dt)
    L.check(L.lib().btsb_maxvit_im2col3_fwd(p(rows), p(col), B, H, W, Cc, code, L.stream_ptr()), "im2col3")
    # F.unfold orders columns (c, ky, kx); ours is (ky, kx, c)
    ref = F.unfold(x, 3, padding=1).view(B, Cc, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * Cc)
    assert torch.equal(col.float().cpu(), ref)
    pool = torch.empty((B * (H // 2) * (W // 2), Cc), device=cuda_dev, dtype=dt)
    L.check(L.lib().btsb_maxvit_avgpool2_fwd(p(rows), p(pool), B, H, W, Cc, code, L.stream_ptr()), "avgpool2")
    ref = F.avg_pool2d(x, 2)
    assert relerr(rows_to_nchw(pool, B, H // 2, W // 2), ref) < (1e-6 if precision == "fp32" else 8e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("stride,H,Cc", [(1, 14, 256), (2, 28, 128), (2, 14, 320), (1, 7, 64)])
def test_dw3_bn_silu_and_se_pool(cuda_dev, precision, stride, H, Cc):
    code, dt, tol = DT[precision]
    g = torch.Generator().manual_seed(3)
    B, W = 3, H
    x = torch.randn(B, Cc, H, W, generator=g).to(dt).float()
    w = torch.randn(Cc, 1, 3, 3, generator=g) / 3
    shift = torch.randn(Cc, generator=g) * 0.2
    ref = F.silu(F.conv2d(x, w, None, stride=stride, padding=1, groups=Cc) + shift.view(1, -1, 1, 1))
    Ho = ref.shape[2]
    out = torch.empty((B * Ho * Ho, Cc), device=cuda_dev, dtype=dt)
    pooled = torch.empty((B, Cc), device=cuda_dev, dtype=torch.float32)
    wk = w.reshape(Cc, 9).t().contiguous().to(cuda_dev)
    xd, sd_ = nchw_to_rows(x, dt, cuda_dev), shift.to(cuda_dev)
    L.check(L.lib().btsb_maxvit_dw3_fwd(p(xd), B, H, W, Cc, stride, p(wk), p(sd_), p(out), p(pooled), code,
                                        L.stream_ptr()), "dw3")
    torch.cuda.synchronize()
    assert relerr(rows_to_nchw(out, B, Ho, Ho), ref) < tol
    assert relerr(pooled.cpu(), ref.mean(dim=(2, 3))) < (1e-5 if precision == "fp32" else 2e-3)


def test_se_gate_and_scale(cuda_dev):
    g = torch.Generator().manual_seed(4)
    B, Cc, R, HW = 5, 256, 16, 9
    pooled = torch.randn(B, Cc, generator=g)
    w1, b1 = torch.randn(R, Cc, generator=g) / 16, torch.randn(R, generator=g) * 0.1
    w2, b2 = torch.randn(Cc, R, generator=g) / 4, torch.randn(Cc, generator=g) * 0.1
    ref = torch.sigmoid(F.linear(F.silu(F.linear(pooled, w1, b1)), w2, b2))
    gate = torch.empty((B, Cc), device=cuda_dev)
    dv = [t.contiguous().to(cuda_dev) for t in (pooled, w1, b1, w2.t(), b2)]
    L.check(L.lib().btsb_maxvit_se_fwd(p(dv[0]), B, Cc, R, p(dv[1]), p(dv[2]), p(dv[3]), p(dv[4]), p(gate),
                                       L.stream_ptr()), "se")
    assert relerr(gate.cpu(), ref) < 1e-5
    for precision in ("fp32", "bf16"):
        code, dt, _ = DT[precision]
        x = torch.randn(B * HW, Cc, generator=g).to(dt)
        xd = x.to(cuda_dev)
        L.check(L.lib().btsb_maxvit_scale_fwd(p(xd), p(gate), B, HW, Cc, code, L.stream_ptr()), "scale")
        want = (x.float().view(B, HW, Cc) * ref.view(B, 1, Cc)).view(B * HW, Cc)
        assert relerr(xd.float().cpu(), want) < (1e-5 if precision == "fp32" else 8e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("Cc", [64, 128, 256, 512])
def test_layernorm_rows_and_lnpool(cuda_dev, precision, Cc):
    code, dt, tol = DT[precision]
    g = torch.Generator().manual_seed(5)
    B, HW = 3, 49
    x = (torch.randn(B * HW, Cc, generator=g) * 2 + 0.5).to(dt)
    w, b = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.1
    ref = F.layer_norm(x.float(), (Cc,), w, b, 1e-6)
    out = torch.empty_like(x, device=cuda_dev)
    xd, wd, bd = x.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev)
    L.check(L.lib().btsb_layernorm_rows_fwd(p(xd), p(wd), p(bd), p(out), B * HW, Cc, code, L.stream_ptr()), "ln")
    assert relerr(out.float().cpu(), ref) < tol
    feat = torch.empty((B, Cc), device=cuda_dev)
    L.check(L.lib().btsb_maxvit_lnpool_fwd(p(xd), p(wd), p(bd), p(feat), B, HW, Cc, code, L.stream_ptr()), "lnpool")
    assert relerr(feat.cpu(), ref.view(B, HW, Cc).mean(dim=1)) < 2e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["block", "grid"])
@pytest.mark.parametrize("H,Cc", [(14, 64), (28, 128), (7, 512)])
def test_window_and_grid_attention(cuda_dev, precision, kind, H, Cc):
    """softmax(q k^T/sqrt(32) + rel-pos bias) v over 7x7 windows / the 7x7 dilated grid, partition and reverse included."""
    code, dt, tol = DT[precision]
    g = torch.Generator().manual_seed(6)
    B, W, heads = 2, H, Cc // 32
    qkv = torch.randn(B, H, W, 3 * Cc, generator=g).to(dt)
    table = torch.randn(169, heads, generator=g) * 0.5
    part = MO.window_partition if kind == "block" else MO.grid_partition
    rev = MO.window_reverse if kind == "block" else MO.grid_reverse
    t = part(qkv.float(), 7).reshape(-1, 49, heads, 96).transpose(1, 2)
    q, k, v = t.chunk(3, dim=3)
    bias = table[MO.rel_pos_index(7).view(-1)].view(49, 49, heads).permute(2, 0, 1).unsqueeze(0)
    attn = ((q * 32 ** -0.5) @ k.transpose(-2, -1) + bias).softmax(dim=-1)
    ref = rev((attn @ v).transpose(1, 2).reshape(-1, 7, 7, Cc), 7, H, W).reshape(B * H * W, Cc)
    out = torch.empty((B * H * W, Cc), device=cuda_dev, dtype=dt)
    qd, td = qkv.view(-1, 3 * Cc).to(cuda_dev), table.to(cuda_dev)
    L.check(L.lib().btsb_maxvit_attn_fwd(p(qd), p(out), B, H, W, Cc, int(kind == "grid"), p(td), code, L.stream_ptr()), "attn")
    torch.cuda.synchronize()
    assert relerr(out.float().cpu(), ref) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gemm_silu_epilogue(cuda_dev, precision):
    code, dt, tol = DT[precision]
    g = torch.Generator().manual_seed(7)
    M, N, K = 777, 256, 64
    a, w = torch.randn(M, K, generator=g).to(dt), (torch.randn(N, K, generator=g) / 8).to(dt)
    b = torch.randn(N, generator=g) * 0.1
    out = torch.empty((M, N), device=cuda_dev, dtype=dt)
    ad, wd, bd = a.to(cuda_dev), w.to(cuda_dev), b.to(cuda_dev)
    L.check(L.lib().btsb_gemm_fwd(p(ad), p(wd), p(bd), None, None, p(out), M, N, K, code, L.EPI_BIAS_SILU, L.stream_ptr()),
            "gemm silu")
    ref = F.silu(a.float() @ w.float().t() + b)
    assert relerr(out.float().cpu(), ref) < tol


# ---------------------------------------------------------------------------------------------------------
# whole models
# ---------------------------------------------------------------------------------------------------------
def _build(case, golden_mv, dev, precision, gain=None):
    cfg, sd = maxvit_case(case, golden_mv, gain)
    cfg = dict(cfg, precision=precision)
    model = getattr(btsbot, cfg["model_name"])(cfg)
    model.load_state_dict(synth.to_torch(sd), strict=True)
    return cfg, sd, model.to(dev).eval()


def _call(model, cfg, img, meta):
    with torch.no_grad():
        if cfg["model_name"] in ("mm_MaxViT", "frozen_fusion"):
            return model(image_input=img, metadata_input=meta)
        return model(input_data=img)


@pytest.mark.parametrize("case", list(CASES))
def test_fp32_logits_match_reference(cuda_dev, golden_mv, example_inputs, case):
    img, meta = maxvit_batch(golden_mv, example_inputs)
    cfg, sd, model = _build(case, golden_mv, cuda_dev, "fp32")
    got = _call(model, cfg, torch.from_numpy(img).to(cuda_dev), torch.from_numpy(meta).to(cuda_dev)).cpu().numpy()
    ref = golden_mv[case]                                       # architectures.py wrappers executed verbatim
    orc = MO.forward(synth.to_torch(sd), cfg, torch.from_numpy(img), torch.from_numpy(meta)).numpy()
    e_ref, e_orc = np.abs(got - ref).max(), np.abs(got - orc).max()
    sure = np.abs(orc) > 1e-4
    print(f"[parity] {case} fp32: max|logit-ref|={e_ref:.3e} max|logit-oracle|={e_orc:.3e} spread {np.ptp(orc):.3f} "
          f"labels compared {int(sure.sum())}/{sure.size}")
    assert e_orc < 1e-4 and e_ref < 1.5e-4
    assert np.array_equal((got > 0)[sure], (orc > 0)[sure])


@pytest.mark.parametrize("case", list(CASES))
def test_bf16_logits_match_reference(cuda_dev, golden_mv, example_inputs, case):
    """2e-2 abs at gain 1 (the north star's bar); on the gain-10 goldens the bar scales with the gain."""
    img, meta = maxvit_batch(golden_mv, example_inputs)
    ti, tm = torch.from_numpy(img), torch.from_numpy(meta)
    cfg, sd, model = _build(case, golden_mv, cuda_dev, "bf16", gain=1.0)
    got = _call(model, cfg, ti.to(cuda_dev), tm.to(cuda_dev)).cpu().numpy()
    orc = MO.forward(synth.to_torch(sd), cfg, ti, tm).numpy()
    err1 = np.abs(got - orc).max()
    sure = np.abs(orc) > 2e-2
    assert err1 < 2e-2
    assert np.array_equal((got > 0)[sure], (orc > 0)[sure])
    gain = float(golden_mv[case + "_cal"][0])
    cfg, sd, model = _build(case, golden_mv, cuda_dev, "bf16")
    got2 = _call(model, cfg, ti.to(cuda_dev), tm.to(cuda_dev)).cpu().numpy()
    err2 = np.abs(got2 - golden_mv[case]).max()
    print(f"[parity] {case} bf16: gain-1 max|err|={err1:.3e} ({int(sure.sum())}/{sure.size} labels compared); "
          f"gain-{gain:.0f} max|err|={err2:.3e} (bar {2e-2 * gain:.2e})")
    assert err2 < 2e-2 * max(1.0, gain)
    # every alert takes part: labels may only flip where the reference logit is smaller than the measured error, and the
    # logits must follow the reference's ordering (the image-only model's spread is inside the tolerance band)
    ref2 = golden_mv[case]
    flips = (got2 > 0) != (ref2 > 0)
    corr = float(np.corrcoef(got2[:, 0], ref2[:, 0])[0, 1])
    print(f"[parity] {case} bf16: {int(flips.sum())}/{flips.size} labels differ over all alerts, corr {corr:.5f}")
    assert not flips.any() or float(np.abs(ref2[flips]).max()) <= err2
    assert flips.mean() <= 0.15 and corr > 0.99


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_trunk_intermediates(cuda_dev, precision):
    """stem, every MBConv output, every window-attention and grid-attention block output, pooled features."""
    from btsbot_b200 import _engine
    cfg = dict(synth.canonical_config("mm_MaxViT", KIND), precision=precision)
    sd = synth.to_torch(synth.make_state_dict(cfg, seed=5))
    B = 3
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(B, start=40).transpose(0, 3, 1, 2)))
    meta = torch.from_numpy(synth.make_metadata(B, start=40))
    cap_o, cap_g = {}, {}
    MO.forward(sd, cfg, img, meta, capture=cap_o)
    scorer = _engine.Scorer(cfg, {k: v.to(cuda_dev) for k, v in sd.items()}, precision)
    scorer(image_input=img.to(cuda_dev), metadata_input=meta.to(cuda_dev), capture=cap_g)
    torch.cuda.synchronize()
    worst, n = 0.0, 0
    for name, ref in cap_o.items():
        if name == "meta":
            continue
        if name == "features":
            got = cap_g[name].float().cpu()
        else:
            rows, h, w = cap_g[name]
            got = rows_to_nchw(rows, B, h, w)
        assert got.shape == ref.shape, name
        rel = relerr(got, ref)
        worst, n = max(worst, rel), n + 1
        assert rel < (3e-5 if precision == "fp32" else 4e-2), (name, rel)
    print(f"[parity] maxvit intermediates {precision}: worst relative error {worst:.3e} over {n} tensors")
    assert n == 2 + 3 * 11 + 1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_chunk_and_shard_invariance(cuda_dev, golden_mv, precision):
    """The trunk walks the batch in chunks; per-alert results must not depend on chunking or on index-range sharding."""
    from btsbot_b200 import _maxvit
    cfg, sd, model = _build("mm_maxvit", golden_mv, cuda_dev, precision)
    n = 11
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(n, start=0).transpose(0, 3, 1, 2))).to(cuda_dev)
    meta = torch.from_numpy(synth.make_metadata(n, start=0)).to(cuda_dev)
    whole = _call(model, cfg, img, meta)
    old = dict(_maxvit.CHUNK)
    try:
        _maxvit.CHUNK.update(fp32=4, bf16=4)
        chunked = _call(model, cfg, img, meta)
    finally:
        _maxvit.CHUNK.update(old)
    parts = torch.cat([_call(model, cfg, img[a:b], meta[a:b]) for a, b in ((0, 1), (1, 6), (6, n))])
    assert torch.equal(whole, chunked) and torch.equal(whole, parts)
    assert torch.isfinite(whole).all()


def test_error_behaviour(cuda_dev, golden_mv):
    cfg, sd, model = _build("mm_maxvit", golden_mv, cuda_dev, "bf16")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(image_input=torch.zeros(1, 3, 63, 63), metadata_input=torch.zeros(1, 25))
    with pytest.raises(ValueError):
        model(image_input=torch.zeros(1, 3, 63, 63, device=cuda_dev), metadata_input=torch.zeros(1, 24, device=cuda_dev))
    assert model(image_input=torch.zeros(0, 3, 63, 63, device=cuda_dev),
                 metadata_input=torch.zeros(0, 25, device=cuda_dev)).shape == (0, 1)


def test_load_HF_model_maxvit_metadata_checkpoint(cuda_dev, golden_mv, example_inputs, tmp_path, monkeypatch):
    """`load_HF_model("maxvit", True, ...)`: the published BTSbot-maxvit-tiny-*-metadata models are frozen_fusion over a
    MaxViT image branch with its head cut to [global_pool] (to_HF.py:143-177, architectures.py:304-308)."""
    import json
    cfg, sd = maxvit_case("ff_maxvit", golden_mv)
    assert cfg["image_model_config"]["model_name"] == "MaxViT"
    mdir = tmp_path / "models" / "BTSbot-maxvit-tiny-randinit-metadata"
    mdir.mkdir(parents=True)
    (mdir / "train_config.json").write_text(json.dumps(cfg))
    torch.save(synth.to_torch(sd), mdir / "pytorch_model.bin")
    monkeypatch.chdir(tmp_path)
    model = btsbot.load_HF_model("maxvit", True, "randinit").eval()
    img, meta = maxvit_batch(golden_mv, example_inputs)
    with torch.no_grad():
        got = model(image_input=torch.from_numpy(img).cuda(), metadata_input=torch.from_numpy(meta).cuda()).cpu().numpy()
    ref = golden_mv["ff_maxvit"]
    print(f"[parity] ff_maxvit via load_HF_model fp32: max|logit-ref|={np.abs(got - ref).max():.3e}")
    assert np.abs(got - ref).max() < 1.5e-4
    # training it: the frozen MaxViT branch feeds the trainable fusion head; an unfrozen MaxViT trunk has no backward
    model.train()
    for p in list(model.image_branch.parameters()) + list(model.meta_branch.parameters()):
        p.requires_grad = False
    lg = model(image_input=torch.from_numpy(img).cuda(), metadata_input=torch.from_numpy(meta).cuda())
    lg.sum().backward()
    assert model.combined_head[0].weight.grad is not None and torch.isfinite(model.combined_head[0].weight.grad).all()
    mm = btsbot.mm_MaxViT(dict(synth.canonical_config("mm_MaxViT", KIND))).cuda().train()
    with pytest.raises(NotImplementedError):
        mm(image_input=torch.from_numpy(img[:2]).cuda(), metadata_input=torch.from_numpy(meta[:2]).cuda())
