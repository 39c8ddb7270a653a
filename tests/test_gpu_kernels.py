"""GPU: every C-ABI kernel against a plain PyTorch fp32 reference of the same op (run on the CPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rows_to_nchw(rows, B, H, W):
    return rows.float().view(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def _nchw_to_rows(x):
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()


def _ln2d(x, w, b):
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, 1e-6).permute(0, 3, 1, 2)


def _report(name, got, ref):
    err = (got - ref).abs().max().item()
    print(f"[parity] {name}: max|err|={err:.3e} ref_absmax={ref.abs().max().item():.3e}")
    return err


@pytest.mark.parametrize("C0,H,W,B", [(80, 63, 63, 5), (64, 63, 63, 3), (80, 64, 37, 2), (128, 16, 16, 9)])
@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_stem(cuda_dev, C0, H, W, B, odt):
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, H, W, generator=g) * 0.02 + 0.016
    w = torch.randn(C0, 3, 4, 4, generator=g) / 7
    b, lw, lb = torch.randn(C0, generator=g) * 0.1, torch.rand(C0, generator=g) + 0.5, torch.randn(C0, generator=g) * 0.1
    ref = _ln2d(F.conv2d(x, w, b, stride=4), lw, lb)
    got = ops.stem(x.to(cuda_dev), w.reshape(C0, 48).t().contiguous().to(cuda_dev), b.to(cuda_dev), lw.to(cuda_dev),
                   lb.to(cuda_dev), odt)
    h, wd = ref.shape[2:]
    err = _report(f"stem C0={C0} {H}x{W} {odt}", _rows_to_nchw(got.cpu(), B, h, wd), ref)
    assert err < (2e-5 if odt == torch.float32 else 3e-2)


@pytest.mark.parametrize("C,H,W,B", [(80, 15, 15, 7), (160, 7, 7, 33), (320, 3, 3, 70), (640, 1, 1, 130),
                                     (64, 15, 15, 3), (128, 7, 7, 5), (256, 3, 3, 5), (512, 1, 1, 5),
                                     (64, 9, 11, 4), (96, 5, 2, 3), (320, 3, 3, 2500), (640, 1, 1, 3000), (48, 3, 3, 40),
                                     (512, 2, 2, 9), (640, 2, 2, 5), (256, 4, 4, 6), (64, 19, 19, 3)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_dwln(cuda_dev, C, H, W, B, dt):
    """dt float16 = the fp16 residual stream of the bf16 mode: fp16 rows in, bf16 rows out (dtype code BF16_XF16); the
    inputs include fp16 subnormals and large magnitudes."""
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, C, H, W, generator=g)
    if dt == torch.float16:
        x.view(-1)[::97] *= 1e-6                      # fp16 subnormals
        x.view(-1)[5::1013] *= 3e3                    # large magnitudes (|x| up to ~1e4)
    if dt != torch.float32:
        x = x.to(dt).float()
    w = torch.randn(C, 1, 7, 7, generator=g) / 7
    b, lw, lb = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    ref = _ln2d(F.conv2d(x, w, b, padding=3, groups=C), lw, lb)
    got = ops.dwln(_nchw_to_rows(x).to(dt).to(cuda_dev), B, H, W, w.reshape(C, 49).t().contiguous().to(cuda_dev),
                   b.to(cuda_dev), lw.to(cuda_dev), lb.to(cuda_dev))
    assert got.dtype == (torch.float32 if dt == torch.float32 else torch.bfloat16)
    err = _report(f"dwln C={C} {H}x{W} {dt}", _rows_to_nchw(got.cpu(), B, H, W), ref)
    # bf16 output rows: half an ulp is 2^-9 of the value (the fp16 case's outliers give LayerNorm outputs up to ~sqrt(C))
    assert err < (3e-5 if dt == torch.float32 else max(4e-2, 2.0 ** -8 * ref.abs().max().item()))


@pytest.mark.parametrize("C,S,B", [(80, 15, 1500), (160, 7, 3000), (64, 15, 1500), (128, 7, 3000)])
def test_dwln_warp_specialised_handoff_is_deterministic(cuda_dev, C, S, B):
    """dwln5 hands the fp32 conv tile from the conv warps to the LayerNorm warps through two mbarriers (compute-sanitizer's
    racecheck does not model mbarrier phases and flags that hand-off, profiles/r02a_san): a persistent grid with ~10 images
    per CTA, launched repeatedly, must give bitwise identical rows every time and agree with the fp32 reference."""
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, S, S, generator=g).bfloat16()
    w = torch.randn(C, 1, 7, 7, generator=g) / 7
    b, lw, lb = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    args = (_nchw_to_rows(x.float()).bfloat16().to(cuda_dev), B, S, S, w.reshape(C, 49).t().contiguous().to(cuda_dev),
            b.to(cuda_dev), lw.to(cuda_dev), lb.to(cuda_dev))
    first = ops.dwln(*args)
    for _ in range(20):
        assert torch.equal(ops.dwln(*args), first)
    ref = _ln2d(F.conv2d(x.float()[:64], w, b, padding=3, groups=C), lw, lb)
    err = _report(f"dwln5 determinism C={C} {S}x{S}", _rows_to_nchw(first[:64 * S * S].cpu(), 64, S, S), ref)
    assert err < 4e-2


@pytest.mark.parametrize("C,H,W,B", [(80, 15, 15, 5), (160, 7, 7, 9), (320, 3, 3, 11), (64, 15, 15, 2), (256, 4, 6, 3)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_lnpatch_then_gemm_is_downsample(cuda_dev, C, H, W, B, dt):
    """dt float16: fp16 residual-stream rows in, bf16 patch matrix, bf16 GEMM writing fp16 rows (BF16_XF16)."""
    from btsbot_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, W, generator=g)
    if dt != torch.float32:
        x = x.to(dt).float()
    lw, lb = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    cout = 2 * C
    w = torch.randn(cout, C, 2, 2, generator=g) / (2 * C ** 0.5)
    if dt != torch.float32:
        w = w.bfloat16().float()
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(_ln2d(x, lw, lb), w, b, stride=2)
    patches = ops.lnpatch(_nchw_to_rows(x).to(dt).to(cuda_dev), B, H, W, lw.to(cuda_dev), lb.to(cuda_dev))
    ho, wo = ref.shape[2:]
    assert patches.shape == (B * ho * wo, 4 * C)
    wdt = torch.float32 if dt == torch.float32 else torch.bfloat16
    assert patches.dtype == wdt
    wt = w.permute(0, 2, 3, 1).reshape(cout, 4 * C).contiguous().to(wdt).to(cuda_dev)
    got = ops.gemm(patches, wt, b.to(cuda_dev), L.EPI_BIAS, out_dtype=dt if dt == torch.float16 else None)
    assert got.dtype == dt
    err = _report(f"downsample C={C} {H}x{W} {dt}", _rows_to_nchw(got.cpu(), B, ho, wo), ref)
    assert err < (3e-5 if dt == torch.float32 else 5e-2)


@pytest.mark.parametrize("C,HW,B,ln", [(640, 1, 70, True), (512, 9, 5, True), (80, 4, 3, False)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_poolln(cuda_dev, C, HW, B, ln, dt):
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B * HW, C, generator=g).to(dt)
    lw, lb = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    pooled = x.float().view(B, HW, C).mean(1)
    ref = F.layer_norm(pooled, (C,), lw, lb, 1e-6) if ln else pooled
    got = ops.poolln(x.to(cuda_dev), B, HW, lw.to(cuda_dev) if ln else None, lb.to(cuda_dev) if ln else None)
    assert _report(f"poolln C={C} HW={HW} {dt}", got.cpu(), ref) < 2e-5


GEMM_SHAPES = [(1000, 320, 80), (1000, 80, 320), (300, 640, 160), (513, 2560, 640), (128, 16, 16), (5000, 256, 1024),
               (77, 160, 640), (1, 640, 2560), (257, 512, 2048), (1125, 64, 256)]


def _gemm_ref(a, w, bias, epi, gamma, res):
    from btsbot_b200 import _lib as L
    v = a.double() @ w.double().t() + bias.double()
    if epi == L.EPI_BIAS_GELU:
        v = F.gelu(v)
    if epi == L.EPI_SCALE_RES:
        v = res.double() + gamma.double() * v
    return v.float()


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_f32(cuda_dev, M, N, K, epi):
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(5)
    a, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    bias, gamma, res = torch.randn(N, generator=g) * 0.1, torch.rand(N, generator=g) + 0.5, torch.randn(M, N, generator=g)
    ref = _gemm_ref(a, w, bias, epi, gamma, res)
    got = ops.gemm(a.to(cuda_dev), w.to(cuda_dev), bias.to(cuda_dev), epi,
                   gamma.to(cuda_dev) if epi == 2 else None, res.to(cuda_dev) if epi == 2 else None)
    assert _report(f"gemm f32 {M}x{N}x{K} epi{epi}", got.cpu(), ref) < 2e-5


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi,xdt", [(0, torch.bfloat16), (1, torch.bfloat16), (2, torch.bfloat16), (0, torch.float16),
                                     (2, torch.float16)])
def test_gemm_bf16_tcgen05(cuda_dev, M, N, K, epi, xdt):
    """xdt float16: bf16 operands, res / out in the fp16 residual stream (dtype code BF16_XF16)."""
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(6)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    bias, gamma = torch.randn(N, generator=g) * 0.1, torch.rand(N, generator=g) + 0.5
    res = torch.randn(M, N, generator=g).to(xdt)
    ref = _gemm_ref(a.float(), w.float(), bias, epi, gamma, res.float())
    got = ops.gemm(a.to(cuda_dev), w.to(cuda_dev), bias.to(cuda_dev), epi,
                   gamma.to(cuda_dev) if epi == 2 else None, res.to(cuda_dev) if epi == 2 else None,
                   out_dtype=xdt if xdt == torch.float16 else None)
    torch.cuda.synchronize()
    assert got.dtype == xdt
    got = got.float().cpu()
    err = _report(f"gemm bf16 {M}x{N}x{K} epi{epi} out {xdt}", got, ref)
    # exact fp32-accumulated product rounded once to bf16: half an ulp of |value| <= ~8 (fp16: 8 times finer)
    if xdt == torch.float16:
        assert err < 4.1e-3 and (got - ref).abs().mean().item() < 5e-4
    else:
        assert err < 3.2e-2 and (got - ref).abs().mean().item() < 4e-3


def test_gemm_bf16_cta_pair_mode_in_subprocess(cuda_dev):
    """The opt-in cta_group::2 path (BTSB_GEMM_2CTA=1: clusters of two CTAs, one M=256 UMMA, B halves staged per CTA)
    must give the same results as the default path; the switch is read once per process, hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from btsbot_b200 import ops\n"
        "g = torch.Generator().manual_seed(6); worst = 0.0\n"
        "for M, N, K, epi in [(1000, 320, 80, 0), (513, 2560, 640, 1), (5000, 256, 1024, 2), (257, 512, 2048, 2), (1, 640, 2560, 0), (73, 1280, 320, 1)]:\n"
        "    a = torch.randn(M, K, generator=g).bfloat16(); w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()\n"
        "    bias, gamma = torch.randn(N, generator=g) * 0.1, torch.rand(N, generator=g) + 0.5\n"
        "    res = torch.randn(M, N, generator=g).bfloat16()\n"
        "    acc = a.float() @ w.float().t() + bias\n"
        "    ref = acc if epi == 0 else (torch.nn.functional.gelu(acc) if epi == 1 else res.float() + gamma * acc)\n"
        "    got = ops.gemm(a.cuda(), w.cuda(), bias.cuda(), epi, gamma.cuda() if epi == 2 else None, res.cuda() if epi == 2 else None)\n"
        "    torch.cuda.synchronize(); worst = max(worst, (got.float().cpu() - ref).abs().max().item())\n"
        "print('PAIR_WORST', worst)\n" % root)
    env = dict(os.environ, BTSB_GEMM_2CTA="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float(out.stdout.strip().split("PAIR_WORST")[-1])
    print(f"[parity] gemm bf16 cta_group::2 path: worst |err| over 6 shapes = {worst:.3e}")
    assert worst < 3.2e-2


@pytest.mark.parametrize("C,M", [(80, 1000), (160, 777), (64, 4096), (128, 129), (80, 128 * 300 + 5), (96, 50),
                                 (320, 1000), (256, 777), (320, 128 * 150 + 5), (256, 128 * 149)])
@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float16])
def test_mlp_fused_tcgen05(cuda_dev, C, M, xdt):
    """fused fc1->GELU->fc2->*gamma->+res vs fp32 math on the same bf16 operands (hidden rounded to bf16 as the
    kernel does before the second GEMM); res / out bf16 or fp16 (the residual stream, dtype code BF16_XF16)."""
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(7)
    y = torch.randn(M, C, generator=g).bfloat16()
    res = torch.randn(M, C, generator=g).to(xdt)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).bfloat16()
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).bfloat16()
    b1, b2 = torch.randn(4 * C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) + 0.5
    hid = F.gelu(y.double() @ w1.double().t() + b1.double()).float().bfloat16()
    ref = (res.double() + gamma.double() * (hid.double() @ w2.double().t() + b2.double())).float()
    got = ops.mlp_fused(y.to(cuda_dev), res.to(cuda_dev), w1.to(cuda_dev), b1.to(cuda_dev), w2.to(cuda_dev),
                        b2.to(cuda_dev), gamma.to(cuda_dev))
    torch.cuda.synchronize()
    assert got.dtype == xdt
    got = got.float().cpu()
    err = _report(f"mlp_fused C={C} M={M} {xdt}", got, ref)
    # the hidden activation's GELU approximation + its bf16 rounding are common to both; the output rounding is 8x finer in fp16
    assert err < (2e-2 if xdt == torch.float16 else 4e-2) and (got - ref).abs().mean().item() < (2.5e-3 if xdt == torch.float16 else 5e-3)


@pytest.mark.parametrize("C,M", [(320, 1000), (256, 777), (320, 128 * 150 + 5), (80, 300)])
@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float16])
def test_mlp_fused_in_place(cuda_dev, C, M, xdt):
    """out == res: the wide variants add gamma * (fc2(...) + b2) to the residual rows with a bulk tensor reduction (bf16 add
    at the L2: the update is rounded to bf16 before the add, one more rounding than the out-of-place kernel); the narrow
    variants stage the rows and simply overwrite them.  Both must agree with the out-of-place result to bf16 rounding."""
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(9)
    y = torch.randn(M, C, generator=g).bfloat16().to(cuda_dev)
    res = torch.randn(M, C, generator=g).to(xdt).to(cuda_dev)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).bfloat16().to(cuda_dev)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).bfloat16().to(cuda_dev)
    b1, b2 = (torch.randn(4 * C, generator=g) * 0.1).to(cuda_dev), (torch.randn(C, generator=g) * 0.1).to(cuda_dev)
    gamma = (torch.rand(C, generator=g) + 0.5).to(cuda_dev)
    ref = ops.mlp_fused(y, res, w1, b1, w2, b2, gamma).float()
    buf = res.clone()
    got = ops.mlp_fused(y, buf, w1, b1, w2, b2, gamma, inplace=True)
    torch.cuda.synchronize()
    assert got.data_ptr() == buf.data_ptr()
    err = (got.float() - ref).abs()
    print(f"[parity] mlp_fused in place C={C} M={M} {xdt}: max|in-place - out-of-place| = {err.max().item():.3e}, "
          f"mean {err.mean().item():.3e}")
    # two roundings instead of one: at most ~1.5 ulp of the result (|values| < 8 -> bf16 ulp <= 2^-5, fp16 ulp <= 2^-8)
    scale = 0.125 if xdt == torch.float16 else 1.0
    assert err.max().item() <= 0.0625 * scale and err.mean().item() < 4e-3 * scale


@pytest.mark.parametrize("C0,H,W,B", [(80, 63, 63, 37), (64, 63, 63, 5), (128, 20, 36, 3), (16, 8, 8, 700), (96, 31, 47, 9),
                                      (80, 63, 63, 1200)])
def test_stem_tcgen05(cuda_dev, C0, H, W, B):
    """tensor-core stem (im2col + GEMM with bias+LayerNorm epilogue) vs fp32 conv+LN on the bf16-rounded operands."""
    from btsbot_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(B, 3, H, W, generator=g) * 0.02 + 0.016).bfloat16().float()
    w = (torch.randn(C0, 3, 4, 4, generator=g) / 7).bfloat16().float()
    b, lw, lb = torch.randn(C0, generator=g) * 0.1, torch.rand(C0, generator=g) + 0.5, torch.randn(C0, generator=g) * 0.1
    ref = _ln2d(F.conv2d(x, w, b, stride=4), lw, lb)
    wp = torch.zeros(C0, 64)
    wp[:, :48] = w.reshape(C0, 48)
    got, patches = ops.stem_tc(x.to(cuda_dev), wp.bfloat16().to(cuda_dev), b.to(cuda_dev), lw.to(cuda_dev), lb.to(cuda_dev))
    torch.cuda.synchronize()
    h, wd = ref.shape[2:]
    # im2col is exact (inputs are bf16-representable): patch (b,oy,ox) row k=(ci,ky,kx)
    unf = F.unfold(x, kernel_size=4, stride=4).transpose(1, 2).reshape(B * h * wd, 48)
    assert torch.equal(patches[:, :48].float().cpu(), unf) and patches[:, 48:].abs().max().item() == 0
    err = _report(f"stem tcgen05 C0={C0} {H}x{W}", _rows_to_nchw(got.cpu(), B, h, wd), ref)
    assert err < 3e-2
    # the one-kernel stem (producer warps build the same im2col rows in shared memory) must give the same rows: the MMA
    # sees identical operands, only the zero K-step 48..63 is skipped
    if C0 in (64, 80, 96):
        fused = ops.stem_fused(x.to(cuda_dev), wp.bfloat16().to(cuda_dev), b.to(cuda_dev), lw.to(cuda_dev), lb.to(cuda_dev))
        torch.cuda.synchronize()
        # same MMA operands; the LayerNorm statistics are summed in a different order (one pass over the row)
        assert (fused.float().cpu() - got.float().cpu()).abs().max() <= 2 ** -6, (fused.float().cpu() - got.float().cpu()).abs().max()
        assert _report(f"stem fused C0={C0} {H}x{W}", _rows_to_nchw(fused.cpu(), B, h, wd), ref) < 3e-2
        # ... and as the opening rows of the fp16 residual stream: the same values rounded to fp16 instead of bf16
        f16 = ops.stem_fused(x.to(cuda_dev), wp.bfloat16().to(cuda_dev), b.to(cuda_dev), lw.to(cuda_dev), lb.to(cuda_dev),
                             out_dtype=torch.float16)
        torch.cuda.synchronize()
        assert f16.dtype == torch.float16
        assert torch.equal(f16.float().bfloat16(), fused) or (f16.float() - fused.float()).abs().max() <= 2 ** -6
        assert _report(f"stem fused fp16 rows C0={C0} {H}x{W}", _rows_to_nchw(f16.cpu(), B, h, wd), ref) < 1.2e-2


def test_gemm_rejects_bad_arguments(cuda_dev):
    from btsbot_b200 import ops
    a = torch.zeros(8, 24, device=cuda_dev, dtype=torch.bfloat16)
    w = torch.zeros(16, 24, device=cuda_dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiples of 16"):
        ops.gemm(a, w, torch.zeros(16, device=cuda_dev))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a.cpu(), w.cpu(), torch.zeros(16))


def test_score_epilogue(cuda_dev):
    from btsbot_b200 import ops
    lg = torch.linspace(-8, 8, 1001).view(-1, 1)
    s, lab = ops.score(lg.to(cuda_dev))
    assert (s.cpu() - torch.sigmoid(lg)).abs().max() < 1e-6
    ref_lab = (torch.sigmoid(lg) > 0.5).to(torch.uint8)
    assert torch.equal(lab.cpu(), ref_lab)


def test_fp16_residual_stream_saturates_instead_of_overflowing(cuda_dev):
    """BTSB_BF16_XF16 stores (GEMM epilogue, fused-MLP epilogue out of place and in place): a value beyond the fp16 range
    becomes +-65504, never inf / NaN -- a checkpoint with an outlier channel degrades gracefully instead of poisoning the
    LayerNorm that reads the row next."""
    from btsbot_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(13)
    M, K, N = 300, 64, 64
    a = torch.randn(M, K, generator=g).bfloat16().to(cuda_dev)
    w = (torch.randn(N, K, generator=g) / 8).bfloat16().to(cuda_dev)
    bias = torch.zeros(N)
    bias[3], bias[7] = 1.0e5, -2.0e5
    out = ops.gemm(a, w, bias.to(cuda_dev), L.EPI_BIAS, out_dtype=torch.float16)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert (out[:, 3] == 65504).all() and (out[:, 7] == -65504).all()
    C = 80
    y = torch.randn(M, C, generator=g).bfloat16().to(cuda_dev)
    res = torch.randn(M, C, generator=g).half()
    res[:, 5] = 65000.0
    res = res.to(cuda_dev)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).bfloat16().to(cuda_dev)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).bfloat16().to(cuda_dev)
    b1 = torch.zeros(4 * C).to(cuda_dev)
    b2 = torch.zeros(C)
    b2[5] = 2000.0
    gamma = torch.ones(C).to(cuda_dev)
    got = ops.mlp_fused(y, res, w1, b1, w2, b2.to(cuda_dev), gamma)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all() and (got[:, 5] == 65504).all()
