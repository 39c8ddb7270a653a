"""CPU: the C-ABI library loads and exports every declared symbol; the Python surface mirrors the reference's;
the product refuses to run without an sm_100 GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import btsbot_b200 as btsbot
from btsbot_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "btsbot_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(btsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/btsbot_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.lib().btsb_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal path")
def test_no_gpu_means_refusal_not_fallback():
    lib = _lib.lib()
    assert lib.btsb_device_ok() == -2
    rc = lib.btsb_score_epilogue(None, 4, None, None, None)
    assert rc == -2 and b"CUDA" in lib.btsb_last_error_string()
    cfg = synth.canonical_config("mm_ConvNeXt", "convnext_pico.d1_in1k")
    model = btsbot.mm_ConvNeXt(cfg).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(image_input=torch.zeros(1, 3, 63, 63), metadata_input=torch.zeros(1, 25))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        btsbot.alert_utils.crop_triplets(np.zeros((1, 63, 63, 3)), 49)


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libbtsbot_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_package_surface_matches_reference_init():
    # btsbot/__init__.py:28-46
    for name in ["architectures", "utils", "alert_utils", "FlexibleDataset", "RandomRightAngleRotation", "make_report",
                 "MaxViT", "ConvNeXt", "mm_MaxViT", "mm_ConvNeXt", "mm_cnn", "um_cnn", "um_nn", "frozen_fusion",
                 "download_HF_model", "load_HF_model", "__version__"]:
        assert hasattr(btsbot, name), name
    for fn in ["crop_norm_cutout", "crop_triplets", "make_triplet", "extract_triplets"]:
        assert callable(getattr(btsbot.alert_utils, fn))
    assert btsbot.architectures.get_model_image_size("maxvit_tiny_rw_224.sw_in1k") == 224
    assert btsbot.architectures.get_model_image_size("maxvit_large_tf_384.in1k") == 384
    assert btsbot.architectures.get_model_image_size("convnext_nano.d1h_in1k") == 224


def test_from_hf_name_mangling():
    f = btsbot.from_HF
    assert f.get_HF_model_link("convnext", True, "randinit") == "nabeelr/BTSbot-convnext-pico-randinit-metadata"
    assert f.get_local_model_dir("maxvit", False, "imagenet") == os.path.join("models", "BTSbot-maxvit-tiny-in1k")
    with pytest.raises(ValueError):
        f.validate_model_params("resnet", False, "randinit")
    with pytest.raises(ValueError):
        f.validate_model_params("convnext", False, "laion")


def test_timm_key_names_and_shapes():
    """Trunk parameter names/shapes follow timm's ConvNeXt so reference checkpoints load."""
    cfg = synth.canonical_config("mm_ConvNeXt", "convnext_nano.d1h_in1k")
    sd = btsbot.mm_ConvNeXt(cfg).state_dict()
    assert sd["convnext_backbone.stem.0.weight"].shape == (80, 3, 4, 4)
    assert sd["convnext_backbone.stages.1.downsample.1.weight"].shape == (160, 80, 2, 2)
    assert sd["convnext_backbone.stages.2.blocks.7.conv_dw.weight"].shape == (320, 1, 7, 7)
    assert sd["convnext_backbone.stages.3.blocks.1.mlp.fc1.weight"].shape == (2560, 640, 1, 1)
    assert sd["convnext_backbone.stages.0.blocks.0.gamma"].shape == (80,)
    assert float(sd["convnext_backbone.stages.0.blocks.0.gamma"][0]) == pytest.approx(1e-6)
    assert not any(k.startswith("convnext_backbone.head") for k in sd)      # non-LS: head is a bare Flatten
    assert not any("stages.0.downsample" in k for k in sd)
    n = sum(v.numel() for k, v in sd.items() if "num_batches" not in k and "running" not in k)
    assert n == 15070643                                                    # 15.071 M params
    pico = btsbot.ConvNeXt(synth.canonical_config("ConvNeXt", "convnext_pico.d1_in1k")).state_dict()
    assert pico["convnext.head.8.weight"].shape == (1, 8) and pico["convnext.head.1.weight"].shape == (512,)


def test_frozen_fusion_surgery():
    cfg = synth.canonical_config("frozen_fusion", "convnext_pico.d1_in1k")
    m = btsbot.frozen_fusion(cfg)
    keys = set(m.state_dict())
    assert "image_branch.convnext.head.1.weight" in keys and "image_branch.convnext.head.3.weight" not in keys
    assert "meta_branch.network.4.weight" in keys and "meta_branch.network.6.weight" not in keys
    assert m.combined_head[0].in_features == 512 + 128


def test_legacy_cnn_pass_through_matches_reference():
    """mm_cnn / um_cnn (PyTorch pass-through classes outside the hot path): same state-dict keys and
    the same logits as the reference's architectures.py:174-274 executed verbatim (tests/golden/legacy_cnn.npz, made by
    `python -m oracle.make_golden --legacy`)."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "legacy_cnn.npz"))
    cfg = dict(conv_kernel=5, conv1_channels=8, conv2_channels=16, conv_dropout1=0.5, conv_dropout2=0.55,
               metadata_cols=list(synth.METADATA_COLS), meta_fc1_neurons=16, meta_dropout=0.25, meta_fc2_neurons=8,
               comb_fc1_neurons=16, comb_fc2_neurons=4, comb_dropout=0.2, fc1_neurons=16, fc2_neurons=4, dropout=0.2)
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(6, start=3000).transpose(0, 3, 1, 2))) * 63.0   # O(1) pixels
    met = torch.from_numpy(synth.make_metadata(6, start=3000))
    for name in ("mm_cnn", "um_cnn"):
        model = getattr(btsbot, name)(dict(cfg)).eval()
        sd = {k[len(name) + 1:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(name + "/")}
        assert set(sd) == set(model.state_dict())
        model.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out = model(image_input=img, metadata_input=met) if name == "mm_cnn" else model(input_data=img)
        assert np.abs(out.numpy() - gold[name]).max() < 1e-5
    # head surgery of frozen_fusion for a um_cnn image branch (architectures.py:314-317)
    m, dim = btsbot.frozen_fusion.remove_branch_head(btsbot.um_cnn(dict(cfg)), "um_cnn")
    assert dim == 16 * 7 * 7 and isinstance(m.head, torch.nn.Identity)


def test_unsupported_models_fail_loudly():
    with pytest.raises(ValueError):
        btsbot.ConvNeXt(dict(synth.canonical_config("ConvNeXt"), model_kind="convnext_base"))


def test_flexible_dataset_and_rotation():
    img = torch.arange(2 * 3 * 5 * 5, dtype=torch.float32).view(2, 3, 5, 5)
    meta, lab = torch.ones(2, 4), torch.tensor([0, 1])
    assert len(btsbot.FlexibleDataset(img, meta, lab)[1]) == 3
    assert len(btsbot.FlexibleDataset(images=img, labels=lab)[0]) == 2
    assert btsbot.FlexibleDataset(metadata=meta, labels=lab)[1][0].shape == (4,)
    np.random.seed(2)
    angles = [int(np.random.choice([0, 90, 180, 270])) for _ in range(8)]
    np.random.seed(2)
    rot = btsbot.RandomRightAngleRotation()
    import torchvision.transforms.v2.functional as TF
    for a in angles:
        assert torch.equal(rot(img[0]), TF.rotate(img[0], a))               # exact permutation, same RNG stream


def test_adamw_batch_struct_matches_header():
    """ctypes mirror of btsb_adamw_batch (multi-tensor AdamW descriptors passed by value to the kernel)."""
    import ctypes
    import re
    from btsbot_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "btsbot_b200.h")).read()
    n = int(re.search(r"#define\s+BTSB_ADAMW_BATCH\s+(\d+)", hdr).group(1))
    assert _lib.ADAMW_BATCH == n
    assert ctypes.sizeof(_lib.AdamwBatch) == 5 * 8 * n + 4 * (n + 1) + 4
    assert ctypes.sizeof(_lib.AdamwBatch) < 4096          # must fit the kernel parameter space next to the scalars


def test_head_params_struct_matches_header(tmp_path):
    """ctypes mirror of btsb_head_params against the C compiler's view of include/btsbot_b200.h (size and the
    offset of every field)."""
    import ctypes
    import subprocess
    from btsbot_b200 import _lib
    names = [f[0] for f in _lib.HeadParams._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "btsbot_b200.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(btsb_head_params));\n'
                   + "".join(f'  printf("%zu\\n", offsetof(btsb_head_params, {n}));\n' for n in names)
                   + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.HeadParams)
    assert out[1:] == [getattr(_lib.HeadParams, n).offset for n in names]


def test_training_precision_switch_is_host_logic_only():
    """precision="bf16" selects the tensor-core training GEMMs by model attribute; building the model needs no GPU."""
    import btsbot_b200 as btsbot
    from btsbot_b200 import synth, _autograd
    cfg = dict(synth.canonical_config("mm_ConvNeXt", "convnext_pico.d1_in1k"), precision="bf16")
    model = btsbot.mm_ConvNeXt(cfg)
    assert model._precision == "bf16" and model.set_precision("fp32")._precision == "fp32"
    assert _autograd._ld(1350) == 1352 and _autograd._ld(8) == 8       # transposed-operand pitch: multiple of 8 (16 B)
    with pytest.raises(ValueError):
        model.set_precision("fp8")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port timed on the host cores) must print ONE JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "alerts/sec" and d["unit"] == "alerts/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "alerts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_kernel_table_roofline_columns():
    """bench.py files every kernel family under the tensor or the HBM roofline and reports the fraction of
    the measured peak it reaches; pure host logic."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.kernel_bound("mlp_fused_320", "bf16") == "tensor"
    assert bench.kernel_bound("gemm_down", "bf16") == "tensor"
    assert bench.kernel_bound("head_gemm", "bf16") == "tensor"
    assert bench.kernel_bound("gemm_down", "fp32") == "hbm"            # CUDA-core fp32 GEMMs: no tensor roofline
    assert bench.kernel_bound("t_gemm_tn", "bf16") == "hbm"
    for name in ("dwln_15x80", "lnpatch", "stem_fused", "meta_head", "poolln"):
        assert bench.kernel_bound(name, "bf16") == "hbm"
    kernels = {"mlp_fused_320": {"gbs": 1000.0, "tflops": 700.0}, "dwln_15x80": {"gbs": 1600.0, "tflops": 40.0}}
    bench.add_roofline_fractions(kernels, {"hbm": 6400.0, "tf_sust": 1400.0}, "bf16")
    assert kernels["mlp_fused_320"]["bound"] == "tensor" and abs(kernels["mlp_fused_320"]["frac"] - 0.5) < 1e-12
    assert kernels["dwln_15x80"]["bound"] == "hbm" and abs(kernels["dwln_15x80"]["frac"] - 0.25) < 1e-12


def _fits_gz(arr, bitpix=-32, bscale=None, bzero=None):
    import gzip
    cards = [f"SIMPLE  = {'T':>20}", f"BITPIX  = {bitpix:>20}", f"NAXIS   = {2:>20}", f"NAXIS1  = {arr.shape[1]:>20}",
             f"NAXIS2  = {arr.shape[0]:>20}"]
    if bscale is not None:
        cards += [f"BSCALE  = {bscale:>20}", f"BZERO   = {bzero:>20} / offset of the stored integers"]
    cards.append("END")
    hdr = "".join(c.ljust(80) for c in cards).ljust(2880).encode("ascii")
    data = arr.astype({-32: ">f4", -64: ">f8", 16: ">i2", 32: ">i4", 8: "u1"}[bitpix]).tobytes()
    return gzip.compress(hdr + data + b"\0" * (-len(data) % 2880))


def test_batched_stamp_ingest_matches_numpy_decode():
    """`btsb_ingest_fits_gz` (host threads; no GPU involved): gunzip + FITS parse of a batch of stamps equals the plain
    gzip + numpy decode the reference's astropy call amounts to (alert_utils.py:141-147), for every pixel type, ragged
    shapes, NaN / inf pixels, and a batch large enough to use the thread pool; malformed input is an error."""
    from btsbot_b200 import alert_utils as au
    g = np.random.default_rng(3)
    a32 = g.standard_normal((63, 63)).astype(np.float32)
    a32[5, 7], a32[9, 9] = np.nan, np.inf
    cases = [(a32, -32, None, None), (g.standard_normal((35, 63)), -64, None, None),
             ((g.standard_normal((20, 17)) * 1000).astype(np.int16), 16, 2.0, 32768.0),
             ((g.standard_normal((1, 1)) * 1e6).astype(np.int32), 32, None, None),
             (np.arange(12, dtype=np.uint8).reshape(3, 4), 8, None, None)]
    blobs = [_fits_gz(*c) for c in cases] * 40                      # 200 stamps
    stamps, hw = au.decode_stamps(blobs, threads=4)
    assert stamps.shape == (200, 63 * 63) and hw.shape == (200, 2)
    for i, (arr, bitpix, bscale, bzero) in enumerate(cases * 40):
        h, w = arr.shape
        assert tuple(hw[i]) == (h, w)
        ref = arr.astype(np.float64) * (bscale or 1.0) + (bzero or 0.0) if bitpix > 0 else arr
        assert np.array_equal(stamps[i, :h * w].reshape(h, w), ref.astype(np.float32), equal_nan=True), (i, bitpix)
    assert au.decode_stamps([])[0].shape == (0, 63 * 63)
    for bad in (b"not a gzip stream", _fits_gz(np.zeros((64, 63), np.float32)), _fits_gz(a32)[:200]):
        with pytest.raises(RuntimeError):
            au.decode_stamps([blobs[0], bad])


def test_to_HF_packaging_round_trip(tmp_path, monkeypatch):
    """to_HF.prep_config / prep_model / config_to_params (to_HF.py:10-43,143-177): a run directory becomes the published
    layout that `load_HF_model` resolves; the Hugging Face upload itself is out of scope."""
    import json
    from btsbot_b200 import to_HF, from_HF
    cfg = synth.canonical_config("frozen_fusion", "convnext_pico.d1_in1k")
    run = tmp_path / "run"
    run.mkdir()
    with pytest.raises(FileNotFoundError):
        to_HF.prep_config(str(run))
    (run / "report.json").write_text(json.dumps({"train_config": cfg, "Training history": {}}))
    assert to_HF.prep_config(str(run)) == cfg and (run / "train_config.json").is_file()
    with pytest.raises(FileNotFoundError):
        to_HF.prep_model(str(run), cfg)
    sd = synth.to_torch(synth.make_state_dict(cfg, seed=4))
    torch.save(sd, run / "best_model.pth")
    monkeypatch.chdir(tmp_path)
    target = to_HF.package_model(str(run))
    assert target == from_HF.get_local_model_dir("convnext", True, "randinit") == os.path.join("models", "BTSbot-convnext-pico-randinit-metadata")
    back = torch.load(os.path.join(target, "pytorch_model.bin"), map_location="cpu")
    assert set(back) == set(sd) and all(torch.equal(back[k].reshape(-1), sd[k].reshape(-1)) for k in sd)
    assert to_HF.config_to_params(cfg) == ("convnext", True, "randinit")
    assert to_HF.config_to_params(synth.canonical_config("MaxViT", "maxvit_tiny_rw_224.sw_in1k")) == ("maxvit", False, "randinit")
    assert to_HF.config_to_params(dict(synth.canonical_config("ConvNeXt", "convnext_pico.d1_in1k"), pretrained=True)) == ("convnext", False, "imagenet")
    assert to_HF.get_HF_basemodel("maxvit", "galaxyzoo") == "mwalmsley/baseline-encoder-regression-maxvit_tiny"
    with pytest.raises(ValueError):
        to_HF.get_HF_basemodel("resnet", "imagenet")


def test_host_pack_bf16_matches_round_to_nearest_even():
    """btsb_host_pack_bf16 (host threads, no GPU involved): bit-equal to torch's float32 -> bfloat16 rounding for normal,
    subnormal, huge and infinite values at every thread count; NaN stays NaN; the pool survives repeated jobs."""
    import torch
    from btsbot_b200 import _lib
    lib = _lib.lib()
    n = 700 * 63 * 63 * 3 + 5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, generator=g)
    x[::1001] = float("nan"); x[1::1003] = float("inf"); x[2::1005] = -float("inf"); x[3::1007] = 1e-40; x[4::1009] = 3.4e38
    x[5::1011] = 1.00390625                         # exactly halfway between two bf16 values: ties to even
    ref = x.to(torch.bfloat16).view(torch.int16)
    for threads in (1, 2, 5, 0, 16, 3):
        out = torch.zeros(n, dtype=torch.int16)
        _lib.check(lib.btsb_host_pack_bf16(x.data_ptr(), out.data_ptr(), n, threads), "host_pack")
        same = (ref == out) | (torch.isnan(x) & torch.isnan(out.view(torch.bfloat16).float()))
        assert bool(same.all()), threads
    assert lib.btsb_host_pack_bf16(None, None, 5, 1) < 0
