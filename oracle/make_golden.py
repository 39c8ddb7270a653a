"""Generate ``tests/golden/*`` by executing the REFERENCE's own source files (build container only).

Run:  python -m oracle.make_golden          (needs /root/reference; never runs on the GPU box)

* `btsbot/architectures.py` is exec'd verbatim with the ``timm`` shim (oracle/timm_shim.py);
* `btsbot/alert_utils.py` is exec'd verbatim with inert stand-ins for ``bson``/``matplotlib``/``astropy``
  (its numeric code -- crop_norm_cutout, crop_triplets, the nan_to_num/normalise/pad tail of make_triplet --
  runs unmodified; only gunzip'd-FITS decoding is replaced by a 2880-byte-block FITS reader below).

Weights come from ``btsbot_b200.synth.make_state_dict`` (numpy Philox, stream-stable), so the fixtures stay
tiny: inputs are regenerated from the seed, only reference OUTPUTS are stored.
"""
import gzip
import importlib.util
import io
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/btsbot"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from btsbot_b200 import synth                       # noqa: E402
from oracle import timm_shim                        # noqa: E402


def load_reference_module(name, stubs=()):
    for s in stubs:
        parts = s.split(".")
        for i in range(1, len(parts) + 1):
            sys.modules.setdefault(".".join(parts[:i]), types.ModuleType(".".join(parts[:i])))
    spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    return spec, mod


def ref_architectures():
    timm_shim.install()
    spec, mod = load_reference_module("architectures")
    spec.loader.exec_module(mod)
    return mod


def fits_bytes(arr: np.ndarray) -> bytes:
    """Minimal single-HDU FITS image, BITPIX -32 (big-endian float32), gzip'd like ZTF stamps."""
    cards = [f"SIMPLE  = {'T':>20}", f"BITPIX  = {-32:>20}", f"NAXIS   = {2:>20}",
             f"NAXIS1  = {arr.shape[1]:>20}", f"NAXIS2  = {arr.shape[0]:>20}", "END"]
    hdr = "".join(c.ljust(80) for c in cards).ljust(2880).encode("ascii")
    data = arr.astype(">f4").tobytes()
    data += b"\0" * (-len(data) % 2880)
    return gzip.compress(hdr + data)


class _FakeHDUList(list):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _fits_open(bio):
    raw = bio.read()
    hdr = raw[:2880].decode("ascii")
    kv = {hdr[i:i + 8].strip(): hdr[i + 10:i + 80].strip() for i in range(0, 2880, 80)}
    w, h = int(kv["NAXIS1"]), int(kv["NAXIS2"])
    data = np.frombuffer(raw[2880:2880 + 4 * w * h], dtype=">f4").reshape(h, w).astype(np.float32)
    return _FakeHDUList([types.SimpleNamespace(data=data.copy())])


def ref_alert_utils():
    spec, mod = load_reference_module(
        "alert_utils", stubs=("bson.json_util", "matplotlib.colors", "matplotlib.pyplot", "astropy.io.fits"))
    sys.modules["bson.json_util"].loads = lambda x: x
    sys.modules["bson.json_util"].dumps = lambda x: x
    sys.modules["matplotlib.colors"].LogNorm = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["astropy.io"].fits = sys.modules["astropy.io.fits"]
    sys.modules["astropy.io.fits"].open = _fits_open
    spec.loader.exec_module(mod)
    return mod


MODEL_CASES = {
    "mm_nano": ("mm_ConvNeXt", "convnext_nano.d1h_in1k", {}),
    "mm_pico": ("mm_ConvNeXt", "convnext_pico.d1_in1k", {}),
    "mm_nano_LS": ("mm_ConvNeXt", "convnext_nano.d1h_in1k", {"train_data_version": "v12LS"}),
    "img_nano": ("ConvNeXt", "convnext_nano.d1h_in1k", {}),
    "img_pico": ("ConvNeXt", "convnext_pico.d1_in1k", {}),
    "ff_pico": ("frozen_fusion", "convnext_pico.d1_in1k", {}),
    "um_nn": ("um_nn", "convnext_pico.d1_in1k", {}),
}


def case_config(case):
    name, kind, extra = MODEL_CASES[case]
    cfg = synth.canonical_config(name, kind)
    cfg.update(extra)
    return cfg


def adversarial_stamps(seed=2):
    """Stamp sets for the make_triplet tail: full, ragged (63,40)/(35,63)/(1,1), NaN/inf pixels, all-zero."""
    g = np.random.default_rng(seed)
    shapes = [((63, 63),) * 3, ((63, 40), (35, 63), (63, 63)), ((1, 1), (63, 63), (20, 17)), ((63, 63),) * 3,
              ((63, 63),) * 3, ((63, 63),) * 3]
    sets = []
    for i, shp in enumerate(shapes):
        st = [(1.0 + g.standard_normal(s)).astype(np.float32) * 100 for s in shp]
        if i == 3:
            st[0][5, 7] = np.nan
            st[1][0, 0] = np.nan
            st[2][10:20, 10:20] = np.nan
        if i == 4:
            st[1][:] = 0.0
        if i == 5:
            st[2][3, 3] = np.inf
        sets.append(st)
    return sets


def main():
    os.makedirs(GOLD, exist_ok=True)
    arch = ref_architectures()
    au = ref_alert_utils()

    # ---- example data (reference fixture for BASELINE config 1), stored as exact float32 -------------
    import pandas as pd
    trip = np.load(os.path.join(REF, "example_data/usage_triplets.npy"))
    assert np.array_equal(trip.astype(np.float32).astype(np.float64), trip)
    cand = pd.read_csv(os.path.join(REF, "example_data/usage_candidates.csv"))
    meta = cand[synth.METADATA_COLS].values.astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "example_inputs.npz"),
                        triplets=trip.astype(np.float32), metadata=meta,
                        labels=cand["label"].values.astype(np.int64))

    # ---- model goldens: reference architectures.py (verbatim) on example + synthetic alerts ----------
    nsyn = 25
    syn_t = synth.make_triplets(nsyn, start=1000)
    syn_m = synth.make_metadata(nsyn, start=1000)
    img = torch.from_numpy(np.ascontiguousarray(np.transpose(np.concatenate([trip.astype(np.float32), syn_t]), (0, 3, 1, 2))))
    met = torch.from_numpy(np.concatenate([meta, syn_m]))
    out = {}
    torch.set_num_threads(8)
    for case in MODEL_CASES:
        cfg = case_config(case)
        model = getattr(arch, cfg["model_name"])(cfg)
        sd = synth.make_state_dict(cfg, seed=2)
        timm_shim.load_timm_keys(model, sd)
        model.eval()

        def run(model):
            with torch.no_grad():
                if cfg["model_name"] in ("mm_ConvNeXt", "frozen_fusion"):
                    return model(image_input=img, metadata_input=met)
                elif cfg["model_name"] == "um_nn":
                    return model(input_data=met)
                return model(input_data=img)
        raw = run(model).numpy().astype(np.float64)
        # calibration constants (stored; tests re-apply them): logits -> min(10, 0.5/std) * (logit - median)
        shift, scale = float(np.median(raw)), float(min(10.0, 0.5 / raw.std()))
        sd = synth.apply_calibration(sd, cfg, scale, shift)
        timm_shim.load_timm_keys(model, sd)
        logits = run(model)
        out[case + "_cal"] = np.array([scale, shift], dtype=np.float64)
        out[case] = logits.numpy().astype(np.float32)
        print(case, logits.shape, float(logits.min()), float(logits.max()),
              "pos frac", float((logits > 0).float().mean()))
        # reference-owned state-dict key names (timm trunk keys excluded: the shim trunk is torchvision-keyed)
        keys = sorted(k for k in model.state_dict() if "features." not in k and ".fc." not in k)
        out[case + "_keys"] = np.array(keys)
    np.savez_compressed(os.path.join(GOLD, "model_logits.npz"), nsyn=nsyn, **out)

    # ---- preprocessing goldens: reference alert_utils.py (verbatim) ----------------------------------
    pre = {}
    t8 = synth.make_triplets(2, start=5000, dtype=np.float64) * 37.5     # un-normalised float64
    for s in (63, 49, 32, 31):
        pre[f"crop{s}_f64"] = au.crop_triplets(t8.copy(), s)
        pre[f"crop{s}_f32"] = au.crop_triplets(t8.astype(np.float32), s)
    for i, stamps in enumerate(adversarial_stamps()):
        alert = {"candidate": {"candid": i}}
        for nm, st in zip(("Science", "Template", "Difference"), stamps):
            alert["cutout" + nm] = {"stampData": fits_bytes(st)}
        with np.errstate(all="ignore"):
            import contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                trip_i, drop = au.make_triplet(alert, normalize=True)
        pre[f"tail{i}"] = trip_i
        pre[f"tail{i}_drop"] = np.array(drop)
    np.savez_compressed(os.path.join(GOLD, "preprocess.npz"), **pre)
    print("wrote", os.listdir(GOLD))


MAXVIT_CASES = {
    "mm_maxvit": ("mm_MaxViT", "maxvit_tiny_rw_224.sw_in1k", {}),
    "img_maxvit": ("MaxViT", "maxvit_tiny_rw_224.sw_in1k", {}),
    # what the published BTSbot-maxvit-tiny-*-metadata checkpoints are (to_HF.py:143-177; architectures.py:304-308)
    "ff_maxvit": ("frozen_fusion", "maxvit_tiny_rw_224.sw_in1k", {}),
}
#: MaxViT is ~40x the FLOPs of ConvNeXt-nano: goldens on 8 shipped example alerts + 8 synthetic ones
MAXVIT_EXAMPLE, MAXVIT_SYN = 8, 8


def maxvit_batch():
    trip = np.load(os.path.join(REF, "example_data/usage_triplets.npy")).astype(np.float32)
    import pandas as pd
    cand = pd.read_csv(os.path.join(REF, "example_data/usage_candidates.csv"))
    meta = cand[synth.METADATA_COLS].values.astype(np.float32)
    sel = np.linspace(0, len(trip) - 1, MAXVIT_EXAMPLE).round().astype(int)       # spans both labels
    t = np.concatenate([trip[sel], synth.make_triplets(MAXVIT_SYN, start=2000)])
    m = np.concatenate([meta[sel], synth.make_metadata(MAXVIT_SYN, start=2000)])
    return sel, np.ascontiguousarray(t.transpose(0, 3, 1, 2)), m


def main_maxvit():
    """Goldens for `architectures.py:25-101` (MaxViT / mm_MaxViT) executed verbatim on the module-based MaxViT twin
    of the timm shim: pins the reference-owned wrapper (bilinear resize, head surgery, metadata branch, fusion head)."""
    arch = ref_architectures()
    sel, img, met = maxvit_batch()
    img, met = torch.from_numpy(img), torch.from_numpy(met)
    out = {"example_idx": sel, "nsyn": MAXVIT_SYN}
    torch.set_num_threads(8)
    for case, (name, kind, extra) in MAXVIT_CASES.items():
        cfg = synth.canonical_config(name, kind)
        cfg.update(extra)
        model = getattr(arch, name)(cfg)
        sd = synth.make_state_dict(cfg, seed=2)
        timm_shim.load_timm_keys(model, sd)
        model.eval()

        def run(model):
            with torch.no_grad():
                return model(image_input=img, metadata_input=met) if name != "MaxViT" else model(input_data=img)
        raw = run(model).numpy().astype(np.float64)
        shift, scale = float(np.median(raw)), float(min(10.0, 0.5 / raw.std()))
        sd = synth.apply_calibration(sd, cfg, scale, shift)
        timm_shim.load_timm_keys(model, sd)
        logits = run(model)
        out[case + "_cal"] = np.array([scale, shift], dtype=np.float64)
        out[case] = logits.numpy().astype(np.float32)
        print(case, tuple(logits.shape), float(logits.min()), float(logits.max()), "pos frac",
              float((logits > 0).float().mean()), "gain", scale)
        out[case + "_keys"] = np.array(sorted(k for k in model.state_dict() if ".head.fc." not in k))
    np.savez_compressed(os.path.join(GOLD, "maxvit_logits.npz"), **out)


LEGACY_CFG = dict(conv_kernel=5, conv1_channels=8, conv2_channels=16, conv_dropout1=0.5, conv_dropout2=0.55,
                  metadata_cols=list(synth.METADATA_COLS), meta_fc1_neurons=16, meta_dropout=0.25, meta_fc2_neurons=8,
                  comb_fc1_neurons=16, comb_fc2_neurons=4, comb_dropout=0.2, fc1_neurons=16, fc2_neurons=4, dropout=0.2)


def main_legacy():
    """Goldens for the legacy classes (`architectures.py:174-274` mm_cnn / um_cnn) executed verbatim: a small instance
    (8 / 16 conv channels) whose seeded torch-initialised state dict travels inside the fixture with the logits."""
    arch = ref_architectures()
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(6, start=3000).transpose(0, 3, 1, 2))) * 63.0   # O(1) pixels
    met = torch.from_numpy(synth.make_metadata(6, start=3000))
    out = {}
    for name in ("mm_cnn", "um_cnn"):
        torch.manual_seed(7)
        model = getattr(arch, name)(dict(LEGACY_CFG)).eval()
        with torch.no_grad():
            for k, v in model.state_dict().items():          # non-trivial BatchNorm statistics
                if k.endswith("running_mean"):
                    v.copy_(torch.from_numpy(synth.METADATA_MOMENTS[:, 0].astype(np.float32)))
                if k.endswith("running_var"):
                    v.copy_(torch.from_numpy((synth.METADATA_MOMENTS[:, 1] ** 2).astype(np.float32)))
            logits = model(image_input=img, metadata_input=met) if name == "mm_cnn" else model(input_data=img)
        for k, v in model.state_dict().items():
            out[f"{name}/{k}"] = v.numpy()
        out[name] = logits.numpy()
        print(name, logits.flatten().tolist())
    np.savez_compressed(os.path.join(GOLD, "legacy_cnn.npz"), **out)


if __name__ == "__main__":
    if "--legacy" in sys.argv:
        main_legacy()
    elif "--maxvit" in sys.argv:
        main_maxvit()
    else:
        main()
