"""Model cases shared by the CPU and GPU parity tests (mirror of oracle/make_golden.MODEL_CASES)."""
from btsbot_b200 import synth

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


def case_state_dict(case, golden_logits, seed=2):
    """numpy state dict with the golden file's calibration applied."""
    cfg = case_config(case)
    scale, shift = golden_logits[case + "_cal"]
    return cfg, synth.apply_calibration(synth.make_state_dict(cfg, seed=seed), cfg, float(scale), float(shift))
