"""Packaging a trained run directory into the published-model layout (`btsbot/to_HF.py:10-43,143-177`), i.e. the inverse
of :func:`btsbot_b200.from_HF.load_HF_model`:

    run directory (train.py output)              published model directory
    ---------------------------------            --------------------------------------------
    report.json["train_config"]          ->      train_config.json
    best_model.pth (state dict)          ->      pytorch_model.bin   (round-tripped through the model class)

Same function names, file names and error behaviour as the reference; the Hugging Face upload itself
(`to_HF.py:180-215`: ``HfApi.create_repo`` / ``upload_file``, model card) is network control plane and not provided --
:func:`package_model` leaves a directory that ``load_HF_model`` reads and the reference's uploader can push unchanged.
"""
import json
import os
import shutil

import torch

from . import architectures
from .from_HF import get_local_model_dir

#: (trunk family, pre-training) -> base model named in the model card (to_HF.py:165-177)
_BASE_MODELS = {
    ("maxvit", "galaxyzoo"): "mwalmsley/baseline-encoder-regression-maxvit_tiny",
    ("maxvit", "imagenet"): "timm/maxvit_tiny_rw_224.sw_in1k",
    ("maxvit", "randinit"): "timm/maxvit_tiny_rw_224.sw_in1k",
    ("convnext", "galaxyzoo"): "mwalmsley/zoobot-encoder-convnext_pico",
    ("convnext", "imagenet"): "timm/convnext_pico.d1_in1k",
    ("convnext", "randinit"): "timm/convnext_pico.d1_in1k",
}


def prep_config(model_dir: str) -> dict:
    """``report.json`` -> ``train_config.json``; returns the config."""
    report_path = os.path.join(model_dir, "report.json")
    if not os.path.exists(report_path):
        raise FileNotFoundError(f"Report file not found: {report_path}")
    with open(report_path) as fh:
        config = json.load(fh)["train_config"]
    with open(os.path.join(model_dir, "train_config.json"), "w") as fh:
        json.dump(config, fh, indent=2)
    return config


def prep_model(model_dir: str, config: dict) -> None:
    """``best_model.pth`` -> ``pytorch_model.bin`` through the model class (a strict load: a checkpoint that does not fit
    the config fails here, not at the user's ``load_HF_model``)."""
    model_path = os.path.join(model_dir, "best_model.pth")
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"Model file not found: {model_path}")
    model = getattr(architectures, config["model_name"])(config)
    model.load_state_dict(torch.load(model_path, map_location=torch.device("cpu")))
    torch.save(model.state_dict(), os.path.join(model_dir, "pytorch_model.bin"))


def config_to_params(config: dict):
    """``(architecture, multi_modal, pretrain)`` of a training config, as `from_HF` names them."""
    multi_modal = config["model_name"] == "frozen_fusion"
    kind = (config["image_model_config"] if multi_modal else config)["model_kind"]
    image_config = config["image_model_config"] if multi_modal else config
    architecture = next((a for a in ("maxvit", "convnext") if a in kind), None)
    if architecture is None:
        raise ValueError("Couldn't understand architecture")
    if "mwalmsley" in kind:
        pretrain = "galaxyzoo"
    elif not image_config.get("pretrained", True):
        pretrain = "randinit"
    elif "in1k" in kind:
        pretrain = "imagenet"
    else:
        raise ValueError("Couldn't understand pre-training regimen")
    return architecture, multi_modal, pretrain


def get_HF_basemodel(arch: str, pretrain: str) -> str:
    try:
        return _BASE_MODELS[(arch, pretrain)]
    except KeyError:
        raise ValueError(f"Invalid architecture: {arch} or pre-training regimen: {pretrain}") from None


def package_model(model_dir: str, models_root: str = "") -> str:
    """``prep_config`` + ``prep_model`` on a run directory, then copy the two published files to
    ``<models_root>/models/BTSbot-<trunk>-<pretraining>[-metadata]`` -- the directory ``load_HF_model`` resolves for the
    same ``(architecture, multi_modal, pretrain)``.  Returns that directory."""
    config = prep_config(model_dir)
    prep_model(model_dir, config)
    target = os.path.join(models_root, get_local_model_dir(*config_to_params(config)))
    os.makedirs(target, exist_ok=True)
    for name in ("pytorch_model.bin", "train_config.json"):
        shutil.copyfile(os.path.join(model_dir, name), os.path.join(target, name))
    return target
