"""Model-directory loader behind the reference's names (`btsbot/from_HF.py:16-81`).

A published BTSbot model is a directory ``models/BTSbot-<trunk>-<pretraining>[-metadata]`` holding
``train_config.json`` (which names the `architectures` class and its config) and ``pytorch_model.bin`` (its state
dict, reference / timm key layout).  Name resolution, error messages and the on-disk layout follow the reference so
that its model directories load unchanged; the Hugging Face download (`from_HF.py:43-56`) needs the network and
``huggingface_hub`` and is only attempted when the files are absent, as in the reference.
"""
import json
import os

import torch

device = "cuda" if torch.cuda.is_available() else "cpu"

#: public short names -> the trunk / pre-training tags used in the published model names (from_HF.py:16-29)
_TRUNKS = {"convnext": "convnext-pico", "maxvit": "maxvit-tiny"}
_PRETRAINING = {"imagenet": "in1k", "galaxyzoo": "galaxyzoo", "randinit": "randinit"}
_HF_OWNER = "nabeelr"
_MODEL_FILES = ("pytorch_model.bin", "train_config.json")


def validate_model_params(architecture: str, multi_modal: bool, pretrain: str):
    """``(architecture, multi_modal, pretrain)`` with the short names expanded; ``ValueError`` on unknown ones."""
    if architecture not in _TRUNKS:
        raise ValueError(f"Invalid architecture: {architecture}")
    if pretrain not in _PRETRAINING:
        raise ValueError(f"Invalid pre-training regimen: {pretrain}")
    return _TRUNKS[architecture], multi_modal, _PRETRAINING[pretrain]


def _model_name(architecture: str, multi_modal: bool, pretrain: str) -> str:
    trunk, multi_modal, tag = validate_model_params(architecture, multi_modal, pretrain)
    return "-".join(["BTSbot", trunk, tag] + (["metadata"] if multi_modal else []))


def get_HF_model_link(architecture: str, multi_modal: bool, pretrain: str) -> str:
    return f"{_HF_OWNER}/{_model_name(architecture, multi_modal, pretrain)}"


def get_local_model_dir(architecture: str, multi_modal: bool, pretrain: str) -> str:
    return os.path.join("models", _model_name(architecture, multi_modal, pretrain))


def download_HF_model(architecture: str, multi_modal: bool, pretrain: str):
    repo = get_HF_model_link(architecture, multi_modal, pretrain)
    target = get_local_model_dir(architecture, multi_modal, pretrain)
    try:
        from huggingface_hub import snapshot_download
    except ImportError as e:  # pragma: no cover
        raise RuntimeError(f"{target} is missing and huggingface_hub is not installed") from e
    print(f"Fetching model from HuggingFace Hub: {repo}")
    os.makedirs(target, exist_ok=True)
    snapshot_download(repo_id=repo, local_dir=target)
    print(f"Model downloaded to {target}")


def load_HF_model(architecture: str, multi_modal: bool, pretrain: str):
    """Build the class ``train_config.json`` names and load ``pytorch_model.bin`` into it (from_HF.py:59-81)."""
    from . import architectures
    model_dir = get_local_model_dir(architecture, multi_modal, pretrain)
    if any(not os.path.isfile(os.path.join(model_dir, name)) for name in _MODEL_FILES):
        print("Model files not present; downloading model...")
        download_HF_model(architecture, multi_modal, pretrain)
    with open(os.path.join(model_dir, "train_config.json")) as fh:
        config = json.load(fh)
    model = getattr(architectures, config["model_name"])(config).to(device)
    weights = torch.load(os.path.join(model_dir, "pytorch_model.bin"), map_location="cpu")
    model.load_state_dict(weights)
    return model
