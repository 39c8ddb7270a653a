"""Model-directory loader with the reference's names and name mangling (`btsbot/from_HF.py:16-81`).

Model files are read from ``models/<BTSbot-...>/{train_config.json,pytorch_model.bin}`` exactly like the
reference.  The Hugging Face download (`from_HF.py:43-56`) needs the network; it is attempted only when
``huggingface_hub`` is importable and the files are absent, as in the reference.
"""
import json
import os

import torch

device = "cuda" if torch.cuda.is_available() else "cpu"


def validate_model_params(architecture: str, multi_modal: bool, pretrain: str):
    if architecture == "convnext":
        architecture = "convnext-pico"
    elif architecture == "maxvit":
        architecture = "maxvit-tiny"
    else:
        raise ValueError(f"Invalid architecture: {architecture}")
    if pretrain == "imagenet":
        pretrain = "in1k"
    elif pretrain not in ["galaxyzoo", "randinit"]:
        raise ValueError(f"Invalid pre-training regimen: {pretrain}")
    return architecture, multi_modal, pretrain


def get_HF_model_link(architecture: str, multi_modal: bool, pretrain: str) -> str:
    architecture, multi_modal, pretrain = validate_model_params(architecture, multi_modal, pretrain)
    return "nabeelr/BTSbot-" + architecture + "-" + pretrain + ("-metadata" if multi_modal else "")


def get_local_model_dir(architecture: str, multi_modal: bool, pretrain: str) -> str:
    return os.path.join("models", get_HF_model_link(architecture, multi_modal, pretrain).split("/")[-1])


def download_HF_model(architecture: str, multi_modal: bool, pretrain: str):
    link = get_HF_model_link(architecture, multi_modal, pretrain)
    model_dir = os.path.join("models", link.split("/")[-1])
    try:
        from huggingface_hub import snapshot_download
    except ImportError as e:  # pragma: no cover
        raise RuntimeError(f"{model_dir} is missing and huggingface_hub is not installed") from e
    print(f"Fetching model from HuggingFace Hub: {link}")
    os.makedirs(model_dir, exist_ok=True)
    snapshot_download(repo_id=link, local_dir=model_dir)
    print(f"Model downloaded to {model_dir}")


def load_HF_model(architecture: str, multi_modal: bool, pretrain: str):
    """Build the model named by ``train_config.json`` and load ``pytorch_model.bin`` (from_HF.py:59-81)."""
    from . import architectures
    model_dir = get_local_model_dir(architecture, multi_modal, pretrain)
    required = ["pytorch_model.bin", "train_config.json"]
    if not all(os.path.isfile(os.path.join(model_dir, f)) for f in required):
        print("Model files not present; downloading model...")
        download_HF_model(architecture, multi_modal, pretrain)
    with open(os.path.join(model_dir, "train_config.json"), "r") as f:
        config = json.load(f)
    model_type = getattr(architectures, config["model_name"])
    model = model_type(config).to(device)
    model.load_state_dict(torch.load(os.path.join(model_dir, "pytorch_model.bin"), map_location=torch.device("cpu")))
    return model
