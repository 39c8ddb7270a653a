"""btsbot_b200 -- B200 (sm_100a) implementation of BTSbot's alert-scoring hot path behind the ``btsbot`` API.

Exports mirror `btsbot/__init__.py:9-46`.  ``import btsbot_b200 as btsbot`` (or :func:`install_as_btsbot`) makes
reference-style code -- ``btsbot.load_HF_model``, ``btsbot.architectures.mm_ConvNeXt(config)``,
``btsbot.FlexibleDataset`` -- run on the hand-written CUDA kernels in ``libbtsbot_b200.so``.
"""
__version__ = "2.0.6+b200.1"

from . import architectures, utils, alert_utils, from_HF, to_HF, synth  # noqa: F401
from .utils import FlexibleDataset, RandomRightAngleRotation, make_report  # noqa: F401
from .architectures import (  # noqa: F401
    MaxViT, ConvNeXt, mm_MaxViT, mm_ConvNeXt, mm_cnn, um_cnn, um_nn, frozen_fusion,
)
from .from_HF import download_HF_model, load_HF_model  # noqa: F401

__all__ = [
    "__version__", "architectures", "utils", "alert_utils", "FlexibleDataset", "RandomRightAngleRotation",
    "make_report", "MaxViT", "ConvNeXt", "mm_MaxViT", "mm_ConvNeXt", "mm_cnn", "um_cnn", "um_nn",
    "frozen_fusion", "download_HF_model", "load_HF_model",
]


def install_as_btsbot():
    """Register this package under the name ``btsbot`` so unmodified reference-style scripts import it."""
    import sys
    sys.modules["btsbot"] = sys.modules[__name__]
    for sub in ("architectures", "utils", "alert_utils", "from_HF", "to_HF"):
        sys.modules["btsbot." + sub] = sys.modules[__name__ + "." + sub]
    return sys.modules[__name__]
