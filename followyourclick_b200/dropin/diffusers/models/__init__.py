"""`diffusers.models` of the drop-in: the engine's AutoencoderKL / UNet2DConditionModel (scripts/inference.py:24
`from diffusers.models import UNet2DConditionModel`); sub-modules (`diffusers.models.attention`, `.vae`, `.resnet`, ...) and the
remaining class names resolve to the reference's vendored files behind this directory (pkgutil.extend_path)."""
import importlib
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from followyourclick_b200.unet import UNet2DConditionModel  # noqa: E402,F401
from followyourclick_b200.vae import AutoencoderKL  # noqa: E402,F401

_HOMES = {"Transformer2DModel": "attention", "PriorTransformer": "prior_transformer", "UNet1DModel": "unet_1d", "UNet2DModel": "unet_2d",
          "VQModel": "vae", "FlaxUNet2DConditionModel": "unet_2d_condition_flax", "FlaxAutoencoderKL": "vae_flax"}


def __getattr__(name):
    if name in _HOMES:
        return getattr(importlib.import_module(f"{__name__}.{_HOMES[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
