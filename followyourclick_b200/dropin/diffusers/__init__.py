"""Shadows the `diffusers` names scripts/inference.py binds (`from diffusers import AutoencoderKL, DDIMScheduler`,
scripts/inference.py:23-36).  Any other attribute (StableDiffusionPipeline, ...) is resolved lazily from the reference's
vendored diffusers 0.11.1, loaded under the private name `_fyc_ref_diffusers` from the next `diffusers/` directory on
sys.path, so converters and the optional T2I first-frame pipeline keep working unchanged."""
import importlib.util
import os
import sys

from followyourclick_b200.scheduling_ddim import DDIMScheduler  # noqa: F401
from followyourclick_b200.vae import AutoencoderKL  # noqa: F401
from followyourclick_b200.unet import UNet2DConditionModel  # noqa: F401  (T2I first-frame generator, scripts/inference.py:195-204)

__version__ = "0.11.1"
_ref = None


def _reference_package():
    global _ref
    if _ref is None:
        here = os.path.dirname(os.path.abspath(__file__))
        for p in sys.path:
            cand = os.path.join(p, "diffusers", "__init__.py")
            if os.path.isfile(cand) and os.path.dirname(os.path.abspath(cand)) != here:
                spec = importlib.util.spec_from_file_location("_fyc_ref_diffusers", cand, submodule_search_locations=[os.path.dirname(cand)])
                mod = importlib.util.module_from_spec(spec)
                sys.modules["_fyc_ref_diffusers"] = mod
                spec.loader.exec_module(mod)
                _ref = mod
                break
        else:
            raise ImportError("no reference `diffusers` package found after the drop-in on sys.path")
    return _ref


def __getattr__(name):
    return getattr(_reference_package(), name)
