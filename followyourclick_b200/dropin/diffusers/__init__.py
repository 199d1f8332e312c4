"""Shadows the `diffusers` names scripts/inference.py binds (`from diffusers import AutoencoderKL, DDIMScheduler`,
scripts/inference.py:23-36) with the engine's classes.  Everything else stays the reference's vendored diffusers 0.11.1:

* sub-modules and sub-packages (`diffusers.utils.import_utils`, `diffusers.schedulers`, `diffusers.models.attention`, ...) resolve
  through `pkgutil.extend_path` to the next `diffusers/` directory on sys.path (the reference tree), so
  `from diffusers.utils.import_utils import is_xformers_available` (scripts/inference.py:34) and the converters' imports
  (`animatediff/utils/convert_from_ckpt.py:34-50`) work unchanged; `diffusers.models` / `diffusers.pipelines` are shim packages of
  the same kind (engine classes first, the reference's own sub-modules behind them);
* other top-level attributes (`StableDiffusionPipeline`, `DiffusionPipeline`, `ModelMixin`, ...) are looked up lazily in the
  reference sub-packages that define them - the reference's top-level `__init__` (which eagerly imports every pipeline) is never run.
"""
import importlib
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
__version__ = "0.11.1"

from followyourclick_b200.scheduling_ddim import DDIMScheduler  # noqa: E402,F401
from followyourclick_b200.unet import UNet2DConditionModel  # noqa: E402,F401  (T2I first-frame generator, scripts/inference.py:195-204)
from followyourclick_b200.vae import AutoencoderKL  # noqa: E402,F401

# where the reference's top-level names live (diffusers/__init__.py of 0.11.1), tried in this order for any other attribute
_LAZY_HOMES = ("diffusers.pipelines", "diffusers.models", "diffusers.schedulers", "diffusers.pipeline_utils", "diffusers.modeling_utils",
               "diffusers.configuration_utils", "diffusers.optimization", "diffusers.training_utils", "diffusers.utils")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    errors = []
    for home in _LAZY_HOMES:
        try:
            mod = importlib.import_module(home)
        except ImportError as e:        # a reference sub-package whose own dependencies are missing in this environment
            errors.append(f"{home}: {e}")
            continue
        try:
            return getattr(mod, name)
        except AttributeError:
            continue
    raise AttributeError(f"module 'diffusers' (followyourclick_b200 drop-in over the reference's diffusers 0.11.1) has no attribute {name!r}"
                         + (f" [{'; '.join(errors)}]" if errors else ""))
