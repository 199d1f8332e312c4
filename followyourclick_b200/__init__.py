"""followyourclick_b200: B200-native (sm_100a) engine for the FollowYourClick denoising hot path.

Public surface = the reference's own class names; see INTEGRATION.md for how
``scripts/inference.py`` mounts them.  Everything below the Python classes is libfyc_sm100a.so (include/fyc.h).
"""
from .unet import UNet2DConditionModel, UNet3DConditionModel, UNet3DConditionOutput, ImageProjModel  # noqa: F401
from .vae import AutoencoderKL  # noqa: F401
from .scheduling_ddim import DDIMScheduler  # noqa: F401
from .pipeline_animation import AnimationPipeline, AnimationPipelineOutput  # noqa: F401
from .ip_adapter import IPAttnProcessor, IPAttnProcessor2_0, MyIPAdapter, MyIPAdapterPlus, Resampler  # noqa: F401

__version__ = "0.1.0"
