"""`diffusers.pipelines` of the drop-in (scripts/inference.py:25 `from diffusers.pipelines import StableDiffusionPipeline`).  The
engine replaces no stock pipeline: every name is the reference's, imported lazily from the sub-package that defines it
(`diffusers/pipelines/<family>/`, reached through pkgutil.extend_path) instead of through the reference's eager
`pipelines/__init__.py`, which imports all twenty pipeline families at once."""
import importlib
import os
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

_FAMILY_BY_PREFIX = (("StableDiffusionPipelineSafe", "stable_diffusion_safe"), ("OnnxStableDiffusion", "stable_diffusion"),
                     ("FlaxStableDiffusion", "stable_diffusion"), ("StableDiffusion", "stable_diffusion"), ("CycleDiffusion", "stable_diffusion"),
                     ("AltDiffusion", "alt_diffusion"), ("VersatileDiffusion", "versatile_diffusion"), ("VQDiffusion", "vq_diffusion"),
                     ("UnCLIP", "unclip"), ("PaintByExample", "paint_by_example"), ("LDMTextToImage", "latent_diffusion"),
                     ("LDMSuperResolution", "latent_diffusion"), ("LDMPipeline", "latent_diffusion_uncond"), ("DDIMPipeline", "ddim"),
                     ("DDPMPipeline", "ddpm"), ("PNDMPipeline", "pndm"), ("RePaint", "repaint"), ("ScoreSdeVe", "score_sde_ve"),
                     ("KarrasVe", "stochastic_karras_ve"), ("DanceDiffusion", "dance_diffusion"), ("AudioDiffusion", "audio_diffusion"),
                     ("Mel", "audio_diffusion"))


def _families():
    seen = []
    for d in __path__:
        if os.path.isdir(d):
            for n in sorted(os.listdir(d)):
                if os.path.isfile(os.path.join(d, n, "__init__.py")) and n not in seen:
                    seen.append(n)
    return seen


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    tried = [fam for prefix, fam in _FAMILY_BY_PREFIX if name.startswith(prefix)]
    for fam in tried + [f for f in _families() if f not in tried]:
        try:
            mod = importlib.import_module(f"{__name__}.{fam}")
        except ImportError:
            if fam in tried:
                raise
            continue
        if hasattr(mod, name):
            return getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
