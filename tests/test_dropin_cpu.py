"""CPU: the drop-in shim packages shadow the hot-path classes and fall through to a reference tree for the rest."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_resolution(tmp_path):
    ref = tmp_path / "ref"
    (ref / "animatediff" / "utils").mkdir(parents=True)
    (ref / "animatediff" / "__init__.py").write_text("")
    (ref / "animatediff" / "utils" / "__init__.py").write_text("")
    (ref / "animatediff" / "utils" / "util.py").write_text("def save_videos_grid():\n    return 'reference'\n")
    (ref / "animatediff" / "models").mkdir()
    (ref / "animatediff" / "models" / "__init__.py").write_text("")
    (ref / "animatediff" / "models" / "unet.py").write_text("class UNet3DConditionModel: origin = 'reference'\n")
    # the reference's vendored diffusers, in miniature: an eager top-level __init__ (never executed by the drop-in), a pipeline family
    # sub-package and utils.import_utils - the places scripts/inference.py:23-34 reaches into
    (ref / "diffusers" / "pipelines" / "stable_diffusion").mkdir(parents=True)
    (ref / "diffusers" / "utils").mkdir()
    (ref / "diffusers" / "__init__.py").write_text("raise ImportError('the reference top-level __init__ must not run')\n")
    (ref / "diffusers" / "pipelines" / "__init__.py").write_text("raise ImportError('eager pipelines/__init__ must not run')\n")
    (ref / "diffusers" / "pipelines" / "stable_diffusion" / "__init__.py").write_text("class StableDiffusionPipeline: origin = 'reference'\n")
    (ref / "diffusers" / "utils" / "__init__.py").write_text("")
    (ref / "diffusers" / "utils" / "import_utils.py").write_text("def is_xformers_available():\n    return 'reference'\n")
    code = ("from animatediff.models.unet import UNet3DConditionModel as U\n"
            "from animatediff.pipelines.pipeline_animation import AnimationPipeline as P\n"
            "from animatediff.utils.util import save_videos_grid\n"
            "from diffusers import AutoencoderKL, DDIMScheduler, StableDiffusionPipeline\n"
            "from diffusers.models import UNet2DConditionModel\n"
            "from diffusers.pipelines import StableDiffusionPipeline as SDP2\n"
            "from diffusers.utils.import_utils import is_xformers_available\n"
            "assert SDP2 is StableDiffusionPipeline and is_xformers_available() == 'reference'\n"
            "assert UNet2DConditionModel.__module__ == 'followyourclick_b200.unet'\n"
            "from ip_adapter import MyIPAdapter, MyIPAdapterPlus\n"
            "from ip_adapter.resampler import Resampler\n"
            "assert MyIPAdapterPlus is not MyIPAdapter and Resampler.__module__ == 'followyourclick_b200.ip_adapter'\n"
            "assert U.__module__ == 'followyourclick_b200.unet' and P.__module__ == 'followyourclick_b200.pipeline_animation'\n"
            "assert save_videos_grid() == 'reference' and StableDiffusionPipeline.origin == 'reference'\n"
            "assert AutoencoderKL.__module__ == 'followyourclick_b200.vae' and DDIMScheduler.__module__ == 'followyourclick_b200.scheduling_ddim'\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "followyourclick_b200", "dropin"), ROOT, str(ref)]))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


REF = "/root/reference"

# Environment shims for THIS container only: packages the reference imports that are not installed here (imageio,
# omegaconf, xformers) and names newer huggingface_hub / transformers releases dropped.  None of them is product code; on a box with
# the reference's own pinned environment they are not needed.
_ENV_STUBS = r'''
import importlib.machinery, sys, types
for n in ("imageio", "omegaconf", "xformers", "xformers.ops"):
    if n not in sys.modules:
        m = types.ModuleType(n); m.__spec__ = importlib.machinery.ModuleSpec(n, None); m.__path__ = []
        sys.modules[n] = m
sys.modules["omegaconf"].OmegaConf = type("OmegaConf", (), {})
import huggingface_hub as hh
for n in ("HfFolder", "cached_download"):
    if not hasattr(hh, n): setattr(hh, n, object)
import transformers
_fe = transformers.CLIPImageProcessor              # (resolving a lazy attribute may re-register sys.modules["transformers"])
for _m in (transformers, sys.modules["transformers"]):
    for _n in ("CLIPFeatureExtractor", "DPTFeatureExtractor"):       # *FeatureExtractor aliases removed in transformers 5
        if _n not in _m.__dict__: setattr(_m, _n, _fe)
'''


def _run_mounted(code, timeout=600):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "followyourclick_b200", "dropin"), ROOT, REF]))
    return subprocess.run([sys.executable, "-c", _ENV_STUBS + code], env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def _need_reference():
    import pytest
    if not os.path.isdir(os.path.join(REF, "animatediff")):
        pytest.skip("/root/reference is only present in the build container")


def test_inference_script_import_block_resolves_against_the_real_tree():
    """The import block of the UNMODIFIED scripts/inference.py (:1-42), executed with PYTHONPATH = dropin : repo : /root/reference:
    the five hot-path names bind to the engine, everything else (utils, converters, stock diffusers pipelines / schedulers /
    import_utils, ip_adapter's other classes) to the reference's own files."""
    _need_reference()
    code = r'''
import os
src = open("/root/reference/scripts/inference.py").read().split("\n")
end = max(i for i, l in enumerate(src[:60]) if l.startswith("from transformers import CLIPVisionModelWithProjection"))
ns = {}
exec(compile("\n".join(src[:end + 1]), "inference_imports", "exec"), ns)
eng = {"UNet3DConditionModel": "followyourclick_b200.unet", "AnimationPipeline": "followyourclick_b200.pipeline_animation",
       "AutoencoderKL": "followyourclick_b200.vae", "DDIMScheduler": "followyourclick_b200.scheduling_ddim",
       "UNet2DConditionModel": "followyourclick_b200.unet", "MyIPAdapter": "followyourclick_b200.ip_adapter",
       "MyIPAdapterPlus": "followyourclick_b200.ip_adapter"}
for name, mod in eng.items():
    assert ns[name].__module__ == mod, (name, ns[name].__module__)
import inspect
ref = {"save_videos_grid": "animatediff/utils/util.py", "convert_ldm_unet_checkpoint": "animatediff/utils/convert_from_ckpt.py",
       "convert_ldm_vae_checkpoint": "animatediff/utils/convert_from_ckpt.py", "convert_lora": "animatediff/utils/convert_lora_safetensor_to_diffusers.py",
       "is_xformers_available": "diffusers/utils/import_utils.py", "StableDiffusionPipeline": "diffusers/pipelines/stable_diffusion/pipeline_stable_diffusion.py"}
for name, f in ref.items():
    got = os.path.realpath(inspect.getsourcefile(ns[name]))
    assert got == os.path.realpath(os.path.join("/root/reference", f)), (name, got)
import diffusers
assert diffusers.__version__ == "0.11.1" and diffusers.DiffusionPipeline.__module__ == "diffusers.pipeline_utils"
from diffusers.schedulers import PNDMScheduler            # reference scheduler package behind the shim
from diffusers.models.attention import CrossAttention      # reference sub-module behind the shim package
assert "reference" in os.path.realpath(inspect.getsourcefile(CrossAttention))
assert ns["DDIMScheduler"].from_config(dict(num_train_timesteps=10, _class_name="x", skip_prk_steps=True)).config.num_train_timesteps == 10
print("mounted ok")
'''
    r = _run_mounted(code)
    assert r.returncode == 0 and "mounted ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_reference_weight_surgery_runs_unchanged_on_engine_models():
    """SURVEY 8f row 4 (load-time weight surgery): the reference's own converters operate on the engine's models without modification,
    because the parameter tree keeps the reference's module paths and `.weight` parameters -
      * `convert_lora` (convert_lora_safetensor_to_diffusers.py:95-157: kohya-style `lora_unet_<path>.lora_down/up.weight`, walks
        `pipeline.unet.__getattr__(name)` and does `layer.weight.data += alpha * up @ down`, 2-D and 1x1-conv 4-D forms),
      * `convert_motion_lora_ckpt_to_diffusers` (:26-51),
      * `load_weights`' motion-module `load_state_dict(strict=False)` (animatediff/utils/util.py:100-109).
    The packed (bf16 / re-laid-out) copies are built lazily at the first forward, so surgery done at load time is what the kernels see;
    `_pack_version` is checked to change on load_state_dict."""
    _need_reference()
    code = r'''
import torch
from animatediff.utils.convert_lora_safetensor_to_diffusers import convert_lora, convert_motion_lora_ckpt_to_diffusers
from animatediff.models.unet import UNet3DConditionModel
from tests.cfgs import mini_unet_ref_kwargs
from followyourclick_b200.synth import load_synth_
unet = load_synth_(UNet3DConditionModel(**mini_unet_ref_kwargs("base")))
assert UNet3DConditionModel.__module__ == "followyourclick_b200.unet"
pipe = type("P", (), {})(); pipe.unet = unet; pipe.text_encoder = torch.nn.Module()
g = torch.Generator().manual_seed(3)
r = lambda *s: torch.randn(*s, generator=g) * 0.1
sd0 = {k: v.clone() for k, v in unet.state_dict().items()}
# kohya-style image LoRA on an attention projection (2-D) and on proj_in (1x1 conv, 4-D)
lora = {"lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight": r(4, 160),
        "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_up.weight": r(160, 4),
        "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.alpha": torch.tensor(4.0),
        "lora_unet_down_blocks_1_attentions_0_proj_in.lora_down.weight": r(4, 320, 1, 1),
        "lora_unet_down_blocks_1_attentions_0_proj_in.lora_up.weight": r(320, 4, 1, 1)}
convert_lora(pipe, lora, alpha=0.8)
k = "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight"
want = sd0[k] + 0.8 * lora["lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_up.weight"] @ lora["lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight"]
assert torch.allclose(unet.state_dict()[k], want, atol=1e-6), "convert_lora 2-D"
k = "down_blocks.1.attentions.0.proj_in.weight"
up, dn = lora["lora_unet_down_blocks_1_attentions_0_proj_in.lora_up.weight"][:, :, 0, 0], lora["lora_unet_down_blocks_1_attentions_0_proj_in.lora_down.weight"][:, :, 0, 0]
assert torch.allclose(unet.state_dict()[k], sd0[k] + 0.8 * (up @ dn)[:, :, None, None], atol=1e-6), "convert_lora 4-D"
# motion LoRA checkpoint (AnimateDiff motion-LoRA key style)
base = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
mlora = {f"module.{base}.processor.to_out_lora.down.weight": r(4, 160), f"module.{base}.processor.to_out_lora.up.weight": r(160, 4)}
convert_motion_lora_ckpt_to_diffusers(pipe, mlora, alpha=0.5)
k = base + ".to_out.0.weight"
assert torch.allclose(unet.state_dict()[k], sd0[k] + 0.5 * mlora[f"module.{base}.processor.to_out_lora.up.weight"] @ mlora[f"module.{base}.processor.to_out_lora.down.weight"], atol=1e-6), "motion lora"
# motion-module checkpoint load (util.load_weights): only motion_modules.* keys, strict=False, nothing unexpected
v0 = unet._pack_version
mm = {kk: torch.zeros_like(vv) for kk, vv in sd0.items() if "motion_modules." in kk and not kk.endswith(".pe")}
missing, unexpected = unet.load_state_dict(mm, strict=False)
assert len(unexpected) == 0 and len(mm) > 0 and unet._pack_version != v0
assert float(unet.state_dict()[base + ".to_q.weight"].abs().max()) == 0.0
print("surgery ok")
'''
    r = _run_mounted(code)
    assert r.returncode == 0 and "surgery ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
