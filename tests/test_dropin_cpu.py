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
    (ref / "diffusers").mkdir()
    (ref / "diffusers" / "__init__.py").write_text("from .sub import StableDiffusionPipeline\nAutoencoderKL = 'reference'\n")
    (ref / "diffusers" / "sub.py").write_text("class StableDiffusionPipeline: origin = 'reference'\n")
    code = ("from animatediff.models.unet import UNet3DConditionModel as U\n"
            "from animatediff.pipelines.pipeline_animation import AnimationPipeline as P\n"
            "from animatediff.utils.util import save_videos_grid\n"
            "from diffusers import AutoencoderKL, DDIMScheduler, StableDiffusionPipeline\n"
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
