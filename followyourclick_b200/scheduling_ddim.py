"""DDIMScheduler: reference call surface (diffusers/schedulers/scheduling_ddim.py:114-422), CUDA step underneath.

The schedule (betas, zero-terminal-SNR rescale, alphas_cumprod, timesteps) is host-side scalar work done once with
the same fp32 torch operations as the reference, so the coefficient tables are bit-identical; ``step`` is one fused
elementwise kernel (fyc_cfg_ddim_step) that also absorbs the classifier-free-guidance combine when the pipeline
calls ``step_cfg``.  No per-step host sync: coefficients are looked up from host copies by integer timestep.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from . import ops
from .modeling import FrozenDict


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def rescale_zero_terminal_snr(betas):
    """scheduling_ddim.py:78-111 (arXiv 2305.08891 Alg. 1)."""
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    first, last = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt -= last
    alphas_bar_sqrt *= first / (first - last)
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
    return 1 - alphas


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                 prediction_type="epsilon", rescale_betas_zero_snr=False, **kwargs):
        if prediction_type not in L.PRED:
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or `v_prediction`")
        self.config = FrozenDict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                 beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                                 set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                 prediction_type=prediction_type, rescale_betas_zero_snr=rescale_betas_zero_snr,
                                 _class_name="DDIMScheduler", _diffusers_version="0.11.1")
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if rescale_betas_zero_snr:
            self.betas = rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._timesteps_host = self.timesteps.tolist()

    # ---- ConfigMixin surface used by the reference scripts (diffusers/configuration_utils.py; scripts/inference.py:199 calls
    # DDIMScheduler.from_pretrained(path, subfolder="scheduler") on the T2I first-frame path)
    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        cfg.update(kwargs)
        accepted = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "trained_betas", "clip_sample", "set_alpha_to_one",
                    "steps_offset", "prediction_type", "rescale_betas_zero_snr")
        return cls(**{k: v for k, v in cfg.items() if k in accepted})       # other schedulers' keys (skip_prk_steps, ...) are ignored like ConfigMixin does

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
        """reads ``scheduler_config.json`` of a local diffusers model folder (no hub access: the engine runs offline)"""
        import json
        import os
        path = pretrained_model_name_or_path if subfolder is None else os.path.join(pretrained_model_name_or_path, subfolder)
        f = path if os.path.isfile(path) else os.path.join(path, "scheduler_config.json")
        if not os.path.isfile(f):
            raise EnvironmentError(f"{f} does not exist (DDIMScheduler.from_pretrained reads a local scheduler_config.json)")
        with open(f) as fh:
            return cls.from_config(json.load(fh), **kwargs)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        timesteps = timesteps + self.config.steps_offset
        self._timesteps_host = [int(t) for t in timesteps]
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def coefs(self, timestep, eta=0.0, guidance=1.0, cfg_pair=None):
        """Scalar coefficients of one step, computed with the reference's fp32 torch expressions (:308-349).  ``cfg_pair``: whether
        the model output holds the [uncond; cond] pair - the caller's own ``guidance_scale > 1.0`` decision on the Python double
        (pipeline_animation.py:599), never re-derived from the fp32-rounded scale inside the kernel."""
        cfg_pair = (guidance > 1.0) if cfg_pair is None else cfg_pair
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        variance = ((1 - a_prev) / b_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5
        return L.DdimCoefs(float(guidance), float(a_t ** 0.5), float(b_t ** 0.5), float(a_prev ** 0.5), float(direction),
                           float(variance ** 0.5 * eta), L.PRED[self.config.prediction_type], int(bool(self.config.clip_sample)),
                           int(bool(cfg_pair)))

    def _noise(self, shape, eta, generator, variance_noise, device):
        if eta <= 0:
            return None
        if variance_noise is not None and generator is not None:
            raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                             " `variance_noise` stays `None`.")
        if variance_noise is None:
            variance_noise = torch.randn(shape, generator=generator, device=device, dtype=torch.float32)
        return variance_noise.to(device=device, dtype=torch.float32).contiguous()

    @torch.no_grad()
    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict=True):
        """scheduling_ddim.py:254-376.  ``timestep`` may be a python int or a tensor element (a CUDA element costs one
        sync, exactly like the reference; the pipeline passes host ints)."""
        if use_clipped_model_output:
            raise NotImplementedError("use_clipped_model_output")
        c = self.coefs(timestep, eta, 1.0)
        x = sample.to(torch.float32).contiguous()
        m = model_output.to(device=x.device, dtype=torch.float32).contiguous()
        prev = ops.cfg_ddim_step(m, x, c, noise=self._noise(x.shape, eta, generator, variance_noise, x.device))
        if not return_dict:
            return (prev,)
        # pred_original_sample (:318-330), as the reference returns it: the same kernel with sqrt(alpha_prev) := 1 and the direction
        # coefficient := 0 leaves exactly x0 (1 * x0 + 0 * eps).  Only this public step() pays for it; the pipeline's fused step_cfg does not.
        c0 = L.DdimCoefs(c.guidance, c.sqrt_alpha_t, c.sqrt_beta_t, 1.0, 0.0, 0.0, c.prediction_type, c.clip_sample, c.cfg_pair)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=ops.cfg_ddim_step(m, x, c0))

    @torch.no_grad()
    def step_cfg(self, model_output_pair, timestep, sample, guidance_scale, eta=0.0, generator=None, variance_noise=None,
                 single_frame_output=None, video_scale=0.0):
        """CFG combine (pipeline_animation.py:763-764) fused with the step: model_output_pair = [uncond; cond].  With
        ``single_frame_output`` (the per-frame prediction, :738-755) the combine is the video_scale form of :757-761."""
        pair = model_output_pair.numel() == 2 * sample.numel()       # the caller's do_cfg decision, visible in the shapes it passes
        c = self.coefs(timestep, eta, guidance_scale, cfg_pair=pair)
        return ops.cfg_ddim_step(model_output_pair, sample, c,
                                 noise=self._noise(sample.shape, eta, generator, variance_noise, sample.device),
                                 single=single_frame_output, video_scale=video_scale)

    def add_noise(self, original_samples, noise, timesteps):
        """scheduling_ddim.py:378-398 (training/inversion helper, not on the sampling path; plain torch)."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def __len__(self):
        return self.config.num_train_timesteps
