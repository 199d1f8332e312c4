"""Oracle: fp32 restatement of ``DDIMScheduler`` (init, set_timesteps, step) and the CFG combine.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: diffusers/schedulers/scheduling_ddim.py
:78-111 (rescale_zero_terminal_snr), :156-212 (__init__), :238-252 (set_timesteps), :254-376 (step);
CFG combine animatediff/pipelines/pipeline_animation.py:763-764.
Arithmetic is done on fp32 torch CPU tensors in the same operation order as the reference so the
known-answer values of SURVEY App. D reproduce bit-for-bit.
"""
import numpy as np
import torch


def default_scheduler_config(**over):
    """configs/inference/inference_img_embed_mask_condition_zero_snr_.yaml:18-26."""
    cfg = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
               steps_offset=1, clip_sample=False, prediction_type="v_prediction",
               rescale_betas_zero_snr=True, set_alpha_to_one=True)
    cfg.update(over)
    return cfg


def make_betas(cfg):
    n = cfg["num_train_timesteps"]
    if cfg["beta_schedule"] == "linear":
        betas = torch.linspace(cfg["beta_start"], cfg["beta_end"], n, dtype=torch.float32)
    elif cfg["beta_schedule"] == "scaled_linear":
        betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, n, dtype=torch.float32) ** 2
    else:
        raise NotImplementedError(cfg["beta_schedule"])
    if cfg["rescale_betas_zero_snr"]:                      # scheduling_ddim.py:90-111
        abar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
        s0, sT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
        abar_sqrt -= sT
        abar_sqrt *= s0 / (s0 - sT)
        abar = abar_sqrt ** 2
        alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
        betas = 1 - alphas
    return betas


class DDIMOracle:
    def __init__(self, cfg=None):
        self.cfg = cfg or default_scheduler_config()
        self.betas = make_betas(self.cfg)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if self.cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.cfg["num_train_timesteps"] // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.cfg["steps_offset"]
        return self.timesteps

    def step(self, model_output, timestep, sample, eta=0.0, variance_noise=None):
        if self.num_inference_steps is None:
            raise ValueError("set_timesteps first")
        t = int(timestep)
        prev_t = t - self.cfg["num_train_timesteps"] // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        pt = self.cfg["prediction_type"]
        if pt == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif pt == "sample":
            x0 = model_output
        elif pt == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            model_output = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(pt)
        if self.cfg["clip_sample"]:
            x0 = torch.clamp(x0, -1, 1)
        variance = ((1 - a_prev) / b_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * model_output
        prev = a_prev ** 0.5 * x0 + direction
        if eta > 0:
            prev = prev + variance ** 0.5 * eta * variance_noise
        return prev


def cfg_combine(noise_pred, guidance_scale):
    """pipeline_animation.py:763-764 (batch order [uncond, cond])."""
    u, c = noise_pred.chunk(2)
    return u + guidance_scale * (c - u)
