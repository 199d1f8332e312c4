"""Diagnostic (not a test): layer-wise bf16-vs-fp32 error growth of the engine on the mini UNet, and sensitivity of the
synthetic network to tiny input perturbations (conditioning of the test problem)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.cfgs import unet_inputs
from tests.engine_helpers import make_unet, unet_forward_kwargs, stats
import followyourclick_b200.synth as synth

for gain in (2.0, 1.0):
    synth._QK_GAIN = gain
    for variant in ("base", "ip"):
        inp = unet_inputs(variant)
        outs, taps = {}, {}
        for dt in (torch.float32, torch.bfloat16):
            unet, _ = make_unet(variant, dt)
            unet._taps = {}
            outs[dt] = unet(inp["sample"].cuda(), inp["timestep"], **unet_forward_kwargs(variant, inp, "cuda")).sample
            taps[dt] = unet._taps
        print(f"gain={gain} variant={variant} final bf16-vs-fp32:", stats(outs[torch.bfloat16], outs[torch.float32]))
        for k in taps[torch.float32]:
            s = stats(taps[torch.bfloat16][k], taps[torch.float32][k])
            print(f"   {k:8s} rel_l2={s['rel_l2']:.4f} ref_max={s['ref_max']:.2f}")
        # conditioning: perturb the input by 1e-4 relative in fp32 mode
        unet, _ = make_unet(variant, torch.float32)
        x = inp["sample"].cuda()
        a = unet(x, inp["timestep"], **unet_forward_kwargs(variant, inp, "cuda")).sample
        b = unet(x * (1 + 1e-4 * torch.randn_like(x)), inp["timestep"], **unet_forward_kwargs(variant, inp, "cuda")).sample
        print("   fp32 sensitivity: input rel perturbation 1e-4 -> output rel_l2", stats(b, a)["rel_l2"])
