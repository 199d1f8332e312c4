"""Diagnostic (ncu target): one full-size UNet forward at the cfg2 shape inside a cudaProfilerStart/Stop range, after two warm-up forwards.

  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'<kernels>' -c <N> -o gpurun_out/prof python tests/diag_profile.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from followyourclick_b200 import UNet3DConditionModel, ops
from followyourclick_b200.synth import synth_on_device_

variant = os.environ.get("VARIANT", "base")
unet = UNet3DConditionModel(**bench.unet_kwargs(False, variant)).to("cuda").to(torch.bfloat16)
synth_on_device_(unet, seed=0)
F, h, w = 16, 64, 64
cp = unet.input_channel_pad()
x = torch.zeros(2, F, h, w, cp, device="cuda").bfloat16()
x[..., :9] = torch.randn(2, F, h, w, 9, device="cuda").bfloat16()
ctx = torch.randn(2, 77, 768, device="cuda")
t = torch.tensor(501, device="cuda")
fps, flow = torch.tensor([2, 2], device="cuda"), torch.tensor([4, 4], device="cuda")
kw = dict(fps_tensor=fps, flow_control=flow, use_fps_condition=True)
if variant == "ip16":
    tokens = torch.randn(2, 16, 768, device="cuda")
    context = unet.prepare_context(ctx, None, True, ip_tokens=tokens)
    kw.update(use_ip_cross_attention=True)
else:
    context = unet.prepare_context(ctx, None, False)
run = lambda: unet.forward_nfhwc(x, t, ctx, context=context, **kw)
for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", bool(torch.isfinite(y).all()))
