"""Diagnostic (not a test): full-size (SD-1.5 + motion modules, 1.28 B params) UNet forward and VAE decode timing on one
GPU, with a per-kernel-family breakdown from CUDA events around every C-ABI call."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followyourclick_b200 import UNet3DConditionModel, AutoencoderKL, ops
from followyourclick_b200.synth import synth_tensor
from oracle.ref_unet import default_unet_config

def gpu_synth_(model, seed=0):
    """fast on-device synthetic weights (values differ from the CPU generator; only used for timing)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    for k, p in model.state_dict(keep_vars=True).items():
        if k.endswith(".pe"): continue
        if p.dim() == 1:
            p.data = (0.05 * torch.randn(p.shape, generator=g, device="cuda")) if k.endswith("bias") else (1 + 0.1 * torch.randn(p.shape, generator=g, device="cuda"))
        else:
            fan_in = p[0].numel()
            p.data = torch.randn(p.shape, generator=g, device="cuda") * fan_in ** -0.5
    model._invalidate()

def main():
    F, h, w = int(os.environ.get("F", 16)), int(os.environ.get("HW", 64)), int(os.environ.get("HW", 64))
    kw = dict(sample_size=64, in_channels=4, out_channels=4, cross_attention_dim=768, attention_head_dim=8,
              use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), unet_use_cross_frame_attention=False,
              unet_use_temporal_attention=False, motion_module_type="Vanilla", use_fps_condition=True,
              use_first_frame_mask_condition_concat=True,
              motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
                                        temporal_position_encoding=True, temporal_position_encoding_max_len=max(24, F), temporal_attention_dim_div=1))
    t0 = time.time()
    unet = UNet3DConditionModel(**kw).to("cuda").to(torch.bfloat16)
    gpu_synth_(unet)
    print("unet build+synth s:", time.time() - t0, "params(M):", sum(p.numel() for p in unet.parameters()) / 1e6)
    cp = unet.input_channel_pad()                       # 16 in tensor-core mode: the 9-channel stem runs on tcgen05 (as in the pipeline)
    x = torch.zeros(2, F, h, w, cp, device="cuda").bfloat16()
    x[..., :9] = torch.randn(2, F, h, w, 9, device="cuda").bfloat16()
    ctx = torch.randn(2, 77, 768, device="cuda")
    t = torch.tensor(501, device="cuda")
    fps, flow = torch.tensor([2, 2], device="cuda"), torch.tensor([4, 4], device="cuda")
    run = lambda: unet.forward_nfhwc(x, t, ctx, fps_tensor=fps, flow_control=flow, use_fps_condition=True)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.time(); e0.record()
    for _ in range(3): y = run()
    e1.record(); host = (time.time() - h0) / 3; torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"UNet fwd F={F} {h}x{w}: {ms:.2f} ms (host enqueue {host*1e3:.1f} ms)  finite={bool(torch.isfinite(y).all())}")
    with ops.profile() as prof: run()
    tot = sum(d["ms"] for d in prof.summary.values())
    for fam, d in sorted(prof.summary.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {fam:20s} {d['ms']:8.2f} ms {100*d['ms']/tot:5.1f}%  launches {d['launches']:4d}  {d['flops']/d['ms']/1e9 if d['ms'] else 0:8.1f} TFLOP/s  {d['bytes']/d['ms']/1e6 if d['ms'] else 0:8.1f} GB/s")
    print("   sum of event times", tot)
    if os.environ.get("SHAPES"):
        ops._prof_shapes = True
        with ops.profile() as prof2: run()
        ops._prof_shapes = False
        rows = sorted(prof2.summary.items(), key=lambda kv: -kv[1]["ms"])
        for fam, d in rows[:110]:
            if d["ms"]:
                print(f"   {fam:46s} {d['ms']:7.3f} ms x{d['launches']:3d}  {d['flops']/d['ms']/1e9:7.1f} TFLOP/s  {d['bytes']/d['ms']/1e6:7.1f} GB/s(alg)")
    # VAE decode of F frames
    vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                        block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, norm_num_groups=32).to("cuda").to(torch.bfloat16)
    gpu_synth_(vae)
    z = torch.randn(F, h, w, 4, device="cuda").bfloat16()
    for _ in range(2): vae.decode_nhwc(z)
    torch.cuda.synchronize(); e0.record(); fr = vae.decode_nhwc(z); e1.record(); torch.cuda.synchronize()
    print(f"VAE decode {F} frames: {e0.elapsed_time(e1):.2f} ms finite={bool(torch.isfinite(fr).all())}")
    with ops.profile() as prof: vae.decode_nhwc(z)
    for fam, d in sorted(prof.summary.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {fam:20s} {d['ms']:8.2f} ms launches {d['launches']:4d}  {d['flops']/d['ms']/1e9 if d['ms'] else 0:8.1f} TFLOP/s")
    if os.environ.get("SHAPES"):
        ops._prof_shapes = True
        with ops.profile() as prof2: vae.decode_nhwc(z)
        ops._prof_shapes = False
        for fam, d in sorted(prof2.summary.items(), key=lambda kv: -kv[1]["ms"])[:25]:
            if d["ms"]:
                print(f"   {fam:46s} {d['ms']:7.3f} ms x{d['launches']:3d}  {d['flops']/d['ms']/1e9:7.1f} TFLOP/s  {d['bytes']/d['ms']/1e6:7.1f} GB/s(alg)")
    print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)

if __name__ == "__main__":
    main()
