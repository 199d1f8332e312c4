"""Diagnostic: LayerNorm / GroupNorm on the level-0 / level-1 activation shapes (timed; also the ncu target for these kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followyourclick_b200 import ops

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for (M, C) in [(131072, 320), (32768, 640), (8192, 1280)]:
    x = torch.randn(M, C, device="cuda").bfloat16(); g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    us = t(lambda: ops.layernorm(x, g, b))
    print(f"layernorm {M}x{C}: {us:.1f} us  {2 * x.numel() * 2 / us / 1e6:.2f} TB/s")
for (NB, R, C) in [(32, 4096, 320), (2, 65536, 320), (32, 1024, 640)]:
    x = torch.randn(NB, R, C, device="cuda").bfloat16(); g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    us = t(lambda: ops.groupnorm(x, g, b, 32, 1e-5, silu=True))
    print(f"groupnorm {NB}x{R}x{C}: {us:.1f} us  {3 * x.numel() * 2 / us / 1e6:.2f} TB/s")
