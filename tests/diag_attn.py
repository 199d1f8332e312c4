"""Diagnostic: the level-0 spatial self-attention (32 frames x 8 heads x 4096 tokens, D = 40) on the tcgen05 kernel; ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followyourclick_b200 import ops
NB, L, heads, D = int(os.environ.get("NB", 32)), 4096, 8, 40
C = heads * D
qk = torch.zeros(NB, L, 2 * heads * 64 + C, dtype=torch.bfloat16, device="cuda")
qk[:, :, :heads * 64].view(NB, L, heads, 64)[..., :D] = torch.randn(NB, L, heads, D, device="cuda")
qk[:, :, heads * 64:2 * heads * 64].view(NB, L, heads, 64)[..., :D] = torch.randn(NB, L, heads, D, device="cuda")
qk[:, :, 2 * heads * 64:] = torch.randn(NB, L, C, device="cuda")
vt = ops.transpose_tokens(qk, 2 * heads * 64, C)
fn = lambda: ops.self_attention_tc(qk, 0, heads * 64, vt, heads, D, D ** -0.5)
for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"self_attention_tc NB={NB} L={L}: {ms:.3f} ms  {4.0 * NB * heads * L * L * D / ms / 1e9:.0f} TFLOP/s (D=40 useful flops)")
