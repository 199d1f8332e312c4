"""Diagnostic: tcgen05 conv / GEMM role counters with CTA pairs on or off (FYC_TC_PAIR=0/1 in the environment)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followyourclick_b200 import ops, _lib
lib = _lib.lib()
lib.fyc_debug_tc_counters.argtypes = [ctypes.c_void_p]
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")

def run(name, fn, flops):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    dbg.zero_(); lib.fyc_debug_tc_counters(dbg.data_ptr()); fn(); torch.cuda.synchronize(); lib.fyc_debug_tc_counters(None)
    d = dbg.view(148, 8).double()
    lead = d[d[:, 4] > 0]
    m = lead.mean(0).tolist() if len(lead) else [0] * 8
    e = d[:, 5].mean().item()
    print(f"{name}: {us:.1f} us {flops/us/1e6:.0f} TFLOP/s | issuers {len(lead)} wait-full {m[2]:.0f} wait-tempty {m[3]:.0f} total {m[4]:.0f} | epi busy {e:.0f}")

for (NB, H, Cin, Cout) in [(32, 64, 320, 320), (32, 64, 640, 640), (32, 32, 640, 640), (32, 32, 1280, 1280)]:
    x = torch.randn(NB, H, H, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, 3, 3, Cin, device="cuda") * (9 * Cin) ** -0.5).bfloat16()
    b = torch.randn(Cout, device="cuda")
    run(f"conv {NB}x{H}x{H} {Cin}->{Cout}", lambda: ops.conv3x3(x, w, bias=b), 2 * NB * H * H * Cout * 9 * Cin)
for (M, N, K) in [(32768, 640, 2560), (131072, 320, 1280), (8192, 1280, 1280)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda").bfloat16()
    run(f"gemm {M}x{N}x{K}r", lambda: ops.gemm(A, W, bias=bias, residual=R), 2 * M * N * K)
