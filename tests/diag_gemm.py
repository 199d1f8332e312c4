"""Diagnostic: per-role cycle counters of the tcgen05 GEMM kernel on the level-0 shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followyourclick_b200 import ops, _lib
lib = _lib.lib()
lib.fyc_debug_tc_counters.argtypes = [ctypes.c_void_p]
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
SHAPES = [(131072, 320, 320, False, False), (131072, 320, 320, True, False), (131072, 2560, 320, False, True), (32768, 640, 640, True, False), (8192, 10240, 1280, False, True)]
if os.environ.get('SHAPE'): SHAPES = [SHAPES[int(os.environ['SHAPE'])]]
for (M, N, K, res, geglu) in SHAPES:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda").bfloat16() if res else None
    for _ in range(3): ops.gemm(A, W, bias=bias, residual=R, geglu=geglu)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(A, W, bias=bias, residual=R, geglu=geglu)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    dbg.zero_(); lib.fyc_debug_tc_counters(dbg.data_ptr()); ops.gemm(A, W, bias=bias, residual=R, geglu=geglu); torch.cuda.synchronize(); lib.fyc_debug_tc_counters(None)
    d = dbg.view(148, 8).double().mean(0).tolist()
    print(f"M={M} N={N} K={K} res={res} geglu={geglu}: {us:.1f} us  {2*M*N*K/us/1e6:.0f} TFLOP/s | "
          f"mma wait-full {d[2]:.0f} wait-tempty {d[3]:.0f} / total {d[4]:.0f} | epi busy w2 {d[5]:.0f} w6 {d[6]:.0f} clk")
