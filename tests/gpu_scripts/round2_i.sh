#!/bin/bash
# round 2, GPU call I: GroupNorm apply kernel (round-robin row blocks, double-buffered loads, tanh SiLU): tests + per-shape probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "groupnorm" > gpurun_out/i_kernels.log 2>&1
echo "kernels exit $?" > gpurun_out/i_box.txt
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -p no:cacheprovider -k "unet or vae" > gpurun_out/i_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/i_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/i_probe.txt 2>&1
tail -3 gpurun_out/i_kernels.log; tail -3 gpurun_out/i_engine.log; cat gpurun_out/i_box.txt; head -16 gpurun_out/i_probe.txt; grep groupnorm gpurun_out/i_probe.txt
