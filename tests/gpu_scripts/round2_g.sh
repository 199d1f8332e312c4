#!/bin/bash
# round 2, GPU call G: head-dim-80 tcgen05 self-attention, per-replica skip reads; tests + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo start > gpurun_out/g_box.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "d80 or self_attention_tcgen05 or two_sources" > gpurun_out/g_newkernels.log 2>&1
echo "new kernels exit $?" >> gpurun_out/g_box.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_zz_late_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/g_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/g_box.txt
timeout 900 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider -k "unet_forward or pipeline or shared" > gpurun_out/g_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/g_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/g_probe.txt 2>&1
FYC_ATTN_D80=0 timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/g_probe_FYC_ATTN_D80_off.txt
timeout 300 python tests/perf_probe.py 2>&1 | head -4 > gpurun_out/g_probe2.txt
FYC_ATTN_D80=0 timeout 300 python tests/perf_probe.py 2>&1 | head -4 > gpurun_out/g_probe2_off.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?" >> gpurun_out/g_box.txt
tail -3 gpurun_out/g_newkernels.log; tail -3 gpurun_out/g_engine.log; tail -3 gpurun_out/g_parity.log; cat gpurun_out/g_box.txt; head -16 gpurun_out/g_probe.txt; head -8 gpurun_out/g_probe_FYC_ATTN_D80_off.txt; cat gpurun_out/g_probe2.txt gpurun_out/g_probe2_off.txt; grep attention gpurun_out/g_probe.txt; head -c 300 gpurun_out/g_bench.json
