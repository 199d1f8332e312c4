#!/bin/bash
# round 2, GPU call J: tcgen05 cross-attention (resident context): kernel tests (under a short timeout), engine tests, A/B probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "cross_attention_tcgen05" > gpurun_out/j_kernels.log 2>&1
rc=$?; echo "kernels exit $rc" > gpurun_out/j_box.txt
if [ $rc -ne 0 ]; then tail -40 gpurun_out/j_kernels.log; export FYC_CROSS_TC=0; echo "cross tc OFF for the rest" >> gpurun_out/j_box.txt; fi
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_zz_late_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/j_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/j_box.txt
timeout 900 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider -k "unet_forward or pipeline_cfg2" > gpurun_out/j_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/j_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/j_probe.txt 2>&1
FYC_CROSS_TC=0 timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/j_probe_FYC_CROSS_TC_off.txt
timeout 300 python tests/perf_probe.py 2>&1 | head -4 > gpurun_out/j_probe2.txt
FYC_CROSS_TC=0 timeout 300 python tests/perf_probe.py 2>&1 | head -4 > gpurun_out/j_probe2_off.txt
tail -4 gpurun_out/j_kernels.log; tail -3 gpurun_out/j_engine.log; tail -3 gpurun_out/j_parity.log; cat gpurun_out/j_box.txt; head -16 gpurun_out/j_probe.txt; head -8 gpurun_out/j_probe_FYC_CROSS_TC_off.txt; cat gpurun_out/j_probe2.txt gpurun_out/j_probe2_off.txt; grep -E "attention" gpurun_out/j_probe.txt | head -14
