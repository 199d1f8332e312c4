#!/bin/bash
# round 2, GPU call E: LN fold with 128-byte aligned augmented weights: kernel tests + A/B probe (same box, back to back, twice)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "layernorm_fold or two_segment" > gpurun_out/e_newkernels.log 2>&1
echo "new kernels exit $?" > gpurun_out/e_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/e_probe.txt 2>&1
SHAPES=1 FYC_LN_FOLD=0 timeout 300 python tests/perf_probe.py > gpurun_out/e_probe_FYC_LN_FOLD_off.txt 2>&1
timeout 300 python tests/perf_probe.py 2>&1 | head -4 > gpurun_out/e_probe2.txt
FYC_LN_FOLD=0 timeout 300 python tests/perf_probe.py 2>&1 | head -4 > gpurun_out/e_probe2_off.txt
tail -3 gpurun_out/e_newkernels.log; cat gpurun_out/e_box.txt; head -16 gpurun_out/e_probe.txt; head -16 gpurun_out/e_probe_FYC_LN_FOLD_off.txt; cat gpurun_out/e_probe2.txt gpurun_out/e_probe2_off.txt
