#!/bin/bash
# round 2, GPU call D: LN fold with the k-step skip, attention diagnostics (what bounds attention_tc2?), probes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo start > gpurun_out/d_box.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "layernorm_fold or two_segment or gemm_bias_residual or geglu or cta_pairs or conv3x3 or conv_head or self_attention_tcgen05" > gpurun_out/d_newkernels.log 2>&1
echo "new kernels exit $?" >> gpurun_out/d_box.txt
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/d_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/d_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/d_probe.txt 2>&1
FYC_LN_FOLD=0 timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/d_probe_FYC_LN_FOLD_off.txt
( for d in 0 1 2 3; do echo "FYC_ATTN_DBG=$d"; FYC_ATTN_DBG=$d timeout 120 python tests/diag_attn.py; done ) > gpurun_out/d_attn_dbg.txt 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench exit $?" >> gpurun_out/d_box.txt
tail -3 gpurun_out/d_newkernels.log; tail -3 gpurun_out/d_engine.log; cat gpurun_out/d_box.txt; head -16 gpurun_out/d_probe.txt; sed -n 2,4p gpurun_out/d_probe_FYC_LN_FOLD_off.txt; cat gpurun_out/d_attn_dbg.txt; grep -E "gL\]|L\]" gpurun_out/d_probe.txt | head
