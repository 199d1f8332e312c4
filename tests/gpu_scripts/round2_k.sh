#!/bin/bash
# round 2, GPU call K: cross-attention v2 (pipelined epilogue, wave-aware grid) tests + probe; ncu --set full evidence of the hot kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "cross_attention_tcgen05" > gpurun_out/k_kernels.log 2>&1
rc=$?; echo "kernels exit $rc" > gpurun_out/k_box.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_zz_late_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/k_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/k_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/k_probe.txt 2>&1
# ncu evidence: the first launches of each hot kernel family of one forward (level-0 shapes), full metric set + source
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gemm_tc_kernel' -c 14 -o gpurun_out/k_prof_gemm python tests/diag_profile.py > gpurun_out/k_ncu_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/k_box.txt
VARIANT=ip16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'attention_tc2|attention_tcg|attention_cx|attention_mma|gn_apply_rows|gn_stats|ln_stats|temporal_attention' -c 24 -o gpurun_out/k_prof_misc python tests/diag_profile.py > gpurun_out/k_ncu_misc.log 2>&1
echo "ncu misc exit $?" >> gpurun_out/k_box.txt
tail -3 gpurun_out/k_kernels.log; tail -3 gpurun_out/k_engine.log; cat gpurun_out/k_box.txt; head -16 gpurun_out/k_probe.txt; grep -E "cross_attention" gpurun_out/k_probe.txt; ls -la gpurun_out/k_prof*.ncu-rep; tail -2 gpurun_out/k_ncu_gemm.log gpurun_out/k_ncu_misc.log
