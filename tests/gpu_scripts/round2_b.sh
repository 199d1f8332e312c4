#!/bin/bash
# round 2, GPU call B: new kernels (LN fold, 4-group tcgen05 attention, two-source GN / GEMM), A/B benches, probe, cfg3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
echo start > gpurun_out/b_box.txt
# the kernels first, each under its own timeout (a hang must not eat the call)
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "layernorm_fold or layernorm_stats or two_sources or two_segment or self_attention_tcgen05 or fused_ip" > gpurun_out/b_newkernels.log 2>&1
echo "new kernels exit $?" >> gpurun_out/b_box.txt
timeout 300 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider -k "L4096" >> gpurun_out/b_newkernels.log 2>&1
echo "attn L4096 exit $?" >> gpurun_out/b_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_full_parity_gpu.py > gpurun_out/b_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/b_box.txt
timeout 1800 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/b_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/b_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/b_probe.txt 2>&1
for v in "FYC_LN_FOLD=0" "FYC_ATTN_G4=0" "FYC_DUAL_SOURCE=0"; do
  env $v timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/b_probe_${v%%=*}_off.txt
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench exit $?" >> gpurun_out/b_box.txt
timeout 600 python bench.py --steps 2 --warmup 2 --workload cfg3 > gpurun_out/b_bench_cfg3.json 2> gpurun_out/b_bench_cfg3.err
echo "bench cfg3 exit $?" >> gpurun_out/b_box.txt
tail -5 gpurun_out/b_newkernels.log; tail -5 gpurun_out/b_gpu_tests.log; tail -4 gpurun_out/b_parity.log; cat gpurun_out/b_box.txt; head -14 gpurun_out/b_probe.txt
