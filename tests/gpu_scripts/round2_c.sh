#!/bin/bash
# round 2, GPU call C: LN fold v2 (mean term as a K block), spill-free epilogue, attention G4 vs G2 under ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo start > gpurun_out/c_box.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "layernorm_fold or layernorm_stats or two_sources or two_segment or gemm_bias_residual or geglu or cta_pairs" > gpurun_out/c_newkernels.log 2>&1
echo "new kernels exit $?" >> gpurun_out/c_box.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_zz_late_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/c_box.txt
timeout 900 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider -k "unet_forward or pipeline_cfg2 or shared" > gpurun_out/c_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/c_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/c_probe.txt 2>&1
FYC_LN_FOLD=0 timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/c_probe_FYC_LN_FOLD_off.txt
FYC_ATTN_G4=0 timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/c_probe_FYC_ATTN_G4_off.txt
for pz in 0 2 3 4; do FYC_ATTN_POLY=$pz timeout 120 python tests/diag_attn.py; done > gpurun_out/c_attn_g4_poly.txt 2>&1
for pz in 2 3; do FYC_ATTN_G4=0 FYC_ATTN_POLY=$pz timeout 120 python tests/diag_attn.py; done > gpurun_out/c_attn_g2_poly.txt 2>&1
NB=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc4 -s 2 -c 1 -o gpurun_out/c_prof_attn_g4 python tests/diag_attn.py > gpurun_out/c_ncu_g4.log 2>&1
NB=4 FYC_ATTN_G4=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 2 -c 1 -o gpurun_out/c_prof_attn_g2 python tests/diag_attn.py > gpurun_out/c_ncu_g2.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench exit $?" >> gpurun_out/c_box.txt
tail -3 gpurun_out/c_newkernels.log; tail -3 gpurun_out/c_engine.log; tail -3 gpurun_out/c_parity.log; cat gpurun_out/c_box.txt; head -16 gpurun_out/c_probe.txt; cat gpurun_out/c_attn_g4_poly.txt gpurun_out/c_attn_g2_poly.txt; ls -la gpurun_out/*.ncu-rep
