#!/bin/bash
# round 2, GPU call F: full validation of the current state + benches + ncu evidence
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
echo start > gpurun_out/f_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_full_parity_gpu.py > gpurun_out/f_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/f_box.txt
timeout 1800 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/f_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/f_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/f_probe.txt 2>&1
FYC_LN_FOLD=0 timeout 300 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/f_probe_FYC_LN_FOLD_off.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench exit $?" >> gpurun_out/f_box.txt
timeout 600 python bench.py --steps 2 --warmup 2 --workload cfg3 > gpurun_out/f_bench_cfg3.json 2> gpurun_out/f_bench_cfg3.err
echo "bench cfg3 exit $?" >> gpurun_out/f_box.txt
# ncu: launch list of one bench step (kernel by kernel), with DRAM bytes; then a full capture of the dominant kernel
FYC_NO_GRAPH=1 FYC_CUPROF=1 timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 1 --ddim-steps 1 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/f_box.txt
tail -3 gpurun_out/f_gpu_tests.log; tail -3 gpurun_out/f_parity.log; cat gpurun_out/f_box.txt; head -16 gpurun_out/f_probe.txt; sed -n 2,4p gpurun_out/f_probe_FYC_LN_FOLD_off.txt; head -c 400 gpurun_out/f_bench.json; wc -l gpurun_out/f_launches.csv
