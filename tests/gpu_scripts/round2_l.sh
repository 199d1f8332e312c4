#!/bin/bash
# round 2, GPU call L: validation of the final state (all -m gpu tests incl. full-size parity), bench, probe, ncu launch list + small
# --set full captures (reports kept well under the 64 MiB gpurun_out limit)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
echo start > gpurun_out/l_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/l_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/l_box.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
echo "bench exit $?" >> gpurun_out/l_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/l_probe.txt 2>&1
FYC_CROSS_TC=0 timeout 200 python tests/perf_probe.py 2>&1 | head -14 > gpurun_out/l_probe_FYC_CROSS_TC_off.txt
FYC_NO_GRAPH=1 FYC_CUPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/l_launches.csv python bench.py --steps 1 --warmup 1 --ddim-steps 1 --no-cpu-baseline > gpurun_out/l_ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/l_box.txt
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gemm_tc_kernel' -s 3 -c 8 -o gpurun_out/l_prof_gemm python tests/diag_profile.py > gpurun_out/l_ncu_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/l_box.txt
VARIANT=ip16 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'attention_tc2|attention_tcg|attention_cx|gn_apply_rows|gn_stats|ln_stats|temporal_attention' -c 12 -o gpurun_out/l_prof_misc python tests/diag_profile.py > gpurun_out/l_ncu_misc.log 2>&1
echo "ncu misc exit $?" >> gpurun_out/l_box.txt
du -sm gpurun_out | cut -f1 > gpurun_out/l_size_mb.txt
if [ "$(cat gpurun_out/l_size_mb.txt)" -gt 55 ]; then rm -f gpurun_out/l_prof_misc.ncu-rep; echo "dropped misc report (size)" >> gpurun_out/l_box.txt; fi
timeout 300 python bench.py --steps 2 --warmup 2 --workload cfg3 --no-cpu-baseline > gpurun_out/l_bench_cfg3.json 2> gpurun_out/l_bench_cfg3.err
echo "bench cfg3 exit $?" >> gpurun_out/l_box.txt
tail -3 gpurun_out/l_gpu_tests.log; cat gpurun_out/l_box.txt; head -16 gpurun_out/l_probe.txt; sed -n 2,4p gpurun_out/l_probe_FYC_CROSS_TC_off.txt; head -c 300 gpurun_out/l_bench.json; echo; wc -l gpurun_out/l_launches.csv; ls -la gpurun_out/*.ncu-rep; du -sm gpurun_out
exit 0
