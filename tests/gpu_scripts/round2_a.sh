#!/bin/bash
# round 2, GPU call A: full-size parity (oracle on the GPU), the whole -m gpu suite, bench (default + shared prefix), per-shape probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_box.txt 2>&1
(nproc; free -g | head -2) >> gpurun_out/a_box.txt 2>&1
timeout 1800 python -m pytest tests/test_full_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/a_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/a_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_full_parity_gpu.py > gpurun_out/a_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/a_box.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?" >> gpurun_out/a_box.txt
FYC_SHARED_PREFIX=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_shared.json 2> gpurun_out/a_bench_shared.err
echo "bench shared exit $?" >> gpurun_out/a_box.txt
SHAPES=1 timeout 300 python tests/perf_probe.py > gpurun_out/a_probe.txt 2>&1
tail -5 gpurun_out/a_parity.log; tail -5 gpurun_out/a_gpu_tests.log; cat gpurun_out/a_box.txt; head -c 600 gpurun_out/a_bench.json
