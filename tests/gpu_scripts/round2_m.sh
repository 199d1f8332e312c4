#!/bin/bash
# round 2, GPU call M: plain GEMM epilogue with the row bias as a template parameter and settled entry prefetches (scoreboard fix):
# GEMM / conv kernel tests, probe, then the full -m gpu suite and the default bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_full.json
echo start > gpurun_out/m_box.txt
timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "gemm or conv or geglu or fold or segment" > gpurun_out/m_kernels.log 2>&1
echo "kernels exit $?" >> gpurun_out/m_box.txt
SHAPES=1 timeout 200 python tests/perf_probe.py > gpurun_out/m_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/m_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/m_box.txt
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
echo "bench exit $?" >> gpurun_out/m_box.txt
tail -2 gpurun_out/m_kernels.log; tail -3 gpurun_out/m_gpu_tests.log; cat gpurun_out/m_box.txt; head -16 gpurun_out/m_probe.txt; grep -E "x320r\]|x640r\]|1280x1280r\]|960x320L|1920x640L|320x320\]" gpurun_out/m_probe.txt; head -c 300 gpurun_out/m_bench.json
exit 0
