#!/bin/bash
# round 2, GPU call N: GroupNorm statistics pass cut into 3 (current) / 1 / 2 waves of resident CTAs - same-box A/B with the kernel tests
# run under each non-default setting
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo start > gpurun_out/n_box.txt
SHAPES=1 FYC_GN_WAVES=3 timeout 120 python tests/perf_probe.py > gpurun_out/n_probe_w3.txt 2>&1
for W in 1 2; do
  FYC_GN_WAVES=$W timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "groupnorm" > gpurun_out/n_tests_w$W.log 2>&1
  echo "W=$W tests exit $?" >> gpurun_out/n_box.txt
  SHAPES=1 FYC_GN_WAVES=$W timeout 100 python tests/perf_probe.py > gpurun_out/n_probe_w$W.txt 2>&1
done
cat gpurun_out/n_box.txt; for W in 3 1 2; do echo "== W=$W"; sed -n 2p gpurun_out/n_probe_w$W.txt; grep -E "^   groupnorm " gpurun_out/n_probe_w$W.txt | head -2; grep -E "groupnorm\[" gpurun_out/n_probe_w$W.txt | head -12; done; tail -2 gpurun_out/n_tests_w1.log gpurun_out/n_tests_w2.log
exit 0
