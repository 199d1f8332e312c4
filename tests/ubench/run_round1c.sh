set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -q --maxfail=25 ) > gpurun_out/pytest_gpu_c.log 2>&1
tail -15 gpurun_out/pytest_gpu_c.log
( time SHAPES=1 timeout 200 python tests/perf_probe.py ) > gpurun_out/probe_c_on.log 2>&1
( time FYC_ZIGZAG=0 FYC_ATTN_SHORTK=0 timeout 200 python tests/perf_probe.py ) > gpurun_out/probe_c_off.log 2>&1
( time FYC_ZIGZAG=0 timeout 200 python tests/perf_probe.py ) > gpurun_out/probe_c_nozz.log 2>&1
grep -E "UNet fwd|VAE decode|layernorm   |groupnorm   |attention   " gpurun_out/probe_c_on.log gpurun_out/probe_c_off.log gpurun_out/probe_c_nozz.log
( time timeout 400 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
head -c 600 gpurun_out/bench_r1c.json
