set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -q --maxfail=25 -x -k "upsample_phases or conv_head or build_unet_input or resampler or hoist" ) > gpurun_out/pytest_new.log 2>&1
tail -5 gpurun_out/pytest_new.log
( time timeout 600 python -m pytest tests -m gpu -q --maxfail=25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
( time SHAPES=1 timeout 300 python tests/perf_probe.py ) > gpurun_out/probe_new.log 2>&1
( time FYC_UP2_PHASES=0 FYC_TC_HEAD=0 timeout 300 python tests/perf_probe.py ) > gpurun_out/probe_old.log 2>&1
grep -E "UNet fwd|VAE decode" gpurun_out/probe_new.log gpurun_out/probe_old.log
( time timeout 400 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
tail -c 1500 gpurun_out/bench_r1b.json
