set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_d.log 2>&1
tail -3 gpurun_out/smoke_d.log
( time timeout 400 python -m pytest tests -m gpu -q --maxfail=25 -k "ragged or short_context or layernorm or groupnorm" ) > gpurun_out/pytest_gpu_d.log 2>&1
tail -6 gpurun_out/pytest_gpu_d.log
( time F=32 HW=96 timeout 300 python tests/perf_probe.py ) > gpurun_out/probe_cfg5_shape.log 2>&1
grep -E "UNet fwd|VAE decode|max mem|Error|error" gpurun_out/probe_cfg5_shape.log
( time timeout 400 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
head -c 400 gpurun_out/bench_r1d.json
( time FYC_CUPROF=1 FYC_NO_GRAPH=1 timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline ) > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/ncu_d.log
wc -l gpurun_out/launches_d.csv
