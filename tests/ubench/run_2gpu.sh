set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_2gpu.txt 2>&1
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 400 gpurun_out/bench_2gpu.err
head -c 700 gpurun_out/bench_2gpu.json
