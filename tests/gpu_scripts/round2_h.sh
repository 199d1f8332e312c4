#!/bin/bash
# round 2, GPU call H (N GPUs of one box): multi-GPU bench lines - cfg2 (clip per GPU + uint8 gather) and cfg5 (camera-LoRA model, 768x768x32f, 50 steps)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/h_box_${N}.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/h_bench_cfg2_n${N}.json 2> gpurun_out/h_bench_cfg2_n${N}.err
echo "cfg2 exit $?" >> gpurun_out/h_box_${N}.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 1 --warmup 1 --workload cfg5 > gpurun_out/h_bench_cfg5_n${N}.json 2> gpurun_out/h_bench_cfg5_n${N}.err
echo "cfg5 exit $?" >> gpurun_out/h_box_${N}.txt
cat gpurun_out/h_box_${N}.txt; tail -c 1500 gpurun_out/h_bench_cfg2_n${N}.json | head -c 700; echo; tail -c 2500 gpurun_out/h_bench_cfg5_n${N}.json | head -c 900; tail -5 gpurun_out/h_bench_cfg5_n${N}.err
