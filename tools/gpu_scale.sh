N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_${N}gpu.log
tail -2 gpurun_out/bench_${N}gpu.log | cut -c1-300
