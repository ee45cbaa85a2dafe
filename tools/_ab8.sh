cd $GRAFT_REPO_ROOT
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload lora_step --steps 5 --warmup 3 > gpurun_out/lora_8gpu_v4.json 2> gpurun_out/lora_8gpu_v4.log
tail -c 900 gpurun_out/lora_8gpu_v4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-latency > gpurun_out/bench_8gpu_v4.json 2> gpurun_out/bench_8gpu_v4.log
tail -c 600 gpurun_out/bench_8gpu_v4.json
