mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --scan-tokens 0 > gpurun_out/bench_n2_v2.json 2> gpurun_out/bench_n2_v2.err; echo rc=$? > gpurun_out/bench_n2_v2.rc
