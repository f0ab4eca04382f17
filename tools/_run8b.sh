TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8.log 2>&1; echo "exit $?"; tail -1 gpurun_out/bench8.log > gpurun_out/r1_bench_8gpu.json; cut -c1-330 gpurun_out/r1_bench_8gpu.json
