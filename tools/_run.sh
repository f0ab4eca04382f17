TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516"
timeout 400 $TR bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_4gpu_final.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench_4gpu_final.log > gpurun_out/r2_bench_4gpu_final.json; cut -c1-330 gpurun_out/r2_bench_4gpu_final.json
