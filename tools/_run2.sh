TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 400 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2.log 2>&1; echo "exit $?"; tail -1 gpurun_out/bench2.log | cut -c1-700
timeout 400 $TR2 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench2_ref.log 2>&1; echo "exit $?"; tail -1 gpurun_out/bench2_ref.log | cut -c1-200
