set -x
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 300 $TR8 tools/check_sharded_sequence.py > gpurun_out/r2_sharded_seq8.log 2>&1; grep '^{' gpurun_out/r2_sharded_seq8.log || tail -25 gpurun_out/r2_sharded_seq8.log
NCCL_DEBUG=WARN timeout 400 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench8.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench8.log | cut -c1-600
timeout 400 $TR4 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench4.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench4.log | cut -c1-600
