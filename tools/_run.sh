set -x
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 150 $TR8 tools/check_sharded_sequence.py > gpurun_out/r2_sharded_seq8b.log 2>&1; grep '^{' gpurun_out/r2_sharded_seq8b.log || tail -25 gpurun_out/r2_sharded_seq8b.log
timeout 200 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench8b.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench8b.log | cut -c1-500
timeout 200 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench4b.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench4b.log | cut -c1-500
