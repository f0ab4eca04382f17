TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR tools/check_sharded_sequence.py > gpurun_out/sharded_seq8.log 2>&1; grep '^{' gpurun_out/sharded_seq8.log || tail -5 gpurun_out/sharded_seq8.log
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 300 $TR4 tools/check_sharded_sequence.py > gpurun_out/sharded_seq4.log 2>&1; grep '^{' gpurun_out/sharded_seq4.log || tail -5 gpurun_out/sharded_seq4.log
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 300 $TR2 tools/check_sharded_sequence.py > gpurun_out/sharded_seq2.log 2>&1; grep '^{' gpurun_out/sharded_seq2.log || tail -5 gpurun_out/sharded_seq2.log
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log | cut -c1-400
timeout 400 $TR tools/bench_streaming.py --db-rows 10000000 --batch 256 --steps 5 > gpurun_out/stream8_v3.log 2>&1; grep '^{' gpurun_out/stream8_v3.log | tail -1
timeout 400 $TR tools/bench_streaming.py --db-rows 10000000 --batch 256 --steps 5 --detect > gpurun_out/stream8_v3d.log 2>&1; grep '^{' gpurun_out/stream8_v3d.log | tail -1
