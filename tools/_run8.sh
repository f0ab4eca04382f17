TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR tools/check_sharded_sequence.py > gpurun_out/sharded_seq8.log 2>&1; grep '^{' gpurun_out/sharded_seq8.log || tail -5 gpurun_out/sharded_seq8.log
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 300 $TR4 tools/check_sharded_sequence.py > gpurun_out/sharded_seq4.log 2>&1; grep '^{' gpurun_out/sharded_seq4.log || tail -5 gpurun_out/sharded_seq4.log
timeout 400 $TR tools/bench_matcher.py --rows 8000000 --batches 32,256,1024 > gpurun_out/matcher8_weak_v2.log 2>&1; grep '^{' gpurun_out/matcher8_weak_v2.log | cut -c1-420
timeout 400 $TR tools/bench_streaming.py --db-rows 10000000 --batch 256 --steps 5 > gpurun_out/stream8_v2.log 2>&1; grep '^{' gpurun_out/stream8_v2.log | tail -1
