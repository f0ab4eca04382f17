TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | head -8
timeout 300 $TR tools/check_sharded.py > gpurun_out/sharded8.log 2>&1; grep "sharded" gpurun_out/sharded8.log | tail -4
timeout 400 $TR tools/bench_matcher.py --rows 8000000 --batches 32,256,1024 > gpurun_out/matcher8_weak.log 2>&1; grep '^{' gpurun_out/matcher8_weak.log | tail -3
timeout 300 $TR tools/bench_matcher.py --rows 1000000 --batches 32,256,1024 > gpurun_out/matcher8_strong.log 2>&1; grep '^{' gpurun_out/matcher8_strong.log | tail -3
timeout 400 $TR tools/bench_streaming.py --db-rows 10000000 --batch 256 --steps 5 > gpurun_out/stream8.log 2>&1; grep '^{' gpurun_out/stream8.log | tail -2; tail -3 gpurun_out/stream8.log
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench8.log 2>&1; grep '^{' gpurun_out/bench8.log | tail -1 | cut -c1-700
