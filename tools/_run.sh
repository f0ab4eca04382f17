set -x
python -m pytest tests/test_gpu_staged.py tests/test_gpu_parity_pixels.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest5.log
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "similarity" 2>&1 | tail -5 >> gpurun_out/r2_pytest5.log
cat gpurun_out/r2_pytest5.log
python tools/ab_probe_side.py > gpurun_out/r2_ab_probe_side.log 2>&1; cat gpurun_out/r2_ab_probe_side.log
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 300 $TR2 tools/check_sharded_sequence.py > gpurun_out/r2_sharded_seq2.log 2>&1; grep '^{' gpurun_out/r2_sharded_seq2.log || tail -25 gpurun_out/r2_sharded_seq2.log
timeout 400 $TR2 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench2.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench2.log | cut -c1-1500
