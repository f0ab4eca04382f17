set -x
python -m pytest tests/test_gpu_staged.py -x -q -m gpu -k "sharded_world1" 2>&1 | tail -4
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 200 $TR2 tools/check_sharded.py > gpurun_out/r2_check_sharded2.log 2>&1; echo "exit $?"; grep "^sharded" gpurun_out/r2_check_sharded2.log || tail -20 gpurun_out/r2_check_sharded2.log
