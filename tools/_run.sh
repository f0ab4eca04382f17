timeout 600 python -m pytest tests -x -q -m gpu -k "parts_sum or sharded_sequence or similarity" > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/check_sharded_sequence.py > gpurun_out/sharded_seq2.log 2>&1; grep '^{' gpurun_out/sharded_seq2.log || tail -5 gpurun_out/sharded_seq2.log
