set -x
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 200 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench8d.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench8d.log | cut -c1-300
timeout 200 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench4d.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench4d.log | cut -c1-300
grep -E "NVLS|Using network|via P2P|NCCL version" gpurun_out/r2_bench8d.log | sort | uniq -c | sort -rn | head -8 > gpurun_out/r2_nccl_info.txt; cat gpurun_out/r2_nccl_info.txt
