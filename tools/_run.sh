timeout 600 python tools/sweep_step_pair.py > gpurun_out/sweep_step_pair.log 2>&1; tail -3 gpurun_out/sweep_step_pair.log
