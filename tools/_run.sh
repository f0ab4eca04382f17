timeout 900 python -m pytest tests -x -q -m gpu -k "match or stream or fullsize or sharded or db" > gpurun_out/pytest_match.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_match.log
timeout 600 python tools/bench_matcher.py --rows 1000000 --dim 4096 --batches 32,256,512,1024,4096 > gpurun_out/matcher_pair.log 2>&1; grep '^{' gpurun_out/matcher_pair.log | cut -c1-330
