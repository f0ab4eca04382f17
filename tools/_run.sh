set -x
python tools/check_config2_full.py --precision auto --out gpurun_out/r2_config2_full_parity.json > gpurun_out/r2_config2_full.log 2>&1
tail -4 gpurun_out/r2_config2_full.log
