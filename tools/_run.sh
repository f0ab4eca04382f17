timeout 900 python -m pytest tests -x -q -m gpu -k "surf" -s > gpurun_out/pytest_surf.log 2>&1; echo "exit $?"; grep -a "^surf\|passed\|failed\|Error\|assert" gpurun_out/pytest_surf.log | head -30
