mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_selfplay.py -x -q -m gpu -k resident > gpurun_out/t_final2.log 2>&1; echo "selfplay gpu tests rc=$?"; tail -4 gpurun_out/t_final2.log
