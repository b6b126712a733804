mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/t_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
