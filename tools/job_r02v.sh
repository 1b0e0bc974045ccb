cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/r02v_gpu_tests.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/r02v_smoke.txt
