cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_ops.py -q -s -x -k "process_gains or process_nlsfs" 2>&1 | tail -12 | tee $O/r02z3_new_tests.txt
