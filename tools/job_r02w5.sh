cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -q -s -k "pitch_lags" 2>&1 | tail -8 | tee $O/r02w5_pitch_tests.txt
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -x -k "pitch_lags and 16-4" 2>&1 | grep -v "^$" | tail -5 | tee -a $O/r02w5_pitch_tests.txt
