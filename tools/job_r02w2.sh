cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -x -k "find_pred_coefs and 8-4" 2>&1 | grep -v "^$" | head -60 > $O/r02w2_sanitizer.txt; head -45 $O/r02w2_sanitizer.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -q -s -k "noise_shape or (find_pred_coefs and 12-2)" 2>&1 | tail -15 | tee $O/r02w2_tests.txt
