cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02w3_pytest.txt 2>&1; tail -3 $O/r02w3_pytest.txt; cat $O/test_errors.json
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -x -k "find_pred_coefs or noise_shape or find_lpc or process_nlsfs or process_gains" 2>&1 | grep -v "^$" | tail -6 | tee $O/r02w3_sanitizer.txt
