cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 420 python -m pytest tests/test_gpu_lpcnet_enc.py tests/test_gpu_lpcnet_plc.py tests/test_gpu_dred.py -q -s -x 2>&1 | tail -40 | tee $O/r02z_new_tests.txt
cp $O/test_errors.json $O/r02z_test_errors.json 2>/dev/null
