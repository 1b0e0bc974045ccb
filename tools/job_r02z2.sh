cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_opus_encode.py tests/test_gpu_dred.py -q -s -x -k "multiframe or unsupported or encode_silk or submit" 2>&1 | tail -30 | tee $O/r02z2_new_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $O/r02z2_smoke.txt
