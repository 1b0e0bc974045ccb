cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -s -k "encode_do_vad" 2>&1 | tail -8 | tee $O/r02u_vad_tests.txt
timeout 200 python tools/bench_silk_analysis.py --only vad_pitch 2>&1 | tail -6 | tee $O/r02u_bench_vad_pitch.json
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -x -k "encode_do_vad and 12-20" 2>&1 | grep -v "^$" | tail -4 | tee -a $O/r02u_vad_tests.txt
