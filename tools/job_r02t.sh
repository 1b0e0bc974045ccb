cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/r02t_gpu_tests.txt
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/r02t_smoke.txt
timeout 120 python tools/bench_silk_analysis.py --only vad_pitch --n 262144 2>&1 | tail -4 | tee $O/r02t_bench_vad_pitch_262144.json
