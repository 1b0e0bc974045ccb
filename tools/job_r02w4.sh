cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee $O/r02w4_smoke.txt
timeout 500 python tools/bench_silk_analysis.py --n 65536 --reps 5 2>&1 | tail -12 | tee $O/r02w4_bench_silk_analysis.json
