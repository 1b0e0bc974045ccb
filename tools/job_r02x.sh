cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_dred.py tests/test_gpu_lpcnet_plc.py -q 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --workload decode > $O/r02x_out.txt 2> $O/r02x_err.txt; wc -l $O/r02x_out.txt; head -c 200 $O/r02x_out.txt; echo
timeout 600 python tools/bench_ops.py --reps 3 > $O/r02x_bench_ops.txt 2>&1; tail -4 $O/r02x_bench_ops.txt | cut -c1-300
