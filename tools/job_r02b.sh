set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.txt 2>&1; tail -5 gpurun_out/r02b_pytest.txt
python __graft_entry__.py smoke > gpurun_out/r02b_smoke.txt 2>&1; tail -4 gpurun_out/r02b_smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 3000 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_ref.json 2>&1; tail -c 1500 gpurun_out/r02b_bench_ref.json
