set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.txt 2>&1; tail -5 gpurun_out/r02b_pytest.txt
python __graft_entry__.py smoke > gpurun_out/r02b_smoke.txt 2>&1; tail -4 gpurun_out/r02b_smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 3000 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_ref.json 2>&1; tail -c 1500 gpurun_out/r02b_bench_ref.json
python tools/float_gap_probe.py > gpurun_out/r02b_float_gap.txt 2>&1; cat gpurun_out/r02b_float_gap.txt
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02b_enc_launches.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 --streams 32768 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ob_celt_enc_frame_rdo -s 4 -c 1 -f -o gpurun_out/r02b_enc_frame_rdo python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 --streams 32768 > /dev/null 2>&1
ls -la gpurun_out
