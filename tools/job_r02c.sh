set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.txt 2>&1; tail -8 gpurun_out/r02c_pytest.txt
python bench.py --workload encode --steps 20 --warmup 3 > gpurun_out/r02c_bench_enc.json 2> gpurun_out/r02c_bench.err; tail -c 2500 gpurun_out/r02c_bench_enc.json; tail -3 gpurun_out/r02c_bench.err
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02c_enc_launches.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 --streams 32768 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ob_celt_enc_frame_rdo -s 4 -c 1 -f -o gpurun_out/r02c_enc_frame_rdo python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 --streams 32768 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ob_celt_enc_analysis -s 4 -c 1 -f -o gpurun_out/r02c_enc_analysis python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 --streams 32768 > /dev/null 2>&1
ls -la gpurun_out
