set -x
cd $GRAFT_REPO_ROOT
for g in none family same; do
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_$g.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 --streams 32768 --enc-group $g > /dev/null 2>&1
done
python bench.py --workload encode --steps 10 --warmup 3 > gpurun_out/r02d_bench_enc.json 2> gpurun_out/r02d_bench.err; tail -c 1500 gpurun_out/r02d_bench_enc.json; tail -3 gpurun_out/r02d_bench.err
python bench.py --workload encode --steps 10 --warmup 3 --enc-group family > gpurun_out/r02d_bench_enc_family.json 2>> gpurun_out/r02d_bench.err
python bench.py --workload encode --steps 10 --warmup 3 --enc-group same > gpurun_out/r02d_bench_enc_same.json 2>> gpurun_out/r02d_bench.err
python -m pytest tests/test_gpu_opus_encode.py tests/test_gpu_encode.py -q 2>&1 | tail -3
