set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.txt 2>&1; tail -8 gpurun_out/r02h_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.txt 2>&1; tail -4 gpurun_out/r02h_smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 6000 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02h_bench_ref.json 2>&1; tail -c 2500 gpurun_out/r02h_bench_ref.json
python tools/float_gap_probe.py > gpurun_out/r02h_float_gap.txt 2>&1; tail -20 gpurun_out/r02h_float_gap.txt
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02h_launches_decode.csv python bench.py --workload decode --steps 2 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02h_launches_encode.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 > /dev/null 2>&1
for k in ob_celt_enc_frame_rdo ob_celt_enc_analysis ob_celt_enc_chains; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/r02h_$k python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 > /dev/null 2>&1
done
for k in ob_celt_dec_pvq ob_celt_dec_entropy ob_celt_dec_synth; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r02h_$k python bench.py --workload decode --steps 2 --warmup 3 > /dev/null 2>&1
done
ls -la gpurun_out
