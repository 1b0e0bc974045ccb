set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_gpu_decode.py tests/test_gpu_plc.py -q 2>&1 | tail -2; cat $O/test_errors.json
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum
ncu --metrics $M --clock-control none --csv --log-file $O/r02i_launches_decode.csv python bench.py --workload decode --steps 2 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file $O/r02i_launches_encode.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 > /dev/null 2>&1
python tools/launch_summary.py $O/r02i_launches_decode.csv ob_ > $O/r02i_launch_summary_decode.txt
python tools/launch_summary.py $O/r02i_launches_encode.csv ob_ > $O/r02i_launch_summary_encode.txt
cat $O/r02i_launch_summary_decode.txt $O/r02i_launch_summary_encode.txt
for k in ob_celt_enc_frame_rdo ob_celt_enc_analysis ob_celt_enc_chains ob_opus_enc_analysis; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o /tmp/$k python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/$k.ncu-rep 65536 40 > $O/r02i_ncu_$k.txt 2>&1
python tools/ncu_lanes.py /tmp/$k.ncu-rep 50 > $O/r02i_lanes_$k.txt 2>&1
done
for k in ob_celt_dec_pvq ob_celt_dec_entropy ob_celt_dec_synth; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o /tmp/$k python bench.py --workload decode --steps 2 --warmup 3 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/$k.ncu-rep 65536 40 > $O/r02i_ncu_$k.txt 2>&1
python tools/ncu_lanes.py /tmp/$k.ncu-rep 50 > $O/r02i_lanes_$k.txt 2>&1
done
cp /tmp/ob_celt_enc_frame_rdo.ncu-rep $O/r02i_enc_frame_rdo.ncu-rep
cp /tmp/ob_celt_enc_analysis.ncu-rep $O/r02i_enc_analysis.ncu-rep
python bench.py --workload decode --steps 20 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks'])"
du -sh $O
