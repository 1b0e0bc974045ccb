cd $GRAFT_REPO_ROOT
O=gpurun_out
for k in ob_celt_enc_frame_rdo ob_celt_enc_analysis; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o /tmp/$k python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 > /dev/null 2>&1
timeout 100 python tools/ncu_summary.py /tmp/$k.ncu-rep 65536 30 > $O/r02y_ncu_$k.txt 2>&1
timeout 100 python tools/ncu_lanes.py /tmp/$k.ncu-rep 30 > $O/r02y_lanes_$k.txt 2>&1
done
for k in ob_celt_dec_synth; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o /tmp/$k python bench.py --workload decode --steps 2 --warmup 3 > /dev/null 2>&1
timeout 100 python tools/ncu_summary.py /tmp/$k.ncu-rep 65536 30 > $O/r02y_ncu_$k.txt 2>&1
done
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r02y_launches_decode.csv python bench.py --workload decode --steps 2 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r02y_launches_encode.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 > /dev/null 2>&1
python tools/launch_summary.py $O/r02y_launches_decode.csv ob_ | tee $O/r02y_launch_summary_decode.txt
python tools/launch_summary.py $O/r02y_launches_encode.csv ob_ | tee $O/r02y_launch_summary_encode.txt
grep -E "gpu__time_duration|issue_active|warp instructions|dram__bytes_(read|write).sum  " $O/r02y_ncu_*.txt | cut -c1-200
