cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "mdct or resampler" 2>&1 | tail -3
timeout 120 python tools/mdct_bench.py | tee $O/r02k_mdct_tma.json
OPUSB200_MDCT_SMEM_FFT=1 timeout 120 python tools/mdct_bench.py | tee $O/r02k_mdct_general.json
run() {
timeout 240 python bench.py --workload encode --steps 6 --warmup 3 --streams $2 > $O/tmp.json 2>> $O/r02k.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json')); e=e.get('encode',e)
    print('$1 streams $2 value %d ms/step %.2f e2e %d sync-e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value'],e['e2e'].get('sync_call_value_rank0',0)), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
except Exception as ex:
    print('$1 failed', ex)
PY
}
( OPUSB200_ENC_AN_TEAMS=1 run teams1 65536
OPUSB200_ENC_AN_TEAMS=3 run teams3 65536
OPUSB200_ENC_AN_TEAMS=4 run teams4 65536 ) 2>&1 | tee $O/r02k_teams.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ob_mdct_forward_tma -s 2 -c 1 -f -o $O/r02k_mdct_tma python tools/mdct_bench.py --reps 2 > /dev/null 2>&1
timeout 100 python tools/ncu_summary.py $O/r02k_mdct_tma.ncu-rep 131072 30 > $O/r02k_ncu_mdct_tma.txt 2>&1
tail -5 $O/r02k.err
