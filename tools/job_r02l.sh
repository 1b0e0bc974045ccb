cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_lpcnet_plc.py tests/test_gpu_fargan.py -q -x 2>&1 | tail -4
run() {
timeout 240 python bench.py --workload encode --steps 6 --warmup 3 --streams $2 > $O/tmp.json 2>> $O/r02l.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json')); e=e.get('encode',e)
    print('$1 streams $2 value %d ms/step %.2f e2e %d sync-e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value'],e['e2e'].get('sync_call_value_rank0',0)), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
except Exception as ex:
    print('$1 failed', ex)
PY
}
( OPUSB200_ENC_AN_TEAMS=1 run inl_teams1 65536
OPUSB200_ENC_AN_TEAMS=3 run inl_teams3 65536
OPUSB200_ENC_AN_TEAMS=8 run inl_teams8 65536 ) 2>&1 | tee $O/r02l_teams.txt
timeout 300 python tools/plc_gap_probe.py 2>&1 | tail -32 | tee $O/r02l_plc_gap.txt
tail -5 $O/r02l.err
