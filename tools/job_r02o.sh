cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "silk" 2>&1 | tail -3
run() {
timeout 300 python bench.py --workload $3 --steps 10 --warmup 3 --streams $2 > $O/tmp.json 2>> $O/r02o.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json')); e=e.get('encode',e) if '$3'=='encode' else e
    print('$1 $3 value %d ms/step %.2f e2e %d sync-e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value'],e['e2e'].get('sync_call_value_rank0',0)), {k.split('_',2)[-1]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
except Exception as ex:
    print('$1 failed', ex)
PY
}
( run l1carve 65536 encode
run d2h_one 65536 decode
OPUSB200_D2H_SPLIT=1 run d2h_split 65536 decode ) 2>&1 | tee $O/r02o_runs.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ob_celt_dec_plc -s 6 -c 1 -f -o /tmp/plc python bench.py --workload decode --steps 4 --warmup 3 --loss 0.05 > /dev/null 2>&1
timeout 100 python tools/ncu_summary.py /tmp/plc.ncu-rep 3277 30 > $O/r02o_ncu_ob_celt_dec_plc.txt 2>&1; head -50 $O/r02o_ncu_ob_celt_dec_plc.txt | tail -22
tail -3 $O/r02o.err
