cd $GRAFT_REPO_ROOT
O=gpurun_out
run() {
timeout 400 python bench.py --workload decode --steps 10 --warmup 3 > $O/tmp.json 2>> $O/r02t.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json'))
    print('$1 decode value %d ms/step %.2f e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value']), {k.split('_',2)[-1]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
except Exception as ex:
    print('$1 failed', ex)
PY
}
( run prefetch
OPUSB200_SYNTH_NO_PREFETCH=1 run noprefetch ) 2>&1 | tee $O/r02t_runs.txt
timeout 300 python -m pytest tests/test_gpu_decode.py -q 2>&1 | tail -2
tail -3 $O/r02t.err
