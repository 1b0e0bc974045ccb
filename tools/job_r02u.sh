cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_opus_encode.py tests/test_gpu_encode.py tests/test_compat_api.py -q 2>&1 | tail -3
run() {
timeout 400 python bench.py --workload encode --steps 10 --warmup 3 > $O/tmp.json 2>> $O/r02u.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json')); e=e.get('encode',e)
    print('$1 encode value %d ms/step %.2f e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value']), {k.split('_',2)[-1]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
except Exception as ex:
    print('$1 failed', ex)
PY
}
( run overlap
OPUSB200_ENC_NO_OVERLAP=1 run no_overlap ) 2>&1 | tee $O/r02u_runs.txt
tail -3 $O/r02u.err
