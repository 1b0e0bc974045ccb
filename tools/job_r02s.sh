cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_opus_encode.py tests/test_gpu_encode.py tests/test_gpu_ops.py -q 2>&1 | tail -3
run() {
timeout 400 python bench.py --workload $2 --steps 10 --warmup 3 > $O/tmp.json 2>> $O/r02s.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json')); e=e.get('encode',e) if '$2'=='encode' else e
    print('$1 $2 value %d ms/step %.2f e2e %d bytes %.1f'%(e['value'],e['ms_per_step'],e['e2e']['value'],e['config'].get('mean_packet_bytes',0)), {k.split('_',2)[-1]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()}, 'cpu', int(e['cpu_baseline']['value']))
except Exception as ex:
    print('$1 failed', ex)
PY
}
( run distinct_binned encode
OPUSB200_ENC_NO_BIN=1 run distinct_unbinned encode ) 2>&1 | tee $O/r02s_runs.txt
tail -3 $O/r02s.err
