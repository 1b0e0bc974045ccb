cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_gpu_opus_encode.py tests/test_gpu_encode.py -q -x 2>&1 | tail -3
run() { # label, streams, extra env already exported
python bench.py --workload encode --steps 6 --warmup 3 --streams $2 > $O/tmp.json 2>> $O/r02j.err
python - <<PY
import json
e=json.load(open('$O/tmp.json')); e=e.get('encode',e)
print('$1 streams $2 value %d ms/step %.2f e2e %d sync-e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value'],e['e2e'].get('sync_call_value_rank0',0)), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
PY
}
( run noinline 65536
OPUSB200_ENC_SYNC=1 run noinline_sync 65536
OPUSB200_ENC_REGS=64 run noinline_r64 131072
cp opus_b200/libopus_b200.so /tmp/new.so; cp opus_b200/libopus_b200_base.so.keep opus_b200/libopus_b200.so
run base 65536
OPUSB200_ENC_SYNC=1 run base_sync 65536
cp /tmp/new.so opus_b200/libopus_b200.so ) 2>&1 | tee $O/r02j_ab.txt
tail -5 $O/r02j.err
