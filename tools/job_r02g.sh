cd $GRAFT_REPO_ROOT
run() { # label, env..., streams
python bench.py --workload encode --steps 6 --warmup 3 --streams $2 > gpurun_out/r02e_tmp.json 2>> gpurun_out/r02g.err
python - <<PY
import json
e=json.load(open('gpurun_out/r02e_tmp.json')); e=e.get('encode',e)
print('$1 streams $2 value %d ms/step %.2f e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value']), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
PY
}
( run base 65536
export OPUSB200_ENC_SYNC=1
run sync 65536
run sync 75776
run sync 131072 ) 2>&1 | tee gpurun_out/r02g_sync.txt
OPUSB200_ENC_SYNC=1 ncu --set full --clock-control none -k regex:ob_celt_enc_frame_rdo -s 4 -c 1 -f -o gpurun_out/r02g_frame_rdo_sync_65536 python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 --streams 65536 > /dev/null 2>&1
OPUSB200_ENC_SYNC=1 python -m pytest tests/test_gpu_opus_encode.py -q 2>&1 | tail -2
tail -5 gpurun_out/r02g.err
