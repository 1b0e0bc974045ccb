cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:ob_celt_enc_frame_rdo -s 4 -c 1 -f -o gpurun_out/r02f_frame_rdo_65536 python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 --streams 65536 > /dev/null 2>&1
OPUSB200_ENC_REGS=64 ncu --set full --clock-control none -k regex:ob_celt_enc_frame_rdo -s 4 -c 1 -f -o gpurun_out/r02f_frame_rdo_r64_131072 python bench.py --workload encode --steps 1 --warmup 3 --enc-warm-frames 6 --streams 131072 > /dev/null 2>&1
for ch in 2 4; do
OPUSB200_ENC_CHUNKS=$ch python bench.py --workload encode --steps 6 --warmup 3 --streams 65536 > gpurun_out/r02e_tmp.json 2>> gpurun_out/r02f.err
python - <<PY
import json
e=json.load(open('gpurun_out/r02e_tmp.json')); e=e.get('encode',e)
print('chunks $ch value %d ms/step %.2f e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value']), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
PY
done 2>&1 | tee gpurun_out/r02f_chunks.txt
tail -5 gpurun_out/r02f.err
