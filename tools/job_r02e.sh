cd $GRAFT_REPO_ROOT
for regs in 128 96 80 64; do for n in 65536 131072; do
OPUSB200_ENC_REGS=$regs python bench.py --workload encode --steps 6 --warmup 3 --streams $n > gpurun_out/r02e_tmp.json 2>> gpurun_out/r02e.err
python - <<PY
import json
e=json.load(open('gpurun_out/r02e_tmp.json')); e=e.get('encode',e)
print('regs $regs streams $n value %d ms/step %.2f e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value']), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
PY
done; done 2>&1 | tee gpurun_out/r02e_regs.txt
OPUSB200_ENC_REGS=64 python bench.py --workload encode --steps 6 --warmup 3 --streams 262144 > gpurun_out/r02e_tmp.json 2>> gpurun_out/r02e.err
python - <<PY | tee -a gpurun_out/r02e_regs.txt
import json
e=json.load(open('gpurun_out/r02e_tmp.json')); e=e.get('encode',e)
print('regs 64 streams 262144 value %d ms/step %.2f e2e %d'%(e['value'],e['ms_per_step'],e['e2e']['value']), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
PY
tail -5 gpurun_out/r02e.err
