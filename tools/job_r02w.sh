cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02w_pytest.txt 2>&1; tail -3 $O/r02w_pytest.txt; cat $O/test_errors.json
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02w_smoke.txt 2>&1; tail -5 $O/r02w_smoke.txt
timeout 700 python bench.py --steps 20 --warmup 3 > $O/r02w_bench.json 2> $O/r02w_bench.err; python - <<PY
import json
d=json.load(open('$O/r02w_bench.json'))
print('decode', int(d['value']), 'e2e', int(d['e2e']['value']), {k[12:]:round(v['ms_per_launch'],2) for k,v in d['roofline']['all_kernels'].items()}, d['clocks'])
e=d['encode']; print('encode', int(e['value']), 'e2e', int(e['e2e']['value']), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()}, 'cpu', int(e['cpu_baseline']['value']))
print('mdct', d.get('mdct_kernel_roofline',{}).get('frac'), 'launches', d['gpu_launches'])
PY
tail -3 $O/r02w_bench.err
