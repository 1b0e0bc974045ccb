cd $GRAFT_REPO_ROOT
O=gpurun_out
run() {
timeout 300 python bench.py --workload encode --steps 6 --warmup 3 --enc-group $1 > $O/tmp.json 2>> $O/r02p.err
python - <<PY
import json
try:
    e=json.load(open('$O/tmp.json')); e=e.get('encode',e)
    print('group=$1 value %d ms/step %.2f'%(e['value'],e['ms_per_step']), {k.split('_',2)[-1]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
except Exception as ex:
    print('$1 failed', ex)
PY
}
( run none; run family; run same ) 2>&1 | tee $O/r02p_groups.txt
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $O/r02p_launches_same.csv python bench.py --workload encode --steps 2 --warmup 3 --enc-warm-frames 4 --enc-group same > /dev/null 2>&1
python tools/launch_summary.py $O/r02p_launches_same.csv ob_ | tee $O/r02p_launch_summary_same.txt
tail -3 $O/r02p.err
