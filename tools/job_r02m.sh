cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02m_pytest.txt 2>&1; tail -6 $O/r02m_pytest.txt; cat $O/test_errors.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02m_smoke.txt 2>&1; tail -4 $O/r02m_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02m_bench.json 2> $O/r02m_bench.err; python - <<PY
import json
d=json.load(open('$O/r02m_bench.json'))
print('decode', int(d['value']), 'e2e', int(d['e2e']['value']), {k[12:]:round(v['ms_per_launch'],2) for k,v in d['roofline']['all_kernels'].items()}, d['clocks'])
e=d['encode']; print('encode', int(e['value']), 'e2e', int(e['e2e']['value']), {k[8:]:round(v['ms_per_launch'],2) for k,v in e['roofline']['all_kernels'].items()})
print('mdct', d.get('mdct_kernel_roofline'))
PY
tail -3 $O/r02m_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02m_bench_ref.json 2>&1; tail -c 600 $O/r02m_bench_ref.json
# compute-sanitizer over the kernels that had no run yet: the concealment kernel P, the Opus-layer encoder at complexity 10, the long-block MDCT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_plc.py::test_random_loss_patterns_vs_reference[2-480-48000]" tests/test_gpu_opus_encode.py::test_submit_wait_matches_the_synchronous_call "tests/test_gpu_ops.py::test_mdct_forward_long_block_kernel[77]" tests/test_gpu_lpcnet_plc.py::test_history_analysis_is_unimplemented_not_faked -q -x > $O/r02m_sanitizer_memcheck.txt 2>&1; echo memcheck rc $?; tail -4 $O/r02m_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_plc.py::test_random_loss_patterns_vs_reference[2-480-48000]" "tests/test_gpu_ops.py::test_mdct_forward_long_block_kernel[77]" -q -x > $O/r02m_sanitizer_racecheck.txt 2>&1; echo racecheck rc $?; tail -4 $O/r02m_sanitizer_racecheck.txt
ls -la $O | head -30
