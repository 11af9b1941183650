timeout 1200 python -m pytest tests/ -q -m gpu -s -o faulthandler_timeout=300 2>&1 | grep -v "^$" > gpurun_out/r2_gpu_tests.log; tail -2 gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -4 gpurun_out/r2_smoke.txt
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.4e ms %.4f e2e %.4e (%.4f ms) frac %.4f root %.4f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['kernels'][1]['kernel_ms']))
for r in d['configs']:
    print(r['tag'], '%.3e'%r['value'], round(r['ms_per_step'],3), round(r['kernel_ms'],3), round(r['root_ms'],4), r['kernel'], round(r['roofline']['frac'],4), 'e2e %.3e'%r['e2e']['value'])
PY
