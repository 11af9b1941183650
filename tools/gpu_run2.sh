timeout 1200 python -m pytest tests/ -q -m gpu -x -o faulthandler_timeout=300 2>&1 | grep -v "^$" | tail -8
timeout 280 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','root_roofline','gpu_launches') if k in d})
for r in d.get('configs',[]):
    print(r['tag'], '%.3e'%r['value'], 'ms', round(r['ms_per_step'],3), 'kernel_ms', round(r['kernel_ms'],3), 'root_ms', round(r['root_ms'],4), r['kernel'], 'frac', round(r['roofline']['frac'],4), 'e2e %.3e'%r['e2e']['value'])
PY
