timeout 600 python bench.py --no-configs --no-cpu-baseline --steps 300 > gpurun_out/r2p_a.json 2>/dev/null
TSP_NO_ZERO_COPY=1 timeout 600 python bench.py --no-configs --no-cpu-baseline --steps 300 > gpurun_out/r2p_b.json 2>/dev/null
python - <<'PY'
import json
for f in ("a","b"):
    d=json.loads(open('gpurun_out/r2p_%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value %.4e (%.4f ms) e2e %.4e" % (d['value'], d['ms_per_step'], d['e2e']['value']), {k:round(v,4) for k,v in d['e2e'].items() if k.endswith('ms') or k.endswith('ms_per_step')})
PY
timeout 900 python -m pytest tests/ -q -m gpu -x -o faulthandler_timeout=300 2>&1 | tail -3
