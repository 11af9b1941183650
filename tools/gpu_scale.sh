# usage: bash tools/gpu_scale.sh N [quick]  (inside gpurun --gpus N): 2-rank NCCL parity test + the bench line at N ranks
N=${1:-2}
O=gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -x -o faulthandler_timeout=200 > $O/r2_multi_gpu_test_n$N.log 2>&1; tail -3 $O/r2_multi_gpu_test_n$N.log
if [ "$2" = "quick" ]; then
  for env in "" "TSP_NO_P2P_GATHER=1"; do
    env $env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 --no-configs > $O/r2_quick_n$N.json 2> $O/r2_quick_n$N.err; echo "bench rc=$? $env"
    python -c "
import json
d=json.loads(open('$O/r2_quick_n$N.json').read().strip().splitlines()[-1])
print('N=%d value %.4e ms %.4f e2e %.4e (%.4f ms) kernel_ms %.4f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))"
  done
  exit 0
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('N=%d value %.4e ms %.4f e2e %.4e kernel_ms %.4f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']))
for r in d.get('configs',[]):
    print(r['tag'], r.get('error') or ('%.3e ms %.3f kernel_ms %.3f roots/gpu %d e2e %.3e' % (r['value'], r['ms_per_step'], r['kernel_ms'], r['roots_per_gpu'], r['e2e']['value'])))
PY
