mkdir -p gpurun_out
MDB_BENCH_QUICK=1 MDB_ENC_WGRAD_SIDE=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/b5_side.log 2>&1
MDB_BENCH_QUICK=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/b5_base.log 2>&1
(MDB_ENC_WGRAD_SIDE=1 timeout 300 python -m pytest tests/test_model_gpu.py tests/test_model_grad_gpu.py -m gpu -x -q > gpurun_out/t5.log 2>&1; echo rc=$? >> gpurun_out/t5.log)
tail -3 gpurun_out/t5.log
for f in gpurun_out/b5_side.log gpurun_out/b5_base.log; do echo $f; python - "$f" <<'P'
import sys,json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('loss'))
    elif 'rror' in l or 'failed' in l: print(l.strip()[:300])
P
done
