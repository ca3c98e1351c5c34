# Round-2 GPU check (one box, one GPU): smoke(), the whole -m gpu suite, the bench line.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t6.log 2>&1; echo rc=$? >> gpurun_out/t6.log)
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b6.log 2> gpurun_out/b6.err
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/t6.log
python - gpurun_out/b6.log <<'P'
import sys,json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('loss'), d.get('batch16',{}).get('value'), (d.get('roofline') or {}).get('frac'))
    elif 'capture failed' in l or 'Error' in l: print(l.strip()[:300])
P
