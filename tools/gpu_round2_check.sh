# Round-2 GPU check (one box, one GPU): the whole -m gpu suite (twice: flakiness), the bench line, the eval-forward workload,
# the launch list of an eager step and ncu --set full of the kernels that had no capture yet.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
for i in 1 2; do (timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t4_$i.log 2>&1; echo rc=$? >> gpurun_out/t4_$i.log); done
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b4.log 2> gpurun_out/b4.err
timeout 300 python bench.py --workload infer --steps 20 --warmup 5 > gpurun_out/b4_infer.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches4.csv python tools/profile_step.py 8 > gpurun_out/p4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:add_ln_bwd|gn_bwd_stats|msda_bwd_vec' -c 6 --profile-from-start off -f -o gpurun_out/misc_full python tools/profile_step.py 8 > gpurun_out/p5.log 2>&1
python tools/ncu_summary.py gpurun_out/misc_full.ncu-rep > gpurun_out/misc_full2.txt 2>&1
rm -f gpurun_out/misc_full.ncu-rep
tail -4 gpurun_out/t4_1.log gpurun_out/t4_2.log
for f in gpurun_out/b4.log gpurun_out/b4_infer.log; do echo $f; python - "$f" <<'P'
import sys,json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('loss'), d.get('batch16',{}).get('value'), (d.get('roofline') or {}).get('frac'))
    elif 'capture failed' in l or 'Error' in l: print(l.strip()[:300])
P
done
tail -3 gpurun_out/b4.err
