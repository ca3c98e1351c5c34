mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/t2.log 2>&1; echo rc=$? >> gpurun_out/t2.log)
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b2.log 2> gpurun_out/b2.err
for v in NO_PDL NO_ENC_FUSED NO_PACK_OVERLAP; do env MDB_BENCH_QUICK=1 MDB_$v=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/b2_$v.log 2>&1; done
env MDB_BENCH_QUICK=1 MDB_NO_PDL=1 MDB_NO_ENC_FUSED=1 MDB_NO_PACK_OVERLAP=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/b2_NONE.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches2.csv python tools/profile_step.py 8 > gpurun_out/p2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:add_ln|gn_bwd_stats|gn_stats|stem_conv|msda_bwd_vec' -c 12 --profile-from-start off -f -o gpurun_out/misc_full python tools/profile_step.py 8 > gpurun_out/p3.log 2>&1
python tools/ncu_summary.py gpurun_out/misc_full.ncu-rep > gpurun_out/misc_full.txt 2>&1
rm -f gpurun_out/misc_full.ncu-rep
tail -12 gpurun_out/t2.log
for f in gpurun_out/b2.log gpurun_out/b2_*.log; do echo $f; python - "$f" <<'P'
import sys,json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('loss'), d['config']['timing'][:60])
    elif 'capture failed' in l or 'Error' in l: print(l.strip()[:300])
P
done
tail -3 gpurun_out/b2.err
