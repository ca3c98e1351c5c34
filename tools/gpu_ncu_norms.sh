mkdir -p gpurun_out
timeout 250 ncu --set full --clock-control none --import-source on -k 'regex:add_ln_bwd' -s 14 -c 3 --profile-from-start off -f -o gpurun_out/ln_big python tools/profile_step.py 8 > gpurun_out/p7.log 2>&1
python tools/ncu_summary.py gpurun_out/ln_big.ncu-rep > gpurun_out/ln_big.txt 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:gn_bwd_stats|gn_bwd_apply' -c 4 --profile-from-start off -f -o gpurun_out/gn_bwd python tools/profile_step.py 8 > gpurun_out/p8.log 2>&1
python tools/ncu_summary.py gpurun_out/gn_bwd.ncu-rep > gpurun_out/gn_bwd.txt 2>&1
rm -f gpurun_out/*.ncu-rep
grep -E "Kernel Name|time_duration|dram_throughput" gpurun_out/ln_big.txt gpurun_out/gn_bwd.txt | cut -c1-160
