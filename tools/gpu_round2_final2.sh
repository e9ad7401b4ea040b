# refresh of the one-GPU numbers after the presorted-subset change (run under gpurun from the repo root)
set -x
timeout 600 python bench.py --full --full-perms 1000 > gpurun_out/full_scdd_r02_final2_n1.json 2> gpurun_out/full_scdd_r02_final2_n1.err; echo "full rc=$?"; cut -c1-600 gpurun_out/full_scdd_r02_final2_n1.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_final2.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches_r02_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_final2.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:fit_kernel -s 3 -c 1 -o gpurun_out/fit_final2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_final2.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/fit_final2.ncu-rep gpurun_out/fit_kernel_r02_final2_ncu.json > /dev/null 2>&1
python tools/ncu_regions.py gpurun_out/fit_final2.ncu-rep > gpurun_out/fit_kernel_r02_final2_regions.txt 2>&1
head -12 gpurun_out/fit_kernel_r02_final2_regions.txt
