# final one-GPU measurement batch of round 2 (run under gpurun from the repo root); every step under its own timeout
set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r02_final.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_tests_r02_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 600 python bench.py --full --full-perms 1000 > gpurun_out/full_scdd_r02_final_n1.json 2> gpurun_out/full_scdd_r02_final_n1.err; echo "full rc=$?"; cat gpurun_out/full_scdd_r02_final_n1.json
timeout 600 python bench.py --config 5 --genes 2500 --steps 3 --warmup 3 > gpurun_out/bench_r02_final_cfg5_n1.json 2> gpurun_out/bench_r02_final_cfg5_n1.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_final.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_final.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:fit_kernel -s 3 -c 1 -o gpurun_out/fit_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/fit_final.ncu-rep gpurun_out/fit_kernel_r02_final_ncu.json > /dev/null 2>&1
python tools/ncu_regions.py gpurun_out/fit_final.ncu-rep > gpurun_out/fit_kernel_r02_final_regions.txt 2>&1
ls -la gpurun_out/fit_final.ncu-rep
