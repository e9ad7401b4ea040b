# one-GPU measurement batch of round 2 (run under gpurun from the repo root)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02f.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r02f.log
bash tools/permgen_probe.sh > gpurun_out/permgen_probe_r02.txt 2>&1; cat gpurun_out/permgen_probe_r02.txt
python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02b_reference.json 2>/dev/null
python bench.py --full --full-perms 1000 > gpurun_out/full_scdd_r02_n1.json 2> gpurun_out/full_scdd_r02_n1.err; echo "full rc=$?"; cat gpurun_out/full_scdd_r02_n1.json
python bench.py --config 5 --genes 2500 --steps 3 --warmup 3 > gpurun_out/bench_r02_cfg5_n1.json 2> gpurun_out/bench_r02_cfg5_n1.err; echo "cfg5 rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
