# eight-GPU measurement batch of round 2 (run under `gpurun --gpus 8` from the repo root)
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | wc -l
timeout 900 $TR bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err; echo "bench n8 rc=$?"
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_r02_n8_reference.json 2> gpurun_out/bench_r02_n8_reference.err; echo "reference n8 rc=$?"
timeout 600 python bench.py --full --full-perms 1000 --gpus 8 > gpurun_out/full_scdd_r02_n8.json 2> gpurun_out/full_scdd_r02_n8.err; echo "full n8 rc=$?"; cat gpurun_out/full_scdd_r02_n8.json
timeout 900 $TR bench.py --gpus 8 --config 5 --steps 2 --warmup 3 --perms-per-step 1 > gpurun_out/bench_r02_cfg5_n8.json 2> gpurun_out/bench_r02_cfg5_n8.err; echo "cfg5 n8 rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sharding or sharded" > gpurun_out/pytest_r02_n8.log 2>&1; echo "pytest n8 rc=$?"; tail -3 gpurun_out/pytest_r02_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/bench_r02_n8.err
