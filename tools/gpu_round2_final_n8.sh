# final eight-GPU measurements of round 2 (run under `gpurun --gpus 8`): the whole scDD() call and config 5
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | wc -l
timeout 420 python bench.py --full --full-perms 1000 --gpus 8 > gpurun_out/full_scdd_r02_final_n8.json 2> gpurun_out/full_scdd_r02_final_n8.err; echo "full n8 rc=$?"; cat gpurun_out/full_scdd_r02_final_n8.json
timeout 600 $TR bench.py --gpus 8 --config 5 --steps 2 --warmup 3 --perms-per-step 1 > gpurun_out/bench_r02_final_cfg5_n8.json 2> gpurun_out/bench_r02_final_cfg5_n8.err; echo "cfg5 n8 rc=$?"; grep '^{' gpurun_out/bench_r02_final_cfg5_n8.json | cut -c1-400
