timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-bf16 > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err; echo rc=$?; tail -c 300 gpurun_out/r02p_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange'], d['config']['parallelism'][:60]); print(d['per_rank_step_ms'])"
