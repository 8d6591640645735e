N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-bf16 > gpurun_out/r02i_bench_n$N.json 2> gpurun_out/r02i_bench_n$N.err; tail -c 200 gpurun_out/r02i_bench_n$N.err
$TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-bf16 --resident-only --no-gather > gpurun_out/r02i_bench_n${N}_nogather.json 2> gpurun_out/r02i_bench_n${N}_nogather.err
$TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --no-bf16 --resident-only > gpurun_out/r02i_bench_n${N}_gather2.json 2> gpurun_out/r02i_bench_n${N}_gather2.err
python bench.py --gpus 1 --steps 30 --warmup 5 --no-bf16 --resident-only --no-cpu-baseline > gpurun_out/r02i_bench_n1_samebox.json 2> gpurun_out/r02i_bench_n1_samebox.err
python - <<PY
import json
for f in ("gpurun_out/r02i_bench_n$N.json", "gpurun_out/r02i_bench_n${N}_nogather.json", "gpurun_out/r02i_bench_n${N}_gather2.json", "gpurun_out/r02i_bench_n1_samebox.json"):
    d = json.load(open(f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), d.get("e2e", {}).get("value"))
    for r in (d.get("per_rank_step_ms") or []): print("   ", r)
PY
python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -p no:cacheprovider -k cpp > gpurun_out/r02i_mgpu_tests_n$N.log 2>&1; tail -14 gpurun_out/r02i_mgpu_tests_n$N.log
