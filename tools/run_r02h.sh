N=${1:-2}
python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02h_mgpu_tests_n$N.log 2>&1; tail -12 gpurun_out/r02h_mgpu_tests_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-bf16 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; tail -c 300 gpurun_out/r02h_bench_n$N.err
$TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-bf16 --resident-only --no-gather > gpurun_out/r02h_bench_n${N}_nogather.json 2> gpurun_out/r02h_bench_n${N}_nogather.err
python - <<PY
import json
for f in ("gpurun_out/r02h_bench_n$N.json", "gpurun_out/r02h_bench_n${N}_nogather.json"):
    d = json.load(open(f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), d.get("e2e", {}).get("value"), d["per_rank_step_ms"])
PY
if [ "$N" = "2" ]; then
  $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h_ref_n$N.json 2> gpurun_out/r02h_ref_n$N.err; cut -c1-400 gpurun_out/r02h_ref_n$N.json
fi
