N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { tag=$1; shift; timeout 110 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 40 --warmup 5 --no-bf16 --resident-only "$@" > gpurun_out/r02m_n${N}_$tag.json 2> gpurun_out/r02m_n${N}_$tag.err; }
run peer1
MSCNN_XCHG_FLAGS_ONLY=1 run peer_flagsonly
run nogather1 --no-gather
python - <<PY
import json
for tag in ("peer1", "peer_flagsonly", "nogather1"):
    f = "gpurun_out/r02m_n${N}_%s.json" % tag
    try:
        d = json.load(open(f)); print(tag, round(d["value"], 1), round(d["ms_per_step"], 3), [(r["sum"], r["min"], r["sm_mhz"], r["power_w_max"]) for r in d["per_rank_step_ms"]])
    except Exception as e: print(tag, "FAILED", e)
PY
