python -m pytest tests/test_graph_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02f_graph.log 2>&1; tail -15 gpurun_out/r02f_graph.log
python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_graph_gpu.py > gpurun_out/r02f_tests.log 2>&1; tail -6 gpurun_out/r02f_tests.log
python tools/bench_configs.py --steps 10 --warmup 3 > gpurun_out/r02f_configs.json 2> gpurun_out/r02f_configs.err; tail -c 400 gpurun_out/r02f_configs.err
python -c "
import json; d=json.load(open('gpurun_out/r02f_configs.json'))
for c in ('config1','config4'):
    for m in ('fp32_faithful','bf16'):
        x=d[c][m]; print(c,m,round(x['ms_per_step'],3),x['graph'],x['kernel_launches_per_step'],x['top_layers_ms'][:4])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['serial_value'], d['e2e_images']['value'])"
