timeout 200 python -m pytest "tests/test_conv_gpu.py::test_conv_fused_pool" "tests/test_net_gpu.py::test_pool_fusion_is_bit_identical" "tests/test_fullsize_parity_gpu.py::test_e2e_8s_768x2560_vs_reference" "tests/test_fullsize_parity_gpu.py::test_e2e_8s_768x2560_unfused_blobs" -m gpu -q -x -p no:cacheprovider > gpurun_out/r02q_tests.log 2>&1; tail -4 gpurun_out/r02q_tests.log
for v in A MSCNN_NO_2CTA_POOL; do
  env $v=1 timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > gpurun_out/r02q_bench_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02q_bench_$v.json')); print('$v', round(d['value'],1), d['layers_ms']['conv3_3'], d['layers_ms']['conv3_2'])"
done
