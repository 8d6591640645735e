python -m pytest tests/test_detect_gpu.py tests/test_multigpu_gpu.py tests/test_net_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02e_tests.log 2>&1; tail -3 gpurun_out/r02e_tests.log
for v in default MSCNN_ROI_BLOCK_PER_ROI; do
  if [ $v = default ]; then env_s="A=1"; else env_s="$v=1"; fi
  env $env_s python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > gpurun_out/r02e_bench_$v.json 2>gpurun_out/r02e_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_$v.json')); print('$v', d['value'], d['layers_ms']['roi_pool_org'], d['layers_ms']['roi_c1'], d['layers_ms']['proposals'])"
done
MSCNN_ROI_BLOCK_PER_ROI=1 ncu -k regex:roi_pool_kernel --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -c 2 --csv --log-file gpurun_out/r02e_roi_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > /dev/null 2>&1
ncu -k regex:"nms_scan_kernel|box_topk_kernel|box_finalize_kernel|nms_mask_kernel|box_decode_kernel|detect_" --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/r02e_box_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > /dev/null 2>&1
