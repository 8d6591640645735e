python -m pytest tests/test_detect_gpu.py tests/test_net_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02d_tests.log 2>&1; tail -3 gpurun_out/r02d_tests.log
for v in default MSCNN_NO_ROI_HINTS MSCNN_NO_ROI_HALVES both; do
  if [ $v = default ]; then env_s=""; elif [ $v = both ]; then env_s="MSCNN_NO_ROI_HINTS=1 MSCNN_NO_ROI_HALVES=1"; else env_s="$v=1"; fi
  env $env_s python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > gpurun_out/r02d_bench_$v.json 2>gpurun_out/r02d_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_$v.json')); print('$v', d['value'], d['layers_ms']['roi_pool_org'], d['layers_ms']['roi_c1'])"
done
ncu -k regex:roi_pool_kernel --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 3 --csv --log-file gpurun_out/r02d_roi_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > /dev/null 2>&1
grep roi_pool gpurun_out/r02d_roi_ncu.csv | cut -d, -f13-15 | tail -8
MSCNN_NO_ROI_HINTS=1 MSCNN_NO_ROI_HALVES=1 ncu -k regex:roi_pool_kernel --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 3 --csv --log-file gpurun_out/r02d_roi_ncu_old.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-bf16 --resident-only > /dev/null 2>&1
grep roi_pool gpurun_out/r02d_roi_ncu_old.csv | cut -d, -f13-15 | tail -4
