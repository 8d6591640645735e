# full GPU suite first (deferred rows, graph, a15 vectors)
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02g_tests.log 2>&1; tail -4 gpurun_out/r02g_tests.log
python tools/bench_configs.py --steps 10 --warmup 3 > gpurun_out/r02g_configs.json 2> gpurun_out/r02g_configs.err; tail -c 300 gpurun_out/r02g_configs.err
# --set full captures
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/r02_conv3_2_pair python tools/profile_one.py conv3_2 > gpurun_out/r02g_ncu1.log 2>&1
$NCU -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/r02_conv1_2_vpool python tools/profile_one.py conv1_2 > gpurun_out/r02g_ncu2.log 2>&1
$NCU -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/r02_conv2_2_vpool python tools/profile_one.py conv2_2 > gpurun_out/r02g_ncu3.log 2>&1
$NCU -k regex:"box_decode|box_topk|nms_mask|nms_scan|box_finalize|head_gather|pool_kernel|roi_pool|detect_|bbnms" -s 21 -c 21 -o gpurun_out/r02_noconv_8s python tools/profile_one.py net8s 2 > gpurun_out/r02g_ncu4.log 2>&1
$NCU -k regex:"deconv2x" -s 1 -c 1 -o gpurun_out/r02_deconv2x python tools/profile_one.py net7s2x 2 > gpurun_out/r02g_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
