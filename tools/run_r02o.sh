timeout 300 python -m pytest tests/test_multigpu_gpu.py tests/test_postprocess_vectors.py -m gpu -q -p no:cacheprovider > gpurun_out/r02o_tests.log 2>&1; tail -3 gpurun_out/r02o_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; tail -2 gpurun_out/r02o_smoke.log
