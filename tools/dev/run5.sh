timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2g_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_test.log
timeout 300 python tools/bench_attn.py > gpurun_out/r2g_attn.log 2>&1
timeout 120 python tools/trace_attn.py 72b-tp4 2 > gpurun_out/r2g_trace_tc.log 2>&1
