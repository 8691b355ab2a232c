python bench.py --steps 12 --warmup 3 --no-companions --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 120 python tools/trace_attn.py 32b 2 > gpurun_out/r2i_trace32.log 2>&1
timeout 120 python tools/trace_attn.py 7b 2 > gpurun_out/r2i_trace7.log 2>&1
