for pg in 16 64 128; do
ASD_PAGE=$pg timeout 300 python tools/bench_attn.py > gpurun_out/r2e_attn_$pg.log 2>&1
done
timeout 120 python tools/trace_attn.py 72b-tp4 2 page=128 > gpurun_out/r2e_trace_tc128.log 2>&1
timeout 120 python tools/trace_attn.py 32b 2 page=128 > gpurun_out/r2e_trace_tc32_128.log 2>&1
timeout 120 python tools/trace_attn.py 7b 2 page=128 > gpurun_out/r2e_trace_tc7_128.log 2>&1
