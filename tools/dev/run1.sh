for mb in 0 48 96 160; do
  python bench.py --steps 12 --warmup 3 --no-companions --no-cpu-baseline --opt attn_prefetch_mb=$mb > gpurun_out/r2c_pf_$mb.json 2> gpurun_out/r2c_pf_$mb.err
done
python tools/trace_attn.py 72b-tp4 2 > gpurun_out/r2c_trace_tc.log 2>&1
python tools/trace_attn.py 72b-tp4 1 > gpurun_out/r2c_trace_mma.log 2>&1
python tools/trace_attn.py 32b 1 > gpurun_out/r2c_trace_mma32.log 2>&1
python tools/trace_attn.py 7b 1 > gpurun_out/r2c_trace_mma7.log 2>&1
