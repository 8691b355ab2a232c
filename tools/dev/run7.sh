timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2t_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_test.log
for t in 4 8 2; do ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|attn|norm|rope|reduce|gather" -s 38 -c 19 --csv --log-file gpurun_out/r2t_c5shard$t.csv python tools/dev/c5_shard.py $t > gpurun_out/r2t.log 2>&1; done
