timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_test_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_test_all.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
