mkdir -p gpurun_out
timeout 300 python tools/shape_sweep.py --fr 1 2 4 --teams 1 2 --ctas 0 2>&1 | grep -v Warning > gpurun_out/n_sweep.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
timeout 900 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?" >> gpurun_out/n_bench.err
tail -n 4 gpurun_out/n_pytest.log; tail -n 2 gpurun_out/n_bench.err
