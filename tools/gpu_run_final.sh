# Round-2 evidence run (B200, one gpurun call): DRAM traffic of every bench workload, the bench line with it, the
# full GPU test-suite, launch list + ncu captures (tools/gpu_profile.sh), smoke.  (The file-to-file probe and the CPU
# reference arm do not depend on the kernels: tools/pipeline_probe.py 17 48, bench.py --impl reference.)
mkdir -p gpurun_out
timeout 1200 python tools/traffic_capture.py > gpurun_out/f_traffic.log 2>&1
cp gpurun_out/r02_dram_traffic.json profiles/r02_dram_traffic.json
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
bash tools/gpu_profile.sh > gpurun_out/f_profile.log 2>&1
timeout 300 python tools/shape_sweep.py --interp cubic linear --fr 1 2 4 --iters 20 2>&1 | grep -v Warning | cut -c1-170 > gpurun_out/f_sweep.jsonl
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f_smoke.log 2>&1
tail -n 3 gpurun_out/f_pytest.log; tail -n 2 gpurun_out/f_bench.err; tail -n 2 gpurun_out/f_smoke.log; cut -c1-200 gpurun_out/f_traffic.log; cat gpurun_out/f_sweep.jsonl
