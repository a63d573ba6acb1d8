mkdir -p gpurun_out
timeout 300 python tools/shape_sweep.py --interp linear cubic --dtype u16 --frames 8 --fr 2 --teams 1 2 2>&1 | grep -v Warning | cut -c1-200 > gpurun_out/c2_sweep.jsonl
timeout 900 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?" >> gpurun_out/c2_bench.err
R360_PROBE_WORKERS=8,16 timeout 900 python tools/pipeline_probe.py 17 48 > gpurun_out/c2_pipeline.jsonl 2> gpurun_out/c2_pipeline.err
cat gpurun_out/c2_sweep.jsonl; tail -n 3 gpurun_out/c2_bench.err; cut -c1-420 gpurun_out/c2_pipeline.jsonl
