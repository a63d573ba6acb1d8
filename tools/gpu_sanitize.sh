# compute-sanitizer memcheck over the GPU suites that drive the round-2 code paths (B200, under gpurun): per-pixel map
# tiles, the large-patch pass, the batched fallback kernel, the work queue, multi-frame items of every kernel, random
# layouts.  Logs land in gpurun_out/san_*.log.
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/san_smoke.log
for t in fallback_overlap multiframe fuzz; do
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_$t.py -q -x > gpurun_out/san_$t.log 2>&1; echo "$t rc=$?" >> gpurun_out/san_$t.log
done
for f in smoke fallback_overlap multiframe fuzz; do grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/san_$f.log | tail -3; done
