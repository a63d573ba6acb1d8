mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum
rm -f gpurun_out/e2_traffic.txt
run() {  # label, args to shape_sweep
  label=$1; shift
  timeout 300 ncu --metrics $M --clock-control none -k regex:remap_tiled -s 3 -c 1 --csv --log-file /tmp/q.csv python tools/shape_sweep.py --iters 1 "$@" > /dev/null 2>&1
  echo "$label $(grep -v '^==' /tmp/q.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print(' '.join('%s=%s' % (dict(zip(h,r))['Metric Name'].split('__')[-1], dict(zip(h,r))['Metric Value']) for r in rows[1:]))")" >> gpurun_out/e2_traffic.txt
}
run linear_default --interp linear --fr 0
run cubic_default --interp cubic --fr 0
run u16_default --interp cubic --dtype u16 --frames 8 --fr 0
R360_ORDER_BAND=64 run linear_band64 --interp linear --fr 0
R360_ORDER_BAND=256 run linear_band256 --interp linear --fr 0
R360_ORDER_BAND=64 run cubic_band64 --interp cubic --fr 0
timeout 300 python tools/shape_sweep.py --interp linear cubic --fr 0 2 4 --iters 20 2>&1 | grep -v Warning | cut -c1-170 > gpurun_out/e2_sweep.jsonl
timeout 120 python tools/shape_sweep.py --interp cubic --dtype u16 --frames 8 --fr 0 --iters 20 2>&1 | grep -v Warning | cut -c1-170 >> gpurun_out/e2_sweep.jsonl
for band in 64 256; do R360_ORDER_BAND=$band timeout 120 python tools/shape_sweep.py --interp linear cubic --fr 0 --iters 20 2>&1 | grep -v Warning | cut -c1-170 | sed "s/^/band=$band /" >> gpurun_out/e2_sweep.jsonl; done
timeout 1500 python -m pytest tests/test_gpu_multiframe.py tests/test_gpu_fallback_overlap.py tests/test_gpu_configs.py tests/test_gpu_parity.py -q -x > gpurun_out/e2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e2_pytest.log
cat gpurun_out/e2_traffic.txt; cat gpurun_out/e2_sweep.jsonl; tail -n 3 gpurun_out/e2_pytest.log
