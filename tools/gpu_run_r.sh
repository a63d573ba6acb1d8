mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
rm -f gpurun_out/r_traffic.txt
for cfg in "1 linear" "0 linear" "1 cubic" "0 cubic"; do
  set -- $cfg
  R360_WALK_CHUNKED=$1 timeout 300 ncu --metrics $M --clock-control none -k regex:remap_tiled -s 3 -c 1 --csv --log-file /tmp/q.csv python tools/shape_sweep.py --interp $2 --fr 4 --iters 1 > /dev/null 2>&1
  echo "$2 chunked=$1 $(grep -v '^==' /tmp/q.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print(' '.join('%s=%s' % (dict(zip(h,r))['Metric Name'].split('__')[-1], dict(zip(h,r))['Metric Value']) for r in rows[1:]))")" >> gpurun_out/r_traffic.txt
  R360_WALK_CHUNKED=$1 timeout 120 python tools/shape_sweep.py --interp $2 --fr 1 2 4 --iters 10 2>&1 | grep -v Warning | cut -c1-150 >> gpurun_out/r_traffic.txt
done
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multiframe.py tests/test_gpu_configs.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
tail -n 3 gpurun_out/r_pytest.log
