mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
rm -f gpurun_out/s_traffic.txt
for band in 1 32 64 128 256; do for interp in linear cubic; do
  R360_ORDER_BAND=$band timeout 300 ncu --metrics $M --clock-control none -k regex:remap_tiled -s 3 -c 1 --csv --log-file /tmp/q.csv python tools/shape_sweep.py --interp $interp --fr 4 --iters 1 > /dev/null 2>&1
  echo "$interp band=$band $(grep -v '^==' /tmp/q.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print(' '.join('%s=%s' % (dict(zip(h,r))['Metric Name'].split('__')[-1], dict(zip(h,r))['Metric Value']) for r in rows[1:]))")" >> gpurun_out/s_traffic.txt
  R360_ORDER_BAND=$band timeout 120 python tools/shape_sweep.py --interp $interp --fr 4 --iters 10 2>&1 | grep -v Warning | cut -c1-150 >> gpurun_out/s_traffic.txt
done; done
