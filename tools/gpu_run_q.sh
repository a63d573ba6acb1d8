mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
rm -f gpurun_out/q2_traffic.txt
for cfg in "0 linear" "32 linear" "64 linear" "96 linear" "64 cubic"; do
  set -- $cfg
  R360_L2_PERSIST_MB=$1 timeout 300 ncu --metrics $M --clock-control none -k regex:remap_tiled -s 3 -c 1 --csv --log-file /tmp/q.csv python tools/shape_sweep.py --interp $2 --fr 4 --iters 1 > /dev/null 2>&1
  echo "$2 persist_MB=$1 $(grep -v '^==' /tmp/q.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print(' '.join('%s=%s' % (dict(zip(h,r))['Metric Name'].split('__')[-1], dict(zip(h,r))['Metric Value']) for r in rows[1:]))")" >> gpurun_out/q2_traffic.txt
  R360_L2_PERSIST_MB=$1 timeout 120 python tools/shape_sweep.py --interp $2 --fr 4 --iters 10 2>&1 | grep -v Warning | cut -c1-150 >> gpurun_out/q2_traffic.txt
done
