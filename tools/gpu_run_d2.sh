mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum
rm -f gpurun_out/d2_traffic.txt
run() {  # label, args to shape_sweep
  label=$1; shift
  timeout 300 ncu --metrics $M --clock-control none -k regex:remap_tiled -s 3 -c 1 --csv --log-file /tmp/q.csv python tools/shape_sweep.py --iters 1 "$@" > /dev/null 2>&1
  echo "$label $(grep -v '^==' /tmp/q.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print(' '.join('%s=%s' % (dict(zip(h,r))['Metric Name'].split('__')[-1], dict(zip(h,r))['Metric Value']) for r in rows[1:]))")" >> gpurun_out/d2_traffic.txt
}
run linear_default --interp linear --fr 4
run linear_ctas1 --interp linear --fr 4 --ctas 1
run linear_teams2_ctas1 --interp linear --fr 4 --teams 2 --ctas 1
run linear_fr2 --interp linear --fr 2
run linear_fr1 --interp linear --fr 1
run linear_frames4 --interp linear --fr 4 --frames 4
run cubic_default --interp cubic --fr 4
run cubic_teams1_ctas2 --interp cubic --fr 4 --teams 1 --ctas 2
run cubic_frames4 --interp cubic --fr 4 --frames 4
R360_BOX_FAMILY=1 run linear_family1_ctas1 --interp linear --fr 4 --ctas 1
R360_ORDER_BAND=64 run linear_band64 --interp linear --fr 4
R360_ORDER_BAND=32 run linear_band32_ctas1 --interp linear --fr 4 --ctas 1
cat gpurun_out/d2_traffic.txt
