mkdir -p gpurun_out
tools/tex_probe > gpurun_out/k_tex_probe.jsonl 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 300 python tools/shape_sweep.py --dtype u16 --frames 8 --fr 1 2 4 --ctas 0 1 --pct 75 2>&1 | grep -v Warning > gpurun_out/k_sweep.jsonl
timeout 600 python tools/pipeline_probe.py 17 24 > gpurun_out/k_pipeline.jsonl 2> gpurun_out/k_pipeline.err
rm -f gpurun_out/k_l2.txt
for cfg in "0 128" "1 128" "2 128" "0 64" "0 0" "0 256" "1 0"; do
  set -- $cfg
  R360_L2_POLICY=$1 R360_L2_PROMO=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:remap_tiled -s 3 -c 1 --csv --log-file /tmp/l2.csv python tools/shape_sweep.py --interp cubic --fr 2 --iters 1 > /dev/null 2>&1
  echo "policy=$1 promo=$2 $(grep -v '^==' /tmp/l2.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
print(' '.join('%s=%s' % (dict(zip(h,r))['Metric Name'].split('__')[-1], dict(zip(h,r))['Metric Value']) for r in rows[1:]))")" >> gpurun_out/k_l2.txt
  R360_L2_POLICY=$1 R360_L2_PROMO=$2 timeout 120 python tools/shape_sweep.py --interp cubic linear --fr 2 --iters 10 2>&1 | grep -v Warning | cut -c1-160 >> gpurun_out/k_l2.txt
done
tail -n 4 gpurun_out/k_pytest.log
