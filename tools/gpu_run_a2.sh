mkdir -p gpurun_out
rm -f gpurun_out/a2_ab.txt
for w in cfg2 cfg2_seam_pole_views cfg4_u16_to_u16 cfg2_other_interp; do for on in 0 1 0 1; do
  R360_COORD_TILES=$on R360_LARGE_PATCH_PASS=$on timeout 300 python bench.py --workload $w --no-variants --no-e2e --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w maps+large=$on', round(d['value']/1e3,1), 'Gpix/s', round(d['ms_per_step'],4), 'ms', d['gpu_launches'], 'launches')" >> gpurun_out/a2_ab.txt
done; done
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/a2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a2_pytest.log
tail -n 4 gpurun_out/a2_pytest.log; cat gpurun_out/a2_ab.txt
