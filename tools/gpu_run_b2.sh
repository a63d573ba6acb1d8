mkdir -p gpurun_out
rm -f gpurun_out/b2_sweep.jsonl
for lib in shipped tools/variants/lib_s8.so; do
  if [ "$lib" = shipped ]; then unset R360_LIBRARY; else export R360_LIBRARY=$PWD/$lib; fi
  timeout 300 python tools/shape_sweep.py --interp linear --fr 2 4 --teams 1 2 --pct 50 75 100 2>&1 | grep -v Warning >> gpurun_out/b2_sweep.jsonl
  timeout 300 python tools/shape_sweep.py --interp cubic --fr 2 4 --teams 1 2 --pct 75 2>&1 | grep -v Warning >> gpurun_out/b2_sweep.jsonl
  timeout 300 python tools/shape_sweep.py --interp cubic --dtype u16 --frames 8 --fr 2 4 --teams 1 2 --pct 75 2>&1 | grep -v Warning >> gpurun_out/b2_sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/b2_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d.get('library'), d.get('interp'), d.get('dtype'), 'fr', d.get('fr'), 'teams', d.get('teams'), 'pct', d.get('pct'), d.get('Gpix_per_s', d.get('error')))
PY
