mkdir -p gpurun_out
timeout 300 python tools/shape_sweep.py --interp cubic --fr 2 4 --teams 1 2 --ctas 1 2 --iters 20 2>&1 | grep -v Warning | cut -c1-170 > gpurun_out/f2_sweep.jsonl
timeout 300 python tools/shape_sweep.py --interp linear --fr 2 4 --teams 1 2 --ctas 1 2 --iters 20 2>&1 | grep -v Warning | cut -c1-170 >> gpurun_out/f2_sweep.jsonl
timeout 300 python tools/shape_sweep.py --interp cubic --dtype u16 --frames 8 --fr 2 4 --teams 1 2 --ctas 1 2 --iters 20 2>&1 | grep -v Warning | cut -c1-170 >> gpurun_out/f2_sweep.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/f2_sweep.jsonl'):
    try: d=json.loads(l[:l.rindex('}')+1] if l.rstrip().endswith('}') else l[:l.index(', "sha"')]+'}')
    except Exception: print(l[:150]); continue
    print(d.get('interp'), d.get('dtype'), 'fr', d.get('fr'), 'teams', d.get('teams'), 'ctas', d.get('ctas'), 'pct', d.get('pct'), d.get('Gpix_per_s', d.get('error')))
PY
