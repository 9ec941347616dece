#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/e22_pytest.txt
cat gpurun_out/e22_pytest.txt
HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e22_new.json 2> gpurun_out/e22_new.err
HB_BENCH_FOLD_SCALE=64 timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e22_new_hard.json 2> gpurun_out/e22_new_hard.err
python - <<'PY'
import json
for f in ('e22_new','e22_new_hard'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -h -i "phase\|rounds" gpurun_out/e22_new.err | tail -3
grep -h -i "phase\|rounds" gpurun_out/e22_new_hard.err | tail -3
