#!/bin/bash
# in-order phase S: parity tests, then A/B against the round-based phase S (HB_DEBUG=128), easy and hard regime
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/e19_pytest.txt
cat gpurun_out/e19_pytest.txt
for v in new old; do
  dbg=0; [ $v = old ] && dbg=128
  HB_DEBUG=$dbg HB_PHASES=1 timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e19_$v.json 2> gpurun_out/e19_$v.err
  HB_DEBUG=$dbg HB_BENCH_FOLD_SCALE=64 timeout 300 python bench.py --no-cpu --steps 10 --warmup 5 > gpurun_out/e19_${v}_hard.json 2> gpurun_out/e19_${v}_hard.err
done
python - <<'PY'
import json
for f in ('e19_new','e19_old','e19_new_hard','e19_old_hard'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -h -i "phase\|rounds" gpurun_out/e19_new.err | tail -6
