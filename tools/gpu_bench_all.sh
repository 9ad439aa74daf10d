#!/bin/bash
# Round-end bench lines on one B200 (under gpurun): the default line (C2), the reference arm, the other configs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --files-full > $O/r2_bench_c2.json 2> $O/bench_c2.err; tail -n 1 $O/bench_c2.err | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/bench_ref.err; tail -n 1 $O/bench_ref.err | cut -c1-200
for c in C1 C3 C4 C5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_$c.json 2> $O/bench_$c.err; tail -n 1 $O/bench_$c.err | cut -c1-200
done
python - <<'PY'
import json
for c in ['c2','reference','C1','C3','C4','C5']:
    try:
        j=json.loads(open('gpurun_out/r2_bench_%s.json' % c).read().strip().splitlines()[-1])
        ph=(j.get('roofline') or {}).get('sliced'); ph = ph and {k:round(v['ms_per_step'],2) for k,v in ph['phases'].items()}
        print(c, round(j['value']/1e9,3), round(j['ms_per_step'],2), j.get('parity_checked'), ph, (j.get('roofline') or {}).get('frac'), (j.get('e2e') or {}).get('value'), j.get('step_wall_ms_rank0'), j.get('clocks'))
        if c == 'c2': print('  cpu', j['cpu_baseline'], '\n  files', j['e2e_from_files'])
    except Exception as e: print(c, 'ERR', e)
PY
