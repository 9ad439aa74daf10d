#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:skm_build" -s 2 -c 1 -f -o $O/c5_build python bench.py --pairs 4000000 --transcripts 3200 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/c5_ncu.log 2>&1
# instrumented library (phase clocks)
cd soapdenovo-trans_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -DSDT_BUILD_PROF -shared -o libsdtgpu.so csrc/sdtgpu.cu csrc/sdt_synth.cu host/kmerset_builder.cpp host/sdt_readpack.c -lpthread 2> /dev/null; cd ..
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/c5_bench_prof.json 2> $O/c5_bench_prof.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c5_bench_prof.json').read().strip().splitlines()[-1])
ph=j['roofline']['sliced']['phases']
print({k:round(v['ms_per_step'],2) for k,v in ph.items()})
p=j['roofline']['sliced']['build_phase_cycles']; print(p)
n=p['items']
for k in ('prepare','insert','compact'): print(k, p[k]/n, 'cycles/item', p[k]/n/1.965e3, 'us')
PY
