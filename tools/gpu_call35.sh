#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
run() { timeout 600 env $1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $2 > $O/c35.json 2> $O/c35.err; tail -n 1 $O/c35.err | cut -c1-200
python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open('gpurun_out/c35.json').read().strip().splitlines()[-1])
    ph=j['roofline']['sliced']['phases']; g=j['roofline']['sliced']['geometry']
    print(sys.argv[1], round(j['value']/1e9,2), round(j['ms_per_step'],2), j['parity_checked'], {k:round(v['ms_per_step'],2) for k,v in ph.items()}, g['n_slices'], g['work_items'], g['retried_items'], g['n_records_merged'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
cd soapdenovo-trans_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -DSDT_NO_MATCH -shared -o libsdtgpu.so csrc/sdtgpu.cu csrc/sdt_synth.cu host/kmerset_builder.cpp host/sdt_readpack.c -lpthread 2>/dev/null; cd ..
run "X=nomatch" ""
run "X=nomatchC5" "--config C5 --no-parity"
