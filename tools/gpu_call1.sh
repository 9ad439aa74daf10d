#!/bin/bash
# round 2, GPU call 1: baseline numbers + fresh profiles of the r1 code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/c1_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "C1 or C2 or C5" > $O/c1_fullsize.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 > $O/c1_bench_base.json 2> $O/c1_bench_base.err
SDTGPU_BUILD_NT=512 SDTGPU_SLICE_SLOTS=1531 timeout 400 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/c1_bench_nt512.json 2> $O/c1_bench_nt512.err
SDTGPU_RESPLIT=1 timeout 300 compute-sanitizer --tool initcheck --track-unused-memory no python -m pytest tests/test_gpu_sliced.py -x -q -m gpu -k "tiny_slices and 31-1" > $O/c1_initcheck.log 2>&1
SDTGPU_RESPLIT=1 timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sliced.py -x -q -m gpu -k "tiny_slices and 31-1" > $O/c1_memcheck.log 2>&1
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:skm_ --csv --log-file $O/c1_traffic_c2.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/c1_traffic_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:skm_emit|skm_dedupe|skm_scatter|skm_build" -s 6 -c 6 -f -o $O/c1_skm python bench.py --pairs 4000000 --transcripts 3200 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/c1_ncu_full.log 2>&1
ls -la $O | tail -20
