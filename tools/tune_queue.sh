for q in 512 1024 1536 2048 3072 4096; do
  echo -n "queue_windows $q: "
  SDTGPU_QUEUE_WINDOWS=$q python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'])"
done
