for s in "" "2,4,8,12,12,12,8" "1,2,4,8,12,14,14" "4,8,10,10,10,10,6" "2,6,10,14,14,10" "3,7,13,13,13" "62" "1,3,6,10,14,14,10"; do
  if [ -z "$s" ]; then unset GCB_SVL_BATCHES; else export GCB_SVL_BATCHES="$s"; fi; python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$s', d['ms_per_step'], d['field_kernel']['kernel_ms'], d['e2e']['ms_per_step'])"
done
