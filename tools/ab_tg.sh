cd /root/repo
for pairs in 125000 250000; do
for mc in 2 3 4 6; do
  echo "== pairs $pairs SA_BATCH_MIN_CHUNKS=$mc"
  SA_BATCH_MIN_CHUNKS=$mc python bench.py --pairs $pairs --steps 20 --warmup 3 --no-cpu --c5 off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'fill ms',round(r['kernel_ms_per_step'],3),'verified',d['verified']['mismatches'])"
done; done
