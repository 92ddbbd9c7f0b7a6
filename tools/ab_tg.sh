cd /root/repo
for tag in "" _oct; do
  export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200$tag.so
  echo "== lib$tag"
  python bench.py --steps 10 --warmup 3 --no-cpu --c5 off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'fill ms',round(r['kernel_ms_per_step'],3),'fill gcups',round(r['fill_only_gcups']),'verified',d['verified'])"
done
export SA_B200_LIB=/root/repo/sequence-alignment-gpu_b200/libsa_b200_oct.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or goldens_default or known_answer or ties or identity or spectrum" 2>&1 | tail -3
