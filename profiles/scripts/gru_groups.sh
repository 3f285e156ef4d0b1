python -m pytest tests -m gpu -x -q -k "gru" 2>&1 | tail -3
for g in 1 2 4 8; do
  python bench.py --workload c3_gru_d256_seq100_items5M_bpr5_b2048 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --set gru_row_groups=$g 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('groups $g: value %.0f samples/s  %.4f ms/step  e2e %.4f ms' % (j['value'], j['ms_per_step'], j['e2e']['ms_per_step']))"
done
