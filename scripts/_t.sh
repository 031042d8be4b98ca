timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 30 --warmup 5 --no-scale-section --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('poisson', round(d['ms_per_step'],4), round(d['roofline']['achieved']))
"
timeout 200 python scripts/bench_extra.py 2>/dev/null | grep '"spgemm"' | cut -c1-200
for pf in 0 6; do SPAM_MERGE_PF=$pf timeout 300 python bench.py --workload rmat22 --steps 5 --warmup 3 --no-scale-section --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rmat22 pf', $pf, round(d['ms_per_step'],3))
"; done
timeout 300 python bench.py --workload rmat22 --steps 5 --warmup 3 --no-scale-section --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rmat22 auto', round(d['ms_per_step'],3))
"
