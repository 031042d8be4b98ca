timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ewise or add or sub or elementwise or fuzz or smoke" 2>&1 | tail -3
for t in 1 0; do SPAM_EWISE_TMA=$t timeout 200 python scripts/bench_extra.py ewise 2>/dev/null | cut -c1-330; done
