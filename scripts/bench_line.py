import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    ph = {k: round(v, 3) for k, v in d["phases_ms"].items()}
    print(f'{d["config"]["workload"]}: ms/step={d["ms_per_step"]:.3f} GFLOP/s={d["value"]:.1f} pipeline_GB/s={d["hbm_gbs_pipeline"]:.0f} phases={ph} launches={d["gpu_launches"]} e2e={d.get("e2e") and round(d["e2e"]["ms_per_step"],2)}')
except Exception as e:
    print("FAIL", e)
