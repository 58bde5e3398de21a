"""Print the headline of a bench.py JSON line (A/B runs)."""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 3), "ms_per_step", round(d["ms_per_step"], 1), "batch", d["config"]["batch_per_gpu"], "lanes",
          d["config"]["lanes_per_gpu"], "launches", d["gpu_launches"], "jumps", d["counters"].get("jumps"))
except Exception as e:  # noqa: BLE001
    print(sys.argv[1], "unreadable:", e)
