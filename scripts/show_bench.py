"""Print the headline fields of a bench.py JSON line (last line of the file given)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("n_gpus", "ms_per_step", "value", "steps", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"))
if d.get("multi_rank_checks"):
    print(d["multi_rank_checks"])
t = d.get("allreduce_timeline") or {}
if t.get("per_rank"):
    print([(r["rank"], r["backward_end_ms"], r["exposed_ms"]) for r in t["per_rank"]])
