"""profiles/traffic.json: DRAM bytes per launch of every kernel, from the ncu launch lists of one steady-state step
(scripts/one_step.py under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`)."""
import json
import sys

sys.path.insert(0, "scripts")
from ncu_launches import load, short  # noqa: E402

out = {}
for path in sys.argv[2:]:
    for d in load(path).values():
        k = short(d["name"]).split("<")[0]
        r = out.setdefault(k, dict(launches=0, dram_bytes=0.0, time_us=0.0))
        r["launches"] += 1
        r["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        r["time_us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
for r in out.values():
    r["dram_bytes_per_launch"] = r["dram_bytes"] / r["launches"]
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
