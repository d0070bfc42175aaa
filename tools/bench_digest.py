#!/usr/bin/env python
"""One line per bench JSON file: the numbers one looks at first."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "FAILED", e)
        continue
    r = d.get("roofline") or {}
    print(path.split("/")[-1], "| impl", d.get("impl", "ours"), "| steps/s %.4f" % d["value"], "| ms/step %.2f" % d["ms_per_step"], "| e2e", (d.get("e2e") or {}).get("value"),
          "| bound", r.get("bound"), r.get("frac"), "| alg-hbm", (r.get("alg_hbm") or {}).get("frac"), "| fp64 exec", ((r.get("candidates") or {}).get("fp64") or {}).get("achieved"),
          "| breakdown", d.get("breakdown_ms"), "| parity", d.get("parity"), "| cpu", d.get("cpu_baseline"), "| shape", d.get("launch_shape"), "| clocks", d.get("clocks"))
