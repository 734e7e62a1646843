#!/usr/bin/env python
"""bench.py's per-kernel numbers only (development aid): python tools/quick_c3.py [label]"""
import json, subprocess, sys, os
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "..", "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline"],
                     capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(sys.argv[1] if len(sys.argv) > 1 else "", "step %.4f e2e %.3f (copy bound %.3f)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["copy_bound_ms"]),
          " ".join("%s=%.4f" % (k["kernel"].split("[")[1].rstrip("]"), k["ms"]) for k in d["roofline"]["per_kernel"]))
except Exception as e:
    print("failed", e, out.stderr[-500:])
