"""Timings of the other BASELINE.json configs (C1, C3, C4, C5-like) through the public API on one B200.
Not the bench line (that is C2, bench.py); these show how the same kernels behave off the headline shape."""
import json, subprocess, sys, os
here = os.path.dirname(os.path.abspath(__file__))
res = []
for name, steps in (("C1", 20), ("C3", 10), ("C3cow", 10), ("C4", 20), ("C4ambient", 20), ("C5", 3)):
    r = subprocess.run([sys.executable, os.path.join(here, "run_config.py"), name, str(steps)], capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    res.append(json.loads(line[-1]) if line else {"config": name, "error": r.stderr[-400:]})
print(json.dumps(res, indent=1))
