"""A/B of an environment switch inside ONE gpurun call (boxes differ by a few % in power-capped clocks): runs
`python tools/op_breakdown.py <B> unet` alternately with VAR=a and VAR=b and prints the graph-replay times.

    python tools/ab_env.py TQ_PDL 0 1 [reps]
"""
import os
import re
import subprocess
import sys

var, a, b = sys.argv[1:4]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
res = {a: [], b: []}
for _ in range(reps):
    for val in (a, b):
        env = dict(os.environ, **{var: val})
        out = subprocess.run([sys.executable, "tools/op_breakdown.py", "256", "unet"], env=env, capture_output=True, text=True).stdout
        m = re.search(r"graph replay: ([\d.]+) ms", out)
        res[val].append(float(m.group(1)) if m else float("nan"))
        print(f"{var}={val}: graph replay {res[val][-1]:.3f} ms", flush=True)
for val in (a, b):
    v = sorted(res[val])
    print(f"# {var}={val}: median {v[len(v) // 2]:.3f} ms  all {v}")
