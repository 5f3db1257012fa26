"""A/B of one debug switch on the same box: alternates `VAR` unset / VAR=1 over fresh processes of tools/graph_test.py
and prints the graph-replayed step time of each run and the medians.

    python tools/ab_env.py PHDFX_NO_WPRE 3 [batch [value ...]]      values default to: unset 1
"""
import os
import re
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    var = sys.argv[1]
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    batch = sys.argv[3] if len(sys.argv) > 3 else "256"
    values = sys.argv[4:] or ["unset", "1"]
    res = {v: [] for v in values}
    for r in range(rounds):
        for val in values:
            env = dict(os.environ)
            env.pop(var, None)
            if val != "unset":
                env[var] = val
            out = subprocess.run([sys.executable, str(ROOT / "tools" / "graph_test.py"), batch], env=env,
                                 capture_output=True, text=True, timeout=600)
            m = re.search(r"graph replay ([0-9.]+) us/step", out.stdout)
            if out.returncode != 0 or not m or "graph result equal: True" not in out.stdout:
                print(out.stdout[-2000:], out.stderr[-2000:])
                sys.exit(1)
            res[val].append(float(m.group(1)))
            print(f"round {r} {var}={val}: {m.group(1)} us/step", flush=True)
    for val, xs in res.items():
        print(f"{var}={val}: median {statistics.median(xs):.1f} us/step, min {min(xs):.1f}, runs {xs}")


if __name__ == "__main__":
    main()
