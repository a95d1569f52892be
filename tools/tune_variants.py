"""Build tuning variants of the library (compile-time macros) and time the config-2 bench with each.
    python tools/tune_variants.py build  NAME=DEF1,DEF2 ...     (here, no GPU needed)
    python tools/tune_variants.py run    NAME ...               (on the GPU box; prints scan ms / step ms)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "actalker_b200", "lib")

if sys.argv[1] == "build":
    from actalker_b200 import build
    for spec in sys.argv[2:]:
        name, _, defs = spec.partition("=")
        out = build.build(out=os.path.join(LIB, f"libactk_{name}.so"), defs=[d for d in defs.split(",") if d])
        print("built", out)
else:
    extra = []
    names = []
    args = sys.argv[2:]
    while args:                      # NAME ... [--flag value ...] (everything from the first --flag goes to bench.py)
        if args[0].startswith("--"):
            extra = args
            break
        names.append(args.pop(0))
    for name in names:
        env = dict(os.environ)
        if name != "default":
            env["ACTK_LIB_PATH"] = os.path.join(LIB, f"libactk_{name}.so")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", "10", *extra],
                           env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{name:24s} scan {d['roofline']['kernel_ms']:.3f} ms  step {d['ms_per_step']:.3f} ms  "
                  f"merge {d['roofline']['merge_ln_ms']:.3f}", flush=True)
        except Exception:
            print(name, "FAILED", r.stderr[-400:], flush=True)
