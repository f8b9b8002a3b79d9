"""A/B of diagnostic switches on the bench window: python tools/ab.py [--reps 5] "<ENV=1 ...>" ...   -> median it/s per configuration.
(The trajectory is chaotic beyond ~30 steps -- FP64 atomics -- so only the driver's short window gives comparable work.)"""
import json, os, statistics, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
reps = 5
if args and args[0] == "--reps":
    reps = int(args[1]); args = args[2:]
for cfg in args:
    env = dict(os.environ)
    for kv in cfg.split():
        k, v = kv.split("=", 1); env[k] = v
    vals, e2e, its = [], [], []
    for _ in range(reps):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline", "--stage-steps", "0"], env=env, capture_output=True, text=True)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        vals.append(d["value"]); e2e.append(d["e2e"]["value"]); its.append(int(d["newton_iterations"]))
    print(f"{cfg:50s} value med {statistics.median(vals):7.1f} [{min(vals):.0f}-{max(vals):.0f}]  e2e med {statistics.median(e2e):7.1f}  its {its}", flush=True)
