"""Per-stage A/B (sb_profile_stages: a stream sync at every stage boundary, so stage times do not depend on the trajectory's
iteration mix): python tools/stage_ab.py "<ENV=1 ...>" ...   -> ms per call of the main stages, steps 8-20 of the bench scene."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cfg in sys.argv[1:]:
    env = dict(os.environ)
    for kv in cfg.split():
        k, v = kv.split("=", 1); env[k] = v
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "stage_profile.py"), "tetdrop", "26", "8", "12"], env=env, capture_output=True, text=True)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    keys = ["intersections", "contact_update", "eval_pgh", "project_to_pd", "assembly_symbolic", "assembly_numeric", "pcg"]
    nd = max(1, d["intersections"]["calls"] + d["contact_update"]["calls"])
    print("   per detection: " + " ".join(f"{k}={d[k]['calls'] / nd:.0f}" for k in ["tile_pairs_pt", "tile_pairs_ee", "tile_pairs_et", "candidates_pt", "candidates_ee", "candidates_et"] if k in d))
    print(f"{cfg:40s} its {d.get('newton_iterations')} " + " ".join(f"{k}={1e3 * d[k]['ms'] / max(1, d[k]['calls']):.0f}us x{d[k]['calls']}" for k in keys if k in d), flush=True)
