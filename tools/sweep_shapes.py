import json, os, subprocess, sys
REPO="/root/repo"
def run(cfg, extra):
    env = dict(os.environ, OB_SPEC_OPTS=cfg)
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "6", "--warmup", "3", "--no-optcg", "--no-cpu-baseline"] + extra, env=env, capture_output=True, text=True, timeout=300)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(" ".join(extra), cfg, "pairs/s %.1f" % d["value"], "phi_a %.4f ms" % d["roofline"]["ms_phi_a"], "phi_t %.4f ms" % d["roofline"]["ms_phi_t"], flush=True)
    except Exception:
        print(cfg, "FAILED", p.stderr[-300:], flush=True)
base="1,4,2,80,8,1,4,16,%d,4,0,1,8,16,2,3,10,96,232,40,128"
for cap in (56, 72, 92, 112):
    run(base % cap, ["--config", "c4share"])
for cfg in ("1,4,2,80,8,1,4,16,56,4,0,1,8,16,2,3,10,96,232,40,128", "1,2,4,80,8,1,4,16,56,4,0,1,8,16,2,3,10,96,232,40,128", "1,4,2,80,8,1,2,16,56,4,0,1,8,16,2,3,10,96,232,40,128", "1,2,4,80,8,1,2,16,56,4,0,1,8,16,2,3,10,96,232,40,128"):
    run(cfg, ["--rows", "125000"])
