"""Option sweeps of the specialised kernels through bench.py (device-resident timing of Phi a / Phi^T).

  python tools/sweep_shapes.py "OPTS[|bench args]" ...      e.g.  "1,4,2,80,8,1,4,16,56|--config c4share"  "1,2,4|--rows 125000"

OPTS is a prefix of OB_SPEC_OPTS (ra,qa,tga,cache_a,wt,rt,pt,cache_t,acc_cap,np,mc,ut,mw,kc,qd,tgd,cache_d,maxcols_d,
nreg_c,nreg_p,psleep); missing fields keep their defaults.  One line per run; results of round 2: profiles/r02_spec_sweeps.txt."""
import json, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cfg, extra):
    env = dict(os.environ, OB_SPEC_OPTS=cfg)
    try:
        p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "6", "--warmup", "3", "--no-optcg", "--no-cpu-baseline"] + extra,
                           env=env, capture_output=True, text=True, timeout=150)
    except subprocess.TimeoutExpired:
        print(" ".join(extra) or "c3", cfg, "TIMEOUT", flush=True)
        return
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(" ".join(extra) or "c3", cfg, "pairs/s %.1f" % d["value"], "phi_a %.4f ms" % d["roofline"]["ms_phi_a"], "phi_t %.4f ms" % d["roofline"]["ms_phi_t"], flush=True)
    except Exception:
        print(cfg, "FAILED", p.stderr[-300:], flush=True)


for arg in sys.argv[1:]:
    cfg, _, extra = arg.partition("|")
    run(cfg, extra.split())
