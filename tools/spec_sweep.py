"""Sweep OB_SPEC_OPTS configurations of the terms-specialised kernels at config C3 (one process per
configuration: the options are read when a module is generated).  usage: spec_sweep.py cfg [cfg ...]"""
import json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cfg in sys.argv[1:]:
    env = dict(os.environ, OB_SPEC_OPTS=cfg)
    try:
      p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "6", "--warmup", "3", "--no-optcg", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True, timeout=120)
    except subprocess.TimeoutExpired:
      print(cfg, "TIMEOUT", flush=True)
      continue
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(cfg, "pairs/s %.1f" % d["value"], "phi_a %.3f ms" % d["roofline"]["ms_phi_a"], "phi_t %.3f ms" % d["roofline"]["ms_phi_t"], flush=True)
    except Exception:
        print(cfg, "FAILED", p.stderr[-400:], flush=True)
