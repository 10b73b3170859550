"""Extracts per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and a few headline metrics of the
named kernels from an `ncu --set full` report and writes them as JSON -- the file bench.py reads `roofline.traffic` from.

  python tools/ncu_traffic.py gpurun_out/x.ncu-rep profiles/r02_traffic.json c3 "ncu command line" """
import csv, io, json, subprocess, sys

rep, out, cfg, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    def val(name, scale_unit=True):
        v = float(d[name].replace(",", ""))
        if scale_unit:
            v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u[name], 1.0)
        return v
    k = d["Kernel Name"]
    e = res.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "duration_us": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    e["duration_us"] += val("gpu__time_duration.sum", False) * {"us": 1, "ms": 1e3, "ns": 1e-3, "msecond": 1e3, "usecond": 1, "nsecond": 1e-3}.get(u["gpu__time_duration.sum"], 1)
    for m, key in (("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                   ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
                   ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
                   ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared_wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared_bank_conflicts"),
                   ("launch__registers_per_thread", "registers"), ("launch__block_size", "block"), ("launch__grid_size", "grid")):
        if m in d:
            e[key] = float(d[m].replace(",", ""))
for e in res.values():
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / e["launches"]
    e["duration_us_per_launch"] = e.pop("duration_us") / e["launches"]
try:
    allc = json.load(open(out))
except Exception:
    allc = {}
allc[cfg] = {"source": rep.split("/")[-1], "command": cmd, "kernels": res}
json.dump(allc, open(out, "w"), indent=1)
print(json.dumps(allc[cfg], indent=1))
