"""SASS evidence for profiles/: mnemonic counts of the built library and of the run-time generated (NVRTC) kernels for
the C3 terms table, plus an excerpt of each hot loop.   python tools/sass_report.py > profiles/r02_sass.md   (no GPU)"""
import re, subprocess, sys, tempfile
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np
import outerbase_b200 as ob
import bench
from outerbase_b200.binding import Library

MNEM = ["UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "DFMA", "DMUL", "DADD", "DMMA", "LDS", "STS", "LDL", "STL", "BRX", "NANOSLEEP", "ATOMG", "RED", "LDTM", "STTM", "UTCHMMA"]


def functions(sass):
    out = {}
    for fn in re.split(r"\n\s*Function : ", sass)[1:]:
        name = fn.split("\n")[0].strip()
        ins = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l.strip()) for l in fn.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
        out[name] = ins
    return out


def table(fns):
    print("| function | instr | max R | " + " | ".join(MNEM) + " |")
    print("|---|---|---|" + "---|" * len(MNEM))
    for name, ins in fns.items():
        txt = "\n".join(ins)
        regs = [int(m) for m in re.findall(r"\bR(\d+)\b", txt)] or [0]
        print(f"| `{name[:60]}` | {len(ins)} | {max(regs)} | " + " | ".join(str(len(re.findall(r"\b" + m, txt))) for m in MNEM) + " |")


def excerpt(ins, pattern, before, after, title):
    idx = [i for i, l in enumerate(ins) if re.search(pattern, l)]
    if not idx:
        return
    i = idx[len(idx) // 2]
    print(f"\n{title}\n```")
    print("\n".join(ins[max(0, i - before):i + after]))
    print("```")


print("# SASS of the shipped kernels (round 2)\n")
so = ob.LIBPATH
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
print(f"## `{so.relative_to(REPO)}` (nvcc -gencode arch=compute_100a,code=sm_100a)\n")
print("arch lines:", sorted(set(re.findall(r"arch = (sm_\w+)", sass))), "\n")
fns = functions(sass)
keep = {k: v for k, v in fns.items() if any(t in k for t in ("phi_a_kernel", "phi_t_kernel", "basis_build_mma", "p2p_allreduce", "cg_stage", "phi_t_reduce", "fp64_peak"))}
table(keep)
for k, v in keep.items():
    if "p2p_allreduce_kernel<true>" in k or "ILb1" in k:
        excerpt(v, r"ST\.E\.128|STG\.E\.128|ST\.E\.STRONG|STG", 6, 10, f"`{k[:50]}`: tagged 16-byte peer stores")
        break
for k, v in keep.items():
    if "basis_build_mma_kernelILi5" in k or "basis_build_mma_kernel<5>" in k:
        excerpt(v, r"DMMA", 4, 12, "`basis_build_mma_kernel<5>`: FP64 tensor-core contraction")
        break

O = Library(REPO / "oracle" / "_build" / "libob_oracle.so", "orc_")
om, terms = bench.setup_model(O)
lib = ob.load_symbols_only()
d = Path(tempfile.mkdtemp())
import os
os.environ["OB_SPEC_OPTS"] = "1,2,4,80,8,1,4,16,56,4,0,1,8,16,2,3,10,96,232,40,128,1"
src_tmap = lib.spec_source(terms)[0]
del os.environ["OB_SPEC_OPTS"]
for tag, src in (("main", lib.spec_source(terms)[0]), ("dot", lib.spec_source_dot(terms)[0]), ("mat", lib.spec_source_mat(terms)[0]),
                 ("tmat", lib.spec_source_tmat(terms)[0]), ("main, option tmap = 1", src_tmap)):
    stem = re.sub(r"\W+", "_", tag)
    (d / f"{stem}.cu").write_text(src)
    subprocess.run(["nvcc", "-arch=sm_100a", "-cubin", "-lineinfo", "-std=c++17", "-o", str(d / f"{stem}.cubin"), str(d / f"{stem}.cu")], check=True, capture_output=True)
    s2 = subprocess.run(["cuobjdump", "-sass", str(d / f"{stem}.cubin")], capture_output=True, text=True).stdout
    f2 = functions(s2)
    print(f"\n## run-time generated module `{tag}` for the C3 terms table (d=10, K=2000; same source NVRTC compiles on the GPU box, here through nvcc -arch=sm_100a)\n")
    table(f2)
    for k, v in f2.items():
        if "tmap" in tag and k != "phi_a_spec":
            continue
        if k == "phi_t_spec":
            excerpt(v, r"USETMAXREG", 3, 4, "`phi_t_spec`: warp-specialised register split (setmaxnreg)")
            excerpt(v, r"UBLKCP", 8, 6, "`phi_t_spec` producer: one bulk copy (TMA engine) per staged column")
            dl = [i for i, l in enumerate(v) if "DFMA" in l]
            i = dl[len(dl) // 2]
            print("\n`phi_t_spec`: one stream's pass loop (register accumulators, one per term)\n```")
            print("\n".join(v[i - 20:i + 25]))
            print("```")
        if k == "phi_a_spec" and "tmap" in tag:
            excerpt(v, r"UTMALDG", 10, 6, "`phi_a_spec`, option tmap: one 2-D tensor copy per run of adjacent basis columns")
            continue
        if k == "phi_am_spec":
            excerpt(v, r"DMMA", 14, 10, "`phi_am_spec`: fragment loads and the FP64 tensor-core contraction of one block")
        if k == "phi_tm_spec":
            excerpt(v, r"DMMA", 14, 10, "`phi_tm_spec`: contraction over the rows of a pass")
        if k == "phi_t_spec" and "tmap" in tag:
            continue
        if k == "phi_a_spec":
            dl = [i for i, l in enumerate(v) if "DFMA" in l]
            i = dl[len(dl) // 2]
            print("\n`phi_a_spec`: Horner body (coefficients as broadcast LDS.128 pairs, register-cached columns)\n```")
            print("\n".join(v[i - 15:i + 20]))
            print("```")
        if k == "phi_d_spec":
            dl = [i for i, l in enumerate(v) if "DFMA" in l]
            i = dl[len(dl) // 3]
            print("\n`phi_d_spec`: reverse-mode walk (3 FMAs per inner edge)\n```")
            print("\n".join(v[i - 10:i + 15]))
            print("```")
