#!/usr/bin/env python3
"""Attribute an ncu source-page export to the device functions of the solve kernel.

usage: python tools/ncu_funcs.py REPORT.ncu-rep [LIB.so] [--kernel fastk]
Joins `ncu -i REPORT --page source --csv --print-source sass` (per-SASS-instruction executed
counts and stall samples) with the function boundaries nvdisasm reports for LIB (which must be
the binary the capture ran). Prints per function: static size, warp instructions executed,
share, stall samples, share, and the top stall reasons.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep = os.path.abspath(args[0])
lib = os.path.abspath(args[1]) if len(args) > 1 else os.path.join(ROOT, "proxsuite_b200", "libpqp_b200.so")
ns = "fastk"
if "--kernel" in sys.argv:
    ns = sys.argv[sys.argv.index("--kernel") + 1]
kern = "_ZN%d%s16pqp_solve_kernelE12PqpSolveArgs" % (len(ns), ns)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, "pqp_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout
bounds, cur, active = [], None, False  # (offset, name)
for line in txt.splitlines():
    if line.startswith(kern + ":"):
        cur, active = "solve_one", True
        continue
    if line.startswith("//-----") and active and kern not in line:
        active = False
    if not active:
        continue
    m = re.match(r"^\$" + re.escape(kern) + r"\$(\S+):", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and cur is not None:
        bounds.append((int(m.group(1), 16), cur))
off2fn = dict(bounds)

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# the export holds one table per profiled kernel; keep the one for our kernel
blocks, curb = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        curb = [line]
        blocks.append(curb)
    elif curb is not None:
        curb.append(line)
sel = [b for b in blocks if (ns + "::pqp_solve_kernel") in b[0]]
if not sel:
    sys.exit("kernel %s not in report" % ns)
rows = list(csv.reader(io.StringIO("\n".join(sel[0][1:]))))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
base = None
st = collections.defaultdict(lambda: collections.Counter())
stall_cols = [h for h in hdr if h.startswith("stall_") and "(" not in h]
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[ci["Address"]], 16)
    if base is None:
        base = addr
    fn = off2fn.get(addr - base, "?")
    s = st[fn]
    s["static"] += 1
    s["inst"] += int(r[ci["Instructions Executed"]] or 0)
    s["thread_inst"] += int(r[ci["Thread Instructions Executed"]] or 0)
    s["samples"] += int(r[ci["# Samples"]] or 0)
    for h in stall_cols:
        try:
            s[h] += int(r[ci[h]] or 0)
        except ValueError:
            pass
tot_i = sum(s["inst"] for s in st.values())
tot_s = sum(s["samples"] for s in st.values())
print("%-28s %7s %12s %6s %6s %8s %6s  top stalls" % ("function", "static", "warp_inst", "share", "lanes", "samples", "share"))
for fn, s in sorted(st.items(), key=lambda kv: -kv[1]["samples"]):
    tops = sorted(((h, s[h]) for h in stall_cols), key=lambda kv: -kv[1])[:4]
    print("%-28s %7d %12d %5.1f%% %6.1f %8d %5.1f%%  %s" % (fn[:28], s["static"], s["inst"], 100.0 * s["inst"] / max(tot_i, 1), s["thread_inst"] / max(s["inst"], 1), s["samples"], 100.0 * s["samples"] / max(tot_s, 1),
                                                   " ".join("%s:%d" % (h.replace("stall_", ""), v) for h, v in tops if v)))
print("total warp_inst %d samples %d" % (tot_i, tot_s))
