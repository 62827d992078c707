#!/usr/bin/env python3
"""Per-device-function SASS statistics of the solve kernel (static instruction mix).

usage: python tools/sass_funcs.py [namespace] [function-substring]
Disassembles proxsuite_b200/libpqp_b200.so with cuobjdump/nvdisasm and prints, for each
non-inlined device function of the chosen kernel, the instruction count and opcode mix.
With a function substring it dumps that function's SASS.
"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ns = sys.argv[1] if len(sys.argv) > 1 else "fastk"
want = sys.argv[2] if len(sys.argv) > 2 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "proxsuite_b200", "libpqp_b200.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, "pqp_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout
kern = "_ZN%d%s16pqp_solve_kernelE12PqpSolveArgs" % (len(ns), ns)
cur, funcs, active = None, collections.OrderedDict(), False
for line in txt.splitlines():
    if line.startswith(kern + ":"):
        cur, active = "solve_one(kernel body)", True
        funcs[cur] = []
        continue
    if line.startswith("//-----") and active and kern not in line:
        active = False
    if not active:
        continue
    m = re.match(r"^\$" + re.escape(kern) + r"\$(\S+):", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
    if m:
        funcs[cur].append(m.group(1).strip())

def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
tot = 0
for name, ins in funcs.items():
    dn = demangle(name)
    if want and want not in dn:
        continue
    ops = collections.Counter()
    for i in ins:
        t = i.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    tot += len(ins)
    print("%-42s %6d  %s" % (dn[:42], len(ins), " ".join("%s:%d" % kv for kv in ops.most_common(12))))
    if want:
        print("\n".join(ins))
print("total", tot)
