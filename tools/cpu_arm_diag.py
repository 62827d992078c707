"""Why did round 1's `bench.py --impl reference` run 3.8x slower than the cpu_baseline leg on the same box?
Runs the same 1024-QP oracle batch in fresh subprocesses under different conditions and prints QP/s for each:
  plain            : no torch import, environment as inherited
  torch_first      : `import torch` before the oracle library is loaded (what the GPU arm's leg does)
  pinned           : OMP_PROC_BIND=true OMP_PLACES=threads OMP_WAIT_POLICY=active (what bench.py now sets)
  omp1_env         : OMP_NUM_THREADS=1 exported by the parent (torchrun does that), thread count passed explicitly
Also prints the host facts that matter (affinity, cgroup quota, NUMA nodes, load)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, time, json
sys.path.insert(0, %r)
if os.environ.get("DIAG_TORCH") == "1":
    import torch
import numpy as np
import bench
H = bench.host_threads()
if os.environ.get("DIAG_PIN") == "1":
    bench.pin_openmp_env(H)
b, flags = bench.make_oracle_batch(int(os.environ.get("DIAG_SAMPLE", "1024")))
out = {}
for T in (H, max(1, H // 2)):
    b.solve(T)
    ts = [b.solve(T) for _ in range(5)]
    out[str(T)] = [round(len(b) / t) for t in ts]
print(json.dumps(dict(mode=os.environ.get("DIAG_MODE"), host_threads=H, qps=out, flags=flags, omp_env={k: v for k, v in os.environ.items() if k.startswith(("OMP", "GOMP", "KMP", "MKL"))})))
''' % ROOT


def facts():
    f = {"affinity": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(), "loadavg": os.getloadavg()}
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/devices/system/node/online"):
        try:
            f[p] = open(p).read().strip()
        except OSError:
            pass
    try:
        f["lscpu"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines() if any(k in l for k in ("Model name", "Socket", "Thread(s)", "NUMA node", "Core(s)"))]
    except Exception:
        pass
    return f


def run(mode, **env):
    e = dict(os.environ, DIAG_MODE=mode, **env)
    p = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=600)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    print(line[-1] if line else ("FAILED " + mode + ": " + p.stderr[-400:]), flush=True)


if __name__ == "__main__":
    print(json.dumps(facts()), flush=True)
    run("plain")
    run("torch_first", DIAG_TORCH="1")
    run("pinned", DIAG_PIN="1")
    run("omp1_env", OMP_NUM_THREADS="1")
    run("plain_again")
