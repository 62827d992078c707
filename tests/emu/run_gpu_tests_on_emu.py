"""TEST INFRASTRUCTURE: run the bodies of the small `-m gpu` parity tests on the CPU emulator (slow; developer aid).
usage: PQP_B200_LIB=tests/emu/libpqp_emu.so python tests/emu/run_gpu_tests_on_emu.py [name-substring ...]"""
import importlib.util
import inspect
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
assert os.environ.get("PQP_B200_LIB", "").endswith("libpqp_emu.so")
from oracle import oracle as O  # noqa: E402
from proxsuite_b200 import proxqp  # noqa: E402

SKIP = ("full_size", "repeated_launches", "overflow_retry", "fused_feed_equals", "fused_feed_odd", "maros", "parity_with_oracle_on_seeded")
spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
t = importlib.util.module_from_spec(spec)
spec.loader.exec_module(t)
want = sys.argv[1:]
failed = 0
for name, fn in inspect.getmembers(t, inspect.isfunction):
    if not name.startswith("test_") or any(s in name for s in SKIP) or (want and not any(w in name for w in want)):
        continue
    params = inspect.signature(fn).parameters
    if "monkeypatch" in params or "kind" in params:
        continue
    kw = {}
    if "px" in params:
        kw["px"] = proxqp
    if "oracle" in params:
        kw["oracle"] = O
    t0 = time.time()
    try:
        fn(**kw)
        print("PASS %-60s %.1f s" % (name, time.time() - t0), flush=True)
    except Exception:
        failed += 1
        print("FAIL %-60s" % name, flush=True)
        traceback.print_exc()
sys.exit(1 if failed else 0)
