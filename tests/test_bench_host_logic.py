"""Host logic of bench.py's CPU arm (no GPU, no oracle): the thread count the reference arm gives the oracle.
Round 1's `--impl reference` arm was 3.8x slower than the cpu_baseline leg of the same box (cgroup CPU quota), and a
later version came out single-threaded on hosts WITHOUT a quota (OMP_PROC_BIND narrows the affinity mask of the calling
thread once libgomp is loaded): both are pinned here."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeBatch:
    def __init__(self, seconds_by_threads):
        self.t = seconds_by_threads
        self.calls = []

    def solve(self, threads):
        self.calls.append(threads)
        return self.t[threads]


def test_host_thread_count_is_read_once(monkeypatch):
    import bench

    h = bench.host_threads()
    assert h >= 1
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: {0}, raising=False)  # what libgomp's binding leaves behind
    assert bench.host_threads() == h
    cands = bench.thread_candidates()
    assert h in cands and max(1, h // 2) in cands


def test_thread_choice_uses_sustained_time_and_respects_the_quota(monkeypatch):
    import bench

    t = {16: 0.110, 32: 0.105, 64: 0.174, 128: 0.150}
    monkeypatch.setattr(bench, "cpu_quota", lambda: 16.0)
    best, tried = bench.best_thread_count(FakeBatch(t), [128, 64, 32, 16])
    assert best == 16 and set(tried) == set(t)  # 32 threads are 4.5 % faster: not enough to leave the quota's count
    best, _ = bench.best_thread_count(FakeBatch({**t, 32: 0.090}), [128, 64, 32, 16])
    assert best == 32  # 18 % faster: taken
    monkeypatch.setattr(bench, "cpu_quota", lambda: None)
    best, _ = bench.best_thread_count(FakeBatch(t), [128, 64, 32, 16])
    assert best == 32  # no quota: ties within 3 % go to fewer threads, 4.5 % is a win
    best, _ = bench.best_thread_count(FakeBatch({64: 0.100, 128: 0.099}), [128, 64])
    assert best == 64
