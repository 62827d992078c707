"""The CUDA kernel SOURCES (proxsuite_b200/csrc/*.cu, *.inl, host state machine included) compiled with g++ against
the functional emulator of tests/emu (cooperative fibers for the threads of a CTA, warp collectives, a fake CUDA
runtime over host memory) and run on the CPU against the oracle: forward parity on every kernel variant (tile /
general / generic layout, fused and stand-alone set-up, box, diagonal Hessian, degenerate, fewer constraint rows than
variables) and the QPLayer backward pass through the C-ABI. This is TEST INFRASTRUCTURE for kernel logic (barrier
pairing, index arithmetic, host state machine); it is never loaded by the package and is no product path — timing,
inter-warp races and hardware behaviour are covered by the `-m gpu` tests only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.check_call(["make", "-C", EMU], stdout=subprocess.DEVNULL)
    lib = os.path.join(EMU, "libpqp_emu.so")
    assert os.path.exists(lib)
    return lib


def run_cases(lib, names, order=None):
    env = dict(os.environ, PQP_B200_LIB=lib)
    env.pop("EMU_ORDER", None)
    if order:
        env["EMU_ORDER"] = order
    env.pop("PQP_LAYOUT", None)
    env.pop("PQP_E2E", None)
    p = subprocess.run([sys.executable, os.path.join(EMU, "run_emu_case.py")] + names, env=env, capture_output=True, text=True, timeout=900)
    recs = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    assert len(recs) == len(names) and all(r["ok"] for r in recs), recs
    return recs


def test_forward_parity_of_every_kernel_variant_on_the_emulator(emu_lib):
    run_cases(emu_lib, ["tile_small", "tile_plain_setup", "tile_eq_guess", "tile_box", "general_odd", "general_diag",
                        "generic_layout", "no_inequalities", "degenerate", "not_strongly_convex", "closest_feasible"])


def test_general_layout_with_fewer_constraint_rows_than_variables(emu_lib):
    """Regression (found with the emulator): P is swept inside the S^-1 region of the general layout, which was sized
    for the dual block only; with n_eq + n_in < n the sweep overran it and corrupted P^-1."""
    run_cases(emu_lib, ["few_rows_generic"])


def test_qplayer_backward_on_the_emulator(emu_lib):
    recs = run_cases(emu_lib, ["backward_eq", "backward_mixed", "backward_dy", "backward_api"])
    assert recs[0]["fd_diff"] < 1e-5 and recs[1]["fd_diff"] < 1e-5  # test/src/dense_backward.cpp acceptance


def test_torch_qp_layer_gradients_on_the_emulator(emu_lib):
    """proxsuite_b200.torch.QPFunction (mirror of proxsuite.torch.qplayer.QPFunction): autograd gradients w.r.t.
    H, g, A, b, C, u against central finite differences through the layer's forward pass."""
    run_cases(emu_lib, ["qplayer", "qplayer_device_api", "qplayer_infeas"])


def test_big_variant_and_its_kkt_fallback_on_the_emulator(emu_lib):
    """The BIG variant of the tile body (packed storage in the workspace, block rank-4 insertion / deletion, Hessian types,
    box rows) and its whole-KKT inverse fallback (PQP_FORCE_KKT=1): oracle parity incl. iteration counters."""
    run_cases(emu_lib, ["big_variant", "big_kkt"])


def test_sharded_batch_of_the_c_abi_on_the_emulator(emu_lib):
    """pqp_sharded_* (include/pqp.h): uneven slices over a device list, bit-identical to one batch; update + re-solve."""
    run_cases(emu_lib, ["sharded"])


@pytest.mark.parametrize("order", ["reverse", "stride"])
def test_results_do_not_depend_on_the_thread_order_between_barriers(emu_lib, order):
    """Order fuzzing: the emulator runs the threads of a CTA in a different order between barriers (EMU_ORDER). A
    kernel whose result changes with that order has a data race or relies on warp lockstep without saying so. Found
    this way: the pivot-block rows of the general kernel's sweep inversion re-read panel entries that a sibling
    thread of the same warp was overwriting (harmless under lockstep, wrong under any other order)."""
    run_cases(emu_lib, ["tile_small", "tile_box", "general_odd", "general_diag", "generic_layout", "few_rows_generic",
                        "backward_mixed"], order=order)
