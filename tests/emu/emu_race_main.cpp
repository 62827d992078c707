// TEST INFRASTRUCTURE: race-detection driver of the CPU emulator. Built with -fsanitize=thread together with the
// kernel sources and emu_runtime.cpp (make -C tests/emu race): every emulated CUDA thread is a TSan fiber and only
// barriers order them, so TSan reports unsynchronised conflicting accesses inside a CTA. Solves a few small QPs of
// each kernel family through the C-ABI (forward + backward).
#include "../../include/pqp.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static int g_initial_guess = PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS;
static int
run(const char* name, int kind, int B, int n, int ne, int ni, int box, int hessian, double sparsity, const char* layout, bool backward)
{
  if (layout)
    setenv("PQP_LAYOUT", layout, 1);
  else
    unsetenv("PQP_LAYOUT");
  const int rows = kind == 2 ? 2 * ni : ni;
  std::vector<double> H((size_t)B * n * n), g((size_t)B * n), A((size_t)B * ne * n), b((size_t)B * ne), C((size_t)B * rows * n), u((size_t)B * rows), l((size_t)B * rows), ub((size_t)B * n), lb((size_t)B * n);
  for (int i = 0; i < B; ++i) {
    if (pqp_random_qp(kind, (uint64_t)i, n, ne, ni, sparsity, 1e-2, &H[(size_t)i * n * n], &g[(size_t)i * n], A.data() + (size_t)i * ne * n, b.data() + (size_t)i * ne, C.data() + (size_t)i * rows * n, u.data() + (size_t)i * rows,
                      l.data() + (size_t)i * rows, &ub[(size_t)i * n], &lb[(size_t)i * n]) != 0) {
      std::printf("%s: generator failed: %s\n", name, pqp_last_error());
      return 1;
    }
  }
  pqp_batch* bt = pqp_batch_create(B, n, ne, rows, box, hessian, PQP_BACKEND_PRIMAL_DUAL_LDLT, 0);
  if (!bt) {
    std::printf("%s: create failed: %s\n", name, pqp_last_error());
    return 1;
  }
  pqp_settings s;
  pqp_batch_settings_get(bt, 0, &s);
  s.eps_abs = 1e-9;
  s.eps_rel = 0;
  s.initial_guess = g_initial_guess;
  pqp_batch_settings_set(bt, -1, &s);
  int rc = pqp_batch_init(bt, 0, B, H.data(), g.data(), A.data(), b.data(), C.data(), l.data(), u.data(), box ? lb.data() : nullptr, box ? ub.data() : nullptr, 1, nullptr, nullptr, nullptr, nullptr);
  if (rc == 0) rc = pqp_batch_solve(bt);
  std::vector<pqp_info> info((size_t)B);
  std::vector<double> x((size_t)B * n);
  if (rc == 0) rc = pqp_batch_results(bt, 0, B, x.data(), nullptr, nullptr, nullptr, nullptr, info.data());
  int solved = 0;
  for (int i = 0; i < B; ++i) solved += info[(size_t)i].status == PQP_SOLVED;
  if (rc == 0 && backward && !box) {
    std::vector<double> loss((size_t)B * (n + ne + rows), 0.0), dg((size_t)B * n);
    for (int i = 0; i < B; ++i) loss[(size_t)i * (n + ne + rows)] = 1.0;
    rc = pqp_batch_backward(bt, 0, B, loss.data(), 1e-9, 1e-7, 1e-7, nullptr, dg.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  std::printf("%s: rc %d, %d/%d solved%s\n", name, rc, solved, B, rc ? pqp_last_error() : "");
  pqp_batch_destroy(bt);
  return rc != 0 || solved != B;
}

int
main(int argc, char** argv)
{
  std::string which = argc > 1 ? argv[1] : "all";
  int bad = 0;
  if (which == "all" || which == "tile_cold") {
    g_initial_guess = PQP_NO_INITIAL_GUESS; // the first active-set change forms S^-1 for equalities + active rows at once
    bad += run("tile_cold", 0, 2, 12, 4, 8, 0, PQP_HESSIAN_DENSE, 0.15, nullptr, false);
    g_initial_guess = PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS;
  }
  if (which == "all" || which == "tile") bad += run("tile", 0, 2, 12, 4, 8, 0, PQP_HESSIAN_DENSE, 0.15, nullptr, false);
  if (which == "all" || which == "generic") bad += run("generic", 0, 2, 12, 4, 8, 0, PQP_HESSIAN_DENSE, 0.15, "generic", true);
  if (which == "all" || which == "compact") bad += run("compact", 0, 2, 12, 4, 8, 0, PQP_HESSIAN_DENSE, 0.15, "compact", false);
  if (which == "all" || which == "box") bad += run("box", 4, 2, 8, 3, 5, 1, PQP_HESSIAN_DENSE, 0.5, nullptr, false);
  if (which == "all" || which == "diag") bad += run("diag", 5, 2, 9, 3, 4, 1, PQP_HESSIAN_DIAGONAL, 0.5, nullptr, false);
  if (which == "g40") bad += run("g40", 0, 1, 40, 20, 40, 0, PQP_HESSIAN_DENSE, 0.15, "generic", false);
  if (which == "t40") bad += run("t40", 0, 1, 40, 20, 40, 0, PQP_HESSIAN_DENSE, 0.15, nullptr, false);
  return bad ? 1 : 0;
}
