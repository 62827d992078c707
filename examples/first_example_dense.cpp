// C++ use of the drop-in: the reference's examples/cpp/first_example_dense.cpp, line for line,
// on proxsuite_b200.hpp (row-major arrays instead of Eigen matrices), followed by the batched
// path (BatchQP + solve_in_parallel) on random QPs of the reference's generator.
//   g++ -std=c++17 -Iinclude examples/first_example_dense.cpp -Lproxsuite_b200 -lpqp_b200 -Wl,-rpath,$PWD/proxsuite_b200
// Exit code 0: solved; 3: no usable CUDA device (there is no CPU fallback).
#include <proxsuite_b200.hpp>

#include <cmath>
#include <cstdio>
#include <vector>

using namespace proxsuite_b200::proxqp;

int
main()
{
  try {
    std::printf("Solve a simple example with inequality constraints using dense ProxQP on the GPU\n");
    const double eps_abs = 1e-9;
    const isize dim = 3, n_eq = 0, n_in = 3;
    const double H[9] = { 13.0, 12.0, -2.0, 12.0, 17.0, 6.0, -2.0, 6.0, 12.0 };
    const double g[3] = { -22.0, -14.5, 13.0 };
    const double C[9] = { 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0 };
    const double l[3] = { -1.0, -1.0, -1.0 };
    const double u[3] = { 1.0, 1.0, 1.0 };

    dense::QP qp(dim, n_eq, n_in);
    qp.settings.eps_abs = eps_abs;
    qp.settings.initial_guess = int(InitialGuessStatus::NO_INITIAL_GUESS);
    qp.init(H, g, nullptr, nullptr, C, l, u);
    qp.solve();
    std::printf("primal residual: %.3e\ndual residual: %.3e\ntotal number of iteration: %lld\n", qp.results.info.pri_res, qp.results.info.dua_res, (long long)qp.results.info.iter);
    std::printf("x = [%.6f %.6f %.6f]  (expected [1 0.5 -1], test/src/cvxpy.py:24-46)\n", qp.results.x[0], qp.results.x[1], qp.results.x[2]);
    bool ok = qp.results.info.status == PQP_SOLVED && std::fabs(qp.results.x[0] - 1.0) < 1e-6 && std::fabs(qp.results.x[1] - 0.5) < 1e-6 && std::fabs(qp.results.x[2] + 1.0) < 1e-6;

    // batched path: 64 random QPs (n = 20, n_eq = 5, n_in = 10), one kernel launch
    const isize B = 64, n = 20, ne = 5, ni = 10;
    dense::BatchQP batch(B);
    std::vector<double> Hm(n * n), gm(n), Am(ne * n), bm(ne), Cm(ni * n), um(ni), lm(ni), ub(n), lb(n);
    for (isize i = 0; i < B; ++i) {
      pqp_random_qp(0 /*dense_strongly_convex_qp*/, std::uint64_t(i), n, ne, ni, 0.15, 1e-2, Hm.data(), gm.data(), Am.data(), bm.data(), Cm.data(), um.data(), lm.data(), ub.data(), lb.data());
      dense::QP& q = batch.init_qp_in_place(n, ne, ni);
      q.settings.eps_abs = eps_abs;
      q.settings.eps_rel = 0;
      q.init(Hm.data(), gm.data(), Am.data(), bm.data(), Cm.data(), lm.data(), um.data());
    }
    dense::solve_in_parallel(batch);
    isize solved = 0;
    double worst = 0;
    for (isize i = 0; i < B; ++i) {
      solved += batch[i].results.info.status == PQP_SOLVED;
      worst = std::fmax(worst, std::fmax(batch[i].results.info.pri_res, batch[i].results.info.dua_res));
    }
    std::printf("BatchQP: %lld / %lld solved, worst residual %.3e\n", (long long)solved, (long long)B, worst);
    ok = ok && solved == B && worst <= eps_abs;
    return ok ? 0 : 1;
  } catch (const std::invalid_argument& e) {
    std::printf("invalid argument: %s\n", e.what());
    return 2;
  } catch (const std::runtime_error& e) {
    std::printf("runtime error (no CPU fallback): %s\n", e.what());
    return 3;
  }
}
