// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Functional emulator of the slice of the CUDA execution model the kernels of
// proxsuite_b200/csrc use, so that the SAME kernel sources (pqp_kernels.cu +
// pqp_fast_body.inl + pqp_solver_body.inl + pqp_capi.cu) can be compiled with
// g++ and run on a CPU for logic regression tests (`-m "not gpu"`), e.g. barrier
// pairing, index arithmetic, memory bounds (ASan), the host state machine.
// It is never built into, loaded by or linked to the package: the product has no
// CPU path (tests/emu/README.md). What it does NOT model: timing, races between
// warps (fibers are cooperative and switch only at barriers / warp collectives),
// caches, the FMA order inside the tensor-core MMA.
//
// A CTA = `block` cooperative fibers (ucontext) on one OS thread, run to the next
// barrier in round-robin order; CTAs of a grid run one after the other.
#pragma once
#define PQP_CPU_EMU 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <ucontext.h>
#include <vector>

// ---- CUDA keywords ------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __builtin_assume(x) ((void)0)

struct double2
{
  double x, y;
};
static inline double2
make_double2(double x, double y)
{
  return double2{ x, y };
}
using std::max;
using std::min;

namespace emu {
struct Idx
{
  unsigned x, y, z;
};
extern Idx thread_idx, block_idx, block_dim, grid_dim;
extern double* const dyn_smem; // dynamic shared memory of the running CTA (16-byte aligned)
unsigned long long globaltimer();
void yield();
void block_barrier();
void warp_barrier();
struct WarpBuf
{
  double d[32], a[32], b[32];
  long long i[32];
};
WarpBuf& warp_buf();
void run_grid(int grid, int block, const std::function<void()>& body);

template<class Arg, class Arg2>
inline void
launch(void (*kern)(Arg), int grid, int block, size_t smem_bytes, const Arg2& arg)
{
  if (smem_bytes > 232448) {
    std::fprintf(stderr, "emu: %zu bytes of dynamic shared memory requested\n", smem_bytes);
    std::abort();
  }
  run_grid(grid, block, [&]() { kern(arg); });
}

template<class A1, class A2, class B1, class B2>
inline void
launch2(void (*kern)(A1, A2), int grid, int block, size_t smem_bytes, const B1& a1, const B2& a2)
{
  if (smem_bytes > 232448) std::abort();
  run_grid(grid, block, [&]() { kern(a1, a2); });
}

inline int
lane_id()
{
  return (int)(thread_idx.x & 31u);
}
// mma.sync.aligned.m8n8k4.row.col.f64: A[g][k] in lane 4g + k, B[k][c] in lane 4c + k, C[g][2t + i] in lane 4g + t
inline void
dmma_8x8x4(double& c0, double& c1, double a, double b)
{
  WarpBuf& w = warp_buf();
  const int l = lane_id(), g = l >> 2, t = l & 3;
  w.a[l] = a;
  w.b[l] = b;
  warp_barrier();
  for (int k = 0; k < 4; ++k) {
    c0 = std::fma(w.a[4 * g + k], w.b[4 * (2 * t) + k], c0);
    c1 = std::fma(w.a[4 * g + k], w.b[4 * (2 * t + 1) + k], c1);
  }
  warp_barrier();
}
} // namespace emu

#define threadIdx (emu::thread_idx)
#define blockIdx (emu::block_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)

// ---- device intrinsics ----------------------------------------------------------
static inline void
__syncthreads()
{
  emu::block_barrier();
}
int __syncthreads_or(int p);
static inline void
__syncwarp(unsigned = 0xffffffffu)
{
  emu::warp_barrier();
}
template<class T>
static inline T
__shfl_xor_sync(unsigned, T v, int o)
{
  emu::WarpBuf& w = emu::warp_buf();
  const int l = emu::lane_id();
  double* slot = w.d;
  static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
  std::memcpy(&slot[l], &v, sizeof(T));
  emu::warp_barrier();
  T r;
  std::memcpy(&r, &slot[l ^ o], sizeof(T));
  emu::warp_barrier();
  return r;
}
static inline unsigned
__reduce_max_sync(unsigned, unsigned v)
{
  emu::WarpBuf& w = emu::warp_buf();
  const int l = emu::lane_id();
  std::memcpy(&w.d[l], &v, sizeof(v));
  emu::warp_barrier();
  unsigned m = 0;
  for (int k = 0; k < 32; ++k) {
    unsigned x;
    std::memcpy(&x, &w.d[k], sizeof(x));
    m = x > m ? x : m;
  }
  emu::warp_barrier();
  return m;
}
static inline int
__double2hiint(double v)
{
  unsigned long long b;
  std::memcpy(&b, &v, 8);
  return (int)(unsigned)(b >> 32);
}
static inline int
__double2loint(double v)
{
  unsigned long long b;
  std::memcpy(&b, &v, 8);
  return (int)(unsigned)(b & 0xffffffffull);
}
static inline double
__hiloint2double(int hi, int lo)
{
  const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned long long)(unsigned)lo;
  double v;
  std::memcpy(&v, &b, 8);
  return v;
}
template<class T>
static inline T
__shfl_sync(unsigned, T v, int src)
{
  emu::WarpBuf& w = emu::warp_buf();
  const int l = emu::lane_id();
  static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
  std::memcpy(&w.d[l], &v, sizeof(T));
  emu::warp_barrier();
  T r;
  std::memcpy(&r, &w.d[src & 31], sizeof(T));
  emu::warp_barrier();
  return r;
}
template<class T>
static inline T
__shfl_up_sync(unsigned, T v, int delta)
{
  emu::WarpBuf& w = emu::warp_buf();
  const int l = emu::lane_id();
  std::memcpy(&w.d[l], &v, sizeof(T));
  emu::warp_barrier();
  T r = v;
  if (l >= delta) std::memcpy(&r, &w.d[l - delta], sizeof(T));
  emu::warp_barrier();
  return r;
}
static inline unsigned
__ballot_sync(unsigned, int p)
{
  emu::WarpBuf& w = emu::warp_buf();
  const int l = emu::lane_id();
  w.i[l] = p ? 1 : 0;
  emu::warp_barrier();
  unsigned m = 0;
  for (int k = 0; k < 32; ++k) m |= (w.i[k] ? 1u : 0u) << k;
  emu::warp_barrier();
  return m;
}
static inline int
__popc(unsigned v)
{
  return __builtin_popcount(v);
}
template<class T>
static inline T
atomicAdd(T* p, T v)
{
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline long long
clock64()
{
  return (long long)emu::globaltimer();
}
template<class T>
static inline T
__ldcg(const T* p)
{
  return *p;
}
template<class T>
static inline void
__stcg(T* p, T v)
{
  *p = v;
}
static inline void
__nanosleep(unsigned)
{
  emu::yield();
}
static inline void
__threadfence()
{
}
static inline bool
__isShared(const void*)
{
  return true;
}
