// TEST INFRASTRUCTURE ONLY — never part of the product library.
//
// Minimal lock-step SIMT emulation so that the kernel SOURCE in
// chemtools_b200/csrc/kernels.cuh can be compiled with g++ and its warp-level
// logic (shuffles, __syncwarp, shared memory carving, the device BDF stepper)
// exercised by the CPU-only test tier (`pytest -m "not gpu"`), where no GPU
// exists.  One std::thread per CUDA thread; __syncwarp/__shfl are barriers.
// The product (libchemtools_b200.so) has no CPU path and never includes this.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CT_SIMT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };

class EmuBarrier {
 public:
  explicit EmuBarrier(int n = 1) : n_(n) {}
  void reset(int n) { n_ = n; count_ = 0; gen_ = 0; }
  void wait() {
    unsigned g = gen_.load(std::memory_order_acquire);
    if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
      count_.store(0, std::memory_order_relaxed);
      gen_.fetch_add(1, std::memory_order_release);
    } else {
      while (gen_.load(std::memory_order_acquire) == g) std::this_thread::yield();
    }
  }
 private:
  int n_;
  std::atomic<int> count_{0};
  std::atomic<unsigned> gen_{0};
};

struct EmuWarp {
  EmuBarrier bar{32};
  alignas(64) unsigned long long xch[32];
};
struct EmuBlock {
  EmuBarrier bar;
  std::vector<EmuWarp> warps;
  unsigned char* smem = nullptr;
};

inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
inline thread_local EmuBlock* emu_blk = nullptr;
inline thread_local EmuWarp* emu_wp = nullptr;
inline thread_local int emu_lane = 0;

inline void __syncthreads() { emu_blk->bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_wp->bar.wait(); }

template <class T>
inline T emu_xchg(T v, int src) {
  static_assert(sizeof(T) <= 8, "emu shuffle: <= 8 bytes");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  emu_wp->xch[emu_lane] = raw;
  emu_wp->bar.wait();
  unsigned long long got = emu_wp->xch[src & 31];
  emu_wp->bar.wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_xchg(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_xchg(v, emu_lane ^ m); }
inline long long __double_as_longlong(double x) { long long r; std::memcpy(&r, &x, 8); return r; }
inline int __any_sync(unsigned, int pred) {
  int v = pred ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) v |= emu_xchg(v, emu_lane ^ o);
  return v;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { return emu_xchg(v, emu_lane - d >= 0 ? emu_lane - d : emu_lane); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu_xchg(v, emu_lane + d < 32 ? emu_lane + d : emu_lane); }

inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long val) { unsigned long long e = cmp; __atomic_compare_exchange_n(p, &e, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); return e; }
inline int atomicCAS(int* p, int cmp, int val) { int e = cmp; __atomic_compare_exchange_n(p, &e, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); return e; }
inline int atomicExch(int* p, int val) { return __atomic_exchange_n(p, val, __ATOMIC_ACQ_REL); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double exp10(double x) { return std::pow(10.0, x); }
inline double __ldg(const double* p) { return *p; }
using std::fabs; using std::fmax; using std::fmin; using std::log; using std::exp; using std::sqrt; using std::pow; using std::log10;

inline thread_local unsigned char* emu_smem_base = nullptr;  // what kernels.cuh's ct_smem resolves to

// Launch: blocks run one after another, threads of a block concurrently.
template <class Kernel, class... Args>
void emu_launch(Kernel k, unsigned grid, unsigned block, size_t smem_bytes, Args... args) {
  for (unsigned b = 0; b < grid; ++b) {
    EmuBlock blk;
    blk.bar.reset((int)block);
    blk.warps = std::vector<EmuWarp>((block + 31) / 32);
    for (unsigned w = 0; w < blk.warps.size(); ++w) blk.warps[w].bar.reset((int)std::min(32u, block - 32 * w));
    std::vector<unsigned char> smem(smem_bytes + 64, 0xFF);  // NaN pattern: real shared memory is NOT zero-initialised
    blk.smem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; ++t) {
      th.emplace_back([&, t, b]() {
        threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
        emu_blk = &blk; emu_wp = &blk.warps[t / 32]; emu_lane = (int)(t % 32); emu_smem_base = blk.smem;
        k(args...);
      });
    }
    for (auto& x : th) x.join();
  }
}
