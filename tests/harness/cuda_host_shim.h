/*
 * cuda_host_shim.h -- just enough of the CUDA device vocabulary to compile
 * yara_b200/csrc/verify.cu and emit.cuh with a plain C++ compiler and run their
 * kernels as ordinary functions, one "thread" (0,0,0) of a 1x1 grid.
 *
 * TEST INFRASTRUCTURE ONLY (tests/harness/vm_host.cpp, tests/test_vm_host_cpu.py): it lets the
 * CPU suite replay the product's verifier code -- literal compare family, the fast hex VM, the
 * fiber VM with its tiers -- against the oracle without a GPU.  Nothing in the product includes it.
 */
#pragma once
#include <stdint.h>
#include <string.h>

#include <cuda_runtime.h>

#ifndef __CUDACC__
#undef __launch_bounds__
#define __launch_bounds__(...)
struct yrb_shim_dim { unsigned x, y, z; };
/* the driver sets these before it calls a kernel as a function; per host thread, so that several
   threads can replay scans at the same time (yr_rules_scan_mem from N threads through the stand-in) */
static thread_local yrb_shim_dim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

template <typename T> static inline T __ldg(const T* p) { return *p; }

template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }

/* a "warp" of ONE lane (resolve_kernels.cuh is compiled with YRB_WARP = 1): collectives are identities */
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
static inline int __any_sync(unsigned, int pred) { return pred; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned) { return v; }
static inline void __syncwarp() {}
static inline void __threadfence_block() {}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int) x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
#endif
