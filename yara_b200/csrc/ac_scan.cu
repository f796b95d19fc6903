// ac_scan.cu -- the Aho-Corasick atom walk of _yr_scanner_scan_mem_block
// (scanner.c:45-176) for sm_100a.
//
// Exactness of chunking (SURVEY.md 8(a) "Chunking exactness"): trie depth is at
// most YR_MAX_ATOM_LENGTH = 4, so the state after consuming byte p is a pure
// function of bytes p-3..p.  A worker that starts at ROOT three bytes before its
// range therefore reaches, from its first owned event on, exactly the states the
// sequential reference walk would be in.
//
// Two kernels:
//   ac_walk_kernel    every thread walks its own chunk through the transition
//                     table (general fallback: any ruleset, any atom lengths)
//   ac_filter_kernel  shared-memory atom prefilter over 16-byte vector loads;
//                     only the (rare) survivors are resolved exactly through the
//                     transition table (see the comment above the kernel)
#include <cstring>

#include "emit.cuh"

// ------------------------------------------------------------------ walk kernel

#define WALK_CHUNK   256
#define WALK_THREADS 128

__global__ void __launch_bounds__(WALK_THREADS)
ac_walk_kernel(DevRules R, ScanView V, CandOut out, int64_t lo, int64_t hi)
{
  const int64_t a = lo + ((int64_t) blockIdx.x * WALK_THREADS + threadIdx.x) * WALK_CHUNK;

  if (a >= hi)
    return;

  const int64_t b = (a + WALK_CHUNK < hi) ? a + WALK_CHUNK : hi;

  int64_t p = a - 3;
  if (p < 0) p = 0;
  if (p < V.g0)
  {
    // the shard was handed over without its 3-byte look-back margin
    atomicExch(&out.st->margin_fault, 1u);
    p = V.g0;
  }

  const uint8_t* __restrict__ d = V.d - V.g0; // d[p] is block byte p
  uint32_t state = 0;

#define CONSUME(byte)                                              \
  {                                                                \
    state = ac_step(R.T, state, (uint32_t) (byte));                \
    p++;                                                           \
    if (p > a && p > V.own_lo && p <= V.own_hi)                    \
    {                                                              \
      const uint32_t mm = __ldg(&R.M[state]);                      \
      if (mm != 0) emit_events(R, V, out, p, mm);                  \
    }                                                              \
  }

  // head: bytes until the address is 16-byte aligned
  while (p < b && (((uintptr_t) (d + p)) & 15) != 0)
    CONSUME(__ldg(d + p));

  // body: 16-byte vector loads
  while (p + 16 <= b)
  {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(d + p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
#pragma unroll
      for (int j = 0; j < 4; j++)
        CONSUME((w[k] >> (8 * j)) & 0xFF);
    }
  }

  while (p < b)
    CONSUME(__ldg(d + p));
#undef CONSUME
}

// The i == 0 event: scanner.c:98 runs once with state == root before any byte is
// consumed (zero-length atoms in the root's match list; also the only event of
// an empty block).
__global__ void ac_event0_kernel(DevRules R, ScanView V, CandOut out)
{
  const uint32_t m0 = __ldg(&R.M[0]);
  if (m0 != 0)
    emit_events(R, V, out, 0, m0);
}

void launch_ac_event0(const DevRules& R, const ScanView& V, const CandOut& out, cudaStream_t s)
{
  ac_event0_kernel<<<1, 1, 0, s>>>(R, V, out);
}

void launch_ac_scan_walk(
    const DevRules& R, const ScanView& V, const CandOut& out, int64_t lo, int64_t hi, cudaStream_t s)
{
  if (hi <= lo)
    return;

  const int64_t nthreads = (hi - lo + WALK_CHUNK - 1) / WALK_CHUNK;
  const unsigned blocks = (unsigned) ((nthreads + WALK_THREADS - 1) / WALK_THREADS);
  ac_walk_kernel<<<blocks, WALK_THREADS, 0, s>>>(R, V, out, lo, hi);
}

// ---------------------------------------------------------------- filter kernel
// (placeholder until the prefilter kernel lands: reports "not usable" so that the
// engine picks the walk kernel)
int build_filter_tables(const void* host_image, int device, FilterTables* out, void** d_storage)
{
  (void) host_image; (void) device;
  memset(out, 0, sizeof(*out));
  *d_storage = nullptr;
  return YRB_SUCCESS;
}

void launch_ac_scan_filter(
    const DevRules& R, const FilterTables& F, const ScanView& V, const CandOut& out, int64_t lo, int64_t hi,
    int num_sms, cudaStream_t s)
{
  (void) F; (void) num_sms;
  launch_ac_scan_walk(R, V, out, lo, hi, s);
}
