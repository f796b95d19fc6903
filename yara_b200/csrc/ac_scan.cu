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
#include <cstdlib>
#include <cstring>

#include "emit.cuh"

// ------------------------------------------------------------------ walk kernel

#define WALK_CHUNK       256   // bytes per thread for large ranges (3 bytes of look-back re-walk each)
#define WALK_CHUNK_SMALL 32    // small ranges are latency bound: shorter dependent chains, more threads
#define WALK_CHUNK_TINY  8     // heads/tails of filter slices (< 1 MiB): ~11 dependent steps instead of 35
#define WALK_THREADS 128

// Two byte ranges in one launch (the head and the tail the filter kernel leaves: blocks [0, blocks1) walk
// [lo, hi), the rest [lo2, hi2)) -- each launch is a latency chain of its own in front of or behind the filter
// kernel, and a 64 KiB scan is nothing but launches.
__global__ void __launch_bounds__(WALK_THREADS)
ac_walk_kernel(DevRules R, ScanView V, CandOut out, int64_t lo, int64_t hi, int chunk, unsigned blocks1 = ~0u, int64_t lo2 = 0,
               int64_t hi2 = 0)
{
  unsigned bx = blockIdx.x;
  if (bx >= blocks1)
  {
    bx -= blocks1;
    lo = lo2;
    hi = hi2;
  }
  const int64_t a = lo + ((int64_t) bx * WALK_THREADS + threadIdx.x) * chunk;

  if (a >= hi)
    return;

  const int64_t b = (a + chunk < hi) ? a + chunk : hi;

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

  const int chunk = (hi - lo) < (1 << 20) ? WALK_CHUNK_TINY : ((hi - lo) < (8 << 20) ? WALK_CHUNK_SMALL : WALK_CHUNK);
  const int64_t nthreads = (hi - lo + chunk - 1) / chunk;
  const unsigned blocks = (unsigned) ((nthreads + WALK_THREADS - 1) / WALK_THREADS);
  ac_walk_kernel<<<blocks, WALK_THREADS, 0, s>>>(R, V, out, lo, hi, chunk);
}

// ---------------------------------------------------------------- filter kernel
//
// Why: the table walk needs >= 2 dependent 4-byte gathers per input byte
// (T[state+index], match_table[state]); at 50 % of HBM bandwidth that would be
// ~26 TB/s of bank-conflicting on-chip gathers (SURVEY.md section 7 "Hard
// parts").  But the walk's OUTPUT is sparse: position i has an event iff
// match_table[S(i)] != 0, and S(i) depends only on the 4-byte window
// W(p) = data[p-3..p], p = i-1.  So the kernel filters windows instead of
// walking states:
//
//   level 1  B3: 2^20-bit bitmap in shared memory, probed ONCE PER TWO
//            POSITIONS with the 3 bytes data[p-2..p] the windows W(p) and W(p+1)
//            have in common; set for the suffix3 and the prefix3 of every hit
//            window.  Branch-free: one IMAD hash, one LDS, the bit lands in a
//            per-thread mask.
//   level 2  B3 again with the OTHER 3-byte key of each of the pair's two windows, then
//            B4: 2^18- or 2^19-bit Bloom filter (k = 4, per kernel variant) in shared
//            memory over the 4-byte window; only for pairs that passed level 1.
//   level 3  exact: open-addressed hash table (window -> match_table value of
//            the state the reference walk would be in), L2-resident; then
//            emit_events() exactly as the walk kernel does.
//
// "Hit windows" are enumerated on the host from the compiled automaton
// (build_filter_tables): every AC state with a non-empty match list and depth
// 3 or 4 contributes the 4-byte windows that end in its path.  Rulesets with hit
// states of depth <= 2 (1-2 byte atoms, or strings with no atom at all, quirk 7)
// would saturate the bitmaps: they are scanned by the walk kernel instead.

#define FT_ROW 528                       // per warp, per u: 12 B pad | 4 B halo | 512 B data
#define HT_MUL 0xC2B2AE35u

// Kernel variants (threads per CTA, 16-byte chunks per thread per tile, log2 of the
// B4 Bloom filter size); all three fit the 227 KB of shared memory with B3.
struct FtVariant { int threads, u, b4_bits, tma_bufs; };   // tma_bufs: 0 = LDG + staging rows (round 1), else TMA buffers per warp
static const FtVariant kVariants[] = {{1024, 4, 18, 0}, {512, 4, 19, 0}, {1024, 2, 19, 0},
                                      {1024, 4, 18, 1}, {768, 4, 18, 1},    // 3-4: variant T
                                      {1024, 4, 0, 1},                      // 5: variant T for dense rulesets (below)
                                      {1024, 4, 0, 2}};                     // 6: variant T for rulesets with short atoms (below)
#define FT_NUM_VARIANTS 7
#define FT_DEFAULT_VARIANT 3
#define FT_DENSE_VARIANT 5
#define FT_SHORT_VARIANT 6
// Dense rulesets (100,000 strings: 190k hit windows): B4 at 2^18 bits would be ~95 % full and filter nothing, and
// B3 at 2^20 bits is a third full (12 % of the pairs pass level 1, and level 2 -- ~100 instructions per survivor
// round and an L2 lookup for what passes -- becomes the kernel).  Variant 5 gives ALL the table space to B3:
// 40,960 words (1.25 x), indexed by the high product of the hash and the word count instead of its top 15 bits,
// and no B4.  Chosen automatically when the standard B3 would be more than FT_DENSE_FILL full.
#define B3_WORDS_DENSE 40960u
// Rulesets with SHORT atoms (round 1: walk kernel, 0.025 of the HBM peak).  A hit state of depth 2 (a 2-byte atom
// "xy") makes every window that ends in xy a hit, a depth-1 state every window that ends in y, and more than 512
// depth-3 states are too many to enumerate their 256 windows each: none of that fits a filter over 3-byte keys
// shared by two positions.  Variant 6 adds B2, an EXACT 2^16-bit bitmap (8 KB of the table space, taken from B3)
// over the last two bytes of a window, probed once per position: it is set for the last two bytes of every such
// state's path (depth 1: for all 256 first bytes).  A position B2 flags gets the exact treatment -- its state is
// walked from the root over its four bytes (exact: depth <= 4), match_table looked up, events emitted -- and B2 is
// authoritative for it: the B3/B4/hash-table path skips flagged positions, so nothing is emitted twice.
// B2_WORDS (2048 words = 2^16 bits) is defined in engine.cuh
#define B3_WORDS_SHORT (B3_WORDS_DENSE - B2_WORDS - 64u)
#define FT_DENSE_FILL 0.10f   // the 20k-string set (0.108) gains 1 % as well (16 GiB: 0.505 -> 0.510); below that nothing changes
static inline uint32_t ft_b3_words(const FtVariant& v)
{ return v.b4_bits != 0 ? B3_WORDS : (v.tma_bufs == 2 ? B3_WORDS_SHORT : B3_WORDS_DENSE); }
static inline uint32_t ft_table_words(const FtVariant& v)   // B3 (+ B2 + pad) as laid out in shared memory
{ return v.b4_bits != 0 ? B3_WORDS : B3_WORDS_DENSE; }
static inline size_t ft_smem_bytes(const FtVariant& v)
{
  return (size_t) ft_table_words(v) * 4 + (v.b4_bits == 0 ? 0 : ((size_t) 1 << (v.b4_bits - 3))) +
         (v.tma_bufs == 0 ? (size_t) (v.threads / 32) * (v.u * FT_ROW + 32)
                          : (size_t) (v.threads / 32) * (16 + 2048 + 16 + 64));
}

struct FilterDev
{
  const uint32_t* b3;
  const uint32_t* b4;
  const uint32_t* b2;  // variant 6 only
  const uint2* ht;     // {window, match_table value}; .y == 0 means empty
  uint32_t ht_bits;
};

static __device__ __forceinline__ uint32_t ht_lookup(const FilterDev& F, uint32_t w)
{
  const uint32_t mask = (1u << F.ht_bits) - 1;
  uint32_t slot = (w * HT_MUL) >> (32 - F.ht_bits);
  for (;;)
  {
    const uint2 e = __ldg(&F.ht[slot]);
    if (e.y == 0) return 0;
    if (e.x == w) return e.y;
    slot = (slot + 1) & mask;
  }
}

// B4 is a Bloom filter with BLOOM_K probes (Kirsch-Mitzenmacher double hashing):
// the level-2 test must be selective enough that the L2-resident exact lookup
// (hundreds of cycles of latency for the whole warp) is reached almost only by
// real hits: 2^19 bits for ~36k windows and k = 4 gives ~0.3 % false positives.
// Integer multiplies run on the FMA pipe; shifts/logic on the (saturated) ALU pipe.
// Inline PTX keeps ptxas from strength-reducing them back into shifts.
// mask * 2 + bit on the FMA pipe (the ALU pipe is the busy one in the filter kernel)
static __device__ __forceinline__ uint32_t push_bit_fma(uint32_t mask, uint32_t bit)
{
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(mask), "r"(bit));
  return r;
}

// Level 2 reuses B3: every hit window has BOTH of its 3-byte keys in it (its last three bytes,
// probed by level 1 when the window is the first of its pair, and its first three, probed when
// it is the second), so for a surviving pair the other key of each window is a second,
// nearly independent test at the level-1 false-positive rate -- one LDS instead of the
// BLOOM_K dependent probes of B4, which only has to vet what passes both.
static __device__ __forceinline__ bool b3_test(const uint32_t* sB3, uint32_t key3)
{
  const uint32_t w = sB3[((key3 << 8) * HASH3_MUL) >> 17];
  return ((w >> (key3 & 31)) & (w >> ((key3 >> 16) & 31)) & 1u) != 0;
}

template <int B4BITS>
static __device__ __forceinline__ bool bloom4_test(const uint32_t* sB4, uint32_t w)
{
  const uint32_t g1 = w * HASH4_MUL, g2 = (w * HASH4_MUL2) | 1u;
  uint32_t g = g1;
#pragma unroll
  for (int k = 0; k < BLOOM_K; k++)
  {
    const uint32_t i4 = g >> (32 - B4BITS);
    if (!((sB4[i4 >> 5] << (i4 & 31)) & 0x80000000u))
      return false;
    g += g2;
  }
  return true;
}

// B3 is a blocked Bloom filter with two bits per key inside ONE 32-bit word (one LDS
// per probe): word = top 15 bits of hash3(key), bit n1 = low 5 bits of the key's first
// byte, bit n2 = low 5 bits of its third byte -- both are the low bits of registers the
// probe already holds, and the shift instruction masks its count to 5 bits for free.
// `(w >> n1) & (w >> n2) & 1` is two shifts and one LOP3; the result is appended to the
// per-thread survivor mask with an integer multiply-add, which runs on the FMA pipe
// (the ALU pipe is the one this kernel saturates).
// B4 (MSB-first bit numbering): word = idx >> 5, n = idx & 31.
template <int FT_THREADS, int FT_U, int B4BITS>
__global__ void __launch_bounds__(FT_THREADS, 1)
ac_filter_kernel(DevRules R, FilterDev F, ScanView V, CandOut out, int64_t lo, int64_t hi)
{
  constexpr int FT_TILE = FT_THREADS * 16 * FT_U;
  constexpr uint32_t B4_WORDS = 1u << (B4BITS - 5);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* sB3 = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* sB4 = sB3 + B3_WORDS;
  uint8_t* stage = smem_raw + (B3_WORDS + B4_WORDS) * 4;

  {
    const uint4* g3 = reinterpret_cast<const uint4*>(F.b3);
    const uint4* g4 = reinterpret_cast<const uint4*>(F.b4);
    uint4* s3 = reinterpret_cast<uint4*>(sB3);
    uint4* s4 = reinterpret_cast<uint4*>(sB4);
    for (uint32_t i = threadIdx.x; i < B3_WORDS / 4; i += FT_THREADS) s3[i] = __ldg(&g3[i]);
    for (uint32_t i = threadIdx.x; i < B4_WORDS / 4; i += FT_THREADS) s4[i] = __ldg(&g4[i]);
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  uint8_t* wstage = stage + warp * (FT_U * FT_ROW);   // this warp's private rows
  uint8_t* wqueue = stage + (FT_THREADS / 32) * (FT_U * FT_ROW) + warp * 32;   // and its survivor queue

  const uint8_t* __restrict__ d = V.d - V.g0; // d[p] = block byte p; (d + lo) is 16-byte aligned
  const int64_t ntiles = (hi - lo) / FT_TILE;
  const uint32_t H3 = HASH3_MUL << 8;

  // [lo, hi) is a whole number of tiles (the host gives the remainder to the walk kernel),
  // so no load needs a bounds check
  const uint8_t* tp = d + lo + (int64_t) blockIdx.x * FT_TILE + (int64_t) threadIdx.x * 16;
  const int64_t tstep = (int64_t) gridDim.x * FT_TILE;

  uint4 v[FT_U];
  uint32_t halo[FT_U];

#define LOAD_TILE(P_)                                                                              \
  _Pragma("unroll") for (int u = 0; u < FT_U; u++)                                                 \
  {                                                                                                \
    v[u] = *reinterpret_cast<const uint4*>((P_) + u * (FT_THREADS * 16));                          \
    halo[u] = 0;                                                                                   \
    if (lane == 0) halo[u] = *reinterpret_cast<const uint32_t*>((P_) + u * (FT_THREADS * 16) - 4); \
  }

  if ((int64_t) blockIdx.x < ntiles) { LOAD_TILE(tp); }

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tp += tstep)
  {

    // ---- level 1: one B3 probe per pair of positions, branch-free
    uint32_t mask = 0;
    const bool has_next = tile + gridDim.x < ntiles;

#define PROBE3(x, y)                                                                \
  {                                                                                 \
    const uint32_t h_ = (x) * H3;                                                   \
    const uint32_t w_ = sB3[h_ >> 17];                                              \
    const uint32_t t_ = __funnelshift_r(w_, 0u, (x)) & __funnelshift_r(w_, 0u, (y)) & 1u; \
    mask = push_bit_fma(mask, t_);                                                  \
  }

#pragma unroll
    for (int u = 0; u < FT_U; u++)
    {
      uint32_t m = __shfl_up_sync(0xFFFFFFFFu, v[u].w, 1);
      if (lane == 0) m = halo[u];
      const uint32_t w0 = v[u].x, w1 = v[u].y, w2 = v[u].z, w3 = v[u].w;

      // stage the chunk (and the warp's halo word) for level 2
      *reinterpret_cast<uint4*>(wstage + u * FT_ROW + 16 + lane * 16) = v[u];
      if (lane == 0) *reinterpret_cast<uint32_t*>(wstage + u * FT_ROW + 12) = m;

      // key = 3 bytes; x holds them in its low 24 bits, y is any register whose low
      // byte is the key's third byte (both bit indices come for free as shift counts)
      const uint32_t f0 = __funnelshift_r(m, w0, 16), f1 = __funnelshift_r(w0, w1, 16);
      const uint32_t f2 = __funnelshift_r(w1, w2, 16), f3 = __funnelshift_r(w2, w3, 16);
      PROBE3(f0, w0);
      PROBE3(w0, f1);
      PROBE3(f1, w1);
      PROBE3(w1, f2);
      PROBE3(f2, w2);
      PROBE3(w2, f3);
      PROBE3(f3, w3);
      PROBE3(w3, w3 >> 16);

      // this chunk's registers are free again (level 2 reads the staged copy): put the next
      // tile's load in flight now, so that HBM latency hides behind the rest of the tile
      if (has_next)
      {
        v[u] = *reinterpret_cast<const uint4*>(tp + tstep + u * (FT_THREADS * 16));
        halo[u] = 0;
        if (lane == 0) halo[u] = *reinterpret_cast<const uint32_t*>(tp + tstep + u * (FT_THREADS * 16) - 4);
      }
    }
#undef PROBE3
    // mask bit (8*FT_U - 1 - slot) belongs to slot = u*8 + pair

    __syncwarp();

    // ---- levels 2 and 3 for the pairs that passed.
    // A lane finds 0.6 survivors per tile on average but the slowest lane of a warp
    // 2.6, and a SIMT loop runs as long as its slowest lane: the warp's survivors are
    // therefore dealt out evenly.  Each lane publishes (lane, n-th survivor) bytes at
    // its prefix-sum position in a 32-entry warp queue; lane e then takes entry e,
    // fetches the owner's mask by shuffle and the window bytes from the owner's
    // staged row.
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, dd);
      if (lane >= (uint32_t) dd) incl += t;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t excl = incl - cnt;
    const bool dense = __any_sync(0xFFFFFFFFu, cnt > 8);   // (lane, n) must fit a byte: n < 8
    const int64_t wpos = (int64_t) (tp - d) - (int64_t) lane * 16;   // block position of the warp's first byte

#define LEVEL23(OL_, S_)                                                                            \
  {                                                                                                 \
    const int u_ = (S_) >> 3;                                                                       \
    const int k_ = ((S_) & 7) * 2;                                                                  \
    /* bytes p-3 .. p+1 of the pair's first position p, from the owner's staged row */              \
    const uint32_t off_ = 16 + (OL_) * 16 + k_ - 3;                                                 \
    const uint8_t* row_ = wstage + u_ * FT_ROW;                                                     \
    const uint32_t x0_ = *reinterpret_cast<const uint32_t*>(row_ + (off_ & ~3u));                   \
    const uint32_t x1_ = *reinterpret_cast<const uint32_t*>(row_ + (off_ & ~3u) + 4);               \
    const uint64_t x_ = ((uint64_t) x1_ << 32) | x0_;                                               \
    const uint32_t sh_ = (off_ & 3u) * 8;                                                           \
    _Pragma("unroll") for (int e_ = 0; e_ < 2; e_++)                                                \
    {                                                                                               \
      const uint32_t w_ = (uint32_t) (x_ >> (sh_ + 8 * e_)); /* W(p + e) */                         \
      if (b3_test(sB3, e_ == 0 ? (w_ & 0xFFFFFFu) : (w_ >> 8)) && bloom4_test<B4BITS>(sB4, w_))     \
      {                                                                                             \
        const uint32_t mv_ = ht_lookup(F, w_);                                                      \
        if (mv_ != 0)                                                                               \
          emit_events<false>(R, V, out, wpos + (int64_t) u_ * (FT_THREADS * 16) + (OL_) * 16 + k_ + e_ + 1, mv_); \
      }                                                                                             \
    }                                                                                               \
  }

    if (!dense)
    {
      for (uint32_t base_e = 0; base_e < total; base_e += 32)   // almost always a single round
      {
        for (uint32_t r = 0; r < cnt; r++)
        {
          const uint32_t e = excl + r - base_e;
          if (e < 32) wqueue[e] = (uint8_t) ((lane << 3) | r);
        }
        __syncwarp();

        const bool active = base_e + lane < total;
        const uint32_t dsc = active ? wqueue[lane] : 0;
        const uint32_t ol = dsc >> 3;
        uint32_t om = __shfl_sync(0xFFFFFFFFu, mask, ol);
        for (uint32_t r = dsc & 7; r > 0; r--)
          om &= ~(0x80000000u >> __clz(om));
        if (active)
        {
          const int sl = __clz(om) - (32 - 8 * FT_U);
          LEVEL23(ol, sl);
        }
        __syncwarp();
      }
    }
    else
    {
      // pathologically dense data: every lane works through its own survivors
      while (mask != 0)
      {
        const int sl = __clz(mask) - (32 - 8 * FT_U);
        mask &= ~(0x80000000u >> (sl + 32 - 8 * FT_U));
        LEVEL23(lane, sl);
      }
    }
#undef LEVEL23

    __syncwarp();
  }
#undef LOAD_TILE
}


// ---- variant T: the same three levels, fed by TMA.
//
// ac_filter_kernel moves every input byte LDG.128 -> registers -> STS.128 (staging for level 2):
// per 2 KB warp-tile that is ~26 global + 20 shared wavefronts on the l1tex data pipe, which the
// kernel saturates (96 %, profiles/r01_ncu_filter_final2_16gib_summary.txt), beside the ~120 of
// the level-1 gathers.  Here each warp owns a buffer of 16 + 2048 bytes; one elected lane
// fetches a whole warp-tile (plus the 16 bytes in front of it: the halo travels with the data,
// no second read) with ONE cp.async.bulk that completes on the buffer's mbarrier.  Level 1 reads
// its chunks back with conflict-free LDS.128; the LDGs, the halo LDGs and all staging stores are
// gone.  The buffer is handed back to TMA as soon as the tile is in registers -- level 2 fetches
// the five bytes of a surviving window from global memory (an L2 hit: the tile has just been
// read) -- so the copy of the next tile has a whole tile's worth of work to land.
// Measured steps (profiles/r02_kernel_sweep*.txt, 4 GiB, scan kernel GB/s; round-1 kernel 2755):
//   level 2 straight from registers, no staging at all        2055 (-23 % shared wavefronts, +42 % instructions)
//   4-slot ring of 512-byte TMA copies + that level 2          1776
//   one TMA copy per warp-tile, refill after level 2           2799 (mbarrier wait = 14 % of stall samples)
//   same with two buffers on 16 warps / 24 warps one buffer    2441 / 2560
//   refill right after the LDS.128s, level 2 from L2           2829 (16 GiB: 2991 vs 2895; no pipe saturated any more:
//                                                                    LSU 82 %, ALU 78 %, issue 74 %)
//   + 16-bit queue descriptors, window-address arithmetic, word address on the FMA pipe: see DESIGN.md
#define FT_T_TILE 2048            // bytes per warp and round
#define FT_T_BUF  (16 + FT_T_TILE)

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
static __device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory"); }
static __device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(mbar), "r"(parity) : "memory");
}

// shared-memory accesses by 32-bit window address: one base register for the whole kernel (a
// generic pointer makes the compiler rebuild the window base from %cluster_ctaid in every
// divergent region), and non-volatile loads of the read-only tables that ptxas may schedule freely
static __device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{ uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ uint32_t lds_u16_v(uint32_t a)
{ uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
static __device__ __forceinline__ void sts_u16_v(uint32_t a, uint32_t v)
{ asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// idx * 4 + base on the FMA pipe (an LEA or SHF + LOP3 would sit on the ALU pipe, the busier one).
// `four` is the constant 4 in a register the compiler cannot see through: with an immediate 4
// ptxas strength-reduces the multiply-add back into an LEA (ALU pipe).
static __device__ __forceinline__ uint32_t word_addr_fma(uint32_t idx, uint32_t four, uint32_t base)
{ uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(idx), "r"(four), "r"(base)); return r; }

template <uint32_t NW>   // B3_WORDS: top 15 bits of the hash; anything else: high product with the word count
static __device__ __forceinline__ bool b3_test_a(uint32_t sB3_a, uint32_t key3)
{
  const uint32_t h = (key3 << 8) * HASH3_MUL;
  const uint32_t idx = NW != B3_WORDS ? __umulhi(h, NW) : (h >> 17);
  const uint32_t w = lds_u32(sB3_a + (idx << 2));
  return ((w >> (key3 & 31)) & (w >> ((key3 >> 16) & 31)) & 1u) != 0;
}

template <int B4BITS>
static __device__ __forceinline__ bool bloom4_test_a(uint32_t sB4_a, uint32_t w)
{
  const uint32_t g1 = w * HASH4_MUL, g2 = (w * HASH4_MUL2) | 1u;
  uint32_t g = g1;
#pragma unroll
  for (int k = 0; k < BLOOM_K; k++)
  {
    const uint32_t i4 = g >> (32 - B4BITS);
    if (!((lds_u32(sB4_a + (i4 >> 5) * 4) << (i4 & 31)) & 0x80000000u))
      return false;
    g += g2;
  }
  return true;
}

template <int NWARPS, int B4BITS, bool SHORT = false>
__global__ void __launch_bounds__(NWARPS * 32, 1)
ac_filter_tma_kernel(DevRules R, FilterDev F, ScanView V, CandOut out, int64_t lo, int64_t hi)
{
  constexpr int FT_THREADS = NWARPS * 32, FT_U = FT_T_TILE / 512;
  constexpr int FT_TILE = NWARPS * FT_T_TILE;         // bytes per CTA and round; warp w owns [w * 2048, (w + 1) * 2048) of it
  constexpr bool DENSE = (B4BITS == 0);               // all table space to B3, no B4 (see FT_DENSE_VARIANT)
  constexpr uint32_t B3W = SHORT ? B3_WORDS_SHORT : (DENSE ? B3_WORDS_DENSE : B3_WORDS);
  constexpr uint32_t TABW = DENSE ? B3_WORDS_DENSE : B3_WORDS;   // B3 (+ B2 + pad, variant 6) in shared memory
  constexpr uint32_t B4_WORDS = DENSE ? 0u : (1u << ((DENSE ? 5 : B4BITS) - 5));
  static_assert(!SHORT || DENSE, "the short-atom variant takes B2 out of the dense layout");
  constexpr int WARP_BYTES = FT_T_BUF + 16 + 64;      // buffer | mbarrier (8 B) + pad | survivor queue (32 x u16)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t* sB3 = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* sB4 = sB3 + TABW;
  uint8_t* stage = smem_raw + (TABW + B4_WORDS) * 4;

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  const uint8_t* wbuf = stage + warp * WARP_BYTES;    // 16 B halo | 4 x 512 B chunks, contiguous
  const uint32_t sB3_a = smem_u32(sB3);
  const uint32_t sB4_a = sB3_a + TABW * 4;
  const uint32_t sB2_a = sB3_a + B3W * 4;   // variant 6: B2 sits behind B3
  const uint32_t wbuf_a = smem_u32(wbuf);
  const uint32_t wbar_a = wbuf_a + FT_T_BUF;
  const uint32_t wq_a = wbuf_a + FT_T_BUF + 16;

  if (lane == 0)
  {
    mbar_init(wbar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  {
    const uint4* g3 = reinterpret_cast<const uint4*>(F.b3);
    const uint4* g4 = reinterpret_cast<const uint4*>(F.b4);
    uint4* s3 = reinterpret_cast<uint4*>(sB3);
    uint4* s4 = reinterpret_cast<uint4*>(sB4);
    for (uint32_t i = threadIdx.x; i < B3W / 4; i += FT_THREADS) s3[i] = __ldg(&g3[i]);
    for (uint32_t i = threadIdx.x; i < B4_WORDS / 4; i += FT_THREADS) s4[i] = __ldg(&g4[i]);
    if (SHORT)
    {
      const uint4* g2 = reinterpret_cast<const uint4*>(F.b2);
      uint4* s2 = reinterpret_cast<uint4*>(sB3 + B3W);
      for (uint32_t i = threadIdx.x; i < B2_WORDS / 4; i += FT_THREADS) s2[i] = __ldg(&g2[i]);
    }
  }
  __syncthreads();

  const uint8_t* __restrict__ d = V.d - V.g0; // d[p] = block byte p; (d + lo) is 16-byte aligned and lo >= 16
  // [lo, hi) is a whole number of WARP-tiles; the warps of a CTA share nothing inside the loop, so the last round
  // may be a partial CTA-tile: this warp has rounds `tile` = blockIdx.x, blockIdx.x + gridDim.x, ... < ntiles
  const int64_t nwt = (hi - lo) / FT_T_TILE;
  const int64_t ntiles = (nwt - (int64_t) warp + NWARPS - 1) / NWARPS;
  const uint32_t H3 = HASH3_MUL << 8;
  const int64_t tstep = (int64_t) gridDim.x * FT_TILE;
  const uint32_t four = 4u + (F.ht_bits >> 31);   // == 4 (ht_bits <= 28), opaque to the compiler: see word_addr_fma

  // this warp's tile of the current round (block bytes [tp - d, tp - d + 2048))
  const uint8_t* tp = d + lo + (int64_t) blockIdx.x * FT_TILE + (int64_t) warp * FT_T_TILE;

  if (lane == 0 && (int64_t) blockIdx.x < ntiles)
  {
    mbar_expect_tx(wbar_a, FT_T_BUF);
    bulk_g2s(wbuf_a, tp - 16, FT_T_BUF, wbar_a);
  }

  uint32_t parity = 0;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tp += tstep)
  {
    mbar_wait(wbar_a, parity);
    parity ^= 1u;

    uint4 vv[FT_U];
    uint32_t mm[FT_U];
#pragma unroll
    for (int u = 0; u < FT_U; u++)
    {
      vv[u] = *reinterpret_cast<const uint4*>(wbuf + 16 + u * 512 + lane * 16);
      mm[u] = __shfl_up_sync(0xFFFFFFFFu, vv[u].w, 1);
      if (lane == 0) mm[u] = *reinterpret_cast<const uint32_t*>(wbuf + 12 + u * 512);
    }

    // The tile is in registers (the shuffles above made every lane wait for its loads) and level 2
    // takes its window bytes from global memory (L2: the tile has just been read), so the buffer
    // can be refilled NOW, a whole tile's worth of work ahead of its next use.
    __syncwarp();
    if (lane == 0 && tile + (int64_t) gridDim.x < ntiles)
    {
      mbar_expect_tx(wbar_a, FT_T_BUF);
      bulk_g2s(wbuf_a, tp + tstep - 16, FT_T_BUF, wbar_a);
    }

    // ---- level 1: one B3 probe per pair of positions, branch-free.  Per probe: IMAD (hash), SHF,
    // IMAD (word address), LDS, 2 SHF + LOP3 (the two bits), IMAD (append): 4 ALU-pipe instructions.
    uint32_t mask = 0;
    uint32_t m2[FT_U];   // variant 6: B2 survivors, 16 positions per chunk

#define PROBE3(x, y)                                                                \
  {                                                                                 \
    const uint32_t h_ = (x) * H3;                                                   \
    const uint32_t w_ = lds_u32(word_addr_fma(DENSE ? __umulhi(h_, B3W) : (h_ >> 17), four, sB3_a)); \
    const uint32_t t_ = __funnelshift_r(w_, 0u, (x)) & __funnelshift_r(w_, 0u, (y)) & 1u; \
    mask = push_bit_fma(mask, t_);                                                  \
  }

#pragma unroll
    for (int u = 0; u < FT_U; u++)
    {
      const uint32_t m = mm[u];
      const uint32_t w0 = vv[u].x, w1 = vv[u].y, w2 = vv[u].z, w3 = vv[u].w;
      const uint32_t f0 = __funnelshift_r(m, w0, 16), f1 = __funnelshift_r(w0, w1, 16);
      const uint32_t f2 = __funnelshift_r(w1, w2, 16), f3 = __funnelshift_r(w2, w3, 16);
      PROBE3(f0, w0);
      PROBE3(w0, f1);
      PROBE3(f1, w1);
      PROBE3(w1, f2);
      PROBE3(f2, w2);
      PROBE3(w2, f3);
      PROBE3(f3, w3);
      PROBE3(w3, w3 >> 16);

      if (SHORT)
      {
        // B2: one exact probe per POSITION with its last two bytes; bit (15 - j) of m2[u] <- byte j of the chunk
        uint32_t t2 = 0;
        const uint32_t ww[5] = {m, w0, w1, w2, w3};
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
          const uint32_t cur = ww[q + 1];
          const uint32_t k0 = __funnelshift_r(ww[q], cur, 24) & 0xFFFFu;   // (prev.b3, cur.b0)
          const uint32_t k1 = cur & 0xFFFFu, k2 = (cur >> 8) & 0xFFFFu, k3 = cur >> 16;
          t2 = t2 * 2 + ((lds_u32(sB2_a + ((k0 >> 5) << 2)) >> (k0 & 31)) & 1u);
          t2 = t2 * 2 + ((lds_u32(sB2_a + ((k1 >> 5) << 2)) >> (k1 & 31)) & 1u);
          t2 = t2 * 2 + ((lds_u32(sB2_a + ((k2 >> 5) << 2)) >> (k2 & 31)) & 1u);
          t2 = t2 * 2 + ((lds_u32(sB2_a + ((k3 >> 5) << 2)) >> (k3 & 31)) & 1u);
        }
        m2[u] = t2;
      }
    }
#undef PROBE3
    // mask bit (31 - sl) belongs to slot sl = u * 8 + pair: clz(mask) is the first slot that passed

    // ---- levels 2 and 3.  A lane finds 0.6 survivors per tile on average but the slowest lane of
    // a warp 2.6, so the warp's survivors are dealt out evenly: lane L's n-th survivor goes to
    // queue position (survivors of lanes < L) + n as a 16-bit descriptor (L << 5 | slot); lane e
    // then takes entry e and fetches the five window bytes itself.
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, dd);
      if (lane >= (uint32_t) dd) incl += t;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t pos = incl - cnt;   // queue position of this lane's next survivor
    uint32_t rem = mask;
    const int64_t wpos = (int64_t) (tp - d);   // block position of the warp-tile's first byte

    for (uint32_t base_e = 0; base_e < total; base_e += 32)   // almost always a single round
    {
      while (rem != 0 && pos < base_e + 32)
      {
        const uint32_t sl = (uint32_t) __clz(rem);
        rem &= ~(0x80000000u >> sl);
        sts_u16_v(wq_a + 2 * (pos - base_e), (lane << 5) | sl);
        pos++;
      }
      __syncwarp();

      if (base_e + lane < total)
      {
        const uint32_t dsc = lds_u16_v(wq_a + 2 * lane);
        // first position p of the pair = tile byte u * 512 + owner * 16 + 2 * pair
        const uint32_t pb = ((dsc & 0x18u) << 6) + ((dsc >> 5) << 4) + ((dsc & 7u) << 1);
        const uint8_t* src = tp + pb - 3;   // bytes p-3 .. p+1
        const uint32_t al = (uint32_t) ((uintptr_t) src & 3u);
        const uint32_t x0 = *reinterpret_cast<const uint32_t*>(src - al);
        const uint32_t x1 = *reinterpret_cast<const uint32_t*>(src - al + 4);
        const uint32_t wa = __funnelshift_r(x0, x1, al * 8);          // W(p)
        const uint32_t wb = __funnelshift_rc(x0, x1, al * 8 + 8);     // W(p + 1)
        // variant 6: a position B2 flags is handled by the exact walk below, never here
        if (b3_test_a<B3W>(sB3_a, wa & 0xFFFFFFu) && (DENSE || bloom4_test_a<DENSE ? 18 : B4BITS>(sB4_a, wa)) &&
            !(SHORT && ((lds_u32(sB2_a + (((wa >> 16) >> 5) << 2)) >> ((wa >> 16) & 31)) & 1u)))
        {
          const uint32_t mv = ht_lookup(F, wa);
          if (mv != 0) emit_events<false>(R, V, out, wpos + pb + 1, mv);
        }
        if (b3_test_a<B3W>(sB3_a, wb >> 8) && (DENSE || bloom4_test_a<DENSE ? 18 : B4BITS>(sB4_a, wb)) &&
            !(SHORT && ((lds_u32(sB2_a + (((wb >> 16) >> 5) << 2)) >> ((wb >> 16) & 31)) & 1u)))
        {
          const uint32_t mv = ht_lookup(F, wb);
          if (mv != 0) emit_events<false>(R, V, out, wpos + pb + 2, mv);
        }
      }
      __syncwarp();
    }

    if (SHORT)
    {
      // ---- the positions B2 flagged: the reference walk's state after byte p is a function of bytes p-3..p (trie
      // depth <= 4), so it is walked from the root over those four bytes -- exact for any atom length
#pragma unroll
      for (int u = 0; u < FT_U; u++)
      {
        uint32_t rem2 = m2[u];
        while (rem2 != 0)
        {
          const uint32_t j = (uint32_t) __clz(rem2) - 16u;   // byte j of this lane's chunk u
          rem2 &= ~(0x80000000u >> (j + 16u));
          const uint32_t pb = (uint32_t) u * 512u + lane * 16u + j;   // position p = tile byte pb
          const uint8_t* src = tp + pb - 3;
          uint32_t st = 0;
#pragma unroll
          for (int q = 0; q < 4; q++) st = ac_step(R.T, st, (uint32_t) __ldg(src + q));
          const uint32_t mv = __ldg(&R.M[st]);
          if (mv != 0) emit_events<false>(R, V, out, wpos + pb + 1, mv);
        }
      }
    }
  }
}


// ---- host: enumerate hit windows of the compiled automaton, build B3 / B4 / hash table
//
// The tables are a pure function of the rules image, and building them is the expensive part of
// a rules upload (a breadth-first walk over the trie plus a few hundred thousand hash inserts).
// They are therefore built once per image: kept in a process-wide registry keyed by a 64-bit
// hash of the image (every device of a yrb_multi, every re-upload of the same YR_RULES gets
// them for an upload), and stored in a disk cache (the device-side analogue of
// the reference's compiled-rules file, arena.c:543/627 -- $YR_B200_CACHE_DIR, default
// $XDG_CACHE_HOME or ~/.cache /yara_b200; YR_B200_CACHE=0 turns the disk part off) so that a new
// process that loads the same rules (yr_rules_load of a .yarc) skips the build.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

struct FilterHost
{
  int usable = 0, has_d1 = 0, has_d2 = 0, root_match = 0, variant = 0;
  uint32_t ht_bits = 0, n_windows = 0;
  float fill3 = 0.f;
  std::vector<uint32_t> b3, b4, b2;
  std::vector<uint2> ht;
};

static int ft_variant_index()
{
  const char* e = getenv("YRB_FILTER_VARIANT");
  int v = e ? atoi(e) : FT_DEFAULT_VARIANT;
  return (v < 0 || v >= FT_NUM_VARIANTS) ? FT_DEFAULT_VARIANT : v;
}

static uint32_t host_ac_step(const uint32_t* T, uint32_t state, uint32_t byte)
{
  const uint32_t index = byte + 1;
  uint32_t t = T[state + index];
  while ((t & 0x1FF) != index)
  {
    if (state != 0) { state = T[state] >> 9; t = T[state + index]; }
    else { t = 0; break; }
  }
  return t >> 9;
}

static void filter_host_build(const void* image, FilterHost* out)
{
  const yrb_image_header* h = (const yrb_image_header*) image;
  const uint32_t* T = yrb_img_trans(image);
  const uint32_t* M = yrb_img_match(image);
  out->variant = ft_variant_index();

  if (h->tables_size < 257)
    return;

  if (M[0] != 0)
  {
    out->root_match = 1; // every position is an event: nothing to filter
    return;
  }

  // breadth-first over the trie (depth <= YR_MAX_ATOM_LENGTH), remembering each state's path
  struct Node { uint32_t state; uint32_t path; int depth; };
  std::vector<Node> level, next;
  std::vector<uint32_t> windows; // hit windows, little-endian: data[p-3] in bits 0-7
  std::vector<uint32_t> d3_paths;
  std::vector<uint32_t> short_keys;   // B2 entries: (second-to-last, last) byte of a short hit state's path, last in bits 8-15
  level.push_back({0, 0, 0});
  size_t n_d3 = 0;

  while (!level.empty())
  {
    next.clear();
    for (const Node& n : level)
    {
      if (n.depth > 0 && M[n.state] != 0)
      {
        if (n.depth == 1)
        {
          out->has_d1 = 1;
          for (uint32_t x = 0; x < 256; x++) short_keys.push_back(x | (n.path << 8));
        }
        else if (n.depth == 2)
        {
          out->has_d2 = 1;
          short_keys.push_back(n.path & 0xFFFF);
        }
        else if (n.depth == 3)
        {
          n_d3++;
          d3_paths.push_back(n.path);
        }
        else
        {
          windows.push_back(n.path);
        }
      }
      if (n.depth < YRB_MAX_ATOM_LENGTH)
      {
        for (uint32_t b = 0; b < 256; b++)
        {
          const uint32_t t = T[n.state + b + 1];
          if ((t & 0x1FF) == b + 1)
            next.push_back({t >> 9, n.path | (b << (8 * n.depth)), n.depth + 1});
        }
      }
    }
    level.swap(next);
  }

  const bool short_ok = !getenv("YRB_FILTER_NO_SHORT") && kVariants[out->variant].tma_bufs != 0;
  if (n_d3 > 512 && short_ok)
    for (uint32_t pth : d3_paths) short_keys.push_back((pth >> 8) & 0xFFFF);   // too many to enumerate: their last two bytes go to B2
  else
    for (uint32_t pth : d3_paths)
      for (uint32_t x = 0; x < 256; x++) windows.push_back(x | (pth << 8));

  if (!short_keys.empty() || n_d3 > 512)
  {
    if (!short_ok)
      return; // not usable: the engine falls back to the walk kernel
    out->variant = FT_SHORT_VARIANT;
    out->b2.assign(B2_WORDS, 0);
    for (uint32_t k2 : short_keys) out->b2[k2 >> 5] |= 1u << (k2 & 31);
    size_t pop2 = 0;
    for (uint32_t w : out->b2) pop2 += (size_t) __builtin_popcount(w);
    if (pop2 > 65536 / 8)
    {
      out->b2.clear();
      return; // more than one position in eight would be walked exactly: the walk kernel does that better
    }
  }
  else if (windows.empty())
    return;

  uint32_t ht_bits = 4;
  while ((1ull << ht_bits) < windows.size() * 2) ht_bits++;
  out->ht.assign((size_t) 1 << ht_bits, make_uint2(0, 0));
  const uint32_t ht_mask = (1u << ht_bits) - 1;

  std::vector<uint32_t> live;   // the windows that really are hits
  for (uint32_t w : windows)
  {
    // the state the reference walk is in after these 4 bytes (ROOT restart is exact: depth <= 4)
    uint32_t st = 0;
    for (int j = 0; j < 4; j++) st = host_ac_step(T, st, (w >> (8 * j)) & 0xFF);
    const uint32_t mv = M[st];
    if (mv == 0) continue;
    live.push_back(w);

    uint32_t slot = (w * HT_MUL) >> (32 - ht_bits);
    while (out->ht[slot].y != 0 && out->ht[slot].x != w) slot = (slot + 1) & ht_mask;
    out->ht[slot] = make_uint2(w, mv);
  }

  // B3 / B4 for the kernel variant; the default variant hands over to the dense one when its B3 would be too full
  for (;;)
  {
    const FtVariant& fv = kVariants[out->variant];
    const int b4_bits = fv.b4_bits;
    const uint32_t b3_words = ft_b3_words(fv);
    out->b3.assign(b3_words, 0);
    out->b4.assign(b4_bits ? ((size_t) 1 << (b4_bits - 5)) : 0, 0);
    for (uint32_t w : live)
    {
      const uint32_t keys[2] = {w >> 8, w & 0xFFFFFF};   // suffix3, prefix3
      for (uint32_t k3 : keys)
      {
        // B3 entry of a 3-byte key: word from the hash (its top 15 bits; dense variant: high product with the word
        // count); TWO bits per key, numbered (bit n = 1u << n): low 5 bits of the key's first byte and of its third
        const uint32_t h = (k3 << 8) * HASH3_MUL;
        const uint32_t idx = b4_bits ? (h >> 17) : (uint32_t) (((uint64_t) h * b3_words) >> 32);
        out->b3[idx] |= (1u << (k3 & 31)) | (1u << ((k3 >> 16) & 31));
      }
      if (b4_bits)
      {
        const uint32_t g1 = w * HASH4_MUL, g2 = (w * HASH4_MUL2) | 1u;
        uint32_t g = g1;
        for (int k = 0; k < BLOOM_K; k++)
        {
          const uint32_t i4 = g >> (32 - b4_bits);
          out->b4[i4 >> 5] |= 0x80000000u >> (i4 & 31);
          g += g2;
        }
      }
    }
    size_t pop3 = 0;
    for (uint32_t i = 0; i < b3_words; i++) pop3 += (size_t) __builtin_popcount(out->b3[i]);
    out->fill3 = (float) ((double) pop3 / ((double) b3_words * 32.0));
    const float dense_fill = getenv("YRB_FILTER_DENSE_FILL") ? (float) atof(getenv("YRB_FILTER_DENSE_FILL")) : FT_DENSE_FILL;
    if (out->variant == FT_DEFAULT_VARIANT && out->fill3 > dense_fill && !getenv("YRB_FILTER_NO_DENSE"))
    {
      out->variant = FT_DENSE_VARIANT;
      continue;
    }
    break;
  }
  out->ht_bits = ht_bits;
  out->n_windows = (uint32_t) windows.size();

  if (out->fill3 > (getenv("YRB_FILTER_MAX_FILL") ? atof(getenv("YRB_FILTER_MAX_FILL")) : 0.45))
  {
    out->b3.clear(); out->b4.clear(); out->ht.clear();
    return; // too dense to filter anything
  }
  out->usable = 1;
}

// ---- the cache: process-wide registry + files

#define FT_CACHE_MAGIC 0x42544659u /* "YFTB" */
#define FT_CACHE_VERSION 5u

struct FtCacheHeader
{
  uint32_t magic, version;
  uint64_t image_hash, image_size;
  int32_t usable, has_d1, has_d2, root_match, variant;
  uint32_t ht_bits, n_windows;
  float fill3;
  uint64_t n_b3, n_b4, n_ht, n_b2;
  uint64_t payload_hash;
};

static uint64_t hash64(const void* p, size_t n, uint64_t seed)
{
  // word-wise multiply-xorshift (wyhash-like mixing); only has to tell rules images apart
  const uint8_t* b = (const uint8_t*) p;
  uint64_t h = seed ^ (n * 0x9E3779B97F4A7C15ULL);
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
  {
    uint64_t w;
    memcpy(&w, b + i, 8);
    h = (h ^ w) * 0xFF51AFD7ED558CCDULL;
    h ^= h >> 32;
  }
  uint64_t tail = 0;
  if (i < n) memcpy(&tail, b + i, n - i);
  h = (h ^ tail) * 0xC4CEB9FE1A85EC53ULL;
  return h ^ (h >> 29);
}

static std::string ft_cache_dir()
{
  const char* off = getenv("YR_B200_CACHE");
  if (off && off[0] == '0') return "";
  std::string dir;
  if (const char* e = getenv("YR_B200_CACHE_DIR")) dir = e;
  else if (const char* x = getenv("XDG_CACHE_HOME")) dir = std::string(x) + "/yara_b200";
  else if (const char* hm = getenv("HOME")) dir = std::string(hm) + "/.cache/yara_b200";
  else return "";
  // mkdir -p of the last two components is all a default path needs
  const size_t cut = dir.rfind('/');
  if (cut != std::string::npos && cut > 0) mkdir(dir.substr(0, cut).c_str(), 0700);
  mkdir(dir.c_str(), 0700);
  return dir;
}

static std::string ft_cache_path(const std::string& dir, uint64_t key, int variant)
{
  char name[96];
  snprintf(name, sizeof(name), "/%016llx-v%d.ftab", (unsigned long long) key, variant);
  return dir + name;
}

static uint64_t ft_payload_hash(const FilterHost& f)
{
  uint64_t h = hash64(f.b3.data(), f.b3.size() * 4, 1);
  h = hash64(f.b4.data(), f.b4.size() * 4, h);
  h = hash64(f.b2.data(), f.b2.size() * 4, h);
  return hash64(f.ht.data(), f.ht.size() * sizeof(uint2), h);
}

static bool ft_cache_load(const std::string& path, uint64_t key, uint64_t image_size, int variant, FilterHost* out)
{
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return false;
  FtCacheHeader hd;
  bool ok = fread(&hd, sizeof(hd), 1, fp) == 1 && hd.magic == FT_CACHE_MAGIC && hd.version == FT_CACHE_VERSION &&
            hd.image_hash == key && hd.image_size == image_size && hd.n_b3 <= B3_WORDS_DENSE &&
            hd.n_b2 <= B2_WORDS &&
            (hd.variant == variant || (variant == FT_DEFAULT_VARIANT && (hd.variant == FT_DENSE_VARIANT || hd.variant == FT_SHORT_VARIANT))) &&
            hd.n_b4 <= (1u << 20) && hd.n_ht <= (1ull << 28);
  if (ok)
  {
    out->usable = hd.usable; out->has_d1 = hd.has_d1; out->has_d2 = hd.has_d2; out->root_match = hd.root_match;
    out->variant = hd.variant; out->ht_bits = hd.ht_bits; out->n_windows = hd.n_windows; out->fill3 = hd.fill3;
    out->b3.resize((size_t) hd.n_b3); out->b4.resize((size_t) hd.n_b4); out->ht.resize((size_t) hd.n_ht);
    out->b2.resize((size_t) hd.n_b2);
    ok = (hd.n_b3 == 0 || fread(out->b3.data(), 4, out->b3.size(), fp) == out->b3.size()) &&
         (hd.n_b4 == 0 || fread(out->b4.data(), 4, out->b4.size(), fp) == out->b4.size()) &&
         (hd.n_b2 == 0 || fread(out->b2.data(), 4, out->b2.size(), fp) == out->b2.size()) &&
         (hd.n_ht == 0 || fread(out->ht.data(), sizeof(uint2), out->ht.size(), fp) == out->ht.size()) &&
         ft_payload_hash(*out) == hd.payload_hash;
  }
  fclose(fp);
  if (!ok) *out = FilterHost();
  return ok;
}

static void ft_cache_store(const std::string& path, uint64_t key, uint64_t image_size, const FilterHost& f)
{
  // write to a private name, then rename: readers never see a partial file
  char tmp[64];
  snprintf(tmp, sizeof(tmp), ".tmp.%ld", (long) getpid());
  const std::string tpath = path + tmp;
  FILE* fp = fopen(tpath.c_str(), "wb");
  if (!fp) return;
  FtCacheHeader hd;
  memset(&hd, 0, sizeof(hd));
  hd.magic = FT_CACHE_MAGIC; hd.version = FT_CACHE_VERSION; hd.image_hash = key; hd.image_size = image_size;
  hd.usable = f.usable; hd.has_d1 = f.has_d1; hd.has_d2 = f.has_d2; hd.root_match = f.root_match; hd.variant = f.variant;
  hd.ht_bits = f.ht_bits; hd.n_windows = f.n_windows; hd.fill3 = f.fill3;
  hd.n_b3 = f.b3.size(); hd.n_b4 = f.b4.size(); hd.n_ht = f.ht.size(); hd.n_b2 = f.b2.size();
  hd.payload_hash = ft_payload_hash(f);
  bool ok = fwrite(&hd, sizeof(hd), 1, fp) == 1 &&
            (f.b3.empty() || fwrite(f.b3.data(), 4, f.b3.size(), fp) == f.b3.size()) &&
            (f.b4.empty() || fwrite(f.b4.data(), 4, f.b4.size(), fp) == f.b4.size()) &&
            (f.b2.empty() || fwrite(f.b2.data(), 4, f.b2.size(), fp) == f.b2.size()) &&
            (f.ht.empty() || fwrite(f.ht.data(), sizeof(uint2), f.ht.size(), fp) == f.ht.size());
  ok = (fclose(fp) == 0) && ok;
  if (!ok || rename(tpath.c_str(), path.c_str()) != 0) unlink(tpath.c_str());
}

static std::mutex g_ft_mu;
static std::map<std::pair<uint64_t, int>, std::shared_ptr<const FilterHost>> g_ft_registry;
static unsigned g_ft_stats[3] = {0, 0, 0};   // served from: registry, disk, built

extern "C" void yrb_filter_cache_stats(unsigned* from_registry, unsigned* from_disk, unsigned* built)
{
  std::lock_guard<std::mutex> lk(g_ft_mu);
  if (from_registry) *from_registry = g_ft_stats[0];
  if (from_disk) *from_disk = g_ft_stats[1];
  if (built) *built = g_ft_stats[2];
}

static std::shared_ptr<const FilterHost> filter_host_tables(const void* image, size_t image_size)
{
  const uint64_t key = hash64(image, image_size, 0x59415241ULL);
  const int variant = ft_variant_index();
  std::lock_guard<std::mutex> lk(g_ft_mu);   // building under the lock: concurrent uploads of one image wait instead of building twice
  auto it = g_ft_registry.find(std::make_pair(key, variant));
  if (it != g_ft_registry.end()) { g_ft_stats[0]++; return it->second; }

  std::shared_ptr<FilterHost> f(new FilterHost());
  // files only for rulesets whose tables take a while to build (tens of thousands of AC slots); a small ruleset
  // is rebuilt in microseconds and would only litter the cache directory (the reference's test suite compiles
  // a fresh rule for each of its ~1,500 assertions)
  const bool worth_a_file = ((const yrb_image_header*) image)->tables_size >= (1u << 15);
  const std::string dir = worth_a_file ? ft_cache_dir() : std::string();
  const std::string path = dir.empty() ? std::string() : ft_cache_path(dir, key, variant);
  if (!path.empty() && ft_cache_load(path, key, image_size, variant, f.get()))
  {
    g_ft_stats[1]++;
  }
  else
  {
    filter_host_build(image, f.get());
    g_ft_stats[2]++;
    if (!path.empty()) ft_cache_store(path, key, image_size, *f);
  }
  if (g_ft_registry.size() >= 16) g_ft_registry.erase(g_ft_registry.begin());   // a handful of rulesets per process at most
  g_ft_registry[std::make_pair(key, variant)] = f;
  return f;
}

int build_filter_tables(const void* image, size_t image_size, int device, FilterTables* out, void** d_storage)
{
  (void) device;
  memset(out, 0, sizeof(*out));
  *d_storage = nullptr;

  const std::shared_ptr<const FilterHost> f = filter_host_tables(image, image_size);
  out->has_d1 = f->has_d1; out->has_d2 = f->has_d2; out->root_match = f->root_match; out->variant = f->variant;
  if (!f->usable)
    return YRB_SUCCESS;

  uint8_t* dptr = nullptr;
  const size_t b3_bytes = f->b3.size() * 4, b4_bytes = f->b4.size() * 4 + f->b2.size() * 4 /* B2 rides behind B4 */,
               ht_bytes = f->ht.size() * sizeof(uint2);
  const size_t total = b3_bytes + b4_bytes + ht_bytes;
  if (cudaMalloc((void**) &dptr, total) != cudaSuccess) { cudaGetLastError(); return YRB_ERROR_INSUFFICIENT_MEMORY; }
  cudaMemcpy(dptr, f->b3.data(), b3_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dptr + b3_bytes, f->b4.data(), f->b4.size() * 4, cudaMemcpyHostToDevice);
  if (!f->b2.empty()) cudaMemcpy(dptr + b3_bytes + f->b4.size() * 4, f->b2.data(), f->b2.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dptr + b3_bytes + b4_bytes, f->ht.data(), ht_bytes, cudaMemcpyHostToDevice);

  out->b3 = reinterpret_cast<const uint32_t*>(dptr);
  out->b4 = reinterpret_cast<const uint32_t*>(dptr + b3_bytes);
  out->b2 = f->b2.empty() ? nullptr : reinterpret_cast<const uint32_t*>(dptr + b3_bytes + f->b4.size() * 4);
  out->ht = reinterpret_cast<const void*>(dptr + b3_bytes + b4_bytes);
  out->ht_bits = f->ht_bits;
  out->n_windows = f->n_windows;
  out->fill3 = f->fill3;
  out->usable = 1;
  *d_storage = dptr;

  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev])
  {
    cudaFuncSetAttribute(ac_filter_kernel<1024, 4, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[0]));
    cudaFuncSetAttribute(ac_filter_kernel<512, 4, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[1]));
    cudaFuncSetAttribute(ac_filter_kernel<1024, 2, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[2]));
    cudaFuncSetAttribute(ac_filter_tma_kernel<32, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[3]));
    cudaFuncSetAttribute(ac_filter_tma_kernel<24, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[4]));
    cudaFuncSetAttribute(ac_filter_tma_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[5]));
    cudaFuncSetAttribute(ac_filter_tma_kernel<32, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ft_smem_bytes(kVariants[6]));
    attr_set[dev] = true;
  }
  return YRB_SUCCESS;
}

// Bytes [lo, hi) in block coordinates.  The filter kernel takes the 16-byte
// aligned middle; the head (which also covers the first bytes of the block, where
// a window would reach before byte 0) and the tail go to the walk kernel.
void launch_ac_scan_filter(
    const DevRules& R, const FilterTables& F, const ScanView& V, const CandOut& out, int64_t lo, int64_t hi,
    int num_sms, cudaStream_t s, cudaEvent_t before, cudaEvent_t after)
{
  if (hi <= lo)
    return;

  const uint8_t* d = V.d - V.g0;
  int64_t flo = lo;
  if (flo < 16) flo = 16;
  if (flo < V.g0 + 16) flo = V.g0 + 16;
  flo += (int64_t) ((16 - ((uintptr_t) (d + flo) & 15)) & 15);
  const FtVariant& fv = kVariants[F.variant];
  const int64_t tile_bytes = (int64_t) fv.threads * 16 * fv.u;
  // whole tiles only, the rest is walked: CTA-tiles for the LDG variants, warp-tiles (2 KB) for the TMA variants
  const int64_t unit = fv.tma_bufs ? (int64_t) FT_T_TILE : tile_bytes;
  const int64_t fhi = flo + (hi - flo) / unit * unit;
  const int64_t ntiles = (fhi - flo + tile_bytes - 1) / tile_bytes;

  if (ntiles < 1)
  {
    launch_ac_scan_walk(R, V, out, lo, hi, s);
    return;
  }

  // head [lo, flo) and tail [fhi, hi) in one launch, in front of the filter kernel
  {
    const int chunk = WALK_CHUNK_TINY;
    const int64_t per_block = (int64_t) WALK_THREADS * chunk;
    const unsigned b1 = (unsigned) ((flo - lo + per_block - 1) / per_block), b2 = (unsigned) ((hi - fhi + per_block - 1) / per_block);
    if (b1 + b2 > 0)
      ac_walk_kernel<<<b1 + b2, WALK_THREADS, 0, s>>>(R, V, out, lo, flo, chunk, b1, fhi, hi);
  }

  FilterDev fd;
  fd.b3 = F.b3;
  fd.b4 = F.b4;
  fd.b2 = F.b2;
  fd.ht = reinterpret_cast<const uint2*>(F.ht);
  fd.ht_bits = F.ht_bits;

  const unsigned grid = (unsigned) (ntiles < num_sms ? ntiles : num_sms);
  const size_t smem = ft_smem_bytes(fv);
  if (before) cudaEventRecord(before, s);
  switch (F.variant)
  {
  case 0: ac_filter_kernel<1024, 4, 18><<<grid, 1024, smem, s>>>(R, fd, V, out, flo, fhi); break;
  case 1: ac_filter_kernel<512, 4, 19><<<grid, 512, smem, s>>>(R, fd, V, out, flo, fhi); break;
  case 2: ac_filter_kernel<1024, 2, 19><<<grid, 1024, smem, s>>>(R, fd, V, out, flo, fhi); break;
  case 3: ac_filter_tma_kernel<32, 18><<<grid, 1024, smem, s>>>(R, fd, V, out, flo, fhi); break;
  case 4: ac_filter_tma_kernel<24, 18><<<grid, 768, smem, s>>>(R, fd, V, out, flo, fhi); break;
  case 5: ac_filter_tma_kernel<32, 0><<<grid, 1024, smem, s>>>(R, fd, V, out, flo, fhi); break;
  case 6: ac_filter_tma_kernel<32, 0, true><<<grid, 1024, smem, s>>>(R, fd, V, out, flo, fhi); break;
  default: ac_filter_kernel<1024, 4, 18><<<grid, 1024, smem, s>>>(R, fd, V, out, flo, fhi); break;
  }
  if (after) cudaEventRecord(after, s);
}
