// engine.cuh -- internal types shared by the kernels of libyrb200.so.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   rules image   one blob, byte-identical to the host image (yrb_image.h)
//   block bytes   the scanned block (or byte-range shard + read margins), u8
//   candidates    Cand[]  16 B: (discovery key, YR_AC_MATCH index); one array
//                 per verifier class (literal / fast hex VM / fiber VM)
//   raw matches   Raw[]   32 B: what _yr_scan_match_callback (scan.c:634) was
//                 called with, plus the discovery key that orders them
//   final matches yrb_match[] 24 B: YR_MATCHES lists, string-major
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "yara_b200.h"

// Discovery key (SURVEY.md 8(a) "Order dependence"): lexicographic in
// (i ascending, position in the state's match chain, callback order inside one
// exhaustive backward exec).  39 bits of i (blocks < 2^38 bytes), 8 bits of chain
// position (the longest YR_AC_MATCH chain of the 100k-string set is 17;
// yrb_rules_create refuses a ruleset with a chain longer than 255), 17 bits of callback
// sequence: the fast hex VM's backward position lists keep duplicates (re.c:2323-2333) and
// one candidate can fire tens of thousands of callbacks (tests/cases.py case 53); a
// candidate that fires more than 131,070 fails the scan loudly instead of saturating.
#define KEY_SEQ_BITS   17
#define KEY_CHAIN_BITS 8
#define KEY_I_SHIFT    (KEY_SEQ_BITS + KEY_CHAIN_BITS)
#define KEY_SEQ_MAX    ((1u << KEY_SEQ_BITS) - 1)
#define KEY_CHAIN_MAX  ((1u << KEY_CHAIN_BITS) - 1)

#define CLS_LITERAL 0
#define CLS_FAST    1
#define CLS_RE      2
#define NUM_CLS     3

// RE_FLAGS_* (re.h:94-102)
#define RE_FLAGS_BACKWARDS  0x04
#define RE_FLAGS_EXHAUSTIVE 0x08
#define RE_FLAGS_WIDE       0x10
#define RE_FLAGS_NO_CASE    0x20
#define RE_FLAGS_DOT_ALL    0x80
#define RE_FLAGS_GREEDY     0x400

struct DevRules
{
  const uint32_t* T;      // transition table
  const uint32_t* M;      // match table
  const yrb_acm* acm;
  const yrb_string* S;
  const uint8_t* lit;
  const uint8_t* code;
  const uint8_t* lower;
  const uint8_t* alter;
  uint32_t tables_size;
  uint32_t num_acm;
  uint32_t num_strings;
  uint32_t stack_words;   // max_re_stack + 1
};

// The bytes a scan can see, in BLOCK coordinates: d[0] is block byte g0, the
// local buffer covers [g0, g1), the whole block is [0, gsize).
struct ScanView
{
  const uint8_t* d;
  int64_t g0;
  int64_t g1;
  int64_t gsize;
  int64_t own_lo;   // events owned: own_lo < i <= own_hi, plus i == 0 if own_lo == 0 ... see ac_scan.cu
  int64_t own_hi;
  uint64_t base;
  int32_t scan_flags;
  // Batch mode (yrb_scan_batch): the "block" is the concatenation of n_seg independent
  // items; item k occupies [seg_start[k], seg_start[k] + seg_len[k]), ascending and disjoint,
  // possibly with a few unowned bytes in between.  0 = ordinary single-block scan.
  int32_t n_seg;
  const int64_t* seg_start;
  const int64_t* seg_len;
};

// the item holding concatenated byte p, or the one before the gap p lies in (-1: before item 0)
static __device__ __forceinline__ int seg_find_in(const int64_t* __restrict__ seg_start, int n_seg, int64_t p)
{
  int lo = 0, hi = n_seg;   // first k in [lo, hi) with seg_start[k] > p
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&seg_start[mid]) > p) hi = mid; else lo = mid + 1;
  }
  return lo - 1;
}

static __device__ __forceinline__ int seg_find(const ScanView& V, int64_t p)
{
  return seg_find_in(V.seg_start, V.n_seg, p);
}

struct Cand
{
  uint64_t key;      // i << KEY_I_SHIFT | chain_pos << KEY_SEQ_BITS
  uint32_t acm_idx;
  uint32_t pad;
};

struct Raw
{
  uint64_t key;       // candidate key | seq
  int64_t offset;     // block-relative match offset
  uint32_t string_idx;
  int32_t length;
  uint32_t misc;      // bits 0-7 xor_key; RAW_SEED: carried over from an earlier block, bits 8-29 = seed index,
                      // key = chain_length (the record was never "discovered" in this block)
  uint32_t group;     // chain head idx for CHAIN_PART strings, else YRB_NONE
};

#define RAW_SEED_IDX_MASK 0x3FFFFFu
#define RAW_SEED      0x80000000u
#define RAW_SEED_MARK 0x40000000u   // kind 1: matches[string] is already non-empty

// Device-side counters / status of one scan (copied to the host with one small D2H).
struct ScanStatus
{
  unsigned long long n_cand[NUM_CLS];
  unsigned long long n_raw;
  unsigned long long err_key_inv; // ~(min over failing candidates of (key & ~seq) | error code); 0 = no error
                                  // (stored inverted so that a zero-filled status needs no second memset)
  unsigned long long n_final;
  unsigned long long n_unconf;
  unsigned long long n_chain_out;
  unsigned int margin_fault;    // a verifier needed bytes outside [g0, g1)
  unsigned int n_too_many;
  unsigned int over_cap;        // some plain string has more than YR_MAX_STRING_MATCHES distinct offsets
  unsigned int ws_overflow;     // a per-candidate work area was too small
  unsigned int survivors;       // filter kernel statistics
  unsigned int n_retry[2];      // candidates queued for the tier-2 verifiers (fast VM, fiber VM)
  int n_sel;                    // surviving records of plain strings (ordered compaction)
  unsigned int any_error;       // some verifier reported an error (see report_error)
  unsigned int dbg[3];          // chain resolver step counters: list-insert walks, prune-loop steps, unvisited-scan steps
  unsigned int n_retry3;        // fast-VM candidates that outgrew tier 2 as well (backward-mode lists hold duplicates)
  unsigned int n_chain_raw;     // Raw records of chained strings (0: the chain kernels are skipped)
  unsigned int err_string1;     // 1 + string of the candidate behind the error (scanner->last_error_string); 0 = none
  unsigned int n_chain_starts;  // chains present among the Raw records (resolve.cu: one warp each)
  unsigned int key_overflow;    // one candidate fired more than KEY_SEQ_MAX match callbacks: its records cannot be ordered
  unsigned int n_retry_re0;     // fiber-VM candidates that outgrew the local-memory tier (verify.cu, tier 0 -> tier 1)
};

// Appends from a warp: the lanes that arrive here together and add to the SAME counter elect one of them to do
// ONE atomicAdd for all (match.any groups them by counter address, popc gives each its rank): on dense data --
// a ruleset whose atoms occur in every other line of a text corpus -- thousands of candidates per tile would
// otherwise queue up on one address.  Returns this lane's slot.
static __device__ __forceinline__ unsigned long long warp_append_slot(unsigned long long* counter)
{
#ifdef YRB_HOST_EMULATION
  return atomicAdd(counter, 1ULL);
#else
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, (unsigned long long) (uintptr_t) counter);
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(peers) - 1;
  unsigned long long base = 0;
  if ((int) lane == leader)
    base = atomicAdd(counter, (unsigned long long) __popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + (unsigned long long) __popc(peers & ((1u << lane) - 1u));
#endif
}

struct CandOut
{
  Cand* cand[NUM_CLS];
  unsigned long long cap[NUM_CLS];
  ScanStatus* st;
};

struct RawOut
{
  Raw* raw;
  unsigned long long cap;
  ScanStatus* st;
  unsigned long long* str_err;  // per string: min over its failing candidates of (candidate key | error code)
  uint8_t* item_dirty;          // batch mode: items whose result needs the general per-item path (else nullptr)
  uint32_t chain_marks_dirty;   // batch mode, YRB_BATCH_BLOCKS_OF_ONE_SCAN: a match callback of a chain part marks its item
};

// Atom prefilter tables of one ruleset (ac_scan.cu, filter kernel).
struct FilterTables
{
  const uint32_t* b3;     // 2^20-bit blocked Bloom filter over 3-byte keys (pair probe, 2 bits per key)
  const uint32_t* b4;     // 2^19-bit Bloom filter (BLOOM_K probes) over 4-byte windows
  const uint32_t* b2;     // variant 6: exact 2^16-bit bitmap over the last two bytes of short hit states' windows
  const void* ht;         // open-addressed {window, match_table value} table
  uint32_t ht_bits;
  uint32_t n_windows;
  float fill3;
  int has_d1, has_d2, root_match, usable, variant;
};

#define B3_BITS 20
#define B3_WORDS (1u << (B3_BITS - 5))
#define B2_WORDS (1u << (16 - 5))
#define HASH3_MUL 0x9E3779B1u
#define HASH4_MUL 0x85EBCA6Bu
#define HASH4_MUL2 0x27D4EB2Fu
#define BLOOM_K 4

__host__ __device__ static inline uint32_t hash3(uint32_t key3) { return ((key3 << 8) * HASH3_MUL) >> (32 - B3_BITS); }

// A verifier error (ERROR_TOO_MANY_RE_FIBERS ...) aborts the reference's scan at that
// candidate -- but only if the reference would have verified the candidate at all
// (scan.c:1018-1023: not when the string was already disabled by the 1M cap, nor in
// fast mode once a SINGLE_MATCH string has matched).  Both conditions are "candidate
// key below a per-string threshold", so the earliest failing candidate per string is
// all that needs to be kept; resolve.cu decides which of them counts.
static __device__ __forceinline__ void report_error(const RawOut& out, uint32_t sidx, uint64_t cand_key, int code)
{
  atomicMin(&out.str_err[sidx], (unsigned long long) ((cand_key & ~(uint64_t) KEY_SEQ_MAX) | (uint64_t) code));
  atomicExch(&out.st->any_error, 1u);
}

// ---- launch helpers implemented across the .cu files -------------------------
void launch_ac_event0(const DevRules& R, const ScanView& V, const CandOut& out, cudaStream_t s);
void launch_ac_scan_walk(const DevRules& R, const ScanView& V, const CandOut& out, int64_t lo, int64_t hi, cudaStream_t s);
void launch_ac_scan_filter(
    const DevRules& R, const FilterTables& F, const ScanView& V, const CandOut& out, int64_t lo, int64_t hi,
    int num_sms, cudaStream_t s, cudaEvent_t before = nullptr, cudaEvent_t after = nullptr);
int build_filter_tables(const void* host_image, size_t image_size, int device, FilterTables* out, void** d_storage);

#define RE_LOCAL_STACK 8   // fiber VM, tier 0 (verify.cu): repeat-counter stack words per fiber held in local memory

struct VerifyWs
{
  void* fast1; size_t fast1_threads;   // tier 1 work areas (small, many threads)
  void* fast2; size_t fast2_threads;   // tier 2 (full limits, few threads)
  void* re1;   size_t re1_threads;
  void* re2;   size_t re2_threads;
  uint32_t* retry_fast;
  uint32_t* retry_re;
  uint32_t* retry_re0;                 // fiber VM, tier 0 -> tier 1 queue (nullptr: no tier 0)
  // Candidates are appended by whichever warp of the scan kernel finds them; verified in that order, the 32
  // lanes of a warp run 32 different strings' code (3.4 lanes active per instruction, a third of the stall
  // samples waiting for instructions: profiles/r02_ncu_verify_re_tier0_cfg5_summary.txt).  Order only matters
  // through the discovery key each candidate carries, so the VM classes are verified in AC-match order: a
  // radix sort by acm index gives sort_perm[c] (nullptr: class not sorted), and a warp's lanes then share the
  // string, its code and most of their control flow.
  uint32_t* sort_k[2];                 // scratch for the sort: keys and values, double buffered, sort_cap entries each
  uint32_t* sort_v[2];
  size_t sort_cap;
  void* sort_tmp; size_t sort_tmp_bytes;
  int sort_cls[NUM_CLS];               // 1: verify this class in sorted order
  uint32_t* retry_fast3;               // tier 2 -> tier 3 queue
  void* fast3; size_t fast3_threads;   // tier 3: allocated only when something reaches it
};

// Side streams and fork/join events of a scanner: the three verifier classes run next to each other, and so
// do the chain replay and the writing-out of the plain strings' lists (all of them latency bound, none fills the GPU).
struct SideStreams
{
  cudaStream_t aux[2];
  cudaEvent_t ev[6];
};

void launch_verify(
    const DevRules& R, const ScanView& V, const CandOut& cands, const unsigned long long n_cand[NUM_CLS],
    const RawOut& out, const VerifyWs& W, cudaStream_t s, unsigned* launches, const SideStreams* side = nullptr);
size_t fast_ws_bytes_per_thread(int tier);
size_t cand_sort_temp_bytes(uint32_t n);
// perm_out (= v_b or v_a, returned) lists the candidate indices in ascending acm order
const uint32_t* launch_cand_sort(
    const Cand* cands, uint32_t n, uint32_t num_acm, uint32_t* k_a, uint32_t* k_b, uint32_t* v_a, uint32_t* v_b, void* tmp,
    size_t tmp_bytes, cudaStream_t s);
void launch_verify_tier2(
    const DevRules& R, const ScanView& V, const CandOut& cands, const RawOut& out, const VerifyWs& W, bool fast, bool re,
    cudaStream_t s, unsigned* launches);
void launch_verify_fast_tier3(
    const DevRules& R, const ScanView& V, const CandOut& cands, const RawOut& out, const VerifyWs& W, int tier,
    cudaStream_t s, unsigned* launches);
size_t re_ws_bytes_per_thread(uint32_t stack_words, int tier);
