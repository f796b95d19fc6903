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

#include <vector>

#include "yara_b200.h"

// Discovery key (SURVEY.md 8(a) "Order dependence"): lexicographic in
// (i ascending, position in the state's match chain, callback order inside one
// exhaustive backward exec).  39 bits of i (blocks < 2^38 bytes); the low 25 bits hold the chain
// position and the callback sequence, split PER RULESET (DevRules::key_seq_bits, set by
// yrb_key_seq_bits() when the rules are uploaded): the chain position gets as many bits as the
// longest YR_AC_MATCH chain of the ruleset needs, at least 8 -- 300 rules that share one string
// share its atom and put 300 entries on one state's chain; the 100k-string synthetic set has 17 --
// and the sequence gets the rest.  With chains of up to 255 entries that is 17 bits of sequence: the
// fast hex VM's backward position lists keep duplicates (re.c:2323-2333) and one candidate can fire
// tens of thousands of callbacks (tests/cases.py case 53); a candidate that fires more than
// 2^key_seq_bits - 2 fails the scan loudly instead of saturating, and a ruleset whose longest
// chain leaves fewer than 8 bits of sequence (more than 131,071 entries) is refused at upload.
#define KEY_LOW_BITS       25
#define KEY_I_SHIFT        KEY_LOW_BITS
#define KEY_CHAIN_BITS_MIN 8
#define KEY_SEQ_BITS_MIN   8
#define KEY_SEQ_MASK(seq_bits)  ((1u << (seq_bits)) - 1)                      // also the largest sequence number
#define KEY_CHAIN_MAX(seq_bits) ((1u << (KEY_LOW_BITS - (seq_bits))) - 1)     // largest chain position

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
  // (the two share the word that used to hold stack_words alone: the struct -- a parameter of every kernel -- keeps its layout)
  uint16_t stack_words;   // max_re_stack + 1
  uint16_t key_seq_bits;  // discovery key: bits of callback sequence (the chain position has KEY_LOW_BITS - key_seq_bits)
};

// Host side, at upload: the longest YR_AC_MATCH chain of a rules image (chains share their tails --
// ahocorasick.c hangs the failure state's list behind a state's own -- so lengths are memoised per entry) ...
static inline uint32_t yrb_longest_match_chain(const void* image)
{
  const yrb_image_header* h = (const yrb_image_header*) image;
  const yrb_acm* acm = yrb_img_acm(image);
  std::vector<uint32_t> len((size_t) h->num_acm + 1, 0);   // len[k]: entries from acm[k-1] to the end of its chain
  std::vector<uint32_t> path;
  uint32_t best = 0;
  for (uint32_t k0 = 1; k0 <= h->num_acm; k0++)
  {
    if (len[k0]) continue;
    path.clear();
    uint32_t k = k0;
    while (k != 0 && k <= h->num_acm && len[k] == 0 && path.size() <= h->num_acm) { path.push_back(k); k = acm[k - 1].next; }
    uint32_t l = (k != 0 && k <= h->num_acm) ? len[k] : 0;
    for (size_t j = path.size(); j-- > 0;) len[path[j]] = ++l;
    if (len[k0] > best) best = len[k0];
  }
  return best;
}

// ... and the split of the key's low bits it asks for: DevRules::key_seq_bits, or 0 when no split leaves
// KEY_SEQ_BITS_MIN bits of sequence (the ruleset is refused).
static inline uint32_t yrb_key_seq_bits(uint32_t longest_chain)
{
  uint32_t chain_bits = KEY_CHAIN_BITS_MIN;
  while (chain_bits < KEY_LOW_BITS && ((1u << chain_bits) - 1) < longest_chain) chain_bits++;
  return KEY_LOW_BITS - chain_bits >= KEY_SEQ_BITS_MIN ? KEY_LOW_BITS - chain_bits : 0;
}

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
  uint64_t key;      // i << KEY_I_SHIFT
  uint32_t acm_idx;
  uint32_t chain_pos; // position of acm_idx in the chain of the state reached at i: < the ruleset's longest chain
};

// The candidate's discovery key, sequence bits zero (the scan kernels leave the chain position beside the key: they
// need not know the ruleset's split).  chain_pos <= KEY_CHAIN_MAX(seq_bits): yrb_key_seq_bits sized the split for it.
static __host__ __device__ __forceinline__ uint64_t cand_key(const Cand& c, uint32_t seq_bits)
{
  return c.key | ((uint64_t) c.chain_pos << seq_bits);
}

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
  unsigned int key_overflow;    // one candidate fired more than 2^key_seq_bits - 2 match callbacks: its records cannot be ordered
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
static __device__ __forceinline__ void report_error(const RawOut& out, uint32_t sidx, uint64_t cand_key, int code, uint32_t seq_bits)
{
  atomicMin(&out.str_err[sidx], (unsigned long long) ((cand_key & ~(uint64_t) KEY_SEQ_MASK(seq_bits)) | (uint64_t) code));
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
