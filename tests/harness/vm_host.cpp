/*
 * vm_host.cpp -- TEST INFRASTRUCTURE: the product's verifier device code
 * (yara_b200/csrc/verify.cu: literal compare family, fast hex VM, fiber VM, their tiers;
 * yara_b200/csrc/emit.cuh: candidate emission) compiled for the CPU through
 * tests/harness/cuda_host_shim.h and driven as one thread of a 1x1 grid.
 *
 * It mirrors what api.cu's scan_impl does between the scan kernel and the resolve stage:
 *   Aho-Corasick walk (ac_step / emit_events) -> Cand[] per class
 *   tier 1 kernels -> tier 2 for what overflowed -> tiers 3 and 4 of the fast VM
 * and hands back the Raw records in discovery-key order plus the earliest verifier error, which
 * tests/test_vm_host_cpu.py compares with the oracle's match-callback trace.
 *
 * vmh_scan() goes on through the resolve stage (yara_b200/csrc/resolve_kernels.cuh: sort keys, the
 * (string, offset) selection, FAST_MODE guard, the chain replay with its look-ahead, verifier error
 * selection) in the order of resolve.cu's launch_resolve, with std::stable_sort standing in for the
 * CUB radix sort and a loop for DeviceSelect::Flagged, and returns what yrb_scan_host would: confirmed
 * and unconfirmed matches and the scan's error.  Every kernel is called once per "thread" of the grid
 * it is launched with on the device; chain_resolve_kernel runs with a single lane per chain.
 * vmh_scan_batch() does the same in batch mode (ScanView.n_seg > 0, the device side of yrb_scan_batch).
 * Not covered here: the scan kernels (prefilter, walk).
 *
 * Built by oracle/Makefile into oracle/_ref-independent tests/harness output (libvmhost.so).
 */
#define YRB_HOST_EMULATION 1
#include "cuda_host_shim.h"

#include <algorithm>
#include <unordered_set>
#include <vector>

#include "../../yara_b200/csrc/verify.cu"
#include "../../yara_b200/csrc/emit.cuh"

#define YRB_WARP 1   /* chain_resolve_kernel: one lane per chain here, the collectives of the shim are identities */
#include "../../yara_b200/csrc/resolve_kernels.cuh"

static void set_grid_1x1()
{
  threadIdx.x = blockIdx.x = 0;
  blockDim.x = gridDim.x = 1;
}

/* calls `f` once per thread of a launch of ceil(n / block) blocks of `block` threads */
template <typename F> static void launch_per_thread(uint64_t n, unsigned block, F f)
{
  blockDim.x = block;
  gridDim.x = (unsigned) ((n + block - 1) / block);
  for (uint64_t t = 0; t < (uint64_t) gridDim.x * block; t++)
  {
    blockIdx.x = (unsigned) (t / block);
    threadIdx.x = (unsigned) (t % block);
    f();
  }
  set_grid_1x1();
}

extern "C" {

typedef struct vmh_raw
{
  uint64_t key;
  int64_t offset;
  uint32_t string_idx;
  int32_t length;
  uint32_t xor_key;
  uint32_t group;
} vmh_raw;

typedef struct vmh_stats
{
  uint64_t n_cand[3];
  uint64_t n_raw;         /* before removing re-emitted duplicates */
  uint32_t retry_fast;    /* candidates that went to tier 2 / tier 2 (fiber VM) / tier 3 / tier 4 */
  uint32_t retry_re;
  uint32_t retry_fast3;
  uint32_t retry_fast4;
  int32_t error;          /* earliest verifier error in discovery order, 0 = none */
  uint32_t error_string;
  uint64_t error_key;     /* discovery key (seq bits cleared) of the failing candidate */
} vmh_stats;

static long long run_verify(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, std::vector<Raw>& kept,
    std::vector<unsigned long long>& str_err_out, vmh_stats* stats, int n_seg = 0, const int64_t* seg_start = nullptr,
    const int64_t* seg_len = nullptr, uint8_t* item_dirty = nullptr);

/* Returns the number of distinct records (may exceed cap: call again with a larger buffer), or -1 on a bad image. */
long long vmh_run(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, vmh_raw* out, size_t cap,
    vmh_stats* stats)
{
  std::vector<Raw> kept;
  std::vector<unsigned long long> str_err;
  const long long rc = run_verify(image, data, size, base, scan_flags, kept, str_err, stats);
  if (rc < 0)
    return rc;
  std::stable_sort(kept.begin(), kept.end(), [](const Raw& a, const Raw& b) { return a.key < b.key; });
  size_t n = 0;
  for (size_t k = 0; k < kept.size(); k++)
  {
    if (n < cap)
    {
      out[n].key = kept[k].key;
      out[n].offset = kept[k].offset;
      out[n].string_idx = kept[k].string_idx;
      out[n].length = kept[k].length;
      out[n].xor_key = kept[k].misc & 0xFF;
      out[n].group = kept[k].group;
    }
    n++;
  }
  return (long long) n;
}

typedef struct vmh_result
{
  int32_t error;
  uint32_t error_string;
  uint64_t n_matches;       /* may exceed the capacity given: call again with more room */
  uint64_t n_unconfirmed;
  uint32_t n_too_many;
  uint32_t over_cap;        /* a string is still over YR_MAX_STRING_MATCHES (batch mode: the whole batch is re-scanned per item) */
  uint64_t first_cut_key;   /* 1M cap: discovery key at which the first string was cut (0 = none was) */
  uint64_t error_key;       /* discovery key | code of the verifier error */
  uint32_t n_cuts;
  uint32_t pad;
} vmh_result;

/* verification + resolve: what yrb_scan_host returns for one block (no seeds, no cap second pass) */
static int scan_common(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, yrb_match* matches, size_t cap_m,
    yrb_match* unconfirmed, size_t cap_u, vmh_result* res, vmh_stats* stats, int n_seg, const int64_t* seg_start,
    const int64_t* seg_len, uint8_t* item_dirty, const yrb_seed* seeds = nullptr, uint32_t n_seeds = 0,
    uint32_t* too_many_out = nullptr);

uint32_t vmh_longest_match_chain(const void* image) { return yrb_longest_match_chain(image); }

/* DevRules::key_seq_bits of an image (engine.cuh), for the stand-in engine's error precedence; 0 = refused */
uint32_t vmh_key_seq_bits(const void* image) { return yrb_key_seq_bits(yrb_longest_match_chain(image)); }

/* one block of a multi-block scan: `seeds` = what earlier blocks left behind (yrb_scan_params.seeds);
   too_many_out: room for 1024 string indices */
int vmh_scan_seeds(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, const yrb_seed* seeds, uint32_t n_seeds,
    yrb_match* matches, size_t cap_m, yrb_match* unconfirmed, size_t cap_u, uint32_t* too_many_out, vmh_result* res,
    vmh_stats* stats)
{
  return scan_common(image, data, size, base, scan_flags, matches, cap_m, unconfirmed, cap_u, res, stats, 0, nullptr, nullptr,
                     nullptr, seeds, n_seeds, too_many_out);
}

int vmh_scan(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, yrb_match* matches, size_t cap_m,
    yrb_match* unconfirmed, size_t cap_u, vmh_result* res, vmh_stats* stats)
{
  return scan_common(image, data, size, base, scan_flags, matches, cap_m, unconfirmed, cap_u, res, stats, 0, nullptr, nullptr, nullptr);
}

/* Batch mode (yrb_scan_batch): `data` is the concatenation, item k = [seg_start[k], seg_start[k] + seg_len[k]).
   matches come back string-major with CONCATENATION offsets (the host side of yrb_scan_batch splits them by item);
   item_dirty[k] != 0: the item needs the per-item path. */
static thread_local int g_chain_marks_dirty = 0;   /* YRB_BATCH_BLOCKS_OF_ONE_SCAN for the next vmh_scan_batch of this thread */
void vmh_set_chain_marks_dirty(int on) { g_chain_marks_dirty = on; }
static thread_local const uint32_t* g_skip_strings = nullptr;   /* yrb_scan_params.skip_strings for the next scan of this thread */
static thread_local uint32_t g_n_skip_strings = 0;
void vmh_set_skip_strings(const uint32_t* strings, uint32_t n) { g_skip_strings = strings; g_n_skip_strings = n; }

int vmh_scan_batch(
    const void* image, const uint8_t* data, size_t size, int n_seg, const int64_t* seg_start, const int64_t* seg_len,
    yrb_match* matches, size_t cap_m, uint8_t* item_dirty, vmh_result* res, vmh_stats* stats)
{
  return scan_common(image, data, size, 0, 0, matches, cap_m, nullptr, 0, res, stats, n_seg, seg_start, seg_len, item_dirty);
}

static int scan_common(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, yrb_match* matches, size_t cap_m,
    yrb_match* unconfirmed, size_t cap_u, vmh_result* res, vmh_stats* stats, int n_seg, const int64_t* seg_start,
    const int64_t* seg_len, uint8_t* item_dirty, const yrb_seed* seeds, uint32_t n_seeds, uint32_t* too_many_out)
{
  const uint32_t* skip = g_skip_strings;   /* one scan's worth */
  const uint32_t n_skip = g_n_skip_strings;
  g_skip_strings = nullptr;
  g_n_skip_strings = 0;
  std::vector<Raw> raw;
  std::vector<unsigned long long> str_err;
  memset(res, 0, sizeof(*res));
  res->error_string = YRB_NONE;
  if (run_verify(image, data, size, base, scan_flags, raw, str_err, stats, n_seg, seg_start, seg_len, item_dirty) < 0)
    return -1;

  const yrb_image_header* h = (const yrb_image_header*) image;
  DevRules R;
  memset(&R, 0, sizeof(R));
  R.T = yrb_img_trans(image);
  R.M = yrb_img_match(image);
  R.acm = yrb_img_acm(image);
  R.S = yrb_img_strings(image);
  R.lit = yrb_img_lit(image);
  R.code = yrb_img_code(image);
  R.lower = yrb_img_lower(image);
  R.alter = yrb_img_alter(image);
  R.tables_size = h->tables_size;
  R.num_acm = h->num_acm;
  R.num_strings = h->num_strings;
  R.stack_words = (uint16_t) (h->max_re_stack + 1);
  R.key_seq_bits = (uint16_t) yrb_key_seq_bits(yrb_longest_match_chain(image));
  if (R.key_seq_bits == 0) return -1;   /* api.cu rules_init refuses such a ruleset */

  /* api.cu seeds_to_raw_kernel: the nodes earlier blocks left behind re-enter as records that sort first */
  std::vector<int64_t> seed_base(n_seeds ? n_seeds : 1);
  for (uint32_t k = 0; k < n_seeds; k++)
  {
    const yrb_seed sd = seeds[k];
    Raw r;
    r.key = (uint64_t) (uint32_t) sd.chain_length;
    r.offset = sd.offset;
    r.string_idx = sd.string_idx;
    r.length = sd.match_length;
    r.misc = (uint32_t) sd.xor_key | (k << 8) | RAW_SEED | (sd.kind ? RAW_SEED_MARK : 0u);
    r.group = sd.string_idx < R.num_strings ? R.S[sd.string_idx].chain_head : YRB_NONE;
    raw.push_back(r);
    seed_base[k] = sd.base;
  }

  const uint32_t n = (uint32_t) raw.size();
  const uint32_t ns = R.num_strings ? R.num_strings : 1;
  ScanStatus st;
  memset(&st, 0, sizeof(st));
  std::vector<unsigned long long> str_min_key(ns, ~0ULL);
  std::vector<uint32_t> str_count(ns, 0);
  std::vector<yrb_match> finals(n + 1), unconf(n + 1);
  std::vector<uint32_t> too_many(1024);

  std::vector<unsigned long long> str_cut_key(ns, ~0ULL);
  std::vector<std::pair<unsigned long long, uint32_t>> cuts;   /* (cut key, string), api.cu run_resolve */
  std::vector<uint64_t> k0(n), k1(n), kF(n);
  std::vector<uint32_t> ix0(n), ix1(n), sel(n), starts(n);
  std::vector<uint8_t> flags(n, 0);

  /* ---- launch_resolve (resolve.cu), kernel by kernel */
  auto resolve_pass = [&](bool use_cut) {
    std::fill(flags.begin(), flags.end(), 0);
    std::fill(str_count.begin(), str_count.end(), 0);
    if (scan_flags & YRB_SCAN_FAST_MODE)
    {
      std::fill(str_min_key.begin(), str_min_key.end(), ~0ULL);
      launch_per_thread(n, 256, [&] { fastmode_min_kernel(raw.data(), n, R, str_min_key.data()); });
    }
    launch_per_thread(n, 256, [&] {
      build_key_a_kernel(raw.data(), n, R, scan_flags, str_min_key.data(), use_cut ? str_cut_key.data() : nullptr, k0.data(),
                         ix0.data());
    });

    /* cub::DeviceRadixSort::SortPairs (stable) */
    std::vector<uint32_t> order(n);
    for (uint32_t j = 0; j < n; j++) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return k0[a] < k0[b]; });
    for (uint32_t j = 0; j < n; j++) { k1[j] = k0[order[j]]; ix1[j] = ix0[order[j]]; }

    launch_per_thread(n, 256, [&] { select_plain_kernel(raw.data(), ix1.data(), k1.data(), n, R, str_count.data(), flags.data()); });
    launch_per_thread(R.num_strings, 256, [&] { check_cap_kernel(str_count.data(), R.num_strings, &st); });

    /* cub::DeviceSelect::Flagged over a counting iterator */
    int n_sel = 0;
    for (uint32_t j = 0; j < n; j++) if (flags[j]) sel[n_sel++] = j;
    st.n_sel = n_sel;
    launch_per_thread(n, 256, [&] { write_plain_kernel(raw.data(), ix1.data(), sel.data(), &st, n, finals.data()); });

    if (h->num_chains > 0)
    {
      std::vector<int32_t> nxt(n), prv(n), clen(n), mlen(n), uvn(n), lists((size_t) 8 * ns);
      std::vector<uint8_t> where(n);
      std::vector<Raw> sraw(n);
      ChainLists L;
      L.nxt = nxt.data(); L.prv = prv.data(); L.clen = clen.data(); L.mlen = mlen.data(); L.where = where.data();
      L.uhead = lists.data(); L.utail = lists.data() + ns; L.ucnt = lists.data() + 2 * (size_t) ns;
      L.mhead = lists.data() + 3 * (size_t) ns; L.mtail = lists.data() + 4 * (size_t) ns; L.mcnt = lists.data() + 5 * (size_t) ns;
      L.uvhead = lists.data() + 6 * (size_t) ns; L.prior = lists.data() + 7 * (size_t) ns;
      L.uvn = uvn.data();
      L.first_conf = (scan_flags & YRB_SCAN_FAST_MODE) ? str_min_key.data() : nullptr;
      launch_per_thread(R.num_strings, 256, [&] { chain_init_kernel(L, R.num_strings); });
      st.n_chain_starts = 0;
      launch_per_thread(n, 256, [&] { chain_gather_kernel(raw.data(), ix1.data(), k1.data(), n, sraw.data(), starts.data(), &st); });
      launch_per_thread(st.n_chain_starts, 128, [&] {
        chain_resolve_kernel(sraw.data(), k1.data(), n, R, L, finals.data(), unconf.data(), &st, too_many.data(), 1024, scan_flags,
                             seed_base.data(), (int64_t) base, n_seg, seg_start, seg_len, item_dirty, starts.data());
      });
    }
  };

  if (n > 0)
  {
    resolve_pass(false);

    if (st.over_cap && n_seg == 0)
    {
      /* api.cu run_resolve + resolve.cu resolve_cap_cut: the cut key of every over-cap string, then once more
         without the cut records */
      for (uint32_t t = 0; t < R.num_strings; t++)
      {
        if (str_count[t] <= YRB_MAX_STRING_MATCHES)
          continue;
        unsigned int n_first = 0;
        launch_per_thread(n, 256, [&] { cap_mark_first_kernel(raw.data(), ix1.data(), k1.data(), n, kF.data(), t, &n_first); });
        if (n_first <= YRB_MAX_STRING_MATCHES)
          continue;
        std::sort(kF.begin(), kF.begin() + n_first);
        launch_per_thread(n, 256, [&] { cap_cut_kernel(raw.data(), n, t, kF.data(), &str_cut_key[t]); });
        if (str_cut_key[t] != ~0ULL)
          cuts.push_back(std::make_pair(str_cut_key[t], t));
      }
      st.n_chain_out = 0;
      st.n_unconf = 0;
      st.n_too_many = 0;
      st.over_cap = 0;
      resolve_pass(true);
      st.over_cap = 0;   /* handled */
    }
  }

  /* cut strings fire their TOO_MANY callback in discovery order of the cut, after the chained strings' */
  std::sort(cuts.begin(), cuts.end());
  for (auto& c : cuts)
    if (st.n_too_many < 1024) too_many[st.n_too_many++] = c.second;
  res->first_cut_key = cuts.empty() ? 0 : cuts[0].first;
  res->n_cuts = (uint32_t) cuts.size();

  /* ---- verifier errors (launch_error_select); in batch mode they only mark their item */
  bool any_err = false;
  if (n_seg > 0) std::fill(str_err.begin(), str_err.end(), ~0ULL);
  for (uint32_t s = 0; s < R.num_strings; s++) any_err = any_err || str_err[s] != ~0ULL;
  /* api.cu run_resolve: strings the reference does not verify in this block because of earlier blocks have no say */
  for (uint32_t k = 0; any_err && k < n_skip; k++)
    if (skip[k] < R.num_strings) str_err[skip[k]] = ~0ULL;
  if (any_err)
  {
    launch_per_thread(R.num_strings, 256, [&] {
      error_select_kernel(str_err.data(), n > 0 ? str_min_key.data() : nullptr, cuts.empty() ? nullptr : str_cut_key.data(), R,
                          scan_flags, &st);
    });
    launch_per_thread(R.num_strings, 256, [&] { error_string_kernel(str_err.data(), R.num_strings, &st); });
    if (st.err_key_inv != 0ULL)
    {
      res->error = (int32_t) (~st.err_key_inv & KEY_SEQ_MASK(R.key_seq_bits));
      res->error_string = st.err_string1 ? st.err_string1 - 1 : YRB_NONE;
      res->error_key = ~st.err_key_inv;
    }
  }

  res->n_matches = (uint64_t) st.n_sel + st.n_chain_out;
  res->n_unconfirmed = st.n_unconf;
  res->n_too_many = st.n_too_many;
  res->over_cap = st.over_cap;
  for (uint64_t k = 0; k < res->n_matches && k < cap_m; k++) matches[k] = finals[k];
  for (uint64_t k = 0; k < res->n_unconfirmed && k < cap_u; k++) unconfirmed[k] = unconf[k];
  for (uint32_t k = 0; too_many_out != nullptr && k < res->n_too_many && k < 1024; k++) too_many_out[k] = too_many[k];
  return 0;
}

static long long run_verify(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, std::vector<Raw>& kept_out,
    std::vector<unsigned long long>& str_err_out, vmh_stats* stats, int n_seg, const int64_t* seg_start,
    const int64_t* seg_len, uint8_t* item_dirty)
{
  set_grid_1x1();
  const yrb_image_header* h = (const yrb_image_header*) image;
  if (h->magic != YRB_IMAGE_MAGIC || h->version != YRB_IMAGE_VERSION)
    return -1;

  DevRules R;
  memset(&R, 0, sizeof(R));
  R.T = yrb_img_trans(image);
  R.M = yrb_img_match(image);
  R.acm = yrb_img_acm(image);
  R.S = yrb_img_strings(image);
  R.lit = yrb_img_lit(image);
  R.code = yrb_img_code(image);
  R.lower = yrb_img_lower(image);
  R.alter = yrb_img_alter(image);
  R.tables_size = h->tables_size;
  R.num_acm = h->num_acm;
  R.num_strings = h->num_strings;
  R.stack_words = (uint16_t) (h->max_re_stack + 1);
  R.key_seq_bits = (uint16_t) yrb_key_seq_bits(yrb_longest_match_chain(image));
  if (R.key_seq_bits == 0) return -1;   /* api.cu rules_init refuses such a ruleset */
  if (getenv("VMH_FORCE_KEY_SEQ_BITS"))   /* test_vm_host_cpu.py: the key check below has teeth */
    R.key_seq_bits = (uint16_t) atoi(getenv("VMH_FORCE_KEY_SEQ_BITS"));

  ScanView V;
  memset(&V, 0, sizeof(V));
  V.d = data;
  V.g0 = 0;
  V.g1 = (int64_t) size;
  V.gsize = (int64_t) size;
  V.own_lo = 0;
  V.own_hi = (int64_t) size;
  V.base = base;
  V.scan_flags = scan_flags;
  V.n_seg = n_seg;          /* batch mode (yrb_scan_batch): the block is a concatenation of items */
  V.seg_start = seg_start;
  V.seg_len = seg_len;

  ScanStatus st;
  memset(&st, 0, sizeof(st));

  /* ---- candidates: the walk of scanner.c:72-163 with the product's transition step and emission */
  std::vector<Cand> cand[NUM_CLS];
  for (int pass = 0; pass < 2; pass++)
  {
    CandOut co;
    for (int c = 0; c < NUM_CLS; c++)
    {
      cand[c].resize((size_t) st.n_cand[c] + 1);
      co.cand[c] = cand[c].data();
      co.cap[c] = pass == 0 ? 0 : cand[c].size();
    }
    memset(&st, 0, sizeof(st));
    co.st = &st;

    uint32_t state = 0;
    for (size_t i = 0;; i++)
    {
      const uint32_t first = R.M[state];
      if (first != 0)
        emit_events(R, V, co, (int64_t) i, first);
      if (i >= size)
        break;
      state = ac_step(R.T, state, data[i]);
    }
  }

  unsigned long long n_cand[NUM_CLS];
  for (int c = 0; c < NUM_CLS; c++) n_cand[c] = st.n_cand[c];

  /* The discovery key has to tell any two candidates apart, and in emission order (here: the order they were
     appended in, class by class; on the device nothing else orders them): position first, then the place in the
     state's YR_AC_MATCH chain -- also beyond the 255 entries an 8-bit chain field would hold. */
  {
    std::vector<std::pair<uint64_t, uint64_t>> keys;   /* (key, emission order within its class) */
    for (int c = 0; c < NUM_CLS; c++)
    {
      for (unsigned long long k = 0; k < n_cand[c]; k++)
      {
        if (cand[c][k].chain_pos > KEY_CHAIN_MAX(R.key_seq_bits)) return -3;
        if (k > 0 && cand_key(cand[c][k], R.key_seq_bits) <= cand_key(cand[c][k - 1], R.key_seq_bits)) return -3;
        keys.push_back(std::make_pair(cand_key(cand[c][k], R.key_seq_bits), k));
      }
    }
    std::sort(keys.begin(), keys.end());
    for (size_t k = 1; k < keys.size(); k++)
      if (keys[k].first == keys[k - 1].first) return -3;
  }

  /* ---- verification, tier by tier (api.cu scan_impl) */
  std::vector<unsigned long long> str_err(R.num_strings ? R.num_strings : 1, ~0ULL);
  std::vector<Raw> raw((size_t) 1 << 16);
  std::vector<uint32_t> retry_fast(n_cand[CLS_FAST] + 1), retry_fast3(n_cand[CLS_FAST] + 1), retry_re(n_cand[CLS_RE] + 1);
  std::vector<uint32_t> retry_re0(n_cand[CLS_RE] + 1);
  /* fiber VM: tier 0 (local-memory pool) -> tier 1 -> tier 2 when the repeat stacks fit, else tier 1 -> tier 2 (api.cu) */
  const bool re_t0 = R.stack_words <= RE_LOCAL_STACK && !getenv("YRB_RE_TIER0_OFF");
  std::vector<uint8_t> ws;
  std::vector<unsigned long long> stage_end;
  std::vector<std::unordered_set<uint64_t>> handed_on;

  for (;;)
  {
    st.n_raw = 0;
    st.n_retry[0] = st.n_retry[1] = st.n_retry3 = st.n_retry_re0 = 0;
    st.any_error = 0;
    st.n_chain_raw = 0;
    std::fill(str_err.begin(), str_err.end(), ~0ULL);
    memset(stats, 0, sizeof(*stats));

    RawOut ro;
    ro.raw = raw.data();
    ro.cap = raw.size();
    ro.st = &st;
    ro.str_err = str_err.data();
    ro.item_dirty = item_dirty;
    ro.chain_marks_dirty = (item_dirty != nullptr && g_chain_marks_dirty) ? 1u : 0u;

    /* A candidate that outgrows a tier is verified again from scratch by the next one, which emits
       its records again (resolve.cu treats equal keys as one record); here the records a tier emitted
       for a candidate it then handed on are simply dropped.  stage_end[k] = n_raw after stage k,
       handed_on[k] = candidates stage k queued for the next tier. */
    stage_end.clear();
    handed_on.clear();
    auto close_stage = [&](const std::vector<uint32_t>* q1, unsigned n1, int cls1, const std::vector<uint32_t>* q2, unsigned n2,
                           int cls2) {
      std::unordered_set<uint64_t> set;
      for (unsigned k = 0; q1 && k < n1; k++) set.insert(cand_key(cand[cls1][(*q1)[k]], R.key_seq_bits) >> R.key_seq_bits);
      for (unsigned k = 0; q2 && k < n2; k++) set.insert(cand_key(cand[cls2][(*q2)[k]], R.key_seq_bits) >> R.key_seq_bits);
      stage_end.push_back(st.n_raw);
      handed_on.push_back(set);
    };

    if (n_cand[CLS_LITERAL])
      verify_literal_kernel(R, V, cand[CLS_LITERAL].data(), n_cand[CLS_LITERAL], ro);

    if (n_cand[CLS_FAST])
    {
      ws.assign(fast_ws_bytes_per_thread(1), 0);
      verify_fast_kernel<1>(R, V, cand[CLS_FAST].data(), n_cand[CLS_FAST], ro, ws.data(), fast_ws_bytes_per_thread(1),
                            FAST_POS_CAP_SMALL, nullptr, retry_fast.data());
    }
    if (n_cand[CLS_RE] && re_t0)
      verify_re_kernel<0>(R, V, cand[CLS_RE].data(), n_cand[CLS_RE], ro, nullptr, 32, RE_FIBERS_LOCAL, nullptr, retry_re0.data());
    else if (n_cand[CLS_RE])
    {
      ws.assign(re_ws_bytes_per_thread(R.stack_words, 1), 0);
      verify_re_kernel<1>(R, V, cand[CLS_RE].data(), n_cand[CLS_RE], ro, ws.data(), re_ws_bytes_per_thread(R.stack_words, 1),
                          RE_FIBERS_SMALL, nullptr, retry_re.data());
    }
    stats->retry_fast = st.n_retry[0];
    if (re_t0)
      close_stage(&retry_fast, st.n_retry[0], CLS_FAST, &retry_re0, st.n_retry_re0, CLS_RE);
    else
      close_stage(&retry_fast, st.n_retry[0], CLS_FAST, &retry_re, st.n_retry[1], CLS_RE);

    if (st.n_retry[0])
    {
      ws.assign(fast_ws_bytes_per_thread(2), 0);
      verify_fast_kernel<2>(R, V, cand[CLS_FAST].data(), 0, ro, ws.data(), fast_ws_bytes_per_thread(2), FAST_POS_CAP,
                            retry_fast.data(), retry_fast3.data());
    }
    if (re_t0 && st.n_retry_re0)
    {
      ws.assign(re_ws_bytes_per_thread(R.stack_words, 1), 0);
      verify_re_kernel<1>(R, V, cand[CLS_RE].data(), 0, ro, ws.data(), re_ws_bytes_per_thread(R.stack_words, 1),
                          RE_FIBERS_SMALL, retry_re0.data(), retry_re.data());
    }
    else if (!re_t0 && st.n_retry[1])
    {
      ws.assign(re_ws_bytes_per_thread(R.stack_words, 2), 0);
      verify_re_kernel<2>(R, V, cand[CLS_RE].data(), 0, ro, ws.data(), re_ws_bytes_per_thread(R.stack_words, 2),
                          YRB_RE_MAX_FIBERS, retry_re.data(), nullptr);
    }
    stats->retry_re = st.n_retry[1];
    stats->retry_fast3 = st.n_retry3;
    close_stage(&retry_fast3, st.n_retry3, CLS_FAST, re_t0 ? &retry_re : nullptr, re_t0 ? st.n_retry[1] : 0, CLS_RE);

    if (re_t0 && st.n_retry[1])
    {
      /* stage 3 of the fiber VM (the fast VM's tier 3 below belongs to the same stage) */
      ws.assign(re_ws_bytes_per_thread(R.stack_words, 2), 0);
      verify_re_kernel<2>(R, V, cand[CLS_RE].data(), 0, ro, ws.data(), re_ws_bytes_per_thread(R.stack_words, 2),
                          YRB_RE_MAX_FIBERS, retry_re.data(), nullptr);
    }

    if (st.n_retry3)
    {
      st.n_retry[0] = 0;
      ws.assign(fast_ws_bytes_per_thread(3), 0);
      verify_fast_kernel<3>(R, V, cand[CLS_FAST].data(), 0, ro, ws.data(), fast_ws_bytes_per_thread(3), FAST_POS_CAP_MID,
                            retry_fast3.data(), retry_fast.data());
      stats->retry_fast4 = st.n_retry[0];
      close_stage(&retry_fast, st.n_retry[0], CLS_FAST, nullptr, 0, 0);
      if (st.n_retry[0])
      {
        ws.assign(fast_ws_bytes_per_thread(4), 0);
        verify_fast_kernel<4>(R, V, cand[CLS_FAST].data(), 0, ro, ws.data(), fast_ws_bytes_per_thread(4), FAST_POS_CAP_HUGE,
                              retry_fast.data(), nullptr);
      }
    }

    if (st.n_raw <= raw.size())
      break;
    raw.resize((size_t) st.n_raw + 1024);   /* like scan_impl: verify again with room for every record */
  }

  for (int c = 0; c < NUM_CLS; c++) stats->n_cand[c] = n_cand[c];
  stats->n_raw = st.n_raw;

  unsigned long long first_err = ~0ULL;
  uint32_t err_string = YRB_NONE;
  for (uint32_t s = 0; s < R.num_strings; s++)
  {
    if (str_err[s] < first_err)
    {
      first_err = str_err[s];
      err_string = s;
    }
  }
  if (first_err != ~0ULL)
  {
    stats->error = (int32_t) (first_err & KEY_SEQ_MASK(R.key_seq_bits));
    stats->error_key = first_err & ~(unsigned long long) KEY_SEQ_MASK(R.key_seq_bits);
    stats->error_string = err_string;
  }

  /* drop what a tier emitted for candidates it handed on */
  raw.resize((size_t) st.n_raw);
  kept_out.clear();
  kept_out.reserve(raw.size());
  size_t stage = 0;
  for (size_t k = 0; k < raw.size(); k++)
  {
    while (stage < stage_end.size() && k >= stage_end[stage]) stage++;
    if (stage < handed_on.size() && handed_on[stage].count(raw[k].key >> R.key_seq_bits))
      continue;
    kept_out.push_back(raw[k]);
  }
  str_err_out = str_err;
  return (long long) kept_out.size();
}

} /* extern "C" */
