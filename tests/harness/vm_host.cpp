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
 * tests/test_vm_host_cpu.py compares with the oracle's match-callback trace.  The resolve stage and
 * the scan kernels use warp collectives and CUB and are covered by the GPU tests only.
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

/* Returns the number of distinct records (may exceed cap: call again with a larger buffer), or -1 on a bad image. */
long long vmh_run(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, vmh_raw* out, size_t cap,
    vmh_stats* stats)
{
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
  R.stack_words = h->max_re_stack + 1;

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

  /* ---- verification, tier by tier (api.cu scan_impl) */
  std::vector<unsigned long long> str_err(R.num_strings ? R.num_strings : 1, ~0ULL);
  std::vector<Raw> raw((size_t) 1 << 16);
  std::vector<uint32_t> retry_fast(n_cand[CLS_FAST] + 1), retry_fast3(n_cand[CLS_FAST] + 1), retry_re(n_cand[CLS_RE] + 1);
  std::vector<uint8_t> ws;
  std::vector<unsigned long long> stage_end;
  std::vector<std::unordered_set<uint64_t>> handed_on;

  for (;;)
  {
    st.n_raw = 0;
    st.n_retry[0] = st.n_retry[1] = st.n_retry3 = 0;
    st.any_error = 0;
    st.n_chain_raw = 0;
    std::fill(str_err.begin(), str_err.end(), ~0ULL);
    memset(stats, 0, sizeof(*stats));

    RawOut ro;
    ro.raw = raw.data();
    ro.cap = raw.size();
    ro.st = &st;
    ro.str_err = str_err.data();
    ro.item_dirty = nullptr;

    /* A candidate that outgrows a tier is verified again from scratch by the next one, which emits
       its records again (resolve.cu treats equal keys as one record); here the records a tier emitted
       for a candidate it then handed on are simply dropped.  stage_end[k] = n_raw after stage k,
       handed_on[k] = candidates stage k queued for the next tier. */
    stage_end.clear();
    handed_on.clear();
    auto close_stage = [&](const std::vector<uint32_t>* q1, unsigned n1, int cls1, const std::vector<uint32_t>* q2, unsigned n2,
                           int cls2) {
      std::unordered_set<uint64_t> set;
      for (unsigned k = 0; q1 && k < n1; k++) set.insert(cand[cls1][(*q1)[k]].key >> KEY_SEQ_BITS);
      for (unsigned k = 0; q2 && k < n2; k++) set.insert(cand[cls2][(*q2)[k]].key >> KEY_SEQ_BITS);
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
    if (n_cand[CLS_RE])
    {
      ws.assign(re_ws_bytes_per_thread(R.stack_words, 1), 0);
      verify_re_kernel<false>(R, V, cand[CLS_RE].data(), n_cand[CLS_RE], ro, ws.data(), re_ws_bytes_per_thread(R.stack_words, 1),
                              RE_FIBERS_SMALL, retry_re.data());
    }
    stats->retry_fast = st.n_retry[0];
    stats->retry_re = st.n_retry[1];
    close_stage(&retry_fast, st.n_retry[0], CLS_FAST, &retry_re, st.n_retry[1], CLS_RE);

    if (st.n_retry[0])
    {
      ws.assign(fast_ws_bytes_per_thread(2), 0);
      verify_fast_kernel<2>(R, V, cand[CLS_FAST].data(), 0, ro, ws.data(), fast_ws_bytes_per_thread(2), FAST_POS_CAP,
                            retry_fast.data(), retry_fast3.data());
    }
    if (st.n_retry[1])
    {
      ws.assign(re_ws_bytes_per_thread(R.stack_words, 2), 0);
      verify_re_kernel<true>(R, V, cand[CLS_RE].data(), 0, ro, ws.data(), re_ws_bytes_per_thread(R.stack_words, 2),
                             YRB_RE_MAX_FIBERS, retry_re.data());
    }
    stats->retry_fast3 = st.n_retry3;
    close_stage(&retry_fast3, st.n_retry3, CLS_FAST, nullptr, 0, 0);

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
    stats->error = (int32_t) (first_err & KEY_SEQ_MAX);
    stats->error_key = first_err & ~(unsigned long long) KEY_SEQ_MAX;
    stats->error_string = err_string;
  }

  /* drop what a tier emitted for candidates it handed on, then discovery order */
  raw.resize((size_t) st.n_raw);
  std::vector<Raw> kept;
  kept.reserve(raw.size());
  size_t stage = 0;
  for (size_t k = 0; k < raw.size(); k++)
  {
    while (stage < stage_end.size() && k >= stage_end[stage]) stage++;
    if (stage < handed_on.size() && handed_on[stage].count(raw[k].key >> KEY_SEQ_BITS))
      continue;
    kept.push_back(raw[k]);
  }
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

} /* extern "C" */
