// resolve.cu -- from Raw records (one per _yr_scan_match_callback call) to the
// final YR_MATCHES lists, on the device:
//   _yr_scan_add_match_to_list                 scan.c:290-368   (a10)
//   _yr_scan_verify_chained_string_match       scan.c:406-632   (a11)
//   _yr_scan_update_match_chain_length         scan.c:257-288
//   FAST_MODE / SINGLE_MATCH guard             scan.c:1021-1023 (a4)
//
// The reference applies these one candidate at a time in discovery order.  Here
// every record carries its discovery key, so the same result is computed as:
//   plain strings   ONE radix sort by (string, offset); of each (string, offset)
//                   group the record with the smallest discovery key survives, or
//                   the one with the largest for GREEDY_REGEXP (replace_if_exists);
//                   ordered compaction keeps the lists in YR_MATCHES order
//   chained strings sorted by (chain, input position) by the same sort, ties put
//                   in key order locally, then replayed by ONE thread per chain
//                   through a verbatim restatement of the list logic (chains are
//                   independent of each other; quirk 10 of SURVEY.md 8(a) rules
//                   out a relational formulation)
//   1M cap          detected per string; see cap_* below
#include <cub/cub.cuh>

#include "resolve.cuh"
#include "resolve_kernels.cuh"

// ------------------------------------------------------------------- host side

void launch_error_select(ResolveWork& w, unsigned long long* str_err, bool have_min_key, cudaStream_t s)
{
  error_select_kernel<<<(w.R.num_strings + 255) / 256, 256, 0, s>>>(
      str_err, have_min_key ? w.str_min_key : nullptr, w.use_cut ? w.str_cut_key : nullptr, w.R, w.scan_flags, w.st);
  error_string_kernel<<<(w.R.num_strings + 255) / 256, 256, 0, s>>>(str_err, w.R.num_strings, w.st);
}


size_t resolve_cub_temp_bytes(uint32_t n)
{
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (uint64_t*) nullptr, (uint64_t*) nullptr, (uint32_t*) nullptr,
                                  (uint32_t*) nullptr, (int) n);
  cub::DeviceSelect::Flagged(nullptr, b2, cub::CountingInputIterator<uint32_t>(0), (uint8_t*) nullptr, (uint32_t*) nullptr,
                             (int*) nullptr, (int) n);
  return (b1 > b2 ? b1 : b2) + 256;
}

#define GRID(n) (unsigned) (((n) + 255) / 256), 256

// ---- candidates in AC-match order for the verifiers (engine.cuh, VerifyWs)
__global__ void cand_keys_kernel(const Cand* cands, uint32_t n, uint32_t* keys, uint32_t* vals)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  keys[j] = cands[j].acm_idx;
  vals[j] = j;
}

size_t cand_sort_temp_bytes(uint32_t n)
{
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (uint32_t*) nullptr, (uint32_t*) nullptr, (uint32_t*) nullptr, (uint32_t*) nullptr, (int) n);
  return b + 256;
}

const uint32_t* launch_cand_sort(
    const Cand* cands, uint32_t n, uint32_t num_acm, uint32_t* k_a, uint32_t* k_b, uint32_t* v_a, uint32_t* v_b, void* tmp,
    size_t tmp_bytes, cudaStream_t s)
{
  int bits = 1;
  while (bits < 32 && (1ull << bits) < (unsigned long long) num_acm) bits++;
  cand_keys_kernel<<<GRID(n), 0, s>>>(cands, n, k_a, v_a);
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_a, k_b, v_a, v_b, (int) n, 0, bits, s);
  return v_b;
}

// Everything from Raw records to final records, queued on the stream without any
// host synchronisation; the caller reads ScanStatus (n_sel, n_chain_out, n_unconf,
// over_cap) afterwards.  With w.use_cut the records at or after each string's cut
// key are dropped (1M cap, second pass).
void launch_resolve(ResolveWork& w, cudaStream_t s, unsigned* launches)
{
  const uint32_t n = w.n_raw;

  if (w.scan_flags & YRB_SCAN_FAST_MODE)
  {
    cudaMemsetAsync(w.str_min_key, 0xFF, sizeof(unsigned long long) * w.R.num_strings, s);
    fastmode_min_kernel<<<GRID(n), 0, s>>>(w.raw, n, w.R, w.str_min_key);
    (*launches)++;
  }

  build_key_a_kernel<<<GRID(n), 0, s>>>(w.raw, n, w.R, w.scan_flags, w.str_min_key, w.use_cut ? w.str_cut_key : nullptr, w.k0, w.ix0);
  size_t b = w.cub_tmp_bytes;
  cub::DeviceRadixSort::SortPairs(w.cub_tmp, b, w.k0, w.k1, w.ix0, w.ix1, (int) n, 0, 64, s); // -> k1, ix1

  // From here on two independent strands work on the sorted records: the plain strings' selection and
  // write-out, and the chain replay.  With a side stream they run next to each other -- the replay (one warp
  // per chain, latency bound) only has to wait for n_sel, the slot its confirmed lists start at.
  const bool fork = w.has_chains && w.side != nullptr;
  cudaStream_t cs = fork ? w.side->aux[0] : s;
  if (fork)
  {
    cudaEventRecord(w.side->ev[3], s);
    cudaStreamWaitEvent(cs, w.side->ev[3], 0);
  }

  ChainLists L;
  if (w.has_chains)   // the ruleset has chained strings AND this block produced records for them (or carries seeds)
  {
    L.nxt = w.c_nxt; L.prv = w.c_prv; L.clen = w.c_clen; L.mlen = w.c_mlen; L.where = w.c_where;
    L.uhead = w.s_lists; L.utail = w.s_lists + w.R.num_strings; L.ucnt = w.s_lists + 2 * (size_t) w.R.num_strings;
    L.mhead = w.s_lists + 3 * (size_t) w.R.num_strings; L.mtail = w.s_lists + 4 * (size_t) w.R.num_strings;
    L.mcnt = w.s_lists + 5 * (size_t) w.R.num_strings;
    L.uvhead = w.s_lists + 6 * (size_t) w.R.num_strings;
    L.prior = w.s_lists + 7 * (size_t) w.R.num_strings;
    L.uvn = w.c_uvn;
    L.first_conf = (w.scan_flags & YRB_SCAN_FAST_MODE) ? w.str_min_key : nullptr;
    chain_init_kernel<<<GRID(w.R.num_strings), 0, cs>>>(L, w.R.num_strings);
    cudaMemsetAsync(&w.st->n_chain_starts, 0, sizeof(unsigned int), cs);
    chain_gather_kernel<<<GRID(n), 0, cs>>>(w.raw, w.ix1, w.k1, n, w.sraw, w.c_scratch_u32, w.st);
  }

  cudaMemsetAsync(w.str_count, 0, sizeof(uint32_t) * w.R.num_strings, s);
  select_plain_kernel<<<GRID(n), 0, s>>>(w.raw, w.ix1, w.k1, n, w.R, w.str_count, w.flags);
  check_cap_kernel<<<GRID(w.R.num_strings), 0, s>>>(w.str_count, w.R.num_strings, w.st);
  b = w.cub_tmp_bytes;
  cub::DeviceSelect::Flagged(w.cub_tmp, b, cub::CountingInputIterator<uint32_t>(0), w.flags, w.ix0, &w.st->n_sel, (int) n, s);
  if (fork)
  {
    cudaEventRecord(w.side->ev[4], s);
    cudaStreamWaitEvent(cs, w.side->ev[4], 0);
  }
  write_plain_kernel<<<GRID(n), 0, s>>>(w.raw, w.ix1, w.ix0, w.st, n, w.finals);
  if (w.finals_host != nullptr && w.n_plain_max > 0)
    cudaMemcpyAsync(w.finals_host, w.finals, (size_t) w.n_plain_max * sizeof(yrb_match), cudaMemcpyDeviceToHost, s);
  (*launches) += 4 + 10 + 3;

  if (w.has_chains)
  {
    // one warp per chain present in the records (at most one chain per string)
    const uint32_t max_chains = w.R.num_strings < n ? w.R.num_strings : n;
    chain_resolve_kernel<<<(max_chains * 32 + 127) / 128, 128, 0, cs>>>(
        w.sraw, w.k1, n, w.R, L, w.finals, w.unconf, w.st, w.too_many, w.too_many_cap, w.scan_flags, w.seed_base, w.cur_base,
        w.n_seg, w.seg_start, w.seg_len, w.item_dirty, w.c_scratch_u32);
    (*launches) += 3;
    if (fork)
    {
      cudaEventRecord(w.side->ev[5], cs);
      cudaStreamWaitEvent(s, w.side->ev[5], 0);
    }
  }
}

// Computes the cut key of one over-cap string (see above).  Synchronises.
int resolve_cap_cut(ResolveWork& w, uint32_t target, cudaStream_t s, unsigned long long* cut_out)
{
  const uint32_t n = w.n_raw;
  unsigned int* d_n = reinterpret_cast<unsigned int*>(w.str_min_key); // scratch, re-initialised by fast mode anyway
  cudaMemsetAsync(d_n, 0, sizeof(unsigned int), s);
  // first-occurrence keys of the target string (k1/ix1 hold the (string, offset) order)
  cap_mark_first_kernel<<<GRID(n), 0, s>>>(w.raw, w.ix1, w.k1, n, w.kF, target, d_n);
  unsigned int n_first = 0;
  cudaMemcpyAsync(&n_first, d_n, sizeof(n_first), cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) return YRB_ERROR_CUDA;
  if (n_first <= YRB_MAX_STRING_MATCHES)
  {
    *cut_out = ~0ULL;
    return YRB_SUCCESS;
  }
  iota_kernel<<<GRID(n_first), 0, s>>>(w.ix0, n_first);
  size_t b = w.cub_tmp_bytes;
  cub::DeviceRadixSort::SortPairs(w.cub_tmp, b, w.kF, w.k0, w.ix0, w.c_scratch_u32, (int) n_first, 0, 64, s);
  unsigned long long* d_cut = w.str_cut_key + target;
  cap_cut_kernel<<<GRID(n), 0, s>>>(w.raw, n, target, w.k0, d_cut);
  cudaMemcpyAsync(cut_out, d_cut, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) return YRB_ERROR_CUDA;
  return YRB_SUCCESS;
}
