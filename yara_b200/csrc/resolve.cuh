// resolve.cuh -- work descriptor of the resolve stage (resolve.cu).
#pragma once

#include "engine.cuh"

struct ResolveWork
{
  // inputs
  const Raw* raw;
  uint32_t n_raw;
  DevRules R;
  int scan_flags;
  int has_chains;
  int use_cut;
  const int64_t* seed_base;   // per seed: YR_MATCH.base (device)
  int64_t cur_base;           // base of the block being scanned
  int n_seg;                  // batch mode (ScanView): the chain replay starts afresh at every item
  const int64_t* seg_start;
  const int64_t* seg_len;
  uint8_t* item_dirty;

  // scratch, all sized for n_raw records unless noted
  uint64_t* k0;
  uint64_t* k1;
  uint64_t* kF;
  uint32_t* ix0;
  uint32_t* ix1;
  uint32_t* c_scratch_u32;
  void* cub_tmp;
  size_t cub_tmp_bytes;
  int32_t* c_nxt;
  int32_t* c_prv;
  int32_t* c_clen;
  int32_t* c_mlen;
  int32_t* c_uvn;
  Raw* sraw;
  uint8_t* c_where;
  uint8_t* flags;

  // per string
  unsigned long long* str_min_key; // num_strings
  unsigned long long* str_cut_key; // num_strings, ~0 = no cut
  uint32_t* str_count;             // num_strings
  int32_t* s_lists;                // 8 * num_strings

  // optional: side stream + events (the chain replay runs next to the plain strings' selection and write-out)
  // and where the plain strings' records go on the host: with `finals_host` set and `n_plain_max` = an upper
  // bound of the number of plain records (records that are not chain parts), launch_resolve queues the copy of
  // finals[0, n_plain_max) right behind write_plain_kernel, while the chain replay is still running
  const SideStreams* side;
  yrb_match* finals_host;
  uint32_t n_plain_max;

  // outputs
  yrb_match* finals;
  yrb_match* unconf;
  uint32_t* too_many;
  uint32_t too_many_cap;
  ScanStatus* st;
};

size_t resolve_cub_temp_bytes(uint32_t n);
void launch_resolve(ResolveWork& w, cudaStream_t s, unsigned* launches);
void launch_error_select(ResolveWork& w, unsigned long long* str_err, bool have_min_key, cudaStream_t s);
int resolve_cap_cut(ResolveWork& w, uint32_t target, cudaStream_t s, unsigned long long* cut_out);
