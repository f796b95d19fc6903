// emit.cuh -- turning an Aho-Corasick hit (position i, state) into verifier
// candidates: the `match_table[state] != 0` branch of scanner.c:98-122 plus the
// data-independent guards of yr_scan_verify_match (scan.c:1015,1025-1027).
#pragma once

#include "engine.cuh"

// One candidate per YR_AC_MATCH of the state's chain with backtrack <= i.
// chain_pos numbers the chain entries so that (i, chain_pos) reproduces the
// order in which the reference would have called yr_scan_verify_match.
// AGGREGATE: the lanes of a warp that append to the same class at the same time share one atomicAdd
// (engine.cuh, warp_append_slot) -- what the walk kernel wants, where every lane walks its own chunk and dense
// data makes many of them emit at once.  The filter kernel reaches this point with one lane in a thousand
// (0.15 events per 2 KB warp-tile) and pays for the collectives' registers in its hot loop (53 instead of 48,
// -2.3 % scanned bytes/s): it appends with a plain atomicAdd.
template <bool AGGREGATE = true>
static __device__ __forceinline__ void emit_events(
    const DevRules& R, const ScanView& V, const CandOut& out, int64_t i, uint32_t first /* M[state], 1-based */)
{
  uint32_t k = first;
  uint32_t chain_pos = 0;
  int64_t rel = i;          // bytes consumed since the start of the scanned unit
  int64_t size = V.gsize;

  if (V.n_seg > 0)
  {
    // Batch mode: the state was reached over the concatenation, so it may be deeper than
    // the item-local one; every entry it adds in front of the item-local chain has
    // backtrack >= its depth > rel and is rejected by the guard below.
    const int sg = seg_find(V, i - 1);
    if (sg < 0) return;
    rel = i - __ldg(&V.seg_start[sg]);
    size = __ldg(&V.seg_len[sg]);
    if (rel > size) return;   // byte i-1 lies in the gap behind item sg
  }

  while (k != 0)
  {
    const yrb_acm m = R.acm[k - 1];

    if ((int64_t) m.backtrack <= rel)
    {
      int64_t offset = rel - (int64_t) m.backtrack;

      // scan.c:1015  `data_size - offset <= 0` (size_t arithmetic: only == 0)
      if (offset != size)
      {
        const uint32_t sflags = R.S[m.string_idx].flags;
        bool ok = true;

        // scan.c:1025-1027 FIXED_OFFSET guard
        if (sflags & YRB_SF_FIXED_OFFSET)
          ok = (R.S[m.string_idx].fixed_offset == (int64_t) (V.base + (uint64_t) offset));

        if (ok)
        {
          int cls = (sflags & YRB_SF_LITERAL) ? CLS_LITERAL : ((sflags & YRB_SF_FAST_REGEXP) ? CLS_FAST : CLS_RE);
          unsigned long long slot = AGGREGATE ? warp_append_slot(&out.st->n_cand[cls]) : atomicAdd(&out.st->n_cand[cls], 1ULL);

          if (slot < out.cap[cls])
          {
            Cand c;
            c.key = (uint64_t) i << KEY_I_SHIFT;
            c.acm_idx = k - 1;
            c.chain_pos = chain_pos;   // shifted into the key by the verifier (cand_key), where the ruleset's split is at hand
            out.cand[cls][slot] = c;
          }
        }
      }
    }

    chain_pos++;
    k = m.next;
  }
}

// One transition of the interleaved table (scanner.c:124-141,
// ahocorasick.h:39-51): slot = target << 9 | (byte + 1); slot 0 of a state is
// its failure link.
static __device__ __forceinline__ uint32_t ac_step(const uint32_t* __restrict__ T, uint32_t state, uint32_t byte)
{
  const uint32_t index = byte + 1;
  uint32_t t = __ldg(&T[state + index]);

  while ((t & 0x1FF) != index)
  {
    if (state != 0)
    {
      state = __ldg(&T[state]) >> 9;
      t = __ldg(&T[state + index]);
    }
    else
    {
      t = 0;
      break;
    }
  }

  return t >> 9;
}
