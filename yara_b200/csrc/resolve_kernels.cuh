// resolve_kernels.cuh -- every kernel and device function of the resolve stage (see resolve.cu
// for the overview).  Kept apart from the host side (CUB sort/select, launch order) so that
// tests/harness/resolve_host.cpp can compile the same device code for the CPU and replay it
// against the oracle (tests/harness/cuda_host_shim.h).
#pragma once

#include "resolve.cuh"

// Lanes that cooperate on one chain in chain_resolve_kernel: a warp on the device, a single lane
// in the CPU replay of the tests (where the collectives degenerate to identities).
#ifndef YRB_WARP
#define YRB_WARP 32
#endif

#define OFFSET_BITS 40
#define FLOOD_RUN 16   // chain replay: consecutive no-op callbacks after which the warp looks ahead
#define CHAIN_ROUND 96 // chain replay: records lane 0 replays before the warp meets again (the other lanes fetch that far ahead)
#define A_CHAIN_BIT (1ULL << 63)
#define A_DROPPED   (~0ULL)

// ----------------------------------------------------------------- key building

__global__ void fastmode_min_kernel(const Raw* raw, uint32_t n, DevRules R, unsigned long long* str_min_key)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const Raw r = raw[j];
  // plain strings: matches[string] becomes non-empty at the string's first matching candidate.  Chained
  // strings: only a chain's HEAD ever gets entries in matches[], and it gets them when a tail completes a
  // chain -- chain_resolve_kernel notes that moment (ChainLists::first_conf); a callback of a chain part is
  // not a match of anything yet.
  if ((R.S[r.string_idx].flags & YRB_SF_SINGLE_MATCH) && r.group == YRB_NONE)
    atomicMin(&str_min_key[r.string_idx], (unsigned long long) (r.key >> R.key_seq_bits));
}

// The sort key: (string, offset) for plain strings; (chain, input position of the
// candidate) for chain parts -- records of one chain must be replayed in discovery
// order, and the few records that share a position are ordered locally by the
// chain kernel.
__global__ void build_key_a_kernel(
    const Raw* raw, uint32_t n, DevRules R, int scan_flags, const unsigned long long* str_min_key,
    const unsigned long long* str_cut_key, uint64_t* kA, uint32_t* ix)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;

  const Raw r = raw[j];
  const uint32_t sflags = R.S[r.string_idx].flags;
  uint64_t a;

  if (r.group != YRB_NONE) // seeds (earlier blocks) come before everything discovered in this block
    a = A_CHAIN_BIT | ((uint64_t) (r.group & 0x7FFFFF) << OFFSET_BITS) |
        ((r.misc & RAW_SEED) ? 0 : (r.key >> KEY_I_SHIFT) + 1);
  else
    a = ((uint64_t) r.string_idx << OFFSET_BITS) | (uint64_t) r.offset;

  // scan.c:1021-1023: in fast mode, once a SINGLE_MATCH string has a match, later
  // candidates are not verified; all callbacks of the first matching candidate stay.
  if ((scan_flags & YRB_SCAN_FAST_MODE) && (sflags & YRB_SF_SINGLE_MATCH) && r.group == YRB_NONE &&
      (r.key >> R.key_seq_bits) != str_min_key[r.string_idx])
    a = A_DROPPED;

  // 1M cap: records discovered at or after the string's cut key never reached a list
  if (str_cut_key != nullptr && r.key >= str_cut_key[r.string_idx])
    a = A_DROPPED;

  kA[j] = a;
  ix[j] = j;
}

// --------------------------------------------------------------- plain strings

// Is sorted record j the one of its (string, offset) group that survives?  The
// group is a run of equal keys in kA; its members are in no particular order.
// first discovered = smallest discovery key (ties, which are re-emissions of the
// same callback by the tier-2 verifier, go to the lowest position).
static __device__ __forceinline__ bool group_pick(
    const Raw* raw, const uint32_t* ix, const uint64_t* kA, uint32_t n, uint32_t j, bool want_last)
{
  const uint64_t a = kA[j];
  if ((j == 0 || kA[j - 1] != a) && (j + 1 == n || kA[j + 1] != a))
    return true; // singleton group: the common case

  const uint64_t my = raw[ix[j]].key;
  for (uint32_t k = j; k-- > 0 && kA[k] == a;)
  {
    const uint64_t o = raw[ix[k]].key;
    if (want_last ? (o > my) : (o <= my)) return false;
  }
  for (uint32_t k = j + 1; k < n && kA[k] == a; k++)
  {
    const uint64_t o = raw[ix[k]].key;
    if (want_last ? (o >= my) : (o < my)) return false;
  }
  return true;
}

__global__ void select_plain_kernel(
    const Raw* raw, const uint32_t* ix, const uint64_t* kA, uint32_t n, DevRules R, uint32_t* str_count, uint8_t* flags)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;

  uint8_t f = 0;
  const uint64_t a = kA[j];

  if (!(a & A_CHAIN_BIT)) // plain, not dropped
  {
    const uint32_t sidx = (uint32_t) (a >> OFFSET_BITS);
    // replace_if_exists (scan.c:314-319) rewrites length/data of the node that was
    // inserted first; xor_key of a greedy regexp is always 0, so "last record" is
    // the same thing.
    const bool greedy = (R.S[sidx].flags & YRB_SF_GREEDY_REGEXP) != 0;
    if (group_pick(raw, ix, kA, n, j, greedy))
    {
      f = 1;
      atomicAdd(&str_count[sidx], 1u);
    }
  }

  flags[j] = f;
}

// sel[k] = sorted position of the k-th surviving plain record (ordered compaction)
__global__ void write_plain_kernel(
    const Raw* raw, const uint32_t* ix, const uint32_t* sel, const ScanStatus* st, uint32_t n, yrb_match* out)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || (int) k >= st->n_sel) return;
  const Raw r = raw[ix[sel[k]]];
  yrb_match m;
  m.offset = r.offset;
  m.string_idx = r.string_idx;
  m.match_length = r.length;
  m.xor_key = (uint8_t) (r.misc & 0xFF);
  m.flags = 0;
  m.seed = 0;
  m.chain_length = 0;
  out[k] = m;
}

__global__ void check_cap_kernel(const uint32_t* str_count, uint32_t num_strings, ScanStatus* st)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num_strings) return;
  if (str_count[s] > YRB_MAX_STRING_MATCHES)
    atomicExch(&st->over_cap, 1u);
}

// -------------------------------------------------------------- chained strings

struct ChainLists
{
  int32_t* nxt;   // per sorted record: list links
  int32_t* prv;
  int32_t* clen;  // YR_MATCH.chain_length
  int32_t* mlen;  // YR_MATCH.match_length (rewritten on confirmation)
  uint8_t* where; // 0 = in no list, 1 = unconfirmed_matches[], 2 = matches[]
  int32_t* uhead; // per string: unconfirmed_matches[s]
  int32_t* utail;
  int32_t* ucnt;
  int32_t* mhead; // per string: matches[s] (only chain heads are used)
  int32_t* mtail;
  int32_t* mcnt;
  // FAST_MODE only (else nullptr): per string, the candidate (key >> key_seq_bits) at which matches[string]
  // became non-empty -- the threshold of the scan.c:1021-1023 guard, which error_select_kernel applies to
  // failing candidates.  The chain kernel fills in the entries of chain heads.
  unsigned long long* first_conf = nullptr;
  int32_t* uvn;    // per record: next in its string's not-yet-visited list / in the ready list
  int32_t* uvhead; // per string
  int32_t* prior;  // per string: matches[s] was already non-empty before this block
};

struct ChainCtx
{
  Raw* sraw;      // the chain records gathered into sorted order (sraw[j] = raw[ix[j]])
  const int64_t* seed_base;
  int64_t cur_base;
  DevRules R;
  ChainLists L;
  unsigned int steps[3];
};

static __device__ __forceinline__ int64_t c_off(const ChainCtx& c, int32_t j) { return c.sraw[j].offset; }
// base + offset, what _yr_scan_add_match_to_list orders and dedups by (scan.c:311-326)
static __device__ __forceinline__ int64_t c_abs(const ChainCtx& c, int32_t j)
{
  const Raw& r = c.sraw[j];
  return ((r.misc & RAW_SEED) ? c.seed_base[(r.misc >> 8) & RAW_SEED_IDX_MASK] : c.cur_base) + r.offset;
}

// scan.c:290-368 over index-linked lists; returns 0 ok / 1 duplicate dropped / 2 too many
static __device__ int clist_add(ChainCtx& c, int32_t* head, int32_t* tail, int32_t* cnt, uint32_t s, int32_t j)
{
  if (cnt[s] == YRB_MAX_STRING_MATCHES)
    return 2;

  const int64_t off = c_abs(c, j);
  int32_t ip = tail[s];

  while (ip >= 0)
  {
    c.steps[0]++;
    const int64_t o = c_abs(c, ip);
    if (off == o) return 1;
    if (off > o) break;
    ip = c.L.prv[ip];
  }

  c.L.prv[j] = ip;
  if (ip >= 0)
  {
    c.L.nxt[j] = c.L.nxt[ip];
    c.L.nxt[ip] = j;
  }
  else
  {
    c.L.nxt[j] = head[s];
    head[s] = j;
  }
  cnt[s]++;
  if (c.L.nxt[j] >= 0)
    c.L.prv[c.L.nxt[j]] = j;
  else
    tail[s] = j;
  return 0;
}

// scan.c:370-389
static __device__ void clist_remove(ChainCtx& c, int32_t* head, int32_t* tail, int32_t* cnt, uint32_t s, int32_t j)
{
  const int32_t p = c.L.prv[j], n = c.L.nxt[j];
  if (p >= 0) c.L.nxt[p] = n;
  if (n >= 0) c.L.prv[n] = p;
  if (head[s] == j) head[s] = n;
  if (tail[s] == j) tail[s] = p;
  cnt[s]--;
  c.L.nxt[j] = c.L.prv[j] = -1;
}

// scan.c:257-288 _yr_scan_update_match_chain_length, recursion turned into an
// explicit stack.
//
// The reference re-scans the whole predecessor list on every call and relies on
// `if (match_to_update->chain_length == chain_length) return;` to stop early.  A
// node of a string that sits d links before the chain's tail is only ever
// updated to the value d, so a node that was updated once ("visited") makes every
// later call on it a no-op (including the recursion below it: quirk 10).  Keeping,
// per string, a list of the nodes NOT yet visited and scanning only those visits
// exactly the same nodes in every call, in O(new nodes) instead of O(all nodes)
// -- which matters for one GPU thread replaying a chain whose parts match
// thousands of times.  Head-string nodes that get visited are exactly the ones
// scan.c:552-583 moves to matches[]: they are collected in a "ready" list.
#define CHAIN_MAX_DEPTH 48

// unvisited list of string t: singly linked through L.uvn, head in L.uvhead[t]; nodes that
// left unconfirmed_matches[] in the meantime (where != 1) are unlinked lazily
static __device__ void update_chain_length(
    ChainCtx& c, uint32_t sidx, int32_t node, int chain_length, int32_t* ready_head)
{
  struct Frame { uint32_t s; int32_t upd; int32_t prev; int32_t cur; int len; };
  Frame fr[CHAIN_MAX_DEPTH];
  int depth = 0;

  uint32_t s = sidx;
  int32_t upd = node;
  int len = chain_length;

  for (;;)
  {
    // "call" update(s, upd, len): upd is an unvisited node of string s
    c.L.clen[upd] = len;
    const yrb_string st = c.R.S[s];
    if (st.chained_to == YRB_NONE)
    {
      c.L.uvn[upd] = *ready_head; // head string: whole chain confirmed up to here
      *ready_head = upd;
    }
    else if (depth < CHAIN_MAX_DEPTH)
    {
      fr[depth].s = s; fr[depth].upd = upd; fr[depth].len = len;
      fr[depth].prev = -1;
      fr[depth].cur = c.L.uvhead[st.chained_to];
      depth++;
    }

    // resume the innermost loop that still has unvisited predecessors to look at
    bool entered = false;
    while (depth > 0 && !entered)
    {
      Frame& f = fr[depth - 1];
      const yrb_string fs = c.R.S[f.s];
      const uint32_t t = fs.chained_to;
      const int64_t upd_off = c_off(c, f.upd);

      while (f.cur >= 0)
      {
        c.steps[2]++;
        const int32_t k = f.cur;
        const int32_t nk = c.L.uvn[k];
        bool unlink = false, visit = false;

        if (c.L.where[k] != 1)
        {
          unlink = true; // pruned or confirmed meanwhile
        }
        else
        {
          const int64_t ending = c_off(c, k) + c.L.mlen[k];
          if (ending + fs.chain_gap_max >= upd_off && ending + fs.chain_gap_min <= upd_off)
            unlink = visit = true;
        }

        if (unlink)
        {
          if (f.prev >= 0) c.L.uvn[f.prev] = nk; else c.L.uvhead[t] = nk;
        }
        else
        {
          f.prev = k;
        }
        f.cur = nk;

        if (visit)
        {
          s = t;
          upd = k;
          len = f.len + 1;
          entered = true;
          break;
        }
      }

      if (!entered)
        depth--;
    }

    if (!entered)
      return;
  }
}

// scan.c:406-632.  Returns 0, or 2 when a list hit the 1M cap.
static __device__ int verify_chained(ChainCtx& c, int32_t j)
{
  const Raw r = c.sraw[j];
  const uint32_t sidx = r.string_idx;
  const yrb_string ms = c.R.S[sidx];
  const int64_t match_offset = r.offset;
  const int32_t match_length = r.length;
  bool add_match = false;

  if (ms.chained_to == YRB_NONE)
  {
    add_match = true;
  }
  else
  {
    const int64_t lowest_offset = c.L.uhead[sidx] >= 0 ? c_off(c, c.L.uhead[sidx]) : match_offset;
    int32_t k = c.L.uhead[ms.chained_to];

    while (k >= 0)
    {
      c.steps[1]++;
      const int32_t next = c.L.nxt[k];
      const int64_t ending = c_off(c, k) + c.L.mlen[k];

      if (ending + ms.chain_gap_max < lowest_offset)
      {
        clist_remove(c, c.L.uhead, c.L.utail, c.L.ucnt, ms.chained_to, k);
        c.L.where[k] = 0;
      }
      else if (ending + ms.chain_gap_max >= match_offset && ending + ms.chain_gap_min <= match_offset)
      {
        add_match = true;
        break;
      }
      k = next;
    }
  }

  if (!add_match)
    return 0;

  if (ms.flags & YRB_SF_CHAIN_TAIL)
  {
    const uint32_t hs = ms.chain_head; // the head string of this chain (flattened at rules upload)

    // scan.c:519-534: every predecessor in range gets chain_length 1 (and recursively
    // 2, 3 ... up the chain); only not-yet-visited nodes can change anything
    int32_t ready = -1;
    {
      const uint32_t t = ms.chained_to;
      int32_t prev = -1, k = c.L.uvhead[t];
      while (k >= 0)
      {
        c.steps[2]++;
        const int32_t nk = c.L.uvn[k];
        bool unlink = false, visit = false;
        if (c.L.where[k] != 1)
        {
          unlink = true;
        }
        else
        {
          const int64_t ending = c_off(c, k) + c.L.mlen[k];
          if (ending + ms.chain_gap_max >= match_offset && ending + ms.chain_gap_min <= match_offset)
            unlink = visit = true;
        }
        if (unlink)
        {
          if (prev >= 0) c.L.uvn[prev] = nk; else c.L.uvhead[t] = nk;
        }
        else
        {
          prev = k;
        }
        if (visit)
          update_chain_length(c, t, k, 1, &ready);
        k = nk;
      }
    }

    // scan.c:546-583: head matches whose whole chain is confirmed move to matches[]
    for (int32_t k = ready; k >= 0;)
    {
      const int32_t next = c.L.uvn[k];
      if (c.L.where[k] == 1)
      {
        clist_remove(c, c.L.uhead, c.L.utail, c.L.ucnt, hs, k);
        c.L.mlen[k] = (int32_t) (match_offset - c_off(c, k) + match_length);
        const int rc = clist_add(c, c.L.mhead, c.L.mtail, c.L.mcnt, hs, k);
        c.L.where[k] = (rc == 0) ? 2 : 0;
        if (rc == 2) return 2;
      }
      k = next;
    }
  }
  else
  {
    c.L.clen[j] = 0;
    c.L.mlen[j] = match_length;
    const int rc = clist_add(c, c.L.uhead, c.L.utail, c.L.ucnt, sidx, j);
    c.L.where[j] = (rc == 0) ? 1 : 0;
    if (rc == 0)
    {
      c.L.uvn[j] = c.L.uvhead[sidx]; // new nodes start unvisited
      c.L.uvhead[sidx] = j;
    }
    if (rc == 2) return 2;
  }

  return 0;
}

__global__ void chain_init_kernel(ChainLists L, uint32_t num_strings)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num_strings) return;
  L.uhead[s] = L.utail[s] = L.mhead[s] = L.mtail[s] = L.uvhead[s] = -1;
  L.prior[s] = 0;
  L.ucnt[s] = L.mcnt[s] = 0;
}

// One thread replays a whole chain, so the records it walks are first gathered into
// sorted order by all threads: sequential reads instead of two dependent random
// gathers per record and per list step.
__global__ void chain_gather_kernel(
    const Raw* raw, const uint32_t* ix, const uint64_t* kA, uint32_t n, Raw* sraw, uint32_t* starts, ScanStatus* st)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint64_t a = kA[j];
  if ((a & A_CHAIN_BIT) && a != A_DROPPED)
  {
    sraw[j] = raw[ix[j]];
    // first record of its chain: one warp of chain_resolve_kernel will replay it
    if (j == 0 || (kA[j - 1] >> OFFSET_BITS) != (a >> OFFSET_BITS))
      starts[atomicAdd(&st->n_chain_starts, 1u)] = j;
  }
}

// One WARP per chain.  Lane 0 replays the chain exactly like the reference would (the
// list logic is order dependent); the other lanes look ahead.  A callback of a non-head
// part whose predecessor's unconfirmed list is empty does nothing at all in
// _yr_scan_verify_chained_string_match (scan.c:452-500: the loop over that list has no
// iterations, add_match stays false) -- and on unlucky data such callbacks come by the
// tens of millions (a short tail part after an unbounded jump matches everywhere, the
// head almost never).  So before every step the warp classifies the next 32 records
// against the CURRENT list state and skips the longest run of such no-ops that ends on
// a group boundary (records sharing an input position are ordered by lane 0 first, so a
// group is skipped only as a whole); everything else goes through lane 0 one group at a
// time.  The state only changes in lane 0's steps, so the classification is exact.
// (capping the registers at 40 for 48 resident warps per SM -- the chains of the bench step need two waves at 80 --
// was measured: 64 registers changes nothing (0.50 ms), at 40 the spills cost more than the second wave (0.55 ms))
__global__ void __launch_bounds__(128)
chain_resolve_kernel(
    Raw* sraw, const uint64_t* kA, uint32_t n, DevRules R, ChainLists L, yrb_match* finals,
    yrb_match* unconf, ScanStatus* st, uint32_t* too_many, uint32_t too_many_cap, int scan_flags,
    const int64_t* seed_base, int64_t cur_base, int n_seg, const int64_t* seg_start, const int64_t* seg_len,
    uint8_t* item_dirty, const uint32_t* starts)
{
  const uint32_t FULL = 0xFFFFFFFFu;                                                    // member mask of the collectives
  const uint32_t ALL = YRB_WARP == 32 ? 0xFFFFFFFFu : ((1u << (YRB_WARP & 31)) - 1u);   // ballot result when every lane votes
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / YRB_WARP;
  const uint32_t lane = threadIdx.x % YRB_WARP;
  if (warp >= st->n_chain_starts) return;

  const uint32_t j0 = starts[warp];
  const uint64_t chain = kA[j0] >> OFFSET_BITS;
  const uint64_t POS_MASK = (1ULL << OFFSET_BITS) - 1;

  ChainCtx c;
  c.sraw = sraw;
  c.seed_base = seed_base;
  c.cur_base = cur_base;
  c.R = R;
  c.L = L;
  c.steps[0] = c.steps[1] = c.steps[2] = 0;

  const uint32_t head_string = sraw[j0].group;
  bool conf_noted = false;              // lane 0: first_conf[head_string] is final
  if (lane == 0 && L.first_conf != nullptr && L.prior[head_string])
  {
    L.first_conf[head_string] = 0;      // matches[head] was non-empty before this block: no candidate of the head is verified
    conf_noted = true;
  }
  uint32_t j = j0;                      // uniform across the warp
  uint64_t prev_key = ~0ULL;            // lane 0's replay state from here on
  uint32_t disabled_string = YRB_NONE;  // strings_temp_disabled after a TOO_MANY (at most one tracked per chain)
  uint64_t cur_cand = ~0ULL;
  bool skip_cand = false;

  // Batch mode: every item is a scan of its own (unconfirmed_matches[] and matches[] start
  // empty, scanner.c:178-204), so at the first record past the current item the head's
  // confirmed list is written out and the lists of the strings replayed so far are emptied.
  int cur_item = -1;
  int64_t item_last = -1;   // largest input position i that still belongs to cur_item (uniform copy)
  uint32_t item_j0 = j0;
  uint32_t touched[16];     // strings lane 0 replayed a record of in the current item
  int n_touched = 0;
  bool touched_overflow = false;

  auto write_confirmed = [&]() {
    const int32_t cnt = L.mcnt[head_string];
    if (cnt <= 0) return;
    unsigned long long slot = (unsigned long long) st->n_sel + atomicAdd(&st->n_chain_out, (unsigned long long) cnt);
    for (int32_t k = L.mhead[head_string]; k >= 0; k = L.nxt[k])
    {
      const Raw r = sraw[k];
      yrb_match m;
      m.offset = r.offset;
      m.string_idx = head_string;
      m.match_length = L.mlen[k];
      m.xor_key = (uint8_t) (r.misc & 0xFF);
      m.flags = (r.misc & RAW_SEED) ? YRB_MATCH_FROM_SEED : 0;
      m.seed = (uint16_t) ((r.misc >> 8) & 0xFFFF);
      if (r.misc & RAW_SEED) m.xor_key = (uint8_t) ((r.misc >> 24) & 0x3F);   // seed index bits 16-21
      m.chain_length = L.clen[k];
      finals[slot++] = m;
    }
  };

  auto reset_string = [&](uint32_t sx) {
    L.uhead[sx] = L.utail[sx] = L.mhead[sx] = L.mtail[sx] = L.uvhead[sx] = -1;
    L.ucnt[sx] = L.mcnt[sx] = 0;
  };

  uint32_t sorted_until = j0;   // records in [j, sorted_until) are in replay order (uniform)
  bool flood = false;           // lane 0 ran into a run of no-op callbacks: look ahead with the whole warp (uniform)

  for (;;)
  {
    if (flood)
    {
      // ---- look-ahead (all lanes): first record at or after j that needs lane 0.  Inside the part
      // that is already in replay order every record stands for itself; beyond it, records that share
      // an input position are not ordered yet, so such a group is skipped only as a whole.
      // O(records / 32) warp steps in total.
      uint32_t scan = j, last_gs = j, stop = j;
      bool found = false, chain_done = false, finished = false;
      uint64_t a_carry = 0;
      while (!finished)
      {
        // eight windows of 32 records per round: their dependent loads (record -> string -> list
        // count) are in flight together, the round costs one latency chain instead of eight
        constexpr int LA = 8;
        uint64_t av[LA];
        bool inv[LA], skv[LA];
        uint32_t sidx[LA], misc[LA];
  #pragma unroll
        for (int u = 0; u < LA; u++)
        {
          const uint32_t jj = scan + (uint32_t) u * YRB_WARP + lane;
          av[u] = ~0ULL;
          sidx[u] = 0;
          misc[u] = 0;
          if (jj < n)
          {
            av[u] = kA[jj];
            sidx[u] = sraw[jj].string_idx;   // meaningful only for chain records (checked below)
            misc[u] = sraw[jj].misc;
          }
        }
  #pragma unroll
        for (int u = 0; u < LA; u++)
        {
          inv[u] = (av[u] >> OFFSET_BITS) == chain;
          skv[u] = false;
          if (inv[u])
          {
            const uint32_t pred = R.S[sidx[u]].chained_to;
            skv[u] = !(misc[u] & RAW_SEED) && pred != YRB_NONE && __ldcg(&L.ucnt[pred]) == 0;
            if (n_seg > 0 && (int64_t) (av[u] & POS_MASK) - 1 > item_last)
              skv[u] = false;   // first record of the next item: lane 0 has to close the current one
          }
        }

  #pragma unroll
        for (int u = 0; u < LA; u++)
        {
          if (finished) break;
          const uint32_t jj = scan + lane;
          const uint64_t a = av[u];
          const bool in_chain = inv[u];
          const unsigned vm = __ballot_sync(FULL, in_chain);   // a prefix of the lanes: a chain's records are contiguous
          const uint32_t nvalid = (uint32_t) __popc(vm);
          if (nvalid == 0)
          {
            chain_done = (scan == j);
            stop = scan;
            finished = true;
            break;
          }
          const unsigned ns = vm & ~__ballot_sync(FULL, skv[u]);

          uint64_t a_prev = __shfl_up_sync(FULL, a, 1);
          if (lane == 0) a_prev = (scan == j) ? ~a : a_carry;   // j itself always starts a group
          const unsigned gm = __ballot_sync(FULL, in_chain && (jj < sorted_until || a != a_prev));

          if (ns)
          {
            const uint32_t fl = (uint32_t) (__ffs(ns) - 1);
            const unsigned upto = gm & (fl >= 31 ? FULL : ((2u << fl) - 1u));
            if (upto) last_gs = scan + 31u - (uint32_t) __clz(upto);
            found = true;
            finished = true;
            break;
          }
          if (gm) last_gs = scan + 31u - (uint32_t) __clz(gm);
          a_carry = __shfl_sync(FULL, a, nvalid - 1);
          stop = scan + nvalid;
          if (nvalid < YRB_WARP)
          {
            finished = true;   // the chain ends inside this window, nothing left for lane 0
            break;
          }
          scan += YRB_WARP;
        }
      }
      if (chain_done)
        break;

      const uint32_t skip_to = found ? last_gs : stop;
      if (skip_to > j)
      {
        for (uint32_t k = j + lane; k < skip_to; k += YRB_WARP)
          L.where[k] = 0;   // in no list
        j = skip_to;
        prev_key = ~0ULL;
        cur_cand = ~0ULL;
        skip_cand = false;
        continue;
      }

      if (j >= sorted_until)
      {
        // a new, possibly huge group of records that share the input position of record j: find its
        // end and check its order with all lanes; lane 0 sorts in the rare unordered case
        const uint64_t a_j = kA[j];
        const bool seed_group = (a_j & POS_MASK) == 0;
        uint32_t r_end = j;
        bool ordered = true;
        for (;;)
        {
          const uint32_t k = r_end + lane;
          const bool same = k < n && kA[k] == a_j;
          const unsigned sm2 = __ballot_sync(FULL, same);
          const bool bad = same && k > j && !seed_group && sraw[k - 1].key > sraw[k].key;
          const unsigned cnt = (sm2 == ALL) ? (unsigned) YRB_WARP : (uint32_t) (__ffs(~sm2) - 1);
          const unsigned bm = __ballot_sync(FULL, bad) & (cnt >= 32 ? FULL : ((1u << cnt) - 1u));
          if (bm) ordered = false;
          r_end += cnt;
          if (cnt < YRB_WARP) break;
        }
        if (!ordered && lane == 0)
        {
          for (uint32_t x = j + 1; x < r_end; x++)
          {
            if (sraw[x - 1].key <= sraw[x].key) continue;
            const Raw vx = sraw[x];
            uint32_t y = x;
            while (y > j && sraw[y - 1].key > vx.key) { sraw[y] = sraw[y - 1]; y--; }
            sraw[y] = vx;
          }
          __threadfence_block();
        }
        __syncwarp();
        sorted_until = r_end;
        continue;   // look again: inside the ordered group single records can be skipped
      }
    }

    // ---- lane 0 replays records from j on, group by group, until the chain ends or it has seen
    // FLOOD_RUN no-op callbacks in a row (then the warp looks ahead instead)
    uint32_t j_next = j, sorted_next = sorted_until;
    int flags_next = 0;   // 1: flood, 2: chain finished, 0: lane 0 used up its round
#if YRB_WARP > 1
    // The replay is one dependent chain of loads per record (sort key, record, its string, the predecessor's list
    // count, the list nodes) at L2 latency each.  While lane 0 replays a round of CHAIN_ROUND records the other
    // lanes run ahead of it over the same records and pull what does not depend on list state into L1.
    if (lane != 0)
    {
      for (uint32_t x = j + lane - 1; x < n && x < j + CHAIN_ROUND + 64; x += YRB_WARP - 1)
      {
        if ((kA[x] >> OFFSET_BITS) != chain) break;
        const uint32_t sx = sraw[x].string_idx;
        const uint32_t pred = R.S[sx].chained_to;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(&sraw[x].string_idx));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(&R.S[sx].chain_gap_max));
        if (pred != YRB_NONE)
        {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(&L.ucnt[pred]));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(&L.utail[pred]));
        }
      }
    }
#endif
    if (lane == 0)
    {
      uint32_t jq = j;
      uint32_t su = sorted_until;
      int noop_run = 0;
      for (;;)
      {
        if (jq >= n || (kA[jq] >> OFFSET_BITS) != chain) { flags_next = 2; break; }

        if (jq >= su)
        {
          if (YRB_WARP > 1 && jq - j >= CHAIN_ROUND) break;   // group boundary: meet the warp, which fetches the next round

          // records of this chain that share the input position: put them in discovery-key order.
          // The seeds -- position 0 -- keep the order the caller gave them in, ascending base + offset
          // per string: each then goes to the tail of its list in O(1); their `key` is a chain length.
          uint32_t r_end = jq + 1;
          while (r_end < n && r_end < jq + 64 && kA[r_end] == kA[jq]) r_end++;
          if (r_end == jq + 64 && r_end < n && kA[r_end] == kA[jq])
          {
            flags_next = 1;   // a long group (one candidate can fire thousands of callbacks): the warp
            break;            // finds its end, checks its order and skips inside it
          }
          const bool seed_group = (kA[jq] & POS_MASK) == 0;
          for (uint32_t x = jq + 1; x < r_end && !seed_group; x++)
          {
            if (sraw[x - 1].key <= sraw[x].key) continue; // already in order: the common case
            const Raw vx = sraw[x];
            uint32_t y = x;
            while (y > jq && sraw[y - 1].key > vx.key) { sraw[y] = sraw[y - 1]; y--; }
            sraw[y] = vx;
          }
          su = r_end;
        }

        for (; jq < su; jq++)
        {
          const Raw r = sraw[jq];

          if (n_seg > 0)
          {
            const int64_t i = (int64_t) (kA[jq] & POS_MASK) - 1;
            if (i > item_last)
            {
              if (cur_item >= 0)
              {
                write_confirmed();
                if (touched_overflow)
                  for (uint32_t k = item_j0; k < jq; k++) reset_string(sraw[k].string_idx);
                else
                  for (int t = 0; t < n_touched; t++) reset_string(touched[t]);
              }
              n_touched = 0;
              touched_overflow = false;
              cur_item = seg_find_in(seg_start, n_seg, i - 1);
              item_last = seg_start[cur_item] + seg_len[cur_item];
              item_j0 = jq;
              prev_key = ~0ULL;
              disabled_string = YRB_NONE;
              cur_cand = ~0ULL;
              skip_cand = false;
            }
          }

          if (!(r.misc & RAW_SEED))
          {
            // scan.c:452-500: nothing to look at in the predecessor's list, nothing happens
            const uint32_t pred = R.S[r.string_idx].chained_to;
            if (pred != YRB_NONE && L.ucnt[pred] == 0)
            {
              L.where[jq] = 0;
              if (++noop_run >= FLOOD_RUN) { flags_next = 1; jq++; break; }
              continue;
            }
          }
          noop_run = 0;

          if (n_seg > 0 && !touched_overflow)
          {
            int t = 0;
            while (t < n_touched && touched[t] != r.string_idx) t++;
            if (t == n_touched)
            {
              if (n_touched < 16) touched[n_touched++] = r.string_idx; else touched_overflow = true;
            }
          }

          L.nxt[jq] = L.prv[jq] = -1;
          L.where[jq] = 0;

          if (r.misc & RAW_SEED)
          {
            // state left by earlier blocks of this scan: goes straight back into its list
            if (r.misc & RAW_SEED_MARK)
            {
              L.prior[r.string_idx] = 1;
            }
            else
            {
              L.clen[jq] = (int32_t) r.key;
              L.mlen[jq] = r.length;
              if (clist_add(c, L.uhead, L.utail, L.ucnt, r.string_idx, (int32_t) jq) == 0)
              {
                L.where[jq] = 1;
                if (L.clen[jq] == 0)
                {
                  L.uvn[jq] = L.uvhead[r.string_idx];
                  L.uvhead[r.string_idx] = (int32_t) jq;
                }
              }
            }
            continue;
          }

          if (r.key == prev_key) continue; // re-emission of the same callback by the tier-2 verifier
          prev_key = r.key;

          if (r.string_idx == disabled_string) continue;

          // scan.c:1021-1023: the FAST_MODE / SINGLE_MATCH guard is evaluated once per
          // candidate, before verification, against matches[string->idx] (only ever
          // non-empty for the chain's head string)
          if ((r.key >> R.key_seq_bits) != cur_cand)
          {
            cur_cand = r.key >> R.key_seq_bits;
            skip_cand = (scan_flags & YRB_SCAN_FAST_MODE) && (R.S[r.string_idx].flags & YRB_SF_SINGLE_MATCH) &&
                        (L.mcnt[r.string_idx] > 0 || L.prior[r.string_idx]);
          }
          if (skip_cand) continue;

          if (verify_chained(c, (int32_t) jq) == 2)
          {
            const uint32_t slot = atomicAdd(&st->n_too_many, 1u);
            if (slot < too_many_cap) too_many[slot] = r.string_idx;
            disabled_string = r.string_idx;
            if (n_seg > 0) item_dirty[cur_item] = 1;   // the callback protocol needs the per-item path
          }
          if (!conf_noted && L.first_conf != nullptr && L.mcnt[head_string] > 0)
          {
            // this callback completed the first chain: from the next candidate on the head string is no
            // longer verified in fast mode -- neither its matches (skip_cand above) nor its errors count
            L.first_conf[head_string] = (unsigned long long) (r.key >> R.key_seq_bits);
            conf_noted = true;
          }
        }
        if (flags_next) break;
      }
      j_next = jq;
      sorted_next = su;
      __threadfence_block();
    }
    __syncwarp();
    j = __shfl_sync(FULL, j_next, 0);
    sorted_until = __shfl_sync(FULL, sorted_next, 0);
    item_last = __shfl_sync(FULL, item_last, 0);
    flags_next = __shfl_sync(FULL, flags_next, 0);
    if (flags_next == 2)
      break;
    flood = flags_next == 1;
  }

  if (lane == 0)
  {
    // confirmed matches of the chain's head string, in list order, after the plain strings' records
    write_confirmed();
    atomicAdd(&st->dbg[0], c.steps[0]);
    atomicAdd(&st->dbg[1], c.steps[1]);
    atomicAdd(&st->dbg[2], c.steps[2]);
    __threadfence_block();
  }
  __syncwarp();

  // what is left in unconfirmed_matches[] of the chain's strings, in record order (a batch item is a
  // whole scan: its leftovers are dropped like yr_scanner_scan_mem_blocks does at the end, scanner.c:565)
  if (n_seg == 0)
  {
    for (uint32_t base0 = j0; base0 < j; base0 += 8 * YRB_WARP)
    {
      // eight windows per round; almost all of them hold nothing (one latency chain per round)
      uint8_t wv[8];
      bool any = false;
#pragma unroll
      for (int u = 0; u < 8; u++)
      {
        const uint32_t k = base0 + (uint32_t) u * YRB_WARP + lane;
        wv[u] = k < j ? L.where[k] : (uint8_t) 0;
        any = any || wv[u] == 1;
      }
      if (!__any_sync(FULL, any)) continue;
#pragma unroll
      for (int u = 0; u < 8; u++)
      {
        const uint32_t k = base0 + (uint32_t) u * YRB_WARP + lane;
        const bool keep = wv[u] == 1;
        const unsigned km = __ballot_sync(FULL, keep);
        if (km == 0) continue;
        unsigned long long slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&st->n_unconf, (unsigned long long) __popc(km));
        slot0 = __shfl_sync(FULL, slot0, 0);
        if (keep)
        {
          const Raw r = sraw[k];
          yrb_match m;
          m.offset = r.offset;
          m.string_idx = r.string_idx;
          m.match_length = L.mlen[k];
          m.xor_key = (uint8_t) (r.misc & 0xFF);
          m.flags = YRB_MATCH_UNCONFIRMED | ((r.misc & RAW_SEED) ? YRB_MATCH_FROM_SEED : 0);
          m.seed = (uint16_t) ((r.misc >> 8) & 0xFFFF);
          if (r.misc & RAW_SEED) m.xor_key = (uint8_t) ((r.misc >> 24) & 0x3F);   // seed index bits 16-21
          m.chain_length = L.clen[k];
          unconf[slot0 + __popc(km & ((1u << lane) - 1u))] = m;
        }
      }
    }
  }
}

__global__ void iota_kernel(uint32_t* p, uint32_t n)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) p[j] = j;
}

// ------------------------------------------------------------ verifier errors

// The scan's error = the earliest failing candidate, in discovery order, that the
// reference would actually have verified (see report_error in engine.cuh).
__global__ void error_select_kernel(
    const unsigned long long* str_err, const unsigned long long* str_min_key, const unsigned long long* str_cut_key,
    DevRules R, int scan_flags, ScanStatus* st)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= R.num_strings) return;
  const unsigned long long e = str_err[s];
  if (e == ~0ULL) return;
  const unsigned long long cand = e >> R.key_seq_bits;

  // scan.c:1021-1023: later candidates of a SINGLE_MATCH string that already matched are not verified
  // (for a chain's head string the threshold is the candidate that completed its first chain; the other
  // parts of a chain never have matches[] entries and keep ~0)
  if (str_min_key != nullptr && (scan_flags & YRB_SCAN_FAST_MODE) && (R.S[s].flags & YRB_SF_SINGLE_MATCH) &&
      str_min_key[s] < cand)
    return;
  // scan.c:1018: neither are candidates of a string disabled by the 1M cap
  if (str_cut_key != nullptr && str_cut_key[s] != ~0ULL && (str_cut_key[s] >> R.key_seq_bits) < cand)
    return;

  atomicMax(&st->err_key_inv, ~e);
}

// scan.c:1085-1086: last_error_string = the string of the failing candidate
__global__ void error_string_kernel(const unsigned long long* str_err, uint32_t num_strings, ScanStatus* st)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num_strings) return;
  if (st->err_key_inv != 0ULL && str_err[s] == ~st->err_key_inv)
    st->err_string1 = s + 1;
}

// ---------------------------------------------------------------- 1M cap (a10)
//
// scan.c:303-307 tests `count == YR_MAX_STRING_MATCHES` BEFORE dedup, in
// discovery order.  For a string whose distinct offsets exceed the cap: the list
// count grows exactly at first-occurrence records; the error fires at the first
// record (any, duplicate or not) discovered after the 1,000,000th of them.  That
// record's key is the string's "cut key": it and everything after it never reach
// the list (the callback disables the string, scan.c:1064-1067).

__global__ void cap_mark_first_kernel(
    const Raw* raw, const uint32_t* ix, const uint64_t* kA, uint32_t n, uint64_t* keys_out, uint32_t target,
    unsigned int* n_out)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint64_t a = kA[j];
  if (a & A_CHAIN_BIT) return;
  if ((uint32_t) (a >> OFFSET_BITS) != target) return;
  if (!group_pick(raw, ix, kA, n, j, false)) return; // not the first discovered of its offset
  keys_out[atomicAdd(n_out, 1u)] = raw[ix[j]].key;
}

__global__ void cap_cut_kernel(
    const Raw* raw, uint32_t n, uint32_t target, const uint64_t* sorted_first_keys, unsigned long long* cut_key)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const Raw r = raw[j];
  if (r.string_idx != target || r.group != YRB_NONE) return;
  const uint64_t k1m = sorted_first_keys[YRB_MAX_STRING_MATCHES - 1];
  if (r.key > k1m)
    atomicMin(cut_key, (unsigned long long) r.key);
}

