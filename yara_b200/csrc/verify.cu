// verify.cu -- per-candidate verification on the device:
//   yr_scan_verify_match dispatch             scan.c:992-1089
//   _yr_scan_verify_literal_match + compares  scan.c:62-255, 887-990
//   _yr_scan_verify_re_match                  scan.c:778-885
//   yr_re_fast_exec (fast hex VM)             re.c:2150-2477
//   yr_re_exec + fiber machinery (Pike VM)    re.c:1214-2068
//   _yr_scan_match_callback, fullword filter  scan.c:634-694
//
// One thread verifies one candidate (candidates are rare: a handful per MiB of
// input).  Everything order-dependent -- list insertion, dedup, the 1M cap,
// fast mode, chained strings -- is deferred to resolve.cu; a verifier only
// appends Raw records that carry the discovery key.
//
// The two VMs are restated verbatim (list order IS the matching semantics, see
// SURVEY.md 8(a) rows a7/a8 and quirk 1); their linked lists live in a
// per-thread work area in global memory, indexed with 16-bit links.
#include "engine.cuh"

// The fiber VM inlined into its three call sites is 150 KB of SASS and its warps spend a third of their stall
// samples waiting for instructions -- but as real functions (__noinline__: 74 KB) re_exec and fiber_sync get the
// candidate's context and the fiber pool by reference, i.e. in local memory instead of registers, and the class
// took 40 % longer (cfg 5, 1 GiB: 0.55 -> 0.76 ms; profiles/r02_kernel_notes.txt).  Kept inline.
#if defined(__CUDACC__) && defined(YRB_VM_NOINLINE)
#define YRB_NOINLINE __noinline__
#else
#define YRB_NOINLINE
#endif

// ----------------------------------------------------------------- data access

struct Ctx
{
  const DevRules* R;
  const uint8_t* d;   // d[p] = block byte p (pointer pre-biased by -g0)
  int64_t g0, g1, gsize;
  RawOut out;
  uint64_t cand_key;
  uint32_t seq;
  int32_t seg;        // batch mode: the item being verified (d, g0, g1, gsize describe that item alone)
  int64_t bias;       // batch mode: start of the item in the concatenation (added to reported offsets)
};

static __device__ __forceinline__ uint8_t byte_at(Ctx& c, int64_t p)
{
  if (p < c.g0 || p >= c.g1)
  {
    atomicExch(&c.out.st->margin_fault, 1u);
    return 0;
  }
  return __ldg(c.d + p);
}

struct CbArgs
{
  uint32_t sidx;
  int forward_matches;
  int full_word;
  uint32_t xor_key;
  uint32_t group;
};

static __device__ __forceinline__ bool is_alnum(uint8_t c)
{
  return (c >= 0x30 && c <= 0x39) || (c >= 0x41 && c <= 0x5a) || (c >= 0x61 && c <= 0x7a);
}

// scan.c:634-765 up to (not including) the list insertion
static __device__ void match_callback(Ctx& c, int64_t match_pos, int32_t match_length, int re_flags, const CbArgs& a)
{
  match_length += a.forward_matches;

  if (a.full_word)
  {
    if (re_flags & RE_FLAGS_WIDE)
    {
      if (match_pos >= 2 && byte_at(c, match_pos - 1) == 0 && is_alnum(byte_at(c, match_pos - 2)))
        return;
      if (match_pos + match_length + 1 < c.gsize && byte_at(c, match_pos + match_length + 1) == 0 &&
          is_alnum(byte_at(c, match_pos + match_length)))
        return;
    }
    else
    {
      if (match_pos >= 1 && is_alnum(byte_at(c, match_pos - 1)))
        return;
      if (match_pos + match_length < c.gsize && is_alnum(byte_at(c, match_pos + match_length)))
        return;
    }
  }

  const uint32_t seq_max = KEY_SEQ_MASK(c.R->key_seq_bits);   // the largest sequence number the ruleset's key layout holds
  const unsigned long long slot = warp_append_slot(&c.out.st->n_raw);

  if (slot < c.out.cap)
  {
    Raw r;
    r.key = c.cand_key | (uint64_t) (c.seq < seq_max ? c.seq : seq_max);
    r.offset = match_pos + c.bias;
    r.string_idx = a.sidx;
    r.length = match_length;
    r.misc = a.xor_key & 0xFF;
    r.group = a.group;
    c.out.raw[slot] = r;
  }

  if (a.group != YRB_NONE)
  {
    atomicAdd(&c.out.st->n_chain_raw, 1u);
    // the items are the blocks of ONE scan: a chain part's match needs the chain state of the blocks before it
    if (c.out.chain_marks_dirty && c.out.item_dirty) c.out.item_dirty[c.seg] = 1;
  }

  if (c.seq >= seq_max)
    atomicExch(&c.out.st->key_overflow, 1u);   // fail loudly (api.cu) rather than order records by a saturated key

  c.seq++;
}

// -------------------------------------------------------------------- opcodes

enum
{
  OP_ANY = 0xA0, OP_LITERAL = 0xA2, OP_MASKED_LITERAL = 0xA4, OP_CLASS = 0xA5,
  OP_WORD_CHAR = 0xA7, OP_NON_WORD_CHAR = 0xA8, OP_SPACE = 0xA9, OP_NON_SPACE = 0xAA,
  OP_DIGIT = 0xAB, OP_NON_DIGIT = 0xAC, OP_MATCH = 0xAD, OP_NOT_LITERAL = 0xAE,
  OP_MASKED_NOT_LITERAL = 0xAF, OP_MATCH_AT_END = 0xB0, OP_MATCH_AT_START = 0xB1,
  OP_WORD_BOUNDARY = 0xB2, OP_NON_WORD_BOUNDARY = 0xB3, OP_REPEAT_ANY_GREEDY = 0xB4,
  OP_REPEAT_ANY_UNGREEDY = 0xB5, OP_SPLIT_A = 0xC0, OP_SPLIT_B = 0xC1, OP_JUMP = 0xC2,
  OP_REPEAT_START_GREEDY = 0xC3, OP_REPEAT_END_GREEDY = 0xC4,
  OP_REPEAT_START_UNGREEDY = 0xC5, OP_REPEAT_END_UNGREEDY = 0xC6
};

static __device__ __forceinline__ uint32_t code_u8(const Ctx& c, uint32_t off) { return __ldg(c.R->code + off); }
static __device__ __forceinline__ uint32_t code_u16(const Ctx& c, uint32_t off)
{ return code_u8(c, off) | (code_u8(c, off + 1) << 8); }
static __device__ __forceinline__ int32_t code_i16(const Ctx& c, uint32_t off) { return (int16_t) code_u16(c, off); }
static __device__ __forceinline__ int32_t code_i32(const Ctx& c, uint32_t off)
{ return (int32_t) (code_u16(c, off) | (code_u16(c, off + 2) << 16)); }

// ---------------------------------------------------------- fast hex VM (a8)

#define FAST_POS_CAP 4608       // tier 2: more than any forward-mode list can hold (<= 4097 distinct positions)
#define FAST_POS_CAP_SMALL 96   // tier 1: what almost every candidate needs
#define FAST_POS_CAP_MID (1 << 16)  // tier 3: backward-mode lists are not kept duplicate-free by the reference
#define FAST_POS_CAP_HUGE (1 << 20) // tier 4: the few candidates that outgrow even that
                                    // (re.c:2323-2333) and can grow far beyond the window
#define WS_RETRY (-2)           // internal: the small work area overflowed, verify again in tier 2

struct FastPos
{
  int32_t input;   // relative to the exec's input position
  int32_t round;
  int32_t prev, next;
};

struct FastWs
{
  FastPos* v;
  int cap;
};

size_t fast_ws_bytes_per_thread(int tier)
{ return sizeof(FastPos) * (size_t) (tier == 1 ? FAST_POS_CAP_SMALL : tier == 2 ? FAST_POS_CAP : tier == 3 ? FAST_POS_CAP_MID : FAST_POS_CAP_HUGE); }

// re.c:2150-2477.  Returns an ERROR_* code.  `pos` = block position of input_data.
static __device__ int re_fast_exec(
    Ctx& c, FastWs* ws, uint32_t code_off, int64_t pos, int64_t input_forwards_size, int64_t input_backwards_size,
    int flags, const CbArgs* cb, int* matches)
{
  FastPos* P = ws->v;
  int n_used = 0;      // bump allocator
  int free_head = -1;  // recycled entries, linked through .next

  const int input_incr = (flags & RE_FLAGS_BACKWARDS) ? -1 : 1;
  int bytes_matched;
  int max_bytes_matched;

  if (flags & RE_FLAGS_BACKWARDS)
    max_bytes_matched = (int) (input_backwards_size < YRB_RE_SCAN_LIMIT ? input_backwards_size : YRB_RE_SCAN_LIMIT);
  else
    max_bytes_matched = (int) (input_forwards_size < YRB_RE_SCAN_LIMIT ? input_forwards_size : YRB_RE_SCAN_LIMIT);

  uint32_t ip = code_off;
  int round = 0;

  if (matches) *matches = -1;

  // Until the first jump the position list holds ONE entry (re.c:2184-2196 seeds it with the exec's input
  // position and only OP_REPEAT_ANY_UNGREEDY adds to it), and nearly every candidate -- a 4-byte atom that
  // happened to occur in unrelated bytes -- dies on the first opcode behind its atom.  That prefix is run on
  // registers: same opcodes, same bounds, same effects as the list loop below with a one-element list; the
  // entry is written to the work area only when a jump or a match is reached.
  int cin0 = (flags & RE_FLAGS_BACKWARDS) ? -1 : 0;
  for (;;)
  {
    const uint32_t op = code_u8(c, ip);
    const int bm = (flags & RE_FLAGS_BACKWARDS) ? (-cin0 - 1) : cin0;
    bool match;
    uint32_t len;
    if (op == OP_ANY)
    {
      match = bm < max_bytes_matched;
      len = 1;
    }
    else if (op == OP_LITERAL || op == OP_NOT_LITERAL)
    {
      match = bm < max_bytes_matched && ((byte_at(c, pos + cin0) == code_u8(c, ip + 1)) == (op == OP_LITERAL));
      len = 2;
    }
    else if (op == OP_MASKED_LITERAL || op == OP_MASKED_NOT_LITERAL)
    {
      match = false;
      if (bm < max_bytes_matched)
      {
        const uint32_t args = code_u16(c, ip + 1);
        match = ((byte_at(c, pos + cin0) & (args >> 8)) == (args & 0xFF)) == (op == OP_MASKED_LITERAL);
      }
      len = 3;
    }
    else
      break;   // OP_REPEAT_ANY_UNGREEDY, OP_MATCH, anything unknown: the list loop takes over
    if (!match)
      return YRB_SUCCESS;   // the only entry is dropped: the list is empty, the exec ends (re.c:2445-2465)
    cin0 += input_incr;
    ip += len;
    round++;
  }

  int first = n_used++;
  P[first].input = cin0;
  P[first].round = round;
  P[first].prev = P[first].next = -1;

  int last = first;

  while (first >= 0)
  {
    int current = first;
    const uint32_t op = code_u8(c, ip);

    while (current >= 0)
    {
      const int next = P[current].next;

      if (P[current].round != round)
      {
        current = next;
        continue;
      }

      const int cin = P[current].input;
      bytes_matched = (flags & RE_FLAGS_BACKWARDS) ? (-cin - 1) : cin;

      bool match = false;

      switch (op)
      {
      case OP_ANY:
        if (bytes_matched >= max_bytes_matched) break;
        match = true;
        P[current].input += input_incr;
        break;

      case OP_LITERAL:
        if (bytes_matched >= max_bytes_matched) break;
        if (byte_at(c, pos + cin) == code_u8(c, ip + 1)) { match = true; P[current].input += input_incr; }
        break;

      case OP_NOT_LITERAL:
        if (bytes_matched >= max_bytes_matched) break;
        if (byte_at(c, pos + cin) != code_u8(c, ip + 1)) { match = true; P[current].input += input_incr; }
        break;

      case OP_MASKED_LITERAL:
      {
        if (bytes_matched >= max_bytes_matched) break;
        const uint32_t args = code_u16(c, ip + 1);
        if ((byte_at(c, pos + cin) & (args >> 8)) == (args & 0xFF)) { match = true; P[current].input += input_incr; }
        break;
      }

      case OP_MASKED_NOT_LITERAL:
      {
        if (bytes_matched >= max_bytes_matched) break;
        const uint32_t args = code_u16(c, ip + 1);
        if ((byte_at(c, pos + cin) & (args >> 8)) != (args & 0xFF)) { match = true; P[current].input += input_incr; }
        break;
      }

      case OP_REPEAT_ANY_UNGREEDY:
      {
        if (bytes_matched >= max_bytes_matched) break;

        const int rmin = (int) code_u16(c, ip + 1);
        const int rmax = (int) code_u16(c, ip + 3);
        const uint32_t next_op = code_u8(c, ip + 5);
        const uint32_t next_lit = code_u8(c, ip + 6);
        int insertion_point = current;

        match = true;

        for (int j = rmin + 1; j <= rmax; j++)
        {
          if (bytes_matched + j >= max_bytes_matched)
            break;

          const int next_input = cin + j * input_incr;

          while (P[insertion_point].next >= 0 && next_input >= P[P[insertion_point].next].input)
            insertion_point = P[insertion_point].next;

          if (next_input == P[insertion_point].input)
            continue;

          if (next_op == OP_LITERAL && next_lit != byte_at(c, pos + next_input))
            continue;

          int ni;
          if (free_head >= 0)
          {
            ni = free_head;
            free_head = P[ni].next;
          }
          else if (n_used < ws->cap)
          {
            ni = n_used++;
          }
          else
          {
            return WS_RETRY;
          }

          P[ni].input = next_input;
          P[ni].round = round + 1;
          P[ni].prev = insertion_point;
          P[ni].next = P[insertion_point].next;
          P[insertion_point].next = ni;
          if (P[ni].next >= 0) P[P[ni].next].prev = ni;
          if (insertion_point == last) last = ni;
        }

        P[current].input += input_incr * rmin;
        break;
      }

      case OP_MATCH:
        if (flags & RE_FLAGS_EXHAUSTIVE)
        {
          int64_t mpos;
          if (flags & RE_FLAGS_BACKWARDS)
          {
            const int64_t lo = pos - input_backwards_size;
            mpos = (pos + cin + 1 > lo) ? pos + cin + 1 : lo;
          }
          else
          {
            mpos = pos;
          }
          match_callback(c, mpos, bytes_matched < max_bytes_matched ? bytes_matched : max_bytes_matched, flags, *cb);
        }
        else
        {
          if (matches) *matches = bytes_matched;
          return YRB_SUCCESS;
        }
        break;

      default:
        return YRB_ERROR_INTERNAL_FATAL;
      }

      if (match)
      {
        P[current].round = round + 1;
      }
      else
      {
        if (current == first) first = P[current].next;
        if (current == last) last = P[current].prev;
        if (P[current].prev >= 0) P[P[current].prev].next = P[current].next;
        if (P[current].next >= 0) P[P[current].next].prev = P[current].prev;
        P[current].next = free_head;
        free_head = current;
      }

      current = next;
    }

    switch (op)
    {
    case OP_ANY: ip += 1; break;
    case OP_LITERAL: case OP_NOT_LITERAL: ip += 2; break;
    case OP_MASKED_LITERAL: case OP_MASKED_NOT_LITERAL: ip += 3; break;
    case OP_REPEAT_ANY_UNGREEDY: ip += 5; break;
    default: break;
    }

    round++;
  }

  return YRB_SUCCESS;
}

// ------------------------------------------------------------- fiber VM (a7)

struct Fiber
{
  uint32_t ip;
  int32_t rc;
  int16_t sp;
  int16_t prev, next;
  int16_t pad;
};

#define SYNC_MAX_DEPTH 24

struct SyncFrame
{
  int last;
  int resume;
  uint32_t mask[4];
};

#ifndef YRB_RE_NO_FASTPATH
#define YRB_RE_NO_FASTPATH 0
#endif
#define RE_STEPS_PER_POSITION_MAX (1u << 20)
#define RE_FIBERS_SMALL 64      // tier 1 pool; tier 2 has the reference's RE_MAX_FIBERS = 1024

struct FiberPool
{
  Fiber* f;          // cap entries
  uint16_t* stacks;  // cap * stack_words
  int cap;
  int stack_words;
  int free_head;
  int head, tail;
  SyncFrame* frames; // tier 2: cap frames in the work area (a frame is only pushed after a split, so depth <= fibers);
                     // tier 1: nullptr, fiber_sync uses SYNC_MAX_DEPTH frames of local memory and asks for tier 2 beyond
};

size_t re_ws_bytes_per_thread(uint32_t stack_words, int tier)
{
  size_t b = (size_t) (tier == 1 ? RE_FIBERS_SMALL : YRB_RE_MAX_FIBERS) * (sizeof(Fiber) + sizeof(uint16_t) * stack_words);
  b = (b + 15) & ~(size_t) 15;
  if (tier != 1) b += (size_t) YRB_RE_MAX_FIBERS * sizeof(SyncFrame);
  return (b + 15) & ~(size_t) 15;
}

static __device__ __forceinline__ uint16_t* fstack(FiberPool& p, int f) { return p.stacks + (size_t) f * p.stack_words; }

// re.c:1214-1248: at most RE_MAX_FIBERS live fibers
static __device__ int fiber_create(FiberPool& p, int* out)
{
  if (p.free_head < 0)
    return p.cap < YRB_RE_MAX_FIBERS ? WS_RETRY : YRB_ERROR_TOO_MANY_RE_FIBERS;

  const int f = p.free_head;
  p.free_head = p.f[f].next;
  p.f[f].ip = 0;
  p.f[f].sp = -1;
  p.f[f].rc = -1;
  p.f[f].next = -1;
  p.f[f].prev = -1;
  *out = f;
  return YRB_SUCCESS;
}

static __device__ __forceinline__ void fiber_release(FiberPool& p, int f)
{
  p.f[f].next = (int16_t) p.free_head;
  p.free_head = f;
}

// re.c:1278-1315
static __device__ bool fiber_exists(FiberPool& p, int target, int last)
{
  if (last < 0)
    return false;

  const Fiber t = p.f[target];
  const uint16_t* ts = fstack(p, target);
  const int stop = p.f[last].next;

  for (int f = p.head; f != stop; f = p.f[f].next)
  {
    const Fiber x = p.f[f];
    if (x.ip == t.ip && x.sp == t.sp && x.rc == t.rc)
    {
      bool equal = true;
      const uint16_t* xs = fstack(p, f);
      for (int i = 0; i <= x.sp; i++)
        if (xs[i] != ts[i]) { equal = false; break; }
      if (equal)
        return true;
    }
  }
  return false;
}

// re.c:1327-1358: the clone goes right after the original
static __device__ int fiber_split(FiberPool& p, int f, int* out)
{
  int n;
  const int rc = fiber_create(p, &n);
  if (rc != YRB_SUCCESS)
    return rc;

  p.f[n].sp = p.f[f].sp;
  p.f[n].ip = p.f[f].ip;
  p.f[n].rc = p.f[f].rc;

  const uint16_t* src = fstack(p, f);
  uint16_t* dst = fstack(p, n);
  for (int i = 0; i <= p.f[f].sp; i++) dst[i] = src[i];

  p.f[n].next = p.f[f].next;
  p.f[n].prev = (int16_t) f;
  if (p.f[f].next >= 0) p.f[p.f[f].next].prev = (int16_t) n;
  p.f[f].next = (int16_t) n;
  if (p.tail == f) p.tail = n;

  *out = n;
  return YRB_SUCCESS;
}

// re.c:1364-1394
static __device__ int fiber_kill(FiberPool& p, int f)
{
  const int next = p.f[f].next;
  const int prev = p.f[f].prev;

  if (prev >= 0) p.f[prev].next = (int16_t) next;
  if (next >= 0) p.f[next].prev = (int16_t) prev;
  if (p.tail == f) p.tail = prev;
  if (p.head == f) p.head = next;

  fiber_release(p, f);
  return next;
}

// re.c:1399-1433
static __device__ void fiber_kill_tail(FiberPool& p, int f)
{
  const int prev = p.f[f].prev;

  if (prev >= 0) p.f[prev].next = -1;
  p.tail = prev;
  if (p.head == f) p.head = -1;

  while (f >= 0)
  {
    const int n = p.f[f].next;
    fiber_release(p, f);
    f = n;
  }
}

static __device__ __forceinline__ void fiber_kill_all(FiberPool& p)
{
  if (p.head >= 0)
    fiber_kill_tail(p, p.head);
}

// re.c:1441-1654 _yr_re_fiber_sync.  The reference recurses for REPEAT_ANY
// (re.c:1626); here the recursion is an explicit frame stack.  The set of
// already-executed split ids (ids < RE_MAX_SPLIT_ID = 128) is a 128-bit mask.

static __device__ YRB_NOINLINE int fiber_sync(Ctx& c, FiberPool& p, int fiber_to_sync)
{
  SyncFrame local_frames[SYNC_MAX_DEPTH];
  SyncFrame* frames = p.frames ? p.frames : local_frames;
  const int max_depth = p.frames ? p.cap : SYNC_MAX_DEPTH;
  int depth = 0;

  uint32_t mask[4] = {0, 0, 0, 0};
  int fiber = fiber_to_sync;
  int last = p.f[fiber_to_sync].next;
  int branch_a, branch_b, rc;

  for (;;)
  {
    if (fiber == last)
    {
      if (depth == 0)
        return YRB_SUCCESS;
      depth--;
      fiber = frames[depth].resume;
      last = frames[depth].last;
      mask[0] = frames[depth].mask[0]; mask[1] = frames[depth].mask[1];
      mask[2] = frames[depth].mask[2]; mask[3] = frames[depth].mask[3];
      continue;
    }

    const uint32_t ip = p.f[fiber].ip;
    const uint32_t opcode = code_u8(c, ip);

    switch (opcode)
    {
    case OP_SPLIT_A:
    case OP_SPLIT_B:
    {
      const uint32_t split_id = code_u8(c, ip + 1) & 127;

      if (mask[split_id >> 5] & (1u << (split_id & 31)))
      {
        fiber = fiber_kill(p, fiber);
      }
      else
      {
        branch_a = fiber;
        rc = fiber_split(p, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;

        if (opcode == OP_SPLIT_B) { const int t = branch_a; branch_a = branch_b; branch_b = t; }

        p.f[branch_a].ip += 4;
        p.f[branch_b].ip += code_i16(c, p.f[branch_b].ip + 2);
        mask[split_id >> 5] |= (1u << (split_id & 31));
      }
      break;
    }

    case OP_REPEAT_START_GREEDY:
    case OP_REPEAT_START_UNGREEDY:
    {
      const uint32_t rmin = code_u16(c, ip + 1);
      const int32_t roff = code_i32(c, ip + 5);
      branch_a = fiber;

      if (rmin == 0)
      {
        rc = fiber_split(p, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;
        if (opcode == OP_REPEAT_START_UNGREEDY) { const int t = branch_a; branch_a = branch_b; branch_b = t; }
        p.f[branch_b].ip += roff;
      }

      p.f[branch_a].sp++;
      if (p.f[branch_a].sp < p.stack_words)
        fstack(p, branch_a)[p.f[branch_a].sp] = 0;
      p.f[branch_a].ip += 9;
      break;
    }

    case OP_REPEAT_END_GREEDY:
    case OP_REPEAT_END_UNGREEDY:
    {
      const uint32_t rmin = code_u16(c, ip + 1);
      const uint32_t rmax = code_u16(c, ip + 3);
      const int32_t roff = code_i32(c, ip + 5);
      const int sp = p.f[fiber].sp;

      if (sp < 0 || sp >= p.stack_words)
        return YRB_ERROR_INTERNAL_FATAL;

      uint16_t* st = fstack(p, fiber);
      st[sp]++;

      if (st[sp] < rmin)
      {
        p.f[fiber].ip += roff;
        break;
      }

      branch_a = fiber;

      if (st[sp] < rmax)
      {
        rc = fiber_split(p, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;
        if (opcode == OP_REPEAT_END_GREEDY) { const int t = branch_a; branch_a = branch_b; branch_b = t; }
        p.f[branch_b].ip += roff;
      }

      p.f[branch_a].sp--;
      p.f[branch_a].ip += 9;
      break;
    }

    case OP_REPEAT_ANY_GREEDY:
    case OP_REPEAT_ANY_UNGREEDY:
    {
      const int rmin = (int) code_u16(c, ip + 1);
      const int rmax = (int) code_u16(c, ip + 3);

      if (p.f[fiber].rc == -1)
        p.f[fiber].rc = 0;

      if (p.f[fiber].rc < rmin)
      {
        p.f[fiber].rc++;
        fiber = p.f[fiber].next;
      }
      else if (p.f[fiber].rc < rmax)
      {
        const int next = p.f[fiber].next;
        branch_a = fiber;
        rc = fiber_split(p, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;
        if (opcode == OP_REPEAT_ANY_UNGREEDY) { const int t = branch_a; branch_a = branch_b; branch_b = t; }

        p.f[branch_a].rc++;
        p.f[branch_b].ip += 5;
        p.f[branch_b].rc = -1;

        // recursive _yr_re_fiber_sync(branch_b), then continue at `next`
        if (depth == max_depth)   // the reference recurses (re.c:1626) as deep as it has fibers
          return p.frames ? YRB_ERROR_INTERNAL_FATAL : WS_RETRY;
        frames[depth].last = last;
        frames[depth].resume = next;
        frames[depth].mask[0] = mask[0]; frames[depth].mask[1] = mask[1];
        frames[depth].mask[2] = mask[2]; frames[depth].mask[3] = mask[3];
        depth++;
        fiber = branch_b;
        last = p.f[branch_b].next;
        mask[0] = mask[1] = mask[2] = mask[3] = 0;
      }
      else
      {
        p.f[fiber].ip += 5;
        p.f[fiber].rc = -1;
      }
      break;
    }

    case OP_JUMP:
      p.f[fiber].ip += code_i16(c, ip + 1);
      break;

    default:
      fiber = p.f[fiber].next;
    }
  }
}

static __device__ __forceinline__ bool is_word_char(Ctx& c, int64_t p, int character_size)
{
  const uint8_t b = byte_at(c, p);
  bool result = is_alnum(b) || b == '_';
  if (character_size == 2)
    result = result && (byte_at(c, p + 1) == 0);
  return result;
}

// One input-consuming opcode against the byte(s) at `input` (the cases of yr_re_exec's switch that read a
// character and move on, re.c:1770-1917, without their prologue); *len = size of the instruction.  Returns
// -1 for any other opcode.
static __device__ __forceinline__ int re_consume_op(
    Ctx& c, uint32_t op, uint32_t ip, int64_t input, int flags, int character_size, uint32_t* len)
{
  bool match;
  switch (op)
  {
  case OP_ANY:
    match = (flags & RE_FLAGS_DOT_ALL) || (byte_at(c, input) != 0x0A);
    *len = 1;
    break;
  case OP_LITERAL:
    if (flags & RE_FLAGS_NO_CASE)
      match = __ldg(c.R->lower + byte_at(c, input)) == __ldg(c.R->lower + code_u8(c, ip + 1));
    else
      match = (byte_at(c, input) == code_u8(c, ip + 1));
    *len = 2;
    break;
  case OP_NOT_LITERAL:
    match = (byte_at(c, input) != code_u8(c, ip + 1));
    *len = 2;
    break;
  case OP_MASKED_LITERAL:
  case OP_MASKED_NOT_LITERAL:
  {
    const uint32_t args = code_u16(c, ip + 1);
    match = ((byte_at(c, input) & (args >> 8)) == (args & 0xFF)) == (op == OP_MASKED_LITERAL);
    *len = 3;
    break;
  }
  case OP_CLASS:
  {
    const uint32_t chr = byte_at(c, input);
    uint32_t r = code_u8(c, ip + 2 + chr / 8) & (1u << (chr % 8));
    if (flags & RE_FLAGS_NO_CASE)
    {
      const uint32_t alt = __ldg(c.R->alter + chr);
      r |= code_u8(c, ip + 2 + alt / 8) & (1u << (alt % 8));
    }
    match = (r != 0);
    if (code_u8(c, ip + 1)) match = !match;
    *len = 34;
    break;
  }
  case OP_WORD_CHAR:
  case OP_NON_WORD_CHAR:
    match = is_word_char(c, input, character_size) == (op == OP_WORD_CHAR);
    *len = 1;
    break;
  case OP_SPACE:
  case OP_NON_SPACE:
  {
    const uint8_t b = byte_at(c, input);
    match = (b == ' ' || b == '\t' || b == '\r' || b == '\n' || b == '\v' || b == '\f') == (op == OP_SPACE);
    *len = 1;
    break;
  }
  case OP_DIGIT:
  case OP_NON_DIGIT:
  {
    const uint8_t b = byte_at(c, input);
    match = (b >= '0' && b <= '9') == (op == OP_DIGIT);
    *len = 1;
    break;
  }
  default:
    return -1;
  }
  return match ? 1 : 0;
}

// opcodes _yr_re_fiber_sync acts on (re.c:1465-1650); at any other opcode it just moves to the next fiber
static __device__ __forceinline__ bool re_is_sync_op(uint32_t op)
{
  return (op >= OP_SPLIT_A && op <= OP_REPEAT_END_UNGREEDY) || op == OP_REPEAT_ANY_GREEDY || op == OP_REPEAT_ANY_UNGREEDY;
}

// re.c:1693-2068 yr_re_exec
static __device__ YRB_NOINLINE int re_exec(
    Ctx& c, FiberPool& p, uint32_t code_off, int64_t pos, int64_t input_forwards_size, int64_t input_backwards_size,
    int flags, const CbArgs* cb, int* matches)
{
  const int character_size = (flags & RE_FLAGS_WIDE) ? 2 : 1;
  int input_incr = character_size;
  int64_t input = pos;
  int bytes_matched, max_bytes_matched;
  int fiber, next_fiber, rc;

  if (matches) *matches = -1;

  if (flags & RE_FLAGS_BACKWARDS)
  {
    max_bytes_matched = (int) (input_backwards_size < YRB_RE_SCAN_LIMIT ? input_backwards_size : YRB_RE_SCAN_LIMIT);
    input -= character_size;
    input_incr = -input_incr;
  }
  else
  {
    max_bytes_matched = (int) (input_forwards_size < YRB_RE_SCAN_LIMIT ? input_forwards_size : YRB_RE_SCAN_LIMIT);
  }

  max_bytes_matched = max_bytes_matched - max_bytes_matched % character_size;
  bytes_matched = 0;

  p.head = p.tail = -1;
  rc = fiber_create(p, &fiber);
  if (rc != YRB_SUCCESS) return rc;

  p.f[fiber].ip = code_off;
  p.head = p.tail = fiber;

  rc = fiber_sync(c, p, fiber);
  if (rc != YRB_SUCCESS) { fiber_kill_all(p); return rc; }

  while (p.head >= 0)
  {
    // ONE live fiber (most regexes most of the time: literal runs between the alternations and loops): there is
    // nothing to deduplicate and nothing to order, and while its opcodes just consume characters
    // _yr_re_fiber_sync has nothing to do either -- that stretch runs on registers, same tests, same bounds, same
    // effects as the general loop below with a one-element list.  The fiber record is written back when an
    // opcode needs the machinery (splits, repeats, jumps, anchors, boundaries, OP_MATCH).
    if (p.head == p.tail && !YRB_RE_NO_FASTPATH)
    {
      fiber = p.head;
      uint32_t fip = p.f[fiber].ip;
      bool synced = false;
      for (;;)
      {
        uint32_t len;
        const uint32_t op = code_u8(c, fip);
        if (op < OP_ANY || op > OP_MASKED_NOT_LITERAL || op == OP_MATCH)   // quick reject before the switch
          break;
        // the prologue of every consuming opcode (re.c:1733-1742)
        if ((bytes_matched >= max_bytes_matched) || (character_size == 2 && byte_at(c, input + 1) != 0))
        {
          fiber_kill(p, fiber);
          return YRB_SUCCESS;
        }
        const int m = re_consume_op(c, op, fip, input, flags, character_size, &len);
        if (m < 0)
          break;
        if (m == 0)
        {
          fiber_kill(p, fiber);
          return YRB_SUCCESS;
        }
        fip += len;
        input += input_incr;
        bytes_matched += character_size;
        if (re_is_sync_op(code_u8(c, fip)))
        {
          p.f[fiber].ip = fip;
          rc = fiber_sync(c, p, fiber);
          if (rc != YRB_SUCCESS) { fiber_kill_all(p); return rc; }
          synced = true;
          break;
        }
      }
      if (synced)
        continue;   // the position is done (the fiber's offspring, if any, runs from the next character on)
      p.f[fiber].ip = fip;
    }

    fiber = p.head;

    while (fiber >= 0)
    {
      next_fiber = p.f[fiber].next;
      if (fiber_exists(p, fiber, p.f[fiber].prev))
        fiber_kill(p, fiber);
      fiber = next_fiber;
    }

    fiber = p.head;
    uint32_t steps = 0;

    while (fiber >= 0)
    {
      const uint32_t ip = p.f[fiber].ip;
      const uint32_t op = code_u8(c, ip);
      int action = 0; // 0 none, 1 continue, 2 kill, 3 kill tail

      // A zero-width loop such as /a(\b)*?b/ can cycle at ONE input position without growing the
      // fiber list: the exit branch dies on a literal mismatch, the loop branch re-splits into the
      // fiber just freed, forever.  yr_re_exec (re.c:1693) has no way out of that either -- the
      // reference spins until killed -- and a scan only ends there because an EARLIER candidate
      // fails with ERROR_TOO_MANY_RE_FIBERS first.  Every candidate is verified here, in
      // parallel, so the cycle is cut after far more steps than a terminating execution can
      // take (each of the <= 1024 fibers runs a handful of zero-width ops per position) and
      // reported as the same explosion; the earliest failing candidate still decides the scan.
      if (++steps > RE_STEPS_PER_POSITION_MAX)
      {
        fiber_kill_all(p);
        return YRB_ERROR_TOO_MANY_RE_FIBERS;
      }
      bool match, kill;

#define PROLOG                                                                                        \
  if ((bytes_matched >= max_bytes_matched) || (character_size == 2 && byte_at(c, input + 1) != 0))    \
  {                                                                                                   \
    action = 2;                                                                                       \
    break;                                                                                            \
  }

      switch (op)
      {
      case OP_ANY:
        PROLOG;
        match = (flags & RE_FLAGS_DOT_ALL) || (byte_at(c, input) != 0x0A);
        action = match ? 0 : 2;
        p.f[fiber].ip += 1;
        break;

      case OP_REPEAT_ANY_GREEDY:
      case OP_REPEAT_ANY_UNGREEDY:
        PROLOG;
        match = (flags & RE_FLAGS_DOT_ALL) || (byte_at(c, input) != 0x0A);
        action = match ? 0 : 2;
        break;

      case OP_LITERAL:
        PROLOG;
        if (flags & RE_FLAGS_NO_CASE)
          match = __ldg(c.R->lower + byte_at(c, input)) == __ldg(c.R->lower + code_u8(c, ip + 1));
        else
          match = (byte_at(c, input) == code_u8(c, ip + 1));
        action = match ? 0 : 2;
        p.f[fiber].ip += 2;
        break;

      case OP_NOT_LITERAL:
        PROLOG;
        match = (byte_at(c, input) != code_u8(c, ip + 1));
        action = match ? 0 : 2;
        p.f[fiber].ip += 2;
        break;

      case OP_MASKED_LITERAL:
      {
        PROLOG;
        const uint32_t args = code_u16(c, ip + 1);
        match = ((byte_at(c, input) & (args >> 8)) == (args & 0xFF));
        action = match ? 0 : 2;
        p.f[fiber].ip += 3;
        break;
      }

      case OP_MASKED_NOT_LITERAL:
      {
        PROLOG;
        const uint32_t args = code_u16(c, ip + 1);
        match = ((byte_at(c, input) & (args >> 8)) != (args & 0xFF));
        action = match ? 0 : 2;
        p.f[fiber].ip += 3;
        break;
      }

      case OP_CLASS:
      {
        PROLOG;
        // re.c:98-112 _yr_re_is_char_in_class; RE_CLASS = negated u8 + bitmap[32]
        const uint32_t chr = byte_at(c, input);
        uint32_t r = code_u8(c, ip + 2 + chr / 8) & (1u << (chr % 8));
        if (flags & RE_FLAGS_NO_CASE)
        {
          const uint32_t alt = __ldg(c.R->alter + chr);
          r |= code_u8(c, ip + 2 + alt / 8) & (1u << (alt % 8));
        }
        match = (r != 0);
        if (code_u8(c, ip + 1)) match = !match;
        action = match ? 0 : 2;
        p.f[fiber].ip += 34;
        break;
      }

      case OP_WORD_CHAR:
        PROLOG;
        match = is_word_char(c, input, character_size);
        action = match ? 0 : 2;
        p.f[fiber].ip += 1;
        break;

      case OP_NON_WORD_CHAR:
        PROLOG;
        match = !is_word_char(c, input, character_size);
        action = match ? 0 : 2;
        p.f[fiber].ip += 1;
        break;

      case OP_SPACE:
      case OP_NON_SPACE:
      {
        PROLOG;
        const uint8_t b = byte_at(c, input);
        match = (b == ' ' || b == '\t' || b == '\r' || b == '\n' || b == '\v' || b == '\f');
        if (op == OP_NON_SPACE) match = !match;
        action = match ? 0 : 2;
        p.f[fiber].ip += 1;
        break;
      }

      case OP_DIGIT:
      {
        PROLOG;
        const uint8_t b = byte_at(c, input);
        match = (b >= '0' && b <= '9');
        action = match ? 0 : 2;
        p.f[fiber].ip += 1;
        break;
      }

      case OP_NON_DIGIT:
      {
        PROLOG;
        const uint8_t b = byte_at(c, input);
        match = !(b >= '0' && b <= '9');
        action = match ? 0 : 2;
        p.f[fiber].ip += 1;
        break;
      }

      case OP_WORD_BOUNDARY:
      case OP_NON_WORD_BOUNDARY:
        if (bytes_matched == 0 && input_backwards_size < character_size)
          match = true;
        else if (bytes_matched >= max_bytes_matched)
          match = true;
        else
          match = is_word_char(c, input, character_size) != is_word_char(c, input - input_incr, character_size);

        if (op == OP_NON_WORD_BOUNDARY) match = !match;
        action = match ? 1 : 2;
        p.f[fiber].ip += 1;
        break;

      case OP_MATCH_AT_START:
        if (flags & RE_FLAGS_BACKWARDS)
          kill = input_backwards_size > (int64_t) bytes_matched;
        else
          kill = input_backwards_size > 0 || (bytes_matched != 0);
        action = kill ? 2 : 1;
        p.f[fiber].ip += 1;
        break;

      case OP_MATCH_AT_END:
        kill = (flags & RE_FLAGS_BACKWARDS) || input_forwards_size > (int64_t) bytes_matched;
        action = kill ? 2 : 1;
        p.f[fiber].ip += 1;
        break;

      case OP_MATCH:
        if (matches) *matches = bytes_matched;

        if (flags & RE_FLAGS_EXHAUSTIVE)
        {
          if (cb != nullptr)
            match_callback(c, (flags & RE_FLAGS_BACKWARDS) ? input + character_size : pos, bytes_matched, flags, *cb);
          action = 2;
        }
        else
        {
          action = 3;
        }
        break;

      default:
        fiber_kill_all(p);
        return YRB_ERROR_INTERNAL_FATAL;
      }
#undef PROLOG

      switch (action)
      {
      case 2:
        fiber = fiber_kill(p, fiber);
        break;
      case 3:
        fiber_kill_tail(p, fiber);
        fiber = -1;
        break;
      case 1:
        rc = fiber_sync(c, p, fiber);
        if (rc != YRB_SUCCESS) { fiber_kill_all(p); return rc; }
        break;
      default:
        next_fiber = p.f[fiber].next;
        rc = fiber_sync(c, p, fiber);
        if (rc != YRB_SUCCESS) { fiber_kill_all(p); return rc; }
        fiber = next_fiber;
      }
    }

    input += input_incr;
    bytes_matched += character_size;
  }

  return YRB_SUCCESS;
}

// ------------------------------------------------------- literal verifier (a5)

static __device__ void verify_literal(Ctx& c, const yrb_acm& m, const yrb_string& s, int64_t offset)
{
  const uint8_t* __restrict__ str = c.R->lit + s.lit_off;
  const uint8_t* __restrict__ lower = c.R->lower;
  const int64_t avail = c.gsize - offset;
  const int64_t len = s.length;
  int forward_matches = 0, flags = 0;
  uint32_t xor_key = 0;
  int64_t i;

  if (s.flags & YRB_SF_FITS_IN_ATOM)
  {
    forward_matches = (int) m.backtrack;
  }
  else if (s.flags & YRB_SF_NO_CASE)
  {
    if ((s.flags & YRB_SF_ASCII) && avail >= len)
    {
      for (i = 0; i < len && __ldg(lower + byte_at(c, offset + i)) == __ldg(lower + __ldg(str + i)); i++) {}
      forward_matches = (i == len) ? (int) i : 0;
    }
    if ((s.flags & YRB_SF_WIDE) && forward_matches == 0 && avail >= len * 2)
    {
      for (i = 0; i < len && __ldg(lower + byte_at(c, offset + 2 * i)) == __ldg(lower + __ldg(str + i)) &&
                  byte_at(c, offset + 2 * i + 1) == 0; i++) {}
      forward_matches = (i == len) ? (int) (i * 2) : 0;
    }
  }
  else
  {
    if ((s.flags & YRB_SF_ASCII) && avail >= len)
    {
      for (i = 0; i < len && byte_at(c, offset + i) == __ldg(str + i); i++) {}
      forward_matches = (i == len) ? (int) i : 0;
    }
    if ((s.flags & YRB_SF_WIDE) && forward_matches == 0 && avail >= len * 2)
    {
      for (i = 0; i < len && byte_at(c, offset + 2 * i) == __ldg(str + i) && byte_at(c, offset + 2 * i + 1) == 0; i++) {}
      forward_matches = (i == len) ? (int) (i * 2) : 0;
    }
    if ((s.flags & YRB_SF_XOR) && forward_matches == 0)
    {
      if ((s.flags & YRB_SF_WIDE) && avail >= len * 2)
      {
        const uint32_t k = byte_at(c, offset) ^ __ldg(str);
        for (i = 0; i < len && byte_at(c, offset + 2 * i) == (__ldg(str + i) ^ k) &&
                    (byte_at(c, offset + 2 * i + 1) ^ k) == 0; i++) {}
        forward_matches = (i == len) ? (int) (i * 2) : 0;
        if (forward_matches > 0) xor_key = k;
      }
      if (forward_matches == 0 && avail >= len)
      {
        const uint32_t k = byte_at(c, offset) ^ __ldg(str);
        for (i = 0; i < len && byte_at(c, offset + i) == (__ldg(str + i) ^ k); i++) {}
        forward_matches = (i == len) ? (int) i : 0;
        if (forward_matches > 0) xor_key = k;
      }
    }
  }

  if (forward_matches == 0)
    return;

  if (forward_matches == s.length * 2)
    flags |= RE_FLAGS_WIDE;
  if (s.flags & YRB_SF_NO_CASE)
    flags |= RE_FLAGS_NO_CASE;

  CbArgs a;
  a.sidx = m.string_idx;
  a.forward_matches = forward_matches;
  a.full_word = (s.flags & YRB_SF_FULL_WORD) != 0;
  a.xor_key = xor_key;
  a.group = s.chain_head;

  match_callback(c, offset, 0, flags, a);
}

// ------------------------------------------------------------ RE verifier (a6)

template <bool FAST>
static __device__ int verify_re(Ctx& c, FastWs* fws, FiberPool* fp, const yrb_acm& m, const yrb_string& s, int64_t offset)
{
  int forward_matches = -1, backward_matches = -1;
  int flags = 0, rc;

  if (s.flags & YRB_SF_GREEDY_REGEXP) flags |= RE_FLAGS_GREEDY;
  if (s.flags & YRB_SF_NO_CASE) flags |= RE_FLAGS_NO_CASE;
  if (s.flags & YRB_SF_DOT_ALL) flags |= RE_FLAGS_DOT_ALL;

  const int64_t fwd = c.gsize - offset;
  const int64_t bwd = offset;

#define EXEC(code, fl, cb, mt) \
  (FAST ? re_fast_exec(c, fws, (code), offset, fwd, bwd, (fl), (cb), (mt)) : re_exec(c, *fp, (code), offset, fwd, bwd, (fl), (cb), (mt)))

  if (s.flags & (YRB_SF_ASCII | YRB_SF_BASE64 | YRB_SF_BASE64_WIDE))
  {
    rc = EXEC(m.fwd_code, flags, nullptr, &forward_matches);
    if (rc != YRB_SUCCESS) return rc;
  }

  if (forward_matches == -1 && (s.flags & YRB_SF_WIDE) && !(s.flags & (YRB_SF_BASE64 | YRB_SF_BASE64_WIDE)))
  {
    flags |= RE_FLAGS_WIDE;
    rc = EXEC(m.fwd_code, flags, nullptr, &forward_matches);
    if (rc != YRB_SUCCESS) return rc;
  }

  if (forward_matches == -1)
    return YRB_SUCCESS;

  if (forward_matches == 0 && m.bwd_code == YRB_NONE)
    return YRB_SUCCESS;

  CbArgs a;
  a.sidx = m.string_idx;
  a.forward_matches = forward_matches;
  a.full_word = (s.flags & YRB_SF_FULL_WORD) != 0;
  a.xor_key = 0;
  a.group = s.chain_head;

  if (m.bwd_code != YRB_NONE)
    return EXEC(m.bwd_code, flags | RE_FLAGS_BACKWARDS | RE_FLAGS_EXHAUSTIVE, &a, &backward_matches);

  match_callback(c, offset, 0, flags, a);
  return YRB_SUCCESS;
#undef EXEC
}

// -------------------------------------------------------------------- kernels

// Returns the position (in the coordinates of c) the candidate's string would start at.
static __device__ __forceinline__ int64_t ctx_init(
    Ctx& c, const DevRules& R, const ScanView& V, const RawOut& out, const Cand& cd, const yrb_acm& m)
{
  const int64_t i = (int64_t) (cd.key >> KEY_I_SHIFT);
  c.R = &R;
  c.out = out;
  c.cand_key = cand_key(cd, R.key_seq_bits);
  c.seq = 0;

  if (V.n_seg > 0)
  {
    c.seg = seg_find(V, i - 1);
    c.bias = __ldg(&V.seg_start[c.seg]);
    c.d = V.d + (c.bias - V.g0);
    c.g0 = 0;
    c.g1 = c.gsize = __ldg(&V.seg_len[c.seg]);
    return i - c.bias - (int64_t) m.backtrack;
  }

  c.seg = 0;
  c.bias = 0;
  c.d = V.d - V.g0;
  c.g0 = V.g0;
  c.g1 = V.g1;
  c.gsize = V.gsize;
  return i - (int64_t) m.backtrack;
}

static __device__ __forceinline__ void report_error_ctx(const Ctx& c, uint32_t sidx, int code)
{
  report_error(c.out, sidx, c.cand_key, code, c.R->key_seq_bits);
  if (c.out.item_dirty) c.out.item_dirty[c.seg] = 1;
}

__global__ void __launch_bounds__(128)
verify_literal_kernel(DevRules R, ScanView V, const Cand* cands, unsigned long long n, RawOut out)
{
  const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

  for (unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
  {
    const Cand cd = cands[t];
    const yrb_acm m = R.acm[cd.acm_idx];
    const yrb_string s = R.S[m.string_idx];
    Ctx c;
    const int64_t at = ctx_init(c, R, V, out, cd, m);
    verify_literal(c, m, s, at);
  }
}

// Two tiers per VM class.  Tier 1 verifies every candidate with a small work area
// (1-2 KB per thread, so hundreds of thousands of candidates are in flight and
// the lists stay cache resident); a candidate whose lists outgrow it is queued
// and verified again from scratch by tier 2 with the reference's full limits.
// Records a candidate emitted before overflowing are emitted again with the same
// discovery keys; resolve.cu treats equal keys as one record.
// TIER 1: every candidate; TIER 2: the retry queue of tier 1; TIER 3: the retry queue of tier 2
// (64k positions, up to 2048 threads); TIER 4: the retry queue of tier 3 (1M positions, 32 threads).
// Tiers 3 and 4 alternate between the two queues: tier 3 reads retry_fast3 / n_retry3 and refills
// retry_fast / n_retry[0] (reset by the host in between), which tier 4 reads.
template <int TIER>
__global__ void __launch_bounds__(128)
verify_fast_kernel(
    DevRules R, ScanView V, const Cand* cands, unsigned long long n, RawOut out, uint8_t* ws, size_t ws_per_thread,
    int cap, const uint32_t* retry, uint32_t* retry_next, const uint32_t* perm = nullptr)   // perm: tier 1, sorted order
{
  const unsigned long long tid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
  FastWs my;
  my.v = reinterpret_cast<FastPos*>(ws + tid * ws_per_thread);
  my.cap = cap;

  if (TIER == 2 || TIER == 4) n = out.st->n_retry[0];
  if (TIER == 3) n = out.st->n_retry3;

  for (unsigned long long t = tid; t < n; t += stride)
  {
    const unsigned long long ci = TIER > 1 ? retry[t] : (perm != nullptr ? perm[t] : t);
    const Cand cd = cands[ci];
    const yrb_acm m = R.acm[cd.acm_idx];
    const yrb_string s = R.S[m.string_idx];
    Ctx c;
    const int64_t at = ctx_init(c, R, V, out, cd, m);
    int rc = verify_re<true>(c, &my, nullptr, m, s, at);
    if (rc == WS_RETRY)
    {
      if (TIER == 1)
      {
        retry_next[atomicAdd(&out.st->n_retry[0], 1u)] = (uint32_t) ci;
        continue;
      }
      if (TIER == 2)
      {
        retry_next[atomicAdd(&out.st->n_retry3, 1u)] = (uint32_t) ci;
        continue;
      }
      if (TIER == 3)
      {
        retry_next[atomicAdd(&out.st->n_retry[0], 1u)] = (uint32_t) ci;
        continue;
      }
      atomicExch(&out.st->ws_overflow, 1u); // the reference would still be calling malloc: give up honestly
      rc = YRB_ERROR_INSUFFICIENT_MEMORY;
    }
    if (rc != YRB_SUCCESS)
      report_error_ctx(c, m.string_idx, rc);
  }
}

// Fiber VM tiers.  TIER 0: every candidate, with a pool of RE_FIBERS_LOCAL fibers in a per-thread LOCAL array:
// local memory is interleaved across the lanes of a warp by the hardware (the same field of the same fiber of 32
// candidates is one 128-byte line) and cached write-back in L1, where the per-thread slices of a global work
// area (1.5 KB apart) cost a line per lane and access.  Nearly every regex lives with a handful of fibers, so this
// is where the class is verified; what outgrows it is queued (retry_out, n_retry_re0) for TIER 1: RE_FIBERS_SMALL
// fibers in the global work area (also the entry tier when the ruleset's repeat stacks are deeper than
// RE_LOCAL_STACK: then retry_in == nullptr and it takes every candidate), which queues for TIER 2 (n_retry[1]):
// the reference's limits, RE_MAX_FIBERS and a frame per fiber.
#define RE_FIBERS_LOCAL 16

template <int TIER>
__global__ void __launch_bounds__(128)
verify_re_kernel(
    DevRules R, ScanView V, const Cand* cands, unsigned long long n, RawOut out, uint8_t* ws, size_t ws_per_thread,
    int cap, const uint32_t* retry_in, uint32_t* retry_out, const uint32_t* perm = nullptr)   // perm: entry tier, sorted order
{
  unsigned long long tid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;

  if (TIER == 0 && ws_per_thread < 32)
  {
    // The candidates of a warp run different regexes over different bytes: the warp executes their paths one
    // after the other (3.4 of 32 lanes active per instruction, profiles/r02_ncu_verify_re*), so a warp's time
    // grows with its candidates while a scan with a few thousand candidates per SM leaves most warp slots
    // empty.  With `ws_per_thread` (< 32) active lanes per warp the same candidates spread over more warps.
    const unsigned lanes = (unsigned) ws_per_thread, lane = threadIdx.x & 31u;
    if (lane >= lanes) return;
    tid = (tid >> 5) * lanes + lane;
    stride = (stride >> 5) * lanes;
  }

  if (TIER == 1 && retry_in != nullptr) n = out.st->n_retry_re0;
  if (TIER == 2) n = out.st->n_retry[1];
  if (tid >= n) return;

  Fiber local_f[TIER == 0 ? RE_FIBERS_LOCAL : 1];
  uint16_t local_s[TIER == 0 ? RE_FIBERS_LOCAL * RE_LOCAL_STACK : 1];

  FiberPool p;
  if (TIER == 0)
  {
    p.f = local_f;
    p.stacks = local_s;
    cap = RE_FIBERS_LOCAL;
  }
  else
  {
    p.f = reinterpret_cast<Fiber*>(ws + tid * ws_per_thread);
    p.stacks = reinterpret_cast<uint16_t*>(p.f + cap);
  }
  p.cap = cap;
  p.stack_words = (int) R.stack_words;
  p.frames = nullptr;
  if (TIER == 2)
  {
    size_t fb = (size_t) cap * (sizeof(Fiber) + sizeof(uint16_t) * R.stack_words);
    fb = (fb + 15) & ~(size_t) 15;
    p.frames = reinterpret_cast<SyncFrame*>(ws + tid * ws_per_thread + fb);
  }

  // free list over the pool (kept consistent across candidates: every exec returns all of its fibers)
  for (int f = 0; f < cap; f++)
    p.f[f].next = (int16_t) ((f + 1 < cap) ? f + 1 : -1);
  p.free_head = 0;
  p.head = p.tail = -1;

  for (unsigned long long t = tid; t < n; t += stride)
  {
    const unsigned long long ci =
        (TIER == 2 || (TIER == 1 && retry_in != nullptr)) ? retry_in[t] : (perm != nullptr ? perm[t] : t);
    const Cand cd = cands[ci];
    const yrb_acm m = R.acm[cd.acm_idx];
    const yrb_string s = R.S[m.string_idx];
    Ctx c;
    const int64_t at = ctx_init(c, R, V, out, cd, m);
    int rc = verify_re<false>(c, nullptr, &p, m, s, at);
    if (rc == WS_RETRY)
    {
      if (TIER == 0)
      {
        retry_out[atomicAdd(&out.st->n_retry_re0, 1u)] = (uint32_t) ci;
        continue;
      }
      if (TIER == 1)
      {
        retry_out[atomicAdd(&out.st->n_retry[1], 1u)] = (uint32_t) ci;
        continue;
      }
      rc = YRB_ERROR_INTERNAL_FATAL;
    }
    if (rc != YRB_SUCCESS)
      report_error_ctx(c, m.string_idx, rc);
  }
}

#ifndef YRB_HOST_EMULATION   // tests/harness/vm_host.cpp compiles the device code above for the CPU and drives it itself
void launch_verify(
    const DevRules& R, const ScanView& V, const CandOut& cands, const unsigned long long n_cand[NUM_CLS],
    const RawOut& out, const VerifyWs& W, cudaStream_t s, unsigned* launches, const SideStreams* side)
{
  const int T = 128;

  // The classes only share atomically appended outputs, so they run side by side: the largest class (by
  // count) stays on the scanner's stream, the others go to the side streams and are joined at the end.
  const int n_classes = (n_cand[CLS_LITERAL] > 0) + (n_cand[CLS_FAST] > 0) + (n_cand[CLS_RE] > 0);
  cudaStream_t cs[NUM_CLS] = {s, s, s};
  bool joined[2] = {false, false};
  if (side != nullptr && n_classes > 1)
  {
    int main_cls = CLS_LITERAL;
    for (int c = 0; c < NUM_CLS; c++) if (n_cand[c] > n_cand[main_cls]) main_cls = c;
    cudaEventRecord(side->ev[0], s);
    int k = 0;
    for (int c = 0; c < NUM_CLS; c++)
    {
      if (c == main_cls || n_cand[c] == 0) continue;
      cs[c] = side->aux[k];
      cudaStreamWaitEvent(side->aux[k], side->ev[0], 0);
      joined[k] = true;
      k++;
    }
  }

  if (n_cand[CLS_LITERAL] > 0)
  {
    unsigned long long blocks = (n_cand[CLS_LITERAL] + T - 1) / T;
    if (blocks > 148 * 64) blocks = 148 * 64;
    verify_literal_kernel<<<(unsigned) blocks, T, 0, cs[CLS_LITERAL]>>>(R, V, cands.cand[CLS_LITERAL], n_cand[CLS_LITERAL], out);
    (*launches)++;
  }

  const uint32_t* perm[NUM_CLS] = {nullptr, nullptr, nullptr};
  {
    size_t at = 0;   // the classes share the scratch: each takes its slice
    for (int c = CLS_FAST; c <= CLS_RE; c++)
    {
      if (!W.sort_cls[c] || n_cand[c] == 0 || at + n_cand[c] > W.sort_cap) continue;
      perm[c] = launch_cand_sort(cands.cand[c], (uint32_t) n_cand[c], R.num_acm, W.sort_k[0] + at, W.sort_k[1] + at,
                                 W.sort_v[0] + at, W.sort_v[1] + at, (uint8_t*) W.sort_tmp + (c == CLS_RE ? W.sort_tmp_bytes / 2 : 0),
                                 W.sort_tmp_bytes / 2, cs[c]);
      at += n_cand[c];
      (*launches) += 4;
    }
  }

  if (n_cand[CLS_FAST] > 0)
  {
    unsigned long long threads = n_cand[CLS_FAST] < W.fast1_threads ? n_cand[CLS_FAST] : W.fast1_threads;
    unsigned blocks = (unsigned) ((threads + T - 1) / T);
    if ((unsigned long long) blocks * T > W.fast1_threads) blocks = (unsigned) (W.fast1_threads / T);
    if (blocks == 0) blocks = 1;
    verify_fast_kernel<1><<<blocks, T, 0, cs[CLS_FAST]>>>(
        R, V, cands.cand[CLS_FAST], n_cand[CLS_FAST], out, (uint8_t*) W.fast1, fast_ws_bytes_per_thread(1),
        FAST_POS_CAP_SMALL, nullptr, W.retry_fast, perm[CLS_FAST]);
    (*launches)++;
  }

  if (n_cand[CLS_RE] > 0)
  {
    unsigned long long threads = n_cand[CLS_RE] < W.re1_threads ? n_cand[CLS_RE] : W.re1_threads;
    unsigned blocks = (unsigned) ((threads + T - 1) / T);
    if ((unsigned long long) blocks * T > W.re1_threads) blocks = (unsigned) (W.re1_threads / T);
    if (blocks == 0) blocks = 1;
    if (W.retry_re0 != nullptr && R.stack_words <= RE_LOCAL_STACK)
    {
      // tier 0 needs no work area: a thread per candidate; tier 1 follows for what it queued (count read on the device)
      // (fewer active lanes per warp spread the candidates over more warps, but the kernel is bound by
      // instruction fetch, not by warp slots: measured 2x slower with 8 lanes, profiles/r02_kernel_notes.txt)
      const unsigned lanes = 32;
      unsigned long long b0 = (n_cand[CLS_RE] * (32 / lanes) + T - 1) / T;
      if (b0 > 148ull * 1024) b0 = 148ull * 1024;
      verify_re_kernel<0><<<(unsigned) b0, T, 0, cs[CLS_RE]>>>(
          R, V, cands.cand[CLS_RE], n_cand[CLS_RE], out, nullptr, lanes, RE_FIBERS_LOCAL, nullptr, W.retry_re0, perm[CLS_RE]);
      verify_re_kernel<1><<<blocks, T, 0, cs[CLS_RE]>>>(
          R, V, cands.cand[CLS_RE], 0, out, (uint8_t*) W.re1, re_ws_bytes_per_thread(R.stack_words, 1),
          RE_FIBERS_SMALL, W.retry_re0, W.retry_re);
      (*launches) += 2;
    }
    else
    {
      verify_re_kernel<1><<<blocks, T, 0, cs[CLS_RE]>>>(
          R, V, cands.cand[CLS_RE], n_cand[CLS_RE], out, (uint8_t*) W.re1, re_ws_bytes_per_thread(R.stack_words, 1),
          RE_FIBERS_SMALL, nullptr, W.retry_re, perm[CLS_RE]);
      (*launches)++;
    }
  }

  for (int k = 0; k < 2; k++)
  {
    if (!joined[k]) continue;
    cudaEventRecord(side->ev[1 + k], side->aux[k]);
    cudaStreamWaitEvent(s, side->ev[1 + k], 0);
  }
}

// Tier 2 (and 3) run only when tier 1 (2) queued something: their work areas are big
// and most scans never need them.
void launch_verify_tier2(
    const DevRules& R, const ScanView& V, const CandOut& cands, const RawOut& out, const VerifyWs& W, bool fast, bool re,
    cudaStream_t s, unsigned* launches)
{
  const int T = 128;
  if (fast)
  {
    verify_fast_kernel<2><<<(unsigned) ((W.fast2_threads + T - 1) / T), T, 0, s>>>(
        R, V, cands.cand[CLS_FAST], 0, out, (uint8_t*) W.fast2, fast_ws_bytes_per_thread(2), FAST_POS_CAP, W.retry_fast,
        W.retry_fast3);
    (*launches)++;
  }
  if (re)
  {
    verify_re_kernel<2><<<(unsigned) ((W.re2_threads + T - 1) / T), T, 0, s>>>(
        R, V, cands.cand[CLS_RE], 0, out, (uint8_t*) W.re2, re_ws_bytes_per_thread(R.stack_words, 2),
        YRB_RE_MAX_FIBERS, W.retry_re, nullptr);
    (*launches)++;
  }
}

void launch_verify_fast_tier3(
    const DevRules& R, const ScanView& V, const CandOut& cands, const RawOut& out, const VerifyWs& W, int tier,
    cudaStream_t s, unsigned* launches)
{
  const unsigned blocks = (unsigned) ((W.fast3_threads + 31) / 32);
  if (tier == 3)
    verify_fast_kernel<3><<<blocks, 32, 0, s>>>(
        R, V, cands.cand[CLS_FAST], 0, out, (uint8_t*) W.fast3, fast_ws_bytes_per_thread(3), FAST_POS_CAP_MID,
        W.retry_fast3, W.retry_fast);
  else
    verify_fast_kernel<4><<<blocks, 32, 0, s>>>(
        R, V, cands.cand[CLS_FAST], 0, out, (uint8_t*) W.fast3, fast_ws_bytes_per_thread(4), FAST_POS_CAP_HUGE,
        W.retry_fast, nullptr);
  (*launches)++;
}
#endif
