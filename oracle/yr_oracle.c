/*
 * yr_oracle.c -- CPU restatement of YARA 4.2.1's byte-scanning hot path over the
 * flat rules image (include/yrb_image.h).
 *
 * TEST INFRASTRUCTURE ONLY -- see yr_oracle.h.  Sequential, single block,
 * written for clarity; every function cites the reference lines it follows
 * (paths relative to /root/reference/libyara).
 */
#include "yr_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ types */

typedef struct
{
  int64_t offset;
  int32_t length;
  int32_t chain_length;
  uint8_t xor_key;
} om_t;

/* YR_MATCHES (types.h:469-475) as a sorted vector */
typedef struct
{
  om_t* v;
  int32_t n, cap;
} olist_t;

/* RE_FIBER (types.h:424-434) in an index-linked pool */
typedef struct
{
  uint32_t ip;
  int32_t sp;
  int32_t rc;
  int32_t prev, next;
  uint16_t* stack;
} ofiber_t;

typedef struct
{
  const yrb_image_header* h;
  const uint32_t* T;
  const uint32_t* M;
  const yrb_acm* acm;
  const yrb_string* S;
  const uint8_t* lit;
  const uint8_t* code;
  const uint8_t* lower;
  const uint8_t* alter;

  const uint8_t* data;
  size_t size;
  uint64_t base;
  int scan_flags;
  int too_many_action;

  olist_t* matches;
  olist_t* unconf;
  uint8_t* disabled;

  /* fiber pool (re.c:1214-1248): live-count cap of RE_MAX_FIBERS */
  ofiber_t* fib;
  uint16_t* fib_stacks;
  int32_t stack_words;
  int32_t free_head;
  int32_t live;
  int32_t head, tail; /* the current exec's fiber list */

  yro_result* r;
} octx_t;

/* callback arguments (scan.c:41-60 CALLBACK_ARGS) */
typedef struct
{
  uint32_t sidx;
  int forward_matches;
  int full_word;
  uint8_t xor_key;
} cbargs_t;

#define RE_FLAGS_BACKWARDS  0x04
#define RE_FLAGS_EXHAUSTIVE 0x08
#define RE_FLAGS_WIDE       0x10
#define RE_FLAGS_NO_CASE    0x20
#define RE_FLAGS_DOT_ALL    0x80
#define RE_FLAGS_GREEDY     0x400

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

static inline uint16_t rd_u16(const uint8_t* p) { return (uint16_t) (p[0] | (p[1] << 8)); }
static inline int16_t rd_i16(const uint8_t* p) { return (int16_t) rd_u16(p); }
static inline int32_t rd_i32(const uint8_t* p)
{ return (int32_t) ((uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24)); }

/* strutils.c:240-244 yr_isalnum */
static inline int is_alnum(uint8_t c)
{ return (c >= 0x30 && c <= 0x39) || (c >= 0x41 && c <= 0x5a) || (c >= 0x61 && c <= 0x7a); }

/* ---------------------------------------------------- match lists (a10) */

/* scan.c:290-368 _yr_scan_add_match_to_list */
static int list_add(olist_t* L, const om_t* m, int replace_if_exists)
{
  if (L->n == YRB_MAX_STRING_MATCHES)
    return YRB_ERROR_TOO_MANY_MATCHES;

  int32_t ip = L->n - 1; /* insertion_point = tail */

  while (ip >= 0)
  {
    if (m->offset == L->v[ip].offset)
    {
      if (replace_if_exists)
        L->v[ip].length = m->length; /* length, data_length, data */
      return YRB_SUCCESS;
    }
    if (m->offset > L->v[ip].offset)
      break;
    ip--;
  }

  if (L->n == L->cap)
  {
    L->cap = L->cap ? L->cap * 2 : 8;
    L->v = (om_t*) realloc(L->v, sizeof(om_t) * (size_t) L->cap);
  }

  memmove(&L->v[ip + 2], &L->v[ip + 1], sizeof(om_t) * (size_t) (L->n - (ip + 1)));
  L->v[ip + 1] = *m;
  L->n++;
  return YRB_SUCCESS;
}

/* scan.c:370-389 _yr_scan_remove_match_from_list */
static void list_remove(olist_t* L, int32_t i)
{
  memmove(&L->v[i], &L->v[i + 1], sizeof(om_t) * (size_t) (L->n - i - 1));
  L->n--;
}

/* ------------------------------------------------- chained strings (a11) */

/* scan.c:257-288 _yr_scan_update_match_chain_length.  match_to_update lives in
   unconf[sidx] at index mi. */
static void update_chain_length(octx_t* c, uint32_t sidx, int32_t mi, int chain_length)
{
  om_t* mu = &c->unconf[sidx].v[mi];

  if (mu->chain_length == chain_length)
    return;

  mu->chain_length = chain_length;

  const yrb_string* s = &c->S[sidx];

  if (s->chained_to == YRB_NONE)
    return;

  olist_t* P = &c->unconf[s->chained_to];
  int64_t upd_offset = mu->offset;

  for (int32_t k = 0; k < P->n; k++)
  {
    int64_t ending_offset = P->v[k].offset + P->v[k].length;

    if (ending_offset + s->chain_gap_max >= upd_offset && ending_offset + s->chain_gap_min <= upd_offset)
      update_chain_length(c, s->chained_to, k, chain_length + 1);
  }
}

/* scan.c:406-632 _yr_scan_verify_chained_string_match */
static int verify_chained(
    octx_t* c, uint32_t sidx, int64_t match_offset, int32_t match_length, uint8_t xor_key)
{
  const yrb_string* ms = &c->S[sidx];
  int add_match = 0;

  if (ms->chained_to == YRB_NONE)
  {
    add_match = 1;
  }
  else
  {
    int64_t lowest_offset = c->unconf[sidx].n > 0 ? c->unconf[sidx].v[0].offset : match_offset;
    olist_t* P = &c->unconf[ms->chained_to];
    int32_t k = 0;

    while (k < P->n)
    {
      int64_t ending_offset = P->v[k].offset + P->v[k].length;

      if (ending_offset + ms->chain_gap_max < lowest_offset)
      {
        list_remove(P, k); /* negatively confirmed; next element slides into k */
        continue;
      }
      else if (
          ending_offset + ms->chain_gap_max >= match_offset &&
          ending_offset + ms->chain_gap_min <= match_offset)
      {
        add_match = 1;
        break;
      }
      k++;
    }
  }

  if (!add_match)
    return YRB_SUCCESS;

  if (ms->flags & YRB_SF_CHAIN_TAIL)
  {
    olist_t* P = &c->unconf[ms->chained_to];

    for (int32_t k = 0; k < P->n; k++)
    {
      int64_t ending_offset = P->v[k].offset + P->v[k].length;

      if (ending_offset + ms->chain_gap_max >= match_offset && ending_offset + ms->chain_gap_min <= match_offset)
        update_chain_length(c, ms->chained_to, k, 1);
    }

    int32_t full_chain_length = 0;
    uint32_t hs = sidx;

    while (c->S[hs].chained_to != YRB_NONE)
    {
      full_chain_length++;
      hs = c->S[hs].chained_to;
    }

    olist_t* H = &c->unconf[hs];
    int32_t k = 0;

    while (k < H->n)
    {
      if (H->v[k].chain_length == full_chain_length)
      {
        om_t m = H->v[k];
        list_remove(H, k);
        m.length = (int32_t) (match_offset - m.offset + match_length);
        int rc = list_add(&c->matches[hs], &m, 0);
        if (rc != YRB_SUCCESS)
          return rc;
        continue;
      }
      k++;
    }
  }
  else
  {
    om_t m;
    m.offset = match_offset;
    m.length = match_length;
    m.chain_length = 0;
    m.xor_key = xor_key;
    return list_add(&c->unconf[sidx], &m, 0);
  }

  return YRB_SUCCESS;
}

/* ------------------------------------------------- match callback (a9) */

/* Optional trace of the _yr_scan_match_callback calls that survive the fullword test, in call order
   (single-threaded test aid: tests/test_vm_host_cpu.py compares it with the records the product's
   verifier code produces). */
static yro_callback* g_trace = NULL;
static size_t g_trace_cap = 0, g_trace_n = 0;

void yro_trace_begin(yro_callback* buf, size_t cap)
{
  g_trace = buf;
  g_trace_cap = cap;
  g_trace_n = 0;
}

size_t yro_trace_end(void)
{
  g_trace = NULL;
  return g_trace_n; /* may exceed cap: the caller retries with a larger buffer */
}

/* scan.c:634-765 _yr_scan_match_callback.  match_pos = match_data - data. */
static int match_callback(octx_t* c, int64_t match_pos, int32_t match_length, int re_flags, const cbargs_t* a)
{
  const yrb_string* s = &c->S[a->sidx];
  const uint8_t* d = c->data;
  int64_t size = (int64_t) c->size;

  match_length += a->forward_matches;

  if (a->full_word)
  {
    if (re_flags & RE_FLAGS_WIDE)
    {
      if (match_pos >= 2 && d[match_pos - 1] == 0 && is_alnum(d[match_pos - 2]))
        return YRB_SUCCESS;
      if (match_pos + match_length + 1 < size && d[match_pos + match_length + 1] == 0 &&
          is_alnum(d[match_pos + match_length]))
        return YRB_SUCCESS;
    }
    else
    {
      if (match_pos >= 1 && is_alnum(d[match_pos - 1]))
        return YRB_SUCCESS;
      if (match_pos + match_length < size && is_alnum(d[match_pos + match_length]))
        return YRB_SUCCESS;
    }
  }

  if (g_trace != NULL)
  {
    if (g_trace_n < g_trace_cap)
    {
      yro_callback* t = &g_trace[g_trace_n];
      t->offset = match_pos;
      t->string_idx = a->sidx;
      t->length = match_length;
      t->xor_key = a->xor_key;
      t->pad = 0;
    }
    g_trace_n++;
  }

  if (s->flags & YRB_SF_CHAIN_PART)
    return verify_chained(c, a->sidx, match_pos, match_length, a->xor_key);

  om_t m;
  m.offset = match_pos;
  m.length = match_length;
  m.chain_length = 0;
  m.xor_key = a->xor_key;
  return list_add(&c->matches[a->sidx], &m, (s->flags & YRB_SF_GREEDY_REGEXP) != 0);
}

/* ------------------------------------------------------ fiber VM (a7) */

static int fiber_create(octx_t* c, int32_t* out) /* re.c:1214-1248 */
{
  if (c->free_head < 0)
    return YRB_ERROR_TOO_MANY_RE_FIBERS; /* pool exhausted: 1024 live fibers */

  int32_t f = c->free_head;
  c->free_head = c->fib[f].next;
  c->live++;
  if ((uint64_t) c->live > c->r->max_fibers)
    c->r->max_fibers = (uint64_t) c->live;

  c->fib[f].ip = 0;
  c->fib[f].sp = -1;
  c->fib[f].rc = -1;
  c->fib[f].next = -1;
  c->fib[f].prev = -1;
  *out = f;
  return YRB_SUCCESS;
}

static void fiber_release(octx_t* c, int32_t f)
{
  c->fib[f].next = c->free_head;
  c->free_head = f;
  c->live--;
}

/* re.c:1278-1315 _yr_re_fiber_exists: an identical fiber between head and last (inclusive) */
static int fiber_exists(octx_t* c, int32_t target, int32_t last)
{
  if (last < 0)
    return 0;

  const ofiber_t* t = &c->fib[target];
  int32_t stop = c->fib[last].next;

  for (int32_t f = c->head; f != stop; f = c->fib[f].next)
  {
    const ofiber_t* x = &c->fib[f];

    if (x->ip == t->ip && x->sp == t->sp && x->rc == t->rc)
    {
      int equal = 1;
      for (int32_t i = 0; i <= x->sp; i++)
        if (x->stack[i] != t->stack[i]) { equal = 0; break; }
      if (equal)
        return 1;
    }
  }
  return 0;
}

/* re.c:1327-1358 _yr_re_fiber_split: clone inserted right after the original */
static int fiber_split(octx_t* c, int32_t f, int32_t* out)
{
  int32_t n;
  int rc = fiber_create(c, &n);
  if (rc != YRB_SUCCESS)
    return rc;

  ofiber_t* F = &c->fib[f];
  ofiber_t* N = &c->fib[n];

  N->sp = F->sp;
  N->ip = F->ip;
  N->rc = F->rc;
  for (int32_t i = 0; i <= F->sp; i++) N->stack[i] = F->stack[i];

  N->next = F->next;
  N->prev = f;
  if (F->next >= 0) c->fib[F->next].prev = n;
  F->next = n;
  if (c->tail == f) c->tail = n;

  *out = n;
  return YRB_SUCCESS;
}

/* re.c:1364-1394 _yr_re_fiber_kill: returns the next fiber */
static int32_t fiber_kill(octx_t* c, int32_t f)
{
  int32_t next = c->fib[f].next;
  int32_t prev = c->fib[f].prev;

  if (prev >= 0) c->fib[prev].next = next;
  if (next >= 0) c->fib[next].prev = prev;
  if (c->tail == f) c->tail = prev;
  if (c->head == f) c->head = next;

  fiber_release(c, f);
  return next;
}

/* re.c:1399-1433 _yr_re_fiber_kill_tail / kill_all */
static void fiber_kill_tail(octx_t* c, int32_t f)
{
  int32_t prev = c->fib[f].prev;

  if (prev >= 0) c->fib[prev].next = -1;
  c->tail = prev;
  if (c->head == f) c->head = -1;

  while (f >= 0)
  {
    int32_t n = c->fib[f].next;
    fiber_release(c, f);
    f = n;
  }
}

static void fiber_kill_all(octx_t* c)
{
  if (c->head >= 0)
    fiber_kill_tail(c, c->head);
}

/* re.c:1441-1654 _yr_re_fiber_sync */
static int fiber_sync(octx_t* c, int32_t fiber_to_sync)
{
  uint8_t splits_executed[YRB_RE_MAX_SPLIT_ID];
  int splits_executed_count = 0;

  int32_t fiber = fiber_to_sync;
  int32_t last = c->fib[fiber_to_sync].next;
  int32_t branch_a, branch_b, next;
  int rc;

  while (fiber != last)
  {
    ofiber_t* F = &c->fib[fiber];
    const uint8_t* ip = c->code + F->ip;
    uint8_t opcode = *ip;

    switch (opcode)
    {
    case OP_SPLIT_A:
    case OP_SPLIT_B:
    {
      uint8_t split_id = ip[1];
      int already = 0;

      for (int i = 0; i < splits_executed_count; i++)
        if (splits_executed[i] == split_id) { already = 1; break; }

      if (already)
      {
        fiber = fiber_kill(c, fiber);
      }
      else
      {
        branch_a = fiber;
        rc = fiber_split(c, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;

        if (opcode == OP_SPLIT_B) { int32_t t = branch_a; branch_a = branch_b; branch_b = t; }

        c->fib[branch_a].ip += 4;
        c->fib[branch_b].ip += rd_i16(c->code + c->fib[branch_b].ip + 2);

        if (splits_executed_count < YRB_RE_MAX_SPLIT_ID)
          splits_executed[splits_executed_count++] = split_id;
      }
      break;
    }

    case OP_REPEAT_START_GREEDY:
    case OP_REPEAT_START_UNGREEDY:
    {
      uint16_t rmin = rd_u16(ip + 1);
      int32_t roff = rd_i32(ip + 5);
      branch_a = fiber;

      if (rmin == 0)
      {
        rc = fiber_split(c, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;
        if (opcode == OP_REPEAT_START_UNGREEDY) { int32_t t = branch_a; branch_a = branch_b; branch_b = t; }
        c->fib[branch_b].ip += roff;
      }

      ofiber_t* A = &c->fib[branch_a];
      A->sp++;
      if (A->sp < c->stack_words) A->stack[A->sp] = 0;
      A->ip += 9;
      break;
    }

    case OP_REPEAT_END_GREEDY:
    case OP_REPEAT_END_UNGREEDY:
    {
      uint16_t rmin = rd_u16(ip + 1);
      uint16_t rmax = rd_u16(ip + 3);
      int32_t roff = rd_i32(ip + 5);

      if (F->sp < 0 || F->sp >= c->stack_words)
        return YRB_ERROR_INTERNAL_FATAL; /* reference would touch stack[-1] */

      F->stack[F->sp]++;

      if (F->stack[F->sp] < rmin)
      {
        F->ip += roff;
        break;
      }

      branch_a = fiber;

      if (F->stack[F->sp] < rmax)
      {
        rc = fiber_split(c, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;
        if (opcode == OP_REPEAT_END_GREEDY) { int32_t t = branch_a; branch_a = branch_b; branch_b = t; }
        c->fib[branch_b].ip += roff;
      }

      c->fib[branch_a].sp--;
      c->fib[branch_a].ip += 9;
      break;
    }

    case OP_REPEAT_ANY_GREEDY:
    case OP_REPEAT_ANY_UNGREEDY:
    {
      uint16_t rmin = rd_u16(ip + 1);
      uint16_t rmax = rd_u16(ip + 3);

      if (F->rc == -1)
        F->rc = 0;

      if (F->rc < rmin)
      {
        F->rc++;
        fiber = F->next;
      }
      else if (F->rc < rmax)
      {
        next = F->next;
        branch_a = fiber;
        rc = fiber_split(c, branch_a, &branch_b);
        if (rc != YRB_SUCCESS) return rc;
        if (opcode == OP_REPEAT_ANY_UNGREEDY) { int32_t t = branch_a; branch_a = branch_b; branch_b = t; }

        c->fib[branch_a].rc++;
        c->fib[branch_b].ip += 5;
        c->fib[branch_b].rc = -1;

        rc = fiber_sync(c, branch_b);
        if (rc != YRB_SUCCESS) return rc;

        fiber = next;
      }
      else
      {
        F->ip += 5;
        F->rc = -1;
      }
      break;
    }

    case OP_JUMP:
      F->ip += rd_i16(ip + 1);
      break;

    default:
      fiber = F->next;
    }
  }

  return YRB_SUCCESS;
}

/* re.c:98-122 */
static int char_in_class(const octx_t* c, const uint8_t* cls, uint8_t chr, int nocase)
{
  const uint8_t* bitmap = cls + 1;
  int result = bitmap[chr / 8] & (1 << (chr % 8));
  if (nocase)
    result |= bitmap[c->alter[chr] / 8] & (1 << (c->alter[chr] % 8));
  if (cls[0])
    result = !result;
  return result != 0;
}

static int is_word_char(const uint8_t* input, int character_size)
{
  int result = is_alnum(*input) || *input == '_';
  if (character_size == 2)
    result = result && (input[1] == 0);
  return result;
}

/* re.c:1693-2068 yr_re_exec.  `pos` = index of input_data in c->data;
   fwd/bwd sizes as in the reference.  cb != NULL <=> callback given. */
static int re_exec(
    octx_t* c, uint32_t code_off, int64_t pos, size_t input_forwards_size, size_t input_backwards_size,
    int flags, const cbargs_t* cb, int* matches)
{
  const uint8_t* input_data = c->data + pos;
  const uint8_t* input;
  int character_size = (flags & RE_FLAGS_WIDE) ? 2 : 1;
  int input_incr = character_size;
  int bytes_matched, max_bytes_matched;
  int32_t fiber, next_fiber;
  int rc;

  if (matches) *matches = -1;

  input = input_data;

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

  c->head = c->tail = -1;
  rc = fiber_create(c, &fiber);
  if (rc != YRB_SUCCESS) return rc;

  c->fib[fiber].ip = code_off;
  c->head = c->tail = fiber;

  rc = fiber_sync(c, fiber);
  if (rc != YRB_SUCCESS) { fiber_kill_all(c); return rc; }

  while (c->head >= 0)
  {
    fiber = c->head;

    while (fiber >= 0)
    {
      next_fiber = c->fib[fiber].next;
      if (fiber_exists(c, fiber, c->fib[fiber].prev))
        fiber_kill(c, fiber);
      fiber = next_fiber;
    }

    fiber = c->head;

    while (fiber >= 0)
    {
      ofiber_t* F = &c->fib[fiber];
      const uint8_t* ip = c->code + F->ip;
      int action = 0; /* 0 none, 1 continue, 2 kill, 3 kill tail */
      int match, kill;

#define PROLOG                                                                     \
  if ((bytes_matched >= max_bytes_matched) || (character_size == 2 && input[1] != 0)) \
  {                                                                                \
    action = 2;                                                                    \
    break;                                                                         \
  }

      switch (*ip)
      {
      case OP_ANY:
        PROLOG;
        match = (flags & RE_FLAGS_DOT_ALL) || (*input != 0x0A);
        action = match ? 0 : 2;
        F->ip += 1;
        break;

      case OP_REPEAT_ANY_GREEDY:
      case OP_REPEAT_ANY_UNGREEDY:
        PROLOG;
        match = (flags & RE_FLAGS_DOT_ALL) || (*input != 0x0A);
        action = match ? 0 : 2;
        break;

      case OP_LITERAL:
        PROLOG;
        if (flags & RE_FLAGS_NO_CASE)
          match = c->lower[*input] == c->lower[ip[1]];
        else
          match = (*input == ip[1]);
        action = match ? 0 : 2;
        F->ip += 2;
        break;

      case OP_NOT_LITERAL:
        PROLOG;
        match = (*input != ip[1]);
        action = match ? 0 : 2;
        F->ip += 2;
        break;

      case OP_MASKED_LITERAL:
      {
        PROLOG;
        uint16_t args = rd_u16(ip + 1);
        match = ((*input & (args >> 8)) == (args & 0xFF));
        action = match ? 0 : 2;
        F->ip += 3;
        break;
      }

      case OP_MASKED_NOT_LITERAL:
      {
        PROLOG;
        uint16_t args = rd_u16(ip + 1);
        match = ((*input & (args >> 8)) != (args & 0xFF));
        action = match ? 0 : 2;
        F->ip += 3;
        break;
      }

      case OP_CLASS:
        PROLOG;
        match = char_in_class(c, ip + 1, *input, flags & RE_FLAGS_NO_CASE);
        action = match ? 0 : 2;
        F->ip += 34;
        break;

      case OP_WORD_CHAR:
        PROLOG;
        match = is_word_char(input, character_size);
        action = match ? 0 : 2;
        F->ip += 1;
        break;

      case OP_NON_WORD_CHAR:
        PROLOG;
        match = !is_word_char(input, character_size);
        action = match ? 0 : 2;
        F->ip += 1;
        break;

      case OP_SPACE:
      case OP_NON_SPACE:
        PROLOG;
        switch (*input)
        {
        case ' ': case '\t': case '\r': case '\n': case '\v': case '\f':
          match = 1;
          break;
        default:
          match = 0;
        }
        if (*ip == OP_NON_SPACE) match = !match;
        action = match ? 0 : 2;
        F->ip += 1;
        break;

      case OP_DIGIT:
        PROLOG;
        match = (*input >= '0' && *input <= '9'); /* isdigit(), C locale */
        action = match ? 0 : 2;
        F->ip += 1;
        break;

      case OP_NON_DIGIT:
        PROLOG;
        match = !(*input >= '0' && *input <= '9');
        action = match ? 0 : 2;
        F->ip += 1;
        break;

      case OP_WORD_BOUNDARY:
      case OP_NON_WORD_BOUNDARY:
        if (bytes_matched == 0 && input_backwards_size < (size_t) character_size)
          match = 1;
        else if (bytes_matched >= max_bytes_matched)
          match = 1;
        else
          match = is_word_char(input, character_size) != is_word_char(input - input_incr, character_size);

        if (*ip == OP_NON_WORD_BOUNDARY) match = !match;
        action = match ? 1 : 2;
        F->ip += 1;
        break;

      case OP_MATCH_AT_START:
        if (flags & RE_FLAGS_BACKWARDS)
          kill = input_backwards_size > (size_t) bytes_matched;
        else
          kill = input_backwards_size > 0 || (bytes_matched != 0);
        action = kill ? 2 : 1;
        F->ip += 1;
        break;

      case OP_MATCH_AT_END:
        kill = (flags & RE_FLAGS_BACKWARDS) || input_forwards_size > (size_t) bytes_matched;
        action = kill ? 2 : 1;
        F->ip += 1;
        break;

      case OP_MATCH:
        if (matches) *matches = bytes_matched;

        if (flags & RE_FLAGS_EXHAUSTIVE)
        {
          if (cb != NULL)
          {
            int64_t mpos = (flags & RE_FLAGS_BACKWARDS) ? (int64_t) (input + character_size - c->data) : pos;
            rc = match_callback(c, mpos, bytes_matched, flags, cb);
            if (rc != YRB_SUCCESS) { fiber_kill_all(c); return rc; }
          }
          action = 2;
        }
        else
        {
          action = 3;
        }
        break;

      default:
        fiber_kill_all(c);
        return YRB_ERROR_INTERNAL_FATAL;
      }
#undef PROLOG

      switch (action)
      {
      case 2:
        fiber = fiber_kill(c, fiber);
        break;
      case 3:
        fiber_kill_tail(c, fiber);
        fiber = -1;
        break;
      case 1:
        rc = fiber_sync(c, fiber);
        if (rc != YRB_SUCCESS) { fiber_kill_all(c); return rc; }
        break;
      default:
        next_fiber = c->fib[fiber].next;
        rc = fiber_sync(c, fiber);
        if (rc != YRB_SUCCESS) { fiber_kill_all(c); return rc; }
        fiber = next_fiber;
      }
    }

    input += input_incr;
    bytes_matched += character_size;
  }

  return YRB_SUCCESS;
}

/* ------------------------------------------------------ fast hex VM (a8) */

typedef struct
{
  int64_t input; /* index into c->data (may be -1 in backward mode) */
  int32_t round;
  int32_t prev, next;
} opos_t;

typedef struct
{
  opos_t* v;
  int32_t n, cap;
  int32_t free_head;
  int32_t live;
} opool_t;

static int32_t pos_create(opool_t* P)
{
  int32_t i;
  if (P->free_head >= 0)
  {
    i = P->free_head;
    P->free_head = P->v[i].next;
  }
  else
  {
    if (P->n == P->cap)
    {
      P->cap = P->cap ? P->cap * 2 : 64;
      P->v = (opos_t*) realloc(P->v, sizeof(opos_t) * (size_t) P->cap);
    }
    i = P->n++;
  }
  P->live++;
  P->v[i].input = 0;
  P->v[i].round = 0;
  P->v[i].next = P->v[i].prev = -1;
  return i;
}

static void pos_release(opool_t* P, int32_t i)
{
  P->v[i].next = P->free_head;
  P->free_head = i;
  P->live--;
}

/* re.c:2150-2477 yr_re_fast_exec */
static int re_fast_exec(
    octx_t* c, uint32_t code_off, int64_t pos, size_t input_forwards_size, size_t input_backwards_size,
    int flags, const cbargs_t* cb, int* matches)
{
  opool_t P;
  memset(&P, 0, sizeof(P));
  P.free_head = -1;

  int input_incr = (flags & RE_FLAGS_BACKWARDS) ? -1 : 1;
  int bytes_matched;
  int max_bytes_matched;
  int rc = YRB_SUCCESS;

  if (flags & RE_FLAGS_BACKWARDS)
    max_bytes_matched = (int) (input_backwards_size < YRB_RE_SCAN_LIMIT ? input_backwards_size : YRB_RE_SCAN_LIMIT);
  else
    max_bytes_matched = (int) (input_forwards_size < YRB_RE_SCAN_LIMIT ? input_forwards_size : YRB_RE_SCAN_LIMIT);

  const uint8_t* ip = c->code + code_off;

  int32_t first = pos_create(&P);
  P.v[first].input = pos;
  if (flags & RE_FLAGS_BACKWARDS)
    P.v[first].input--;

  int32_t last = first;
  int round = 0;

  if (matches) *matches = -1;

  while (first >= 0)
  {
    int32_t current = first;

    while (current >= 0)
    {
      int32_t next = P.v[current].next;

      if (P.v[current].round != round)
      {
        current = next;
        continue;
      }

      int64_t cin = P.v[current].input;

      bytes_matched = (flags & RE_FLAGS_BACKWARDS) ? (int) (pos - cin - 1) : (int) (cin - pos);

      int match = 0;

      switch (*ip)
      {
      case OP_ANY:
        if (bytes_matched >= max_bytes_matched) break;
        match = 1;
        P.v[current].input += input_incr;
        break;

      case OP_LITERAL:
        if (bytes_matched >= max_bytes_matched) break;
        if (c->data[cin] == ip[1]) { match = 1; P.v[current].input += input_incr; }
        break;

      case OP_NOT_LITERAL:
        if (bytes_matched >= max_bytes_matched) break;
        if (c->data[cin] != ip[1]) { match = 1; P.v[current].input += input_incr; }
        break;

      case OP_MASKED_LITERAL:
      {
        if (bytes_matched >= max_bytes_matched) break;
        uint16_t args = rd_u16(ip + 1);
        if ((c->data[cin] & (args >> 8)) == (args & 0xFF)) { match = 1; P.v[current].input += input_incr; }
        break;
      }

      case OP_MASKED_NOT_LITERAL:
      {
        if (bytes_matched >= max_bytes_matched) break;
        uint16_t args = rd_u16(ip + 1);
        if ((c->data[cin] & (args >> 8)) != (args & 0xFF)) { match = 1; P.v[current].input += input_incr; }
        break;
      }

      case OP_REPEAT_ANY_UNGREEDY:
      {
        if (bytes_matched >= max_bytes_matched) break;

        int rmin = rd_u16(ip + 1);
        int rmax = rd_u16(ip + 3);
        const uint8_t* next_opcode = ip + 5;
        int32_t insertion_point = current;

        match = 1;

        for (int j = rmin + 1; j <= rmax; j++)
        {
          if (bytes_matched + j >= max_bytes_matched)
            break;

          int64_t next_input = cin + (int64_t) j * input_incr;

          while (P.v[insertion_point].next >= 0 && next_input >= P.v[P.v[insertion_point].next].input)
            insertion_point = P.v[insertion_point].next;

          if (next_input == P.v[insertion_point].input)
            continue;

          if (*next_opcode == OP_LITERAL && next_opcode[1] != c->data[next_input])
            continue;

          int32_t ni = pos_create(&P);
          P.v[ni].input = next_input;
          P.v[ni].round = round + 1;
          P.v[ni].prev = insertion_point;
          P.v[ni].next = P.v[insertion_point].next;
          P.v[insertion_point].next = ni;
          if (P.v[ni].next >= 0) P.v[P.v[ni].next].prev = ni;
          if (insertion_point == last) last = ni;
        }

        if ((uint64_t) P.live > c->r->max_positions)
          c->r->max_positions = (uint64_t) P.live;

        P.v[current].input += (int64_t) input_incr * rmin;
        break;
      }

      case OP_MATCH:
        if (flags & RE_FLAGS_EXHAUSTIVE)
        {
          int64_t mpos;
          if (flags & RE_FLAGS_BACKWARDS)
          {
            int64_t lo = pos - (int64_t) input_backwards_size;
            mpos = cin + 1 > lo ? cin + 1 : lo;
          }
          else
          {
            mpos = pos;
          }
          int ml = bytes_matched < max_bytes_matched ? bytes_matched : max_bytes_matched;
          rc = match_callback(c, mpos, ml, flags, cb);
          if (rc != YRB_SUCCESS) goto _done;
        }
        else
        {
          if (matches) *matches = bytes_matched;
          goto _done;
        }
        break;

      default:
        rc = YRB_ERROR_INTERNAL_FATAL;
        goto _done;
      }

      if (match)
      {
        P.v[current].round = round + 1;
      }
      else
      {
        if (current == first) first = P.v[current].next;
        if (current == last) last = P.v[current].prev;
        if (P.v[current].prev >= 0) P.v[P.v[current].prev].next = P.v[current].next;
        if (P.v[current].next >= 0) P.v[P.v[current].next].prev = P.v[current].prev;
        pos_release(&P, current);
      }

      current = next;
    }

    switch (*ip)
    {
    case OP_ANY: ip += 1; break;
    case OP_LITERAL: case OP_NOT_LITERAL: ip += 2; break;
    case OP_MASKED_LITERAL: case OP_MASKED_NOT_LITERAL: ip += 3; break;
    case OP_REPEAT_ANY_UNGREEDY: ip += 5; break;
    default: break;
    }

    round++;
  }

_done:
  free(P.v);
  return rc;
}

/* --------------------------------------------------- verification (a4-a6) */

/* scan.c:778-885 _yr_scan_verify_re_match */
static int verify_re(octx_t* c, const yrb_acm* m, size_t offset)
{
  const yrb_string* s = &c->S[m->string_idx];
  int forward_matches = -1, backward_matches = -1;
  int flags = 0, rc;
  int fast = (s->flags & YRB_SF_FAST_REGEXP) != 0;

  if (s->flags & YRB_SF_GREEDY_REGEXP) flags |= RE_FLAGS_GREEDY;
  if (s->flags & YRB_SF_NO_CASE) flags |= RE_FLAGS_NO_CASE;
  if (s->flags & YRB_SF_DOT_ALL) flags |= RE_FLAGS_DOT_ALL;

#define EXEC(code, fl, cb, mt)                                                                  \
  (fast ? re_fast_exec(c, (code), (int64_t) offset, c->size - offset, offset, (fl), (cb), (mt)) \
        : re_exec(c, (code), (int64_t) offset, c->size - offset, offset, (fl), (cb), (mt)))

  if (s->flags & (YRB_SF_ASCII | YRB_SF_BASE64 | YRB_SF_BASE64_WIDE))
  {
    rc = EXEC(m->fwd_code, flags, NULL, &forward_matches);
    if (rc != YRB_SUCCESS) return rc;
  }

  if (forward_matches == -1 && (s->flags & YRB_SF_WIDE) && !(s->flags & (YRB_SF_BASE64 | YRB_SF_BASE64_WIDE)))
  {
    flags |= RE_FLAGS_WIDE;
    rc = EXEC(m->fwd_code, flags, NULL, &forward_matches);
    if (rc != YRB_SUCCESS) return rc;
  }

  if (forward_matches == -1)
    return YRB_SUCCESS;

  if (forward_matches == 0 && m->bwd_code == YRB_NONE)
    return YRB_SUCCESS;

  cbargs_t a;
  a.sidx = m->string_idx;
  a.forward_matches = forward_matches;
  a.full_word = (s->flags & YRB_SF_FULL_WORD) != 0;
  a.xor_key = 0;

  if (m->bwd_code != YRB_NONE)
    return EXEC(m->bwd_code, flags | RE_FLAGS_BACKWARDS | RE_FLAGS_EXHAUSTIVE, &a, &backward_matches);

  return match_callback(c, (int64_t) offset, 0, flags, &a);
#undef EXEC
}

/* scan.c:62-255 compare family; scan.c:887-990 _yr_scan_verify_literal_match */
static int verify_literal(octx_t* c, const yrb_acm* m, size_t offset)
{
  const yrb_string* s = &c->S[m->string_idx];
  const uint8_t* str = c->lit + s->lit_off;
  const uint8_t* d = c->data + offset;
  size_t avail = c->size - offset;
  size_t len = (size_t) s->length;
  int forward_matches = 0, flags = 0;
  uint8_t xor_key = 0;
  size_t i;

  if (s->flags & YRB_SF_FITS_IN_ATOM)
  {
    forward_matches = (int) m->backtrack;
  }
  else if (s->flags & YRB_SF_NO_CASE)
  {
    if (s->flags & YRB_SF_ASCII)
    {
      if (avail >= len)
      {
        for (i = 0; i < len && c->lower[d[i]] == c->lower[str[i]]; i++) {}
        forward_matches = (i == len) ? (int) i : 0;
      }
    }
    if ((s->flags & YRB_SF_WIDE) && forward_matches == 0)
    {
      if (avail >= len * 2)
      {
        for (i = 0; i < len && c->lower[d[2 * i]] == c->lower[str[i]] && d[2 * i + 1] == 0; i++) {}
        forward_matches = (i == len) ? (int) (i * 2) : 0;
      }
    }
  }
  else
  {
    if (s->flags & YRB_SF_ASCII)
    {
      if (avail >= len)
      {
        for (i = 0; i < len && d[i] == str[i]; i++) {}
        forward_matches = (i == len) ? (int) i : 0;
      }
    }
    if ((s->flags & YRB_SF_WIDE) && forward_matches == 0)
    {
      if (avail >= len * 2)
      {
        for (i = 0; i < len && d[2 * i] == str[i] && d[2 * i + 1] == 0; i++) {}
        forward_matches = (i == len) ? (int) (i * 2) : 0;
      }
    }
    if ((s->flags & YRB_SF_XOR) && forward_matches == 0)
    {
      if (s->flags & YRB_SF_WIDE)
      {
        if (avail >= len * 2)
        {
          uint8_t k = d[0] ^ str[0];
          for (i = 0; i < len && d[2 * i] == (str[i] ^ k) && (d[2 * i + 1] ^ k) == 0; i++) {}
          forward_matches = (i == len) ? (int) (i * 2) : 0;
          if (forward_matches > 0) xor_key = k;
        }
      }
      if (forward_matches == 0)
      {
        if (avail >= len)
        {
          uint8_t k = d[0] ^ str[0];
          for (i = 0; i < len && d[i] == (str[i] ^ k); i++) {}
          forward_matches = (i == len) ? (int) i : 0;
          if (forward_matches > 0) xor_key = k;
        }
      }
    }
  }

  if (forward_matches == 0)
    return YRB_SUCCESS;

  if (forward_matches == s->length * 2)
    flags |= RE_FLAGS_WIDE;
  if (s->flags & YRB_SF_NO_CASE)
    flags |= RE_FLAGS_NO_CASE;

  cbargs_t a;
  a.sidx = m->string_idx;
  a.forward_matches = forward_matches;
  a.full_word = (s->flags & YRB_SF_FULL_WORD) != 0;
  a.xor_key = xor_key;

  return match_callback(c, (int64_t) offset, 0, flags, &a);
}

/* scan.c:992-1089 yr_scan_verify_match */
static int verify_match(octx_t* c, uint32_t acm_idx, size_t offset)
{
  const yrb_acm* m = &c->acm[acm_idx];
  const yrb_string* s = &c->S[m->string_idx];
  int rc;

  c->r->n_candidates++;

  if (c->size - offset <= 0) /* size_t arithmetic: true only when offset == size */
    return YRB_SUCCESS;

  if (c->disabled[m->string_idx])
    return YRB_SUCCESS;

  if ((c->scan_flags & YRB_SCAN_FAST_MODE) && (s->flags & YRB_SF_SINGLE_MATCH) && c->matches[m->string_idx].n > 0)
    return YRB_SUCCESS;

  if ((s->flags & YRB_SF_FIXED_OFFSET) && s->fixed_offset != (int64_t) (c->base + offset))
    return YRB_SUCCESS;

  if (s->flags & YRB_SF_LITERAL)
    rc = verify_literal(c, m, offset);
  else
    rc = verify_re(c, m, offset);

  if (rc == YRB_ERROR_TOO_MANY_MATCHES)
  {
    yro_result* r = c->r;
    r->too_many = (uint32_t*) realloc(r->too_many, sizeof(uint32_t) * (r->n_too_many + 1));
    r->too_many[r->n_too_many++] = m->string_idx;

    if (c->too_many_action == 0)
    {
      c->disabled[m->string_idx] = 1;
      rc = YRB_SUCCESS;
    }
  }

  if (rc != YRB_SUCCESS)
    c->r->error_string = m->string_idx;

  return rc;
}

/* ------------------------------------------------------------ AC walk (a1) */

typedef int (*event_fn)(void* user, uint64_t i, uint32_t acm_idx);

/* scanner.c:45-176 _yr_scanner_scan_mem_block, minus the timeout check */
static int ac_walk(
    const uint32_t* T, const uint32_t* M, const yrb_acm* acm, const uint8_t* data, size_t size, event_fn fn,
    void* user)
{
  size_t i = 0;
  uint32_t state = 0;
  int rc;

  for (;;)
  {
    if (M[state] != 0)
    {
      uint32_t k = M[state];
      while (k != 0)
      {
        const yrb_acm* m = &acm[k - 1];
        if (m->backtrack <= i)
        {
          rc = fn(user, (uint64_t) i, k - 1);
          if (rc != YRB_SUCCESS) return rc;
        }
        k = m->next;
      }
    }

    if (i >= size)
      break;

    uint32_t index = (uint32_t) data[i++] + 1;
    uint32_t t = T[state + index];

    while ((t & 0x1FF) != index)
    {
      if (state != 0)
      {
        state = T[state] >> 9;
        t = T[state + index];
      }
      else
      {
        t = 0;
        break;
      }
    }

    state = t >> 9;
  }

  return YRB_SUCCESS;
}

static int scan_event(void* user, uint64_t i, uint32_t acm_idx)
{
  octx_t* c = (octx_t*) user;
  return verify_match(c, acm_idx, (size_t) (i - c->acm[acm_idx].backtrack));
}

/* -------------------------------------------------------------- public API */

static void ctx_bind_image(octx_t* c, const void* image)
{
  c->h = (const yrb_image_header*) image;
  c->T = yrb_img_trans(image);
  c->M = yrb_img_match(image);
  c->acm = yrb_img_acm(image);
  c->S = yrb_img_strings(image);
  c->lit = yrb_img_lit(image);
  c->code = yrb_img_code(image);
  c->lower = yrb_img_lower(image);
  c->alter = yrb_img_alter(image);
}

static size_t dump_lists(const olist_t* L, uint32_t n, yrb_match** out, uint8_t flag)
{
  size_t total = 0, k = 0;
  for (uint32_t i = 0; i < n; i++) total += (size_t) L[i].n;

  yrb_match* v = (yrb_match*) calloc(total ? total : 1, sizeof(yrb_match));

  for (uint32_t i = 0; i < n; i++)
  {
    for (int32_t j = 0; j < L[i].n; j++)
    {
      v[k].offset = L[i].v[j].offset;
      v[k].string_idx = i;
      v[k].match_length = L[i].v[j].length;
      v[k].xor_key = L[i].v[j].xor_key;
      v[k].flags = flag;
      v[k].chain_length = L[i].v[j].chain_length;
      k++;
    }
  }

  *out = v;
  return total;
}

yro_result* yro_scan_block(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, int too_many_action)
{
  octx_t c;
  memset(&c, 0, sizeof(c));
  ctx_bind_image(&c, image);

  if (c.h->magic != YRB_IMAGE_MAGIC || c.h->version != YRB_IMAGE_VERSION)
    return NULL;

  uint32_t ns = c.h->num_strings;

  c.data = data;
  c.size = size;
  c.base = base;
  c.scan_flags = scan_flags;
  c.too_many_action = too_many_action;
  c.matches = (olist_t*) calloc(ns ? ns : 1, sizeof(olist_t));
  c.unconf = (olist_t*) calloc(ns ? ns : 1, sizeof(olist_t));
  c.disabled = (uint8_t*) calloc(ns ? ns : 1, 1);

  c.stack_words = (int32_t) c.h->max_re_stack + 1;
  c.fib = (ofiber_t*) calloc(YRB_RE_MAX_FIBERS, sizeof(ofiber_t));
  c.fib_stacks = (uint16_t*) calloc((size_t) YRB_RE_MAX_FIBERS * (size_t) c.stack_words, sizeof(uint16_t));

  for (int32_t f = 0; f < YRB_RE_MAX_FIBERS; f++)
  {
    c.fib[f].stack = c.fib_stacks + (size_t) f * (size_t) c.stack_words;
    c.fib[f].next = (f + 1 < YRB_RE_MAX_FIBERS) ? f + 1 : -1;
  }
  c.free_head = 0;
  c.head = c.tail = -1;

  yro_result* r = (yro_result*) calloc(1, sizeof(yro_result));
  r->error_string = YRB_NONE;
  c.r = r;

  r->error = ac_walk(c.T, c.M, c.acm, data, size, scan_event, &c);

  r->n_matches = dump_lists(c.matches, ns, &r->matches, 0);
  r->n_unconfirmed = dump_lists(c.unconf, ns, &r->unconfirmed, YRB_MATCH_UNCONFIRMED);

  for (uint32_t i = 0; i < ns; i++)
  {
    free(c.matches[i].v);
    free(c.unconf[i].v);
  }
  free(c.matches);
  free(c.unconf);
  free(c.disabled);
  free(c.fib);
  free(c.fib_stacks);

  return r;
}

void yro_result_free(yro_result* r)
{
  if (r == NULL) return;
  free(r->matches);
  free(r->unconfirmed);
  free(r->too_many);
  free(r);
}

typedef struct
{
  uint64_t* pos;
  uint32_t* acm;
  size_t cap, n;
} evsink_t;

static int sink_event(void* user, uint64_t i, uint32_t acm_idx)
{
  evsink_t* s = (evsink_t*) user;
  if (s->n < s->cap)
  {
    s->pos[s->n] = i;
    s->acm[s->n] = acm_idx;
  }
  s->n++;
  return YRB_SUCCESS;
}

size_t yro_ac_events(
    const void* image, const uint8_t* data, size_t size, uint64_t* ev_pos, uint32_t* ev_acm, size_t cap)
{
  evsink_t s = {ev_pos, ev_acm, cap, 0};
  ac_walk(yrb_img_trans(image), yrb_img_match(image), yrb_img_acm(image), data, size, sink_event, &s);
  return s.n;
}
