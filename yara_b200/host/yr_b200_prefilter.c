/*
 * yr_b200_prefilter.c -- condition-VM pre-filter: the step AFTER the
 * byte-scanning path.
 *
 * Once the scan itself costs microseconds, what is left of a small-file scan is
 * yr_execute_code (exec.c:418): it runs every rule's condition for every file, and for a
 * 100,000-string / 2,000-rule set that is ~1 ms of pushing 50 string pointers per rule
 * only to learn that none of them matched (measured: 1.45 ms per yr_scanner_scan_mem of a
 * 16-byte buffer, 0.39 ms with the conditions short-circuited).
 *
 * A rule can be skipped -- OP_INIT_RULE takes its jump exactly as it does for a disabled
 * rule (exec.c:1154-1175), leaving rule_matches_flags untouched -- when
 *   (1) its condition is a pure function of the match state of its own strings: the
 *       bytecode between OP_INIT_RULE and OP_MATCH_RULE contains no opcode that reads the
 *       scanned data, filesize/entrypoint, module objects, external variables or other
 *       rules (classified once per YR_RULES by decoding the instruction stream);
 *   (2) it is not GLOBAL (a false global rule has a side effect, exec.c:1193);
 *   (3) none of its strings has a match in this scan (a per-scanner byte per rule, set where
 *       matches are created: materialise() in the GPU glue, _yr_scan_match_callback in the
 *       CPU validation build);
 *   (4) it is known to be FALSE in that situation.  That is not derived, it is observed:
 *       the first time such a rule is evaluated with none of its strings matched, the value
 *       OP_MATCH_RULE pops is remembered (a pure rule has only one value there).
 *
 * On by default; YR_B200_RULE_PREFILTER=0 turns it off (the hooks then cost one
 * thread-local compare per rule).  No GPU code here; the same file is compiled into
 * libyara_b200.so and into oracle/_ref/libyara_refpf.so (reference + these hooks only), which
 * runs the reference's own test programs with the filter on.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <yara/exec.h>
#include <yara/rules.h>
#include <yara/types.h>

#include "yr_b200.h"

enum
{
  PF_IMPURE = 0,    /* never skipped */
  PF_UNKNOWN = 1,   /* pure; value with no string matched not seen yet */
  PF_SKIPPABLE = 2, /* pure; false when none of its strings matched */
  PF_TRUE = 3       /* pure; true when none of its strings matched (e.g. `not $a`) */
};

typedef struct pf_rules
{
  YR_RULES* rules;
  uint32_t n;
  uint32_t walked; /* rules whose code the classifier reached (== n unless the code layout surprised it) */
  uint8_t* state;
  struct pf_rules* next;
} pf_rules;

typedef struct pf_scan
{
  YR_SCAN_CONTEXT* ctx;
  uint32_t n;
  uint8_t* hit; /* per rule: some string of it got a match in the current scan */
  struct pf_scan* next;
} pf_scan;

static pthread_mutex_t g_pf_lock = PTHREAD_MUTEX_INITIALIZER;
static pf_rules* g_pf_rules = NULL;
static pf_scan* g_pf_scans = NULL;
static uint64_t g_pf_evaluated = 0, g_pf_skipped = 0;

static __thread YR_SCAN_CONTEXT* t_exec_ctx = NULL; /* set between begin() and end() */
static __thread uint8_t* t_exec_state = NULL;
static __thread uint8_t* t_exec_hit = NULL;
static __thread uint64_t t_evaluated = 0, t_skipped = 0;
static __thread YR_SCAN_CONTEXT* t_note_ctx = NULL; /* one-entry cache of the registry lookup */
static __thread pf_scan* t_note_scan = NULL;

int yr_b200_prefilter_enabled(void)
{
  static int on = -1;
  if (on < 0)
  {
    const char* e = getenv("YR_B200_RULE_PREFILTER"); /* on unless set to 0 */
    on = (e != NULL && e[0] == '0') ? 0 : 1;
  }
  return on;
}

/* Operand bytes of the opcodes a pure condition may contain (exec.c, each case's `ip +=`);
   -1: the opcode makes the rule impure (or is unknown). */
static int pure_operand_size(uint8_t op)
{
  switch (op)
  {
  /* stack and logic */
  case OP_NOP: case OP_AND: case OP_OR: case OP_NOT: case OP_POP: case OP_PUSH_U: case OP_DEFINED:
  case OP_BITWISE_NOT: case OP_BITWISE_AND: case OP_BITWISE_OR: case OP_BITWISE_XOR: case OP_SHL: case OP_SHR:
  case OP_MOD:
  /* match state of the rule's own strings */
  case OP_COUNT: case OP_COUNT_IN: case OP_LENGTH: case OP_OFFSET: case OP_FOUND: case OP_FOUND_AT: case OP_FOUND_IN:
  case OP_OF_FOUND_IN:
  /* loops over integer ranges / enumerations and string sets */
  case OP_ITER_NEXT: case OP_ITER_START_INT_RANGE: case OP_ITER_START_INT_ENUM: case OP_ITER_START_STRING_SET:
  case OP_ITER_CONDITION: case OP_ITER_END:
    return 0;
  case OP_PUSH_8:
    return 1;
  case OP_PUSH_16:
    return 2;
  case OP_PUSH_32:
    return 4;
  case OP_JFALSE: case OP_JFALSE_P: case OP_JTRUE: case OP_JTRUE_P: case OP_JNUNDEF: case OP_JUNDEF_P: case OP_JL_P:
  case OP_JLE_P: case OP_JZ: case OP_JZ_P:
    return 4;
  case OP_PUSH: case OP_INT_TO_DBL: case OP_OF: case OP_OF_PERCENT: case OP_INCR_M: case OP_CLEAR_M: case OP_ADD_M:
  case OP_POP_M: case OP_PUSH_M: case OP_SET_M: case OP_SWAPUNDEF:
    return 8;
  default:
    break;
  }
  if (op >= OP_INT_BEGIN && op <= OP_INT_END) return 0;
  if (op >= OP_DBL_BEGIN && op <= OP_DBL_END) return 0;
  return -1;
}

/* Walks the condition code: OP_IMPORT* then per rule OP_INIT_RULE <jump:i32> <rule:u32> ... OP_MATCH_RULE
   <rule:u64>, finally OP_HALT (parser.c / grammar.y emit it in this shape). */
static uint32_t classify(YR_RULES* rules, uint8_t* state)
{
  const uint8_t* ip = rules->code_start;
  uint32_t walked = 0;
  memset(state, PF_IMPURE, rules->num_rules);

  if (ip == NULL)
    return 0;

  for (;;)
  {
    const uint8_t op = *ip;

    if (op == OP_IMPORT)
    {
      ip += 1 + sizeof(uint64_t);
      continue;
    }
    if (op != OP_INIT_RULE)
      return walked; /* OP_HALT, or a layout this code does not know: everything left stays impure */

    int32_t off;
    uint32_t idx;
    memcpy(&off, ip + 1, sizeof(off));
    memcpy(&idx, ip + 1 + sizeof(int32_t), sizeof(idx));
    if (off <= 0 || idx >= rules->num_rules)
      return walked;

    const uint8_t* end = ip + off; /* the jump is relative to the opcode (exec.c:159-186) */
    const uint8_t* p = ip + 1 + sizeof(int32_t) + sizeof(uint32_t);
    int pure = !RULE_IS_GLOBAL(&rules->rules_table[idx]);

    while (pure && p < end)
    {
      if (*p == OP_MATCH_RULE)
      {
        uint64_t r;
        memcpy(&r, p + 1, sizeof(r));
        p += 1 + sizeof(uint64_t);
        pure = (p == end && r == idx);
        break;
      }
      const int sz = pure_operand_size(*p);
      if (sz < 0)
        pure = 0;
      else
        p += 1 + sz;
    }
    if (pure && p != end)
      pure = 0;

    state[idx] = pure ? PF_UNKNOWN : PF_IMPURE;
    walked++;
    ip = end;
  }
}

static pf_rules* rules_entry_for(YR_RULES* rules)
{
  pf_rules* e;
  for (e = g_pf_rules; e != NULL; e = e->next)
    if (e->rules == rules)
      return e;

  e = (pf_rules*) calloc(1, sizeof(*e));
  if (e == NULL)
    return NULL;
  e->rules = rules;
  e->n = rules->num_rules;
  e->state = (uint8_t*) malloc(e->n ? e->n : 1);
  if (e->state == NULL)
  {
    free(e);
    return NULL;
  }
  e->walked = classify(rules, e->state);
  e->next = g_pf_rules;
  g_pf_rules = e;
  return e;
}

/* Entries are recycled, never freed (a thread's one-entry cache may still point at one), so
   reading e->ctx is always safe; an entry belongs to ctx only while e->ctx == ctx. */
static pf_scan* scan_entry_for(YR_SCAN_CONTEXT* ctx)
{
  const uint32_t n = ctx->rules->num_rules;

  if (t_note_ctx == ctx && t_note_scan != NULL && t_note_scan->ctx == ctx && t_note_scan->n == n)
    return t_note_scan;

  pf_scan* e;
  pf_scan* spare = NULL;
  pthread_mutex_lock(&g_pf_lock);
  for (e = g_pf_scans; e != NULL; e = e->next)
  {
    if (e->ctx == ctx)
      break;
    if (e->ctx == NULL && spare == NULL)
      spare = e;
  }

  if (e == NULL)
  {
    e = spare;
    if (e == NULL)
    {
      e = (pf_scan*) calloc(1, sizeof(*e));
      if (e != NULL)
      {
        e->next = g_pf_scans;
        g_pf_scans = e;
      }
    }
    if (e != NULL)
    {
      e->ctx = ctx;
      e->n = 0; /* forces the (re)allocation below */
    }
  }

  if (e != NULL && e->n != n)
  {
    uint8_t* h = (uint8_t*) realloc(e->hit, n ? n : 1);
    if (h != NULL)
    {
      e->hit = h;
      e->n = n;
      memset(e->hit, 0, n ? n : 1);
    }
    else
    {
      e->ctx = NULL;
      e = NULL;
    }
  }
  pthread_mutex_unlock(&g_pf_lock);

  t_note_ctx = ctx;
  t_note_scan = e;
  return e;
}

void yr_b200_prefilter_note(YR_SCAN_CONTEXT* ctx, YR_STRING* string)
{
  if (!yr_b200_prefilter_enabled())
    return;
  pf_scan* e = scan_entry_for(ctx);
  if (e != NULL && string->rule_idx < e->n)
    e->hit[string->rule_idx] = 1;
}

/* the last answer of yr_b200_prefilter_skip (see there) */
static __thread uint32_t t_asked_rule = 0xFFFFFFFFu;
static __thread int t_asked_answer = 0;
static __thread int t_asked_once = 0;

void yr_b200_prefilter_begin(YR_SCAN_CONTEXT* ctx)
{
  t_exec_ctx = NULL;
  if (!yr_b200_prefilter_enabled())
    return;

  pf_scan* se = scan_entry_for(ctx);
  pthread_mutex_lock(&g_pf_lock);
  pf_rules* re = rules_entry_for(ctx->rules);
  pthread_mutex_unlock(&g_pf_lock);
  if (se == NULL || re == NULL || re->n != se->n)
    return;

  t_exec_state = re->state;
  t_exec_hit = se->hit;
  t_evaluated = t_skipped = 0;
  t_asked_once = 0;
  t_exec_ctx = ctx;
}

/* OP_INIT_RULE asks TWICE (exec.c:1168 for the jump, exec.c:1172 for the operand it steps over when the jump is not
   taken) and both answers must be the same one: the state table is shared by every thread that scans with these
   rules, and another thread may learn "skippable" between the two questions -- the jump not taken and the operand
   not stepped over leaves ip inside the instruction (seen as ERROR_COULD_NOT_MAP_FILE from the exception handler,
   or worse).  The first answer is remembered per thread and repeated for the second question. */

int yr_b200_prefilter_skip(YR_SCAN_CONTEXT* ctx, uint32_t rule_idx)
{
  if (t_exec_ctx != ctx)
    return 0;
  if (t_asked_once && t_asked_rule == rule_idx)
  {
    t_asked_once = 0;
    return t_asked_answer;
  }
  const int answer = (*(volatile uint8_t*) &t_exec_state[rule_idx] == PF_SKIPPABLE && !t_exec_hit[rule_idx]) ? 1 : 0;
  t_asked_rule = rule_idx;
  t_asked_answer = answer;
  t_asked_once = 1;
  if (answer)
    t_skipped++;
  return answer;
}

void yr_b200_prefilter_learn(YR_SCAN_CONTEXT* ctx, uint32_t rule_idx, int result)
{
  if (t_exec_ctx != ctx)
    return;
  t_evaluated++;
  /* racing threads store the same value: the rule is a function of its strings' match state only */
  if (t_exec_state[rule_idx] == PF_UNKNOWN && !t_exec_hit[rule_idx])
    *(volatile uint8_t*) &t_exec_state[rule_idx] = result ? PF_TRUE : PF_SKIPPABLE;
}

void yr_b200_prefilter_end(YR_SCAN_CONTEXT* ctx)
{
  if (t_exec_ctx != ctx)
    return;
  t_asked_once = 0;
  __atomic_fetch_add(&g_pf_evaluated, t_evaluated, __ATOMIC_RELAXED);
  __atomic_fetch_add(&g_pf_skipped, t_skipped, __ATOMIC_RELAXED);
  memset(t_exec_hit, 0, ctx->rules->num_rules);
  t_exec_ctx = NULL;
}

void yr_b200_prefilter_stats(uint64_t* evaluated, uint64_t* skipped)
{
  *evaluated = __atomic_load_n(&g_pf_evaluated, __ATOMIC_RELAXED);
  *skipped = __atomic_load_n(&g_pf_skipped, __ATOMIC_RELAXED);
}

/* how the rules of a YR_RULES were classified: counts of impure / not yet known / skippable / true,
   and (counts[4]) how many rules the classifier reached */
void yr_b200_prefilter_classes(YR_RULES* rules, uint32_t counts[5])
{
  memset(counts, 0, 5 * sizeof(uint32_t));
  pthread_mutex_lock(&g_pf_lock);
  pf_rules* e = rules_entry_for(rules);
  if (e != NULL)
  {
    for (uint32_t i = 0; i < e->n; i++) counts[e->state[i] & 3]++;
    counts[4] = e->walked;
  }
  pthread_mutex_unlock(&g_pf_lock);
}

void yr_b200_prefilter_scanner_release(YR_SCAN_CONTEXT* ctx)
{
  if (t_note_ctx == ctx)
  {
    t_note_ctx = NULL;
    t_note_scan = NULL;
  }
  pthread_mutex_lock(&g_pf_lock);
  for (pf_scan* e = g_pf_scans; e != NULL; e = e->next)
  {
    if (e->ctx == ctx)
    {
      e->ctx = NULL; /* recycled by the next scanner */
      break;
    }
  }
  pthread_mutex_unlock(&g_pf_lock);
}

void yr_b200_prefilter_rules_release(YR_RULES* rules)
{
  pthread_mutex_lock(&g_pf_lock);
  for (pf_rules** pp = &g_pf_rules; *pp != NULL; pp = &(*pp)->next)
  {
    if ((*pp)->rules == rules)
    {
      pf_rules* e = *pp;
      *pp = e->next;
      free(e->state);
      free(e);
      break;
    }
  }
  pthread_mutex_unlock(&g_pf_lock);
}
