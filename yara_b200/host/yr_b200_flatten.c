/*
 * yr_b200_flatten.c -- "rules upload", host half: turn a compiled YR_RULES
 * (pointers into a 12-buffer YR_ARENA) into the flat, pointer-free rules image
 * of include/yrb_image.h.  Pure C, no CUDA; compiled against the reference's
 * own headers so that struct layouts cannot drift.
 *
 * Reads: rules->ac_transition_table, ac_match_table,
 * ac_match_pool, strings_table, num_strings, arena (types.h:552-622).
 */
#include <stdlib.h>
#include <string.h>
#include <yara.h>
#include <yara/arena.h>
#include <yara/compiler.h>
#include <yara/globals.h>
#include <yara/re.h>

#include "yr_b200.h"
#include "yrb_image.h"

static uint64_t align16(uint64_t x) { return (x + 15) & ~(uint64_t) 15; }

/* Byte length of one RE instruction (re.h:65-92, packed args re.c:69-87). 0 = unknown. */
static int re_insn_len(uint8_t op)
{
  switch (op)
  {
  case RE_OPCODE_ANY:
  case RE_OPCODE_WORD_CHAR:
  case RE_OPCODE_NON_WORD_CHAR:
  case RE_OPCODE_SPACE:
  case RE_OPCODE_NON_SPACE:
  case RE_OPCODE_DIGIT:
  case RE_OPCODE_NON_DIGIT:
  case RE_OPCODE_MATCH:
  case RE_OPCODE_MATCH_AT_END:
  case RE_OPCODE_MATCH_AT_START:
  case RE_OPCODE_WORD_BOUNDARY:
  case RE_OPCODE_NON_WORD_BOUNDARY:
    return 1;
  case RE_OPCODE_LITERAL:
  case RE_OPCODE_NOT_LITERAL:
    return 2;
  case RE_OPCODE_MASKED_LITERAL:
  case RE_OPCODE_MASKED_NOT_LITERAL:
  case RE_OPCODE_JUMP:
    return 3;
  case RE_OPCODE_CLASS:
    return 1 + (int) sizeof(RE_CLASS);
  case RE_OPCODE_REPEAT_ANY_GREEDY:
  case RE_OPCODE_REPEAT_ANY_UNGREEDY:
    return 5;
  case RE_OPCODE_SPLIT_A:
  case RE_OPCODE_SPLIT_B:
    return 4;
  case RE_OPCODE_REPEAT_START_GREEDY:
  case RE_OPCODE_REPEAT_END_GREEDY:
  case RE_OPCODE_REPEAT_START_UNGREEDY:
  case RE_OPCODE_REPEAT_END_UNGREEDY:
    return 9;
  }
  return 0;
}

/* Deepest REPEAT_START nesting met walking linearly from `off` to the MATCH
   that ends the program: bounds RE_FIBER.sp (re.c:1551,1580). */
static uint32_t re_stack_depth(const uint8_t* code, size_t code_size, size_t off)
{
  int depth = 0, max_depth = 0;

  while (off < code_size)
  {
    uint8_t op = code[off];
    int len = re_insn_len(op);

    if (len == 0 || op == RE_OPCODE_MATCH)
      break;

    if (op == RE_OPCODE_REPEAT_START_GREEDY || op == RE_OPCODE_REPEAT_START_UNGREEDY)
    {
      depth++;
      if (depth > max_depth)
        max_depth = depth;
    }
    else if (op == RE_OPCODE_REPEAT_END_GREEDY || op == RE_OPCODE_REPEAT_END_UNGREEDY)
    {
      depth--;
    }

    off += len;
  }

  return (uint32_t) max_depth;
}

int yr_b200_rules_flatten(YR_RULES* rules, void** image_out, size_t* size_out)
{
  YR_ARENA* arena = rules->arena;

  const size_t trans_bytes = arena->buffers[YR_AC_TRANSITION_TABLE].used;
  const uint32_t tables_size = (uint32_t) (trans_bytes / sizeof(YR_AC_TRANSITION));
  const uint32_t num_acm =
      (uint32_t) (arena->buffers[YR_AC_STATE_MATCHES_POOL].used / sizeof(YR_AC_MATCH));
  const uint32_t num_strings = rules->num_strings;
  const uint8_t* code_base = arena->buffers[YR_RE_CODE_SECTION].data;
  const size_t code_size = arena->buffers[YR_RE_CODE_SECTION].used;

  size_t lit_size = 0;

  for (uint32_t i = 0; i < num_strings; i++)
  {
    YR_STRING* s = &rules->strings_table[i];
    if (STRING_IS_LITERAL(s))
      lit_size += (size_t) s->length;
  }

  yrb_image_header h;
  memset(&h, 0, sizeof(h));
  h.magic = YRB_IMAGE_MAGIC;
  h.version = YRB_IMAGE_VERSION;
  h.tables_size = tables_size;
  h.num_acm = num_acm;
  h.num_strings = num_strings;
  h.num_rules = rules->num_rules;
  h.lit_size = (uint32_t) lit_size;
  h.code_size = (uint32_t) code_size;

  uint64_t off = align16(sizeof(h));
  h.off_trans = off;   off = align16(off + (uint64_t) tables_size * 4);
  h.off_match = off;   off = align16(off + (uint64_t) tables_size * 4);
  h.off_acm = off;     off = align16(off + (uint64_t) num_acm * sizeof(yrb_acm));
  h.off_strings = off; off = align16(off + (uint64_t) num_strings * sizeof(yrb_string));
  h.off_lit = off;     off = align16(off + lit_size + 16);
  h.off_code = off;    off = align16(off + code_size + 16);
  h.off_lower = off;   off += 256;
  h.off_alter = off;   off += 256;
  h.total_size = off;

  uint8_t* img = (uint8_t*) calloc(1, (size_t) off);

  if (img == NULL)
    return ERROR_INSUFFICIENT_MEMORY;

  memcpy(img + h.off_trans, rules->ac_transition_table, (size_t) tables_size * 4);
  memcpy(img + h.off_match, rules->ac_match_table, (size_t) tables_size * 4);

  if (code_size > 0)
    memcpy(img + h.off_code, code_base, code_size);

  memcpy(img + h.off_lower, yr_lowercase, 256);
  memcpy(img + h.off_alter, yr_altercase, 256);

  yrb_acm* acm = (yrb_acm*) (img + h.off_acm);
  uint32_t max_stack = 0;

  for (uint32_t i = 0; i < num_acm; i++)
  {
    const YR_AC_MATCH* m = &rules->ac_match_pool[i];

    acm[i].string_idx = m->string->idx;
    acm[i].fwd_code = m->forward_code ? (uint32_t) (m->forward_code - code_base) : YRB_NONE;
    acm[i].bwd_code = m->backward_code ? (uint32_t) (m->backward_code - code_base) : YRB_NONE;
    acm[i].next = m->next ? (uint32_t) (m->next - rules->ac_match_pool) + 1 : 0;
    acm[i].backtrack = m->backtrack;

    if (m->forward_code)
    {
      uint32_t d = re_stack_depth(code_base, code_size, acm[i].fwd_code);
      if (d > max_stack) max_stack = d;
    }
    if (m->backward_code)
    {
      uint32_t d = re_stack_depth(code_base, code_size, acm[i].bwd_code);
      if (d > max_stack) max_stack = d;
    }
  }

  yrb_string* strs = (yrb_string*) (img + h.off_strings);
  uint8_t* lit = img + h.off_lit;
  uint32_t lit_off = 0, num_chains = 0;

  for (uint32_t i = 0; i < num_strings; i++)
  {
    YR_STRING* s = &rules->strings_table[i];

    strs[i].flags = s->flags;
    strs[i].length = s->length;
    strs[i].chained_to = s->chained_to ? s->chained_to->idx : YRB_NONE;
    strs[i].chain_gap_min = s->chain_gap_min;
    strs[i].chain_gap_max = s->chain_gap_max;
    strs[i].fixed_offset = s->fixed_offset;
    strs[i].rule_idx = s->rule_idx;
    strs[i].lit_off = YRB_NONE;
    strs[i].chain_head = YRB_NONE;

    if (STRING_IS_LITERAL(s))
    {
      strs[i].lit_off = lit_off;
      memcpy(lit + lit_off, s->string, (size_t) s->length);
      lit_off += (uint32_t) s->length;
    }

    if (STRING_IS_CHAIN_PART(s))
    {
      YR_STRING* head = s;
      while (head->chained_to != NULL) head = head->chained_to;
      strs[i].chain_head = head->idx;
      if (head == s)
        num_chains++;
    }
  }

  h.max_re_stack = max_stack;
  h.num_chains = num_chains;
  memcpy(img, &h, sizeof(h));

  *image_out = img;
  *size_out = (size_t) off;

  return ERROR_SUCCESS;
}

void yr_b200_image_free(void* image) { free(image); }
