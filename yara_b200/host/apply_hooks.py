#!/usr/bin/env python3
"""Build-time application of the INTEGRATION.md patch.

Reads one reference source file where it lies (never copied into this
repository), applies the hook lines a libyara maintainer would add, and writes
the result into the scratch build directory.  Each edit must match exactly once;
anything else means the reference drifted and the build stops.

  scanner.c  #include "yr_b200.h"
             scanner.c:495   _yr_scanner_scan_mem_block -> yr_b200_scan_mem_block
             scanner.c:287   yr_scanner_destroy: yr_b200_scanner_release(scanner)
             scanner.c:660   yr_scanner_scan_mem: yr_b200_single_block_scan(scanner, 1/0) around the scan
                             (parked file-batch results are complete one-block results)
             scanner.c:502   behind the block loop: yr_b200_blocks_flush(scanner) (small blocks of a multi-block
                             scan are deferred and scanned as one batch)
             scanner.c:545   at _exit, where the scan's matches and notebook go: yr_b200_blocks_discard(scanner)
                             (a scan that ended early -- timeout, callback abort, exception -- leaves nothing deferred)
  rules.c    #include "yr_b200.h"
             rules.c:515     yr_rules_destroy: yr_b200_rules_release(rules)
  condition-VM pre-filter (yr_b200_prefilter.c, on unless YR_B200_RULE_PREFILTER=0):
  scanner.c  scanner.c:518   yr_b200_prefilter_begin/end around yr_execute_code
  exec.c     exec.c:1168     OP_INIT_RULE: `|| yr_b200_prefilter_skip(...)` next to RULE_IS_DISABLED
             exec.c:1191     OP_MATCH_RULE: yr_b200_prefilter_learn(...)

The "*_pf" variants build oracle/_ref/libyara_refpf.so, the CPU validation library of the
pre-filter: the reference with ONLY the pre-filter hooks (no GPU path), matches noted in
scan.c:642 (_yr_scan_match_callback).
"""
import sys

EXEC_CALL = (
    "      { result = yr_execute_code(scanner); },\n",
    "      { yr_b200_prefilter_begin(scanner); result = yr_execute_code(scanner); yr_b200_prefilter_end(scanner); },\n",
)
DESTROY_ANCHOR = "  RE_FIBER* fiber = scanner->re_fiber_pool.fibers.head;\n"
RULES_ANCHOR = "  YR_EXTERNAL_VARIABLE* external = rules->ext_vars_table;\n\n  while (!EXTERNAL_VARIABLE_IS_NULL(external))\n  {\n    if (external->type == EXTERNAL_VARIABLE_TYPE_MALLOC_STRING)\n      yr_free(external->value.s);"

EDITS = {
    "scanner_pf": [
        ('#include "exception.h"\n', '#include "exception.h"\n#include "yr_b200.h"\n'),
        EXEC_CALL,
        (DESTROY_ANCHOR, "  yr_b200_prefilter_scanner_release(scanner);\n\n" + DESTROY_ANCHOR),
    ],
    "rules_pf": [
        ("#include <yara/utils.h>\n", '#include <yara/utils.h>\n#include "yr_b200.h"\n'),
        (RULES_ANCHOR, RULES_ANCHOR.replace("\n\n  while", "\n\n  yr_b200_prefilter_rules_release(rules);\n\n  while", 1)),
    ],
    "scan_pf": [
        ("#include <yara/utils.h>\n", '#include <yara/utils.h>\n#include "yr_b200.h"\n'),
        (
            "  YR_STRING* string = callback_args->string;\n  YR_MATCH* new_match;\n",
            "  YR_STRING* string = callback_args->string;\n  YR_MATCH* new_match;\n\n  yr_b200_prefilter_note(callback_args->context, string);\n",
        ),
    ],
    "exec": [
        ("#include <yara/utils.h>\n", '#include <yara/utils.h>\n#include "yr_b200.h"\n'),
        (
            "      ip = jmp_if(RULE_IS_DISABLED(current_rule), ip);\n",
            "      ip = jmp_if(\n          RULE_IS_DISABLED(current_rule) || yr_b200_prefilter_skip(context, current_rule_idx), ip);\n",
        ),
        (
            "      if (!RULE_IS_DISABLED(current_rule))\n        ip += sizeof(uint32_t);\n",
            "      if (!(RULE_IS_DISABLED(current_rule) || yr_b200_prefilter_skip(context, current_rule_idx)))\n        ip += sizeof(uint32_t);\n",
        ),
        (
            "      if (!is_undef(r1) && r1.i)\n        yr_bitmask_set(context->rule_matches_flags, r2.i);\n",
            "      yr_b200_prefilter_learn(context, (uint32_t) r2.i, !is_undef(r1) && r1.i);\n\n      if (!is_undef(r1) && r1.i)\n        yr_bitmask_set(context->rule_matches_flags, r2.i);\n",
        ),
    ],
    "scanner": [
        ('#include "exception.h"\n', '#include "exception.h"\n#include "yr_b200.h"\n'),
        (
            "{ result = _yr_scanner_scan_mem_block(scanner, data, block); }",
            "{ result = yr_b200_scan_mem_block(scanner, data, block); }",
        ),
        (
            "  RE_FIBER* fiber = scanner->re_fiber_pool.fibers.head;\n",
            "  yr_b200_scanner_release(scanner);\n\n  RE_FIBER* fiber = scanner->re_fiber_pool.fibers.head;\n",
        ),
        EXEC_CALL,
        (
            "  result = iterator->last_error;\n\n  if (result != ERROR_SUCCESS)\n    goto _exit;\n",
            "  result = yr_b200_blocks_flush(scanner); /* blocks whose scan yr_b200_scan_mem_block deferred */\n\n"
            "  if (result != ERROR_SUCCESS)\n    goto _exit;\n\n"
            "  result = iterator->last_error;\n\n  if (result != ERROR_SUCCESS)\n    goto _exit;\n",
        ),
        (
            "  if (result != ERROR_BLOCK_NOT_READY)\n  {\n    _yr_scanner_clean_matches(scanner);\n",
            "  if (result != ERROR_BLOCK_NOT_READY)\n  {\n    yr_b200_blocks_discard(scanner); /* a scan that ended early leaves nothing deferred */\n"
            "    _yr_scanner_clean_matches(scanner);\n",
        ),
        (
            "  int result = yr_scanner_scan_mem_blocks(scanner, &iterator);\n",
            "  yr_b200_single_block_scan(scanner, 1); /* yr_scanner_scan_mem: this block is the whole scan */\n"
            "  int result = yr_scanner_scan_mem_blocks(scanner, &iterator);\n"
            "  yr_b200_single_block_scan(scanner, 0);\n",
        ),
    ],
    "rules": [
        ("#include <yara/utils.h>\n", '#include <yara/utils.h>\n#include "yr_b200.h"\n'),
        (
            "  YR_EXTERNAL_VARIABLE* external = rules->ext_vars_table;\n\n  while (!EXTERNAL_VARIABLE_IS_NULL(external))\n  {\n    if (external->type == EXTERNAL_VARIABLE_TYPE_MALLOC_STRING)\n      yr_free(external->value.s);",
            "  YR_EXTERNAL_VARIABLE* external = rules->ext_vars_table;\n\n  yr_b200_rules_release(rules);\n\n  while (!EXTERNAL_VARIABLE_IS_NULL(external))\n  {\n    if (external->type == EXTERNAL_VARIABLE_TYPE_MALLOC_STRING)\n      yr_free(external->value.s);",
        ),
    ],
}


def main():
    which, src, dst = sys.argv[1:4]
    text = open(src).read()
    for old, new in EDITS[which]:
        if text.count(old) != 1:
            sys.exit("apply_hooks: %r matches %d times in %s" % (old[:50], text.count(old), src))
        text = text.replace(old, new)
    open(dst, "w").write(text)


if __name__ == "__main__":
    main()
