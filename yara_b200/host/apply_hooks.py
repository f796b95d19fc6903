#!/usr/bin/env python3
"""Build-time application of the INTEGRATION.md patch.

Reads one reference source file where it lies (never copied into this
repository), applies the hook lines a libyara maintainer would add, and writes
the result into the scratch build directory.  Each edit must match exactly once;
anything else means the reference drifted and the build stops.

  scanner.c  #include "yr_b200.h"
             scanner.c:495   _yr_scanner_scan_mem_block -> yr_b200_scan_mem_block
             scanner.c:287   yr_scanner_destroy: yr_b200_scanner_release(scanner)
  rules.c    #include "yr_b200.h"
             rules.c:515     yr_rules_destroy: yr_b200_rules_release(rules)
"""
import sys

EDITS = {
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
