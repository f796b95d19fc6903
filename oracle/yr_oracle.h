/*
 * yr_oracle.h -- CPU restatement of YARA 4.2.1's byte-scanning hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the product (yara_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  The restatement is checked against the reference's own
 * libyara built from /root/reference (oracle/_ref/libyara_ref.so) on the
 * known-answer cases of tests/test-rules.c re-stated in tests/test_oracle_*.py and
 * on seeded synthetic rulesets/corpora, and against the golden fixtures under
 * tests/golden/ that were generated from that same reference build.
 */
#ifndef YR_ORACLE_H
#define YR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "yrb_image.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yro_result
{
  int error;               /* ERROR_* of the scan (0 = success) */
  uint32_t error_string;   /* scanner->last_error_string idx, YRB_NONE if none */
  size_t n_matches;        /* confirmed matches, grouped by string asc, list order inside */
  yrb_match* matches;
  size_t n_unconfirmed;    /* leftovers of unconfirmed_matches[], same grouping */
  yrb_match* unconfirmed;
  uint32_t n_too_many;     /* strings for which CALLBACK_MSG_TOO_MANY_MATCHES fired, in firing order */
  uint32_t* too_many;
  uint64_t n_candidates;   /* calls to yr_scan_verify_match (scanner.c:111,153) */
  uint64_t max_fibers;     /* peak live fibers in any yr_re_exec */
  uint64_t max_positions;  /* peak list length in any yr_re_fast_exec */
} yro_result;

/*
 * One _yr_scanner_scan_mem_block (scanner.c:45-176) over data[0..size) with
 * block base `base`, starting from empty match lists.
 *   scan_flags        SCAN_FLAGS_* (only FAST_MODE matters on this path)
 *   too_many_action   what the user callback returns for TOO_MANY_MATCHES
 *                     (0 = CALLBACK_CONTINUE, else abort with error 30)
 */
yro_result* yro_scan_block(
    const void* image,
    const uint8_t* data,
    size_t size,
    uint64_t base,
    int scan_flags,
    int too_many_action);

void yro_result_free(yro_result* r);

/* Trace of the match callbacks of the next yro_scan_block calls (see yr_oracle.c). */
typedef struct yro_callback
{
  int64_t offset;
  uint32_t string_idx;
  int32_t length;
  uint32_t xor_key;
  uint32_t pad;
} yro_callback;

void yro_trace_begin(yro_callback* buf, size_t cap);
size_t yro_trace_end(void);

/* Only the AC walk: emits (i, acm_index) for every yr_scan_verify_match call,
   in discovery order.  Returns the count; writes at most cap pairs. */
size_t yro_ac_events(
    const void* image, const uint8_t* data, size_t size, uint64_t* ev_pos, uint32_t* ev_acm, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
