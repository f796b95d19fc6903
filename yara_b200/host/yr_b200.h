/*
 * yr_b200.h -- libyara-side glue of the B200 scan path (the code a libyara
 * maintainer would place in libyara/cuda/).  Everything here speaks libyara
 * types; everything below it (include/yara_b200.h) speaks plain pointers.
 *
 * The three hooks are what INTEGRATION.md patches into the reference:
 *   scanner.c:495   result = yr_b200_scan_mem_block(scanner, data, block);
 *   scanner.c:287   yr_b200_scanner_release(scanner);   (yr_scanner_destroy)
 *   rules.c:515     yr_b200_rules_release(rules);       (yr_rules_destroy)
 */
#ifndef YR_B200_H
#define YR_B200_H

#include <stddef.h>
#include <stdint.h>
#include <yara/types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rules upload, host half (yr_b200_flatten.c). Caller frees with yr_b200_image_free. */
int yr_b200_rules_flatten(YR_RULES* rules, void** image_out, size_t* size_out);
void yr_b200_image_free(void* image);

/* Replaces _yr_scanner_scan_mem_block (scanner.c:45-176); same signature,
   same return codes, same side effects on scanner->matches[] /
   unconfirmed_matches[] / strings_temp_disabled / last_error_string. */
int yr_b200_scan_mem_block(
    YR_SCANNER* scanner,
    const uint8_t* block_data,
    YR_MEMORY_BLOCK* block);

/* yr_b200_scan_mem_block may defer the small blocks of a multi-block scan (process memory, custom iterators) and
   scan them as one engine batch; yr_scanner_scan_mem_blocks calls this behind its block loop (scanner.c:502) so that
   nothing is left deferred.  Same return codes as a block scan.  $YR_B200_BLOCK_BATCH=<bytes> sets the largest block
   that is deferred (default 262144; 0 turns the batching off). */
int yr_b200_blocks_flush(YR_SCANNER* scanner);

/* scanner.c:545, at _exit next to _yr_scanner_clean_matches: a scan that did not reach the flush above (timeout,
   callback abort, exception in a block) drops what it had deferred, so that a reused YR_SCANNER starts its next scan
   with nothing pending. */
void yr_b200_blocks_discard(YR_SCANNER* scanner);

void yr_b200_scanner_release(YR_SCANNER* scanner);

/* yr_scanner_scan_mem brackets its scan with (scanner, 1) / (scanner, 0): only then is a block known to
   be the whole scan, which is what a parked file-batch result stands for. */
void yr_b200_single_block_scan(YR_SCANNER* scanner, int on);
void yr_b200_rules_release(YR_RULES* rules);

/*
 * File batches (optional accelerator; the libyara API itself is unchanged).  A caller
 * with many buffers to scan -- the file queue of cli/yara.c:365-457 -- hands them over
 * first; the GPU scans them with one launch sequence and the per-buffer results are
 * parked in the batch.  While the batch is alive, yr_rules_scan_mem / yr_scanner_scan_mem
 * of the same rules over one of those buffers (same pointer and size, block base 0,
 * no SCAN_FLAGS_FAST_MODE) takes its string matches from the batch instead of scanning;
 * condition evaluation and callbacks run as always.  Buffers whose result depends on
 * scan state (chained strings, verifier errors, the 1M cap) are not parked and are
 * scanned on demand as before, so results never differ from unbatched scans.
 * The buffers must stay unchanged while the batch is alive; destroy the batch only
 * when no scan is using it.
 */
typedef struct YR_B200_BATCH YR_B200_BATCH;

int yr_b200_batch_scan(
    YR_RULES* rules, const uint8_t* const* buffers, const size_t* sizes, uint32_t n, int flags,
    int one_allocation, /* non-zero: all buffers lie in ascending order inside one allocation */
    YR_B200_BATCH** batch);
void yr_b200_batch_destroy(YR_B200_BATCH* batch);
uint32_t yr_b200_batch_rescans(const YR_B200_BATCH* batch); /* buffers that were not parked */
uint64_t yr_b200_batch_hits(void);                          /* scans served from parked results so far */

/*
 * Condition-VM pre-filter (yr_b200_prefilter.c; on unless YR_B200_RULE_PREFILTER=0).
 * Hook sites: scanner.c:518 (begin/end around yr_execute_code), exec.c:1168-1173 (skip test next
 * to RULE_IS_DISABLED in OP_INIT_RULE), exec.c:1185 (learn in OP_MATCH_RULE); matches are noted
 * where they are created (materialise() in the glue).
 */
int yr_b200_prefilter_enabled(void);
void yr_b200_prefilter_note(YR_SCAN_CONTEXT* context, YR_STRING* string);
void yr_b200_prefilter_begin(YR_SCAN_CONTEXT* context);
int yr_b200_prefilter_skip(YR_SCAN_CONTEXT* context, uint32_t rule_idx);
void yr_b200_prefilter_learn(YR_SCAN_CONTEXT* context, uint32_t rule_idx, int result);
void yr_b200_prefilter_end(YR_SCAN_CONTEXT* context);
void yr_b200_prefilter_stats(uint64_t* evaluated, uint64_t* skipped);
void yr_b200_prefilter_classes(YR_RULES* rules, uint32_t counts[5]);
void yr_b200_prefilter_scanner_release(YR_SCAN_CONTEXT* context);
void yr_b200_prefilter_rules_release(YR_RULES* rules);

/* Request coalescing ($YR_B200_COALESCE=N, off by default): concurrent one-block scans of the same
   rules share one yrb_scan_batch call; how many batches served how many scans so far. */
void yr_b200_coalesce_stats(uint64_t* batches, uint64_t* scans);

/* Which CUDA device the glue uses for new scanners (default: $YR_B200_DEVICE or 0). */
void yr_b200_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif
