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

void yr_b200_scanner_release(YR_SCANNER* scanner);
void yr_b200_rules_release(YR_RULES* rules);

/* Which CUDA device the glue uses for new scanners (default: $YR_B200_DEVICE or 0). */
void yr_b200_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif
