/*
 * yara_b200.h -- C ABI of the B200 (sm_100a) scan engine: libyrb200.so.
 *
 * Plain pointers and sizes only; no libyara types, no torch types.  This is the
 * layer the reference's `libyara/cuda` glue binds (yara_b200/host/yr_b200_glue.c
 * is that glue; INTEGRATION.md shows the three lines it adds to the reference).
 *
 * Reference interface each entry point stands in for (libyara 4.2.1):
 *
 *   yrb_rules_create / _release   "rules upload": the tables yr_rules_from_arena
 *                                 wires up (rules.c:326-370) and that
 *                                 _yr_scanner_scan_mem_block reads
 *                                 (scanner.c:61-63); lifetime of yr_rules_destroy
 *                                 (rules.c:515).
 *   yrb_scanner_create / _destroy per-scanner scan state: yr_scanner_create /
 *                                 yr_scanner_destroy (scanner.c:205,287) -- the
 *                                 re_fiber_pool / re_fast_exec_position_pool /
 *                                 matches[] of YR_SCAN_CONTEXT (types.h:740-821)
 *                                 become device work areas.
 *   yrb_scan_host                 _yr_scanner_scan_mem_block(scanner, block_data,
 *                                 block) (scanner.c:45-176) + everything it
 *                                 calls: yr_scan_verify_match (scan.c:992),
 *                                 _yr_scan_verify_literal_match (scan.c:887),
 *                                 _yr_scan_verify_re_match (scan.c:778),
 *                                 yr_re_exec / yr_re_fast_exec (re.c:1693,2150),
 *                                 _yr_scan_match_callback (scan.c:634),
 *                                 _yr_scan_add_match_to_list (scan.c:290),
 *                                 _yr_scan_verify_chained_string_match (scan.c:406).
 *   yrb_scan_device               same, for a block already resident in HBM
 *                                 (no reference equivalent; used for the
 *                                 HBM-resident throughput figure and by callers
 *                                 that produce data on the device).
 *   yrb_scan_batch                N independent buffers -- N x yr_rules_scan_mem over a file
 *                                 queue (cli/yara.c:365-457), each of which would run
 *                                 _yr_scanner_scan_mem_block from the root state -- with one
 *                                 launch sequence; results per buffer equal yrb_scan_host's.
 *   yrb_shard_*                   byte-range sharding of ONE block over several
 *                                 GPUs/processes; no reference
 *                                 equivalent (a buffer is always scanned by one
 *                                 thread there, cli/yara.c:1246).
 *   yrb_multi_*                   the same sharding driven from ONE process: what
 *                                 yr_rules_scan_mem (rules.c:172-207) of the drop-in
 *                                 calls for a large block when several GPUs are
 *                                 configured; the reference's way to use more
 *                                 workers is per file only (cli/yara.c:1561-1610).
 *
 * Errors are the reference's ERROR_* values (error.h:40-109): 0 success,
 * 1 insufficient memory, 30 too many matches, 46 too many RE fibers, 31 internal.
 * There is NO CPU fallback: without a usable CUDA device every call fails with
 * YRB_ERROR_NO_DEVICE.
 */
#ifndef YARA_B200_H
#define YARA_B200_H

#include <stddef.h>
#include <stdint.h>

#include "yrb_image.h"

#ifdef __cplusplus
extern "C" {
#endif

#define YRB_ERROR_NO_DEVICE     1001
#define YRB_ERROR_BAD_IMAGE     1002
#define YRB_ERROR_CUDA          1003
#define YRB_ERROR_BAD_ARGUMENT  1004
#define YRB_ERROR_MARGIN        1005 /* a shard's read margin was too small for a verification */

typedef struct yrb_rules yrb_rules;
typedef struct yrb_scanner yrb_scanner;

/* Which AC scan kernel to use (yrb_scanner_set_option YRB_OPT_SCAN_KERNEL). */
#define YRB_KERNEL_AUTO   0
#define YRB_KERNEL_WALK   1 /* per-thread chunked walk of the transition table */
#define YRB_KERNEL_FILTER 2 /* shared-memory atom prefilter + exact walk of survivors */

#define YRB_OPT_SCAN_KERNEL   1
#define YRB_OPT_H2D_CHUNK     2 /* bytes per H2D slice in yrb_scan_host (default 64 MiB) */
#define YRB_OPT_SHARD_ON_DEVICE 3 /* 1: a defer_resolve scan leaves its raw records in device memory (no D2H) */

struct yrb_seed;

typedef struct yrb_scan_params
{
  uint64_t base;           /* block->base (only added to fixed_offset checks, scan.c:1025) */
  int32_t scan_flags;      /* SCAN_FLAGS_* (scan.h:38-42) */
  int32_t too_many_action; /* what the user's callback would answer to
                              CALLBACK_MSG_TOO_MANY_MATCHES: 0 = CALLBACK_CONTINUE
                              (string gets disabled, scan.c:1064-1067), else abort */
  /* Byte-range shard of a larger block.  For a whole block leave all zero.
     data[0] is byte `data_offset` of the block, whose full length is
     `block_size`; this call owns the AC events i in (own_begin, own_end]
     (plus i == 0 when own_begin == 0).  Bytes outside the owned range are
     read-only margin (3 bytes before for the AC state, up to
     max(YR_RE_SCAN_LIMIT, 2*longest literal) either side for verification). */
  uint64_t data_offset;
  uint64_t block_size;
  uint64_t own_begin;
  uint64_t own_end;
  int32_t defer_resolve;   /* 1: stop after verification; fetch raw records with
                              yrb_shard_raw and resolve all shards together with
                              yrb_resolve_raw */
  uint32_t n_seeds;        /* multi-block scans: what earlier blocks of this scan left in
                              unconfirmed_matches[] of chained strings (scanner.c keeps those
                              lists across blocks, types.h:806-816) */
  const struct yrb_seed* seeds;
  uint64_t timeout_ns;     /* 0 = none.  Otherwise the scan gives up with ERROR_SCAN_TIMEOUT once this much time
                              has passed since the call began.  The reference looks at its stopwatch every 4096
                              input bytes (scanner.c:74-81); yrb_scan_host looks between H2D slices, keeping at
                              most four slices queued ahead, so a timeout takes effect within a few slices. */
  uint32_t n_skip_strings; /* multi-block scans: strings whose candidates the reference does not verify in this block
                              at all (scan.c:1018-1023) because of what EARLIER blocks of the scan did -- disabled by
                              the 1,000,000-match cap there, or SINGLE_MATCH strings that already matched, in fast
                              mode (heads of chained strings excepted: those travel as seeds of kind 1).  A verifier
                              error of such a string does not end the scan.  Their match records are the caller's to
                              drop.  0 / NULL: none. */
  uint32_t reserved2;
  const uint32_t* skip_strings;
} yrb_scan_params;

/* One YR_MATCH of unconfirmed_matches[string] carried over from an earlier block of
   the same scan (kind 0), or the fact that matches[string] is already non-empty
   (kind 1: the FAST_MODE guard of scan.c:1021-1023).  At most YRB_MAX_SEEDS (4M) per block. */
typedef struct yrb_seed
{
  int64_t base;          /* YR_MATCH.base: base of the block it was found in */
  int64_t offset;        /* YR_MATCH.offset, relative to THAT block (the reference compares
                            these raw offsets across blocks, scan.c:474-488: reproduced) */
  uint32_t string_idx;
  int32_t match_length;
  int32_t chain_length;
  uint8_t xor_key;
  uint8_t kind;
  uint16_t reserved;
} yrb_seed;

typedef struct yrb_scan_stats
{
  uint64_t bytes;          /* bytes whose AC events were owned by this call */
  uint64_t n_candidates;   /* yr_scan_verify_match calls that passed the cheap guards */
  uint64_t n_raw;          /* _yr_scan_match_callback calls that recorded a match */
  uint32_t kernel_launches;
  uint32_t scan_kernel;    /* YRB_KERNEL_* actually used */
  float ms_h2d;            /* CUDA-event times of the last call */
  float ms_scan;
  float ms_verify;
  float ms_resolve;
  float ms_total;
  float ms_kernel;         /* the last filter-kernel launch alone (0 when the walk kernel ran): the roofline figure */
  uint32_t reserved;
} yrb_scan_stats;

typedef struct yrb_scan_result
{
  int32_t error;           /* ERROR_* of the block scan */
  uint32_t error_string;   /* idx for scanner->last_error_string, YRB_NONE if none */
  size_t n_matches;        /* confirmed matches: the records of one string are contiguous and in
                              YR_MATCHES list order (ascending offset); plain strings come string-major,
                              the lists of chained strings' heads after them */
  const yrb_match* matches;
  size_t n_unconfirmed;    /* unconfirmed chain-part matches left at end of block */
  const yrb_match* unconfirmed;
  uint32_t n_too_many;     /* strings that hit YR_MAX_STRING_MATCHES, in firing order */
  const uint32_t* too_many;
  uint32_t seeds_consumed; /* 1: `unconfirmed` is the complete new state of the chained strings' lists
                              (the seeds that survived are in it, flagged YRB_MATCH_FROM_SEED) */
  uint32_t reserved;
  yrb_scan_stats stats;
} yrb_scan_result;
/* Pointers inside a result are owned by the scanner and stay valid until its
   next scan/resolve call or its destruction. */

int yrb_device_count(void);
const char* yrb_version(void);

int yrb_rules_create(const void* image, size_t image_size, int device, yrb_rules** out);
void yrb_rules_retain(yrb_rules* rules);
void yrb_rules_release(yrb_rules* rules);

/* What the atom prefilter made of this ruleset: whether the filter kernel is used at
   all, how many 4-byte hit windows it holds, how full its first-level bitmap is. */
typedef struct yrb_filter_info
{
  int32_t usable;       /* 0: every scan of these rules runs the table-walk kernel */
  int32_t variant;
  uint32_t n_windows;
  float fill3;          /* fraction of set bits in B3; a pair of positions passes level 1 with probability ~fill3^2 */
  int32_t has_depth1;   /* hit states of depth 1 / 2 (atoms of one or two bytes), zero-length atoms in the root state */
  int32_t has_depth2;
  int32_t root_match;
  int32_t reserved;
} yrb_filter_info;
int yrb_rules_filter_info(const yrb_rules* rules, yrb_filter_info* out);

int yrb_scanner_create(yrb_rules* rules, yrb_scanner** out);
void yrb_scanner_destroy(yrb_scanner* scanner);
int yrb_scanner_set_option(yrb_scanner* scanner, int option, uint64_t value);
/* Device bytes held by the scanner's block buffer; yrb_scanner_trim frees it (it is
   re-allocated by the next scan).  Lets a pool of idle scanners bound its HBM use. */
size_t yrb_scanner_footprint(yrb_scanner* scanner);
void yrb_scanner_trim(yrb_scanner* scanner);

/* Scan one block held in host memory: H2D (sliced, overlapped with the scan
   kernel), AC scan, verification, resolve, D2H of the match records. */
int yrb_scan_host(
    yrb_scanner* scanner, const uint8_t* data, size_t size, const yrb_scan_params* params,
    yrb_scan_result* result);

/* Same for a block already in this scanner's device memory. */
int yrb_scan_device(
    yrb_scanner* scanner, const uint8_t* d_data, size_t size, const yrb_scan_params* params,
    yrb_scan_result* result);

/*
 * File batches (BASELINE config 4): many independent buffers --
 * what cli/yara.c:365-457 feeds yr_rules_scan_file one at a time -- scanned with ONE
 * launch sequence instead of ~20 launches and a few synchronisations each.  The items are
 * concatenated in HBM; the Aho-Corasick guard `backtrack <= i` (scanner.c:104) is applied
 * with i counted from the start of the item, which makes the candidates of every item
 * exactly those of a scan that started at the root state there, and the verifiers see
 * each item's own bounds.
 *
 * The batch runs the common, order-independent part of the resolve only.  An item whose
 * result could depend on the rest (a chained string matched in it, a verifier raised an
 * error in it, a string went over YR_MAX_STRING_MATCHES, SCAN_FLAGS_FAST_MODE is set,
 * or the ruleset has zero-length atoms) is flagged YRB_ITEM_RESCAN and carries no
 * matches: the caller scans it with yrb_scan_host, the general path.  For every other
 * item the records equal those yrb_scan_host would return for it with base 0.
 *
 * With YRB_BATCH_ONE_ALLOCATION, items lying in ascending order inside one host
 * allocation (at most 25 % of the span unused) keep their layout on the device and travel
 * in large copies; anything else is packed at 16-byte boundaries with one copy per item.
 * Pinned memory (yrb_host_alloc) is what makes either fast.
 */
typedef struct yrb_batch_item
{
  const uint8_t* data;
  uint64_t size;
} yrb_batch_item;

#define YRB_ITEM_DONE   0
#define YRB_ITEM_RESCAN 1

typedef struct yrb_batch_result
{
  uint32_t n_items;
  uint32_t n_rescan;          /* items flagged YRB_ITEM_RESCAN */
  const uint8_t* item_state;  /* n_items entries */
  const uint64_t* item_first; /* n_items + 1: item k owns matches[item_first[k] .. item_first[k+1]) */
  const yrb_match* matches;   /* per item: string-major, ascending offsets relative to the item */
  yrb_scan_stats stats;
} yrb_batch_result;

#define YRB_BATCH_ONE_ALLOCATION 1 /* all items lie inside ONE host allocation: the bytes between
                                     them are readable and may be copied along */
#define YRB_BATCH_BLOCKS_OF_ONE_SCAN 2 /* the items are consecutive blocks of ONE multi-block scan (scanner.c:467-500):
                                          chained strings connect across them, so an item in which a chain part
                                          matched is flagged YRB_ITEM_RESCAN as well (its result needs the chain
                                          state the blocks before it left; the caller scans it with seeds) */

int yrb_scan_batch(
    yrb_scanner* scanner, const yrb_batch_item* items, uint32_t n_items, int scan_flags, int batch_flags,
    yrb_batch_result* result);

/* Sharded scans: raw (pre-resolve) records of the last defer_resolve scan, as
   opaque 32-byte records in host memory ... */
int yrb_shard_raw(yrb_scanner* scanner, const void** records, size_t* n_records);
/* ... and the global resolve over the concatenated records of all shards. */
int yrb_resolve_raw(
    yrb_scanner* scanner, const void* records, size_t n_records, const yrb_scan_params* params,
    yrb_scan_result* result);

/* Shards scanned by different PROCESSES (one per GPU): records stay on the devices.  A scanner's export
   buffer is created once and its CUDA IPC handle (64 bytes) given to the resolving process, which opens it
   once; after every defer_resolve scan yrb_shard_export_publish copies the shard's records into it (device to
   device) and says how many there are; the resolving process gathers all shards with peer copies and runs the
   one global resolve.  Nothing but the record COUNT crosses the host. */
int yrb_shard_export_create(yrb_scanner* scanner, size_t capacity_records, void* ipc_handle_out, size_t handle_bytes);
void yrb_shard_export_destroy(yrb_scanner* scanner);
int yrb_shard_export_ptr(yrb_scanner* scanner, void** d_ptr); /* the export buffer in the owner's own address space */
int yrb_shard_export_publish(yrb_scanner* scanner, size_t* n_records, int* any_error);
int yrb_shard_import_open(int device, const void* ipc_handle, size_t handle_bytes, void** d_ptr);
int yrb_shard_import_close(int device, void* d_ptr);
int yrb_resolve_raw_device(
    yrb_scanner* scanner, uint32_t n_parts, const void* const* d_parts, const size_t* counts,
    const yrb_scan_params* params, yrb_scan_result* result);

/*
 * ONE block over several GPUs of this node, driven from one process.  Every
 * device holds a replica of the rules image and one scanner; a block is cut into byte-range
 * shards (boundaries at multiples of 4096) that additionally read yrb_multi_margin() bytes either
 * side; each device copies, scans and verifies its shard concurrently (one persistent host thread
 * per device); the raw records are brought to the first device with peer copies and resolved there
 * once.  No collective, no NCCL.  The result equals yrb_scan_host's on one device, bit for bit.
 * Multi-block scans (params->n_seeds > 0) are not sharded: YRB_ERROR_BAD_ARGUMENT.
 */
typedef struct yrb_multi yrb_multi;
int yrb_multi_create(const void* image, size_t image_size, const int* devices, int n_devices, yrb_multi** out);
void yrb_multi_destroy(yrb_multi* multi);
int yrb_multi_device_count(const yrb_multi* multi);
uint64_t yrb_multi_margin(const yrb_multi* multi);
/* the shard plan for a block of block_size bytes: device g reads [g0[g], g1[g]) and owns the
   Aho-Corasick events of [own0[g], own1[g]); arrays of yrb_multi_device_count() entries */
int yrb_multi_plan(
    const yrb_multi* multi, uint64_t block_size, uint64_t* g0, uint64_t* g1, uint64_t* own0, uint64_t* own1);
int yrb_multi_scan_host(
    yrb_multi* multi, const uint8_t* data, size_t size, const yrb_scan_params* params, yrb_scan_result* result);
/* shards already resident: d_shards[g] (memory of device g) holds block bytes [g0[g], g1[g]) */
int yrb_multi_scan_device(
    yrb_multi* multi, const uint8_t* const* d_shards, const uint64_t* g0, const uint64_t* g1, const uint64_t* own0,
    const uint64_t* own1, uint64_t block_size, const yrb_scan_params* params, yrb_scan_result* result);

/* Copies min(match_length, max_match_data) bytes of every match of the last
   yrb_scan_device result into out (packed back to back, offsets[] = start of
   each, offsets[n] = total): the data copy of scan.c:723-737 for callers whose
   block lives in HBM. */
int yrb_gather_match_data(
    yrb_scanner* scanner, const uint8_t* d_data, uint32_t max_match_data, uint8_t* out, size_t out_cap,
    uint64_t* offsets);

/* The same for the last yrb_scan_host result, from the scanner's own resident copy of the block: data[offsets[k] ..
   offsets[k + 1]) are the first min(match_length, max_match_data) bytes of match k (the order of result.matches).
   Both arrays are pinned host memory owned by the scanner, valid until its next call.  YRB_ERROR_BAD_ARGUMENT when
   the last call was not a yrb_scan_host with matches.  The libyara glue uses it for large results: one sequential
   read instead of a cache and TLB miss per match in a multi-GiB host block. */
int yrb_last_match_data(yrb_scanner* scanner, uint32_t max_match_data, const uint8_t** data, const uint64_t** offsets);

/* The atom-prefilter tables are a pure function of the rules image; they are built once per image
   and kept in a process-wide registry and in a disk cache keyed by a hash of the image
   ($YR_B200_CACHE_DIR, default ~/.cache/yara_b200; YR_B200_CACHE=0: no files) -- the device-side
   counterpart of loading compiled rules instead of compiling them (yr_rules_load, rules.c:372-431,
   arena.c:543).  How often yrb_rules_create found them in the registry / on disk / had to build. */
void yrb_filter_cache_stats(unsigned* from_registry, unsigned* from_disk, unsigned* built);

/* Utilities for benchmarks/tests: device memory owned by the scanner's device,
   counter-based corpus bytes (identical to yara_b200.synth.rand_bytes) and
   planting.  They exist so that a 16 GiB corpus never has to cross PCIe. */
int yrb_dev_alloc(int device, size_t bytes, void** d_ptr);
int yrb_dev_free(int device, void* d_ptr);
int yrb_dev_fill_random(int device, void* d_ptr, size_t bytes, uint64_t seed, uint64_t stream_offset);
int yrb_dev_write(int device, void* d_ptr, size_t d_offset, const void* src, size_t bytes);
int yrb_dev_read(int device, const void* d_ptr, size_t d_offset, void* dst, size_t bytes);
int yrb_dev_flush_l2(int device);
/* n byte strings, string k = bytes[starts[k] .. starts[k+1]), written at d_ptr + offsets[k] */
int yrb_dev_scatter(
    int device, void* d_ptr, size_t d_size, uint32_t n, const uint64_t* offsets, const uint64_t* starts,
    const uint8_t* bytes);
/* pinned host memory (cudaHostAlloc) for end-to-end measurements */
int yrb_host_alloc(size_t bytes, void** h_ptr);
int yrb_host_free(void* h_ptr);

#ifdef __cplusplus
}
#endif
#endif
