/*
 * yrb_image.h -- the flat "rules image": every table the byte-scanning hot
 * path reads, with all pointers turned into indices/offsets, in ONE contiguous
 * relocatable blob.  It is what crosses the C-ABI (include/yara_b200.h) and
 * what gets uploaded to HBM verbatim.
 *
 * Producer: yara_b200/host/yr_b200_flatten.c (reads a compiled YR_RULES).
 * Consumers: the CUDA engine (yara_b200/csrc), the CPU oracle restatement
 * (oracle/yr_oracle.c), the golden fixtures under tests/golden/.
 *
 * What each section mirrors in the reference (libyara 4.2.1):
 *   trans   <- rules->ac_transition_table  (include/yara/ahocorasick.h:39-51,
 *              built at ahocorasick.c:525-643).  u32 slot = target<<9 | index,
 *              index 0 = failure link, 1..256 = byte+1.
 *   match   <- rules->ac_match_table (types.h:603-611): 1-based index of the
 *              first YR_AC_MATCH of the state, 0 = none.  Only tables_size
 *              u32 are meaningful (the arena buffer is over-allocated,
 *              ahocorasick.c:448-449).
 *   acm     <- rules->ac_match_pool, YR_AC_MATCH (types.h:324-344).
 *   strings <- rules->strings_table, YR_STRING (types.h:238-287).
 *   lit     <- string->string bytes of STRING_FLAGS_LITERAL strings (YR_SZ_POOL)
 *   code    <- arena buffer YR_RE_CODE_SECTION (compiler.h:64), byte-identical,
 *              so forward_code/backward_code become offsets into it.
 *   lower / alter <- yr_lowercase[], yr_altercase[] (libyara.c:245-258); taken
 *              from the host process because tolower() is locale dependent.
 */
#ifndef YRB_IMAGE_H
#define YRB_IMAGE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YRB_IMAGE_MAGIC   0x31425259u /* "YRB1" */
#define YRB_IMAGE_VERSION 2u
#define YRB_NONE          0xFFFFFFFFu

/* limits.h of the reference, baked in (SURVEY.md section 5 "Config / flags") */
#define YRB_MAX_ATOM_LENGTH        4
#define YRB_MAX_STRING_MATCHES     1000000
#define YRB_RE_SCAN_LIMIT          4096
#define YRB_RE_MAX_FIBERS          1024
#define YRB_RE_MAX_SPLIT_ID        128
#define YRB_DEFAULT_MAX_MATCH_DATA 512

/* STRING_FLAGS_* (types.h:66-89) */
#define YRB_SF_HEXADECIMAL   0x02u
#define YRB_SF_NO_CASE       0x04u
#define YRB_SF_ASCII         0x08u
#define YRB_SF_WIDE          0x10u
#define YRB_SF_REGEXP        0x20u
#define YRB_SF_FAST_REGEXP   0x40u
#define YRB_SF_FULL_WORD     0x80u
#define YRB_SF_SINGLE_MATCH  0x200u
#define YRB_SF_LITERAL       0x400u
#define YRB_SF_FITS_IN_ATOM  0x800u
#define YRB_SF_CHAIN_PART    0x2000u
#define YRB_SF_CHAIN_TAIL    0x4000u
#define YRB_SF_FIXED_OFFSET  0x8000u
#define YRB_SF_GREEDY_REGEXP 0x10000u
#define YRB_SF_DOT_ALL       0x20000u
#define YRB_SF_XOR           0x80000u
#define YRB_SF_PRIVATE       0x100000u
#define YRB_SF_BASE64        0x200000u
#define YRB_SF_BASE64_WIDE   0x400000u

/* SCAN_FLAGS_* (scan.h:38-42) */
#define YRB_SCAN_FAST_MODE 1

/* error codes (error.h:40-109) */
#define YRB_SUCCESS                  0
#define YRB_ERROR_INSUFFICIENT_MEMORY 1
#define YRB_ERROR_SCAN_TIMEOUT        26
#define YRB_ERROR_TOO_MANY_MATCHES    30
#define YRB_ERROR_INTERNAL_FATAL      31
#define YRB_ERROR_TOO_MANY_RE_FIBERS  46

typedef struct yrb_image_header
{
  uint32_t magic;
  uint32_t version;
  uint64_t total_size;    /* bytes, header included */
  uint32_t tables_size;   /* AC slots (u32 each) in trans[] and match[] */
  uint32_t num_acm;       /* YR_AC_MATCH entries */
  uint32_t num_strings;
  uint32_t num_rules;
  uint32_t lit_size;      /* bytes in lit[] */
  uint32_t code_size;     /* bytes in code[] */
  uint32_t max_re_stack;  /* deepest REPEAT_START nesting in code[] (fiber stack bound) */
  uint32_t num_chains;    /* number of chain heads (strings with CHAIN_PART and no chained_to) */
  uint64_t off_trans;     /* all offsets from the start of the blob, 16-byte aligned */
  uint64_t off_match;
  uint64_t off_acm;
  uint64_t off_strings;
  uint64_t off_lit;
  uint64_t off_code;
  uint64_t off_lower;     /* 256 bytes */
  uint64_t off_alter;     /* 256 bytes */
} yrb_image_header;

/* YR_AC_MATCH (types.h:324-344), 20 bytes */
typedef struct yrb_acm
{
  uint32_t string_idx;
  uint32_t fwd_code;  /* offset in code[], YRB_NONE if NULL */
  uint32_t bwd_code;  /* offset in code[], YRB_NONE if NULL */
  uint32_t next;      /* 1-based index of next entry in acm[], 0 = end of chain */
  uint32_t backtrack; /* u16 in the reference */
} yrb_acm;

/* YR_STRING (types.h:238-287), 40 bytes */
typedef struct yrb_string
{
  uint32_t flags;
  int32_t length;
  uint32_t lit_off;    /* offset in lit[] (LITERAL strings), else YRB_NONE */
  uint32_t chained_to; /* string idx, YRB_NONE if NULL */
  int32_t chain_gap_min;
  int32_t chain_gap_max;
  int64_t fixed_offset;
  uint32_t rule_idx;
  uint32_t chain_head; /* idx of the chain's head string (== own idx for heads), YRB_NONE if not CHAIN_PART */
} yrb_string;

static inline const uint32_t* yrb_img_trans(const void* img)
{ return (const uint32_t*) ((const uint8_t*) img + ((const yrb_image_header*) img)->off_trans); }
static inline const uint32_t* yrb_img_match(const void* img)
{ return (const uint32_t*) ((const uint8_t*) img + ((const yrb_image_header*) img)->off_match); }
static inline const yrb_acm* yrb_img_acm(const void* img)
{ return (const yrb_acm*) ((const uint8_t*) img + ((const yrb_image_header*) img)->off_acm); }
static inline const yrb_string* yrb_img_strings(const void* img)
{ return (const yrb_string*) ((const uint8_t*) img + ((const yrb_image_header*) img)->off_strings); }
static inline const uint8_t* yrb_img_lit(const void* img)
{ return (const uint8_t*) img + ((const yrb_image_header*) img)->off_lit; }
static inline const uint8_t* yrb_img_code(const void* img)
{ return (const uint8_t*) img + ((const yrb_image_header*) img)->off_code; }
static inline const uint8_t* yrb_img_lower(const void* img)
{ return (const uint8_t*) img + ((const yrb_image_header*) img)->off_lower; }
static inline const uint8_t* yrb_img_alter(const void* img)
{ return (const uint8_t*) img + ((const yrb_image_header*) img)->off_alter; }

/*
 * One resolved match, as the engine (and the oracle restatement) hand it back:
 * exactly the fields scan.c:714-756 would have put into a YR_MATCH, minus the
 * data copy, which the caller takes from its own block buffer.  The records of
 * one string are contiguous and in YR_MATCHES list order (ascending offset,
 * unique offsets; scan.c:290-368).
 */
typedef struct yrb_match
{
  int64_t offset;       /* block-relative, like YR_MATCH.offset */
  uint32_t string_idx;
  int32_t match_length;
  uint8_t xor_key;      /* YRB_MATCH_FROM_SEED: bits 16-21 of the seed index instead (the node keeps its own key) */
  uint8_t flags;        /* YRB_MATCH_* */
  uint16_t seed;        /* YRB_MATCH_FROM_SEED: bits 0-15 of the index of the yrb_seed this node was */
  int32_t chain_length; /* only meaningful for unconfirmed chain matches */
} yrb_match;

#define YRB_MATCH_UNCONFIRMED 1 /* chain part awaiting its tail */
#define YRB_MATCH_FROM_SEED   2 /* a node carried over from an earlier block (yrb_scan_params.seeds) */
#define YRB_MAX_SEEDS         0x3FFFFFu
#define YRB_MATCH_SEED_INDEX(m) ((uint32_t) (m)->seed | ((uint32_t) ((m)->xor_key & 0x3F) << 16))

#ifdef __cplusplus
}
#endif
#endif
