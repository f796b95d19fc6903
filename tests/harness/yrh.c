/*
 * yrh.c -- parity dump harness (TEST INFRASTRUCTURE).
 *
 * Linked once against oracle/_ref/libyara_ref.so (-> libyrh_ref.so, the oracle
 * side) and once against yara_b200/libyara_b200.so (-> libyrh_b200.so, the
 * product side).  It only uses the public libyara API plus a walk of
 * scanner->matches[] at CALLBACK_MSG_SCAN_FINISHED, when match lists are still
 * alive (scanner.c:555-566) -- SURVEY.md section 7 step 0.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <yara.h>

int yr_b200_rules_flatten(YR_RULES* rules, void** image_out, size_t* size_out);
void yr_b200_image_free(void* image);

typedef struct
{
  int64_t base;
  int64_t offset;
  uint64_t data_hash; /* FNV-1a 64 of the data_length copied bytes */
  uint32_t string_idx;
  int32_t match_length;
  int32_t data_length;
  uint8_t xor_key;
  uint8_t is_private;
  uint16_t pad;
} YRH_MATCH;

typedef struct
{
  YR_RULES* rules;
  YRH_MATCH* out;
  size_t cap;
  size_t n;
  uint8_t* rule_bits;
  int too_many_action; /* what to return on CALLBACK_MSG_TOO_MANY_MATCHES */
  int too_many_calls;
  uint32_t too_many_first; /* string idx of the first TOO_MANY callback */
} YRH_CTX;

static uint64_t fnv1a(const uint8_t* p, size_t n)
{
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++)
  {
    h ^= p[i];
    h *= 1099511628211ULL;
  }
  return h;
}

static int yrh_callback(YR_SCAN_CONTEXT* context, int message, void* message_data, void* user_data)
{
  YRH_CTX* c = (YRH_CTX*) user_data;

  if (message == CALLBACK_MSG_RULE_MATCHING)
  {
    YR_RULE* rule = (YR_RULE*) message_data;
    if (c->rule_bits)
      c->rule_bits[rule - c->rules->rules_table] = 1;
  }
  else if (message == CALLBACK_MSG_TOO_MANY_MATCHES)
  {
    if (c->too_many_calls == 0)
      c->too_many_first = ((YR_STRING*) message_data)->idx;
    c->too_many_calls++;
    return c->too_many_action;
  }
  else if (message == CALLBACK_MSG_SCAN_FINISHED)
  {
    for (uint32_t i = 0; i < c->rules->num_strings; i++)
    {
      for (YR_MATCH* m = context->matches[i].head; m != NULL; m = m->next)
      {
        if (c->n < c->cap)
        {
          YRH_MATCH* o = &c->out[c->n];
          o->base = m->base;
          o->offset = m->offset;
          o->string_idx = i;
          o->match_length = m->match_length;
          o->data_length = m->data_length;
          o->xor_key = m->xor_key;
          o->is_private = m->is_private;
          o->pad = 0;
          o->data_hash = fnv1a(m->data, m->data_length > 0 ? (size_t) m->data_length : 0);
        }
        c->n++;
      }
    }
  }

  return CALLBACK_CONTINUE;
}

int yrh_initialize(void) { return yr_initialize(); }
int yrh_finalize(void) { return yr_finalize(); }

static void yrh_compile_error(
    int level, const char* file, int line, const YR_RULE* rule, const char* msg, void* user)
{
  char* buf = (char*) user;
  if (level == YARA_ERROR_LEVEL_ERROR && buf && buf[0] == 0)
    snprintf(buf, 256, "line %d: %s", line, msg);
}

int yrh_compile(const char* text, void** rules_out, char* errbuf)
{
  YR_COMPILER* compiler = NULL;
  YR_RULES* rules = NULL;
  int rc;

  if (errbuf) errbuf[0] = 0;

  rc = yr_compiler_create(&compiler);
  if (rc != ERROR_SUCCESS) return rc;

  yr_compiler_set_callback(compiler, yrh_compile_error, errbuf);

  if (yr_compiler_add_string(compiler, text, NULL) != 0)
  {
    yr_compiler_destroy(compiler);
    return ERROR_SYNTAX_ERROR;
  }

  rc = yr_compiler_get_rules(compiler, &rules);
  yr_compiler_destroy(compiler);

  if (rc == ERROR_SUCCESS)
    *rules_out = rules;

  return rc;
}

void yrh_rules_destroy(void* rules) { yr_rules_destroy((YR_RULES*) rules); }

void yrh_rules_info(void* rules_, uint32_t* num_strings, uint32_t* num_rules, uint32_t* tables_size)
{
  YR_RULES* rules = (YR_RULES*) rules_;
  YR_RULES_STATS stats;
  *num_strings = rules->num_strings;
  *num_rules = rules->num_rules;
  yr_rules_get_stats(rules, &stats);
  *tables_size = stats.ac_tables_size;
}

int yrh_set_max_match_data(uint32_t v) { return yr_set_configuration(YR_CONFIG_MAX_MATCH_DATA, &v); }

int yrh_flatten(void* rules, void** image, size_t* size)
{
  return yr_b200_rules_flatten((YR_RULES*) rules, image, size);
}

void yrh_free_image(void* image) { yr_b200_image_free(image); }

/* One single-block scan through the public API; dumps every string's match list. */
int yrh_scan(
    void* rules_, const uint8_t* buf, size_t size, int flags, int timeout, int too_many_action,
    YRH_MATCH* out, size_t cap, size_t* n_out, uint8_t* rule_bits, int* too_many_calls)
{
  YRH_CTX c;
  memset(&c, 0, sizeof(c));
  c.rules = (YR_RULES*) rules_;
  c.out = out;
  c.cap = cap;
  c.rule_bits = rule_bits;
  c.too_many_action = too_many_action;

  if (rule_bits)
    memset(rule_bits, 0, c.rules->num_rules);

  int rc = yr_rules_scan_mem(c.rules, buf, size, flags, yrh_callback, &c, timeout);

  *n_out = c.n;
  if (too_many_calls) *too_many_calls = c.too_many_calls;
  return rc;
}

/* yr_scanner_scan_mem with a time budget in NANOSECONDS (the public API takes whole seconds,
   scanner.c:371-374: timeout * 1000000000): tests of ERROR_SCAN_TIMEOUT that do not need
   multi-second scans */
int yrh_scan_timeout_ns(void* rules_, const uint8_t* buf, size_t size, uint64_t timeout_ns)
{
  YR_SCANNER* scanner = NULL;
  int rc = yr_scanner_create((YR_RULES*) rules_, &scanner);
  if (rc != ERROR_SUCCESS) return rc;
  YRH_CTX c;
  memset(&c, 0, sizeof(c));
  c.rules = (YR_RULES*) rules_;
  yr_scanner_set_callback(scanner, yrh_callback, &c);
  scanner->timeout = timeout_ns;
  rc = yr_scanner_scan_mem(scanner, buf, size);
  yr_scanner_destroy(scanner);
  return rc;
}

/* ---- multi-block scan: the same fixed-size / overlapping block iterator the
   reference's own tests use (tests/util.c:55-293), restated. ---- */
typedef struct
{
  const uint8_t* buf;
  size_t size, block_size, overlap, pos;
  YR_MEMORY_BLOCK block;
  size_t slow_from;      /* fetching a block at or behind this offset takes slow_us (0: never) */
  unsigned slow_us;
} YRH_ITER;

static const uint8_t* yrh_fetch(YR_MEMORY_BLOCK* b)
{
  YRH_ITER* it = (YRH_ITER*) b->context;
  if (it->slow_us != 0 && b->base >= it->slow_from)
    usleep(it->slow_us);
  return it->buf + b->base;
}

static YR_MEMORY_BLOCK* yrh_iter_at(YRH_ITER* it)
{
  if (it->pos >= it->size && !(it->size == 0 && it->pos == 0))
    return NULL;
  size_t n = it->size - it->pos;
  if (n > it->block_size) n = it->block_size;
  it->block.base = it->pos;
  it->block.size = n;
  it->block.context = it;
  it->block.fetch_data = yrh_fetch;
  return &it->block;
}

static YR_MEMORY_BLOCK* yrh_first(YR_MEMORY_BLOCK_ITERATOR* iter)
{
  YRH_ITER* it = (YRH_ITER*) iter->context;
  it->pos = 0;
  iter->last_error = ERROR_SUCCESS;
  return yrh_iter_at(it);
}

static YR_MEMORY_BLOCK* yrh_next(YR_MEMORY_BLOCK_ITERATOR* iter)
{
  YRH_ITER* it = (YRH_ITER*) iter->context;
  iter->last_error = ERROR_SUCCESS;
  if (it->pos + it->block_size >= it->size)
    return NULL;
  it->pos += it->block_size - it->overlap;
  return yrh_iter_at(it);
}

static uint64_t yrh_file_size(YR_MEMORY_BLOCK_ITERATOR* iter)
{
  return ((YRH_ITER*) iter->context)->size;
}

int yrh_scan_blocks(
    void* rules_, const uint8_t* buf, size_t size, size_t block_size, size_t overlap, int flags,
    int too_many_action, YRH_MATCH* out, size_t cap, size_t* n_out, uint8_t* rule_bits,
    int* too_many_calls)
{
  YRH_CTX c;
  YRH_ITER it;
  YR_MEMORY_BLOCK_ITERATOR iter;
  YR_SCANNER* scanner;

  memset(&c, 0, sizeof(c));
  memset(&it, 0, sizeof(it));
  memset(&iter, 0, sizeof(iter));

  c.rules = (YR_RULES*) rules_;
  c.out = out;
  c.cap = cap;
  c.rule_bits = rule_bits;
  c.too_many_action = too_many_action;

  if (rule_bits)
    memset(rule_bits, 0, c.rules->num_rules);

  it.buf = buf;
  it.size = size;
  it.block_size = block_size;
  it.overlap = overlap;
  iter.context = &it;
  iter.first = yrh_first;
  iter.next = yrh_next;
  iter.file_size = yrh_file_size;

  int rc = yr_scanner_create(c.rules, &scanner);
  if (rc != ERROR_SUCCESS) return rc;

  yr_scanner_set_callback(scanner, yrh_callback, &c);
  yr_scanner_set_flags(scanner, flags);
  rc = yr_scanner_scan_mem_blocks(scanner, &iter);
  yr_scanner_destroy(scanner);

  *n_out = c.n;
  if (too_many_calls) *too_many_calls = c.too_many_calls;
  return rc;
}

/* ---- multi-block scan over an arbitrary list of regions: block k is `sizes[k]` bytes at buf + offs[k], reported with
   base bases[k] -- any order, gaps, overlaps, repeats, empty blocks -- the shapes process-memory and custom iterators
   produce (proc/linux.c:145,309). ---- */
typedef struct
{
  const uint8_t* buf;
  const size_t* offs;
  const size_t* sizes;
  const uint64_t* bases;
  size_t n, k;
  YR_MEMORY_BLOCK block;
} YRH_REGIONS;

static const uint8_t* yrh_regions_fetch(YR_MEMORY_BLOCK* b)
{
  YRH_REGIONS* it = (YRH_REGIONS*) b->context;
  return it->buf + it->offs[it->k];
}

static YR_MEMORY_BLOCK* yrh_regions_at(YRH_REGIONS* it)
{
  if (it->k >= it->n) return NULL;
  it->block.base = it->bases[it->k];
  it->block.size = it->sizes[it->k];
  it->block.context = it;
  it->block.fetch_data = yrh_regions_fetch;
  return &it->block;
}

static YR_MEMORY_BLOCK* yrh_regions_first(YR_MEMORY_BLOCK_ITERATOR* iter)
{
  YRH_REGIONS* it = (YRH_REGIONS*) iter->context;
  it->k = 0;
  iter->last_error = ERROR_SUCCESS;
  return yrh_regions_at(it);
}

static YR_MEMORY_BLOCK* yrh_regions_next(YR_MEMORY_BLOCK_ITERATOR* iter)
{
  YRH_REGIONS* it = (YRH_REGIONS*) iter->context;
  iter->last_error = ERROR_SUCCESS;
  it->k++;
  return yrh_regions_at(it);
}

int yrh_scan_regions(
    void* rules_, const uint8_t* buf, size_t n, const size_t* offs, const size_t* sizes, const uint64_t* bases, int flags,
    YRH_MATCH* out, size_t cap, size_t* n_out, uint8_t* rule_bits)
{
  YRH_CTX c;
  YRH_REGIONS it;
  YR_MEMORY_BLOCK_ITERATOR iter;
  YR_SCANNER* scanner;

  memset(&c, 0, sizeof(c));
  memset(&it, 0, sizeof(it));
  memset(&iter, 0, sizeof(iter));
  c.rules = (YR_RULES*) rules_;
  c.out = out;
  c.cap = cap;
  c.rule_bits = rule_bits;
  if (rule_bits)
    memset(rule_bits, 0, c.rules->num_rules);
  it.buf = buf;
  it.offs = offs;
  it.sizes = sizes;
  it.bases = bases;
  it.n = n;
  iter.context = &it;
  iter.first = yrh_regions_first;
  iter.next = yrh_regions_next;

  int rc = yr_scanner_create(c.rules, &scanner);
  if (rc != ERROR_SUCCESS) return rc;
  yr_scanner_set_callback(scanner, yrh_callback, &c);
  yr_scanner_set_flags(scanner, flags);
  rc = yr_scanner_scan_mem_blocks(scanner, &iter);
  yr_scanner_destroy(scanner);
  *n_out = c.n;
  return rc;
}

/* ---- ONE scanner, two multi-block scans: the first runs out of its time budget part-way (the iterator turns slow
   at `slow_from`, scanner.c:74-81), the second scans other bytes to the end.  Nothing of the first scan -- blocks the
   drop-in had deferred, match lists -- may show up in the second (scanner.c:540-552 cleans up behind every scan). */
int yrh_scan_blocks_after_timed_out_scan(
    void* rules_, const uint8_t* buf1, size_t size1, size_t slow_from, uint64_t timeout_ns, const uint8_t* buf2, size_t size2,
    size_t block_size, YRH_MATCH* out, size_t cap, size_t* n_out, uint8_t* rule_bits, int* rc_first)
{
  YRH_CTX c;
  YRH_ITER it;
  YR_MEMORY_BLOCK_ITERATOR iter;
  YR_SCANNER* scanner;

  memset(&c, 0, sizeof(c));
  memset(&it, 0, sizeof(it));
  memset(&iter, 0, sizeof(iter));
  c.rules = (YR_RULES*) rules_;
  iter.context = &it;
  iter.first = yrh_first;
  iter.next = yrh_next;
  iter.file_size = yrh_file_size;

  int rc = yr_scanner_create(c.rules, &scanner);
  if (rc != ERROR_SUCCESS) return rc;
  yr_scanner_set_callback(scanner, yrh_callback, &c);

  it.buf = buf1;
  it.size = size1;
  it.block_size = block_size;
  it.slow_from = slow_from;
  it.slow_us = (unsigned) (timeout_ns / 1000) * 4 + 2000;
  scanner->timeout = timeout_ns;
  *rc_first = yr_scanner_scan_mem_blocks(scanner, &iter);

  memset(&c, 0, sizeof(c));
  c.rules = (YR_RULES*) rules_;
  c.out = out;
  c.cap = cap;
  c.rule_bits = rule_bits;
  if (rule_bits)
    memset(rule_bits, 0, c.rules->num_rules);
  yr_scanner_set_callback(scanner, yrh_callback, &c);
  it.buf = buf2;
  it.size = size2;
  it.slow_us = 0;
  scanner->timeout = 0;
  rc = yr_scanner_scan_mem_blocks(scanner, &iter);
  yr_scanner_destroy(scanner);
  *n_out = c.n;
  return rc;
}

/* ---- ONE YR_SCANNER (yr_scanner_create ... yr_scanner_destroy) over a sequence of scans, the way a long-lived
   consumer of the scanner API uses it: scan k is yr_scanner_scan_mem (modes[k] == 0) or yr_scanner_scan_mem_blocks
   with blocks of modes[k] bytes and a quarter of that as overlap, with flags[k] set just before.  All matches go to
   `out` one scan after the other (n_per_scan[k] of them), rule verdicts to rule_bits[k * num_rules ...]. */
int yrh_scanner_sequence(
    void* rules_, int n, const uint8_t** bufs, const size_t* sizes, const int* modes, const int* flags, YRH_MATCH* out,
    size_t cap, size_t* n_per_scan, uint8_t* rule_bits, int* rcs)
{
  YRH_CTX c;
  YRH_ITER it;
  YR_MEMORY_BLOCK_ITERATOR iter;
  YR_SCANNER* scanner;
  size_t used = 0;

  memset(&iter, 0, sizeof(iter));
  iter.context = &it;
  iter.first = yrh_first;
  iter.next = yrh_next;
  iter.file_size = yrh_file_size;

  int rc = yr_scanner_create((YR_RULES*) rules_, &scanner);
  if (rc != ERROR_SUCCESS) return rc;

  for (int k = 0; k < n; k++)
  {
    memset(&c, 0, sizeof(c));
    c.rules = (YR_RULES*) rules_;
    c.out = out + used;
    c.cap = cap - used;
    c.rule_bits = rule_bits + (size_t) k * c.rules->num_rules;
    memset(c.rule_bits, 0, c.rules->num_rules);
    yr_scanner_set_callback(scanner, yrh_callback, &c);
    yr_scanner_set_flags(scanner, flags[k]);
    if (modes[k] == 0)
    {
      rcs[k] = yr_scanner_scan_mem(scanner, bufs[k], sizes[k]);
    }
    else
    {
      memset(&it, 0, sizeof(it));
      it.buf = bufs[k];
      it.size = sizes[k];
      it.block_size = (size_t) modes[k];
      it.overlap = (size_t) modes[k] / 4;
      rcs[k] = yr_scanner_scan_mem_blocks(scanner, &iter);
    }
    n_per_scan[k] = c.n;
    used += c.n < c.cap ? c.n : c.cap;
  }
  yr_scanner_destroy(scanner);
  return ERROR_SUCCESS;
}

/* ---- CPU baseline: N scanner threads over one shared YR_RULES, one contiguous
   slice each -- the structure `yara -p N` uses (cli/yara.c:1561-1610). ---- */
typedef struct
{
  YR_RULES* rules;
  const uint8_t* buf;
  size_t size;
  int flags;
  int rc;
} YRH_SLICE;

static int yrh_null_callback(YR_SCAN_CONTEXT* c, int m, void* d, void* u) { return CALLBACK_CONTINUE; }

static void* yrh_slice_thread(void* arg)
{
  YRH_SLICE* s = (YRH_SLICE*) arg;
  s->rc = yr_rules_scan_mem(s->rules, s->buf, s->size, s->flags, yrh_null_callback, NULL, 0);
  return NULL;
}

double yrh_scan_mt(void* rules_, const uint8_t* buf, size_t size, int nthreads, int flags, int* rc_out)
{
  pthread_t* th = (pthread_t*) calloc((size_t) nthreads, sizeof(pthread_t));
  YRH_SLICE* sl = (YRH_SLICE*) calloc((size_t) nthreads, sizeof(YRH_SLICE));
  struct timespec t0, t1;
  size_t per = size / (size_t) nthreads;
  int rc = 0;

  clock_gettime(CLOCK_MONOTONIC, &t0);

  for (int i = 0; i < nthreads; i++)
  {
    sl[i].rules = (YR_RULES*) rules_;
    sl[i].buf = buf + per * (size_t) i;
    sl[i].size = (i == nthreads - 1) ? size - per * (size_t) i : per;
    sl[i].flags = flags;
    pthread_create(&th[i], NULL, yrh_slice_thread, &sl[i]);
  }

  for (int i = 0; i < nthreads; i++)
  {
    pthread_join(th[i], NULL);
    if (sl[i].rc != 0) rc = sl[i].rc;
  }

  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(sl);

  if (rc_out) *rc_out = rc;
  return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

#ifdef YRH_B200
/* file batches through the drop-in's glue (yara_b200/host/yr_b200.h) */
typedef struct YR_B200_BATCH YR_B200_BATCH;
int yr_b200_batch_scan(
    YR_RULES* rules, const uint8_t* const* buffers, const size_t* sizes, uint32_t n, int flags, int one_allocation,
    YR_B200_BATCH** batch);
void yr_b200_batch_destroy(YR_B200_BATCH* batch);
uint32_t yr_b200_batch_rescans(const YR_B200_BATCH* batch);

int yrh_batch_begin(
    void* rules, const uint8_t* const* buffers, const size_t* sizes, uint32_t n, int flags, int one_allocation,
    void** batch, uint32_t* n_rescans)
{
  YR_B200_BATCH* b = NULL;
  int rc = yr_b200_batch_scan((YR_RULES*) rules, buffers, sizes, n, flags, one_allocation, &b);
  *batch = b;
  if (rc == 0 && n_rescans != NULL) *n_rescans = yr_b200_batch_rescans(b);
  return rc;
}

void yrh_batch_end(void* batch) { yr_b200_batch_destroy((YR_B200_BATCH*) batch); }

uint64_t yr_b200_batch_hits(void);
uint64_t yrh_batch_hits(void) { return yr_b200_batch_hits(); }

void yr_b200_coalesce_stats(uint64_t* batches, uint64_t* scans);
void yrh_coalesce_stats(uint64_t* batches, uint64_t* scans) { yr_b200_coalesce_stats(batches, scans); }
#endif

#if defined(YRH_B200) || defined(YRH_PREFILTER)
/* condition-VM pre-filter counters (yara_b200/host/yr_b200_prefilter.c) */
void yr_b200_prefilter_stats(uint64_t* evaluated, uint64_t* skipped);
void yr_b200_prefilter_classes(YR_RULES* rules, uint32_t counts[5]);
int yr_b200_prefilter_enabled(void);

int yrh_prefilter_enabled(void) { return yr_b200_prefilter_enabled(); }
void yrh_prefilter_stats(uint64_t* evaluated, uint64_t* skipped) { yr_b200_prefilter_stats(evaluated, skipped); }
void yrh_prefilter_classes(void* rules, uint32_t* counts) { yr_b200_prefilter_classes((YR_RULES*) rules, counts); }
#endif
