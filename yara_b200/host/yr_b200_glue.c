/*
 * yr_b200_glue.c -- libyara-side glue of the B200 scan path ("libyara/cuda" of
 * BASELINE.json's north star).  It is the ONLY code that knows both worlds:
 * YR_SCANNER / YR_RULES / YR_MATCH on one side, the plain C ABI of
 * include/yara_b200.h on the other.
 *
 *   yr_b200_scan_mem_block   replaces _yr_scanner_scan_mem_block (scanner.c:45-176)
 *                            at its single call site (scanner.c:493-496)
 *   materialise()            builds the YR_MATCH nodes scan.c:714-756 would have
 *                            built, from the notebook, and links them into
 *                            scanner->matches[] (list_insert: the ordering and
 *                            dedup-by-address of scan.c:290-368, so that multi-block
 *                            scans merge exactly like the reference)
 *   list_full()              CALLBACK_MSG_TOO_MANY_MATCHES as scan.c:1054-1073
 *                            issues it, on the caller's thread
 *
 * There is no CPU fallback: if the CUDA engine cannot be used the block scan
 * fails with ERROR_INTERNAL_FATAL_ERROR and says why on stderr.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <yara.h>
#include <yara/bitmask.h>
#include <yara/notebook.h>
#include <yara/stopwatch.h>

#include "yara_b200.h"
#include "yr_b200.h"

typedef struct rules_entry
{
  YR_RULES* rules;
  yrb_rules* dev;
  int device;
  yrb_multi* multi;            /* all configured GPUs ($YR_B200_DEVICES), created on the first large block */
  pthread_mutex_t multi_lock;  /* one sharded scan (and its materialisation) at a time */
  uint32_t* chain_strings; /* idx of every STRING_FLAGS_CHAIN_PART string */
  uint32_t n_chain_strings;
  int has_fixed_offset;    /* some string is only verified at one absolute offset (scan.c:1025-1027: base + offset) */
  struct rules_entry* next;
} rules_entry;

/* a small block of a multi-block scan whose scan is deferred (see blocks_flush below) */
typedef struct pend_block
{
  uint64_t base;
  size_t size;
  size_t off; /* of its bytes in scanner_entry.pend_data */
} pend_block;

typedef struct scanner_entry
{
  YR_SCANNER* scanner;
  yrb_scanner* dev;
  yrb_rules* dev_rules;
  rules_entry* re;
  struct scanner_entry* next;
  /* deferred blocks: copies of their bytes, packed at 16-byte boundaries, in scan order */
  uint8_t* pend_data;
  size_t pend_cap, pend_used;
  pend_block* pend;
  uint32_t n_pend, cap_pend;
  const void* pend_scan; /* scanner->matches_notebook of the scan they belong to (a new scan gets a new notebook) */
} scanner_entry;

/* Device scanners (stream, block buffer, work areas, pinned result buffers) are
   expensive to create and yr_rules_scan_mem creates a YR_SCANNER per call
   (rules.c:196): released ones are parked here and handed to the next scanner of
   the same rules. */
#define YR_B200_POOL_MAX 32
typedef struct pool_entry
{
  yrb_scanner* dev;
  yrb_rules* dev_rules;
} pool_entry;
static pool_entry g_pool[YR_B200_POOL_MAX];
static int g_pool_n = 0;
/* idle scanners may keep at most this many bytes of block buffers between them */
#define YR_B200_POOL_BYTES ((size_t) 48 << 30)

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static rules_entry* g_rules = NULL;
static scanner_entry* g_scanners = NULL;
static int g_device = -1;

void yr_b200_set_device(int device) { g_device = device; }

static int current_device(void)
{
  if (g_device >= 0)
    return g_device;
  const char* e = getenv("YR_B200_DEVICE");
  return e ? atoi(e) : 0;
}

/* ---- several GPUs for ONE block: $YR_B200_DEVICES = "0,1,2,3" or "all".  A block of
   at least $YR_B200_MULTI_MIN_MIB (default 256) MiB that is the only block of its scan is cut into
   byte-range shards, one per device (yrb_multi_scan_host); everything else stays on one device. */
static int g_multi_n = -1;
static int g_multi_dev[64];

static int multi_devices(const int** devs)
{
  if (g_multi_n < 0)
  {
    int n = 0;
    const char* e = getenv("YR_B200_DEVICES");
    if (e != NULL && strcmp(e, "all") == 0)
    {
      n = yrb_device_count();
      if (n > 64) n = 64;
      for (int i = 0; i < n; i++) g_multi_dev[i] = i;
    }
    else if (e != NULL)
    {
      while (*e != 0 && n < 64)
      {
        char* end = NULL;
        long v = strtol(e, &end, 10);
        if (end == e) break;
        int dup = 0;
        for (int i = 0; i < n; i++) dup |= g_multi_dev[i] == (int) v;
        if (!dup && v >= 0) g_multi_dev[n++] = (int) v;
        e = (*end == ',') ? end + 1 : end;
      }
    }
    g_multi_n = n;
  }
  *devs = g_multi_dev;
  return g_multi_n;
}

static size_t multi_min_bytes(void)
{
  static size_t v = 0;
  if (v == 0)
  {
    const char* e = getenv("YR_B200_MULTI_MIN_MIB");
    long mib = e ? atol(e) : 256;
    v = (size_t) (mib < 1 ? 1 : mib) << 20;
  }
  return v;
}

static int fail_loudly(const char* what, int rc)
{
  fprintf(stderr, "yara_b200: %s failed (code %d); the B200 scan path has no CPU fallback\n", what, rc);
  return ERROR_INTERNAL_FATAL_ERROR;
}

/* rules upload, lazily on first scan, shared by all scanners of this YR_RULES */
static int get_rules(YR_RULES* rules, rules_entry** out)
{
  int device = current_device();
  rules_entry* e;

  for (e = g_rules; e != NULL; e = e->next)
  {
    if (e->rules == rules && e->device == device)
    {
      *out = e;
      return ERROR_SUCCESS;
    }
  }

  void* image = NULL;
  size_t image_size = 0;
  struct timespec ta, tb, tc;
  clock_gettime(CLOCK_MONOTONIC, &ta);
  int rc = yr_b200_rules_flatten(rules, &image, &image_size);
  if (rc != ERROR_SUCCESS)
    return rc;
  clock_gettime(CLOCK_MONOTONIC, &tb);

  yrb_rules* dev = NULL;
  rc = yrb_rules_create(image, image_size, device, &dev);
  yr_b200_image_free(image);
  clock_gettime(CLOCK_MONOTONIC, &tc);
  if (getenv("YRB_TIMING"))
    fprintf(stderr, "yara_b200: rules of %u strings: flatten %.1f ms, upload (CUDA context on first use, image copy, filter tables) %.1f ms\n",
            rules->num_strings, (tb.tv_sec - ta.tv_sec) * 1e3 + (tb.tv_nsec - ta.tv_nsec) * 1e-6,
            (tc.tv_sec - tb.tv_sec) * 1e3 + (tc.tv_nsec - tb.tv_nsec) * 1e-6);
  if (rc != YRB_SUCCESS)
    return fail_loudly("rules upload (yrb_rules_create)", rc);

  e = (rules_entry*) calloc(1, sizeof(*e));
  if (e == NULL)
  {
    yrb_rules_release(dev);
    return ERROR_INSUFFICIENT_MEMORY;
  }
  e->rules = rules;
  e->dev = dev;
  e->device = device;
  pthread_mutex_init(&e->multi_lock, NULL);

  for (uint32_t i = 0; i < rules->num_strings; i++)
  {
    if (STRING_IS_CHAIN_PART(&rules->strings_table[i]))
      e->n_chain_strings++;
    if (STRING_IS_FIXED_OFFSET(&rules->strings_table[i]))
      e->has_fixed_offset = 1;
  }

  if (e->n_chain_strings > 0)
  {
    e->chain_strings = (uint32_t*) malloc(sizeof(uint32_t) * e->n_chain_strings);
    if (e->chain_strings == NULL)
    {
      /* without this list multi-block chain matches would be resolved without their seeds: fail instead */
      yrb_rules_release(dev);
      free(e);
      return ERROR_INSUFFICIENT_MEMORY;
    }
    uint32_t k = 0;
    for (uint32_t i = 0; i < rules->num_strings; i++)
      if (STRING_IS_CHAIN_PART(&rules->strings_table[i]))
        e->chain_strings[k++] = i;
  }

  e->next = g_rules;
  g_rules = e;
  *out = e;
  return ERROR_SUCCESS;
}

/* the sharded engine of this YR_RULES over all configured devices, created on first use */
static int get_multi(rules_entry* re, yrb_multi** out)
{
  int rc = ERROR_SUCCESS;
  pthread_mutex_lock(&g_lock);
  if (re->multi == NULL)
  {
    const int* devs = NULL;
    const int n = multi_devices(&devs);
    void* image = NULL;
    size_t image_size = 0;
    rc = yr_b200_rules_flatten(re->rules, &image, &image_size);
    if (rc == ERROR_SUCCESS)
    {
      int r2 = yrb_multi_create(image, image_size, devs, n, &re->multi);
      yr_b200_image_free(image);
      if (r2 != YRB_SUCCESS)
        rc = fail_loudly("rules upload to $YR_B200_DEVICES (yrb_multi_create)", r2);
    }
  }
  *out = re->multi;
  pthread_mutex_unlock(&g_lock);
  return rc;
}

/* ---- pool of idle device scanners; call with g_lock held */
static yrb_scanner* pool_take(yrb_rules* dev_rules)
{
  for (int i = 0; i < g_pool_n; i++)
  {
    if (g_pool[i].dev_rules == dev_rules)
    {
      yrb_scanner* dev = g_pool[i].dev;
      g_pool[i] = g_pool[--g_pool_n];
      return dev;
    }
  }
  return NULL;
}

static void pool_put(yrb_scanner* dev, yrb_rules* dev_rules)
{
  if (g_pool_n < YR_B200_POOL_MAX)
  {
    size_t held = yrb_scanner_footprint(dev);
    for (int i = 0; i < g_pool_n; i++) held += yrb_scanner_footprint(g_pool[i].dev);
    if (held > YR_B200_POOL_BYTES)
      yrb_scanner_trim(dev);
    g_pool[g_pool_n].dev = dev;
    g_pool[g_pool_n].dev_rules = dev_rules;
    g_pool_n++;
  }
  else
  {
    yrb_scanner_destroy(dev);
  }
}

/* ---- file batches (yr_b200_batch_scan): results parked until the per-file scans ask for them */
struct YR_B200_BATCH
{
  YR_RULES* rules;
  uint32_t n;
  const uint8_t** data;
  size_t* sizes;
  uint32_t* order;   /* item indices sorted by data pointer */
  uint8_t* state;
  uint64_t* first;
  yrb_match* matches;
  struct YR_B200_BATCH* next;
};

static YR_B200_BATCH* volatile g_batches = NULL;

/* the scanner this thread is running a ONE-block scan for (yr_scanner_scan_mem), if any */
static __thread YR_SCANNER* t_single_block_scanner = NULL;

void yr_b200_single_block_scan(YR_SCANNER* scanner, int on) { t_single_block_scanner = on ? scanner : NULL; }

static uint64_t g_batch_hits = 0; /* scans served from parked results (tests, statistics) */
uint64_t yr_b200_batch_hits(void) { return __atomic_load_n(&g_batch_hits, __ATOMIC_RELAXED); }

static const uint8_t** g_sort_keys;
static int cmp_by_ptr(const void* a, const void* b)
{
  const uint8_t* pa = g_sort_keys[*(const uint32_t*) a];
  const uint8_t* pb = g_sort_keys[*(const uint32_t*) b];
  return pa < pb ? -1 : (pa > pb ? 1 : 0);
}

static void batch_free(YR_B200_BATCH* b)
{
  if (b == NULL) return;
  free((void*) b->data); free(b->sizes); free(b->order); free(b->state); free(b->first); free(b->matches);
  free(b);
}

int yr_b200_batch_scan(
    YR_RULES* rules, const uint8_t* const* buffers, const size_t* sizes, uint32_t n, int flags, int one_allocation,
    YR_B200_BATCH** batch)
{
  *batch = NULL;
  YR_B200_BATCH* b = (YR_B200_BATCH*) calloc(1, sizeof(*b));
  yrb_batch_item* items = (yrb_batch_item*) malloc(sizeof(yrb_batch_item) * (n ? n : 1));
  if (b != NULL)
  {
    b->rules = rules;
    b->n = n;
    b->data = (const uint8_t**) malloc(sizeof(uint8_t*) * (n ? n : 1));
    b->sizes = (size_t*) malloc(sizeof(size_t) * (n ? n : 1));
    b->order = (uint32_t*) malloc(sizeof(uint32_t) * (n ? n : 1));
    b->state = (uint8_t*) malloc(n ? n : 1);
    b->first = (uint64_t*) calloc((size_t) n + 1, sizeof(uint64_t));
  }
  if (b == NULL || items == NULL || !b->data || !b->sizes || !b->order || !b->state || !b->first)
  {
    free(items);
    batch_free(b);
    return ERROR_INSUFFICIENT_MEMORY;
  }
  for (uint32_t k = 0; k < n; k++)
  {
    b->data[k] = items[k].data = buffers[k];
    b->sizes[k] = sizes[k];
    items[k].size = sizes[k];
    b->order[k] = k;
  }

  /* a device scanner for the duration of the call */
  rules_entry* re = NULL;
  yrb_scanner* dev = NULL;
  pthread_mutex_lock(&g_lock);
  int rc = get_rules(rules, &re);
  if (rc == ERROR_SUCCESS)
    dev = pool_take(re->dev);
  pthread_mutex_unlock(&g_lock);
  if (rc == ERROR_SUCCESS && dev == NULL)
  {
    int r2 = yrb_scanner_create(re->dev, &dev);
    if (r2 != YRB_SUCCESS)
      rc = fail_loudly("yrb_scanner_create", r2);
  }
  if (rc != ERROR_SUCCESS)
  {
    free(items);
    batch_free(b);
    return rc;
  }

  struct timespec t0, t1;
  const int timing = getenv("YRB_TIMING") != NULL;
  if (timing) clock_gettime(CLOCK_MONOTONIC, &t0);
  yrb_batch_result res;
  int r2 = yrb_scan_batch(dev, items, n, flags, one_allocation ? YRB_BATCH_ONE_ALLOCATION : 0, &res);
  free(items);
  if (timing)
  {
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "yara_b200: batch of %u buffers: engine %.3f ms (device %.3f ms: scan+h2d %.3f verify %.3f resolve %.3f), %u launches\n",
            n, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6, res.stats.ms_total, res.stats.ms_scan,
            res.stats.ms_verify, res.stats.ms_resolve, res.stats.kernel_launches);
  }
  if (r2 == YRB_SUCCESS)
  {
    memcpy(b->state, res.item_state, n);
    memcpy(b->first, res.item_first, sizeof(uint64_t) * ((size_t) n + 1));
    const size_t total = (size_t) res.item_first[n];
    b->matches = (yrb_match*) malloc(sizeof(yrb_match) * (total ? total : 1));
    if (b->matches == NULL)
      rc = ERROR_INSUFFICIENT_MEMORY;
    else if (total)
      memcpy(b->matches, res.matches, sizeof(yrb_match) * total);
  }
  else
  {
    rc = r2 < 1000 ? r2 : fail_loudly("yrb_scan_batch", r2);
  }

  pthread_mutex_lock(&g_lock);
  pool_put(dev, re->dev);
  if (rc == ERROR_SUCCESS)
  {
    g_sort_keys = b->data;   /* under g_lock: qsort has no context argument */
    qsort(b->order, n, sizeof(uint32_t), cmp_by_ptr);
    b->next = g_batches;
    g_batches = b;
  }
  pthread_mutex_unlock(&g_lock);

  if (rc != ERROR_SUCCESS)
  {
    batch_free(b);
    return rc;
  }
  *batch = b;
  return ERROR_SUCCESS;
}

void yr_b200_batch_destroy(YR_B200_BATCH* batch)
{
  if (batch == NULL) return;
  pthread_mutex_lock(&g_lock);
  for (YR_B200_BATCH* volatile* pp = &g_batches; *pp != NULL; pp = &(*pp)->next)
  {
    if (*pp == batch)
    {
      *pp = batch->next;
      break;
    }
  }
  pthread_mutex_unlock(&g_lock);
  batch_free(batch);
}

uint32_t yr_b200_batch_rescans(const YR_B200_BATCH* batch)
{
  uint32_t c = 0;
  for (uint32_t k = 0; k < batch->n; k++) c += batch->state[k] != YRB_ITEM_DONE;
  return c;
}

/* The parked records of (rules, data, size), if a live batch scanned exactly that buffer. */
static int batch_lookup(YR_RULES* rules, const uint8_t* data, size_t size, const yrb_match** recs, size_t* n_recs)
{
  int found = 0;
  pthread_mutex_lock(&g_lock);
  for (YR_B200_BATCH* b = g_batches; b != NULL && !found; b = b->next)
  {
    if (b->rules != rules) continue;
    uint32_t lo = 0, hi = b->n;
    while (lo < hi)
    {
      uint32_t mid = (lo + hi) / 2;
      if (b->data[b->order[mid]] < data) lo = mid + 1; else hi = mid;
    }
    for (; lo < b->n && b->data[b->order[lo]] == data; lo++)
    {
      const uint32_t k = b->order[lo];
      if (b->sizes[k] == size && b->state[k] == YRB_ITEM_DONE)
      {
        *recs = b->matches + b->first[k];
        *n_recs = (size_t) (b->first[k + 1] - b->first[k]);
        found = 1;
        break;
      }
    }
  }
  pthread_mutex_unlock(&g_lock);
  return found;
}

/* ---- request coalescing ($YR_B200_COALESCE=N, off by default).  `yara -p N` has N threads that each
   scan one file at a time (cli/yara.c:1246-1290); every such scan is ~20 launches and a few
   synchronisations for microseconds of GPU work.  With coalescing, one-block scans of the same YR_RULES
   that are in flight at the same time are handed to ONE yrb_scan_batch call: the first thread to arrive
   leads -- it takes everything pending (its own request included, at most N), runs the batch and hands
   every waiter its records -- and whoever arrives while it is busy forms the next batch.  Results are
   those of yrb_scan_batch, i.e. of a scan of the buffer alone; items it flags YRB_ITEM_RESCAN fall
   through to the ordinary path.  No change of the libyara API or of the CLI. */
typedef struct co_req
{
  YR_RULES* rules;
  const uint8_t* data;
  size_t size;
  int flags;
  int state; /* 0 pending, 1 records ready, 2 scan it yourself, 3 failed (rc) */
  int rc;
  yrb_match* recs;
  size_t n_recs;
  struct co_req* next;
} co_req;

static pthread_mutex_t g_co_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_co_cond = PTHREAD_COND_INITIALIZER;
static co_req* g_co_head = NULL;
static co_req* g_co_tail = NULL;
static int g_co_leading = 0;
static int g_co_max = -1;
static uint64_t g_co_batches = 0, g_co_scans = 0;

static int co_max(void)
{
  if (g_co_max < 0)
  {
    const char* e = getenv("YR_B200_COALESCE");
    int v = e ? atoi(e) : 0;
    g_co_max = v < 2 ? 0 : (v > 1024 ? 1024 : v);
  }
  return g_co_max;
}

void yr_b200_coalesce_stats(uint64_t* batches, uint64_t* scans)
{
  *batches = __atomic_load_n(&g_co_batches, __ATOMIC_RELAXED);
  *scans = __atomic_load_n(&g_co_scans, __ATOMIC_RELAXED);
}

/* What the leader found out about one request.  Results are built here, in leader-local storage, and
   only PUBLISHED into the waiters' co_req (which live on the waiters' stacks) by co_submit under
   g_co_lock, state last: a waiter that sees state != 0 may return and pop its co_req at once, so
   nothing may be written to a request after its state, and nothing outside the lock. */
typedef struct co_out
{
  int state;
  int rc;
  yrb_match* recs;
  size_t n_recs;
} co_out;

/* the leader's part: one engine batch over reqs[0..n), all of one YR_RULES and one flag set.
   Reads the requests (immutable while pending), writes only outs[]. */
static void co_run_batch(co_req* const* reqs, co_out* outs, uint32_t n)
{
  int rc = ERROR_SUCCESS;
  rules_entry* re = NULL;
  yrb_scanner* dev = NULL;
  yrb_batch_item* items = (yrb_batch_item*) malloc(sizeof(yrb_batch_item) * n);

  memset(outs, 0, sizeof(co_out) * n);

  if (items == NULL)
    rc = ERROR_INSUFFICIENT_MEMORY;

  if (rc == ERROR_SUCCESS)
  {
    pthread_mutex_lock(&g_lock);
    rc = get_rules(reqs[0]->rules, &re);
    if (rc == ERROR_SUCCESS)
      dev = pool_take(re->dev);
    pthread_mutex_unlock(&g_lock);
    if (rc == ERROR_SUCCESS && dev == NULL)
    {
      int r2 = yrb_scanner_create(re->dev, &dev);
      if (r2 != YRB_SUCCESS)
        rc = fail_loudly("yrb_scanner_create", r2);
    }
  }

  if (rc == ERROR_SUCCESS)
  {
    for (uint32_t k = 0; k < n; k++)
    {
      items[k].data = reqs[k]->data;
      items[k].size = reqs[k]->size;
    }
    yrb_batch_result res;
    int r2 = yrb_scan_batch(dev, items, n, reqs[0]->flags, 0, &res);
    if (r2 != YRB_SUCCESS)
    {
      rc = r2 < 1000 ? r2 : fail_loudly("yrb_scan_batch", r2);
    }
    else
    {
      for (uint32_t k = 0; k < n; k++)
      {
        co_out* o = &outs[k];
        if (res.item_state[k] != YRB_ITEM_DONE)
        {
          o->state = 2;
          continue;
        }
        o->n_recs = (size_t) (res.item_first[k + 1] - res.item_first[k]);
        o->recs = (yrb_match*) malloc(sizeof(yrb_match) * (o->n_recs ? o->n_recs : 1));
        if (o->recs == NULL)
        {
          o->rc = ERROR_INSUFFICIENT_MEMORY;
          o->state = 3;
          continue;
        }
        memcpy(o->recs, res.matches + res.item_first[k], sizeof(yrb_match) * o->n_recs);
        o->state = 1;
      }
    }
    pthread_mutex_lock(&g_lock);
    pool_put(dev, re->dev);
    pthread_mutex_unlock(&g_lock);
  }

  free(items);
  if (rc != ERROR_SUCCESS)
  {
    for (uint32_t k = 0; k < n; k++)
    {
      free(outs[k].recs);
      outs[k].recs = NULL;
      outs[k].n_recs = 0;
      outs[k].rc = rc;
      outs[k].state = 3;
    }
  }
  __atomic_fetch_add(&g_co_batches, 1, __ATOMIC_RELAXED);
  __atomic_fetch_add(&g_co_scans, n, __ATOMIC_RELAXED);
}

/* Blocks until r is resolved one way or the other; returns r->state. */
static int co_submit(co_req* r)
{
  co_req* mine[1024];
  co_out* outs = NULL;

  pthread_mutex_lock(&g_co_lock);
  r->next = NULL;
  if (g_co_tail) g_co_tail->next = r; else g_co_head = r;
  g_co_tail = r;

  while (r->state == 0)
  {
    if (g_co_leading)
    {
      pthread_cond_wait(&g_co_cond, &g_co_lock);
      continue;
    }

    /* lead: my own request plus whatever else is pending for the same rules and flags, up to the cap
       (one slot stays reserved for my own request until the walk has met it) */
    g_co_leading = 1;
    uint32_t n = 0;
    const uint32_t cap = (uint32_t) co_max();
    int seen_me = 0;
    co_req** pp = &g_co_head;
    co_req* last = NULL;
    while (*pp != NULL)
    {
      co_req* q = *pp;
      int take;
      if (q == r)
      {
        take = 1;
        seen_me = 1;
      }
      else
      {
        take = q->rules == r->rules && q->flags == r->flags && n < cap - (seen_me ? 0u : 1u);
      }
      if (take)
      {
        *pp = q->next;
        q->next = NULL;
        mine[n++] = q;
      }
      else
      {
        last = q;
        pp = &q->next;
      }
    }
    g_co_tail = last;
    pthread_mutex_unlock(&g_co_lock);

    outs = (co_out*) malloc(sizeof(co_out) * n);
    if (outs != NULL)
      co_run_batch(mine, outs, n);

    /* publish under the lock, state last; after this line a waiter's co_req may be gone */
    pthread_mutex_lock(&g_co_lock);
    for (uint32_t k = 0; k < n; k++)
    {
      co_req* q = mine[k];
      if (outs != NULL)
      {
        q->rc = outs[k].rc;
        q->recs = outs[k].recs;
        q->n_recs = outs[k].n_recs;
        q->state = outs[k].state;
      }
      else
      {
        q->rc = ERROR_INSUFFICIENT_MEMORY;
        q->state = 3;
      }
    }
    free(outs);
    outs = NULL;
    g_co_leading = 0;
    pthread_cond_broadcast(&g_co_cond);
  }
  const int st = r->state;
  pthread_mutex_unlock(&g_co_lock);
  return st;
}

static __thread scanner_entry* t_last_entry = NULL; /* the entry get_scanner found or made for this thread's last call */

static int get_scanner(YR_SCANNER* scanner, yrb_scanner** out, rules_entry** re_out)
{
  scanner_entry* e;
  int rc = ERROR_SUCCESS;

  pthread_mutex_lock(&g_lock);

  for (e = g_scanners; e != NULL; e = e->next)
  {
    if (e->scanner == scanner)
    {
      *out = e->dev;
      *re_out = e->re;
      t_last_entry = e;
      pthread_mutex_unlock(&g_lock);
      return ERROR_SUCCESS;
    }
  }

  rules_entry* re = NULL;
  yrb_rules* dev_rules = NULL;
  rc = get_rules(scanner->rules, &re);
  if (rc == ERROR_SUCCESS)
    dev_rules = re->dev;

  if (rc == ERROR_SUCCESS)
  {
    yrb_scanner* dev = pool_take(dev_rules);
    int r2 = YRB_SUCCESS;

    if (dev == NULL)
      r2 = yrb_scanner_create(dev_rules, &dev);

    if (r2 != YRB_SUCCESS)
    {
      rc = fail_loudly("yrb_scanner_create", r2);
    }
    else
    {
      e = (scanner_entry*) calloc(1, sizeof(*e));
      if (e == NULL)
      {
        pool_put(dev, dev_rules);
        pthread_mutex_unlock(&g_lock);
        return ERROR_INSUFFICIENT_MEMORY;
      }
      e->scanner = scanner;
      e->dev = dev;
      e->dev_rules = dev_rules;
      e->re = re;
      e->next = g_scanners;
      g_scanners = e;
      t_last_entry = e;
      *out = dev;
      *re_out = re;
    }
  }

  pthread_mutex_unlock(&g_lock);
  return rc;
}

void yr_b200_scanner_release(YR_SCANNER* scanner)
{
  yr_b200_prefilter_scanner_release(scanner);
  pthread_mutex_lock(&g_lock);
  scanner_entry** pp = &g_scanners;
  while (*pp != NULL)
  {
    if ((*pp)->scanner == scanner)
    {
      scanner_entry* e = *pp;
      *pp = e->next;
      pool_put(e->dev, e->dev_rules);
      free(e->pend_data);
      free(e->pend);
      free(e);
      break;
    }
    pp = &(*pp)->next;
  }
  pthread_mutex_unlock(&g_lock);
}

void yr_b200_rules_release(YR_RULES* rules)
{
  yr_b200_prefilter_rules_release(rules);
  pthread_mutex_lock(&g_lock);
  rules_entry** pp = &g_rules;
  while (*pp != NULL)
  {
    if ((*pp)->rules == rules)
    {
      rules_entry* e = *pp;
      *pp = e->next;
      for (int i = 0; i < g_pool_n;)
      {
        if (g_pool[i].dev_rules == e->dev)
        {
          yrb_scanner_destroy(g_pool[i].dev);
          g_pool[i] = g_pool[--g_pool_n];
        }
        else
        {
          i++;
        }
      }
      if (e->multi != NULL)
        yrb_multi_destroy(e->multi);
      pthread_mutex_destroy(&e->multi_lock);
      yrb_rules_release(e->dev);
      free(e->chain_strings);
      free(e);
      continue;
    }
    pp = &(*pp)->next;
  }
  pthread_mutex_unlock(&g_lock);
}

/* One node into a string's address-ordered match list; the outcome the reference's list insert gives
   (scan.c:290-368): a full list refuses the node (ERROR_TOO_MANY_MATCHES, nothing touched); a node at an
   address the list already holds is dropped -- or, for greedy regexps (overwrite_equal), lends its lengths
   and data to the one that is there; anything else is linked behind the last node with a smaller address.
   The engine hands records over in ascending order, so the walk from the tail ends at once for the
   records of one block; it only goes further when blocks come out of address order. */
static int list_insert(YR_MATCH* m, YR_MATCHES* list, int overwrite_equal)
{
  if (list->count == YR_MAX_STRING_MATCHES)
    return ERROR_TOO_MANY_MATCHES;

  const int64_t at = m->base + m->offset;
  YR_MATCH* after = list->tail; /* m goes behind this one; NULL: m becomes the head */

  for (; after != NULL; after = after->prev)
  {
    const int64_t there = after->base + after->offset;

    if (there < at)
      break;

    if (there == at)
    {
      if (overwrite_equal)
      {
        after->match_length = m->match_length;
        after->data_length = m->data_length;
        after->data = m->data;
      }
      return ERROR_SUCCESS;
    }
  }

  YR_MATCH* before = after != NULL ? after->next : list->head;

  m->prev = after;
  m->next = before;
  *(after != NULL ? &after->next : &list->head) = m;
  *(before != NULL ? &before->prev : &list->tail) = m;
  list->count++;

  return ERROR_SUCCESS;
}

/* A string whose list is full: the callback decides (scan.c:1054-1073). CALLBACK_CONTINUE silences the
   string for the rest of the scan; anything else ends the scan with the string recorded. */
static int list_full(YR_SCANNER* scanner, YR_STRING* string)
{
  if (scanner->callback(scanner, CALLBACK_MSG_TOO_MANY_MATCHES, (void*) string, scanner->user_data) != CALLBACK_CONTINUE)
  {
    scanner->last_error_string = string;
    return ERROR_TOO_MANY_MATCHES;
  }

  yr_bitmask_set(scanner->strings_temp_disabled, string->idx);
  return ERROR_SUCCESS;
}

/* match data of records[k] packed by the engine (yrb_last_match_data), for the next materialise() of this thread:
   data[offs[k] .. offs[k + 1]).  NULL: copy from the block. */
static __thread const uint8_t* t_packed_data = NULL;
static __thread const uint64_t* t_packed_offs = NULL;

static int materialise(
    YR_SCANNER* scanner, const uint8_t* block_data, YR_MEMORY_BLOCK* block, const yrb_match* recs, size_t n,
    YR_MATCHES* lists, int confirmed, uint32_t max_match_data, YR_MATCH** seed_nodes)
{
  const uint8_t* packed = t_packed_data;
  const uint64_t* packed_offs = t_packed_offs;
  t_packed_data = NULL;
  t_packed_offs = NULL;

  YR_RULES* rules = scanner->rules;

  for (size_t i = 0; i < n; i++)
  {
    const yrb_match* r = &recs[i];
    YR_STRING* string = &rules->strings_table[r->string_idx];

    yr_b200_prefilter_note(scanner, string); /* condition-VM pre-filter: this rule has something to look at */

    /* the data copies below touch the block at random offsets (a cache and TLB miss each
       in a multi-GiB block): get the next ones moving */
    if (packed == NULL && i + 12 < n && recs[i + 12].offset >= 0 && (size_t) recs[i + 12].offset < block->size)
      __builtin_prefetch(block_data + recs[i + 12].offset);

    if (yr_bitmask_is_set(scanner->strings_temp_disabled, string->idx))
      continue; /* disabled in an earlier block of this scan (scan.c:1018) */

    if ((r->flags & YRB_MATCH_FROM_SEED) && seed_nodes != NULL)
    {
      /* a YR_MATCH node an earlier block created: it keeps its base and offset */
      YR_MATCH* node = seed_nodes[YRB_MATCH_SEED_INDEX(r)];
      node->prev = node->next = NULL;
      node->chain_length = r->chain_length;

      if (confirmed)
      {
        /* scan.c:561-576: total length up to the tail, data re-copied from the CURRENT block at
           the node's own (block-relative) offset; the reference reads whatever lies there */
        node->match_length = r->match_length;
        node->data_length = r->match_length < (int32_t) max_match_data ? r->match_length : (int32_t) max_match_data;
        if (node->data_length < 0) node->data_length = 0;
        uint8_t* d = (uint8_t*) yr_notebook_alloc(scanner->matches_notebook, node->data_length > 0 ? node->data_length : 1);
        if (d == NULL)
          return ERROR_INSUFFICIENT_MEMORY;
        memset(d, 0, node->data_length);
        if (node->offset >= 0 && (size_t) node->offset < block->size)
        {
          size_t avail = block->size - (size_t) node->offset;
          memcpy(d, block_data + node->offset, avail < (size_t) node->data_length ? avail : (size_t) node->data_length);
        }
        node->data = d;
      }

      int rcs = list_insert(node, &lists[string->idx], 0);
      if (rcs == ERROR_TOO_MANY_MATCHES)
      {
        rcs = list_full(scanner, string);
        if (rcs != ERROR_SUCCESS)
          return rcs;
      }
      continue;
    }

    YR_MATCH* m = (YR_MATCH*) yr_notebook_alloc(scanner->matches_notebook, sizeof(YR_MATCH));
    if (m == NULL)
      return ERROR_INSUFFICIENT_MEMORY;

    m->base = block->base;
    m->offset = r->offset;
    m->match_length = r->match_length;
    m->data_length = r->match_length < (int32_t) max_match_data ? r->match_length : (int32_t) max_match_data;
    m->prev = NULL;
    m->next = NULL;
    m->chain_length = r->chain_length;
    m->is_private = STRING_IS_PRIVATE(string);
    m->xor_key = r->xor_key;
    m->data = NULL;

    if (m->data_length > 0)
    {
      uint8_t* d = (uint8_t*) yr_notebook_alloc(scanner->matches_notebook, m->data_length);
      if (d == NULL)
        return ERROR_INSUFFICIENT_MEMORY;
      if (packed != NULL && packed_offs[i + 1] - packed_offs[i] == (uint64_t) m->data_length)
        memcpy(d, packed + packed_offs[i], m->data_length);
      else
        memcpy(d, block_data + r->offset, m->data_length);
      m->data = d;
    }
    else if (m->data_length < 0)
    {
      m->data_length = 0;
    }

    int rc = list_insert(m, &lists[string->idx], confirmed && STRING_IS_GREEDY_REGEXP(string));

    if (rc == ERROR_TOO_MANY_MATCHES)
    {
      /* only reachable when earlier blocks already filled the list */
      rc = list_full(scanner, string);
      if (rc != ERROR_SUCCESS)
        return rc;
    }
  }

  return ERROR_SUCCESS;
}

/* What earlier blocks of this scan left in unconfirmed_matches[] of the chained
   strings (and which chain heads already have confirmed matches): handed to the
   engine so that chains whose parts lie in different blocks resolve exactly as in
   the reference, which keeps these lists across blocks (types.h:806-816). */
static int collect_seeds(YR_SCANNER* scanner, rules_entry* re, yrb_seed** seeds_out, YR_MATCH*** nodes_out, uint32_t* n_out)
{
  uint32_t n = 0;
  *seeds_out = NULL;
  *nodes_out = NULL;
  *n_out = 0;

  for (uint32_t k = 0; k < re->n_chain_strings; k++)
  {
    uint32_t idx = re->chain_strings[k];
    n += (uint32_t) scanner->unconfirmed_matches[idx].count;
    if (scanner->matches[idx].head != NULL)
      n++;
  }

  if (n == 0)
    return ERROR_SUCCESS;
  if (n > YRB_MAX_SEEDS)
    return fail_loudly("more than 4M chain-part matches carried between blocks", ERROR_INTERNAL_FATAL_ERROR);

  yrb_seed* seeds = (yrb_seed*) calloc(n, sizeof(yrb_seed));
  YR_MATCH** nodes = (YR_MATCH**) calloc(n, sizeof(YR_MATCH*));
  if (seeds == NULL || nodes == NULL)
  {
    free(seeds);
    free(nodes);
    return ERROR_INSUFFICIENT_MEMORY;
  }

  uint32_t i = 0;
  for (uint32_t k = 0; k < re->n_chain_strings; k++)
  {
    uint32_t idx = re->chain_strings[k];
    for (YR_MATCH* m = scanner->unconfirmed_matches[idx].head; m != NULL && i < n; m = m->next)
    {
      seeds[i].base = m->base;
      seeds[i].offset = m->offset;
      seeds[i].string_idx = idx;
      seeds[i].match_length = m->match_length;
      seeds[i].chain_length = m->chain_length;
      seeds[i].xor_key = m->xor_key;
      seeds[i].kind = 0;
      nodes[i] = m;
      i++;
    }
    if (scanner->matches[idx].head != NULL && i < n)
    {
      seeds[i].string_idx = idx;
      seeds[i].kind = 1;
      i++;
    }
  }

  *seeds_out = seeds;
  *nodes_out = nodes;
  *n_out = i;
  return ERROR_SUCCESS;
}

static int scan_block_now(
    YR_SCANNER* scanner, const uint8_t* block_data, YR_MEMORY_BLOCK* block, yrb_scanner* dev, rules_entry* re,
    uint32_t max_match_data);

static int match_data_gather_on(void)
{
  static int v = -1;
  if (v < 0)
  {
    const char* e = getenv("YR_B200_GATHER_MATCH_DATA");
    v = e ? atoi(e) != 0 : 1;
  }
  return v;
}

/* blocks of at most this many bytes are deferred and batched ($YR_B200_BLOCK_BATCH, bytes; 0 = off) */
static size_t block_batch_max(void)
{
  static long v = -1;
  if (v < 0)
  {
    const char* e = getenv("YR_B200_BLOCK_BATCH");
    v = e ? atol(e) : (256 << 10);
    if (v < 0) v = 0;
  }
  return (size_t) v;
}

static int pend_append(scanner_entry* e, YR_SCANNER* scanner, const uint8_t* data, YR_MEMORY_BLOCK* block)
{
  const size_t off = (e->pend_used + 15) & ~(size_t) 15;
  if (off + block->size + 16 > e->pend_cap)
  {
    size_t ncap = e->pend_cap ? e->pend_cap : ((size_t) 64 << 10);
    while (ncap < off + block->size + 16) ncap *= 2;
    uint8_t* nd = (uint8_t*) realloc(e->pend_data, ncap);
    if (nd == NULL) return ERROR_INSUFFICIENT_MEMORY;
    e->pend_data = nd;
    e->pend_cap = ncap;
  }
  if (e->n_pend == e->cap_pend)
  {
    const uint32_t nc = e->cap_pend ? e->cap_pend * 2 : 64;
    pend_block* np = (pend_block*) realloc(e->pend, sizeof(pend_block) * nc);
    if (np == NULL) return ERROR_INSUFFICIENT_MEMORY;
    e->pend = np;
    e->cap_pend = nc;
  }
  if (off > e->pend_used) memset(e->pend_data + e->pend_used, 0, off - e->pend_used);
  if (block->size) memcpy(e->pend_data + off, data, block->size);
  e->pend[e->n_pend].base = block->base;
  e->pend[e->n_pend].size = block->size;
  e->pend[e->n_pend].off = off;
  e->n_pend++;
  e->pend_used = off + block->size;
  e->pend_scan = (const void*) scanner->matches_notebook;
  return ERROR_SUCCESS;
}

/* The deferred blocks, in order: ONE yrb_scan_batch over their copies.  A block the batch answers for (no chained
   string matched in it, no verifier error, nothing near the match cap: include/yara_b200.h) gets its YR_MATCH nodes
   from the batch records -- dedup against earlier blocks by base + offset, the cap across blocks and temporarily
   disabled strings are handled where the nodes are linked (materialise), as for any block.  A block the batch
   flags is scanned on its own, now, with the chain state of everything before it as seeds: exactly what would have
   happened without the batch.  Blocks without chain-part matches neither read nor change that state, so the order
   of the results is the order of the blocks. */
static int blocks_flush(YR_SCANNER* scanner, scanner_entry* e, uint32_t max_match_data)
{
  const uint32_t n = e->n_pend;
  int rc = ERROR_SUCCESS;
  if (n == 0)
    return ERROR_SUCCESS;
  if (e->pend_scan != (const void*) scanner->matches_notebook)
  {
    e->n_pend = 0;
    e->pend_used = 0;
    return ERROR_SUCCESS;
  }

  yrb_batch_item* items = (yrb_batch_item*) malloc(sizeof(yrb_batch_item) * n);
  uint8_t* state = (uint8_t*) malloc(n);
  uint64_t* first = (uint64_t*) malloc(sizeof(uint64_t) * ((size_t) n + 1));
  yrb_match* recs = NULL;
  if (items == NULL || state == NULL || first == NULL)
    rc = ERROR_INSUFFICIENT_MEMORY;

  if (rc == ERROR_SUCCESS)
  {
    for (uint32_t k = 0; k < n; k++)
    {
      items[k].data = e->pend_data + e->pend[k].off;
      items[k].size = e->pend[k].size;
    }
    yrb_batch_result br;
    /* a handful of blocks is cheaper one by one: a device scanner's first batch allocates its pinned segment tables
       (milliseconds), and the reference's test suite makes a new ruleset -- hence a new scanner -- per assertion */
    int r2 = n >= 8 ? yrb_scan_batch(e->dev, items, n, scanner->flags, YRB_BATCH_ONE_ALLOCATION | YRB_BATCH_BLOCKS_OF_ONE_SCAN, &br)
                    : -1;
    if (r2 == YRB_SUCCESS)
    {
      /* the records live in the device scanner's buffers, which the per-block scans below reuse: take a copy */
      const uint64_t total = br.item_first[n];
      recs = (yrb_match*) malloc(sizeof(yrb_match) * (total ? total : 1));
      if (recs == NULL)
        rc = ERROR_INSUFFICIENT_MEMORY;
      else
      {
        memcpy(state, br.item_state, n);
        memcpy(first, br.item_first, sizeof(uint64_t) * ((size_t) n + 1));
        memcpy(recs, br.matches, sizeof(yrb_match) * total);
      }
    }
    else if (r2 == -1 || r2 < 1000)
      memset(state, YRB_ITEM_RESCAN, n); /* a single block, or the batch refused (its own error codes): one by one */
    else
      rc = fail_loudly("yrb_scan_batch", r2);
  }

  for (uint32_t k = 0; k < n && rc == ERROR_SUCCESS; k++)
  {
    YR_MEMORY_BLOCK blk;
    memset(&blk, 0, sizeof(blk));
    blk.base = e->pend[k].base;
    blk.size = e->pend[k].size;
    const uint8_t* data = e->pend_data + e->pend[k].off;
    if (scanner->timeout > 0 && yr_stopwatch_elapsed_ns(&scanner->stopwatch) > scanner->timeout)
      rc = ERROR_SCAN_TIMEOUT;
    else if (state[k] == YRB_ITEM_DONE && recs != NULL)
      rc = materialise(scanner, data, &blk, recs + first[k], (size_t) (first[k + 1] - first[k]), scanner->matches, 1,
                       max_match_data, NULL);
    else
      rc = scan_block_now(scanner, data, &blk, e->dev, e->re, max_match_data);
  }

  free(items);
  free(state);
  free(first);
  free(recs);
  e->n_pend = 0;
  e->pend_used = 0;
  return rc;
}

/* scanner.c, behind the block loop of yr_scanner_scan_mem_blocks: whatever is still deferred */
int yr_b200_blocks_flush(YR_SCANNER* scanner)
{
  uint32_t max_match_data;
  scanner_entry* e;
  pthread_mutex_lock(&g_lock);
  for (e = g_scanners; e != NULL; e = e->next)
    if (e->scanner == scanner)
      break;
  pthread_mutex_unlock(&g_lock);
  if (e == NULL || e->n_pend == 0)
    return ERROR_SUCCESS;
  FAIL_ON_ERROR(yr_get_configuration_uint32(YR_CONFIG_MAX_MATCH_DATA, &max_match_data));
  return blocks_flush(scanner, e, max_match_data);
}

/* scanner.c:545 (_exit): the scan is over, whatever its result; blocks still deferred belong to it and go with it */
void yr_b200_blocks_discard(YR_SCANNER* scanner)
{
  scanner_entry* e;
  pthread_mutex_lock(&g_lock);
  for (e = g_scanners; e != NULL; e = e->next)
    if (e->scanner == scanner)
      break;
  pthread_mutex_unlock(&g_lock);
  if (e != NULL)
  {
    e->n_pend = 0;
    e->pend_used = 0;
    e->pend_scan = NULL;
  }
}

int yr_b200_scan_mem_block(YR_SCANNER* scanner, const uint8_t* block_data, YR_MEMORY_BLOCK* block)
{
  yrb_scanner* dev = NULL;
  rules_entry* re = NULL;
  uint32_t max_match_data;
  int rc;

  if (scanner->timeout > 0 && yr_stopwatch_elapsed_ns(&scanner->stopwatch) > scanner->timeout)
    return ERROR_SCAN_TIMEOUT;

  FAIL_ON_ERROR(yr_get_configuration_uint32(YR_CONFIG_MAX_MATCH_DATA, &max_match_data));

  /* already scanned as part of a live file batch?  (only items whose result is complete and
     independent of scan state are parked there; the rest fall through to the scan below).  A parked
     result is that of a whole one-block scan -- chain parts left unconfirmed at its end are dropped --
     so it only stands in for yr_scanner_scan_mem, never for the first block of a longer scan. */
  if (g_batches != NULL && t_single_block_scanner == scanner && block->base == 0 &&
      !(scanner->flags & SCAN_FLAGS_FAST_MODE))
  {
    const yrb_match* recs = NULL;
    size_t n_recs = 0;
    if (batch_lookup(scanner->rules, block_data, block->size, &recs, &n_recs))
    {
      __atomic_fetch_add(&g_batch_hits, 1, __ATOMIC_RELAXED);
      return materialise(scanner, block_data, block, recs, n_recs, scanner->matches, 1, max_match_data, NULL);
    }
  }

  /* one-block scans in flight on several threads share one engine batch (off unless $YR_B200_COALESCE) */
  if (co_max() > 1 && t_single_block_scanner == scanner && block->base == 0 && block->size > 0 &&
      !(scanner->flags & SCAN_FLAGS_FAST_MODE))
  {
    co_req r;
    memset(&r, 0, sizeof(r));
    r.rules = scanner->rules;
    r.data = block_data;
    r.size = block->size;
    r.flags = scanner->flags;
    const int st = co_submit(&r);
    if (st == 1)
    {
      rc = materialise(scanner, block_data, block, r.recs, r.n_recs, scanner->matches, 1, max_match_data, NULL);
      free(r.recs);
      return rc;
    }
    if (st == 3)
      return r.rc;
    /* st == 2: the batch could not answer for this buffer: scan it on its own */
  }

  rc = get_scanner(scanner, &dev, &re);
  if (rc != ERROR_SUCCESS)
    return rc;
  scanner_entry* entry = t_last_entry;

  /* ---- the small blocks of a multi-block scan (process memory, custom iterators: scanner.c:467-500,
     proc/linux.c:309) are not scanned one launch sequence each but collected and scanned as ONE engine batch
     (blocks_flush) -- when the next large block arrives, when enough have piled up, or behind the loop. */
  if (block_batch_max() > 0 && entry != NULL && t_single_block_scanner != scanner && block->size <= block_batch_max() &&
      !(scanner->flags & SCAN_FLAGS_FAST_MODE) && !re->has_fixed_offset /* a batch item has no base */)
  {
    if (entry->n_pend > 0 && entry->pend_scan != (const void*) scanner->matches_notebook)
      entry->n_pend = 0, entry->pend_used = 0; /* left behind by a scan that ended in an exception */
    if (entry->n_pend >= 1024 || entry->pend_used + block->size + 16 > ((size_t) 8 << 20))
    {
      rc = blocks_flush(scanner, entry, max_match_data);
      if (rc != ERROR_SUCCESS)
        return rc;
    }
    if (pend_append(entry, scanner, block_data, block) == ERROR_SUCCESS)
      return ERROR_SUCCESS;
    /* out of memory for the copy: scan it now */
  }
  if (entry != NULL && entry->n_pend > 0)
  {
    rc = blocks_flush(scanner, entry, max_match_data); /* keep the order of the blocks */
    if (rc != ERROR_SUCCESS)
      return rc;
  }
  return scan_block_now(scanner, block_data, block, dev, re, max_match_data);
}

/* scan.c:1018-1023, the part of it that depends on EARLIER blocks of the scan: a string the 1,000,000-match cap
   disabled there, or -- in fast mode -- a SINGLE_MATCH string that matched there.  (Heads of chained strings are
   the engine's business: their state travels as seeds.) */
static int not_verified_because_of_earlier_blocks(YR_SCANNER* scanner, YR_STRING* string)
{
  if (yr_bitmask_is_set(scanner->strings_temp_disabled, string->idx))
    return 1;
  return (scanner->flags & SCAN_FLAGS_FAST_MODE) && STRING_IS_SINGLE_MATCH(string) && !STRING_IS_CHAIN_PART(string) &&
         scanner->matches[string->idx].head != NULL;
}

/* One block through the engine, now: H2D + scan + verify + resolve with the chain state of the earlier blocks as
   seeds, YR_MATCH lists built from the records. */
static int scan_block_now(
    YR_SCANNER* scanner, const uint8_t* block_data, YR_MEMORY_BLOCK* block, yrb_scanner* dev, rules_entry* re,
    uint32_t max_match_data)
{
  int rc;
  yrb_scan_params params;
  memset(&params, 0, sizeof(params));
  params.base = block->base;
  params.scan_flags = scanner->flags;
  params.too_many_action = 0; /* the user's answer is only known after asking: see below */
  if (scanner->timeout > 0)
  {
    /* scanner.c:74-81: what is left of the scan's time budget; the engine checks it between H2D slices */
    uint64_t elapsed = yr_stopwatch_elapsed_ns(&scanner->stopwatch);
    params.timeout_ns = elapsed < scanner->timeout ? scanner->timeout - elapsed : 1;
  }

  yrb_seed* seeds = NULL;
  YR_MATCH** seed_nodes = NULL;
  uint32_t n_seeds = 0;
  if (re->n_chain_strings > 0 && re->chain_strings != NULL)
  {
    rc = collect_seeds(scanner, re, &seeds, &seed_nodes, &n_seeds);
    if (rc != ERROR_SUCCESS)
      return rc;
    params.seeds = seeds;
    params.n_seeds = n_seeds;
  }

  static int timing = -1;
  if (timing < 0) timing = getenv("YRB_TIMING") != NULL;
  struct timespec t0, t1, t2;
  if (timing) clock_gettime(CLOCK_MONOTONIC, &t0);

  yrb_scan_result res;
  const int* multi_devs = NULL;
  rules_entry* locked = NULL;   /* holds re->multi_lock until the records have been materialised */
  if (n_seeds == 0 && multi_devices(&multi_devs) > 1 && block->size >= multi_min_bytes())
  {
    yrb_multi* multi = NULL;
    rc = get_multi(re, &multi);
    if (rc != ERROR_SUCCESS)
    {
      free(seeds);
      free(seed_nodes);
      return rc;
    }
    pthread_mutex_lock(&re->multi_lock);
    locked = re;
    rc = yrb_multi_scan_host(multi, block_data, block->size, &params, &res);
  }
  else
  {
    rc = yrb_scan_host(dev, block_data, block->size, &params, &res);
  }
  if (timing) clock_gettime(CLOCK_MONOTONIC, &t1);
  free(seeds);
#define YR_B200_UNLOCK_MULTI() do { if (locked != NULL) { pthread_mutex_unlock(&locked->multi_lock); locked = NULL; } } while (0)
  if (rc != YRB_SUCCESS)
  {
    YR_B200_UNLOCK_MULTI();
    free(seed_nodes);
    return rc < 1000 ? rc : fail_loudly("yrb_scan_host", rc);
  }

  /* A verifier error of a string whose candidates the reference would not have verified in this block, because of
     what EARLIER blocks of the scan did to it (scan.c:1018-1023), is not the scan's error.  The engine knows a block
     at a time; it reports the earliest failing candidate, so only when that one belongs to such a string is there
     anything to do: the block is scanned once more with those strings named (yrb_scan_params.skip_strings), and
     whatever comes back then stands. */
  uint32_t* skip_strings = NULL;
  if (res.error != 0 && res.error != ERROR_TOO_MANY_MATCHES && res.error_string != YRB_NONE &&
      res.error_string < scanner->rules->num_strings && params.skip_strings == NULL &&
      not_verified_because_of_earlier_blocks(scanner, &scanner->rules->strings_table[res.error_string]))
  {
    YR_B200_UNLOCK_MULTI();
    uint32_t n_skip = 0;
    skip_strings = (uint32_t*) malloc(sizeof(uint32_t) * scanner->rules->num_strings);
    if (skip_strings == NULL)
    {
      free(seed_nodes);
      return ERROR_INSUFFICIENT_MEMORY;
    }
    for (uint32_t i = 0; i < scanner->rules->num_strings; i++)
      if (not_verified_because_of_earlier_blocks(scanner, &scanner->rules->strings_table[i]))
        skip_strings[n_skip++] = i;
    params.skip_strings = skip_strings;
    params.n_skip_strings = n_skip;
    if (re->n_chain_strings > 0 && re->chain_strings != NULL)
    {
      /* the seeds again: the first call's copy is gone (the YR_MATCH nodes they stand for are untouched) */
      free(seed_nodes);
      seed_nodes = NULL;
      rc = collect_seeds(scanner, re, &seeds, &seed_nodes, &n_seeds);
      if (rc != ERROR_SUCCESS)
      {
        free(skip_strings);
        return rc;
      }
      params.seeds = seeds;
      params.n_seeds = n_seeds;
    }
    rc = yrb_scan_host(dev, block_data, block->size, &params, &res);
    if (re->n_chain_strings > 0 && re->chain_strings != NULL)
      free(seeds);
    free(skip_strings);
    if (rc != YRB_SUCCESS)
    {
      free(seed_nodes);
      return rc < 1000 ? rc : fail_loudly("yrb_scan_host", rc);
    }
  }

  if (res.error != 0)
  {
    YR_B200_UNLOCK_MULTI();
    free(seed_nodes);
    if (res.error_string != YRB_NONE)
      scanner->last_error_string = &scanner->rules->strings_table[res.error_string];
    return res.error;
  }

  if (res.seeds_consumed)
  {
    /* `unconfirmed` is the complete new state of the chained strings' lists: rebuild them */
    for (uint32_t k = 0; k < re->n_chain_strings; k++)
    {
      YR_MATCHES* l = &scanner->unconfirmed_matches[re->chain_strings[k]];
      l->head = l->tail = NULL;
      l->count = 0;
    }
  }

  /* In fast mode a SINGLE_MATCH string that already matched in an earlier block is
     not verified again (scan.c:1021-1023): drop this block's records for it.  Not so the
     heads of chained strings: their confirmed records are chains that a TAIL completed -- tails
     never have matches[] entries and are always verified -- possibly with a head match an earlier
     block left unconfirmed; that the head's own candidates are no longer verified is known to the
     engine through the seeds (RAW_SEED_MARK). */
  if (scanner->flags & SCAN_FLAGS_FAST_MODE)
  {
    yrb_match* recs = (yrb_match*) res.matches;
    size_t k = 0;
    for (size_t i = 0; i < res.n_matches; i++)
    {
      YR_STRING* string = &scanner->rules->strings_table[recs[i].string_idx];
      if (STRING_IS_SINGLE_MATCH(string) && !STRING_IS_CHAIN_PART(string) && scanner->matches[string->idx].head != NULL)
        continue;
      recs[k++] = recs[i];
    }
    res.n_matches = k;
  }

  /* Large results of a single-device scan: the engine gathers the match data from its resident copy of the block and
     hands it over packed -- materialise then reads sequentially instead of taking a cache and TLB miss per match
     in a multi-GiB host block (28 of the 340 ms of a 16 GiB scan with 262k matches). */
  if (locked == NULL && n_seeds == 0 && res.n_matches >= 8192 && max_match_data > 0 && !(scanner->flags & SCAN_FLAGS_FAST_MODE) &&
      match_data_gather_on())
  {
    const uint8_t* pd = NULL;
    const uint64_t* po = NULL;
    if (yrb_last_match_data(dev, max_match_data, &pd, &po) == YRB_SUCCESS)
    {
      t_packed_data = pd;
      t_packed_offs = po;
    }
  }

  rc = materialise(
      scanner, block_data, block, res.matches, res.n_matches, scanner->matches, 1, max_match_data, seed_nodes);

  if (rc == ERROR_SUCCESS)
    rc = materialise(
        scanner, block_data, block, res.unconfirmed, res.n_unconfirmed, scanner->unconfirmed_matches, 0,
        max_match_data, seed_nodes);

  free(seed_nodes);
  if (rc != ERROR_SUCCESS)
  {
    YR_B200_UNLOCK_MULTI();
    return rc;
  }

  /* strings that hit YR_MAX_STRING_MATCHES inside this block, in firing order */
  for (uint32_t i = 0; i < res.n_too_many; i++)
  {
    YR_STRING* string = &scanner->rules->strings_table[res.too_many[i]];
    if (yr_bitmask_is_set(scanner->strings_temp_disabled, string->idx))
      continue;
    rc = list_full(scanner, string);
    if (rc != ERROR_SUCCESS)
    {
      YR_B200_UNLOCK_MULTI();
      return rc;
    }
  }
  YR_B200_UNLOCK_MULTI();

  if (timing)
  {
    clock_gettime(CLOCK_MONOTONIC, &t2);
    static __thread struct timespec t_prev_end = {0, 0};
    const double gap_ms = t_prev_end.tv_sec ? (t0.tv_sec - t_prev_end.tv_sec) * 1e3 + (t0.tv_nsec - t_prev_end.tv_nsec) * 1e-6 : 0.0;
    t_prev_end = t2;
    fprintf(stderr,
            "yara_b200: block %zu B: since the previous block of this thread %.3f ms (libyara, caller), engine %.3f ms (scan+h2d %.3f verify %.3f resolve %.3f), materialise %.3f ms, "
            "%zu matches, %llu candidates, %u launches\n",
            block->size, gap_ms, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6, res.stats.ms_scan,
            res.stats.ms_verify, res.stats.ms_resolve, (t2.tv_sec - t1.tv_sec) * 1e3 + (t2.tv_nsec - t1.tv_nsec) * 1e-6,
            res.n_matches, (unsigned long long) res.stats.n_candidates, res.stats.kernel_launches);
  }

  if (scanner->timeout > 0 && yr_stopwatch_elapsed_ns(&scanner->stopwatch) > scanner->timeout)
    return ERROR_SCAN_TIMEOUT;

  return ERROR_SUCCESS;
}
