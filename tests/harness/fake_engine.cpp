/*
 * fake_engine.cpp -- TEST INFRASTRUCTURE ONLY.  A stand-in for libyrb200.so that implements the part
 * of include/yara_b200.h the libyara glue calls (rules, scanners, yrb_scan_host, yrb_scan_batch) by
 * running the product's own verify / resolve device code on the CPU (tests/harness/vm_host.cpp), so
 * that the -m "not gpu" suite can drive yara_b200/host/yr_b200_glue.c -- materialise, the TOO_MANY
 * protocol, seeds across blocks, parked batch results, the pre-filter's note site -- end to end
 * through the unchanged libyara API.
 *
 * It is built into tests/harness/libyrb200cpu.so, and the drop-in's own objects are linked against
 * it as tests/harness/libyara_b200cpu.so (yara_b200/Makefile); nothing outside tests/ loads either.  The product library
 * (yara_b200/libyrb200.so) has no CPU path and fails with YRB_ERROR_NO_DEVICE without a GPU.
 *
 * Not replayed: the scan kernels (the walk of vm_host.cpp stands in for them), shards, device-resident data.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "yara_b200.h"

extern "C" {

/* tests/harness/vm_host.cpp */
typedef struct vmh_result
{
  int32_t error;
  uint32_t error_string;
  uint64_t n_matches;
  uint64_t n_unconfirmed;
  uint32_t n_too_many;
  uint32_t over_cap;
  uint64_t first_cut_key;
  uint64_t error_key;
  uint32_t n_cuts;
  uint32_t pad;
} vmh_result;

typedef struct vmh_stats
{
  uint64_t n_cand[3];
  uint64_t n_raw;
  uint32_t retry_fast, retry_re, retry_fast3, retry_fast4;
  int32_t error;
  uint32_t error_string;
  uint64_t error_key;
} vmh_stats;

int vmh_scan_seeds(
    const void* image, const uint8_t* data, size_t size, uint64_t base, int scan_flags, const yrb_seed* seeds, uint32_t n_seeds,
    yrb_match* matches, size_t cap_m, yrb_match* unconfirmed, size_t cap_u, uint32_t* too_many_out, vmh_result* res,
    vmh_stats* stats);
int vmh_scan_batch(
    const void* image, const uint8_t* data, size_t size, int n_seg, const int64_t* seg_start, const int64_t* seg_len,
    yrb_match* matches, size_t cap_m, uint8_t* item_dirty, vmh_result* res, vmh_stats* stats);
void vmh_set_chain_marks_dirty(int on);
void vmh_set_skip_strings(const uint32_t* strings, uint32_t n);
uint32_t vmh_key_seq_bits(const void* image);

struct yrb_rules
{
  std::vector<uint8_t> image;
  int refs;
  bool root_match;
  uint32_t key_seq_bits;
};

struct yrb_scanner
{
  yrb_rules* rules;
  std::vector<yrb_match> m, u, bm;
  std::vector<uint32_t> tm;
  std::vector<uint8_t> state;
  std::vector<uint64_t> first;
  /* yrb_last_match_data: the "resident copy" of the last yrb_scan_host block and its packed match data */
  std::vector<uint8_t> last_block, g_data;
  std::vector<uint64_t> g_off;
  size_t last_n = 0;
  bool last_host = false;
};

int yrb_device_count(void) { return 1; }
const char* yrb_version(void) { return "yara_b200 TEST stand-in (device code replayed on the CPU)"; }

int yrb_rules_create(const void* image, size_t image_size, int device, yrb_rules** out)
{
  (void) device;
  const yrb_image_header* h = (const yrb_image_header*) image;
  if (image_size < sizeof(*h) || h->magic != YRB_IMAGE_MAGIC || h->version != YRB_IMAGE_VERSION || h->total_size != image_size)
    return YRB_ERROR_BAD_ARGUMENT;
  yrb_rules* r = new yrb_rules;
  r->image.assign((const uint8_t*) image, (const uint8_t*) image + image_size);
  r->refs = 1;
  r->root_match = h->tables_size > 0 && yrb_img_match(r->image.data())[0] != 0;
  r->key_seq_bits = vmh_key_seq_bits(r->image.data());
  if (r->key_seq_bits == 0) { delete r; return YRB_ERROR_BAD_IMAGE; }   /* as api.cu rules_init */
  *out = r;
  return YRB_SUCCESS;
}

void yrb_rules_retain(yrb_rules* r) { if (r) r->refs++; }
void yrb_rules_release(yrb_rules* r) { if (r && --r->refs == 0) delete r; }

int yrb_scanner_create(yrb_rules* rules, yrb_scanner** out)
{
  yrb_scanner* s = new yrb_scanner;
  s->rules = rules;
  rules->refs++;
  *out = s;
  return YRB_SUCCESS;
}

void yrb_scanner_destroy(yrb_scanner* s)
{
  if (!s) return;
  yrb_rules_release(s->rules);
  delete s;
}

int yrb_scanner_set_option(yrb_scanner*, int, uint64_t) { return YRB_SUCCESS; }
size_t yrb_scanner_footprint(yrb_scanner*) { return 0; }
void yrb_scanner_trim(yrb_scanner*) {}

static unsigned g_host_scans = 0, g_batch_calls = 0, g_match_data_calls = 0;
unsigned yrb_test_match_data_calls(void) { return g_match_data_calls; }

int yrb_last_match_data(yrb_scanner* s, uint32_t max_match_data, const uint8_t** data, const uint64_t** offsets)
{
  if (!s || !data || !offsets || !s->last_host || s->last_n == 0) return YRB_ERROR_BAD_ARGUMENT;
  __atomic_fetch_add(&g_match_data_calls, 1, __ATOMIC_RELAXED);
  s->g_off.assign(s->last_n + 1, 0);
  s->g_data.clear();
  for (size_t k = 0; k < s->last_n; k++)
  {
    const int32_t l = s->m[k].match_length;
    const uint64_t len = (uint64_t) (l < 0 ? 0 : (l > (int32_t) max_match_data ? (int32_t) max_match_data : l));
    for (uint64_t j = 0; j < len; j++)
    {
      const int64_t p = s->m[k].offset + (int64_t) j;
      s->g_data.push_back((p >= 0 && (size_t) p < s->last_block.size()) ? s->last_block[(size_t) p] : 0);
    }
    s->g_off[k + 1] = s->g_data.size();
  }
  s->g_data.push_back(0);
  *data = s->g_data.data();
  *offsets = s->g_off.data();
  return YRB_SUCCESS;
}

unsigned yrb_test_host_scans(void) { return g_host_scans; }
unsigned yrb_test_batch_calls(void) { return g_batch_calls; }

int yrb_scan_host(yrb_scanner* s, const uint8_t* data, size_t size, const yrb_scan_params* p, yrb_scan_result* out)
{
  if (!s || !p || !out || p->block_size != 0 || p->defer_resolve)
    return YRB_ERROR_BAD_ARGUMENT;
  __atomic_fetch_add(&g_host_scans, 1, __ATOMIC_RELAXED);
  memset(out, 0, sizeof(*out));
  out->error_string = YRB_NONE;
  static const uint8_t dummy = 0;
  if (data == nullptr) data = &dummy;

  vmh_result res;
  vmh_stats st;
  size_t cap = 1 << 12;
  s->tm.assign(1024, 0);
  for (;;)
  {
    s->m.resize(cap);
    s->u.resize(cap);
    vmh_set_skip_strings(p->skip_strings, p->skip_strings ? p->n_skip_strings : 0);
    if (vmh_scan_seeds(s->rules->image.data(), data, size, p->base, p->scan_flags, p->seeds, p->n_seeds, s->m.data(), cap,
                       s->u.data(), cap, s->tm.data(), &res, &st) != 0)
      return YRB_ERROR_BAD_ARGUMENT;
    if (res.n_matches <= cap && res.n_unconfirmed <= cap)
      break;
    cap = (size_t) std::max(res.n_matches, res.n_unconfirmed);
  }
  /* error precedence as in api.cu run_resolve: the earliest failing candidate in discovery order
     (scanner.c:111) vs. the earliest TOO_MANY abort */
  const unsigned long long SEQ = (1ULL << s->rules->key_seq_bits) - 1;   /* KEY_SEQ_MASK (engine.cuh) */
  unsigned long long first_abort = ~0ULL;
  const uint32_t ntm = res.n_too_many > 1024 ? 1024 : res.n_too_many;
  if (p->too_many_action != 0 && ntm > 0)
  {
    first_abort = res.n_cuts ? res.first_cut_key : 0;
    out->error = YRB_ERROR_TOO_MANY_MATCHES;
    out->error_string = s->tm[0];
  }
  if (res.error != 0 && (res.error_key | SEQ) < (first_abort | SEQ))
  {
    out->error = res.error;
    out->error_string = res.error_string;
  }
  s->last_block.assign(data, data + size);
  s->last_n = (size_t) res.n_matches;
  s->last_host = true;
  out->n_matches = (size_t) res.n_matches;
  out->matches = s->m.data();
  out->n_unconfirmed = (size_t) res.n_unconfirmed;
  out->unconfirmed = s->u.data();
  out->n_too_many = ntm;
  out->too_many = s->tm.data();
  out->seeds_consumed = p->n_seeds > 0 ? 1 : 0;
  out->stats.bytes = size;
  out->stats.n_candidates = st.n_cand[0] + st.n_cand[1] + st.n_cand[2];
  out->stats.n_raw = st.n_raw;
  return YRB_SUCCESS;
}

/* the host half of yrb_scan_batch (api.cu), restated over the CPU replay of its device half */
/* yrb_multi_*: the stand-in has one "device"; a sharded scan must equal a single-device scan by
   definition (include/yara_b200.h), so it IS one here.  What this exercises on the CPU is the
   glue's routing of large single-block scans ($YR_B200_DEVICES), its locking and its release path. */
struct yrb_multi
{
  yrb_rules* rules;
  yrb_scanner* scanner;
  int n;
};
static unsigned g_multi_scans = 0;
unsigned yrb_test_multi_scans(void) { return g_multi_scans; }

int yrb_multi_create(const void* image, size_t image_size, const int* devices, int n_devices, yrb_multi** out)
{
  if (!devices || n_devices < 1) return YRB_ERROR_BAD_ARGUMENT;
  yrb_multi* m = new yrb_multi;
  m->n = n_devices;
  int rc = yrb_rules_create(image, image_size, 0, &m->rules);
  if (rc != YRB_SUCCESS) { delete m; return rc; }
  yrb_scanner_create(m->rules, &m->scanner);
  *out = m;
  return YRB_SUCCESS;
}

void yrb_multi_destroy(yrb_multi* m)
{
  if (!m) return;
  yrb_scanner_destroy(m->scanner);
  yrb_rules_release(m->rules);
  delete m;
}

int yrb_multi_scan_host(yrb_multi* m, const uint8_t* data, size_t size, const yrb_scan_params* p, yrb_scan_result* out)
{
  if (p->n_seeds > 0) return YRB_ERROR_BAD_ARGUMENT;
  g_multi_scans++;
  return yrb_scan_host(m->scanner, data, size, p, out);
}

int yrb_scan_batch(
    yrb_scanner* s, const yrb_batch_item* items, uint32_t n_items, int scan_flags, int batch_flags, yrb_batch_result* out)
{
  if (!s || !out || (n_items > 0 && !items))
    return YRB_ERROR_BAD_ARGUMENT;
  __atomic_fetch_add(&g_batch_calls, 1, __ATOMIC_RELAXED);
  s->last_host = false;
  memset(out, 0, sizeof(*out));
  out->n_items = n_items;
  s->state.assign(n_items ? n_items : 1, YRB_ITEM_DONE);
  s->first.assign((size_t) n_items + 1, 0);
  s->bm.clear();
  out->item_state = s->state.data();
  out->item_first = s->first.data();
  out->matches = s->bm.data();
  if (n_items == 0)
    return YRB_SUCCESS;

  if ((scan_flags & YRB_SCAN_FAST_MODE) || s->rules->root_match)
  {
    std::fill(s->state.begin(), s->state.end(), (uint8_t) YRB_ITEM_RESCAN);
    out->n_rescan = n_items;
    return YRB_SUCCESS;
  }

  std::vector<int64_t> start(n_items), len(n_items);
  std::vector<uint8_t> concat;
  int64_t total = 0;
  for (uint32_t k = 0; k < n_items; k++)
  {
    start[k] = total;
    len[k] = (int64_t) items[k].size;
    total += (len[k] + 15) & ~(int64_t) 15;
  }
  concat.assign((size_t) total + 16, 0xA5);
  for (uint32_t k = 0; k < n_items; k++)
    if (len[k]) memcpy(concat.data() + start[k], items[k].data, (size_t) len[k]);

  std::vector<uint8_t> dirty(n_items, 0);
  vmh_result res;
  vmh_stats st;
  size_t cap = 1 << 12;
  for (;;)
  {
    s->m.resize(cap);
    std::fill(dirty.begin(), dirty.end(), 0);
    vmh_set_chain_marks_dirty((batch_flags & YRB_BATCH_BLOCKS_OF_ONE_SCAN) != 0);
    if (vmh_scan_batch(s->rules->image.data(), concat.data(), (size_t) total, (int) n_items, start.data(), len.data(), s->m.data(),
                       cap, dirty.data(), &res, &st) != 0)
      return YRB_ERROR_BAD_ARGUMENT;
    if (res.n_matches <= cap)
      break;
    cap = (size_t) res.n_matches;
  }
  if (res.over_cap)
    std::fill(dirty.begin(), dirty.end(), 1);

  std::vector<uint32_t> item_of((size_t) res.n_matches);
  for (size_t t = 0; t < res.n_matches; t++)
  {
    const uint32_t k = (uint32_t) (std::upper_bound(start.begin(), start.end(), s->m[t].offset) - start.begin()) - 1;
    item_of[t] = k;
    if (!dirty[k]) s->first[(size_t) k + 1]++;
  }
  for (uint32_t k = 0; k < n_items; k++)
  {
    if (dirty[k]) { s->state[k] = YRB_ITEM_RESCAN; out->n_rescan++; }
    s->first[(size_t) k + 1] += s->first[k];
  }
  s->bm.resize((size_t) s->first[n_items]);
  std::vector<uint64_t> cur(s->first.begin(), s->first.end() - 1);
  for (size_t t = 0; t < res.n_matches; t++)
  {
    const uint32_t k = item_of[t];
    if (dirty[k]) continue;
    yrb_match m = s->m[t];
    m.offset -= start[k];
    s->bm[(size_t) cur[k]++] = m;
  }
  out->matches = s->bm.data();
  return YRB_SUCCESS;
}

} /* extern "C" */
