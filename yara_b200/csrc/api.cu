// api.cu -- the extern "C" layer of libyrb200.so (include/yara_b200.h):
// rules upload, per-scanner device state, and the block-scan pipeline
//
//   H2D slices --> AC scan (per slice) --> verify --> resolve --> D2H records
//
// One CUDA stream per scanner (plus a copy stream for the H2D slices), so the
// up-to-YR_MAX_THREADS scanners that may share one YR_RULES (limits.h:45-51)
// run concurrently against one read-only device image.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cub/cub.cuh>

#include "engine.cuh"
#include "resolve.cuh"

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      if (getenv("YRB_DEBUG")) fprintf(stderr, "yrb200: %s -> %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return (e_ == cudaErrorMemoryAllocation) ? YRB_ERROR_INSUFFICIENT_MEMORY : YRB_ERROR_CUDA;  \
    }                                                                                             \
  } while (0)

static bool trace_on() { static int t = -1; if (t < 0) t = getenv("YRB_TRACE") != nullptr; return t != 0; }
#define TRACE(...) do { if (trace_on()) { fprintf(stderr, "yrb200 trace: " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)

// ------------------------------------------------------------------- structs

struct yrb_rules
{
  size_t l2_window_bytes = 0;   // leading bytes of d_image under the persisting L2 window of every scanner stream
  int device;
  std::atomic<int> refs;
  void* d_image;
  size_t image_size;
  yrb_image_header h;
  DevRules R;
  FilterTables F;
  void* d_filter;
  int num_sms;
};

// Device scratch comes from the device's stream-ordered memory pool (cudaMallocAsync, release threshold = never):
// cudaMalloc / cudaFree synchronise the whole device and serialise on the driver, which a single scanner never
// notices but sixteen scanner threads growing their buffers at the same time do -- `yara -p 16` through the drop-in
// spent 100-300 ms per file in allocations during its first scans (profiles/r02_kernel_notes.txt).  The pool
// hands freed blocks back without either.  YRB_POOLED_ALLOC=0 goes back to cudaMalloc.
static bool pooled_alloc()
{
  static const bool on = !(getenv("YRB_POOLED_ALLOC") && atoi(getenv("YRB_POOLED_ALLOC")) == 0);
  return on;
}

static cudaError_t dev_malloc(void** p, size_t bytes)
{
  if (!pooled_alloc()) return cudaMalloc(p, bytes);
  cudaError_t e = cudaMallocAsync(p, bytes, cudaStreamPerThread);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);   // usable on any stream from here on
  return e;
}

static void dev_free(void* p)
{
  // every user of the block has finished: scans synchronise their streams before they return, and buffers only
  // grow or go between scans
  if (pooled_alloc()) { if (cudaFreeAsync(p, cudaStreamPerThread) != cudaSuccess) { cudaGetLastError(); cudaFree(p); } }
  else cudaFree(p);
}

static void dev_pool_setup(int device)
{
  static std::atomic<unsigned long long> done{0};
  if (!pooled_alloc() || device < 0 || device >= 64 || (done.load() >> device) & 1ull) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
  {
    unsigned long long never = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never);
  }
  cudaGetLastError();
  done.fetch_or(1ull << device);
}

template <typename T>
struct DevBuf
{
  T* p = nullptr;
  size_t cap = 0; // elements
  // Pool memory is private to its device: a cudaMemcpyPeerAsync out of it is staged through the host (measured at 2
  // GPUs: gather + global resolve of a sharded scan 0.52 -> 0.70 ms).  Buffers that other devices copy from -- the
  // Raw records and the per-string error keys of a shard -- are therefore plain cudaMalloc memory.
  bool plain = false;

  // Small, frequently regrown scratch doubles; anything beyond 64 MiB (the resident block above
  // all: a 16 GiB block must not take 32 GiB of HBM, nor a 130 GB one fail) is sized exactly,
  // rounded up to 2 MiB.
  int ensure(size_t n, bool keep = false)
  {
    if (n <= cap) return YRB_SUCCESS;
    size_t ncap = cap ? cap : 1;
    while (ncap < n) ncap *= 2;
    if (ncap * sizeof(T) > ((size_t) 64 << 20))
    {
      const size_t gran = ((size_t) 2 << 20) / sizeof(T) ? ((size_t) 2 << 20) / sizeof(T) : 1;
      ncap = (n + gran - 1) / gran * gran;
    }
    T* np = nullptr;
    cudaError_t e = plain ? cudaMalloc((void**) &np, ncap * sizeof(T)) : dev_malloc((void**) &np, ncap * sizeof(T));
    if (e != cudaSuccess) { cudaGetLastError(); return YRB_ERROR_INSUFFICIENT_MEMORY; }
    if (keep && p && cap) cudaMemcpy(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice);
    if (p) { if (plain) cudaFree(p); else dev_free(p); }
    p = np;
    cap = ncap;
    return YRB_SUCCESS;
  }
  void release() { if (p) { if (plain) cudaFree(p); else dev_free(p); } p = nullptr; cap = 0; }
};

template <typename T>
struct PinBuf
{
  T* p = nullptr;
  size_t cap = 0;

  int ensure(size_t n)
  {
    if (n <= cap) return YRB_SUCCESS;
    size_t ncap = cap ? cap : 1;
    while (ncap < n) ncap *= 2;
    T* np = nullptr;
    if (cudaHostAlloc((void**) &np, ncap * sizeof(T), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
    { cudaGetLastError(); return YRB_ERROR_INSUFFICIENT_MEMORY; }
    if (p) cudaFreeHost(p);
    p = np;
    cap = ncap;
    return YRB_SUCCESS;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct yrb_scanner
{
  yrb_rules* rules;
  int device;
  cudaStream_t stream, copy_stream;
  SideStreams side = {};       // verifier classes / chain replay next to each other (engine.cuh)
  bool use_side = true;        // YRB_SIDE_STREAMS=0: everything on `stream` (measurements)
  cudaEvent_t ev[8];   // 0-3 phases, 5 copy fork, 6-7 around the last filter-kernel launch
  bool kernel_timed = false;
  std::vector<cudaEvent_t> slice_ev;
  std::vector<cudaEvent_t> slice_done;   // recorded behind each slice's scan kernel (timeout checks)

  int opt_kernel = YRB_KERNEL_AUTO;
  size_t opt_h2d_chunk = 64u << 20;
  bool shard_of_multi = false;   // yrb_multi worker: a deferred scan leaves its Raw records (and error keys) on the device
  uint32_t shard_n_chain_raw = 0, shard_any_error = 0;

  DevBuf<uint8_t> d_data;
  DevBuf<Cand> d_cand[NUM_CLS];
  DevBuf<Raw> d_raw;
  DevBuf<uint8_t> ws_fast1, ws_fast2, ws_fast3, ws_re1, ws_re2;
  DevBuf<uint32_t> retry_fast, retry_fast3, retry_re, retry_re0;
  DevBuf<uint32_t> cs_k0, cs_k1, cs_v0, cs_v1;   // candidates in AC-match order (VerifyWs::sort_*)
  DevBuf<uint8_t> cs_tmp;
  VerifyWs vws = {};
  ScanStatus* d_status = nullptr;
  ScanStatus* h_status = nullptr;

  // resolve scratch
  DevBuf<uint64_t> k0, k1, kF;
  DevBuf<uint32_t> ix0, ix1, scr32;
  DevBuf<uint8_t> cub_tmp;
  DevBuf<int32_t> c_nxt, c_prv, c_clen, c_mlen, c_uvn, s_lists;
  DevBuf<uint8_t> c_where, flags8;
  DevBuf<Raw> sraw;
  DevBuf<yrb_seed> d_seeds;
  DevBuf<int64_t> seed_base;
  DevBuf<yrb_match> finals, unconf;
  DevBuf<unsigned long long> str_min_key, str_cut_key, str_err;
  DevBuf<uint32_t> str_count, too_many;

  PinBuf<yrb_match> h_finals, h_unconf;
  PinBuf<uint32_t> h_too_many, h_str_count;
  PinBuf<Raw> h_raw;
  size_t h_raw_n = 0;

  // batch scans (yrb_scan_batch)
  DevBuf<int64_t> seg_tab;
  DevBuf<uint8_t> item_dirty;
  PinBuf<int64_t> h_seg_tab;
  PinBuf<uint8_t> h_item_dirty;
  std::vector<uint8_t> b_state;
  std::vector<uint64_t> b_first;
  std::vector<yrb_match> b_matches;

  // last device-resident result (for yrb_gather_match_data)
  size_t last_n_final = 0;
  ScanView last_view;
  DevBuf<uint64_t> g_offsets;
  DevBuf<uint8_t> g_data;
  PinBuf<uint64_t> h_g_offsets;   // yrb_last_match_data: the same, in pinned host memory owned by the scanner
  PinBuf<uint8_t> h_g_data;
  bool last_scan_resident = false;   // the last scan copied its block into d_data (yrb_scan_host) and it is still there
};

// ------------------------------------------------------------------- helpers

extern "C" int yrb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char* yrb_version(void) { return "yara_b200 0.1 (sm_100a; libyara 4.2.1 hot path)"; }

// ---------------------------------------------------------------------- rules

// The discovery key orders the entries of a state's YR_AC_MATCH chain by their position (emit.cuh); how many of
// the key's bits that takes is a property of the ruleset (engine.cuh: yrb_key_seq_bits).  A chain too long for
// any split would make later entries share one key and the "first wins / last wins" rules of the resolve
// ambiguous: such a ruleset is refused loudly instead of saturating.
extern "C" void yrb_rules_release(yrb_rules* r);

static int rules_init(yrb_rules* r, const void* image, size_t image_size, int device)
{
  const yrb_image_header* h = (const yrb_image_header*) image;
  CK(cudaSetDevice(device));

  r->device = device;
  r->refs = 1;
  r->image_size = image_size;
  r->h = *h;
  r->d_image = nullptr;
  r->d_filter = nullptr;

  if (h->max_re_stack >= 0xFFFFu) return YRB_ERROR_BAD_IMAGE;   // DevRules::stack_words is 16 bits (real rulesets: a handful)
  const uint32_t longest_chain = yrb_longest_match_chain(image);
  r->R.key_seq_bits = (uint16_t) yrb_key_seq_bits(longest_chain);
  if (r->R.key_seq_bits == 0)
  {
    fprintf(stderr, "yrb200: a state's YR_AC_MATCH chain has %u entries (more than %u); the discovery key cannot order them\n",
            longest_chain, KEY_CHAIN_MAX(KEY_SEQ_BITS_MIN));
    return YRB_ERROR_BAD_IMAGE;
  }

  dev_pool_setup(device);
  CK(cudaMalloc(&r->d_image, image_size));
  CK(cudaMemcpy(r->d_image, image, image_size, cudaMemcpyHostToDevice));

  const uint8_t* b = (const uint8_t*) r->d_image;
  r->R.T = (const uint32_t*) (b + h->off_trans);
  r->R.M = (const uint32_t*) (b + h->off_match);
  r->R.acm = (const yrb_acm*) (b + h->off_acm);
  r->R.S = (const yrb_string*) (b + h->off_strings);
  r->R.lit = b + h->off_lit;
  r->R.code = b + h->off_code;
  r->R.lower = b + h->off_lower;
  r->R.alter = b + h->off_alter;
  r->R.tables_size = h->tables_size;
  r->R.num_acm = h->num_acm;
  r->R.num_strings = h->num_strings;
  r->R.stack_words = (uint16_t) (h->max_re_stack + 1);

  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  r->num_sms = prop.multiProcessorCount;

  // The rules image (transition/match tables, YR_AC_MATCH entries, strings, RE code: a few MB that every
  // candidate touches) is kept in L2 under a persisting access-policy window on each scanner's streams, so that
  // the block bytes streaming through -- tens of GB per scan -- do not evict it (north star: "the transition
  // table ... in an L2-persisting window").  The carve-out is set once per device; YRB_L2_PERSIST=0 turns it off.
  static const bool persist_on = !(getenv("YRB_L2_PERSIST") && atoi(getenv("YRB_L2_PERSIST")) == 0);
  r->l2_window_bytes = 0;
  if (persist_on && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0)
  {
    size_t want = image_size;
    if (want > (size_t) prop.accessPolicyMaxWindowSize) want = (size_t) prop.accessPolicyMaxWindowSize;
    if (want > (size_t) prop.persistingL2CacheMaxSize) want = (size_t) prop.persistingL2CacheMaxSize;
    size_t cur = 0;
    cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
    if (cur < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); want = cur; }
    r->l2_window_bytes = want;
  }

  memset(&r->F, 0, sizeof(r->F));
  int rc = build_filter_tables(image, image_size, device, &r->F, &r->d_filter);
  if (rc != YRB_SUCCESS) return rc;

  // cudaMemcpy from pageable memory may return before the DMA has landed, and the
  // scanners run on non-blocking streams that do not order after the legacy stream
  CK(cudaDeviceSynchronize());
  return YRB_SUCCESS;
}

extern "C" int yrb_rules_create(const void* image, size_t image_size, int device, yrb_rules** out)
{
  if (image == nullptr || out == nullptr || image_size < sizeof(yrb_image_header))
    return YRB_ERROR_BAD_ARGUMENT;

  const yrb_image_header* h = (const yrb_image_header*) image;
  if (h->magic != YRB_IMAGE_MAGIC || h->version != YRB_IMAGE_VERSION || h->total_size != image_size)
    return YRB_ERROR_BAD_IMAGE;

  if (device < 0 || device >= yrb_device_count())
    return YRB_ERROR_NO_DEVICE;

  yrb_rules* r = new (std::nothrow) yrb_rules();
  if (!r) return YRB_ERROR_INSUFFICIENT_MEMORY;

  r->refs = 1;
  r->device = device;
  const int rc = rules_init(r, image, image_size, device);
  if (rc != YRB_SUCCESS)
  {
    yrb_rules_release(r);   // frees whatever was built
    return rc;
  }
  *out = r;
  return YRB_SUCCESS;
}

extern "C" void yrb_rules_retain(yrb_rules* r) { if (r) r->refs++; }

extern "C" int yrb_rules_filter_info(const yrb_rules* r, yrb_filter_info* out)
{
  if (!r || !out) return YRB_ERROR_BAD_ARGUMENT;
  out->usable = r->F.usable;
  out->variant = r->F.variant;
  out->n_windows = r->F.n_windows;
  out->fill3 = r->F.fill3;
  out->has_depth1 = r->F.has_d1;
  out->has_depth2 = r->F.has_d2;
  out->root_match = r->F.root_match;
  return YRB_SUCCESS;
}

extern "C" void yrb_rules_release(yrb_rules* r)
{
  if (!r) return;
  if (--r->refs == 0)
  {
    cudaSetDevice(r->device);
    if (r->d_image) cudaFree(r->d_image);
    if (r->d_filter) cudaFree(r->d_filter);
    delete r;
  }
}

// -------------------------------------------------------------------- scanner

extern "C" void yrb_scanner_destroy(yrb_scanner* s);

static int scanner_init(yrb_scanner* s)
{
  s->d_raw.plain = true;     // what the first device of a sharded scan copies from its peers (DevBuf::plain)
  s->str_err.plain = true;
  CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 8; i++) CK(cudaEventCreate(&s->ev[i]));
  for (int i = 0; i < 2; i++) CK(cudaStreamCreateWithFlags(&s->side.aux[i], cudaStreamNonBlocking));
  for (int i = 0; i < 6; i++) CK(cudaEventCreateWithFlags(&s->side.ev[i], cudaEventDisableTiming));
  if (const char* e = getenv("YRB_SIDE_STREAMS")) s->use_side = atoi(e) != 0;
  if (s->rules->l2_window_bytes > 0)
  {
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    av.accessPolicyWindow.base_ptr = s->rules->d_image;
    av.accessPolicyWindow.num_bytes = s->rules->l2_window_bytes;
    av.accessPolicyWindow.hitRatio = 1.0f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStream_t all[3] = {s->stream, s->side.aux[0], s->side.aux[1]};
    for (cudaStream_t st : all)
      if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
  }
  CK(cudaMalloc((void**) &s->d_status, sizeof(ScanStatus)));
  CK(cudaHostAlloc((void**) &s->h_status, sizeof(ScanStatus), cudaHostAllocDefault));
  memset(s->h_status, 0, sizeof(ScanStatus));
  return YRB_SUCCESS;
}

extern "C" int yrb_scanner_create(yrb_rules* rules, yrb_scanner** out)
{
  if (!rules || !out) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(rules->device));

  yrb_scanner* s = new (std::nothrow) yrb_scanner();   // value-initialised: every handle starts out null
  if (!s) return YRB_ERROR_INSUFFICIENT_MEMORY;

  s->rules = rules;
  s->device = rules->device;
  yrb_rules_retain(rules);
  if (const char* e = getenv("YRB_SCAN_KERNEL")) // test hook: 1 = walk, 2 = filter
    s->opt_kernel = atoi(e);

  const int rc = scanner_init(s);
  if (rc != YRB_SUCCESS)
  {
    yrb_scanner_destroy(s);   // tolerates a partially built scanner
    return rc;
  }
  *out = s;
  return YRB_SUCCESS;
}

extern "C" void yrb_shard_export_destroy(yrb_scanner* s);

extern "C" void yrb_scanner_destroy(yrb_scanner* s)
{
  if (!s) return;
  yrb_shard_export_destroy(s);
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
  for (int i = 0; i < 2; i++) if (s->side.aux[i]) { cudaStreamSynchronize(s->side.aux[i]); cudaStreamDestroy(s->side.aux[i]); }
  for (int i = 0; i < 6; i++) if (s->side.ev[i]) cudaEventDestroy(s->side.ev[i]);
  s->d_data.release();
  for (int c = 0; c < NUM_CLS; c++) s->d_cand[c].release();
  s->d_raw.release(); s->ws_fast1.release(); s->ws_fast2.release(); s->ws_re1.release(); s->ws_re2.release();
  s->retry_fast.release(); s->retry_fast3.release(); s->retry_re.release(); s->retry_re0.release(); s->ws_fast3.release();
  s->cs_k0.release(); s->cs_k1.release(); s->cs_v0.release(); s->cs_v1.release(); s->cs_tmp.release();
  s->k0.release(); s->k1.release(); s->kF.release(); s->ix0.release(); s->ix1.release(); s->scr32.release();
  s->cub_tmp.release(); s->c_nxt.release(); s->c_prv.release(); s->c_clen.release(); s->c_mlen.release(); s->c_uvn.release(); s->sraw.release(); s->d_seeds.release(); s->seed_base.release();
  s->s_lists.release(); s->c_where.release(); s->flags8.release(); s->finals.release(); s->unconf.release();
  s->str_min_key.release(); s->str_cut_key.release(); s->str_err.release(); s->str_count.release(); s->too_many.release();
  s->g_offsets.release(); s->g_data.release(); s->h_g_offsets.release(); s->h_g_data.release();
  s->seg_tab.release(); s->item_dirty.release(); s->h_seg_tab.release(); s->h_item_dirty.release();
  s->h_finals.release(); s->h_unconf.release(); s->h_too_many.release(); s->h_str_count.release(); s->h_raw.release();
  if (s->d_status) cudaFree(s->d_status);
  if (s->h_status) cudaFreeHost(s->h_status);
  for (auto e : s->slice_ev) cudaEventDestroy(e);
  for (auto e : s->slice_done) cudaEventDestroy(e);
  for (int i = 0; i < 8; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  if (s->stream) cudaStreamDestroy(s->stream);
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  cudaGetLastError();
  yrb_rules_release(s->rules);
  delete s;
}

extern "C" size_t yrb_scanner_footprint(yrb_scanner* s)
{
  return s ? s->d_data.cap : 0;
}

extern "C" void yrb_scanner_trim(yrb_scanner* s)
{
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  cudaStreamSynchronize(s->copy_stream);
  s->d_data.release();
  if (pooled_alloc())
  {
    // give the block buffer back to the device, not just to the pool (this is what trimming is for)
    cudaMemPool_t pool;
    cudaStreamSynchronize(cudaStreamPerThread);
    if (cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
  }
}

extern "C" int yrb_scanner_set_option(yrb_scanner* s, int option, uint64_t value)
{
  if (!s) return YRB_ERROR_BAD_ARGUMENT;
  switch (option)
  {
  case YRB_OPT_SCAN_KERNEL:
    if (value > YRB_KERNEL_FILTER) return YRB_ERROR_BAD_ARGUMENT;
    s->opt_kernel = (int) value;
    return YRB_SUCCESS;
  case YRB_OPT_H2D_CHUNK:
    if (value < 4096) return YRB_ERROR_BAD_ARGUMENT;
    s->opt_h2d_chunk = (size_t) value & ~(size_t) 255;
    return YRB_SUCCESS;
  case YRB_OPT_SHARD_ON_DEVICE:
    s->shard_of_multi = value != 0;
    return YRB_SUCCESS;
  }
  return YRB_ERROR_BAD_ARGUMENT;
}

// Multi-block scans: the unconfirmed chain-part matches earlier blocks left behind
// re-enter the resolve stage as records that sort before anything found in this block.
__global__ void seeds_to_raw_kernel(const yrb_seed* seeds, uint32_t n, DevRules R, Raw* out, int64_t* seed_base)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const yrb_seed sd = seeds[k];
  Raw r;
  r.key = (uint64_t) (uint32_t) sd.chain_length;
  r.offset = sd.offset;
  r.string_idx = sd.string_idx;
  r.length = sd.match_length;
  r.misc = (uint32_t) sd.xor_key | (k << 8) | RAW_SEED | (sd.kind ? RAW_SEED_MARK : 0u);   // k < 2^22
  r.group = sd.string_idx < R.num_strings ? R.S[sd.string_idx].chain_head : YRB_NONE;
  out[k] = r;
  seed_base[k] = sd.base;
}

// ------------------------------------------------------------- scan pipeline

static int pick_kernel(const yrb_scanner* s)
{
  if (s->opt_kernel == YRB_KERNEL_WALK) return YRB_KERNEL_WALK;
  if (getenv("YRB_DEBUG") && !s->rules->F.usable)
    fprintf(stderr, "yrb200: prefilter not usable (root_match=%d d1=%d d2=%d), using the walk kernel\n",
            s->rules->F.root_match, s->rules->F.has_d1, s->rules->F.has_d2);
  if (s->rules->F.usable) return YRB_KERNEL_FILTER;
  return YRB_KERNEL_WALK;
}

static void launch_scan_slice(yrb_scanner* s, int kernel, const ScanView& V, const CandOut& co, int64_t lo, int64_t hi)
{
  if (kernel == YRB_KERNEL_FILTER)
  {
    launch_ac_scan_filter(s->rules->R, s->rules->F, V, co, lo, hi, s->rules->num_sms, s->stream, s->ev[6], s->ev[7]);
    s->kernel_timed = true;
  }
  else
    launch_ac_scan_walk(s->rules->R, V, co, lo, hi, s->stream);
}

static int read_status(yrb_scanner* s)
{
  CK(cudaMemcpyAsync(s->h_status, s->d_status, sizeof(ScanStatus), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return YRB_SUCCESS;
}

static int ensure_resolve_scratch(yrb_scanner* s, uint32_t n)
{
  const uint32_t ns = s->rules->R.num_strings ? s->rules->R.num_strings : 1;
  int rc = 0;
  rc |= s->k0.ensure(n); rc |= s->k1.ensure(n); rc |= s->kF.ensure(n);
  rc |= s->ix0.ensure(n); rc |= s->ix1.ensure(n); rc |= s->scr32.ensure(n);
  rc |= s->flags8.ensure(n); rc |= s->finals.ensure(n); rc |= s->unconf.ensure(n);
  rc |= s->cub_tmp.ensure(resolve_cub_temp_bytes(n));
  if (s->rules->h.num_chains > 0)
  {
    rc |= s->c_nxt.ensure(n); rc |= s->c_prv.ensure(n); rc |= s->c_clen.ensure(n); rc |= s->c_mlen.ensure(n);
    rc |= s->c_where.ensure(n); rc |= s->c_uvn.ensure(n); rc |= s->sraw.ensure(n); rc |= s->s_lists.ensure((size_t) 8 * ns);
  }
  rc |= s->str_min_key.ensure(ns); rc |= s->str_cut_key.ensure(ns); rc |= s->str_count.ensure(ns);
  rc |= s->too_many.ensure(1024);
  return rc ? YRB_ERROR_INSUFFICIENT_MEMORY : YRB_SUCCESS;
}

static void fill_work(yrb_scanner* s, ResolveWork& w, const Raw* raw, uint32_t n, int scan_flags)
{
  w.raw = raw;
  w.n_raw = n;
  w.R = s->rules->R;
  w.scan_flags = scan_flags;
  w.has_chains = s->rules->h.num_chains > 0;   // refined by the callers: chain records or seeds present
  w.use_cut = 0;
  w.seed_base = s->seed_base.p;
  w.cur_base = 0;
  w.n_seg = 0;
  w.seg_start = w.seg_len = nullptr;
  w.item_dirty = nullptr;
  w.k0 = s->k0.p; w.k1 = s->k1.p; w.kF = s->kF.p; w.ix0 = s->ix0.p; w.ix1 = s->ix1.p; w.c_scratch_u32 = s->scr32.p;
  w.cub_tmp = s->cub_tmp.p; w.cub_tmp_bytes = s->cub_tmp.cap;
  w.c_nxt = s->c_nxt.p; w.c_prv = s->c_prv.p; w.c_clen = s->c_clen.p; w.c_mlen = s->c_mlen.p; w.c_uvn = s->c_uvn.p; w.sraw = s->sraw.p; w.c_where = s->c_where.p;
  w.flags = s->flags8.p;
  w.str_min_key = s->str_min_key.p; w.str_cut_key = s->str_cut_key.p; w.str_count = s->str_count.p;
  w.s_lists = s->s_lists.p;
  w.finals = s->finals.p; w.unconf = s->unconf.p;
  w.side = s->use_side ? &s->side : nullptr;
  w.finals_host = nullptr;
  w.n_plain_max = 0;
  w.too_many = s->too_many.p; w.too_many_cap = 1024;
  w.st = s->d_status;
}

// Resolve `n` raw records resident at d_raw; fills result (matches, unconfirmed, too_many, error).
// `n_chain_records`: how many of the records are chain parts (Raw.group set), or -1 when the caller does not know.
static int run_resolve(yrb_scanner* s, const Raw* d_raw, uint32_t n, const yrb_scan_params* params, yrb_scan_result* res,
                       unsigned* launches, bool chain_records, int64_t n_chain_records = -1)
{
  ScanStatus* hs = s->h_status;
  res->n_matches = res->n_unconfirmed = 0;
  res->n_too_many = 0;
  std::vector<std::pair<unsigned long long, uint32_t>> cuts; // (cut key, string)

  if (n > 0)
  {
    int rc = ensure_resolve_scratch(s, n);
    if (rc != YRB_SUCCESS) return rc;
    if (s->h_finals.ensure(n) || s->h_too_many.ensure(2048)) return YRB_ERROR_INSUFFICIENT_MEMORY;

    ResolveWork w;
    fill_work(s, w, d_raw, n, params->scan_flags);
    w.cur_base = (int64_t) params->base;
    if (!chain_records) w.has_chains = 0;   // nothing for the chain kernels to do: three launches less

    for (int pass = 0;; pass++)
    {
      // One queue of kernels, the result records (at most n) and the status come back with ONE synchronisation.
      // finals = [plain strings' lists: n_sel records][confirmed lists of chain heads: n_chain_out records].  At
      // most n_plain = n - (chain parts) records are plain, so finals[0, n_plain) can leave for the host as soon
      // as it is written, while the chain replay (the long pole of the resolve) is still running on the side
      // stream; what the replay adds lies in [n_sel, n) and follows.  Only when records were deduplicated
      // (n_sel < n_plain) does the chain part begin inside the range that was copied early: one more small copy.
      uint32_t n_plain = 0;   // chain parts among the records, count unknown: everything is copied after the replay
      if (!w.has_chains)
        n_plain = n;
      else if (n_chain_records >= 0 && (uint64_t) n_chain_records <= n)
        n_plain = n - (uint32_t) n_chain_records;
      w.finals_host = s->h_finals.p;
      w.n_plain_max = n_plain;
      launch_resolve(w, s->stream, launches);
      if (n_plain < n)
        CK(cudaMemcpyAsync(s->h_finals.p + n_plain, s->finals.p + n_plain, (size_t) (n - n_plain) * sizeof(yrb_match),
                           cudaMemcpyDeviceToHost, s->stream));
      rc = read_status(s);
      if (rc != YRB_SUCCESS) return rc;
      if (n_plain < n && (uint32_t) hs->n_sel < n_plain && hs->n_chain_out > 0)
      {
        size_t hi = (size_t) hs->n_sel + (size_t) hs->n_chain_out;
        if (hi > n_plain) hi = n_plain;
        TRACE("resolve: %u plain records deduplicated, chain part re-copied from %d", n_plain - (uint32_t) hs->n_sel, hs->n_sel);
        CK(cudaMemcpyAsync(s->h_finals.p + hs->n_sel, s->finals.p + hs->n_sel, (hi - (size_t) hs->n_sel) * sizeof(yrb_match),
                           cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
      }
      if (!hs->over_cap || pass == 1) break;

      // rare, exact path: find each over-cap string's cut key, then resolve again without the cut records
      const uint32_t ns = s->rules->R.num_strings;
      if (s->h_str_count.ensure(ns) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
      CK(cudaMemsetAsync(s->str_cut_key.p, 0xFF, sizeof(unsigned long long) * ns, s->stream));
      CK(cudaMemcpyAsync(s->h_str_count.p, s->str_count.p, sizeof(uint32_t) * ns, cudaMemcpyDeviceToHost, s->stream));
      CK(cudaStreamSynchronize(s->stream));
      for (uint32_t t = 0; t < ns; t++)
      {
        if (s->h_str_count.p[t] > YRB_MAX_STRING_MATCHES)
        {
          unsigned long long cut = ~0ULL;
          rc = resolve_cap_cut(w, t, s->stream, &cut);
          if (rc != YRB_SUCCESS) return rc;
          if (cut != ~0ULL) cuts.push_back(std::make_pair(cut, t));
        }
      }
      w.use_cut = 1;
      CK(cudaMemsetAsync(&s->d_status->n_chain_out, 0, sizeof(unsigned long long), s->stream));
      CK(cudaMemsetAsync(&s->d_status->n_unconf, 0, sizeof(unsigned long long), s->stream));
      CK(cudaMemsetAsync(&s->d_status->n_too_many, 0, sizeof(unsigned int), s->stream));
      CK(cudaMemsetAsync(&s->d_status->over_cap, 0, sizeof(unsigned int), s->stream));
    }

    const size_t nf = (size_t) hs->n_sel + (size_t) hs->n_chain_out, nu = (size_t) hs->n_unconf;
    uint32_t ntm = hs->n_too_many > 1024 ? 1024 : hs->n_too_many;

    if (nu || ntm)
    {
      if (s->h_unconf.ensure(nu ? nu : 1)) return YRB_ERROR_INSUFFICIENT_MEMORY;
      if (nu) CK(cudaMemcpyAsync(s->h_unconf.p, s->unconf.p, nu * sizeof(yrb_match), cudaMemcpyDeviceToHost, s->stream));
      if (ntm) CK(cudaMemcpyAsync(s->h_too_many.p, s->too_many.p, ntm * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
      CK(cudaStreamSynchronize(s->stream));
    }

    // cut strings fire their TOO_MANY callback in discovery order of the cut
    std::sort(cuts.begin(), cuts.end());
    for (auto& c : cuts) if (ntm < 2048) s->h_too_many.p[ntm++] = c.second;

    if (getenv("YRB_TIMING"))
      fprintf(stderr, "yrb200: resolve %u raw records; chain steps: insert walks %u, prune %u, unvisited scans %u\n", n,
              hs->dbg[0], hs->dbg[1], hs->dbg[2]);

    res->n_matches = nf;
    res->n_unconfirmed = nu;
    res->n_too_many = ntm;
    s->last_n_final = nf;
  }
  else
  {
    s->last_n_final = 0;
  }

  res->matches = s->h_finals.p;
  res->unconfirmed = s->h_unconf.p;
  res->too_many = s->h_too_many.p;

  if (hs->any_error)
  {
    // rare: pick the earliest failing candidate the reference would really have verified.  Strings it does not
    // verify in this block because of what earlier blocks of the scan did (params->skip_strings) have no say.
    for (uint32_t k = 0; k < params->n_skip_strings && params->skip_strings != nullptr; k++)
      if (params->skip_strings[k] < s->rules->R.num_strings)
        CK(cudaMemsetAsync(s->str_err.p + params->skip_strings[k], 0xFF, sizeof(unsigned long long), s->stream));
    ResolveWork w;
    fill_work(s, w, d_raw, n, params->scan_flags);
    w.use_cut = !cuts.empty();
    launch_error_select(w, s->str_err.p, n > 0, s->stream);
    int rc = read_status(s);
    if (rc != YRB_SUCCESS) return rc;
  }

  // error precedence: the earliest failing candidate in discovery order
  // (scanner.c:111 GOTO_EXIT_ON_ERROR) vs. the earliest TOO_MANY abort
  unsigned long long first_abort = ~0ULL;
  if (params->too_many_action != 0 && res->n_too_many > 0)
  {
    first_abort = cuts.empty() ? 0 : cuts[0].first;
    res->error = YRB_ERROR_TOO_MANY_MATCHES;
    res->error_string = res->too_many[0];
  }
  const unsigned long long err_key = ~hs->err_key_inv;
  const unsigned long long seq_mask = KEY_SEQ_MASK(s->rules->R.key_seq_bits);
  if (hs->err_key_inv != 0ULL && (err_key | seq_mask) < (first_abort | seq_mask))
  {
    res->error = (int32_t) (err_key & seq_mask);
    res->error_string = hs->err_string1 ? hs->err_string1 - 1 : YRB_NONE;
  }
  return YRB_SUCCESS;
}

// Layout of one yrb_scan_batch call: where each item sits in the concatenation.
struct BatchPlan
{
  const yrb_batch_item* items;
  uint32_t n;
  const int64_t* start;     // host copy of the segment starts (pinned, n entries)
  bool one_allocation;      // the concatenation IS the host layout: copy it in large slices
  const int64_t* d_start;
  const int64_t* d_len;
  uint8_t* d_dirty;
  bool blocks_of_one_scan;  // YRB_BATCH_BLOCKS_OF_ONE_SCAN: items with chain-part matches are flagged for the per-item path
};

static int scan_impl(
    yrb_scanner* s, const uint8_t* h_data, const uint8_t* d_data_in, size_t size, const yrb_scan_params* params,
    yrb_scan_result* res, const BatchPlan* bp = nullptr)
{
  if (!s || !params || !res || (size > 0 && !h_data && !d_data_in)) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(s->device));

  memset(res, 0, sizeof(*res));
  res->error_string = YRB_NONE;
  unsigned launches = 0;
  s->kernel_timed = false;
  s->last_scan_resident = (h_data != nullptr && bp == nullptr);

  ScanView V;
  memset(&V, 0, sizeof(V));
  V.g0 = (int64_t) params->data_offset;
  V.g1 = V.g0 + (int64_t) size;
  V.gsize = params->block_size ? (int64_t) params->block_size : (int64_t) size;
  V.own_lo = params->block_size ? (int64_t) params->own_begin : 0;
  V.own_hi = params->block_size ? (int64_t) params->own_end : (int64_t) size;
  V.base = params->base;
  V.scan_flags = params->scan_flags;
  if (bp)
  {
    V.n_seg = (int32_t) bp->n;
    V.seg_start = bp->d_start;
    V.seg_len = bp->d_len;
  }

  if (V.gsize >= (1LL << 38) || V.g1 > V.gsize || V.own_hi > V.gsize || V.own_lo > V.own_hi)
    return YRB_ERROR_BAD_ARGUMENT;

  const uint8_t* d_data = d_data_in;
  if (h_data)
  {
    int rc = s->d_data.ensure(size ? size + 16 : 16);
    if (rc != YRB_SUCCESS) return rc;
    d_data = s->d_data.p;
  }
  V.d = d_data;

  // bytes whose events this call owns, clipped to what is locally present
  int64_t lo = V.own_lo > V.g0 ? V.own_lo : V.g0;
  int64_t hi = V.own_hi < V.g1 ? V.own_hi : V.g1;
  // events i in (own_lo, own_hi] are produced by consuming bytes [own_lo, own_hi)
  if (hi < lo) hi = lo;
  res->stats.bytes = (uint64_t) (hi - lo);

  const int kernel = pick_kernel(s);
  res->stats.scan_kernel = (uint32_t) kernel;

  // candidate capacities: generous first guess, exact retry on overflow
  size_t cap_guess = (size_t) ((hi - lo) / 2048) + 65536;
  for (int c = 0; c < NUM_CLS; c++)
    if (s->d_cand[c].ensure(cap_guess) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;

  CK(cudaEventRecord(s->ev[0], s->stream));

  // ---- H2D slices, each followed by its AC scan slice
  const size_t chunk = s->opt_h2d_chunk;
  size_t n_slices = (h_data && !(bp && !bp->one_allocation)) ? (size + chunk - 1) / chunk : 0;
  while (s->slice_ev.size() < n_slices)
  {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s->slice_ev.push_back(e);
  }
  while (params->timeout_ns != 0 && s->slice_done.size() < n_slices)
  {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s->slice_done.push_back(e);
  }
  const std::chrono::steady_clock::time_point t_begin = std::chrono::steady_clock::now();

  ScanStatus* hs = s->h_status;

  {
    const uint32_t ns = s->rules->R.num_strings ? s->rules->R.num_strings : 1;
    if (s->str_err.ensure(ns) != YRB_SUCCESS || s->str_min_key.ensure(ns) != YRB_SUCCESS ||
        s->str_cut_key.ensure(ns) != YRB_SUCCESS || s->str_count.ensure(ns) != YRB_SUCCESS)
      return YRB_ERROR_INSUFFICIENT_MEMORY;
  }

  for (int attempt = 0;; attempt++)
  {
    CK(cudaMemsetAsync(s->d_status, 0, sizeof(ScanStatus), s->stream));

    CandOut co;
    for (int c = 0; c < NUM_CLS; c++) { co.cand[c] = s->d_cand[c].p; co.cap[c] = s->d_cand[c].cap; }
    co.st = s->d_status;

    if (V.own_lo == 0 && !bp)   // batches are refused for rulesets with root-state matches
    {
      launch_ac_event0(s->rules->R, V, co, s->stream);
      launches++;
    }

    if (bp && !bp->one_allocation && attempt == 0)
    {
      // one copy per item into its packed place; a scan slice per ~chunk bytes of items
      CK(cudaEventRecord(s->ev[5], s->stream));
      CK(cudaStreamWaitEvent(s->copy_stream, s->ev[5], 0));
      size_t n_groups = 0;
      int64_t glo = 0;
      for (uint32_t k = 0; k < bp->n; k++)
      {
        const uint64_t nb = bp->items[k].size;
        if (nb) CK(cudaMemcpyAsync(s->d_data.p + bp->start[k], bp->items[k].data, nb, cudaMemcpyHostToDevice, s->copy_stream));
        const int64_t end = bp->start[k] + (int64_t) nb;
        if (k + 1 == bp->n || (size_t) (end - glo) >= chunk)
        {
          const int64_t ghi = (k + 1 == bp->n) ? (int64_t) size : end;
          if (s->slice_ev.size() <= n_groups)
          {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->slice_ev.push_back(e);
          }
          CK(cudaEventRecord(s->slice_ev[n_groups], s->copy_stream));
          CK(cudaStreamWaitEvent(s->stream, s->slice_ev[n_groups], 0));
          n_groups++;
          if (ghi > glo) { launch_scan_slice(s, kernel, V, co, glo, ghi); launches++; }
          glo = ghi;
        }
      }
    }
    else if (h_data && attempt == 0)
    {
      CK(cudaEventRecord(s->ev[5], s->stream));
      CK(cudaStreamWaitEvent(s->copy_stream, s->ev[5], 0)); // d_data may still be read by the previous scan
      for (size_t k = 0; k < n_slices; k++)
      {
        if (params->timeout_ns != 0 && k >= 4)
        {
          // scanner.c:74-81 at slice granularity: at most four slices are queued ahead of the one that has
          // finished, so the clock is looked at while the device still has work -- and when time is up the
          // host stops feeding it
          CK(cudaEventSynchronize(s->slice_done[k - 4]));
          const uint64_t elapsed = (uint64_t) std::chrono::duration_cast<std::chrono::nanoseconds>(
                                       std::chrono::steady_clock::now() - t_begin).count();
          if (elapsed > params->timeout_ns)
          {
            cudaStreamSynchronize(s->copy_stream);
            cudaStreamSynchronize(s->stream);
            return YRB_ERROR_SCAN_TIMEOUT;
          }
        }
        const size_t o = k * chunk;
        const size_t nb = (o + chunk <= size) ? chunk : size - o;
        CK(cudaMemcpyAsync(s->d_data.p + o, h_data + o, nb, cudaMemcpyHostToDevice, s->copy_stream));
        CK(cudaEventRecord(s->slice_ev[k], s->copy_stream));
        CK(cudaStreamWaitEvent(s->stream, s->slice_ev[k], 0));
        // local bytes [o, o+nb) are block bytes [g0+o, g0+o+nb)
        int64_t slo = V.g0 + (int64_t) o, shi = V.g0 + (int64_t) (o + nb);
        if (slo < lo) slo = lo;
        if (shi > hi) shi = hi;
        if (shi > slo) { launch_scan_slice(s, kernel, V, co, slo, shi); launches++; }
        if (params->timeout_ns != 0) CK(cudaEventRecord(s->slice_done[k], s->stream));
      }
    }
    else if (hi > lo)
    {
      launch_scan_slice(s, kernel, V, co, lo, hi);
      launches++;
    }

    CK(cudaEventRecord(s->ev[1], s->stream));
    int rc = read_status(s);
    if (rc != YRB_SUCCESS) return rc;

    bool overflow = false;
    for (int c = 0; c < NUM_CLS; c++)
    {
      if (hs->n_cand[c] > s->d_cand[c].cap)
      {
        overflow = true;
        if (s->d_cand[c].ensure((size_t) hs->n_cand[c] + 1024) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
      }
    }
    if (!overflow) break;
    if (attempt > 3) return YRB_ERROR_INTERNAL_FATAL;
  }

  if (hs->margin_fault)
    return YRB_ERROR_MARGIN;

  unsigned long long n_cand[NUM_CLS];
  unsigned long long total_cand = 0;
  for (int c = 0; c < NUM_CLS; c++) { n_cand[c] = hs->n_cand[c]; total_cand += n_cand[c]; }
  res->stats.n_candidates = total_cand;

  if (total_cand == 0 && !params->defer_resolve)
  {
    // nothing to verify (the common case for small benign blocks): no further launches or syncs
    s->last_n_final = 0;
    s->last_view = V;
    res->matches = s->h_finals.p;
    res->unconfirmed = s->h_unconf.p;
    res->too_many = s->h_too_many.p;
    res->stats.kernel_launches = launches;
    cudaEventElapsedTime(&res->stats.ms_scan, s->ev[0], s->ev[1]);
    res->stats.ms_total = res->stats.ms_scan;
    res->stats.ms_h2d = h_data ? res->stats.ms_scan : 0.f;
    return YRB_SUCCESS;
  }

  // ---- verification work areas: tier 1 sized by the candidate count, tier 2 fixed
  if (n_cand[CLS_FAST] > 0)
  {
    size_t want = (size_t) ((n_cand[CLS_FAST] + 127) / 128) * 128;
    const size_t maxt = (size_t) s->rules->num_sms * 2048;
    if (want > maxt) want = maxt;
    if (s->ws_fast1.ensure(want * fast_ws_bytes_per_thread(1)) != YRB_SUCCESS ||
        s->retry_fast.ensure((size_t) n_cand[CLS_FAST]) != YRB_SUCCESS ||
        s->retry_fast3.ensure((size_t) n_cand[CLS_FAST]) != YRB_SUCCESS)
      return YRB_ERROR_INSUFFICIENT_MEMORY;
    s->vws.retry_fast3 = s->retry_fast3.p;
    s->vws.fast1 = s->ws_fast1.p; s->vws.fast1_threads = s->ws_fast1.cap / fast_ws_bytes_per_thread(1) / 128 * 128;
    s->vws.retry_fast = s->retry_fast.p;
  }
  if (n_cand[CLS_RE] > 0)
  {
    const uint32_t sw = s->rules->R.stack_words;
    // tier 0 (fiber pool in local memory) takes every candidate when the repeat stacks fit; tier 1 then only sees
    // what outgrew it and gets by with far fewer threads (and a far smaller work area)
    static const bool tier0_on = !(getenv("YRB_RE_TIER0") && atoi(getenv("YRB_RE_TIER0")) == 0);
    const bool tier0 = tier0_on && sw <= RE_LOCAL_STACK;
    size_t want = (size_t) ((n_cand[CLS_RE] + 127) / 128) * 128;
    const size_t maxt = (size_t) s->rules->num_sms * (tier0 ? 256 : 2048);
    if (want > maxt) want = maxt;
    if (s->ws_re1.ensure(want * re_ws_bytes_per_thread(sw, 1)) != YRB_SUCCESS ||
        s->retry_re.ensure((size_t) n_cand[CLS_RE]) != YRB_SUCCESS ||
        (tier0 && s->retry_re0.ensure((size_t) n_cand[CLS_RE]) != YRB_SUCCESS))
      return YRB_ERROR_INSUFFICIENT_MEMORY;
    s->vws.re1 = s->ws_re1.p; s->vws.re1_threads = want;
    s->vws.retry_re = s->retry_re.p;
    s->vws.retry_re0 = tier0 ? s->retry_re0.p : nullptr;
  }

  if (s->d_raw.ensure((size_t) total_cand * 2 + 4096) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;

  {
    // the VM classes are verified in AC-match order when there are enough candidates for warps to share strings
    static const int sort_mask = getenv("YRB_SORT_CANDS") ? atoi(getenv("YRB_SORT_CANDS")) : (1 << CLS_RE);   // the fast-hex class does not repay its sort (profiles/r02_kernel_notes.txt)
    size_t need = 0, nmax = 0;
    for (int c = 0; c < NUM_CLS; c++)
    {
      s->vws.sort_cls[c] = ((sort_mask >> c) & 1) && c != CLS_LITERAL && n_cand[c] >= 2048 && n_cand[c] < (1ull << 31);
      if (s->vws.sort_cls[c]) { need += (size_t) n_cand[c]; if (n_cand[c] > nmax) nmax = (size_t) n_cand[c]; }
    }
    if (need > 0)
    {
      const size_t tb = (cand_sort_temp_bytes((uint32_t) nmax) + 255) & ~(size_t) 255;
      if (s->cs_k0.ensure(need) != YRB_SUCCESS || s->cs_k1.ensure(need) != YRB_SUCCESS || s->cs_v0.ensure(need) != YRB_SUCCESS ||
          s->cs_v1.ensure(need) != YRB_SUCCESS || s->cs_tmp.ensure(2 * tb) != YRB_SUCCESS)
        return YRB_ERROR_INSUFFICIENT_MEMORY;
      s->vws.sort_k[0] = s->cs_k0.p; s->vws.sort_k[1] = s->cs_k1.p;
      s->vws.sort_v[0] = s->cs_v0.p; s->vws.sort_v[1] = s->cs_v1.p;
      s->vws.sort_cap = need;
      s->vws.sort_tmp = s->cs_tmp.p;
      s->vws.sort_tmp_bytes = 2 * tb;
    }
    else
      s->vws.sort_cap = 0;
  }

  for (int attempt = 0;; attempt++)
  {
    CandOut co;
    for (int c = 0; c < NUM_CLS; c++) { co.cand[c] = s->d_cand[c].p; co.cap[c] = s->d_cand[c].cap; }
    co.st = s->d_status;
    RawOut ro;
    ro.raw = s->d_raw.p;
    ro.cap = s->d_raw.cap;
    ro.st = s->d_status;
    ro.str_err = s->str_err.p;
    ro.item_dirty = bp ? bp->d_dirty : nullptr;
    ro.chain_marks_dirty = (bp && bp->blocks_of_one_scan) ? 1u : 0u;
    if (s->rules->h.code_size > 0) // only RE / hex strings can raise verifier errors
      CK(cudaMemsetAsync(s->str_err.p, 0xFF, sizeof(unsigned long long) * s->rules->R.num_strings, s->stream));

    TRACE("verify tier 1: %llu literal, %llu fast, %llu re candidates", n_cand[0], n_cand[1], n_cand[2]);
    launch_verify(s->rules->R, V, co, n_cand, ro, s->vws, s->stream, &launches, s->use_side ? &s->side : nullptr);
    int rc = read_status(s);
    if (rc != YRB_SUCCESS) return rc;
    TRACE("tier 1 done: retry fast %u re %u, raw %llu, any_error %u", hs->n_retry[0], hs->n_retry[1], hs->n_raw, hs->any_error);

    if (hs->n_retry[0] > 0 || hs->n_retry[1] > 0)
    {
      // tier 2: the reference's full limits, only for what outgrew tier 1, only as many threads as needed
      const size_t maxt2 = (size_t) s->rules->num_sms * 128;
      if (hs->n_retry[0] > 0)
      {
        size_t t2 = ((size_t) hs->n_retry[0] + 127) / 128 * 128;
        if (t2 > maxt2) t2 = maxt2;
        if (s->ws_fast2.ensure(t2 * fast_ws_bytes_per_thread(2)) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
        s->vws.fast2 = s->ws_fast2.p; s->vws.fast2_threads = t2;
      }
      if (hs->n_retry[1] > 0)
      {
        size_t t2 = ((size_t) hs->n_retry[1] + 127) / 128 * 128;
        if (t2 > maxt2) t2 = maxt2;
        if (s->ws_re2.ensure(t2 * re_ws_bytes_per_thread(s->rules->R.stack_words, 2)) != YRB_SUCCESS)
          return YRB_ERROR_INSUFFICIENT_MEMORY;
        s->vws.re2 = s->ws_re2.p; s->vws.re2_threads = t2;
      }
      launch_verify_tier2(s->rules->R, V, co, ro, s->vws, hs->n_retry[0] > 0, hs->n_retry[1] > 0, s->stream, &launches);
      rc = read_status(s);
      if (rc != YRB_SUCCESS) return rc;
      TRACE("tier 2 done: retry3 %u, raw %llu, any_error %u", hs->n_retry3, hs->n_raw, hs->any_error);
    }

    if (hs->n_retry3 > 0)
    {
      // rare: position lists that outgrew tier 2 (backward-mode duplicates).  Tier 3: 64k positions
      // (768 KB) per thread, as many threads as there are candidates up to 2048; tier 4: 1M
      // positions (12 MB) per thread for what is left, 32 threads.
      size_t t3 = hs->n_retry3 < 2048 ? hs->n_retry3 : 2048;
      t3 = (t3 + 31) / 32 * 32;
      if (s->ws_fast3.ensure(t3 * fast_ws_bytes_per_thread(3)) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
      s->vws.fast3 = s->ws_fast3.p;
      s->vws.fast3_threads = t3;
      CK(cudaMemsetAsync(&s->d_status->n_retry[0], 0, sizeof(unsigned int), s->stream));
      launch_verify_fast_tier3(s->rules->R, V, co, ro, s->vws, 3, s->stream, &launches);
      rc = read_status(s);
      if (rc != YRB_SUCCESS) return rc;
      TRACE("tier 3 done: %u candidates left for tier 4", hs->n_retry[0]);
      if (hs->n_retry[0] > 0)
      {
        const size_t t4 = 32;
        if (s->ws_fast3.ensure(t4 * fast_ws_bytes_per_thread(4)) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
        s->vws.fast3 = s->ws_fast3.p;
        s->vws.fast3_threads = t4;
        launch_verify_fast_tier3(s->rules->R, V, co, ro, s->vws, 4, s->stream, &launches);
        rc = read_status(s);
        if (rc != YRB_SUCCESS) return rc;
      }
      // up to 1.5 GB that almost no scan needs: do not keep it in a (possibly pooled) scanner
      if (s->ws_fast3.cap > ((size_t) 256 << 20)) { s->ws_fast3.release(); s->vws.fast3 = nullptr; s->vws.fast3_threads = 0; }
    }
    CK(cudaEventRecord(s->ev[2], s->stream));

    if (hs->n_raw <= s->d_raw.cap) break;
    if (attempt > 2) return YRB_ERROR_INTERNAL_FATAL;
    if (s->d_raw.ensure((size_t) hs->n_raw + 1024) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
    CK(cudaMemsetAsync(&s->d_status->n_raw, 0, sizeof(unsigned long long), s->stream));
    CK(cudaMemsetAsync(&s->d_status->n_retry[0], 0, 2 * sizeof(unsigned int), s->stream));
    CK(cudaMemsetAsync(&s->d_status->n_retry_re0, 0, sizeof(unsigned int), s->stream));
    CK(cudaMemsetAsync(&s->d_status->any_error, 0, sizeof(unsigned int), s->stream));
    CK(cudaMemsetAsync(&s->d_status->n_retry3, 0, 2 * sizeof(unsigned int), s->stream));   // n_retry3, n_chain_raw
  }

  if (hs->key_overflow)
  {
    fprintf(stderr, "yrb200: one candidate produced more than %u match callbacks; the discovery key cannot order them\n",
            KEY_SEQ_MASK(s->rules->R.key_seq_bits) - 1);
    return YRB_ERROR_INTERNAL_FATAL;
  }

  if (hs->margin_fault)
    return YRB_ERROR_MARGIN;

  const unsigned long long n_raw = hs->n_raw;
  if (n_raw >= (1ULL << 31)) return YRB_ERROR_INSUFFICIENT_MEMORY;
  res->stats.n_raw = n_raw;
  s->last_view = V;

  uint32_t n_resolve = (uint32_t) n_raw;
  if (params->n_seeds > 0 && !params->defer_resolve)
  {
    if (params->n_seeds > YRB_MAX_SEEDS || params->seeds == nullptr || s->rules->h.num_chains == 0) return YRB_ERROR_BAD_ARGUMENT;
    if (s->d_seeds.ensure(params->n_seeds) != YRB_SUCCESS || s->seed_base.ensure(params->n_seeds) != YRB_SUCCESS ||
        s->d_raw.ensure((size_t) n_raw + params->n_seeds, true) != YRB_SUCCESS)
      return YRB_ERROR_INSUFFICIENT_MEMORY;
    CK(cudaMemcpyAsync(s->d_seeds.p, params->seeds, sizeof(yrb_seed) * params->n_seeds, cudaMemcpyHostToDevice, s->stream));
    seeds_to_raw_kernel<<<(params->n_seeds + 255) / 256, 256, 0, s->stream>>>(
        s->d_seeds.p, params->n_seeds, s->rules->R, s->d_raw.p + n_raw, s->seed_base.p);
    launches++;
    n_resolve += params->n_seeds;
  }

  int rc = YRB_SUCCESS;
  if (bp)
  {
    // the order-independent part of the resolve only; items that need more are flagged dirty
    if (n_raw > 0)
    {
      rc = ensure_resolve_scratch(s, (uint32_t) n_raw);
      if (rc != YRB_SUCCESS) return rc;
      if (s->h_finals.ensure((size_t) n_raw)) return YRB_ERROR_INSUFFICIENT_MEMORY;
      ResolveWork w;
      fill_work(s, w, s->d_raw.p, (uint32_t) n_raw, 0);
      w.has_chains = hs->n_chain_raw > 0;
      w.n_seg = (int) bp->n;
      w.seg_start = bp->d_start;
      w.seg_len = bp->d_len;
      w.item_dirty = bp->d_dirty;
      launch_resolve(w, s->stream, &launches);
      CK(cudaMemcpyAsync(s->h_finals.p, s->finals.p, (size_t) n_raw * sizeof(yrb_match), cudaMemcpyDeviceToHost, s->stream));
      rc = read_status(s);
      if (rc != YRB_SUCCESS) return rc;
      res->n_matches = (size_t) hs->n_sel + (size_t) hs->n_chain_out;
      res->matches = s->h_finals.p;
    }
    CK(cudaMemcpyAsync(s->h_item_dirty.p, bp->d_dirty, bp->n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaEventRecord(s->ev[3], s->stream));
    CK(cudaStreamSynchronize(s->stream));
  }
  else if (params->defer_resolve && s->shard_of_multi)
  {
    // one device's share of a yrb_multi scan: the records stay in d_raw, the per-string error keys in
    // str_err; yrb_multi brings both to the resolving device with peer copies
    s->h_raw_n = (size_t) n_raw;
    s->shard_n_chain_raw = hs->n_chain_raw;
    s->shard_any_error = hs->any_error;
    CK(cudaEventRecord(s->ev[3], s->stream));
    CK(cudaStreamSynchronize(s->stream));
  }
  else if (params->defer_resolve)
  {
    if (s->h_raw.ensure(n_raw ? (size_t) n_raw : 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
    if (n_raw) CK(cudaMemcpyAsync(s->h_raw.p, s->d_raw.p, (size_t) n_raw * sizeof(Raw), cudaMemcpyDeviceToHost, s->stream));
    s->h_raw_n = (size_t) n_raw;
    if (hs->any_error)
    {
      // shard: report the earliest failing candidate; the global guards are not known here
      ResolveWork w;
      fill_work(s, w, s->d_raw.p, (uint32_t) n_raw, 0);
      launch_error_select(w, s->str_err.p, false, s->stream);
      int rc2 = read_status(s);
      if (rc2 != YRB_SUCCESS) return rc2;
      if (hs->err_key_inv != 0ULL)
      {
        res->error = (int32_t) (~hs->err_key_inv & KEY_SEQ_MASK(s->rules->R.key_seq_bits));
        res->error_string = hs->err_string1 ? hs->err_string1 - 1 : YRB_NONE;
      }
    }
    CK(cudaEventRecord(s->ev[3], s->stream));
    CK(cudaStreamSynchronize(s->stream));
  }
  else
  {
    TRACE("resolve: %u records, chain records %u", n_resolve, hs->n_chain_raw);
    rc = run_resolve(s, s->d_raw.p, n_resolve, params, res, &launches, hs->n_chain_raw > 0 || n_resolve > (uint32_t) n_raw,
                     (int64_t) hs->n_chain_raw + (int64_t) (n_resolve - (uint32_t) n_raw));
    TRACE("resolve done: rc %d error %d", rc, res->error);
    res->seeds_consumed = (n_resolve > (uint32_t) n_raw) ? 1 : 0;
    CK(cudaEventRecord(s->ev[3], s->stream));
    CK(cudaStreamSynchronize(s->stream));
  }

  res->stats.kernel_launches = launches;
  cudaEventElapsedTime(&res->stats.ms_scan, s->ev[0], s->ev[1]);
  cudaEventElapsedTime(&res->stats.ms_verify, s->ev[1], s->ev[2]);
  cudaEventElapsedTime(&res->stats.ms_resolve, s->ev[2], s->ev[3]);
  cudaEventElapsedTime(&res->stats.ms_total, s->ev[0], s->ev[3]);
  res->stats.ms_h2d = h_data ? res->stats.ms_scan : 0.f;
  if (s->kernel_timed && cudaEventElapsedTime(&res->stats.ms_kernel, s->ev[6], s->ev[7]) != cudaSuccess)
  { cudaGetLastError(); res->stats.ms_kernel = 0.f; }
  return rc;
}

extern "C" int yrb_scan_host(
    yrb_scanner* s, const uint8_t* data, size_t size, const yrb_scan_params* params, yrb_scan_result* result)
{
  static const uint8_t dummy = 0;
  return scan_impl(s, data ? data : &dummy, nullptr, size, params, result);
}

extern "C" int yrb_scan_device(
    yrb_scanner* s, const uint8_t* d_data, size_t size, const yrb_scan_params* params, yrb_scan_result* result)
{
  static const uint8_t dummy = 0;
  return scan_impl(s, nullptr, d_data ? d_data : &dummy, size, params, result);
}

extern "C" int yrb_scan_batch(
    yrb_scanner* s, const yrb_batch_item* items, uint32_t n_items, int scan_flags, int batch_flags,
    yrb_batch_result* out)
{
  if (!s || !out || (n_items > 0 && !items) || n_items >= (1u << 30)) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(s->device));
  memset(out, 0, sizeof(*out));
  out->n_items = n_items;

  s->b_state.assign(n_items ? n_items : 1, YRB_ITEM_DONE);
  s->b_first.assign((size_t) n_items + 1, 0);
  s->b_matches.clear();
  out->item_state = s->b_state.data();
  out->item_first = s->b_first.data();
  out->matches = s->b_matches.data();
  if (n_items == 0) return YRB_SUCCESS;

  // results that depend on order or on per-string scan state are the per-item path's business
  if ((scan_flags & YRB_SCAN_FAST_MODE) || s->rules->F.root_match)
  {
    std::fill(s->b_state.begin(), s->b_state.end(), (uint8_t) YRB_ITEM_RESCAN);
    out->n_rescan = n_items;
    return YRB_SUCCESS;
  }

  if (s->h_seg_tab.ensure((size_t) 2 * n_items) || s->h_item_dirty.ensure(n_items) ||
      s->seg_tab.ensure((size_t) 2 * n_items) || s->item_dirty.ensure(n_items))
    return YRB_ERROR_INSUFFICIENT_MEMORY;

  int64_t* start = s->h_seg_tab.p;
  int64_t* len = s->h_seg_tab.p + n_items;
  bool one_alloc = (batch_flags & YRB_BATCH_ONE_ALLOCATION) != 0;
  const uint8_t* base = nullptr;
  int64_t total = 0;

  if (one_alloc)
  {
    // keep the host layout: start[k] = data[k] - data of the first non-empty item
    int64_t used = 0;
    for (uint32_t k = 0; k < n_items; k++)
    {
      len[k] = (int64_t) items[k].size;
      if (len[k] == 0) { start[k] = total; continue; }
      if (!items[k].data) return YRB_ERROR_BAD_ARGUMENT;
      if (!base) base = items[k].data;
      const int64_t o = (int64_t) (items[k].data - base);
      if (o < total) { one_alloc = false; break; }   // not ascending / overlapping: pack instead
      start[k] = o;
      total = o + len[k];
      used += len[k];
    }
    if (one_alloc && total > used + used / 4 + 4096) one_alloc = false;
  }
  if (!one_alloc)
  {
    total = 0;
    for (uint32_t k = 0; k < n_items; k++)
    {
      if (items[k].size && !items[k].data) return YRB_ERROR_BAD_ARGUMENT;
      start[k] = total;
      len[k] = (int64_t) items[k].size;
      total += (len[k] + 15) & ~(int64_t) 15;
    }
  }
  if (total >= (1LL << 38)) return YRB_ERROR_BAD_ARGUMENT;

  CK(cudaMemcpyAsync(s->seg_tab.p, s->h_seg_tab.p, sizeof(int64_t) * 2 * n_items, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemsetAsync(s->item_dirty.p, 0, n_items, s->stream));
  memset(s->h_item_dirty.p, 0, n_items);

  BatchPlan bp;
  bp.items = items;
  bp.n = n_items;
  bp.start = start;
  bp.one_allocation = one_alloc;
  bp.d_start = s->seg_tab.p;
  bp.d_len = s->seg_tab.p + n_items;
  bp.d_dirty = s->item_dirty.p;
  bp.blocks_of_one_scan = (batch_flags & YRB_BATCH_BLOCKS_OF_ONE_SCAN) != 0;

  yrb_scan_params params;
  memset(&params, 0, sizeof(params));
  params.scan_flags = scan_flags;
  yrb_scan_result res;
  static const uint8_t dummy = 0;
  int rc = scan_impl(s, one_alloc && base ? base : &dummy, nullptr, (size_t) total, &params, &res, &bp);
  if (rc != YRB_SUCCESS) return rc;
  out->stats = res.stats;

  const ScanStatus* hs = s->h_status;
  if (res.n_matches > 0 && hs->over_cap)
    memset(s->h_item_dirty.p, 1, n_items);   // some string's list would be truncated somewhere: sort it out per item

  // distribute the string-major records over their items (stable, so each item stays string-major)
  std::vector<uint32_t> item_of(res.n_matches);
  for (size_t t = 0; t < res.n_matches; t++)
  {
    const int64_t o = res.matches[t].offset;
    uint32_t k = (uint32_t) (std::upper_bound(start, start + n_items, o) - start) - 1;
    item_of[t] = k;
    if (!s->h_item_dirty.p[k]) s->b_first[(size_t) k + 1]++;
  }
  for (uint32_t k = 0; k < n_items; k++)
  {
    if (s->h_item_dirty.p[k]) { s->b_state[k] = YRB_ITEM_RESCAN; out->n_rescan++; }
    s->b_first[(size_t) k + 1] += s->b_first[k];
  }
  s->b_matches.resize((size_t) s->b_first[n_items]);
  std::vector<uint64_t> cur(s->b_first.begin(), s->b_first.end() - 1);
  for (size_t t = 0; t < res.n_matches; t++)
  {
    const uint32_t k = item_of[t];
    if (s->h_item_dirty.p[k]) continue;
    yrb_match m = res.matches[t];
    m.offset -= start[k];
    s->b_matches[(size_t) cur[k]++] = m;
  }
  out->matches = s->b_matches.data();
  s->last_n_final = 0;
  return YRB_SUCCESS;
}

extern "C" int yrb_shard_raw(yrb_scanner* s, const void** records, size_t* n_records)
{
  if (!s || !records || !n_records) return YRB_ERROR_BAD_ARGUMENT;
  *records = s->h_raw.p;
  *n_records = s->h_raw_n;
  return YRB_SUCCESS;
}

extern "C" int yrb_resolve_raw(
    yrb_scanner* s, const void* records, size_t n_records, const yrb_scan_params* params, yrb_scan_result* res)
{
  if (!s || !params || !res || (n_records && !records) || n_records >= (1ULL << 31)) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(s->device));
  memset(res, 0, sizeof(*res));
  res->error_string = YRB_NONE;
  unsigned launches = 0;

  if (s->d_raw.ensure(n_records ? n_records : 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  CK(cudaEventRecord(s->ev[2], s->stream));
  CK(cudaMemsetAsync(s->d_status, 0, sizeof(ScanStatus), s->stream));
  if (n_records)
    CK(cudaMemcpyAsync(s->d_raw.p, records, n_records * sizeof(Raw), cudaMemcpyHostToDevice, s->stream));
  int rc = read_status(s);
  if (rc != YRB_SUCCESS) return rc;
  rc = run_resolve(s, s->d_raw.p, (uint32_t) n_records, params, res, &launches, true);
  CK(cudaEventRecord(s->ev[3], s->stream));
  CK(cudaStreamSynchronize(s->stream));
  res->stats.n_raw = n_records;
  res->stats.kernel_launches = launches;
  cudaEventElapsedTime(&res->stats.ms_resolve, s->ev[2], s->ev[3]);
  return rc;
}

// ------------------------------------------------- one block over several GPUs
//
// byte-range shards with read margins, no exchange step, no collective.  One
// process drives all devices (the drop-in's yr_rules_scan_mem is one call on one thread): a
// persistent worker thread per device runs that device's H2D + scan + verify; the Raw records
// (tiny compared with the input) are brought to the first device with cudaMemcpyPeerAsync and
// resolved there ONCE -- dedup across atoms, the 1M cap and chained strings are global.
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>

struct MultiJob
{
  const uint8_t* h_data = nullptr;   // host block (whole), or
  const uint8_t* d_shard = nullptr;  // this device's resident shard (points at block byte g0)
  uint64_t g0 = 0, g1 = 0, own0 = 0, own1 = 0, block_size = 0;
  yrb_scan_params params;
  yrb_scan_result res;
  int rc = 0;
};

struct MultiWorker
{
  int device = 0;
  yrb_rules* rules = nullptr;
  yrb_scanner* scanner = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<int> state{0};   // 0 idle, 1 job posted, 2 job done, 3 quit
  std::atomic<int> asleep{0};  // the worker went to sleep on cv (a poster has to notify)
  MultiJob job;
};

// A sharded scan of a few milliseconds is handed to the workers and back through `state`; the wake-up of a sleeping
// thread (tens of microseconds, twice per scan) would be a visible share of it.  Both sides therefore spin on the
// atomic for a short while -- long enough for back-to-back scans -- before they fall back to the condition variable.
#define MULTI_SPIN_NS 300000

template <typename Pred>
static bool spin_until(Pred pred)
{
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  for (int i = 0;; i++)
  {
    if (pred()) return true;
    if ((i & 63) == 63 && std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() > MULTI_SPIN_NS)
      return false;
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

struct yrb_multi
{
  std::vector<MultiWorker*> w;
  uint64_t margin = 0;
  std::mutex scan_mu;   // one scan at a time
  DevBuf<unsigned long long> err_tmp;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

static void multi_worker_main(MultiWorker* w)
{
  cudaSetDevice(w->device);
  for (;;)
  {
    if (!spin_until([w] { const int st = w->state.load(std::memory_order_acquire); return st == 1 || st == 3; }))
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->asleep.store(1);
      w->cv.wait(lk, [w] { const int st = w->state.load(); return st == 1 || st == 3; });
      w->asleep.store(0);
    }
    if (w->state.load() == 3) return;

    MultiJob& j = w->job;
    yrb_scan_params p = j.params;
    p.data_offset = j.g0;
    p.block_size = j.block_size;
    p.own_begin = j.own0;
    p.own_end = j.own1;
    p.defer_resolve = 1;
    p.n_seeds = 0;
    p.seeds = nullptr;
    static const uint8_t dummy = 0;
    const size_t len = (size_t) (j.g1 - j.g0);
    if (j.h_data)
      j.rc = scan_impl(w->scanner, len ? j.h_data + j.g0 : &dummy, nullptr, len, &p, &j.res);
    else
      j.rc = scan_impl(w->scanner, nullptr, j.d_shard ? j.d_shard : &dummy, len, &p, &j.res);

    {
      std::lock_guard<std::mutex> lk(w->mu);   // the waiter checks `state` under this mutex before it sleeps
      w->state.store(2, std::memory_order_release);
    }
    w->cv.notify_all();
  }
}

static uint64_t image_margin(const void* image)
{
  const yrb_image_header* h = (const yrb_image_header*) image;
  const yrb_string* S = yrb_img_strings(image);
  int64_t longest = 0;
  for (uint32_t i = 0; i < h->num_strings; i++)
    if (S[i].length > longest) longest = S[i].length;
  uint64_t m = (uint64_t) (longest * 2 > YRB_RE_SCAN_LIMIT ? longest * 2 : YRB_RE_SCAN_LIMIT) + 64;
  return (m + 4095) & ~(uint64_t) 4095;
}

extern "C" void yrb_multi_destroy(yrb_multi* m)
{
  if (!m) return;
  for (MultiWorker* w : m->w)
  {
    if (w->th.joinable())
    {
      { std::lock_guard<std::mutex> lk(w->mu); w->state.store(3); }
      w->cv.notify_all();
      w->th.join();
    }
    if (w->scanner) yrb_scanner_destroy(w->scanner);
    if (w->rules) yrb_rules_release(w->rules);
    delete w;
  }
  if (!m->w.empty()) cudaSetDevice(m->w[0]->device);
  m->err_tmp.release();
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  delete m;
}

extern "C" int yrb_multi_create(const void* image, size_t image_size, const int* devices, int n_devices, yrb_multi** out)
{
  // (a device may be listed more than once: it then gets several shards, each with its own scanner --
  // no use in production, but it lets one GPU exercise the whole sharded flow)
  if (!image || !out || !devices || n_devices < 1 || n_devices > 64) return YRB_ERROR_BAD_ARGUMENT;

  yrb_multi* m = new (std::nothrow) yrb_multi();
  if (m) m->err_tmp.plain = true;   // destination of peer copies (DevBuf::plain)
  if (!m) return YRB_ERROR_INSUFFICIENT_MEMORY;

  for (int g = 0; g < n_devices; g++)
  {
    MultiWorker* w = new (std::nothrow) MultiWorker();
    if (!w) { yrb_multi_destroy(m); return YRB_ERROR_INSUFFICIENT_MEMORY; }
    m->w.push_back(w);
    w->device = devices[g];
    int rc = yrb_rules_create(image, image_size, devices[g], &w->rules);
    if (rc == YRB_SUCCESS) rc = yrb_scanner_create(w->rules, &w->scanner);
    if (rc != YRB_SUCCESS) { yrb_multi_destroy(m); return rc; }
    w->scanner->shard_of_multi = true;
  }
  m->margin = image_margin(image);

  // records travel device -> first device directly where the topology allows it (NVLink/NVSwitch)
  cudaSetDevice(devices[0]);
  for (int g = 1; g < n_devices; g++)
  {
    int can = 0;
    if (devices[g] != devices[0] && cudaDeviceCanAccessPeer(&can, devices[0], devices[g]) == cudaSuccess && can)
      cudaDeviceEnablePeerAccess(devices[g], 0);
    cudaGetLastError();
  }
  if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess)
  { cudaGetLastError(); yrb_multi_destroy(m); return YRB_ERROR_CUDA; }

  for (MultiWorker* w : m->w) w->th = std::thread(multi_worker_main, w);
  *out = m;
  return YRB_SUCCESS;
}

extern "C" int yrb_multi_device_count(const yrb_multi* m) { return m ? (int) m->w.size() : 0; }
extern "C" uint64_t yrb_multi_margin(const yrb_multi* m) { return m ? m->margin : 0; }

// Shard g owns the events of bytes [own0, own1) and reads [g0, g1): boundaries at multiples of 4096.
extern "C" int yrb_multi_plan(const yrb_multi* m, uint64_t block_size, uint64_t* g0, uint64_t* g1, uint64_t* own0, uint64_t* own1)
{
  if (!m || !g0 || !g1 || !own0 || !own1) return YRB_ERROR_BAD_ARGUMENT;
  const uint64_t n = m->w.size();
  const uint64_t per = ((block_size + n - 1) / n + 4095) & ~(uint64_t) 4095;
  for (uint64_t g = 0; g < n; g++)
  {
    const uint64_t lo = g * per < block_size ? g * per : block_size;
    const uint64_t hi = (g + 1) * per < block_size ? (g + 1) * per : block_size;
    own0[g] = lo;
    own1[g] = hi;
    g0[g] = lo > m->margin ? lo - m->margin : 0;
    g1[g] = hi + m->margin < block_size ? hi + m->margin : block_size;
  }
  return YRB_SUCCESS;
}

__global__ void err_min_merge_kernel(unsigned long long* dst, const unsigned long long* src, uint32_t n)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && src[k] < dst[k]) dst[k] = src[k];
}

static int multi_run(yrb_multi* m, const yrb_scan_params* params, yrb_scan_result* res)
{
  const size_t n = m->w.size();
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();

  for (MultiWorker* w : m->w)
  {
    { std::lock_guard<std::mutex> lk(w->mu); w->state.store(1, std::memory_order_release); }
    if (w->asleep.load()) w->cv.notify_all();
  }
  for (MultiWorker* w : m->w)
  {
    if (!spin_until([w] { return w->state.load(std::memory_order_acquire) == 2; }))
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [w] { return w->state.load() == 2; });
    }
    w->state.store(0);
  }
  std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();

  memset(res, 0, sizeof(*res));
  res->error_string = YRB_NONE;
  size_t total = 0;
  uint32_t chain_raw = 0, any_error = 0;
  int64_t chain_raw_total = 0;
  for (MultiWorker* w : m->w)
  {
    if (w->job.rc != YRB_SUCCESS) return w->job.rc;
    total += w->scanner->h_raw_n;
    chain_raw |= w->scanner->shard_n_chain_raw;
    chain_raw_total += w->scanner->shard_n_chain_raw;
    any_error |= w->scanner->shard_any_error;
    res->stats.bytes += w->job.res.stats.bytes;
    res->stats.n_candidates += w->job.res.stats.n_candidates;
    res->stats.kernel_launches += w->job.res.stats.kernel_launches;
    if (w->job.res.stats.ms_scan > res->stats.ms_scan) res->stats.ms_scan = w->job.res.stats.ms_scan;
    if (w->job.res.stats.ms_verify > res->stats.ms_verify) res->stats.ms_verify = w->job.res.stats.ms_verify;
    if (w->job.res.stats.ms_h2d > res->stats.ms_h2d) res->stats.ms_h2d = w->job.res.stats.ms_h2d;
    if (w->job.res.stats.ms_kernel > res->stats.ms_kernel) res->stats.ms_kernel = w->job.res.stats.ms_kernel;
  }
  res->stats.scan_kernel = m->w[0]->job.res.stats.scan_kernel;
  if (total >= (1ULL << 31)) return YRB_ERROR_INSUFFICIENT_MEMORY;

  // ---- gather on the first device, then ONE global resolve
  yrb_scanner* s0 = m->w[0]->scanner;
  CK(cudaSetDevice(s0->device));
  if (s0->d_raw.ensure(total ? total : 1, true) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  CK(cudaEventRecord(m->ev0, s0->stream));
  size_t at = s0->h_raw_n;
  {
    // the shards' records arrive over several streams at once: seven small peer copies queued on one stream are
    // seven copy latencies one after the other
    cudaStream_t lanes[4] = {s0->stream, s0->copy_stream, s0->side.aux[0], s0->side.aux[1]};
    bool used[4] = {false, false, false, false};
    for (int k = 1; k < 4; k++) CK(cudaStreamWaitEvent(lanes[k], m->ev0, 0));
    for (size_t g = 1; g < n; g++)
    {
      yrb_scanner* sg = m->w[g]->scanner;
      const int k = (int) (g & 3);
      if (sg->h_raw_n)
      {
        CK(cudaMemcpyPeerAsync(s0->d_raw.p + at, s0->device, sg->d_raw.p, sg->device, sg->h_raw_n * sizeof(Raw), lanes[k]));
        used[k] = true;
      }
      at += sg->h_raw_n;
    }
    for (int k = 1; k < 4; k++)
    {
      if (!used[k]) continue;
      CK(cudaEventRecord(s0->side.ev[k - 1], lanes[k]));
      CK(cudaStreamWaitEvent(s0->stream, s0->side.ev[k - 1], 0));
    }
  }
  if (any_error)
  {
    const uint32_t ns = s0->rules->R.num_strings;
    if (m->err_tmp.ensure(ns ? ns : 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
    for (size_t g = 1; g < n; g++)
    {
      yrb_scanner* sg = m->w[g]->scanner;
      if (!sg->shard_any_error) continue;
      CK(cudaMemcpyPeerAsync(m->err_tmp.p, s0->device, sg->str_err.p, sg->device, sizeof(unsigned long long) * ns, s0->stream));
      if (!s0->shard_any_error)
      {
        // the first device saw no error itself: its own table may never have been initialised
        CK(cudaMemcpyAsync(s0->str_err.p, m->err_tmp.p, sizeof(unsigned long long) * ns, cudaMemcpyDeviceToDevice, s0->stream));
        s0->shard_any_error = 1;
      }
      else
        err_min_merge_kernel<<<(ns + 255) / 256, 256, 0, s0->stream>>>(s0->str_err.p, m->err_tmp.p, ns);
    }
    const unsigned int one = 1;
    CK(cudaMemcpyAsync(&s0->d_status->any_error, &one, sizeof(one), cudaMemcpyHostToDevice, s0->stream));
    s0->h_status->any_error = 1;
  }

  unsigned launches = 0;
  yrb_scan_params p = *params;
  p.n_seeds = 0;
  p.seeds = nullptr;
  int rc = run_resolve(s0, s0->d_raw.p, (uint32_t) total, &p, res, &launches, chain_raw != 0, chain_raw_total);
  CK(cudaEventRecord(m->ev1, s0->stream));
  CK(cudaStreamSynchronize(s0->stream));
  std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();

  res->stats.n_raw = total;
  res->stats.kernel_launches += launches;
  cudaEventElapsedTime(&res->stats.ms_resolve, m->ev0, m->ev1);
  res->stats.ms_total = std::chrono::duration<float, std::milli>(t2 - t0).count();
  (void) t1;
  s0->last_n_final = 0;   // yrb_gather_match_data is a single-device facility
  return rc;
}

extern "C" int yrb_multi_scan_host(
    yrb_multi* m, const uint8_t* data, size_t size, const yrb_scan_params* params, yrb_scan_result* result)
{
  if (!m || !params || !result || (size > 0 && !data)) return YRB_ERROR_BAD_ARGUMENT;
  if (params->n_seeds > 0 || params->block_size != 0 || params->defer_resolve) return YRB_ERROR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lk(m->scan_mu);
  const size_t n = m->w.size();
  if (n == 1 || size < n * 4 * (size_t) m->margin)
  {
    // not worth sharding (and an empty block has exactly one i == 0 event): the first device alone
    yrb_scanner* s0 = m->w[0]->scanner;
    s0->shard_of_multi = false;
    static const uint8_t dummy = 0;
    const int rc = scan_impl(s0, data ? data : &dummy, nullptr, size, params, result);
    s0->shard_of_multi = true;
    return rc;
  }
  std::vector<uint64_t> g0(n), g1(n), o0(n), o1(n);
  yrb_multi_plan(m, size, g0.data(), g1.data(), o0.data(), o1.data());
  for (size_t g = 0; g < n; g++)
  {
    MultiJob& j = m->w[g]->job;
    j = MultiJob();
    j.h_data = data;
    j.g0 = g0[g]; j.g1 = g1[g]; j.own0 = o0[g]; j.own1 = o1[g];
    j.block_size = size;
    j.params = *params;
  }
  return multi_run(m, params, result);
}

extern "C" int yrb_multi_scan_device(
    yrb_multi* m, const uint8_t* const* d_shards, const uint64_t* g0, const uint64_t* g1, const uint64_t* own0,
    const uint64_t* own1, uint64_t block_size, const yrb_scan_params* params, yrb_scan_result* result)
{
  if (!m || !params || !result || !d_shards || !g0 || !g1 || !own0 || !own1) return YRB_ERROR_BAD_ARGUMENT;
  if (params->n_seeds > 0 || params->defer_resolve) return YRB_ERROR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lk(m->scan_mu);
  for (size_t g = 0; g < m->w.size(); g++)
  {
    if (g1[g] < g0[g] || g1[g] > block_size || own1[g] < own0[g]) return YRB_ERROR_BAD_ARGUMENT;
    MultiJob& j = m->w[g]->job;
    j = MultiJob();
    j.d_shard = d_shards[g];
    j.g0 = g0[g]; j.g1 = g1[g]; j.own0 = own0[g]; j.own1 = own1[g];
    j.block_size = block_size;
    j.params = *params;
  }
  return multi_run(m, params, result);
}

// ---- shards scanned by DIFFERENT PROCESSES (one process per GPU): the records stay on the devices.
// Every scanner owns an export buffer whose CUDA IPC handle is handed to the resolving process once;
// after each deferred scan the shard's Raw records (and nothing else) are copied into it device-to-device,
// and the resolving process pulls them straight from there -- peer copies over NVLink, no host staging.

struct ShardExport
{
  Raw* d = nullptr;
  size_t cap = 0;
};
static std::mutex g_export_mu;
static std::map<yrb_scanner*, ShardExport> g_exports;

extern "C" int yrb_shard_export_create(yrb_scanner* s, size_t capacity_records, void* ipc_handle_out, size_t handle_bytes)
{
  if (!s || !ipc_handle_out || handle_bytes < sizeof(cudaIpcMemHandle_t) || capacity_records == 0) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(s->device));
  ShardExport e;
  CK(cudaMalloc((void**) &e.d, capacity_records * sizeof(Raw)));
  e.cap = capacity_records;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, e.d) != cudaSuccess) { cudaGetLastError(); cudaFree(e.d); return YRB_ERROR_CUDA; }
  memcpy(ipc_handle_out, &h, sizeof(h));
  s->shard_of_multi = true;
  std::lock_guard<std::mutex> lk(g_export_mu);
  auto it = g_exports.find(s);
  if (it != g_exports.end()) cudaFree(it->second.d);
  g_exports[s] = e;
  return YRB_SUCCESS;
}

extern "C" int yrb_shard_export_ptr(yrb_scanner* s, void** d_ptr)
{
  if (!s || !d_ptr) return YRB_ERROR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lk(g_export_mu);
  auto it = g_exports.find(s);
  if (it == g_exports.end()) return YRB_ERROR_BAD_ARGUMENT;
  *d_ptr = it->second.d;
  return YRB_SUCCESS;
}

extern "C" void yrb_shard_export_destroy(yrb_scanner* s)
{
  std::lock_guard<std::mutex> lk(g_export_mu);
  auto it = g_exports.find(s);
  if (it == g_exports.end()) return;
  cudaSetDevice(s->device);
  cudaFree(it->second.d);
  g_exports.erase(it);
}

// After a defer_resolve scan: the records into the export buffer; *n_records = how many, *any_error = 1 if a
// verifier reported an error in this shard (the caller then falls back to the host path, yrb_shard_raw).
extern "C" int yrb_shard_export_publish(yrb_scanner* s, size_t* n_records, int* any_error)
{
  if (!s || !n_records) return YRB_ERROR_BAD_ARGUMENT;
  ShardExport e;
  {
    std::lock_guard<std::mutex> lk(g_export_mu);
    auto it = g_exports.find(s);
    if (it == g_exports.end()) return YRB_ERROR_BAD_ARGUMENT;
    e = it->second;
  }
  if (s->h_raw_n > e.cap) return YRB_ERROR_INSUFFICIENT_MEMORY;
  CK(cudaSetDevice(s->device));
  if (s->h_raw_n)
    CK(cudaMemcpyAsync(e.d, s->d_raw.p, s->h_raw_n * sizeof(Raw), cudaMemcpyDeviceToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  *n_records = s->h_raw_n;
  if (any_error) *any_error = (int) s->shard_any_error;
  return YRB_SUCCESS;
}

extern "C" int yrb_shard_import_open(int device, const void* ipc_handle, size_t handle_bytes, void** d_ptr)
{
  if (!ipc_handle || !d_ptr || handle_bytes < sizeof(cudaIpcMemHandle_t)) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return YRB_SUCCESS;
}

extern "C" int yrb_shard_import_close(int device, void* d_ptr)
{
  CK(cudaSetDevice(device));
  CK(cudaIpcCloseMemHandle(d_ptr));
  return YRB_SUCCESS;
}

// Records are pulled by a kernel of the RESOLVING device that loads straight through the peer mapping
// (NVLink P2P loads).  A cudaMemcpy from another process's allocation is executed in a context on the
// SOURCE GPU, where it time-slices with that process's own kernels -- measured at N = 2: the copy waited for
// the peer's next 2.6 ms scan kernel (gather + resolve 2.0 ms instead of 0.45).
__global__ void gather_records_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < n16; k += stride)
    dst[k] = src[k];
}

// The global resolve over record arrays that live in device memory (this scanner's own export buffer,
// buffers imported from other processes, or plain device pointers): gathered with device-to-device copies.
extern "C" int yrb_resolve_raw_device(
    yrb_scanner* s, uint32_t n_parts, const void* const* d_parts, const size_t* counts, const yrb_scan_params* params,
    yrb_scan_result* res)
{
  if (!s || !params || !res || (n_parts && (!d_parts || !counts))) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(s->device));
  memset(res, 0, sizeof(*res));
  res->error_string = YRB_NONE;
  size_t total = 0;
  for (uint32_t k = 0; k < n_parts; k++) total += counts[k];
  if (total >= (1ULL << 31)) return YRB_ERROR_BAD_ARGUMENT;
  unsigned launches = 0;

  // (the scanner's own records come from its export buffer, never from d_raw itself)
  if (s->d_raw.ensure(total ? total : 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  CK(cudaEventRecord(s->ev[2], s->stream));
  CK(cudaMemsetAsync(s->d_status, 0, sizeof(ScanStatus), s->stream));
  size_t at = 0;
  for (uint32_t k = 0; k < n_parts; k++)
  {
    if (counts[k])
    {
      const size_t n16 = counts[k] * (sizeof(Raw) / 16);
      const unsigned blocks = (unsigned) std::min<size_t>((n16 + 255) / 256, 148 * 4);
      gather_records_kernel<<<blocks, 256, 0, s->stream>>>(
          reinterpret_cast<uint4*>(s->d_raw.p + at), reinterpret_cast<const uint4*>(d_parts[k]), n16);
      launches++;
    }
    at += counts[k];
  }
  CK(cudaGetLastError());
  memset(s->h_status, 0, sizeof(ScanStatus));
  int rc = run_resolve(s, s->d_raw.p, (uint32_t) total, params, res, &launches, true);
  CK(cudaEventRecord(s->ev[3], s->stream));
  CK(cudaStreamSynchronize(s->stream));
  res->stats.n_raw = total;
  res->stats.kernel_launches = launches;
  cudaEventElapsedTime(&res->stats.ms_resolve, s->ev[2], s->ev[3]);
  return rc;
}

// ------------------------------------------------------------ match data copy

__global__ void match_len_kernel(const yrb_match* m, uint32_t n, uint32_t max_match_data, uint64_t* lens)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t l = m[j].match_length;
  lens[j] = (uint64_t) (l < 0 ? 0 : (l > (int32_t) max_match_data ? (int32_t) max_match_data : l));
}

__global__ void match_copy_kernel(const yrb_match* m, uint32_t n, const uint64_t* offs, ScanView V, uint8_t* out)
{
  // one warp per match
  const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (j >= n) return;
  const uint64_t o0 = offs[j], o1 = offs[j + 1];
  const int64_t src = m[j].offset;
  for (uint64_t k = lane; k < o1 - o0; k += 32)
  {
    const int64_t p = src + (int64_t) k;
    out[o0 + k] = (p >= V.g0 && p < V.g1) ? V.d[p - V.g0] : 0;
  }
}

extern "C" int yrb_gather_match_data(
    yrb_scanner* s, const uint8_t* d_data, uint32_t max_match_data, uint8_t* out, size_t out_cap, uint64_t* offsets)
{
  if (!s || !offsets) return YRB_ERROR_BAD_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const uint32_t n = (uint32_t) s->last_n_final;
  offsets[0] = 0;
  if (n == 0) return YRB_SUCCESS;

  if (s->g_offsets.ensure((size_t) n + 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint64_t*) nullptr, (uint64_t*) nullptr, (int) n + 1);
  if (s->cub_tmp.ensure(tb + 256) != YRB_SUCCESS || s->kF.ensure((size_t) n + 1) != YRB_SUCCESS)
    return YRB_ERROR_INSUFFICIENT_MEMORY;

  CK(cudaMemsetAsync(s->kF.p, 0, ((size_t) n + 1) * sizeof(uint64_t), s->stream));
  match_len_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->finals.p, n, max_match_data, s->kF.p);
  size_t tb2 = s->cub_tmp.cap;
  cub::DeviceScan::ExclusiveSum(s->cub_tmp.p, tb2, s->kF.p, s->g_offsets.p, (int) n + 1, s->stream);
  CK(cudaMemcpyAsync(offsets, s->g_offsets.p, ((size_t) n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));

  const uint64_t total = offsets[n];
  if (total > out_cap) return YRB_ERROR_INSUFFICIENT_MEMORY;
  if (total == 0) return YRB_SUCCESS;
  if (s->g_data.ensure((size_t) total) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;

  ScanView V = s->last_view;
  V.d = d_data;
  match_copy_kernel<<<(unsigned) (((uint64_t) n * 32 + 255) / 256), 256, 0, s->stream>>>(s->finals.p, n, s->g_offsets.p, V, s->g_data.p);
  CK(cudaMemcpyAsync(out, s->g_data.p, (size_t) total, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return YRB_SUCCESS;
}

// The match data of the last yrb_scan_host result, gathered on the device from the scanner's resident copy of the
// block and handed over packed, in pinned memory owned by the scanner: for a caller that has to copy
// min(match_length, max_match_data) bytes per match out of a multi-GiB host block (scan.c:723-737 does, so the glue
// must) that turns a cache and TLB miss per match into a sequential read.
extern "C" int yrb_last_match_data(yrb_scanner* s, uint32_t max_match_data, const uint8_t** data, const uint64_t** offsets)
{
  if (!s || !data || !offsets) return YRB_ERROR_BAD_ARGUMENT;
  if (!s->last_scan_resident || s->last_n_final == 0 || s->last_view.n_seg > 0) return YRB_ERROR_BAD_ARGUMENT;
  const size_t n = s->last_n_final;
  if (s->h_g_offsets.ensure(n + 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  // sizes first (the packed total decides the size of the pinned buffer), then the bytes
  CK(cudaSetDevice(s->device));
  if (s->g_offsets.ensure(n + 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint64_t*) nullptr, (uint64_t*) nullptr, (int) n + 1);
  if (s->cub_tmp.ensure(tb + 256) != YRB_SUCCESS || s->kF.ensure(n + 1) != YRB_SUCCESS) return YRB_ERROR_INSUFFICIENT_MEMORY;
  CK(cudaMemsetAsync(s->kF.p, 0, (n + 1) * sizeof(uint64_t), s->stream));
  match_len_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s->stream>>>(s->finals.p, (uint32_t) n, max_match_data, s->kF.p);
  size_t tb2 = s->cub_tmp.cap;
  cub::DeviceScan::ExclusiveSum(s->cub_tmp.p, tb2, s->kF.p, s->g_offsets.p, (int) n + 1, s->stream);
  CK(cudaMemcpyAsync(s->h_g_offsets.p, s->g_offsets.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  const uint64_t total = s->h_g_offsets.p[n];
  if (s->g_data.ensure((size_t) total + 1) != YRB_SUCCESS || s->h_g_data.ensure((size_t) total + 1) != YRB_SUCCESS)
    return YRB_ERROR_INSUFFICIENT_MEMORY;
  if (total > 0)
  {
    ScanView V = s->last_view;
    V.d = s->d_data.p;
    match_copy_kernel<<<(unsigned) (((uint64_t) n * 32 + 255) / 256), 256, 0, s->stream>>>(s->finals.p, (uint32_t) n, s->g_offsets.p, V, s->g_data.p);
    CK(cudaMemcpyAsync(s->h_g_data.p, s->g_data.p, (size_t) total, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
  }
  *data = s->h_g_data.p;
  *offsets = s->h_g_offsets.p;
  return YRB_SUCCESS;
}

// ------------------------------------------------------------------ utilities

extern "C" int yrb_dev_alloc(int device, size_t bytes, void** d_ptr)
{
  if (device < 0 || device >= yrb_device_count()) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  CK(cudaMalloc(d_ptr, bytes ? bytes : 1));
  return YRB_SUCCESS;
}

extern "C" int yrb_dev_free(int device, void* d_ptr)
{
  if (device < 0 || device >= yrb_device_count()) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  CK(cudaFree(d_ptr));
  return YRB_SUCCESS;
}

// byte k of the stream = (splitmix64(seed + (k >> 3)) >> (8 * (k & 7))) & 0xFF  -- yara_b200/synth.py rand_bytes
__global__ void fill_random_kernel(uint8_t* p, size_t bytes, uint64_t seed, uint64_t stream_offset)
{
  const size_t nwords = (bytes + 7) / 8 + 1;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  const uint64_t w0 = stream_offset >> 3;
  const unsigned shift_bytes = (unsigned) (stream_offset & 7);

  for (size_t w = (size_t) blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride)
  {
    uint64_t z = (w0 + w + seed + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    // word w covers stream bytes [8*(w0+w), +8) = local bytes [8*w - shift_bytes, +8)
    const int64_t l0 = (int64_t) (8 * w) - (int64_t) shift_bytes;
    if (shift_bytes == 0 && l0 + 8 <= (int64_t) bytes && ((uintptr_t) (p + l0) & 7) == 0)
    {
      *reinterpret_cast<uint64_t*>(p + l0) = z;
    }
    else
    {
      for (int j = 0; j < 8; j++)
      {
        const int64_t l = l0 + j;
        if (l >= 0 && l < (int64_t) bytes) p[l] = (uint8_t) (z >> (8 * j));
      }
    }
  }
}

extern "C" int yrb_dev_fill_random(int device, void* d_ptr, size_t bytes, uint64_t seed, uint64_t stream_offset)
{
  if (device < 0 || device >= yrb_device_count()) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  if (bytes == 0) return YRB_SUCCESS;
  fill_random_kernel<<<148 * 8, 256>>>((uint8_t*) d_ptr, bytes, seed, stream_offset);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return YRB_SUCCESS;
}

// plants: n byte strings (concatenated in `bytes`, string k = [starts[k], starts[k+1])) written at d_ptr + offsets[k]
__global__ void scatter_kernel(uint8_t* dst, size_t dst_size, uint32_t n, const uint64_t* offsets, const uint64_t* starts, const uint8_t* bytes)
{
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (k >= n) return;
  const uint64_t s0 = starts[k], s1 = starts[k + 1], o = offsets[k];
  for (uint64_t j = lane; j < s1 - s0; j += 32)
    if (o + j < dst_size) dst[o + j] = bytes[s0 + j];
}

extern "C" int yrb_dev_scatter(
    int device, void* d_ptr, size_t d_size, uint32_t n, const uint64_t* offsets, const uint64_t* starts, const uint8_t* bytes)
{
  if (device < 0 || device >= yrb_device_count()) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  if (n == 0) return YRB_SUCCESS;
  uint64_t *d_off = nullptr, *d_st = nullptr;
  uint8_t* d_b = nullptr;
  const size_t nb = (size_t) starts[n];
  CK(cudaMalloc((void**) &d_off, sizeof(uint64_t) * n));
  CK(cudaMalloc((void**) &d_st, sizeof(uint64_t) * ((size_t) n + 1)));
  CK(cudaMalloc((void**) &d_b, nb ? nb : 1));
  CK(cudaMemcpy(d_off, offsets, sizeof(uint64_t) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_st, starts, sizeof(uint64_t) * ((size_t) n + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, bytes, nb, cudaMemcpyHostToDevice));
  scatter_kernel<<<(unsigned) (((uint64_t) n * 32 + 255) / 256), 256>>>((uint8_t*) d_ptr, d_size, n, d_off, d_st, d_b);
  CK(cudaDeviceSynchronize());
  cudaFree(d_off); cudaFree(d_st); cudaFree(d_b);
  return YRB_SUCCESS;
}

extern "C" int yrb_host_alloc(size_t bytes, void** h_ptr)
{
  if (yrb_device_count() < 1) return YRB_ERROR_NO_DEVICE;
  CK(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return YRB_SUCCESS;
}

extern "C" int yrb_host_free(void* h_ptr)
{
  CK(cudaFreeHost(h_ptr));
  return YRB_SUCCESS;
}

extern "C" int yrb_dev_write(int device, void* d_ptr, size_t d_offset, const void* src, size_t bytes)
{
  if (device < 0 || device >= yrb_device_count()) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  CK(cudaMemcpy((uint8_t*) d_ptr + d_offset, src, bytes, cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize()); // pageable source: the DMA may still be in flight when cudaMemcpy returns
  return YRB_SUCCESS;
}

extern "C" int yrb_dev_read(int device, const void* d_ptr, size_t d_offset, void* dst, size_t bytes)
{
  if (device < 0 || device >= yrb_device_count()) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  CK(cudaMemcpy(dst, (const uint8_t*) d_ptr + d_offset, bytes, cudaMemcpyDeviceToHost));
  return YRB_SUCCESS;
}

__global__ void flush_kernel(uint32_t* p, size_t n)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (uint32_t) i;
}

extern "C" int yrb_dev_flush_l2(int device)
{
  static void* buf[16] = {nullptr};
  const size_t bytes = 256u << 20; // > 126 MB L2
  if (device < 0 || device >= yrb_device_count() || device >= 16) return YRB_ERROR_NO_DEVICE;
  CK(cudaSetDevice(device));
  if (!buf[device]) CK(cudaMalloc(&buf[device], bytes));
  flush_kernel<<<148 * 4, 256>>>((uint32_t*) buf[device], bytes / 4);
  CK(cudaDeviceSynchronize());
  return YRB_SUCCESS;
}
