// mixdir_b200 host engine + C ABI (include/mixdir_b200.h).
//
// Drives the coordinate-ascent loop of mixdir_vi() (/root/reference/R/mixdir_vi.R:58-115)
// and mixdir_vi_dp() (/root/reference/R/mixdir_vi_dp.R:66-129) entirely on the GPU(s):
// per iteration  fused E+M pass -> [deterministic partial reduction] -> [NCCL all-reduce of the
// statistics buffer] -> k_iter (parameter/ELBO step, convergence test, next prior).  The host
// queues iterations ahead and polls a mapped status word; once the device sets `stop` the queued
// kernels return at once.  No CPU fallback: every entry point fails with MIXDIR_B200_CUDA_ERROR
// if no usable device is present.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/mixdir_b200.h"
#include "kernels.cuh"
#include "fused.cuh"
#include "units.cuh"
#include "seg.cuh"
#include "seg_ws.cuh"
#include "iter.cuh"
#include "predict.cuh"
#include "predict_units.cuh"
#include "features.cuh"

using namespace mxd;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      char _b[512];                                                                        \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
               __FILE__, __LINE__);                                                        \
      return fail(MIXDIR_B200_CUDA_ERROR, _b);                                             \
    }                                                                                      \
  } while (0)

// every entry point that touches a device restores the caller's current device on return
struct DeviceGuard {
  int dev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      dev = -1;
      cudaGetLastError();
    }
  }
  ~DeviceGuard() {
    if (dev >= 0) cudaSetDevice(dev);
  }
};

#define RET_IF(expr)                   \
  do {                                 \
    int _s = (expr);                   \
    if (_s != MIXDIR_B200_OK) return _s; \
  } while (0)

// ---------------------------------------------------------------------------
// NCCL, resolved at run time (torch's bundled libnccl.so.2 if it is already in
// the process, else the system one) so the library loads on machines without it.
// ---------------------------------------------------------------------------
namespace nccl {
typedef struct ncclComm* comm_t;
struct unique_id { char internal[128]; };
enum { kSum = 0, kMax = 2 };
enum { kInt64 = 4, kDouble = 8 };  // ncclInt64 = 4, ncclFloat64 = 8 (nccl.h)
typedef int (*GetUniqueId_t)(unique_id*);
typedef int (*CommInitRank_t)(comm_t*, int, unique_id, int);
typedef int (*CommInitAll_t)(comm_t*, int, const int*);
typedef int (*CommDestroy_t)(comm_t);
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, comm_t, cudaStream_t);
typedef int (*Group_t)(void);
typedef const char* (*GetErrorString_t)(int);
static void* lib = nullptr;
static GetUniqueId_t GetUniqueId;
static CommInitRank_t CommInitRank;
static CommInitAll_t CommInitAll;
static CommDestroy_t CommDestroy;
static AllReduce_t AllReduce;
static Group_t GroupStart, GroupEnd;
static GetErrorString_t GetErrorString;

static int load() {
  if (lib) return MIXDIR_B200_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(MIXDIR_B200_CUDA_ERROR, std::string("cannot load libnccl.so.2: ") + dlerror());
#define SYM(name)                                                                      \
  name = (name##_t)dlsym(h, "nccl" #name);                                             \
  if (!name) return fail(MIXDIR_B200_CUDA_ERROR, "libnccl lacks nccl" #name);
  SYM(GetUniqueId) SYM(CommInitRank) SYM(CommInitAll) SYM(CommDestroy) SYM(AllReduce)
  SYM(GetErrorString)
#undef SYM
  GroupStart = (Group_t)dlsym(h, "ncclGroupStart");
  GroupEnd = (Group_t)dlsym(h, "ncclGroupEnd");
  if (!GroupStart || !GroupEnd) return fail(MIXDIR_B200_CUDA_ERROR, "libnccl lacks ncclGroup*");
  lib = h;
  return MIXDIR_B200_OK;
}
}  // namespace nccl

#define NCCL_TRY(expr)                                                                    \
  do {                                                                                    \
    int _r = (expr);                                                                      \
    if (_r != 0) {                                                                        \
      char _b[512];                                                                       \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, nccl::GetErrorString(_r),   \
               __FILE__, __LINE__);                                                       \
      return fail(MIXDIR_B200_CUDA_ERROR, _b);                                            \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
// What the fused kernels gather/scatter: one element per feature, or per PAIR of features
// (kernels.cuh "Units").  identity: elements == (padded) features, the raw code matrix and the
// feature tables are used as they are.
struct UnitPlan {
  bool identity = true;
  int U = 0;      // elements per packed row, padding included (multiple of 4 when packed)
  int WPRu = 0;   // 32-bit words per (packed) row
  int NR2 = 0;    // rows of the unit tables
  int n_pairs = 0;
  std::vector<UnitDesc> desc;  // [U]
  std::vector<int> gbase;      // [U] first unit-table row of element e
  std::vector<int2> t2src;     // [NR2] feature-table rows a unit row is the sum of
  std::vector<MargDesc> marg;  // [NR]  unit rows a feature-table row collects
  // engine 4 (k_em_units): shared-memory placement, element e at bank position e mod Q
  int Q = 1, srows = 0;
  std::vector<int> erows;      // [U] table rows of element e (0: padding)
  std::vector<int> slot;       // [NR2] unit-table row -> shared-memory slot (row * Q + position)
  std::vector<int> eoff_rot;   // [4][U] byte offsets, rotated (units.cuh)
};

// engine 5 (k_em_seg): layout of a tile's segment lists (seg.cuh)
struct SegPlan {
  bool fs = false;  // feature split: the two CTAs of a cluster own half of the features each
  bool ws = false;  // warp-specialised kernel (seg_ws.cuh): fs only, 1024 threads per CTA
  bool gs = false;  // global statistics: no statistics table in shared memory, every (tile, feature,
                    // code) sum goes straight to the global table (RED.ADD.U64); what lets K = 64 fit
  int G = 0;        // group padding, entries
  int TB = 0;       // bytes per tile
  // per M-step owner (fs: 2, else 1): block inside a tile's lists, jobs, statistics rows
  int blk_off[2] = {0, 0}, blk_size[2] = {0, 0};
  int job0[2] = {0, 0}, njobs[2] = {0, 0};
  int sc0[2] = {0, 0}, SC[2] = {0, 0};
  std::vector<int4> jobs;         // {idx_off, cnt_off (inside the owner's block), C_j | feature << 8, first statistics row}
  std::vector<int> srow_to_stat;  // shared-memory statistics row -> feature-table row, owner after owner
};

struct Plan {
  int engine = 0;  // 1 generic, 2 fused, 4 units, 5 sorted segments
  int KP = 0;
  int KC = 0, P = 1, RG = 8, W = 0;
  int FPG = 4;  // engine 4: elements per M-step rotation group
  int n_clusters = 0;
  size_t smem = 0;
  int R = 0;  // rows per tile
  UnitPlan up;
  SegPlan sp;
};

// final pass / predict over units (predict_units.cuh): its own pair plan and packed codes --
// no statistics table competes for shared memory here, so more classes fit one CTA than in the fit
struct PredPlan {
  bool valid = false;
  int K = 0, KC = 0, P = 1, W = 16, R = 256;
  size_t smem = 0;
  UnitPlan up;
};

struct Shard {
  int device = 0;
  bool view = false;  // shares the uploaded data of another handle's shard (mixdir_b200_fit_batch)
  cudaStream_t stream = nullptr;
  int64_t row0 = 0, n_rows = 0;
  uint8_t* d_codes = nullptr;
  int* d_ncat = nullptr;
  int* d_rowoff = nullptr;
  int* d_rowoff_pad = nullptr;
  int64_t* d_ragoff = nullptr;
  nccl::comm_t comm = nullptr;
  bool owns_comm = true;
  int sm_count = 0, smem_optin = 0;
  // workspace (sized for ws_K / ws_KP)
  int ws_K = 0, ws_KP = 0, ws_iter = 0, ws_clusters = 0, ws_ctas = 0;
  double *d_T = nullptr, *d_phi = nullptr, *d_logU = nullptr, *d_stats = nullptr;
  double *d_logprior = nullptr, *d_cs_a = nullptr, *d_part = nullptr;
  double *d_trace = nullptr, *d_rag_a = nullptr, *d_rag_b = nullptr, *d_lambda = nullptr;
  double *d_lambda_pad = nullptr, *d_prior_out = nullptr;
  double *d_Spart = nullptr, *d_cspart = nullptr, *d_Hpart = nullptr;
  int *d_underflow = nullptr, *d_warn = nullptr;
  unsigned int* d_ticket = nullptr;
  // k_iter work list: one block per element (feature or feature pair)
  int2* d_elem = nullptr;
  int2* d_erange = nullptr;
  int n_elem = 0;
  std::map<const void*, size_t> smem_attr;  // kernels whose dynamic shared memory opt-in is set on this device
  // units (feature pairs): descriptors + packed codes, kept while the plan stays the same
  std::vector<UnitDesc> unit_sig;  // plan the buffers below were built for
  bool unit_valid = false;
  UnitDesc* d_udesc = nullptr;
  int* d_ugbase = nullptr;
  int2* d_t2src = nullptr;
  MargDesc* d_marg = nullptr;
  int* d_slot = nullptr;           // engine 4: unit-table row -> shared-memory slot
  int* d_eoff = nullptr;           // engine 4: rotated element offsets
  uint8_t* d_ucodes = nullptr;     // [n_rows + kRowPad][WPRu] words (null: identity plan)
  int64_t packed_rows = 0;         // rows of d_ucodes already packed
  unsigned char* d_seg = nullptr;  // engine 5: [n_tiles][TB] segment lists
  int4* d_jobs = nullptr;
  int* d_srow2stat = nullptr;
  int64_t seg_tiles = 0;           // tiles of d_seg already built
  double* d_T2 = nullptr;          // [NR2][KP] (workspace)
  int ws_NR2 = 0;
  // predict over units: plan signature + device copies, kept while the plan stays the same
  std::vector<UnitDesc> pred_sig;
  bool pred_valid = false;
  UnitDesc* d_pdesc = nullptr;
  int2* d_pt2src = nullptr;
  int* d_pslot = nullptr;
  int* d_peoff = nullptr;
  uint8_t* d_pcodes = nullptr;
  double* d_pT2 = nullptr;
  size_t pT2_cap = 0;
  cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr;
  PriorState* d_prior = nullptr;
  IterStatus* d_status = nullptr;
  IterStatus* h_status = nullptr;  // pinned + mapped: k_iter writes it, the host polls it
  cudaEvent_t ev_loop0 = nullptr, ev_loop1 = nullptr;
  std::vector<cudaEvent_t> ev_em;
  std::vector<cudaEvent_t> ev_ring;  // completion of the iterations in flight
  // asynchronous upload: chunks are copied on copy_stream, each followed by an event; the
  // compute stream waits for (and scans) a chunk only when it first needs its rows
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> up_ev;
  std::vector<int64_t> up_end;  // exclusive row end of chunk c
  size_t up_seen = 0;           // chunks already waited for + scanned on the compute stream
  ScanOut* d_scan = nullptr;    // NA count / largest code / validity, accumulated over chunks
  unsigned int* d_na_feat = nullptr;  // [J] NA cells per feature (null: J too large to count)
  // page-locked, device-mapped scratch for the small per-fit inputs: kernels pull them over
  // PCIe themselves, so they do not queue behind the bulk upload on the H2D copy engine
  unsigned char* h_pin = nullptr;
  size_t h_pin_bytes = 0, h_pin_used = 0;
};

struct mixdir_b200_handle {
  int J = 0, CB = 1, row_bytes = 0, WPR = 0, NR = 0;
  int64_t n_rows = 0;        // rows held by this handle (all shards)
  int64_t n_missing = 0;     // global
  int max_code = 0;          // global
  int64_t sum_ncat = 0;
  std::vector<int> ncat, rowoff, rowoff_pad;
  std::vector<int64_t> ragoff1;  // ragged offsets for K = 1 (scaled by K at use)
  std::vector<uint8_t> has_na;   // [J] 0: the feature holds no NA anywhere (valid once scan_resolved)
  std::vector<Shard> shards;
  int rank = 0, world = 1;   // multi-process flavour
  int64_t n_rows_global = 0; // rows of all ranks together (multi-process flavour; else = n_rows)
  int64_t row_offset = 0;    // global number of this rank's first row (seeds k_em_seg's rounding hash:
                             // the same rows get the same rounding whatever the number of ranks)
  int engine_pref = 0;
  int pairing = 1;           // 1: pair features into units where shared memory allows (default)
  bool scan_resolved = false;
  bool bad = false;  // the upload scan found an invalid code: every later call reports it
  mixdir_b200_timing timing = {};
  bool time_em = true;  // CUDA events around every E+M pass (em_kernel_ms); off inside a concurrent batch
  // mixdir_b200_fit_batch: handles that share this one's uploaded data, each with its own stream and
  // workspace, one per concurrent fit (kept for later batches, freed with the handle)
  std::vector<mixdir_b200_handle*> views;
};

static bool needs_allreduce(const mixdir_b200_handle* h) {
  return h->world > 1 || h->shards.size() > 1;
}

// all-reduce `count` doubles/int64 in place on every shard's stream
static int allreduce(mixdir_b200_handle* h, std::vector<void*>& bufs, size_t count, int dtype,
                     int op) {
  if (!needs_allreduce(h)) return MIXDIR_B200_OK;
  NCCL_TRY(nccl::GroupStart());
  for (size_t s = 0; s < h->shards.size(); s++) {
    Shard& sh = h->shards[s];
    int r = nccl::AllReduce(bufs[s], bufs[s], count, dtype, op, sh.comm, sh.stream);
    if (r != 0) {
      nccl::GroupEnd();
      return fail(MIXDIR_B200_CUDA_ERROR, std::string("ncclAllReduce: ") + nccl::GetErrorString(r));
    }
  }
  NCCL_TRY(nccl::GroupEnd());
  return MIXDIR_B200_OK;
}

struct Plan;
static int allreduce_stats(mixdir_b200_handle* h, std::vector<void*>& bufs, size_t tab, const Plan& pl);

// ---------------------------------------------------------------------------
// block cache: device and page-locked allocations are recycled across handles (cudaMalloc,
// cudaFree and page-locking cost milliseconds; a host that re-creates handles, e.g. per
// repetition or per batch, should not pay them every time).  mixdir_b200_release_cache() gives
// everything back.
// ---------------------------------------------------------------------------
namespace cache {
static std::mutex mu;
static std::multimap<std::pair<int, size_t>, void*> dev_free;          // (device, bytes) -> block
static std::unordered_map<void*, std::pair<int, size_t>> dev_live;     // block -> (device, bytes)
static std::multimap<size_t, void*> pin_free;
static std::unordered_map<void*, size_t> pin_live;
static size_t dev_cached = 0;
static const size_t kDevCap = (size_t)48 << 30;  // beyond this, blocks go back to the driver

static cudaError_t dev_get(void** p, size_t bytes) {
  bytes = (bytes + 511) & ~(size_t)511;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = dev_free.find({dev, bytes});
    if (it != dev_free.end()) {
      *p = it->second;
      dev_free.erase(it);
      dev_cached -= bytes;
      dev_live[*p] = {dev, bytes};
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {  // make room and retry once
    cudaGetLastError();
    std::vector<void*> drop;
    {
      std::lock_guard<std::mutex> g(mu);
      for (auto& kv : dev_free) drop.push_back(kv.second);
      dev_free.clear();
      dev_cached = 0;
    }
    for (void* q : drop) cudaFree(q);
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(mu);
    dev_live[*p] = {dev, bytes};
  }
  return e;
}
static void dev_put(void* p) {
  std::pair<int, size_t> k;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = dev_live.find(p);
    if (it == dev_live.end()) {
      cudaFree(p);
      return;
    }
    k = it->second;
    dev_live.erase(it);
    if (dev_cached + k.second <= kDevCap) {
      dev_free.insert({k, p});
      dev_cached += k.second;
      return;
    }
  }
  cudaFree(p);
}
static cudaError_t pin_get(void** p, size_t bytes) {
  bytes = (bytes + 4095) & ~(size_t)4095;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = pin_free.lower_bound(bytes);
    if (it != pin_free.end() && it->first <= 2 * bytes + (1 << 16)) {
      *p = it->second;
      pin_live[*p] = it->first;
      pin_free.erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(mu);
    pin_live[*p] = bytes;
  }
  return e;
}
static void pin_put(void* p) {
  std::lock_guard<std::mutex> g(mu);
  auto it = pin_live.find(p);
  if (it == pin_live.end()) return;
  pin_free.insert({it->second, p});
  pin_live.erase(it);
}
static void release_all() {
  std::vector<void*> d, h;
  {
    std::lock_guard<std::mutex> g(mu);
    for (auto& kv : dev_free) d.push_back(kv.second);
    for (auto& kv : pin_free) h.push_back(kv.second);
    dev_free.clear();
    pin_free.clear();
    dev_cached = 0;
  }
  for (void* q : d) cudaFree(q);
  for (void* q : h) cudaFreeHost(q);
}
}  // namespace cache

extern "C" int mixdir_b200_release_cache(void) {
  cache::release_all();
  return MIXDIR_B200_OK;
}

template <typename T>
static int dalloc(T** p, size_t n) {
  CU_TRY(cache::dev_get((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return MIXDIR_B200_OK;
}
template <typename T>
static void dfree(T*& p) {
  if (p) cache::dev_put((void*)p);
  p = nullptr;
}

static void free_workspace(Shard& s) {
  dfree(s.d_T); dfree(s.d_phi); dfree(s.d_logU); dfree(s.d_stats); dfree(s.d_logprior);
  dfree(s.d_cs_a); dfree(s.d_part); dfree(s.d_trace); dfree(s.d_rag_a);
  dfree(s.d_rag_b); dfree(s.d_lambda); dfree(s.d_lambda_pad); dfree(s.d_prior_out);
  dfree(s.d_Spart); dfree(s.d_cspart); dfree(s.d_Hpart); dfree(s.d_ticket);
  dfree(s.d_underflow); dfree(s.d_warn); dfree(s.d_prior); dfree(s.d_status); dfree(s.d_T2);
  s.ws_K = s.ws_KP = s.ws_iter = s.ws_clusters = s.ws_ctas = s.ws_NR2 = 0;
}

static void free_shard(Shard& s) {
  cudaSetDevice(s.device);
  if (s.copy_stream) cudaStreamSynchronize(s.copy_stream);
  if (s.stream) cudaStreamSynchronize(s.stream);
  free_workspace(s);
  if (s.view) {  // the data belongs to the parent handle
    s.d_codes = nullptr; s.d_ncat = nullptr; s.d_rowoff = nullptr; s.d_rowoff_pad = nullptr;
    s.d_scan = nullptr; s.d_na_feat = nullptr;
  }
  dfree(s.d_codes); dfree(s.d_ncat); dfree(s.d_rowoff); dfree(s.d_rowoff_pad);
  dfree(s.d_ragoff); dfree(s.d_scan);
  dfree(s.d_na_feat);
  dfree(s.d_udesc); dfree(s.d_ugbase); dfree(s.d_t2src); dfree(s.d_marg); dfree(s.d_ucodes);
  dfree(s.d_slot); dfree(s.d_eoff); dfree(s.d_elem); dfree(s.d_erange);
  dfree(s.d_seg); dfree(s.d_jobs); dfree(s.d_srow2stat);
  s.unit_valid = false;
  dfree(s.d_pdesc); dfree(s.d_pt2src); dfree(s.d_pslot); dfree(s.d_peoff); dfree(s.d_pcodes); dfree(s.d_pT2);
  s.pred_valid = false;
  s.pT2_cap = 0;
  if (s.ev_p0) cudaEventDestroy(s.ev_p0);
  if (s.ev_p1) cudaEventDestroy(s.ev_p1);
  s.ev_p0 = s.ev_p1 = nullptr;
  if (s.copy_stream) {
    cudaStreamSynchronize(s.copy_stream);  // an upload may still be in flight
    cudaStreamDestroy(s.copy_stream);
  }
  s.copy_stream = nullptr;
  for (auto e : s.up_ev) cudaEventDestroy(e);
  s.up_ev.clear();
  if (s.h_status) cache::pin_put(s.h_status);
  s.h_status = nullptr;
  if (s.h_pin) cache::pin_put(s.h_pin);
  s.h_pin = nullptr;
  for (auto e : s.ev_em) cudaEventDestroy(e);
  s.ev_em.clear();
  for (auto e : s.ev_ring) cudaEventDestroy(e);
  s.ev_ring.clear();
  if (s.ev_loop0) cudaEventDestroy(s.ev_loop0);
  if (s.ev_loop1) cudaEventDestroy(s.ev_loop1);
  s.ev_loop0 = s.ev_loop1 = nullptr;
  if (s.comm && s.owns_comm && nccl::lib) nccl::CommDestroy(s.comm);
  s.comm = nullptr;
  if (s.stream) cudaStreamDestroy(s.stream);
  s.stream = nullptr;
}

// ---------------------------------------------------------------------------
// create
// ---------------------------------------------------------------------------
static int init_geometry(mixdir_b200_handle* h, int code_bytes, int J, const int* ncat) {
  if (code_bytes != 1 && code_bytes != 2) return fail(MIXDIR_B200_BAD_ARGUMENT, "code_bytes must be 1 or 2");
  if (J < 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "n_features must be >= 1");
  if (!ncat) return fail(MIXDIR_B200_BAD_ARGUMENT, "ncat is NULL");
  h->J = J;
  h->CB = code_bytes;
  h->row_bytes = (J * code_bytes + 3) & ~3;
  h->WPR = h->row_bytes / 4;
  h->ncat.assign(ncat, ncat + J);
  h->rowoff.resize(J);
  h->ragoff1.resize(J);
  int nr = 0;
  int64_t ro = 0;
  const int limit = code_bytes == 1 ? 255 : 65535;
  for (int j = 0; j < J; j++) {
    // mixdir() stops on a column without levels (R/mixdir.R:112-115)
    if (ncat[j] < 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "a feature has no categories (ncat[j] < 1)");
    if (ncat[j] > limit) return fail(MIXDIR_B200_BAD_ARGUMENT, "ncat[j] does not fit code_bytes");
    h->rowoff[j] = nr;
    nr += ncat[j] + 1;
    h->ragoff1[j] = ro;
    ro += ncat[j];
  }
  h->NR = nr;
  h->sum_ncat = ro;
  const int fpw = 4 / code_bytes;
  h->rowoff_pad.assign((size_t)h->WPR * fpw, h->rowoff[0]);  // padding features -> NA row of feature 0
  for (int j = 0; j < J; j++) h->rowoff_pad[j] = h->rowoff[j];
  return MIXDIR_B200_OK;
}

constexpr int kMaxNaFeat = 8192;  // per-feature NA counters of the scan live in shared memory

static int init_shard_static(mixdir_b200_handle* h, Shard& s) {
  CU_TRY(cudaSetDevice(s.device));
  int cc_major = 0;  // (cudaGetDeviceProperties costs milliseconds; these attribute reads do not)
  CU_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, s.device));
  if (cc_major < 10)
    return fail(MIXDIR_B200_CUDA_ERROR, "mixdir_b200 needs a Blackwell-class GPU (built for sm_100a)");
  CU_TRY(cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, s.device));
  CU_TRY(cudaDeviceGetAttribute(&s.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s.device));
  CU_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
  RET_IF(dalloc(&s.d_scan, 1));
  CU_TRY(cudaMemsetAsync(s.d_scan, 0, sizeof(ScanOut), s.stream));
  if (h->J <= kMaxNaFeat) {
    RET_IF(dalloc(&s.d_na_feat, (size_t)h->J));
    CU_TRY(cudaMemsetAsync(s.d_na_feat, 0, (size_t)h->J * sizeof(unsigned int), s.stream));
  }
  CU_TRY(cudaEventCreate(&s.ev_loop0));
  CU_TRY(cudaEventCreate(&s.ev_loop1));
  CU_TRY(cache::pin_get((void**)&s.h_status, sizeof(IterStatus)));
  const int J = h->J;
  RET_IF(dalloc(&s.d_ncat, J));
  RET_IF(dalloc(&s.d_rowoff, J));
  RET_IF(dalloc(&s.d_rowoff_pad, h->rowoff_pad.size()));
  RET_IF(dalloc(&s.d_ragoff, J));
  CU_TRY(cudaMemcpy(s.d_ncat, h->ncat.data(), J * sizeof(int), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(s.d_rowoff, h->rowoff.data(), J * sizeof(int), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(s.d_rowoff_pad, h->rowoff_pad.data(), h->rowoff_pad.size() * sizeof(int),
                    cudaMemcpyHostToDevice));
  const size_t bytes = (size_t)(s.n_rows + kRowPad) * h->row_bytes;
  RET_IF(dalloc(&s.d_codes, bytes));
  // Zero what the upload does not write (on the copy stream, i.e. before the copies): the pad
  // rows always, and the whole buffer when rows carry padding bytes (J*CB not a multiple of 4) --
  // a stale byte there would be read as a code of a padding feature.
  if (h->J * h->CB == h->row_bytes)
    CU_TRY(cudaMemsetAsync(s.d_codes + (size_t)s.n_rows * h->row_bytes, 0,
                           (size_t)kRowPad * h->row_bytes, s.copy_stream));
  else
    CU_TRY(cudaMemsetAsync(s.d_codes, 0, bytes, s.copy_stream));
  return MIXDIR_B200_OK;
}

// rows per upload chunk: ~8 chunks, multiples of 4096 rows (16-byte aligned tile starts)
static int64_t upload_chunk_rows(int64_t n_rows) {
  int64_t c = (n_rows + 7) / 8;
  c = (c + 4095) / 4096 * 4096;
  return std::max<int64_t>(c, 4096);
}

static int push_chunk_event(Shard& s, int64_t row_end) {
  cudaEvent_t e;
  CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU_TRY(cudaEventRecord(e, s.copy_stream));
  s.up_ev.push_back(e);
  s.up_end.push_back(row_end);
  return MIXDIR_B200_OK;
}

// Asynchronous upload of packed codes: chunked copies on the copy stream.  Returns without
// waiting; with page-locked source memory the copies overlap the first E+M pass.
static int upload_codes(mixdir_b200_handle* h, Shard& s, const uint8_t* src_rows) {
  if (s.n_rows == 0) return MIXDIR_B200_OK;
  const size_t w = (size_t)h->J * h->CB;
  const int64_t chunk = upload_chunk_rows(s.n_rows);
  for (int64_t r0 = 0; r0 < s.n_rows; r0 += chunk) {
    const int64_t nr = std::min(chunk, s.n_rows - r0);
    uint8_t* dst = s.d_codes + (size_t)r0 * h->row_bytes;
    const uint8_t* src = src_rows + (size_t)r0 * w;
    if ((int)w == h->row_bytes)
      CU_TRY(cudaMemcpyAsync(dst, src, w * nr, cudaMemcpyHostToDevice, s.copy_stream));
    else
      CU_TRY(cudaMemcpy2DAsync(dst, h->row_bytes, src, w, w, nr, cudaMemcpyHostToDevice,
                               s.copy_stream));
    RET_IF(push_chunk_event(s, r0 + nr));
  }
  return MIXDIR_B200_OK;
}

// Same for the reference's fp64 column-major matrix (what mixdir() hands to mixdir_vi(),
// R/mixdir.R:117-118): the codes are packed ON THE HOST, by a few threads, into page-locked staging
// buffers (fp64 column-major -> 1- or 2-byte row-major: 8x fewer bytes cross PCIe than copying the
// doubles, and a pageable 4.8 GB source would otherwise crawl through the driver's bounce buffers),
// each chunk is sent as soon as it is packed.  Invalid cells (not NA and not a positive integer
// code) raise the scan's `bad` flag.  MIXDIR_B200_HOST_THREADS overrides the thread count.
static int host_pack_threads() {
  if (const char* e = getenv("MIXDIR_B200_HOST_THREADS")) {
    const int t = atoi(e);
    if (t >= 1) return std::min(t, 64);
  }
  unsigned hc = std::thread::hardware_concurrency();
  if (!hc) hc = 1;
  // one process per GPU (torchrun): the ranks of a node share its cores
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
    const int lw = atoi(e);
    if (lw > 1) hc = std::max(1u, hc / (unsigned)lw);
  }
  return (int)std::max(1u, std::min(16u, hc));
}

// hostpack.cpp (built by g++): AVX2 packer for 1-byte codes
extern "C" int mxd_have_avx2();
extern "C" int mxd_pack_u8_avx2(const double* X, int64_t ld, int64_t r0, int64_t r1, int J, int row_bytes,
                                unsigned char* dst, int limit);

template <typename CT>
static void pack_block(const double* X, int64_t ld, int64_t r0, int64_t r1, int J, int row_bytes,
                       unsigned char* dst /* row r0 of the chunk buffer */, int limit, int* bad) {
  if (sizeof(CT) == 1 && mxd_have_avx2() && !getenv("MIXDIR_B200_SCALAR_PACK")) {
    if (mxd_pack_u8_avx2(X, ld, r0, r1, J, row_bytes, dst, limit)) __atomic_store_n(bad, 1, __ATOMIC_RELAXED);
    return;
  }
  constexpr int64_t RB = 256;  // rows per block: 256 x J codes stay in L1 while the columns stream
  int b = 0;
  for (int64_t i0 = r0; i0 < r1; i0 += RB) {
    const int64_t n = std::min(RB, r1 - i0);
    for (int j = 0; j < J; j++) {
      const double* col = X + (int64_t)j * ld + i0;
      unsigned char* out = dst + (i0 - r0) * row_bytes + (size_t)j * sizeof(CT);
      for (int64_t i = 0; i < n; i++) {
        const double v = col[i];
        int x = 0;
        if (v == v) {  // not NA / NaN
          x = (int)v;
          if (v < 1.0 || v != (double)x || x > limit) {
            b = 1;
            x = 0;
          }
        }
        *reinterpret_cast<CT*>(out + i * row_bytes) = (CT)x;
      }
    }
  }
  if (b) __atomic_store_n(bad, 1, __ATOMIC_RELAXED);
}

static int upload_r_matrix(mixdir_b200_handle* h, Shard& s, const double* X, int64_t ld) {
  if (s.n_rows == 0) return MIXDIR_B200_OK;
  const int J = h->J;
  const int64_t chunk = upload_chunk_rows(s.n_rows);
  const size_t cbytes = (size_t)chunk * h->row_bytes;
  unsigned char* stage[2] = {nullptr, nullptr};
  cudaEvent_t sent[2] = {nullptr, nullptr};
  int bad = 0, status = MIXDIR_B200_OK;
  auto body = [&]() -> int {
    for (int b = 0; b < 2; b++) {
      CU_TRY(cache::pin_get((void**)&stage[b], cbytes));
      CU_TRY(cudaEventCreateWithFlags(&sent[b], cudaEventDisableTiming));
    }
    const int T = host_pack_threads();
    const int limit = h->CB == 1 ? 255 : 65535;
    int c = 0;
    for (int64_t r0 = 0; r0 < s.n_rows; r0 += chunk, c++) {
      const int64_t nr = std::min(chunk, s.n_rows - r0);
      unsigned char* st = stage[c & 1];
      if (c >= 2) CU_TRY(cudaEventSynchronize(sent[c & 1]));  // its previous copy has left the buffer
      if (h->J * h->CB != h->row_bytes) memset(st, 0, (size_t)nr * h->row_bytes);  // row padding bytes
      std::vector<std::thread> pool;
      const int64_t per = ((nr + T - 1) / T + 255) / 256 * 256;
      for (int t = 0; t < T; t++) {
        const int64_t a = std::min<int64_t>(nr, t * per), b2 = std::min<int64_t>(nr, (t + 1) * per);
        if (a >= b2) break;
        const double* Xs = X + s.row0 + r0;
        unsigned char* d = st + (size_t)a * h->row_bytes;
        if (h->CB == 1)
          pool.emplace_back(pack_block<uint8_t>, Xs, ld, a, b2, J, h->row_bytes, d, limit, &bad);
        else
          pool.emplace_back(pack_block<uint16_t>, Xs, ld, a, b2, J, h->row_bytes, d, limit, &bad);
      }
      for (auto& th : pool) th.join();
      CU_TRY(cudaMemcpyAsync(s.d_codes + (size_t)r0 * h->row_bytes, st, (size_t)nr * h->row_bytes,
                             cudaMemcpyHostToDevice, s.copy_stream));
      CU_TRY(cudaEventRecord(sent[c & 1], s.copy_stream));
      RET_IF(push_chunk_event(s, r0 + nr));
    }
    if (bad) {  // seen by the upload scan's consumers (k_em_*, resolve_scan)
      static const int one = 1;
      CU_TRY(cudaMemcpyAsync(&s.d_scan->bad, &one, sizeof(int), cudaMemcpyHostToDevice, s.copy_stream));
      RET_IF(push_chunk_event(s, s.n_rows));
    }
    CU_TRY(cudaStreamSynchronize(s.copy_stream));  // the staging buffers go back to the cache
    return MIXDIR_B200_OK;
  };
  status = body();
  for (int b = 0; b < 2; b++) {
    if (sent[b]) cudaEventDestroy(sent[b]);
    if (stage[b]) cache::pin_put(stage[b]);
  }
  return status;
}

// make upload chunks [s.up_seen, upto] visible to the compute stream and scan them
// (NA count, largest code = the reference's n_cat, validity)
static int consume_upload(mixdir_b200_handle* h, Shard& s, size_t upto) {
  CU_TRY(cudaSetDevice(s.device));
  for (; s.up_seen <= upto && s.up_seen < s.up_ev.size(); s.up_seen++) {
    const size_t c = s.up_seen;
    CU_TRY(cudaStreamWaitEvent(s.stream, s.up_ev[c], 0));
    const int64_t r0 = c == 0 ? 0 : s.up_end[c - 1];
    const int64_t nr = s.up_end[c] - r0;
    if (nr <= 0) continue;  // a marker behind the last chunk (invalid cells found by the host packer)
    const uint8_t* base = s.d_codes + (size_t)r0 * h->row_bytes;
    const int grid = (int)std::min<int64_t>(s.sm_count * 8, (nr * h->WPR + 255) / 256);
    const size_t sm = s.d_na_feat ? (size_t)h->J * sizeof(unsigned int) : 0;
    if (h->CB == 1)
      k_scan_codes<1><<<grid, 256, sm, s.stream>>>(base, h->row_bytes, nr, h->J, s.d_ncat, s.d_scan, s.d_na_feat);
    else
      k_scan_codes<2><<<grid, 256, sm, s.stream>>>(base, h->row_bytes, nr, h->J, s.d_ncat, s.d_scan, s.d_na_feat);
    CU_TRY(cudaGetLastError());
  }
  return MIXDIR_B200_OK;
}

// wait for the whole upload and bring NA count / largest code / validity to the host
// (collective over ranks in the multi-process flavour)
static int bad_codes() {
  return fail(MIXDIR_B200_BAD_ARGUMENT,
              "a code exceeds its feature's number of categories (or X holds a value that is "
              "not a positive integer code or NA)");
}

static int resolve_scan(mixdir_b200_handle* h) {
  if (h->scan_resolved) return h->bad ? bad_codes() : MIXDIR_B200_OK;
  int64_t miss = 0;
  int mx = 0, bad = 0;
  for (auto& s : h->shards) {
    if (!s.up_ev.empty()) RET_IF(consume_upload(h, s, s.up_ev.size() - 1));
  }
  std::vector<long long> na_any(h->J, 0);  // > 0: the feature holds an NA on some shard / rank
  std::vector<unsigned int> na_cnt(h->J);
  for (auto& s : h->shards) {
    CU_TRY(cudaSetDevice(s.device));
    ScanOut o;
    CU_TRY(cudaMemcpyAsync(&o, s.d_scan, sizeof o, cudaMemcpyDeviceToHost, s.stream));
    if (s.d_na_feat)
      CU_TRY(cudaMemcpyAsync(na_cnt.data(), s.d_na_feat, (size_t)h->J * sizeof(unsigned int),
                             cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    for (int j = 0; j < h->J; j++) na_any[j] += s.d_na_feat ? (long long)(na_cnt[j] != 0) : 1;
    miss += (int64_t)o.n_missing;
    mx = std::max(mx, o.max_code);
    bad |= o.bad;
  }
  if (h->world > 1) {  // combine over ranks (multi-process flavour: one shard)
    Shard& s = h->shards[0];
    CU_TRY(cudaSetDevice(s.device));
    long long* d = nullptr;
    RET_IF(dalloc(&d, 3));
    long long v[3] = {(long long)miss, (long long)mx, (long long)bad};
    CU_TRY(cudaMemcpyAsync(d, v, sizeof v, cudaMemcpyHostToDevice, s.stream));
    NCCL_TRY(nccl::AllReduce(d, d, 1, nccl::kInt64, nccl::kSum, s.comm, s.stream));
    NCCL_TRY(nccl::AllReduce(d + 1, d + 1, 2, nccl::kInt64, nccl::kMax, s.comm, s.stream));
    CU_TRY(cudaMemcpyAsync(v, d, sizeof v, cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    dfree(d);
    miss = v[0];
    mx = (int)v[1];
    bad = (int)v[2];
    long long* dn = nullptr;
    RET_IF(dalloc(&dn, (size_t)h->J));
    CU_TRY(cudaMemcpyAsync(dn, na_any.data(), (size_t)h->J * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
    NCCL_TRY(nccl::AllReduce(dn, dn, (size_t)h->J, nccl::kInt64, nccl::kMax, s.comm, s.stream));
    CU_TRY(cudaMemcpyAsync(na_any.data(), dn, (size_t)h->J * sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    dfree(dn);
  }
  h->has_na.assign(h->J, 1);
  for (int j = 0; j < h->J; j++) h->has_na[j] = na_any[j] != 0;
  h->n_missing = miss;
  h->max_code = mx;
  h->scan_resolved = true;
  h->bad = bad != 0;
  return h->bad ? bad_codes() : MIXDIR_B200_OK;
}

static int make_shards(mixdir_b200_handle* h, int64_t n_rows, int n_devices, const int* device_ids) {
  int ndev = n_devices <= 0 ? 1 : n_devices;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(MIXDIR_B200_CUDA_ERROR, "no CUDA device available (mixdir_b200 has no CPU path)");
  if (ndev > count && !device_ids) return fail(MIXDIR_B200_BAD_ARGUMENT, "n_devices exceeds the visible GPUs");
  h->n_rows = n_rows;
  h->shards.resize(ndev);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < ndev; d++) {
    Shard& s = h->shards[d];
    s.device = device_ids ? device_ids[d] : (ndev == 1 ? cur : d);
    s.row0 = n_rows * d / ndev;  // contiguous row blocks
    s.n_rows = n_rows * (d + 1) / ndev - s.row0;
  }
  if (ndev > 1) {
    RET_IF(nccl::load());
    std::vector<nccl::comm_t> comms(ndev);
    std::vector<int> devs(ndev);
    for (int d = 0; d < ndev; d++) devs[d] = h->shards[d].device;
    NCCL_TRY(nccl::CommInitAll(comms.data(), ndev, devs.data()));
    for (int d = 0; d < ndev; d++) h->shards[d].comm = comms[d];
  }
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_destroy(mixdir_b200_handle* h) {
  if (!h) return MIXDIR_B200_OK;
  DeviceGuard guard;
  for (auto* v : h->views) mixdir_b200_destroy(v);
  h->views.clear();
  for (auto& s : h->shards) free_shard(s);
  delete h;
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_create(const void* codes, int code_bytes, int64_t n_rows, int n_features,
                                  const int* ncat, int n_devices, const int* device_ids,
                                  mixdir_b200_handle** out) {
  if (!out) return fail(MIXDIR_B200_BAD_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (n_rows < 0 || (n_rows > 0 && !codes)) return fail(MIXDIR_B200_BAD_ARGUMENT, "codes is NULL");
  DeviceGuard guard;
  mixdir_b200_handle* h = new mixdir_b200_handle();
  int st = init_geometry(h, code_bytes, n_features, ncat);
  if (st == MIXDIR_B200_OK) st = make_shards(h, n_rows, n_devices, device_ids);
  for (size_t i = 0; st == MIXDIR_B200_OK && i < h->shards.size(); i++) {
    Shard& s = h->shards[i];
    st = init_shard_static(h, s);
    if (st == MIXDIR_B200_OK)
      st = upload_codes(h, s, (const uint8_t*)codes + (size_t)s.row0 * n_features * code_bytes);
  }
  if (st != MIXDIR_B200_OK) {
    std::string keep = g_err;
    mixdir_b200_destroy(h);
    g_err = keep;
    return st;
  }
  *out = h;
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_create_from_r_matrix(const double* X, int64_t n_rows, int n_features,
                                                const int* ncat, int n_devices,
                                                const int* device_ids, mixdir_b200_handle** out) {
  if (!out) return fail(MIXDIR_B200_BAD_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (n_rows < 0 || (n_rows > 0 && !X)) return fail(MIXDIR_B200_BAD_ARGUMENT, "X is NULL");
  if (!ncat || n_features < 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "bad ncat / n_features");
  int cb = 1;
  for (int j = 0; j < n_features; j++)
    if (ncat[j] > 255) cb = 2;
  DeviceGuard guard;
  mixdir_b200_handle* h = new mixdir_b200_handle();
  int st = init_geometry(h, cb, n_features, ncat);
  if (st == MIXDIR_B200_OK) st = make_shards(h, n_rows, n_devices, device_ids);
  auto body = [&]() -> int {
    for (auto& s : h->shards) {
      RET_IF(init_shard_static(h, s));
      RET_IF(upload_r_matrix(h, s, X, n_rows));
    }
    return MIXDIR_B200_OK;
  };
  if (st == MIXDIR_B200_OK) st = body();
  if (st != MIXDIR_B200_OK) {
    std::string keep = g_err;
    mixdir_b200_destroy(h);
    g_err = keep;
    return st;
  }
  *out = h;
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_nccl_unique_id(void* out128) {
  if (!out128) return fail(MIXDIR_B200_BAD_ARGUMENT, "out128 is NULL");
  RET_IF(nccl::load());
  nccl::unique_id id;
  NCCL_TRY(nccl::GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return MIXDIR_B200_OK;
}

struct mixdir_b200_comm {
  int device = 0, rank = 0, world = 1;
  nccl::comm_t comm = nullptr;
};

extern "C" int mixdir_b200_comm_init(int device, int rank, int world, const void* unique_id128,
                                     mixdir_b200_comm** out) {
  if (!out) return fail(MIXDIR_B200_BAD_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(MIXDIR_B200_BAD_ARGUMENT, "bad rank/world");
  if (world > 1 && !unique_id128) return fail(MIXDIR_B200_BAD_ARGUMENT, "unique_id128 is NULL");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(MIXDIR_B200_CUDA_ERROR, "no CUDA device available (mixdir_b200 has no CPU path)");
  if (device < 0 || device >= count) return fail(MIXDIR_B200_BAD_ARGUMENT, "bad device");
  DeviceGuard guard;
  mixdir_b200_comm* c = new mixdir_b200_comm();
  c->device = device;
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    int st = nccl::load();
    if (st != MIXDIR_B200_OK) {
      delete c;
      return st;
    }
    cudaSetDevice(device);
    nccl::unique_id id;
    memcpy(&id, unique_id128, sizeof id);
    int r = nccl::CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
      delete c;
      return fail(MIXDIR_B200_CUDA_ERROR, std::string("ncclCommInitRank: ") + nccl::GetErrorString(r));
    }
  }
  *out = c;
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_comm_destroy(mixdir_b200_comm* c) {
  if (!c) return MIXDIR_B200_OK;
  if (c->comm && nccl::lib) nccl::CommDestroy(c->comm);
  delete c;
  return MIXDIR_B200_OK;
}

// multi-process flavour: the global number of this rank's first row = rows of the lower ranks
// (one small all-reduce at create time; every rank creates its handles in the same order anyway)
static int exchange_row_offset(mixdir_b200_handle* h) {
  if (h->world <= 1) return MIXDIR_B200_OK;
  Shard& s = h->shards[0];
  CU_TRY(cudaSetDevice(s.device));
  std::vector<long long> cnt(h->world, 0);
  cnt[h->rank] = (long long)h->n_rows;
  long long* d = nullptr;
  RET_IF(dalloc(&d, (size_t)h->world));
  CU_TRY(cudaMemcpyAsync(d, cnt.data(), cnt.size() * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
  NCCL_TRY(nccl::AllReduce(d, d, (size_t)h->world, nccl::kInt64, nccl::kSum, s.comm, s.stream));
  CU_TRY(cudaMemcpyAsync(cnt.data(), d, cnt.size() * sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
  CU_TRY(cudaStreamSynchronize(s.stream));
  dfree(d);
  h->row_offset = 0;
  h->n_rows_global = 0;
  for (int r = 0; r < h->world; r++) h->n_rows_global += cnt[r];
  for (int r = 0; r < h->rank; r++) h->row_offset += cnt[r];
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_create_dist(const void* codes_local, int code_bytes, int64_t n_rows_local,
                                       int n_features, const int* ncat, mixdir_b200_comm* comm,
                                       mixdir_b200_handle** out) {
  if (!out) return fail(MIXDIR_B200_BAD_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!comm) return fail(MIXDIR_B200_BAD_ARGUMENT, "comm is NULL");
  if (n_rows_local < 0 || (n_rows_local > 0 && !codes_local))
    return fail(MIXDIR_B200_BAD_ARGUMENT, "codes_local is NULL");
  DeviceGuard guard;
  mixdir_b200_handle* h = new mixdir_b200_handle();
  h->rank = comm->rank;
  h->world = comm->world;
  int st = init_geometry(h, code_bytes, n_features, ncat);
  auto body = [&]() -> int {
    h->n_rows = n_rows_local;
    h->shards.resize(1);
    Shard& s = h->shards[0];
    s.device = comm->device;
    s.row0 = 0;
    s.n_rows = n_rows_local;
    s.comm = comm->comm;
    s.owns_comm = false;
    RET_IF(init_shard_static(h, s));
    RET_IF(exchange_row_offset(h));
    return upload_codes(h, s, (const uint8_t*)codes_local);
  };
  if (st == MIXDIR_B200_OK) st = body();
  if (st != MIXDIR_B200_OK) {
    std::string keep = g_err;
    mixdir_b200_destroy(h);
    g_err = keep;
    return st;
  }
  *out = h;
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_create_dist_from_r_matrix(const double* X_local, int64_t n_rows_local,
                                                     int n_features, const int* ncat,
                                                     mixdir_b200_comm* comm, mixdir_b200_handle** out) {
  if (!out) return fail(MIXDIR_B200_BAD_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!comm) return fail(MIXDIR_B200_BAD_ARGUMENT, "comm is NULL");
  if (n_rows_local < 0 || (n_rows_local > 0 && !X_local)) return fail(MIXDIR_B200_BAD_ARGUMENT, "X_local is NULL");
  if (!ncat || n_features < 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "bad ncat / n_features");
  int cb = 1;
  for (int j = 0; j < n_features; j++)
    if (ncat[j] > 255) cb = 2;
  DeviceGuard guard;
  mixdir_b200_handle* h = new mixdir_b200_handle();
  h->rank = comm->rank;
  h->world = comm->world;
  int st = init_geometry(h, cb, n_features, ncat);
  auto body = [&]() -> int {
    h->n_rows = n_rows_local;
    h->shards.resize(1);
    Shard& s = h->shards[0];
    s.device = comm->device;
    s.row0 = 0;
    s.n_rows = n_rows_local;
    s.comm = comm->comm;
    s.owns_comm = false;
    RET_IF(init_shard_static(h, s));
    RET_IF(exchange_row_offset(h));
    return upload_r_matrix(h, s, X_local, n_rows_local);
  };
  if (st == MIXDIR_B200_OK) st = body();
  if (st != MIXDIR_B200_OK) {
    std::string keep = g_err;
    mixdir_b200_destroy(h);
    g_err = keep;
    return st;
  }
  *out = h;
  return MIXDIR_B200_OK;
}

extern "C" int64_t mixdir_b200_n_rows(const mixdir_b200_handle* h) { return h ? h->n_rows : -1; }
extern "C" int mixdir_b200_max_code(const mixdir_b200_handle* h) {
  DeviceGuard guard;
  if (!h || resolve_scan(const_cast<mixdir_b200_handle*>(h)) != MIXDIR_B200_OK) return -1;
  return h->max_code;
}
extern "C" int64_t mixdir_b200_n_missing(const mixdir_b200_handle* h) {
  DeviceGuard guard;
  if (!h || resolve_scan(const_cast<mixdir_b200_handle*>(h)) != MIXDIR_B200_OK) return -1;
  return h->n_missing;
}
extern "C" const char* mixdir_b200_last_error(void) { return g_err.c_str(); }
extern "C" int mixdir_b200_abi_version(void) { return MIXDIR_B200_ABI_VERSION; }
extern "C" size_t mixdir_b200_ragged_size(int n_features, const int* ncat, int n_latent) {
  size_t s = 0;
  for (int j = 0; j < n_features; j++) s += (size_t)ncat[j];
  return s * (size_t)n_latent;
}
extern "C" int mixdir_b200_set_engine(mixdir_b200_handle* h, int engine) {
  if (!h || engine < 0 || engine > 5 || engine == 3) return fail(MIXDIR_B200_BAD_ARGUMENT, "bad engine");
  h->engine_pref = engine;
  return MIXDIR_B200_OK;
}
extern "C" int mixdir_b200_set_pairing(mixdir_b200_handle* h, int on) {
  if (!h) return fail(MIXDIR_B200_BAD_ARGUMENT, "NULL handle");
  h->pairing = on ? 1 : 0;
  return MIXDIR_B200_OK;
}
extern "C" int mixdir_b200_last_timing(const mixdir_b200_handle* h, mixdir_b200_timing* out) {
  if (!h || !out) return fail(MIXDIR_B200_BAD_ARGUMENT, "NULL argument");
  *out = h->timing;
  return MIXDIR_B200_OK;
}

// ---------------------------------------------------------------------------
// planning the fused kernel: classes per CTA (KC), cluster size (P), row groups
// per warp (RG), warps (W) such that table + statistics slices fit shared memory
// ---------------------------------------------------------------------------
typedef void (*fused_fn)(FusedArgs);
typedef void (*units_fn)(UnitArgs);
typedef void (*seg_fn)(SegArgs);
constexpr int kWsRegsE = 96, kWsRegsM = 32;  // k_em_seg_ws: registers per thread of the E-step / M-step warps
static void layout_units(UnitPlan* up, int KC);

static fused_fn pick_fused(int KC, int CB, int RG) {
#define PICK(kc, cb, rg) \
  if (KC == kc && CB == cb && RG == rg) return k_em_fused<kc, cb, rg>;
  PICK(32, 1, 8) PICK(32, 1, 4) PICK(32, 1, 2)
  PICK(16, 1, 8) PICK(16, 1, 4) PICK(16, 1, 2)
  PICK(8, 1, 8) PICK(8, 1, 4) PICK(8, 1, 2)
  PICK(32, 2, 8) PICK(32, 2, 4) PICK(32, 2, 2)
  PICK(16, 2, 8) PICK(16, 2, 4) PICK(16, 2, 2)
#undef PICK
  return nullptr;
}

static units_fn pick_units(int KC, int RG, int FPG, int W);
static seg_fn pick_seg(int KC, int W, bool fs, bool ws = false) {
  if (KC == 16 && fs && ws) return k_em_seg_ws<kWsRegsE, kWsRegsM>;
  if (KC == 16 && fs) return k_em_seg<16, 512, true>;
  if (KC == 16) return k_em_seg<16, 512, false>;
  if (KC == 32) return W > 16 ? k_em_seg<32, 640, false> : k_em_seg<32, 512, false>;
  return nullptr;
}
static int plan_threads(const Plan& p) { return (p.engine == 5 && p.sp.ws) ? 1024 : 32 * p.W; }
static const void* plan_fn(const mixdir_b200_handle* h, const Plan& p) {
  if (p.engine == 5) return (const void*)pick_seg(p.KC, p.W, p.sp.fs, p.sp.ws);
  if (p.engine == 4) return (const void*)pick_units(p.KC, p.RG, p.FPG, p.W);
  return (const void*)pick_fused(p.KC, h->CB, p.RG);
}

// Dynamic shared memory above 48 KB is an opt-in PER DEVICE: set it on the shard's device (which
// must be current) the first time a kernel is planned or launched there.
static cudaError_t ensure_smem_attr(Shard& s, const void* fn, size_t bytes) {
  auto it = s.smem_attr.find(fn);
  if (it != s.smem_attr.end() && it->second >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) s.smem_attr[fn] = bytes;
  return e;
}

// how many clusters of this plan can be resident at once (0: cannot be scheduled);
// the shard's device must be current
static int plan_occupancy(const mixdir_b200_handle* h, Shard& s, const Plan& p) {
  const void* fn = plan_fn(h, p);
  if (!fn) return 0;
  cudaError_t e = ensure_smem_attr(s, fn, p.smem);
  int nclus = 0;
  if (e == cudaSuccess) {
    if (p.P > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(p.P * s.sm_count);
      cfg.blockDim = dim3(plan_threads(p));
      cfg.dynamicSmemBytes = p.smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = p.P;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      e = cudaOccupancyMaxActiveClusters(&nclus, fn, &cfg);
    } else {
      int nb = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, plan_threads(p), p.smem);
      nclus = nb * s.sm_count;
    }
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return nclus;
}

// identity unit plan: one element per (padded) feature, raw codes, feature tables
static void identity_units(const mixdir_b200_handle* h, UnitPlan* up) {
  *up = UnitPlan();
  up->identity = true;
  up->U = (int)h->rowoff_pad.size();
  up->WPRu = h->WPR;
  up->NR2 = h->NR;
  up->gbase = h->rowoff_pad;
  up->desc.assign(up->U, UnitDesc{-1, -1, 1, 0});
  for (int j = 0; j < h->J; j++) up->desc[j] = UnitDesc{j * h->CB, -1, 1, 0};
  up->marg.assign(h->NR, MargDesc{0, 0, 0, 0});
  for (int j = 0; j < h->J; j++)
    for (int c = 1; c <= h->ncat[j]; c++) up->marg[h->rowoff[j] + c] = MargDesc{h->rowoff[j] + c, 0, 1, 0};
}

// Unit plan with m feature pairs (1-byte codes only).  The 2m features with the fewest
// categories are paired smallest-with-largest, which minimises the added table rows
// sum (C_a+1)(C_b+1) for that many pairs; every pair code must still fit a byte.
// Returns false if some pair would need more than 256 codes.
// has_na (nullable, engine 4 only): features with has_na[j] == 0 hold no NA anywhere (upload
// scan) and get no NA row -- their unit-local code is x - 1 -- which shrinks the pair tables.
static bool make_units(const mixdir_b200_handle* h, int m, UnitPlan* up, bool always_packed = false,
                       const uint8_t* has_na = nullptr) {
  const int J = h->J;
  if (m <= 0 && !always_packed) {
    identity_units(h, up);
    return true;
  }
  if (m < 0) m = 0;
  if (h->CB != 1 || 2 * m > J) return false;
  std::vector<int> lev(J), sub(J);  // unit-local codes per feature, and what is subtracted from x
  for (int j = 0; j < J; j++) {
    sub[j] = (has_na && !has_na[j]) ? 1 : 0;
    lev[j] = h->ncat[j] + 1 - sub[j];
  }
  std::vector<int> order(J);
  for (int j = 0; j < J; j++) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lev[a] < lev[b]; });
  std::vector<int> partner(J, -1);
  for (int i = 0; i < m; i++) {
    const int a = order[i], b = order[2 * m - 1 - i];
    if (lev[a] * lev[b] > 256) return false;
    partner[a] = b;
    partner[b] = a;
  }
  *up = UnitPlan();
  up->identity = false;
  up->n_pairs = m;
  up->marg.assign(h->NR, MargDesc{0, 0, 0, 0});
  int nr2 = 0;
  for (int j = 0; j < J; j++) {
    const int b = partner[j];
    if (b >= 0 && b < j) continue;  // emitted with its lower-index partner
    const int Ca = h->ncat[j], La = lev[j], sa = sub[j];
    if (b < 0) {
      up->desc.push_back(UnitDesc{j, -1, 1, sa});
      up->gbase.push_back(nr2);
      up->erows.push_back(La);
      for (int u = 0; u < La; u++) up->t2src.push_back(make_int2(h->rowoff[j] + u + sa, -1));
      for (int c = 1; c <= Ca; c++) up->marg[h->rowoff[j] + c] = MargDesc{nr2 + c - sa, 0, 1, 0};
      nr2 += La;
    } else {
      const int Cb = h->ncat[b], Lb = lev[b], sb = sub[b];
      up->desc.push_back(UnitDesc{j, b, Lb, sa | (sb << 1)});
      up->gbase.push_back(nr2);
      up->erows.push_back(La * Lb);
      for (int ua = 0; ua < La; ua++)
        for (int ub = 0; ub < Lb; ub++)
          up->t2src.push_back(make_int2(h->rowoff[j] + ua + sa, h->rowoff[b] + ub + sb));
      for (int c = 1; c <= Ca; c++) up->marg[h->rowoff[j] + c] = MargDesc{nr2 + (c - sa) * Lb, 1, Lb, 0};
      for (int c = 1; c <= Cb; c++) up->marg[h->rowoff[b] + c] = MargDesc{nr2 + (c - sb), Lb, La, 0};
      nr2 += La * Lb;
    }
  }
  up->NR2 = nr2;
  while (up->desc.size() % 4) {  // padding elements: code 0 -> the all-NA row of element 0 (scratch)
    up->desc.push_back(UnitDesc{-1, -1, 1, 0});
    up->gbase.push_back(up->gbase[0]);
    up->erows.push_back(0);
  }
  up->U = (int)up->desc.size();
  up->WPRu = up->U / 4;
  return true;
}

// host-only view of a unit plan (no device needed): lets the CPU tests check the invariants
// the device code relies on.  Arrays may be NULL (sizes only).
extern "C" int mixdir_b200_unit_plan(int n_features, const int* ncat, int n_pairs,
                                     const unsigned char* has_na, int* n_elements,
                                     int* n_table_rows, int* desc, int* gbase, int* t2src, int* marg) {
  mixdir_b200_handle h;
  RET_IF(init_geometry(&h, 1, n_features, ncat));
  UnitPlan up;
  if (n_pairs < 0 || !make_units(&h, n_pairs, &up, has_na != nullptr, has_na))
    return fail(MIXDIR_B200_BAD_ARGUMENT, "this many pairs cannot be formed (a pair code must fit a byte)");
  if (n_elements) *n_elements = up.U;
  if (n_table_rows) *n_table_rows = up.NR2;
  if (desc) memcpy(desc, up.desc.data(), up.desc.size() * sizeof(UnitDesc));
  if (gbase) memcpy(gbase, up.gbase.data(), up.gbase.size() * sizeof(int));
  if (t2src && !up.identity) memcpy(t2src, up.t2src.data(), up.t2src.size() * sizeof(int2));
  if (marg) memcpy(marg, up.marg.data(), up.marg.size() * sizeof(MargDesc));
  return MIXDIR_B200_OK;
}

// engine 4 placement: element e sits at bank position e mod Q of the shared-memory table rows
// (units.cuh).  Re-orders the elements so that the Q stacks are equally high, row 0 of every
// stack is the zero / scratch row of the padding elements.
static void layout_units(UnitPlan* up, int KC) {
  const int Q = KC >= 16 ? 1 : 16 / KC;
  const int U = up->U, cap = U / Q;
  std::vector<int> real;
  for (int e = 0; e < U; e++)
    if (up->desc[e].off_a >= 0) real.push_back(e);
  std::stable_sort(real.begin(), real.end(), [&](int a, int b) { return up->erows[a] > up->erows[b]; });
  std::vector<int> stack(Q, 1), cnt(Q, 0);
  std::vector<std::vector<std::pair<int, int>>> members(Q);  // (old element, first row in its stack)
  for (int e : real) {
    int q = -1;
    for (int c = 0; c < Q; c++)
      if (cnt[c] < cap && (q < 0 || stack[c] < stack[q])) q = c;
    members[q].push_back({e, stack[q]});
    stack[q] += up->erows[e];
    cnt[q]++;
  }
  std::vector<UnitDesc> desc(U, UnitDesc{-1, -1, 1, 0});
  std::vector<int> gbase(U, 0), erows(U, 0), eoff(U, 0);
  up->slot.assign(up->NR2, 0);
  for (int q = 0; q < Q; q++) {
    for (int i = 0; i < cap; i++) {
      const int e = i * Q + q;
      if (i < (int)members[q].size()) {
        const int old = members[q][i].first, row0 = members[q][i].second;
        desc[e] = up->desc[old];
        gbase[e] = up->gbase[old];
        erows[e] = up->erows[old];
        eoff[e] = (row0 * Q + q) * KC * 8;
        for (int c = 0; c < erows[e]; c++) up->slot[gbase[e] + c] = (row0 + c) * Q + q;
      } else {
        eoff[e] = q * KC * 8;  // padding element: code 0 -> row 0 of its position
      }
    }
  }
  up->desc = desc;
  up->gbase = gbase;
  up->erows = erows;
  up->Q = Q;
  up->srows = *std::max_element(stack.begin(), stack.end());
  up->eoff_rot.assign((size_t)4 * U, 0);
  for (int r = 0; r < 4; r++)
    for (int e = 0; e < U; e++) up->eoff_rot[(size_t)r * U + e] = eoff[(e & ~3) | ((e + r) & 3)];
}

static units_fn pick_units(int KC, int RG, int FPG, int W) {
#define PICK(kc, rg, fpg)                   \
  if (KC == kc && RG == rg && FPG == fpg) \
    return W > 16 ? k_em_units<kc, rg, fpg, 640> : k_em_units<kc, rg, fpg, 512>;
  PICK(32, 8, 2) PICK(32, 4, 2) PICK(32, 2, 2) PICK(32, 8, 4) PICK(32, 4, 4) PICK(32, 2, 4)
  PICK(16, 8, 2) PICK(16, 4, 2) PICK(16, 2, 2) PICK(16, 8, 4) PICK(16, 4, 4) PICK(16, 2, 4)
  PICK(8, 8, 4) PICK(8, 4, 4) PICK(8, 2, 4)
#undef PICK
  return nullptr;
}

// engine 4: k_em_units.  Search (KC, RG, pairs) for the cheapest plan that fits shared memory.
// MIXDIR_B200_FORCE_PLAN="KC,RG,pairs[,FPG[,W]]" (tuning aid, tools/sweep_plans.py) restricts the
// search; -1 leaves a field free.
static bool plan_units(const mixdir_b200_handle* h, Shard& s, int K, Plan* out) {
  if (h->CB != 1) return false;
  int fKC = -1, fRG = -1, fM = -1, fFPG = -1, fW = -1;
  if (const char* f = getenv("MIXDIR_B200_FORCE_PLAN"))
    sscanf(f, "%d,%d,%d,%d,%d", &fKC, &fRG, &fM, &fFPG, &fW);
  // NA rows only for features that hold NAs -- known once the upload scan has been read back
  // (not while the first, chunked iteration overlaps the upload)
  const uint8_t* na = (h->scan_resolved && (int)h->has_na.size() == h->J && !getenv("MIXDIR_B200_KEEP_NA_ROWS"))
                          ? h->has_na.data() : nullptr;
  Plan best;
  double best_cost = 1e300;
  const int kcs[3] = {32, 16, 8};
  const int rgs[3] = {8, 4, 2};
  for (int KC : kcs) {
    if (fKC > 0 && KC != fKC) continue;
    const int RPW = 32 / KC;
    const int P = (K + KC - 1) / KC;
    if (P > 8) continue;
    for (int RG : rgs) {
      if (fRG > 0 && RG != fRG) continue;
      if (RG * RPW > 32) continue;
      UnitPlan up;
      Plan c;
      bool found = false;
      auto shape = [&](UnitPlan& u) -> size_t {
        layout_units(&u, KC);
        c.FPG = RPW > 2 ? 4 : (u.U / 4 >= 16 ? 4 : 2);
        if (fFPG > 0 && fFPG >= RPW) c.FPG = fFPG;
        // warps: at most one per rotation group; 20 when the tiles of 20 warps still fit (more
        // read-modify-writes in flight: 13.58 -> 13.09 ms at S4), else 16
        const int caps[2] = {20, 16};
        for (int cap : caps) {
          c.W = std::min(cap, u.U / c.FPG);
          if (c.W > 8) c.W &= ~3;  // whole warps per scheduler: 17 or 18 warps run slower than 16
          if (fW > 0) c.W = std::min(c.W, fW);
          while (c.W > 1 && ((size_t)c.W * RG * RPW * u.WPRu) % 4 != 0) c.W--;
          if (((size_t)c.W * RG * RPW * u.WPRu) % 4 != 0) return (size_t)-1;
          const size_t tot = UnitSmem(u.srows, KC, c.W, RG, u.WPRu, P).total;
          if (tot <= (size_t)s.smem_optin || cap == 16) return tot;
        }
        return (size_t)-1;
      };
      // Only the number of code words per row (4 elements each) matters: for every word count,
      // fewest first, try the smallest number of pairs that reaches it.
      const int m_hi = fM >= 0 ? std::min(fM, h->J / 2) : (h->pairing ? h->J / 2 : 0);
      const int w_lo = (h->J - m_hi + 3) / 4, w_hi = (h->J + 3) / 4;
      for (int w = w_lo; w <= w_hi; w++) {
        const int m = fM >= 0 ? m_hi : std::max(0, h->J - 4 * w);
        if (m > m_hi || !make_units(h, m, &up, true, na)) continue;
        const size_t tot = shape(up);
        if (tot > (size_t)s.smem_optin) continue;
        c.smem = tot;
        found = true;
        break;
      }
      if (!found) continue;
      c.engine = 4; c.KC = KC; c.P = P; c.RG = RG; c.KP = P * KC;
      c.R = c.W * RG * RPW;
      c.up = up;
      if (!pick_units(KC, RG, c.FPG, c.W)) continue;
      c.n_clusters = plan_occupancy(h, s, c);
      if (c.n_clusters < 1) continue;
      // Cost model, in shared-memory cycles per row summed over the CTAs of a cluster and divided
      // by the fraction of SMs the cluster shape can occupy.  Per CTA and row: 6 wavefronts per
      // element and 256 bytes of classes, plus the work every CTA repeats per row whatever its class
      // slice (code words, soft-max, exchange, barriers; fitted to the plan sweeps in
      // profiles/README.md: ~65 cycles at KC=16, ~52 at KC=8); fewer row groups per lane hide less
      // latency; per-tile barriers are spread over the tile's rows.
      const double used = std::min(1.0, (double)c.n_clusters * P / s.sm_count);
      const int G = up.U / c.FPG;
      const double ovh = KC == 8 ? 52.0 : (KC == 16 ? 65.0 : 80.0);
      const double rgpen = RG == 8 ? 1.0 : (RG == 4 ? 1.08 : 1.30);
      // fewer warps keep fewer shared-memory requests in flight (sweeps: 8 warps x1.26, 12 x1.07,
      // 16 x1.0, 20 x0.965)
      const double wpen = c.W >= 20 ? 0.965 : (c.W >= 16 ? 1.0 : (c.W >= 12 ? 1.07 : (c.W >= 8 ? 1.26 : 1.5)));
      const double per_row = P * (up.U * 6.0 / RPW + ovh) * rgpen * wpen / used;
      const double per_tile = P * (G * 60.0 + 1500.0) / c.R;
      const double cost = per_row + per_tile;
      if (cost < best_cost) {
        best_cost = cost;
        best = c;
      }
    }
  }
  if (best.engine != 4) return false;
  *out = best;
  return true;
}

// the per-iteration exchange step: sum the statistics buffer over shards / ranks.  Engine 5 keeps
// S as int64 fixed point (exact, order-independent sums) and the tail [cs | H | underflow | nNA |
// bad] as doubles.
static int allreduce_stats(mixdir_b200_handle* h, std::vector<void*>& bufs, size_t tab, const Plan& pl) {
  if (!needs_allreduce(h)) return MIXDIR_B200_OK;
  if (pl.engine != 5) return allreduce(h, bufs, tab + pl.KP + 4, nccl::kDouble, nccl::kSum);
  // k_em_seg: the statistics table AND the class masses are 64-bit integers (exact sums: N GPUs
  // give the bits of one); only [H | underflow | nNA | bad] are doubles
  RET_IF(allreduce(h, bufs, tab + 2 * pl.KP, nccl::kInt64, nccl::kSum));
  std::vector<void*> tails(bufs.size());
  for (size_t i = 0; i < bufs.size(); i++) tails[i] = (double*)bufs[i] + tab + 2 * pl.KP;
  return allreduce(h, tails, 4, nccl::kDouble, nccl::kSum);
}

// engine 5: layout of the segment lists of a tile of R rows (seg.cuh).  fs: the features are dealt
// alternately to the two CTAs of a cluster (the C_j pattern of neighbouring features balances).
static void layout_seg(const mixdir_b200_handle* h, int R, int G, bool fs, bool gs, int W, SegPlan* sp) {
  const int J = h->J;
  *sp = SegPlan();
  sp->fs = fs;
  sp->gs = gs;
  sp->G = G;
  const int owners = fs ? 2 : 1;
  int blk = 0;
  for (int o = 0; o < owners; o++) {
    std::vector<int> feats;
    for (int j = 0; j < J; j++)
      if (!fs || j % 2 == o) feats.push_back(j);
    // warp w works on jobs w, w + W, ...: order them by falling C_j (more categories = more groups
    // = more per-group work) and deal the rounds in snake order so the warps finish together
    std::stable_sort(feats.begin(), feats.end(), [&](int a, int b) { return h->ncat[a] > h->ncat[b]; });
    for (size_t r0 = (size_t)W; r0 < feats.size(); r0 += 2 * (size_t)W)
      std::reverse(feats.begin() + r0, feats.begin() + std::min(feats.size(), r0 + W));
    sp->job0[o] = (int)sp->jobs.size();
    sp->njobs[o] = (int)feats.size();
    sp->sc0[o] = (int)sp->srow_to_stat.size();
    int cnt_bytes = 0, sc = 0;  // a job's group counts start at a 4-byte boundary
    for (int j : feats) {
      cnt_bytes += (h->ncat[j] + 3) & ~3;
      sc += h->ncat[j];
    }
    sp->SC[o] = gs ? 0 : sc;
    int off = (cnt_bytes + 15) & ~15, c0 = 0, s0 = 0;
    for (int j : feats) {
      // first statistics row: inside the owner's shared-memory table, or (gs) of the global feature table
      sp->jobs.push_back(make_int4(off, c0, h->ncat[j] | (j << 8), gs ? h->rowoff[j] + 1 : s0));
      for (int c = 1; c <= h->ncat[j] && !gs; c++) sp->srow_to_stat.push_back(h->rowoff[j] + c);
      c0 += (h->ncat[j] + 3) & ~3;
      s0 += h->ncat[j];
      off += (R + (G - 1) * std::min(h->ncat[j], R) + 15) & ~15;
    }
    sp->blk_off[o] = blk;
    sp->blk_size[o] = off;
    blk += off;
  }
  sp->TB = blk;
}

// engine 5: k_em_seg.  Classes per CTA 16 or 32, 8 row groups per lane, tiles of at most 255 rows
// (row numbers are bytes, 255 is the dummy).  MIXDIR_B200_FORCE_PLAN="KC,-,pairs,-,W" restricts
// the search as for the unit kernel.
static bool plan_seg_gs(const mixdir_b200_handle* h, Shard& s, int K, int gs, Plan* out) {
  if (h->CB != 1) return false;
  int fKC = -1, fRG = -1, fM = -1, fFPG = -1, fW = -1;
  if (const char* f = getenv("MIXDIR_B200_FORCE_PLAN"))
    sscanf(f, "%d,%d,%d,%d,%d", &fKC, &fRG, &fM, &fFPG, &fW);
  const uint8_t* na = (h->scan_resolved && (int)h->has_na.size() == h->J && !getenv("MIXDIR_B200_KEEP_NA_ROWS"))
                          ? h->has_na.data() : nullptr;
  Plan best;
  double best_cost = 1e300;
  // MIXDIR_B200_SEG_GLOBAL_STATS=1: no statistics table in shared memory, every (tile, feature, code)
  // sum goes to the global table with RED.ADD.U64.  That lets K = 32 run in ONE CTA (no cluster) and
  // K = 64, J = 100 in clusters of 4 x 16 classes -- both measured SLOWER than the default plans
  // (DESIGN.md section 8).
  // Automatic use: SMALL problems only, where the freed shared memory buys larger tiles and with them a
  // whole round of tiles less (BASELINE configs[2], 70k x 60, K = 10: 160-row tiles in 3 rounds -> 240-row
  // tiles in 2, E+M pass 106 -> 94 us), see plan_seg.
  const int kcs[2] = {32, 16};
  for (int KC : kcs) {
    if (fKC > 0 && KC != fKC) continue;
    if (KC == 32 && K <= 16) continue;
    const int RPW = 32 / KC, RG = 8;
    const int P = (K + KC - 1) / KC;
    if (P > 8) continue;
    const int Wmax = fW > 0 ? fW : (KC == 16 ? 15 : 20);
    for (int W = Wmax; W >= 4; W--) {
      if (fW > 0 && W != fW) continue;
      const int R = W * RG * RPW;
      if (R > kSegDummyRow) continue;
     {
      Plan c;
      const bool fs = KC == 16 && P == 2 && !getenv("MIXDIR_B200_NO_FEATURE_SPLIT");
      // warp-specialised kernel (seg_ws.cuh): 16 M-step warps next to the W E-step warps
      const bool ws = fs && !gs && W <= 15 && getenv("MIXDIR_B200_SEG_WS") != nullptr;
      layout_seg(h, R, fs ? 4 : 4 * RPW, fs, gs != 0, ws ? 16 : W, &c.sp);
      c.sp.ws = ws;
      const int maxSC = std::max(c.sp.SC[0], c.sp.SC[1]), maxJobs = std::max(c.sp.njobs[0], c.sp.njobs[1]);
      const int maxBlk = std::max(c.sp.blk_size[0], c.sp.blk_size[1]);
      UnitPlan up;
      bool found = false;
      const int m_hi = fM >= 0 ? std::min(fM, h->J / 2) : (h->pairing ? h->J / 2 : 0);
      const int w_lo = (h->J - m_hi + 3) / 4, w_hi = (h->J + 3) / 4;
      for (int w = w_lo; w <= w_hi; w++) {
        const int m = fM >= 0 ? m_hi : std::max(0, h->J - 4 * w);
        if (m > m_hi || !make_units(h, m, &up, true, na)) continue;
        layout_units(&up, KC);
        if (((size_t)R * up.WPRu) % 4 != 0) continue;  // a tile is one bulk copy: 16-byte multiple
        const size_t tot = SegSmem(up.srows, maxSC, fs ? 32 : KC, KC, W, RG, up.WPRu, maxJobs, maxBlk).total;
        if (tot > (size_t)s.smem_optin) continue;
        c.smem = tot;
        found = true;
        break;
      }
      if (!found) continue;
      c.engine = 5; c.KC = KC; c.P = P; c.RG = RG; c.W = W; c.KP = P * KC; c.R = R; c.FPG = 0;
      c.up = up;
      c.n_clusters = plan_occupancy(h, s, c);
      if (c.n_clusters < 1) continue;
      // shared-memory cycles per row and CTA: one 128-byte wavefront per (row, element) and 16
      // classes in the E-step; the M-step reads 4 bytes per (row, feature, class) but is bound by
      // instruction issue (~0.8 cycles per row and feature at KC=16, ~1.5 at KC=32); warps idle in
      // the last round of the feature rotation; per-row overhead (soft-max, exchange) as in plan_units
      const double used = std::min(1.0, (double)c.n_clusters * P / s.sm_count);
      const int rounds = (maxJobs + W - 1) / W;
      const double e_cost = up.U * (KC / 16.0);
      const double m_cost = (fs ? 1.5 : (KC == 16 ? 1.6 : 1.5)) * rounds * W;
      const double ovh = KC == 16 ? 65.0 : 80.0;
      // the time of a tile does not depend on the number of warps (each warp walks its own 16 rows and
      // its own features; sweep 5..15 warps: time x warps constant), so fewer warps = smaller tiles cost
      // proportionally more
      const double wpen = (KC == 16 ? 15.0 : 20.0) / W;
      const double cost = P * (e_cost + m_cost + ovh) * wpen / used + P * 3000.0 / R;
      if (cost < best_cost) {
        best_cost = cost;
        best = c;
      }
     }
    }
  }
  if (best.engine != 5) return false;
  *out = best;
  return true;
}

// rounds of tiles ONE GPU's persistent grid would walk over ALL rows under a plan.  Deliberately not per
// shard: the plan must not depend on the number of GPUs, or a fit on N GPUs would add its logit terms in
// another order (another pair plan) than on one and lose its bit-identity with it.
static int64_t seg_rounds(const mixdir_b200_handle* h, const Plan& p) {
  const int64_t total = h->n_rows_global > 0 ? h->n_rows_global : h->n_rows * (int64_t)h->world;
  const int64_t tiles = (total + p.R - 1) / p.R;
  return (tiles + p.n_clusters - 1) / std::max(p.n_clusters, 1);
}

static bool plan_seg(const mixdir_b200_handle* h, Shard& s, int K, Plan* out) {
  if (getenv("MIXDIR_B200_SEG_GLOBAL_STATS")) return plan_seg_gs(h, s, K, 1, out);  // A/B switch
  Plan a;
  if (!plan_seg_gs(h, s, K, 0, &a)) return false;
  *out = a;
  const int64_t ra = seg_rounds(h, a);
  if (ra <= 4 && !getenv("MIXDIR_B200_SEG_NO_GLOBAL_STATS")) {  // a small problem: would larger tiles save a round?
    Plan b;
    if (plan_seg_gs(h, s, K, 1, &b) && b.KC == a.KC && b.P == a.P && seg_rounds(h, b) < ra) *out = b;
  }
  return true;
}

// host-only view of the unit kernel's shared-memory placement for KC classes per CTA (CPU tests)
extern "C" int mixdir_b200_unit_layout(int n_features, const int* ncat, int n_pairs,
                                       const unsigned char* has_na, int classes_per_cta,
                                       int* n_elements, int* n_table_rows,
                                       int* smem_rows, int* desc, int* gbase, int* slot,
                                       int* eoff_rot) {
  if (classes_per_cta != 4 && classes_per_cta != 8 && classes_per_cta != 16 && classes_per_cta != 32)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "classes_per_cta must be 4, 8, 16 or 32");
  mixdir_b200_handle h;
  RET_IF(init_geometry(&h, 1, n_features, ncat));
  UnitPlan up;
  if (n_pairs < 0 || !make_units(&h, n_pairs, &up, true, has_na))
    return fail(MIXDIR_B200_BAD_ARGUMENT, "this many pairs cannot be formed (a pair code must fit a byte)");
  layout_units(&up, classes_per_cta);
  if (n_elements) *n_elements = up.U;
  if (n_table_rows) *n_table_rows = up.NR2;
  if (smem_rows) *smem_rows = up.srows;
  if (desc) memcpy(desc, up.desc.data(), up.desc.size() * sizeof(UnitDesc));
  if (gbase) memcpy(gbase, up.gbase.data(), up.gbase.size() * sizeof(int));
  if (slot) memcpy(slot, up.slot.data(), up.slot.size() * sizeof(int));
  if (eoff_rot) memcpy(eoff_rot, up.eoff_rot.data(), up.eoff_rot.size() * sizeof(int));
  return MIXDIR_B200_OK;
}

// engine 2: k_em_fused (row-owned M-step, read-modify-write per row and unit)
static bool plan_fused(const mixdir_b200_handle* h, Shard& s, int K, Plan* out) {
  Plan best;
  double best_cost = 1e300;
  const int kcs[3] = {32, 16, 8};
  for (int KC : kcs) {
    const int RPW = 32 / KC, FPW = 4 / h->CB;
    if (RPW > FPW) continue;
    const int P = (K + KC - 1) / KC;
    if (P > 8) continue;
    const int RG0 = P <= 2 ? 8 : (P <= 4 ? 4 : 2);
    // most feature pairs whose unit tables (+ tiles) still fit shared memory
    UnitPlan up;
    bool found = false;
    int W = 0, RG = RG0;
    size_t smem = 0;
    // a row tile is one bulk copy: its byte count (and the second buffer's offset) must be a
    // multiple of 16
    auto shape = [&](const UnitPlan& u) {
      W = std::min(16, u.WPRu);
      RG = RG0;
      while (((size_t)W * RG * RPW * u.WPRu) % 4 != 0 && RG < 8) RG *= 2;
      return FusedSmem(u.NR2, KC, W, RG, u.WPRu, FPW, P);
    };
    const int m_hi = (h->pairing && h->CB == 1) ? h->J / 2 : 0;
    const int w_lo = (h->J - m_hi + 3) / 4, w_hi = (h->J + 3) / 4;
    for (int w = m_hi ? w_lo : w_hi; w <= w_hi; w++) {
      // smallest number of pairs that gives w code words per row (none: the raw code matrix)
      const int m = m_hi ? std::max(0, h->J - 4 * w) : 0;
      if (m > m_hi || !make_units(h, m, &up)) continue;
      FusedSmem L = shape(up);
      if (L.total > (size_t)s.smem_optin) continue;
      smem = L.total;
      found = true;
      break;
    }
    if (!found) continue;
    Plan c;
    c.engine = 2; c.KC = KC; c.P = P; c.RG = RG; c.W = W; c.KP = P * KC;
    c.smem = smem; c.R = W * RG * RPW;
    c.n_clusters = plan_occupancy(h, s, c);
    if (c.n_clusters < 1) continue;
    // cost ~ shared-memory wavefronts per row over the SMs in use: P CTAs x elements per row /
    // rows per instruction x conflict factor
    const double conflict = KC == 8 ? 1.5 : 1.0;
    const double used = std::min(1.0, (double)c.n_clusters * P / s.sm_count);
    const double cost = P * conflict * (up.WPRu * FPW) / RPW / used + 1e-3 * P;
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
      best.up = up;
    }
  }
  if (best.engine != 2) return false;
  *out = best;
  return true;
}

// engine_pref: 0 automatic (segment kernel, else unit kernel, else fused, else generic), 1 generic,
// 2 fused, 4 units, 5 sorted segments.
// Planning happens on the first shard's device (all shards of a handle are the same GPU model).
static int plan_engine(mixdir_b200_handle* h, Shard& s, int K, Plan* out) {
  CU_TRY(cudaSetDevice(s.device));
  const int pref = h->engine_pref;
  // Automatic mode takes the segment kernel from 16384 rows on: below, rounding the responsibilities
  // to 32 bits can cost more than 1e-9 of the ELBO (seg.cuh), and such fits are launch-bound anyway.
  // MIXDIR_B200_FP64_STATS=1 keeps fp64 statistics (unit kernel) in automatic mode.
  static const bool fp64_stats = getenv("MIXDIR_B200_FP64_STATS") != nullptr;
  const int64_t total_rows = h->n_rows_global > 0 ? h->n_rows_global : h->n_rows * (int64_t)h->world;
  const bool seg_ok = pref == 5 || (!fp64_stats && total_rows >= 16384);
  if ((pref == 0 || pref == 5) && seg_ok && plan_seg(h, s, K, out)) return MIXDIR_B200_OK;
  if (pref == 5)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "segment kernel does not fit this problem shape (K, sum C_j) "
                                          "or the codes are 2 bytes wide");
  if ((pref == 0 || pref == 4) && plan_units(h, s, K, out)) return MIXDIR_B200_OK;
  if (pref == 4)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "unit kernel does not fit this problem shape (K, sum C_j) "
                                          "or the codes are 2 bytes wide");
  if ((pref == 0 || pref == 2) && plan_fused(h, s, K, out)) return MIXDIR_B200_OK;
  if (pref == 2)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "fused kernel does not fit this problem shape (K, sum C_j)");
  Plan g;
  g.engine = 1;
  g.KP = (K + 7) & ~7;
  identity_units(h, &g.up);
  *out = g;
  return MIXDIR_B200_OK;
}

static int ensure_workspace(mixdir_b200_handle* h, Shard& s, int K, const Plan& pl, int max_iter) {
  CU_TRY(cudaSetDevice(s.device));
  const int KP = pl.KP;
  const int nclus = pl.engine >= 2 ? pl.n_clusters : 0;
  const int nctas = nclus * pl.P;
  if (s.ws_K == K && s.ws_KP == KP && s.ws_iter >= max_iter && s.ws_clusters >= nclus &&
      s.ws_ctas >= nctas && s.ws_NR2 >= pl.up.NR2 && (pl.engine == 5 || nclus == 0 || s.d_Spart))
    return MIXDIR_B200_OK;
  free_workspace(s);
  const size_t tab = (size_t)h->NR * KP;
  const size_t tab2 = (size_t)pl.up.NR2 * KP;
  const size_t rag = (size_t)h->sum_ncat * K;
  RET_IF(dalloc(&s.d_T, tab));
  RET_IF(dalloc(&s.d_phi, tab));
  RET_IF(dalloc(&s.d_logU, tab));
  RET_IF(dalloc(&s.d_stats, tab + 2 * KP + 4));  // (k_em_seg: two words per class mass)
  RET_IF(dalloc(&s.d_logprior, KP));
  RET_IF(dalloc(&s.d_cs_a, kMaxClasses));
  RET_IF(dalloc(&s.d_part, (size_t)h->J * 3));
  RET_IF(dalloc(&s.d_trace, std::max(max_iter, 1)));
  RET_IF(dalloc(&s.d_rag_a, rag));
  RET_IF(dalloc(&s.d_rag_b, rag));
  RET_IF(dalloc(&s.d_lambda, K));
  RET_IF(dalloc(&s.d_lambda_pad, KP));
  RET_IF(dalloc(&s.d_prior_out, 2 * K));
  RET_IF(dalloc(&s.d_ticket, 1));
  RET_IF(dalloc(&s.d_underflow, 1));
  RET_IF(dalloc(&s.d_warn, 1));
  RET_IF(dalloc(&s.d_prior, 1));
  RET_IF(dalloc(&s.d_status, 1));
  if (nclus > 0) {
    if (pl.engine != 5) RET_IF(dalloc(&s.d_Spart, (size_t)nclus * tab2));
    RET_IF(dalloc(&s.d_cspart, (size_t)nclus * KP));
    RET_IF(dalloc(&s.d_Hpart, nctas));
  }
  CU_TRY(cudaMemsetAsync(s.d_T, 0, tab * sizeof(double), s.stream));
  CU_TRY(cudaMemsetAsync(s.d_phi, 0, tab * sizeof(double), s.stream));
  CU_TRY(cudaMemsetAsync(s.d_logU, 0, tab * sizeof(double), s.stream));
  CU_TRY(cudaMemsetAsync(s.d_ticket, 0, sizeof(unsigned int), s.stream));
  CU_TRY(cudaMemsetAsync(s.d_underflow, 0, sizeof(int), s.stream));
  CU_TRY(cudaMemsetAsync(s.d_status, 0, sizeof(IterStatus), s.stream));
  s.ws_K = K;
  s.ws_KP = KP;
  s.ws_iter = std::max(max_iter, 1);
  s.ws_clusters = nclus;
  s.ws_ctas = nctas;
  RET_IF(dalloc(&s.d_T2, tab2));
  s.ws_NR2 = pl.up.NR2;
  return MIXDIR_B200_OK;
}

__global__ void k_copy_words(uint64_t* __restrict__ dst, const uint64_t* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// start a new batch of staged inputs (the compute stream is idle between API calls)
static int pin_reset(Shard& s, size_t need) {
  CU_TRY(cudaSetDevice(s.device));
  if (s.h_pin_bytes < need) {
    CU_TRY(cudaStreamSynchronize(s.stream));
    if (s.h_pin) cache::pin_put(s.h_pin);
    s.h_pin = nullptr;
    s.h_pin_bytes = 0;
    CU_TRY(cache::pin_get((void**)&s.h_pin, need));
    s.h_pin_bytes = need;
  }
  s.h_pin_used = 0;
  return MIXDIR_B200_OK;
}

// host array -> device array through the mapped scratch (bytes must be a multiple of 8)
static int pin_send(Shard& s, void* dst_dev, const void* src_host, size_t bytes) {
  if (bytes == 0) return MIXDIR_B200_OK;
  if (s.h_pin_used + bytes > s.h_pin_bytes) return fail(MIXDIR_B200_CUDA_ERROR, "staging buffer overflow");
  unsigned char* p = s.h_pin + s.h_pin_used;
  memcpy(p, src_host, bytes);
  s.h_pin_used += (bytes + 63) & ~(size_t)63;
  void* dp = nullptr;
  CU_TRY(cudaHostGetDevicePointer(&dp, p, 0));
  const size_t n = bytes / 8;
  const int grid = (int)std::min<size_t>((n + 255) / 256, 64);
  k_copy_words<<<grid, 256, 0, s.stream>>>((uint64_t*)dst_dev, (const uint64_t*)dp, n);
  CU_TRY(cudaGetLastError());
  return MIXDIR_B200_OK;
}

static int upload_ragoff(mixdir_b200_handle* h, Shard& s, int K) {
  std::vector<int64_t> ro(h->J);
  for (int j = 0; j < h->J; j++) ro[j] = h->ragoff1[j] * K;
  return pin_send(s, s.d_ragoff, ro.data(), h->J * sizeof(int64_t));
}

static size_t pad8(size_t b) { return (b + 7) & ~(size_t)7; }
static size_t unit_stage_bytes(const UnitPlan& up) {
  return pad8(up.desc.size() * sizeof(UnitDesc)) + pad8(up.gbase.size() * sizeof(int)) +
         pad8(up.t2src.size() * sizeof(int2)) + pad8(up.marg.size() * sizeof(MargDesc)) +
         pad8(up.slot.size() * sizeof(int)) + pad8(up.eoff_rot.size() * sizeof(int)) +
         2 * pad8(up.desc.size() * sizeof(int2)) + 10 * 64;
}
static size_t seg_stage_bytes(const SegPlan& sp) {
  return pad8(sp.jobs.size() * sizeof(int4)) + pad8(sp.srow_to_stat.size() * sizeof(int)) + 2 * 64;
}

template <typename T>
static int send_vec(Shard& s, T** dptr, const std::vector<T>& vec) {
  const size_t bytes = pad8(vec.size() * sizeof(T));
  RET_IF(dalloc((unsigned char**)dptr, bytes));
  std::vector<unsigned char> tmp(bytes, 0);
  memcpy(tmp.data(), vec.data(), vec.size() * sizeof(T));
  return pin_send(s, *dptr, tmp.data(), bytes);
}

// Device copies of the unit plan (+ the packed code buffer) for this shard; kept across fits
// while the plan stays the same.  Call after pin_reset (the arrays go through the staging area).
static int ensure_units(mixdir_b200_handle* h, Shard& s, const Plan& pl) {
  const UnitPlan& up = pl.up;
  CU_TRY(cudaSetDevice(s.device));
  std::vector<UnitDesc> sig = up.desc;  // element order + what the placement depends on
  sig.push_back(UnitDesc{pl.engine, pl.KC, up.Q, up.srows});
  sig.push_back(UnitDesc{pl.R, pl.RG, pl.W, pl.FPG});
  sig.push_back(UnitDesc{pl.sp.fs ? 1 : 0, pl.sp.G, pl.sp.TB, (pl.sp.gs ? 1 : 0) | (pl.sp.ws ? 2 : 0)});
  if (s.unit_valid && s.unit_sig.size() == sig.size() &&
      memcmp(s.unit_sig.data(), sig.data(), sig.size() * sizeof(UnitDesc)) == 0)
    return MIXDIR_B200_OK;
  s.unit_valid = false;
  dfree(s.d_udesc); dfree(s.d_ugbase); dfree(s.d_t2src); dfree(s.d_marg); dfree(s.d_ucodes);
  dfree(s.d_slot); dfree(s.d_eoff); dfree(s.d_elem); dfree(s.d_erange);
  dfree(s.d_seg); dfree(s.d_jobs); dfree(s.d_srow2stat);
  s.packed_rows = 0;
  s.seg_tiles = 0;
  {  // k_iter's work list: the real elements (feature, or the two features of a pair)
    std::vector<int2> elem, erange;
    for (int e = 0; e < up.U; e++) {
      const UnitDesc& d = up.desc[e];
      if (d.off_a < 0) continue;
      if (up.identity) {
        elem.push_back(make_int2(d.off_a / h->CB, -1));
        erange.push_back(make_int2(0, 0));
      } else {
        elem.push_back(make_int2(d.off_a, d.off_b < 0 ? -1 : d.off_b));
        erange.push_back(make_int2(up.gbase[e], up.erows[e]));
      }
    }
    s.n_elem = (int)elem.size();
    RET_IF(send_vec(s, &s.d_elem, elem));
    RET_IF(send_vec(s, &s.d_erange, erange));
  }
  RET_IF(send_vec(s, &s.d_udesc, up.desc));
  RET_IF(send_vec(s, &s.d_ugbase, up.gbase));
  RET_IF(send_vec(s, &s.d_marg, up.marg));
  if (!up.identity) {
    RET_IF(send_vec(s, &s.d_t2src, up.t2src));
    const size_t rowb = (size_t)up.WPRu * 4;
    // engine 4 stores the packed codes tile by tile ([tile][word][row in tile]): whole tiles,
    // the tail of the last one zeroed; engine 2 row-major with kRowPad zero rows behind
    const int64_t R = pl.engine >= 4 ? pl.R : 1;
    const int64_t full = s.n_rows / R * R, rows_alloc = (s.n_rows + R - 1) / R * R + kRowPad;
    RET_IF(dalloc(&s.d_ucodes, (size_t)rows_alloc * rowb));
    CU_TRY(cudaMemsetAsync(s.d_ucodes + (size_t)full * rowb, 0, (size_t)(rows_alloc - full) * rowb, s.stream));
  }
  if (pl.engine >= 4) {
    RET_IF(send_vec(s, &s.d_slot, up.slot));
    RET_IF(send_vec(s, &s.d_eoff, up.eoff_rot));
  }
  if (pl.engine == 5) {
    RET_IF(send_vec(s, &s.d_jobs, pl.sp.jobs));
    RET_IF(send_vec(s, &s.d_srow2stat, pl.sp.srow_to_stat));
    const int64_t n_tiles = (s.n_rows + pl.R - 1) / pl.R;
    RET_IF(dalloc(&s.d_seg, (size_t)std::max<int64_t>(n_tiles, 1) * pl.sp.TB));
  }
  s.unit_sig = sig;
  s.unit_valid = true;
  return MIXDIR_B200_OK;
}

// pack raw rows [packed_rows, upto) into unit codes (their upload must have been consumed)
static int pack_units(mixdir_b200_handle* h, Shard& s, const Plan& pl, int64_t upto, int* launches) {
  if (pl.up.identity || s.packed_rows >= upto) return MIXDIR_B200_OK;
  const int64_t r0 = s.packed_rows, nr = upto - r0;
  const int WPRu = pl.up.WPRu;
  const int grid = (int)std::min<int64_t>((int64_t)s.sm_count * 16, (nr * WPRu + 255) / 256);
  k_pack_units<<<grid, 256, 0, s.stream>>>(s.d_codes + (size_t)r0 * h->row_bytes, h->row_bytes, nr,
                                           s.d_udesc, WPRu, pl.engine >= 4 ? pl.R : 0,
                                           reinterpret_cast<uint32_t*>(s.d_ucodes) + (size_t)r0 * WPRu);
  CU_TRY(cudaGetLastError());
  s.packed_rows = upto;
  if (launches) *launches += 1;
  return MIXDIR_B200_OK;
}

// engine 5: build the segment lists of the tiles that end at or before row `upto` (or hold the
// last rows of the shard); their raw codes must have been consumed from the upload
static int pack_seg(mixdir_b200_handle* h, Shard& s, const Plan& pl, int64_t upto, int* launches) {
  if (pl.engine != 5) return MIXDIR_B200_OK;
  const int64_t t1 = upto >= s.n_rows ? (s.n_rows + pl.R - 1) / pl.R : upto / pl.R;
  if (t1 <= s.seg_tiles) return MIXDIR_B200_OK;
  const int64_t t0 = s.seg_tiles;
  PackSegArgs a;
  a.raw = s.d_codes + (size_t)t0 * pl.R * h->row_bytes;
  a.row_bytes = h->row_bytes;
  a.n_rows = s.n_rows - t0 * pl.R;
  a.n_tiles = (int)(t1 - t0);
  a.R = pl.R; a.J = h->J; a.G = pl.sp.G; a.TB = pl.sp.TB;
  a.split = pl.sp.njobs[0];
  a.blk_off[0] = pl.sp.blk_off[0]; a.blk_off[1] = pl.sp.blk_off[1];
  a.jobs = s.d_jobs;
  a.out = s.d_seg + (size_t)t0 * pl.sp.TB;
  const int grid = (int)std::min<int64_t>(a.n_tiles, (int64_t)s.sm_count * 32);
  k_pack_seg<<<grid, 256, 0, s.stream>>>(a);
  CU_TRY(cudaGetLastError());
  s.seg_tiles = t1;
  if (launches) *launches += 1;
  return MIXDIR_B200_OK;
}

// unit table T2 from the feature table T (unit-test door; a fit builds it in k_iter)
static int build_unit_table(Shard& s, const Plan& pl) {
  if (pl.up.identity) return MIXDIR_B200_OK;
  const size_t n = (size_t)pl.up.NR2 * pl.KP;
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)s.sm_count * 4);
  k_build_T2<<<grid, 256, 0, s.stream>>>(s.d_t2src, pl.up.NR2, pl.KP, s.d_T, s.d_T2);
  CU_TRY(cudaGetLastError());
  return MIXDIR_B200_OK;
}

// One E+M pass on a shard: statistics land in s.d_stats (local, before any all-reduce).
// While an asynchronous upload is still in flight the pass runs CHUNKED: the compute stream
// waits for upload chunk c (and scans it), then launches the kernel on the rows that are
// complete so far, so the first iteration overlaps the host-to-device copy.
static int launch_em(mixdir_b200_handle* h, Shard& s, int K, const Plan& pl, double* zeta_out,
                     bool scatter, cudaEvent_t ev0, cudaEvent_t ev1, int* launches) {
  CU_TRY(cudaSetDevice(s.device));
  const size_t tab = (size_t)h->NR * pl.KP;
  const size_t tab2 = (size_t)pl.up.NR2 * pl.KP;
  const UnitPlan& up = pl.up;
  const uint8_t* ucodes = up.identity ? s.d_codes : s.d_ucodes;
  const size_t urow_bytes = up.identity ? (size_t)h->row_bytes : (size_t)up.WPRu * 4;
  const int* stop = &s.d_status->stop;
  // row ranges of this pass: one per pending upload chunk, else the whole shard
  std::vector<std::pair<int64_t, int64_t>> ranges;
  std::vector<size_t> range_chunk;  // upload chunk to consume before range i (or SIZE_MAX)
  const int64_t align = pl.engine >= 2 ? pl.R : 1;
  if (s.up_seen < s.up_ev.size()) {
    int64_t prev = 0;
    for (size_t c = 0; c < s.up_ev.size(); c++) {
      const bool last = c + 1 == s.up_ev.size();
      const int64_t end = last ? s.n_rows : (s.up_end[c] / align) * align;
      if (c < s.up_seen && !last) continue;  // (cannot happen today; keeps the logic general)
      if (end > prev || last) {
        ranges.push_back({prev, std::max(end, prev)});
        range_chunk.push_back(c);
        prev = std::max(end, prev);
      }
    }
  } else {
    ranges.push_back({0, s.n_rows});
    range_chunk.push_back(SIZE_MAX);
  }
  const bool chunked = ranges.size() > 1;

  if (pl.engine >= 2 && s.n_rows > 0) {
    if (chunked) {  // partial buffers accumulate over the launches
      if (pl.engine != 5) {
        CU_TRY(cudaMemsetAsync(s.d_Spart, 0, (size_t)pl.n_clusters * tab2 * sizeof(double), s.stream));
        CU_TRY(cudaMemsetAsync(s.d_cspart, 0, (size_t)pl.n_clusters * pl.KP * sizeof(double), s.stream));
      }
      CU_TRY(cudaMemsetAsync(s.d_Hpart, 0, (size_t)pl.n_clusters * pl.P * sizeof(double), s.stream));
    }
    if (ev0) CU_TRY(cudaEventRecord(ev0, s.stream));
    int max_clus = 0;
    for (size_t i = 0; i < ranges.size(); i++) {
      if (range_chunk[i] != SIZE_MAX) RET_IF(consume_upload(h, s, range_chunk[i]));
      const int64_t r0 = ranges[i].first, nr = ranges[i].second - ranges[i].first;
      if (nr <= 0) continue;
      RET_IF(pack_units(h, s, pl, ranges[i].second, launches));
      RET_IF(pack_seg(h, s, pl, ranges[i].second, launches));
      const int n_tiles = (int)((nr + pl.R - 1) / pl.R);
      const int nclus = std::min(pl.n_clusters, n_tiles);
      max_clus = std::max(max_clus, nclus);
      FusedArgs a;
      a.codes = reinterpret_cast<const uint32_t*>(ucodes + (size_t)r0 * urow_bytes);
      a.n_rows = nr;
      a.n_tiles = n_tiles;
      a.WPR = up.WPRu;
      a.NR = up.NR2;
      a.K = K;
      a.KP = pl.KP;
      a.P = pl.P;
      a.rowoff_pad = s.d_ugbase;
      a.T = up.identity ? s.d_T : s.d_T2;
      a.logprior = s.d_logprior;
      a.Spart = s.d_Spart;
      a.cspart = s.d_cspart;
      a.Hpart = s.d_Hpart;
      a.underflow = s.d_underflow;
      a.accumulate = chunked ? 1 : 0;
      a.bad = &s.d_scan->bad;
      a.stop = stop;
      CU_TRY(ensure_smem_attr(s, plan_fn(h, pl), pl.smem));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nclus * pl.P);
      cfg.blockDim = dim3(plan_threads(pl));
      cfg.dynamicSmemBytes = pl.smem;
      cfg.stream = s.stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = pl.P;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (pl.engine == 5) {
        SegArgs g;
        g.codes = a.codes;
        g.seg = s.d_seg + (size_t)(r0 / pl.R) * pl.sp.TB;
        g.n_rows = nr;
        g.row_base = h->row_offset + s.row0 + r0;
        g.n_tiles = n_tiles;
        g.WPR = up.WPRu;
        g.NR2 = up.NR2;
        g.srows = up.srows;
        g.K = K;
        g.KP = pl.KP;
        g.P = pl.P;
        g.TB = pl.sp.TB;
        for (int o = 0; o < 2; o++) {
          g.blk_off[o] = pl.sp.blk_off[o]; g.blk_size[o] = pl.sp.blk_size[o];
          g.job0[o] = pl.sp.job0[o]; g.njobs[o] = pl.sp.njobs[o];
          g.sc0[o] = pl.sp.sc0[o]; g.SC[o] = pl.sp.SC[o];
        }
        g.gs = pl.sp.gs ? 1 : 0;
        g.ws_we = pl.W;
        g.ws_debug = getenv("MIXDIR_B200_WS_DEBUG") ? atoi(getenv("MIXDIR_B200_WS_DEBUG")) : 0;
        g.maxSC = std::max(pl.sp.SC[0], pl.sp.SC[1]);
        g.maxJobs = std::max(pl.sp.njobs[0], pl.sp.njobs[1]);
        g.maxBlk = std::max(pl.sp.blk_size[0], pl.sp.blk_size[1]);
        g.eoff_rot = s.d_eoff;
        g.slot_of_row = s.d_slot;
        g.T2 = s.d_T2;
        g.logprior = s.d_logprior;
        g.jobs = s.d_jobs;
        g.srow_to_stat = s.d_srow2stat;
        g.stats = reinterpret_cast<unsigned long long*>(s.d_stats);
        g.cs_fixed = g.stats + tab;
        g.stats_tail = s.d_stats + tab + 2 * pl.KP;
        g.Hpart = s.d_Hpart;
        g.n_hsum = (chunked ? pl.n_clusters : nclus) * pl.P;
        g.ticket = s.d_ticket;
        g.underflow = s.d_underflow;
        g.scan = s.d_scan;
        g.accumulate = chunked ? 1 : 0;
        g.bad = &s.d_scan->bad;
        g.stop = stop;
        CU_TRY(cudaLaunchKernelEx(&cfg, pick_seg(pl.KC, pl.W, pl.sp.fs, pl.sp.ws), g));
      } else if (pl.engine == 4) {
        UnitArgs u;
        u.codes = a.codes;
        u.n_rows = nr;
        u.n_tiles = n_tiles;
        u.WPR = up.WPRu;
        u.NR2 = up.NR2;
        u.srows = up.srows;
        u.K = K;
        u.KP = pl.KP;
        u.P = pl.P;
        u.eoff_rot = s.d_eoff;
        u.slot_of_row = s.d_slot;
        u.T2 = s.d_T2;
        u.logprior = s.d_logprior;
        u.Spart = s.d_Spart;
        u.cspart = s.d_cspart;
        u.Hpart = s.d_Hpart;
        u.underflow = s.d_underflow;
        u.accumulate = chunked ? 1 : 0;
        u.bad = &s.d_scan->bad;
        u.stop = stop;
        CU_TRY(cudaLaunchKernelEx(&cfg, pick_units(pl.KC, pl.RG, pl.FPG, pl.W), u));
      } else {
        CU_TRY(cudaLaunchKernelEx(&cfg, pick_fused(pl.KC, h->CB, pl.RG), a));
      }
      *launches += 1;
    }
    if (ev1) CU_TRY(cudaEventRecord(ev1, s.stream));
    h->timing.grid_ctas = max_clus * pl.P;
    if (pl.engine == 5) return MIXDIR_B200_OK;  // statistics already in place (atomics + last CTA)
    ReduceUnitsArgs r;
    r.NR = h->NR;
    r.NR2 = up.NR2;
    r.KP = pl.KP;
    r.n_clusters = chunked ? pl.n_clusters : max_clus;
    r.n_ctas = r.n_clusters * pl.P;
    r.Spart = s.d_Spart;
    r.cspart = s.d_cspart;
    r.Hpart = s.d_Hpart;
    r.underflow = s.d_underflow;
    r.scan = s.d_scan;
    r.marg = s.d_marg;
    r.stats = s.d_stats;
    r.stop = stop;
    k_reduce_units<<<dim3(h->NR, (pl.KP + 31) / 32), 256, 0, s.stream>>>(r);
    CU_TRY(cudaGetLastError());
    *launches += 1;
  } else {
    CU_TRY(cudaMemsetAsync(s.d_stats, 0, (tab + pl.KP + 4) * sizeof(double), s.stream));
    if (ev0) CU_TRY(cudaEventRecord(ev0, s.stream));
    for (size_t i = 0; i < ranges.size(); i++) {
      if (range_chunk[i] != SIZE_MAX) RET_IF(consume_upload(h, s, range_chunk[i]));
      const int64_t r0 = ranges[i].first, nr = ranges[i].second - ranges[i].first;
      if (nr <= 0) continue;
      GenericArgs g;
      g.codes = s.d_codes + (size_t)r0 * h->row_bytes;
      g.row_bytes = h->row_bytes;
      g.n_rows = nr;
      g.J = h->J;
      g.K = K;
      g.KP = pl.KP;
      g.rowoff = s.d_rowoff;
      g.T = s.d_T;
      g.logprior = s.d_logprior;
      g.stats = s.d_stats;
      g.NR = h->NR;
      g.underflow = s.d_underflow;
      g.zeta_out = zeta_out;
      g.zeta_ld = s.n_rows;
      g.zeta_row0 = r0;
      g.do_scatter = scatter ? 1 : 0;
      g.bad = &s.d_scan->bad;
      g.stop = stop;
      const int grid = s.sm_count * 8;
      if (h->CB == 1)
        k_em_generic<1><<<grid, 256, 0, s.stream>>>(g);
      else
        k_em_generic<2><<<grid, 256, 0, s.stream>>>(g);
      CU_TRY(cudaGetLastError());
      *launches += 1;
    }
    if (ev1) CU_TRY(cudaEventRecord(ev1, s.stream));
    k_tail_to_stats<<<1, 1, 0, s.stream>>>(s.d_underflow, s.d_scan, s.d_stats + tab + pl.KP + 1, stop);
    CU_TRY(cudaGetLastError());
    *launches += 1;
  }
  return MIXDIR_B200_OK;
}

static int check_fit_args(const mixdir_b200_handle* h, const mixdir_b200_fit_params* p) {
  if (!h || !p) return fail(MIXDIR_B200_BAD_ARGUMENT, "NULL handle or params");
  if (p->n_latent < 1 || p->n_latent > kMaxClasses)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "n_latent must be in 1..256");
  if (p->max_iter < 0) return fail(MIXDIR_B200_BAD_ARGUMENT, "max_iter must be >= 0");
  if (!(p->alpha1 > 0) || !(p->beta > 0) || (p->dp && !(p->alpha2 > 0)))
    return fail(MIXDIR_B200_BAD_ARGUMENT, "alpha and beta must be positive");
  if (isnan(p->epsilon)) return fail(MIXDIR_B200_BAD_ARGUMENT, "epsilon is NaN");
  return MIXDIR_B200_OK;
}

static int run_predict(mixdir_b200_handle* h, int K, int KP, double* class_prob, int32_t* pred_class);

extern "C" int mixdir_b200_fit(mixdir_b200_handle* h, const mixdir_b200_fit_params* p,
                               const double* colsum_init, const double* phi_init,
                               double* elbo_trace, int* n_iter, int* converged, double* elbo,
                               double* lambda, double* prior_param, double* phi,
                               double* category_prob, double* class_prob, int32_t* pred_class,
                               int* warn_flags) {
  RET_IF(check_fit_args(h, p));
  if (!colsum_init || !phi_init) return fail(MIXDIR_B200_BAD_ARGUMENT, "colsum_init / phi_init is NULL");
  DeviceGuard guard;
  const int K = p->n_latent, J = h->J;
  // em_kernel_ms costs two event records per iteration: a few microseconds of a 40 us iteration on a
  // small fit, nothing on a large one.  Small fits are timed only on request.
  static const bool force_em_timing = getenv("MIXDIR_B200_TIME_EM") != nullptr;
  const bool time_em_saved = h->time_em;
  struct Restore { mixdir_b200_handle* h; bool v; ~Restore() { h->time_em = v; } } restore{h, time_em_saved};
  if (h->time_em && !force_em_timing && h->n_rows < 65536) h->time_em = false;
  // the table needs n_cat = max(X) (R/mixdir_vi.R:8) before the first pass: when the caller
  // does not pass it, wait for the upload scan (no upload/compute overlap in that case)
  if (p->n_cat_bound <= 0) RET_IF(resolve_scan(h));
  const int n_cat_bound = p->n_cat_bound > 0 ? p->n_cat_bound : h->max_code;
  Plan pl;
  RET_IF(plan_engine(h, h->shards[0], K, &pl));
  const size_t tab = (size_t)h->NR * pl.KP;
  const size_t rag = (size_t)h->sum_ncat * K;
  const int jk_grid = (J * K + 127) / 128;
  h->timing = mixdir_b200_timing();
  h->timing.engine = pl.engine;
  h->timing.classes_per_cta = pl.KC;
  h->timing.cluster_size = pl.P;
  h->timing.smem_bytes = (int)pl.smem;
  h->timing.unit_elements = pl.up.U;
  h->timing.unit_pairs = pl.up.n_pairs;
  h->timing.unit_table_rows = pl.up.NR2;

  // constants of the ELBO that depend on the hyper-parameters only (host libm; ~1 ulp from R's)
  double termC_const = 0;
  for (int j = 0; j < J; j++) termC_const += lgamma(h->ncat[j] * p->beta) - h->ncat[j] * lgamma(p->beta);
  termC_const *= K;
  const double termA_const = lgamma(K * p->alpha1) - K * lgamma(p->alpha1);

  // ---- initial state on every shard: class masses + table from phi_init ----
  auto iter_args = [&](Shard& s, int mode, bool poll) -> IterArgs {
    IterArgs ia;
    ia.J = J; ia.K = K; ia.KP = pl.KP; ia.dp = p->dp; ia.mode = mode;
    ia.n_cat_bound = n_cat_bound; ia.NR = h->NR; ia.NR2 = pl.up.NR2;
    ia.stats_fixed = pl.engine == 5 ? 1 : 0;
    ia.beta = p->beta; ia.a1 = p->alpha1; ia.a2 = p->alpha2; ia.epsilon = p->epsilon;
    ia.termC_const = termC_const; ia.termA_const = termA_const;
    ia.ncat = s.d_ncat; ia.rowoff = s.d_rowoff; ia.elem = s.d_elem; ia.erange = s.d_erange;
    ia.t2src = pl.up.identity ? nullptr : s.d_t2src;
    ia.stats = s.d_stats; ia.phi = s.d_phi; ia.T = s.d_T; ia.T2 = s.d_T2; ia.part = s.d_part;
    ia.ticket = s.d_ticket; ia.underflow_flag = s.d_underflow; ia.st = s.d_prior;
    ia.logprior = s.d_logprior; ia.cs_cur = s.d_cs_a; ia.trace = s.d_trace; ia.status = s.d_status;
    ia.status_host = nullptr;
    if (poll) cudaHostGetDevicePointer((void**)&ia.status_host, s.h_status, 0);
    return ia;
  };
  for (auto& s : h->shards) {
    RET_IF(ensure_workspace(h, s, K, pl, p->max_iter));
    RET_IF(pin_reset(s, (rag + K + J) * 8 + 4096 + unit_stage_bytes(pl.up) + seg_stage_bytes(pl.sp)));
    RET_IF(upload_ragoff(h, s, K));
    RET_IF(ensure_units(h, s, pl));
    if (pl.engine == 5)  // the E+M kernel ADDS into the statistics; k_iter clears them after use
      CU_TRY(cudaMemsetAsync(s.d_stats, 0, (tab + 2 * pl.KP + 4) * sizeof(double), s.stream));
    while (h->time_em && (int)s.ev_em.size() < 2 * std::max(p->max_iter, 1)) {
      cudaEvent_t e;
      CU_TRY(cudaEventCreate(&e));
      s.ev_em.push_back(e);
    }
    RET_IF(pin_send(s, s.d_cs_a, colsum_init, K * sizeof(double)));
    RET_IF(pin_send(s, s.d_rag_a, phi_init, rag * sizeof(double)));
    CU_TRY(cudaMemsetAsync(s.d_status, 0, sizeof(IterStatus), s.stream));
    CU_TRY(cudaMemsetAsync(s.d_warn, 0, sizeof(int), s.stream));
    CU_TRY(cudaMemsetAsync(s.d_underflow, 0, sizeof(int), s.stream));
    memset(s.h_status, 0, sizeof(IterStatus));
    ScatterRagArgs ra = {J, K, pl.KP, s.d_ncat, s.d_rowoff, s.d_ragoff, s.d_rag_a, s.d_phi, 0, 0.0};
    k_ragged_to_table<<<jk_grid, 128, 0, s.stream>>>(ra);
    // table T (+ unit table T2) from phi_init, prior of the first iteration from colsum_init
    k_iter<<<s.n_elem, kIterThreads, 0, s.stream>>>(iter_args(s, 0, false));
    CU_TRY(cudaGetLastError());
  }

  // ---- the coordinate-ascent loop (R/mixdir_vi.R:58-115) ----
  // The host never waits for an iteration: it keeps up to kDepth iterations queued and reads the
  // status word k_iter mirrors into mapped page-locked memory.  Once the device has set `stop`
  // (converged / underflow / invalid codes) the kernels still in the queue return at once.
  constexpr int kDepth = 8;
  int iters = 0, conv = 0, warn = 0, status = MIXDIR_B200_OK, launches = 0;
  double last_elbo = NAN;
  for (auto& s : h->shards) {
    CU_TRY(cudaSetDevice(s.device));
    CU_TRY(cudaEventRecord(s.ev_loop0, s.stream));
  }
  std::vector<void*> stat_bufs(h->shards.size());
  Shard& s0 = h->shards[0];
  while ((int)s0.ev_ring.size() < kDepth) {
    cudaEvent_t e;
    CU_TRY(cudaSetDevice(s0.device));
    CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s0.ev_ring.push_back(e);
  }
  std::vector<int> iter_launches;
  for (int it = 0; it < p->max_iter; it++) {
    if (it >= kDepth) {
      CU_TRY(cudaEventSynchronize(s0.ev_ring[it % kDepth]));  // iteration it - kDepth is done
      if (((volatile IterStatus*)s0.h_status)->stop) break;
    }
    int l0 = 0;
    for (size_t si = 0; si < h->shards.size(); si++) {
      Shard& s = h->shards[si];
      int l = 0;
      RET_IF(launch_em(h, s, K, pl, nullptr, true, h->time_em ? s.ev_em[2 * it] : nullptr,
                       h->time_em ? s.ev_em[2 * it + 1] : nullptr, &l));
      stat_bufs[si] = s.d_stats;
      if (si == 0) l0 += l;
    }
    RET_IF(allreduce_stats(h, stat_bufs, tab, pl));
    for (size_t si = 0; si < h->shards.size(); si++) {
      Shard& s = h->shards[si];
      CU_TRY(cudaSetDevice(s.device));
      k_iter<<<s.n_elem, kIterThreads, 0, s.stream>>>(iter_args(s, 1, si == 0));
      CU_TRY(cudaGetLastError());
      if (si == 0) l0 += 1;
    }
    CU_TRY(cudaSetDevice(s0.device));
    CU_TRY(cudaEventRecord(s0.ev_ring[it % kDepth], s0.stream));
    iter_launches.push_back(l0);
  }
  for (auto& s : h->shards) {
    CU_TRY(cudaSetDevice(s.device));
    CU_TRY(cudaEventRecord(s.ev_loop1, s.stream));
  }
  for (auto& s : h->shards) {
    CU_TRY(cudaSetDevice(s.device));
    CU_TRY(cudaStreamSynchronize(s.stream));
  }
  {
    CU_TRY(cudaSetDevice(s0.device));
    const IterStatus fin = *s0.h_status;
    iters = fin.iter;
    conv = fin.converged;
    warn = fin.warn;
    if (iters > 0) last_elbo = fin.elbo;
    if (fin.bad) status = MIXDIR_B200_BAD_ARGUMENT;  // the upload scan found an invalid code
    else if (fin.underflow) status = MIXDIR_B200_UNDERFLOW;
    for (int it = 0; it < iters; it++) launches += iter_launches[it];
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, s0.ev_loop0, s0.ev_loop1));
    h->timing.fit_loop_ms = ms;
    double em = 0;
    for (int it = 0; h->time_em && it < iters; it++) {
      CU_TRY(cudaEventElapsedTime(&ms, s0.ev_em[2 * it], s0.ev_em[2 * it + 1]));
      em += ms;
    }
    h->timing.em_kernel_ms = em;
    h->timing.iterations = iters;
    h->timing.em_launches = iters;
    h->timing.total_launches = launches;
    if (elbo_trace && iters > 0)
      CU_TRY(cudaMemcpy(elbo_trace, s0.d_trace, iters * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (n_iter) *n_iter = iters;
  if (converged) *converged = conv;
  if (elbo) *elbo = last_elbo;
  if (status == MIXDIR_B200_BAD_ARGUMENT) {
    if (warn_flags) *warn_flags = 0;
    return bad_codes();
  }
  if (status != MIXDIR_B200_OK) {
    if (warn_flags) *warn_flags = warn;
    return fail(status, "zeta underflow: some row has rowSums(zeta) == 0");
  }
  RET_IF(resolve_scan(h));  // upload complete by now; validates the codes when max_iter == 0

  // ---- results (R/mixdir_vi.R:119-155) ----
  for (size_t si = 0; si < h->shards.size(); si++) {
    Shard& s = h->shards[si];
    CU_TRY(cudaSetDevice(s.device));
    ResultArgs ra = {J, K, pl.KP, p->dp, p->alpha1, p->alpha2, s.d_ncat, s.d_rowoff, s.d_ragoff,
                     s.d_phi, s.d_cs_a, s.d_logU, s.d_rag_a, s.d_rag_b, s.d_lambda,
                     s.d_lambda_pad, s.d_prior_out, s.d_warn};
    k_results<<<jk_grid, 128, 0, s.stream>>>(ra);
    CU_TRY(cudaGetLastError());
    if (si == 0) {
      int w2 = 0;
      if (lambda) CU_TRY(cudaMemcpyAsync(lambda, s.d_lambda, K * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
      if (prior_param)
        CU_TRY(cudaMemcpyAsync(prior_param, s.d_prior_out, (p->dp ? 2 : 1) * K * sizeof(double),
                               cudaMemcpyDeviceToHost, s.stream));
      if (phi) CU_TRY(cudaMemcpyAsync(phi, s.d_rag_a, rag * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
      if (category_prob)
        CU_TRY(cudaMemcpyAsync(category_prob, s.d_rag_b, rag * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
      CU_TRY(cudaMemcpyAsync(&w2, s.d_warn, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
      CU_TRY(cudaStreamSynchronize(s.stream));
      warn |= w2;
    }
  }
  if (warn_flags) *warn_flags = warn;
  if (class_prob || pred_class) RET_IF(run_predict(h, K, pl.KP, class_prob, pred_class));
  return MIXDIR_B200_OK;
}

// A handle that SHARES the uploaded data of `h` (one shard, upload resolved) and owns a stream and
// a workspace: one concurrent fit of mixdir_b200_fit_batch.
static int make_view(mixdir_b200_handle* h, mixdir_b200_handle** out) {
  const Shard& ps = h->shards[0];
  mixdir_b200_handle* v = new mixdir_b200_handle();
  v->J = h->J; v->CB = h->CB; v->row_bytes = h->row_bytes; v->WPR = h->WPR; v->NR = h->NR;
  v->n_rows = h->n_rows; v->n_missing = h->n_missing; v->max_code = h->max_code;
  v->sum_ncat = h->sum_ncat;
  v->ncat = h->ncat; v->rowoff = h->rowoff; v->rowoff_pad = h->rowoff_pad; v->ragoff1 = h->ragoff1;
  v->has_na = h->has_na;
  v->engine_pref = h->engine_pref; v->pairing = h->pairing;
  v->scan_resolved = true; v->bad = h->bad;
  v->shards.resize(1);
  Shard& s = v->shards[0];
  s.view = true;
  s.device = ps.device; s.row0 = ps.row0; s.n_rows = ps.n_rows;
  s.sm_count = ps.sm_count; s.smem_optin = ps.smem_optin;
  s.d_codes = ps.d_codes; s.d_ncat = ps.d_ncat; s.d_rowoff = ps.d_rowoff;
  s.d_rowoff_pad = ps.d_rowoff_pad; s.d_scan = ps.d_scan; s.d_na_feat = ps.d_na_feat;
  auto body = [&]() -> int {
    CU_TRY(cudaSetDevice(s.device));
    CU_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreate(&s.ev_loop0));
    CU_TRY(cudaEventCreate(&s.ev_loop1));
    CU_TRY(cache::pin_get((void**)&s.h_status, sizeof(IterStatus)));
    RET_IF(dalloc(&s.d_ragoff, (size_t)h->J));
    return MIXDIR_B200_OK;
  };
  const int st = body();
  if (st != MIXDIR_B200_OK) {
    std::string keep = g_err;
    mixdir_b200_destroy(v);
    g_err = keep;
    return st;
  }
  *out = v;
  return MIXDIR_B200_OK;
}

// concurrent fits of one batch: small fits are bound by launch latency and the dependent chain of
// short kernels (tens of microseconds per iteration on a handful of SMs), so several of them share
// the GPU without slowing each other; a fit whose row pass fills the GPU gains nothing
static int batch_streams(const mixdir_b200_handle* h, int n_fits) {
  int c = (int)std::min<int64_t>(8, std::max<int64_t>(1, ((int64_t)1 << 21) / std::max<int64_t>(h->n_rows, 1)));
  if (const char* e = getenv("MIXDIR_B200_BATCH_STREAMS")) {
    const int t = atoi(e);
    if (t >= 1) c = std::min(t, 32);
  }
  if (needs_allreduce(h)) c = 1;  // collectives on one communicator must stay in one order
  return std::max(1, std::min(c, n_fits));
}

extern "C" int mixdir_b200_fit_batch(mixdir_b200_handle* h, const mixdir_b200_fit_params* p, int n_fits,
                                     const double* colsum_init, const double* phi_init,
                                     double* elbo_trace, int* n_iter, int* converged, double* elbo,
                                     double* lambda, double* prior_param, double* phi,
                                     double* category_prob, int* warn_flags, int* status,
                                     int* best, double* class_prob, int32_t* pred_class) {
  RET_IF(check_fit_args(h, p));
  if (n_fits < 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "n_fits must be >= 1");
  if (!colsum_init || !phi_init) return fail(MIXDIR_B200_BAD_ARGUMENT, "colsum_init / phi_init is NULL");
  if (!elbo || !lambda || !category_prob)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "elbo, lambda and category_prob are needed to select the best fit");
  DeviceGuard guard;
  const int K = p->n_latent;
  const size_t rag = (size_t)h->sum_ncat * K;
  const size_t nprior = (size_t)(p->dp ? 2 : 1) * K;
  const int C = batch_streams(h, n_fits);
  if (C > 1) {
    RET_IF(resolve_scan(h));  // the views share a completed upload
    while ((int)h->views.size() < C - 1) {
      mixdir_b200_handle* v = nullptr;
      RET_IF(make_view(h, &v));
      h->views.push_back(v);
    }
    for (auto* v : h->views) {
      v->engine_pref = h->engine_pref;
      v->pairing = h->pairing;
      v->time_em = false;  // (two event records per iteration and fit: host API calls the workers contend for)
    }
  }
  h->time_em = C == 1;
  std::vector<int> st(n_fits, MIXDIR_B200_OK);
  std::vector<std::string> msg(n_fits);
  std::vector<mixdir_b200_timing> tms(C);
  auto worker = [&](int w) {
    mixdir_b200_handle* hw = w == 0 ? h : h->views[w - 1];
    double loop_ms = 0, em_ms = 0;
    int iters = 0, eml = 0, tl = 0;
    for (int i = w; i < n_fits; i += C) {
      st[i] = mixdir_b200_fit(hw, p, colsum_init + (size_t)i * K, phi_init + (size_t)i * rag,
                              elbo_trace ? elbo_trace + (size_t)i * std::max(p->max_iter, 1) : nullptr,
                              n_iter ? n_iter + i : nullptr, converged ? converged + i : nullptr,
                              elbo + i, lambda + (size_t)i * K,
                              prior_param ? prior_param + (size_t)i * nprior : nullptr,
                              phi ? phi + (size_t)i * rag : nullptr, category_prob + (size_t)i * rag,
                              nullptr, nullptr, warn_flags ? warn_flags + i : nullptr);
      if (st[i] != MIXDIR_B200_OK) msg[i] = g_err;
      loop_ms += hw->timing.fit_loop_ms; em_ms += hw->timing.em_kernel_ms;
      iters += hw->timing.iterations; eml += hw->timing.em_launches; tl += hw->timing.total_launches;
    }
    tms[w] = hw->timing;
    tms[w].fit_loop_ms = loop_ms; tms[w].em_kernel_ms = em_ms;
    tms[w].iterations = iters; tms[w].em_launches = eml; tms[w].total_launches = tl;
  };
  if (C == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 1; w < C; w++) pool.emplace_back(worker, w);
    worker(0);
    for (auto& t : pool) t.join();
  }
  // timing of the batch: the engine description of worker 0, sums over every fit of the batch
  // (fit_loop_ms therefore counts concurrent loops once each, not the batch's elapsed time)
  h->timing = tms[0];
  for (int w = 1; w < C; w++) {
    h->timing.fit_loop_ms += tms[w].fit_loop_ms; h->timing.em_kernel_ms += tms[w].em_kernel_ms;
    h->timing.iterations += tms[w].iterations; h->timing.em_launches += tms[w].em_launches;
    h->timing.total_launches += tms[w].total_launches;
  }
  h->time_em = true;
  if (status) std::copy(st.begin(), st.end(), status);
  // an error in any repetition ends mixdir() (stop() inside mixdir_vi, R/mixdir.R:120-134)
  for (int i = 0; i < n_fits; i++)
    if (st[i] != MIXDIR_B200_OK) {
      if (best) *best = -1;
      return fail(st[i], msg[i]);
    }
  // `if (result$ELBO < result_tmp$ELBO) result <- result_tmp` (R/mixdir.R:129-133): first maximum
  int b = 0;
  for (int i = 1; i < n_fits; i++)
    if (elbo[b] < elbo[i]) b = i;
  if (best) *best = b;
  if (class_prob || pred_class)  // the arithmetic of the final pass == predict_class (R/predict.R:88-103)
    return mixdir_b200_predict(h, K, lambda + (size_t)b * K, category_prob + (size_t)b * rag, class_prob,
                               pred_class);
  return MIXDIR_B200_OK;
}

// ---------------------------------------------------------------------------
// large results to (pageable) host memory: device -> page-locked staging in 64 MB pieces, each piece
// copied out by a few host threads while the next one crosses PCIe.  (cudaMemcpy to pageable memory
// bounces through the driver at 3-4 GB/s: 0.68 s for the 2.56 GB class_prob of 10M x 32.)
// Copies `height` runs of `width` bytes; the runs are dst_pitch / src_pitch bytes apart.
// ---------------------------------------------------------------------------
static void host_copy_parallel(unsigned char* dst, const unsigned char* src, size_t bytes, int T) {
  if (bytes < ((size_t)4 << 20) || T <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = ((bytes + T - 1) / T + 4095) & ~(size_t)4095;
  for (int t = 0; t < T; t++) {
    const size_t o = (size_t)t * per;
    if (o >= bytes) break;
    pool.emplace_back([=]() { memcpy(dst + o, src + o, std::min(per, bytes - o)); });
  }
  for (auto& th : pool) th.join();
}

static int d2h_bulk(Shard& s, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width,
                    size_t height) {
  const size_t total = width * height;
  if (total == 0) return MIXDIR_B200_OK;
  if (total < ((size_t)8 << 20)) {
    CU_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, height, cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    return MIXDIR_B200_OK;
  }
  const size_t kPiece = (size_t)64 << 20;
  // pieces: whole runs when a run is smaller than a piece, else slices of one run
  struct Piece { size_t run, off, len, nruns; };
  std::vector<Piece> pieces;
  if (width >= kPiece) {
    for (size_t r = 0; r < height; r++)
      for (size_t o = 0; o < width; o += kPiece) pieces.push_back({r, o, std::min(kPiece, width - o), 1});
  } else {
    const size_t per = std::max<size_t>(1, kPiece / width);
    for (size_t r = 0; r < height; r += per) pieces.push_back({r, 0, width, std::min(per, height - r)});
  }
  unsigned char* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  const int T = std::min(host_pack_threads(), 8);
  auto body = [&]() -> int {
    for (int b = 0; b < 2; b++) {
      CU_TRY(cache::pin_get((void**)&stage[b], kPiece));
      CU_TRY(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
    }
    auto issue = [&](size_t i) -> int {
      const Piece& p = pieces[i];
      const unsigned char* sp = (const unsigned char*)src + p.run * src_pitch + p.off;
      if (p.nruns == 1)
        CU_TRY(cudaMemcpyAsync(stage[i & 1], sp, p.len, cudaMemcpyDeviceToHost, s.stream));
      else
        CU_TRY(cudaMemcpy2DAsync(stage[i & 1], width, sp, src_pitch, width, p.nruns, cudaMemcpyDeviceToHost, s.stream));
      CU_TRY(cudaEventRecord(ev[i & 1], s.stream));
      return MIXDIR_B200_OK;
    };
    RET_IF(issue(0));
    for (size_t i = 0; i < pieces.size(); i++) {
      if (i + 1 < pieces.size()) RET_IF(issue(i + 1));  // its buffer was emptied in the previous round
      CU_TRY(cudaEventSynchronize(ev[i & 1]));
      const Piece& p = pieces[i];
      unsigned char* dp = (unsigned char*)dst + p.run * dst_pitch + p.off;
      if (p.nruns == 1 || dst_pitch == width) {
        host_copy_parallel(dp, stage[i & 1], p.len * p.nruns, T);
      } else {
        for (size_t r = 0; r < p.nruns; r++) host_copy_parallel(dp + r * dst_pitch, stage[i & 1] + r * width, width, T);
      }
    }
    return MIXDIR_B200_OK;
  };
  const int st = body();
  if (st != MIXDIR_B200_OK) cudaStreamSynchronize(s.stream);
  for (int b = 0; b < 2; b++) {
    if (ev[b]) cudaEventDestroy(ev[b]);
    if (stage[b]) cache::pin_put(stage[b]);
  }
  return st;
}

// ---------------------------------------------------------------------------
// final pass / predict
// ---------------------------------------------------------------------------
typedef void (*predu_fn)(PredictUnitArgs);
static predu_fn pick_predict_units(int KC) {
  return KC == 8 ? k_predict_units<8> : (KC == 16 ? k_predict_units<16> : k_predict_units<32>);
}

// unit plan of the predict kernel for K classes: as many feature pairs as shared memory allows next
// to the tile buffers; classes per CTA and cluster size by a cost in shared-memory wavefronts per row
static bool plan_predict(const mixdir_b200_handle* h, const Shard& s, int K, PredPlan* out) {
  if (h->CB != 1 || getenv("MIXDIR_B200_OLD_PREDICT")) return false;
  const uint8_t* na = (h->scan_resolved && (int)h->has_na.size() == h->J) ? h->has_na.data() : nullptr;
  PredPlan best;
  double best_cost = 1e300;
  const int kcs[3] = {32, 16, 8};
  for (int KC : kcs) {
    if (KC > 8 && K <= KC / 2) continue;  // a narrower slice holds all classes
    const int P = (K + KC - 1) / KC;
    if (P > 8) continue;
    PredPlan c;
    c.K = K; c.KC = KC; c.P = P; c.W = 16; c.R = 256;
    bool found = false;
    const int m_hi = h->pairing ? h->J / 2 : 0;
    const int w_lo = (h->J - m_hi + 3) / 4, w_hi = (h->J + 3) / 4;
    for (int w = w_lo; w <= w_hi && !found; w++) {
      const int m = std::max(0, h->J - 4 * w);
      if (m > m_hi || !make_units(h, m, &c.up, true, na)) continue;
      layout_units(&c.up, KC);
      c.smem = PredictUnitSmem(c.up.srows, KC, c.R, c.up.WPRu, P).total;
      found = c.smem <= (size_t)s.smem_optin;
    }
    if (!found) continue;
    const int n_clusters = s.sm_count / P;
    const double used = (double)n_clusters * P / s.sm_count;
    const double cost = P * (c.up.U * (KC / 16.0) + (KC == 8 ? 40.0 : 60.0)) / used;
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
      best.valid = true;
    }
  }
  if (!best.valid) return false;
  *out = best;
  return true;
}

// device copies of the predict plan + the packed codes of this shard (kept while the plan is the same)
static int ensure_predict_units(mixdir_b200_handle* h, Shard& s, const PredPlan& pp) {
  const UnitPlan& up = pp.up;
  std::vector<UnitDesc> sig = up.desc;
  sig.push_back(UnitDesc{pp.KC, up.Q, up.srows, pp.R});
  if (s.pred_valid && s.pred_sig.size() == sig.size() &&
      memcmp(s.pred_sig.data(), sig.data(), sig.size() * sizeof(UnitDesc)) == 0)
    return MIXDIR_B200_OK;
  s.pred_valid = false;
  CU_TRY(cudaStreamSynchronize(s.stream));  // the staging area is reused below
  dfree(s.d_pdesc); dfree(s.d_pt2src); dfree(s.d_pslot); dfree(s.d_peoff); dfree(s.d_pcodes);
  RET_IF(pin_reset(s, unit_stage_bytes(up) + 4096));
  RET_IF(send_vec(s, &s.d_pdesc, up.desc));
  RET_IF(send_vec(s, &s.d_pt2src, up.t2src));
  RET_IF(send_vec(s, &s.d_pslot, up.slot));
  RET_IF(send_vec(s, &s.d_peoff, up.eoff_rot));
  const size_t rowb = (size_t)up.WPRu * 4;
  const int64_t R = pp.R, rows_alloc = (s.n_rows + R - 1) / R * R + kRowPad;
  RET_IF(dalloc(&s.d_pcodes, (size_t)rows_alloc * rowb));
  const int64_t full = s.n_rows / R * R;
  CU_TRY(cudaMemsetAsync(s.d_pcodes + (size_t)full * rowb, 0, (size_t)(rows_alloc - full) * rowb, s.stream));
  if (s.n_rows > 0) {
    const int grid = (int)std::min<int64_t>((int64_t)s.sm_count * 16, (s.n_rows * up.WPRu + 255) / 256);
    k_pack_units<<<grid, 256, 0, s.stream>>>(s.d_codes, h->row_bytes, s.n_rows, s.d_pdesc, up.WPRu, pp.R,
                                             reinterpret_cast<uint32_t*>(s.d_pcodes));
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaStreamSynchronize(s.stream));  // the staged arrays have been read
  s.pred_sig = sig;
  s.pred_valid = true;
  return MIXDIR_B200_OK;
}

// final pass on every shard with the tables already in d_logU (row stride KP) / d_lambda_pad
static int run_predict(mixdir_b200_handle* h, int K, int KP, double* class_prob, int32_t* pred_class) {
  h->timing.final_pass_kernel_ms = 0;
  h->timing.final_pass_engine = 0;
  for (auto& s : h->shards) {
    if (s.n_rows == 0) continue;
    CU_TRY(cudaSetDevice(s.device));
    double* d_cp = nullptr;
    int32_t* d_pc = nullptr;
    if (class_prob) RET_IF(dalloc(&d_cp, (size_t)s.n_rows * K));
    if (pred_class) RET_IF(dalloc(&d_pc, (size_t)s.n_rows));
    if (!s.ev_p0) {
      CU_TRY(cudaEventCreate(&s.ev_p0));
      CU_TRY(cudaEventCreate(&s.ev_p1));
    }
    int engine = 0;
    PredPlan pp;
    if (plan_predict(h, s, K, &pp)) {  // packed unit codes + pair table of log U in shared memory
      RET_IF(ensure_predict_units(h, s, pp));
      const size_t tab2 = (size_t)pp.up.NR2 * KP;
      if (s.pT2_cap < tab2) {
        dfree(s.d_pT2);
        RET_IF(dalloc(&s.d_pT2, tab2));
        s.pT2_cap = tab2;
      }
      predu_fn fn = pick_predict_units(pp.KC);
      CU_TRY(ensure_smem_attr(s, (const void*)fn, pp.smem));
      CU_TRY(cudaEventRecord(s.ev_p0, s.stream));
      const int g2 = (int)std::min<size_t>((tab2 + 255) / 256, (size_t)s.sm_count * 4);
      k_build_T2<<<g2, 256, 0, s.stream>>>(s.d_pt2src, pp.up.NR2, KP, s.d_logU, s.d_pT2);
      CU_TRY(cudaGetLastError());
      PredictUnitArgs a;
      a.codes = reinterpret_cast<const uint32_t*>(s.d_pcodes);
      a.n_rows = s.n_rows;
      a.n_tiles = (int)((s.n_rows + pp.R - 1) / pp.R);
      a.R = pp.R;
      a.WPR = pp.up.WPRu;
      a.NR2 = pp.up.NR2;
      a.srows = pp.up.srows;
      a.K = K;
      a.KP = KP;
      a.P = pp.P;
      a.eoff_rot = s.d_peoff;
      a.slot_of_row = s.d_pslot;
      a.T2 = s.d_pT2;
      a.lambda = s.d_lambda_pad;
      a.class_prob = d_cp;
      a.pred_class = d_pc;
      const int nclus = std::max(1, std::min(s.sm_count / pp.P, a.n_tiles));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nclus * pp.P);
      cfg.blockDim = dim3(32 * pp.W);
      cfg.dynamicSmemBytes = pp.smem;
      cfg.stream = s.stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = pp.P;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      CU_TRY(cudaLaunchKernelEx(&cfg, fn, a));
      engine = 6;
    } else {
      // 2-byte codes: log-U table in shared memory when K <= 32 and it fits, else the general kernel
      const int KC = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
      const int FPW = 4 / h->CB, W = 16, RPW = 32 / KC;
      const PredictSmem L(h->NR, KC, W, 8, h->WPR, FPW);
      typedef void (*pfn)(PredictSmemArgs);
      pfn fn = nullptr;
      if (K <= 32 && RPW <= FPW * 8 && L.total <= (size_t)s.smem_optin) {
        if (h->CB == 1) fn = KC == 8 ? k_predict_smem<8, 1> : (KC == 16 ? k_predict_smem<16, 1> : k_predict_smem<32, 1>);
        else fn = KC == 8 ? k_predict_smem<8, 2> : (KC == 16 ? k_predict_smem<16, 2> : k_predict_smem<32, 2>);
        if (ensure_smem_attr(s, (const void*)fn, L.total) != cudaSuccess) {
          cudaGetLastError();
          fn = nullptr;
        }
      }
      CU_TRY(cudaEventRecord(s.ev_p0, s.stream));
      if (fn) {
        const int R = W * 8 * RPW;
        PredictSmemArgs a;
        a.codes = reinterpret_cast<const uint32_t*>(s.d_codes);
        a.n_rows = s.n_rows;
        a.n_tiles = (int)((s.n_rows + R - 1) / R);
        a.WPR = h->WPR;
        a.NR = h->NR;
        a.K = K;
        a.KP = KP;
        a.rowoff_pad = s.d_rowoff_pad;
        a.logU = s.d_logU;
        a.lambda = s.d_lambda_pad;
        a.class_prob = d_cp;
        a.pred_class = d_pc;
        const int grid = std::min(s.sm_count, a.n_tiles);
        fn<<<grid, 32 * W, L.total, s.stream>>>(a);
        engine = 7;
      } else {
        PredictArgs a = {s.d_codes, h->row_bytes, s.n_rows, h->J, K, KP, s.d_rowoff, s.d_logU,
                         s.d_lambda_pad, d_cp, d_pc};
        const int grid = s.sm_count * 8;
        if (h->CB == 1)
          k_predict<1><<<grid, 256, 0, s.stream>>>(a);
        else
          k_predict<2><<<grid, 256, 0, s.stream>>>(a);
        engine = 8;
      }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(s.ev_p1, s.stream));
    int st = MIXDIR_B200_OK;
    if (class_prob)  // shard block of an n_rows x K column-major matrix
      st = d2h_bulk(s, class_prob + s.row0, (size_t)h->n_rows * sizeof(double), d_cp,
                    (size_t)s.n_rows * sizeof(double), (size_t)s.n_rows * sizeof(double), (size_t)K);
    if (st == MIXDIR_B200_OK && pred_class)
      st = d2h_bulk(s, pred_class + s.row0, 0, d_pc, 0, (size_t)s.n_rows * sizeof(int32_t), 1);
    if (st == MIXDIR_B200_OK && cudaStreamSynchronize(s.stream) != cudaSuccess) st = fail(MIXDIR_B200_CUDA_ERROR, "final pass failed");
    dfree(d_cp);
    dfree(d_pc);
    RET_IF(st);
    if (&s == &h->shards[0]) {
      float ms = 0;
      CU_TRY(cudaEventElapsedTime(&ms, s.ev_p0, s.ev_p1));
      h->timing.final_pass_kernel_ms = ms;
      h->timing.final_pass_engine = engine;
    }
  }
  return MIXDIR_B200_OK;
}

extern "C" int mixdir_b200_predict(mixdir_b200_handle* h, int n_latent, const double* lambda,
                                   const double* category_prob, double* class_prob,
                                   int32_t* pred_class) {
  if (!h || !lambda || !category_prob) return fail(MIXDIR_B200_BAD_ARGUMENT, "NULL argument");
  const int K = n_latent;
  if (K < 1 || K > kMaxClasses) return fail(MIXDIR_B200_BAD_ARGUMENT, "n_latent must be in 1..256");
  DeviceGuard guard;
  RET_IF(resolve_scan(h));
  Plan pl;
  pl.engine = 1;
  pl.KP = (K + 7) & ~7;
  const size_t rag = (size_t)h->sum_ncat * K;
  const int jk_grid = (h->J * K + 127) / 128;
  std::vector<double> lpad(pl.KP, 0.0);
  for (int k = 0; k < K; k++) lpad[k] = lambda[k];
  for (auto& s : h->shards) {
    RET_IF(ensure_workspace(h, s, K, pl, 1));
    RET_IF(pin_reset(s, (size_t)h->J * 8 + 4096));
    RET_IF(upload_ragoff(h, s, K));
    CU_TRY(cudaMemcpyAsync(s.d_rag_a, category_prob, rag * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaMemcpyAsync(s.d_lambda_pad, lpad.data(), pl.KP * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    ScatterRagArgs ra = {h->J, K, pl.KP, s.d_ncat, s.d_rowoff, s.d_ragoff, s.d_rag_a, s.d_logU, 1, 0.0};
    k_ragged_to_table<<<jk_grid, 128, 0, s.stream>>>(ra);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s.stream));  // lpad is a local
  }
  return run_predict(h, K, pl.KP, class_prob, pred_class);
}

// adjusted Rand index of a K x K contingency table (Hubert & Arabie; what mcclust::arandi computes)
static double adjusted_rand(const unsigned long long* c, int K) {
  auto c2 = [](double x) { return x * (x - 1.0) / 2.0; };
  std::vector<double> a(K, 0.0), b(K, 0.0);
  double n = 0, sij = 0;
  for (int i = 0; i < K; i++)
    for (int j = 0; j < K; j++) {
      const double v = (double)c[(size_t)i * K + j];
      a[i] += v;
      b[j] += v;
      n += v;
      sij += c2(v);
    }
  double sa = 0, sb = 0;
  for (int i = 0; i < K; i++) {
    sa += c2(a[i]);
    sb += c2(b[i]);
  }
  const double correc = sa * sb / c2(n);
  return (sij - correc) / (0.5 * sa + 0.5 * sb - correc);
}

extern "C" int mixdir_b200_feature_divergence(mixdir_b200_handle* h, int n_latent, const double* lambda,
                                              const double* category_prob,
                                              const unsigned char* active, int measure,
                                              const int64_t* rows, int64_t n_rows_sub, double* out) {
  if (!h || !lambda || !category_prob || !active || !out)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "NULL argument");
  const int K = n_latent, J = h->J;
  if (K < 1 || K > kMaxClasses) return fail(MIXDIR_B200_BAD_ARGUMENT, "n_latent must be in 1..256");
  if (measure != 0 && measure != 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "measure must be 0 (JS) or 1 (ARI)");
  if (rows && n_rows_sub < 0) return fail(MIXDIR_B200_BAD_ARGUMENT, "negative subsample size");
  if (rows && h->world > 1)
    return fail(MIXDIR_B200_BAD_ARGUMENT, "row subsamples are not supported by multi-process handles");
  if (rows)
    for (int64_t i = 0; i < n_rows_sub; i++)
      if (rows[i] < 0 || rows[i] >= h->n_rows) return fail(MIXDIR_B200_BAD_ARGUMENT, "subsample row out of range");
  const size_t ncont = (size_t)(J + 1) * K * K;
  if (measure == 1 && ncont > ((size_t)1 << 27))
    return fail(MIXDIR_B200_BAD_ARGUMENT, "ARI contingency tables too large ((J+1) * K^2 > 2^27)");
  DeviceGuard guard;
  RET_IF(resolve_scan(h));
  Plan pl;
  pl.engine = 1;
  pl.KP = (K + 7) & ~7;
  const size_t rag = (size_t)h->sum_ncat * K;
  const int jk_grid = (J * K + 127) / 128;
  std::vector<double> lpad(pl.KP, 0.0);
  for (int k = 0; k < K; k++) lpad[k] = lambda[k];
  const size_t nout = (size_t)J + 1;
  std::vector<void*> bufs(h->shards.size());
  struct Tmp { unsigned char* act = nullptr; int64_t* rows = nullptr; double* partial = nullptr;
               double* res = nullptr; unsigned long long* cont = nullptr; };
  std::vector<Tmp> tmp(h->shards.size());
  auto cleanup = [&]() {
    for (auto& t : tmp) { dfree(t.act); dfree(t.rows); dfree(t.partial); dfree(t.res); dfree(t.cont); }
  };
  int status = MIXDIR_B200_OK;
  for (size_t si = 0; si < h->shards.size() && status == MIXDIR_B200_OK; si++) {
    Shard& s = h->shards[si];
    Tmp& t = tmp[si];
    status = [&]() -> int {
      RET_IF(ensure_workspace(h, s, K, pl, 1));
      RET_IF(pin_reset(s, (size_t)J * 8 + 4096));
      RET_IF(upload_ragoff(h, s, K));
      CU_TRY(cudaMemcpyAsync(s.d_rag_a, category_prob, rag * sizeof(double), cudaMemcpyHostToDevice, s.stream));
      CU_TRY(cudaMemcpyAsync(s.d_lambda_pad, lpad.data(), pl.KP * sizeof(double), cudaMemcpyHostToDevice, s.stream));
      ScatterRagArgs ra = {J, K, pl.KP, s.d_ncat, s.d_rowoff, s.d_ragoff, s.d_rag_a, s.d_logU, 1, 0.0};
      k_ragged_to_table<<<jk_grid, 128, 0, s.stream>>>(ra);
      CU_TRY(cudaGetLastError());
      // one warp per row; the per-warp accumulators [J+1] live in shared memory
      int warps = (int)std::min<size_t>(8, (size_t)48 * 1024 / (nout * sizeof(double)));
      if (warps < 1) return fail(MIXDIR_B200_BAD_ARGUMENT, "too many features for feature_divergence (J > 6143)");
      const int64_t n = rows ? n_rows_sub : s.n_rows;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)s.sm_count * 8, (n + warps - 1) / warps));
      RET_IF(dalloc(&t.act, (size_t)J));
      RET_IF(dalloc(&t.partial, (size_t)grid * nout));
      RET_IF(dalloc(&t.res, nout));
      CU_TRY(cudaMemcpyAsync(t.act, active, J, cudaMemcpyHostToDevice, s.stream));
      if (rows) {
        RET_IF(dalloc(&t.rows, (size_t)std::max<int64_t>(n_rows_sub, 1)));
        CU_TRY(cudaMemcpyAsync(t.rows, rows, n_rows_sub * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
      }
      if (measure == 1) {
        RET_IF(dalloc(&t.cont, ncont));
        CU_TRY(cudaMemsetAsync(t.cont, 0, ncont * sizeof(unsigned long long), s.stream));
      }
      FeatureDivArgs a = {s.d_codes, h->row_bytes, s.n_rows, s.row0, t.rows, n_rows_sub, J, K, pl.KP,
                          s.d_rowoff, s.d_logU, s.d_lambda_pad, t.act, measure, t.partial, t.cont};
      const size_t smem = (size_t)warps * nout * sizeof(double);
      if (s.n_rows > 0) {
        if (h->CB == 1)
          k_feature_div<1><<<grid, 32 * warps, smem, s.stream>>>(a);
        else
          k_feature_div<2><<<grid, 32 * warps, smem, s.stream>>>(a);
        CU_TRY(cudaGetLastError());
        k_feature_div_reduce<<<(int)((nout + 127) / 128), 128, 0, s.stream>>>(t.partial, grid, (int)nout, t.res);
        CU_TRY(cudaGetLastError());
      } else {
        CU_TRY(cudaMemsetAsync(t.res, 0, nout * sizeof(double), s.stream));
      }
      bufs[si] = measure == 1 ? (void*)t.cont : (void*)t.res;
      CU_TRY(cudaStreamSynchronize(s.stream));  // lpad / host inputs are read by the copies above
      return MIXDIR_B200_OK;
    }();
  }
  if (status == MIXDIR_B200_OK)
    status = measure == 1 ? allreduce(h, bufs, ncont, nccl::kInt64, nccl::kSum)
                          : allreduce(h, bufs, nout, nccl::kDouble, nccl::kSum);
  if (status == MIXDIR_B200_OK) {
    status = [&]() -> int {
      Shard& s = h->shards[0];
      CU_TRY(cudaSetDevice(s.device));
      if (measure == 0) {
        CU_TRY(cudaMemcpyAsync(out, tmp[0].res, nout * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        CU_TRY(cudaStreamSynchronize(s.stream));
      } else {
        std::vector<unsigned long long> cont(ncont);
        CU_TRY(cudaMemcpyAsync(cont.data(), tmp[0].cont, ncont * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, s.stream));
        CU_TRY(cudaStreamSynchronize(s.stream));
        for (int i = 0; i <= J; i++) out[i] = 1.0 - adjusted_rand(cont.data() + (size_t)i * K * K, K);
      }
      for (int i = 0; i < J; i++)
        if (!active[i]) out[i] = NAN;
      return MIXDIR_B200_OK;
    }();
  }
  cleanup();
  return status;
}

extern "C" int mixdir_b200_estep(mixdir_b200_handle* h, int n_latent, const double* log_prior,
                                 const double* table, int engine, double* S, double* colsum,
                                 double* entropy, int* underflow, double* zeta) {
  if (!h || !log_prior || !table) return fail(MIXDIR_B200_BAD_ARGUMENT, "NULL argument");
  const int K = n_latent;
  if (K < 1 || K > kMaxClasses) return fail(MIXDIR_B200_BAD_ARGUMENT, "n_latent must be in 1..256");
  if (engine < 0 || engine > 5 || engine == 3) return fail(MIXDIR_B200_BAD_ARGUMENT, "bad engine");
  DeviceGuard guard;
  RET_IF(resolve_scan(h));
  const int keep = h->engine_pref;
  if (engine) h->engine_pref = engine;
  Plan pl;
  int st = plan_engine(h, h->shards[0], K, &pl);
  h->engine_pref = keep;
  RET_IF(st);
  const size_t tab = (size_t)h->NR * pl.KP;
  const size_t rag = (size_t)h->sum_ncat * K;
  const int jk_grid = (h->J * K + 127) / 128;
  std::vector<double> lp(pl.KP, -INFINITY);
  for (int k = 0; k < K; k++) lp[k] = log_prior[k];
  std::vector<void*> bufs(h->shards.size());
  for (size_t si = 0; si < h->shards.size(); si++) {
    Shard& s = h->shards[si];
    RET_IF(ensure_workspace(h, s, K, pl, 1));
    RET_IF(pin_reset(s, (size_t)h->J * 8 + 4096 + unit_stage_bytes(pl.up) + seg_stage_bytes(pl.sp)));
    RET_IF(upload_ragoff(h, s, K));
    RET_IF(ensure_units(h, s, pl));
    CU_TRY(cudaMemcpyAsync(s.d_rag_a, table, rag * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaMemcpyAsync(s.d_logprior, lp.data(), pl.KP * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaMemsetAsync(s.d_status, 0, sizeof(IterStatus), s.stream));
    CU_TRY(cudaMemsetAsync(s.d_underflow, 0, sizeof(int), s.stream));
    ScatterRagArgs ra = {h->J, K, pl.KP, s.d_ncat, s.d_rowoff, s.d_ragoff, s.d_rag_a, s.d_T, 0, 0.0};
    k_ragged_to_table<<<jk_grid, 128, 0, s.stream>>>(ra);
    CU_TRY(cudaGetLastError());
    RET_IF(build_unit_table(s, pl));
    CU_TRY(cudaStreamSynchronize(s.stream));
    int l = 0;
    double* d_zeta = nullptr;
    if (zeta && s.n_rows > 0) {
      RET_IF(dalloc(&d_zeta, (size_t)s.n_rows * K));
      if (pl.engine >= 2) {  // responsibilities come from the generic kernel (no scatter)
        Plan g = pl;
        g.engine = 1;
        RET_IF(launch_em(h, s, K, g, d_zeta, false, nullptr, nullptr, &l));
      }
    }
    // k_em_seg ADDS into the statistics (and the generic pass above has used the buffer)
    if (pl.engine == 5) CU_TRY(cudaMemsetAsync(s.d_stats, 0, (tab + 2 * pl.KP + 4) * sizeof(double), s.stream));
    RET_IF(launch_em(h, s, K, pl, pl.engine == 1 ? d_zeta : nullptr, true, nullptr, nullptr, &l));
    if (d_zeta) {
      CU_TRY(cudaMemcpy2DAsync(zeta + s.row0, (size_t)h->n_rows * sizeof(double), d_zeta,
                               (size_t)s.n_rows * sizeof(double), (size_t)s.n_rows * sizeof(double),
                               K, cudaMemcpyDeviceToHost, s.stream));
      CU_TRY(cudaStreamSynchronize(s.stream));
      dfree(d_zeta);
    }
    bufs[si] = s.d_stats;
  }
  RET_IF(allreduce_stats(h, bufs, tab, pl));
  {
    Shard& s = h->shards[0];
    CU_TRY(cudaSetDevice(s.device));
    if (pl.engine == 5) {
      k_fixed_to_double<<<64, 256, 0, s.stream>>>(s.d_stats, tab);
      CU_TRY(cudaGetLastError());
    }
    ScatterRagArgs ra = {h->J, K, pl.KP, s.d_ncat, s.d_rowoff, s.d_ragoff, nullptr, s.d_stats, 0, 0.0};
    k_table_to_ragged<<<jk_grid, 128, 0, s.stream>>>(ra, s.d_rag_b);
    CU_TRY(cudaGetLastError());
    const int csw = pl.engine == 5 ? 2 : 1;  // k_em_seg: two integer words per class mass (seg.cuh)
    std::vector<double> tail(csw * pl.KP + 2);
    if (S) CU_TRY(cudaMemcpyAsync(S, s.d_rag_b, rag * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaMemcpyAsync(tail.data(), s.d_stats + tab, tail.size() * sizeof(double),
                           cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    if (colsum) {
      if (pl.engine == 5) {
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(tail.data());
        for (int k = 0; k < K; k++) colsum[k] = cs_value(w[k], w[pl.KP + k]);
      } else {
        memcpy(colsum, tail.data(), K * sizeof(double));
      }
    }
    if (entropy) *entropy = tail[csw * pl.KP];
    if (underflow) *underflow = tail[csw * pl.KP + 1] > 0 ? 1 : 0;
  }
  h->timing.engine = pl.engine;
  h->timing.classes_per_cta = pl.KC;
  h->timing.cluster_size = pl.P;
  h->timing.smem_bytes = (int)pl.smem;
  h->timing.unit_elements = pl.up.U;
  h->timing.unit_pairs = pl.up.n_pairs;
  h->timing.unit_table_rows = pl.up.NR2;
  return MIXDIR_B200_OK;
}
