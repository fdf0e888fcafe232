// Library identity and error plumbing for lib3dvlp_b200.so (see include/pointnet2_b200.h).
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace pn2 {

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return PN2_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);  // racing threads store the same value
  if (!v) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    __atomic_store_n(&cached[dev], v, __ATOMIC_RELAXED);
  }
  return v;
}

void keep_workspace_pool() {
  static int ready[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  if (__atomic_load_n(&ready[dev], __ATOMIC_ACQUIRE)) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);  // idempotent
  }
  __atomic_store_n(&ready[dev], 1, __ATOMIC_RELEASE);
}

// One int per device, set by a kernel whose bounded mbarrier wait timed out (a protocol bug must
// surface as an error, never as a hung GPU).  Polled by pn2_poll_device_error().
static int* g_err_flag[64] = {nullptr};

// Entry points are called from several host threads (DataParallel workers, the autograd engine): the
// first caller per device publishes the flag with a compare-and-swap, a loser frees its allocation.
int* device_err_flag() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  int* cur = __atomic_load_n(&g_err_flag[dev], __ATOMIC_ACQUIRE);
  if (cur) return cur;
  int* p = nullptr;
  if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
  if (cudaMemset(p, 0, sizeof(int)) != cudaSuccess) { cudaFree(p); return nullptr; }
  int* expected = nullptr;
  if (__atomic_compare_exchange_n(&g_err_flag[dev], &expected, p, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) return p;
  cudaFree(p);
  return expected;
}

// ---- host-side statistics (pn2_get_stat): which kernel variant a call really launched ------------
static const char* const kStatName[STAT_COUNT] = {"fps_cluster_launches", "fps_bucket_launches", "fps_pair_launches",
                                                   "bq_grid_launches", "bq_small_grid_launches", "bq_scan_launches",
                                                   "nn_small_launches", "nn_tiled_launches", "wgrad_tma_launches",
                                                   "wgrad_cpasync_launches"};
static long long g_stat[STAT_COUNT] = {0};
void stat_add(Stat s, long long v) { __atomic_fetch_add(&g_stat[s], v, __ATOMIC_RELAXED); }

// ---- run-time options ----------------------------------------------------------------------------
static const char* const kOptName[OPT_COUNT] = {"fps_bucket", "bq_grid", "sa_tiles_per_group", "wgrad_tma"};
static const char* const kOptEnv[OPT_COUNT] = {"PN2_FPS_BUCKET", "PN2_BQ_GRID", "PN2_SA_TILES_PER_GROUP", "PN2_WGRAD_TMA"};
static const int kOptDefault[OPT_COUNT] = {0, 1, 0, 1};
static int g_opt[OPT_COUNT] = {-1, -1, -1, -1};
// per-thread override (pn2_set_option_thread): -1 = follow the process-wide value.  A host thread that
// captures a CUDA graph with other settings does not disturb calls made by other threads meanwhile.
static thread_local int t_opt[OPT_COUNT] = {-1, -1, -1, -1};

int option_value(Option o) {
  if (t_opt[o] >= 0) return t_opt[o];
  int v = __atomic_load_n(&g_opt[o], __ATOMIC_RELAXED);
  if (v < 0) {
    const char* e = getenv(kOptEnv[o]);
    v = e ? atoi(e) : kOptDefault[o];
    if (v < 0) v = 0;
    __atomic_store_n(&g_opt[o], v, __ATOMIC_RELAXED);
  }
  return v;
}

}  // namespace pn2

extern "C" int pn2_set_option(const char* name, int value) {
  for (int i = 0; i < pn2::OPT_COUNT; ++i) {
    if (name && strcmp(name, pn2::kOptName[i]) == 0) {
      __atomic_store_n(&pn2::g_opt[i], value < 0 ? 0 : value, __ATOMIC_RELAXED);
      return PN2_OK;
    }
  }
  pn2::set_error("set_option: unknown option '%s' (known: fps_bucket, bq_grid, sa_tiles_per_group, wgrad_tma)", name ? name : "(null)");
  return PN2_ERR_INVALID;
}

extern "C" int pn2_set_option_thread(const char* name, int value) {
  for (int i = 0; i < pn2::OPT_COUNT; ++i) {
    if (name && strcmp(name, pn2::kOptName[i]) == 0) {
      pn2::t_opt[i] = value < 0 ? -1 : value;
      return PN2_OK;
    }
  }
  pn2::set_error("set_option_thread: unknown option '%s'", name ? name : "(null)");
  return PN2_ERR_INVALID;
}

extern "C" long long pn2_get_stat(const char* name) {
  for (int i = 0; i < pn2::STAT_COUNT; ++i)
    if (name && strcmp(name, pn2::kStatName[i]) == 0) return __atomic_load_n(&pn2::g_stat[i], __ATOMIC_RELAXED);
  return -1;
}

extern "C" int pn2_get_option(const char* name) {
  for (int i = 0; i < pn2::OPT_COUNT; ++i)
    if (name && strcmp(name, pn2::kOptName[i]) == 0) return pn2::option_value(static_cast<pn2::Option>(i));
  return -1;
}

extern "C" int pn2_poll_device_error(int* flag_out, void* stream) {
  PN2_RANGE("pn2_poll_device_error");
  int* p = pn2::device_err_flag();
  if (!p) { pn2::set_error("poll_device_error: no flag"); return PN2_ERR_CUDA; }
  int v = 0;
  PN2_CUDA(cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, pn2::as_stream(stream)));
  PN2_CUDA(cudaStreamSynchronize(pn2::as_stream(stream)));
  if (v != 0) PN2_CUDA(cudaMemsetAsync(p, 0, sizeof(int), pn2::as_stream(stream)));
  if (flag_out) *flag_out = v;
  return PN2_OK;
}

extern "C" int pn2_abi_version(void) { return 1; }

extern "C" const char* pn2_last_error(void) { return pn2::g_err; }

extern "C" int pn2_device_arch(int* major, int* minor) {
  int dev = 0;
  PN2_CUDA(cudaGetDevice(&dev));
  int ma = 0, mi = 0;
  PN2_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  PN2_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (major) *major = ma;
  if (minor) *minor = mi;
  return PN2_OK;
}
