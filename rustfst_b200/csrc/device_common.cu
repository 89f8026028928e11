// device_common.cu — device discovery and host<->HBM marshalling of CSR FSTs.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "device_common.cuh"

namespace b200 {
void configure_device_pool(int dev);
namespace {
std::once_flag g_once;
int g_sms = 0;
std::string g_init_error;

void init_device() {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_init_error = std::string("librustfst_b200: no CUDA device available (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                   "); this library has no CPU fallback";
    cudaGetLastError();
    return;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  configure_device_pool(dev);
}
}  // namespace

// Keep freed blocks in the stream-ordered pool: the per-wave scratch buffers are re-allocated constantly.
void configure_device_pool(int dev) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  } else {
    cudaGetLastError();
  }
}

// The persistent cooperative kernels (compose, trim, relaxation, DFS order) each want every SM: two of them from
// different calls can become co-resident, or one of them only PARTLY resident, and then spin at their grid barriers
// while they wait for each other's CTAs to leave (measured: 15 ms per compose with two callers on a good run, 150 ms on
// a bad one).  So the kernel section of a call is exclusive per process; uploads and downloads stay outside it and
// overlap with another caller's kernels.
std::recursive_mutex& device_exclusive() {
  static std::recursive_mutex m;
  return m;
}

namespace {
struct ThreadStreams {
  std::vector<cudaStream_t> idle;
  ~ThreadStreams() { for (cudaStream_t s : idle) cudaStreamDestroy(s); }
};
thread_local ThreadStreams t_streams;
}  // namespace
cudaStream_t acquire_thread_stream() {
  if (!t_streams.idle.empty()) {
    cudaStream_t s = t_streams.idle.back();
    t_streams.idle.pop_back();
    return s;
  }
  cudaStream_t s = nullptr;
  B200_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  return s;
}
void release_thread_stream(cudaStream_t s) {
  cudaStreamSynchronize(s);  // nothing of this call may still be running when the next call reuses the stream
  t_streams.idle.push_back(s);
}

void require_device() {
  std::call_once(g_once, init_device);
  if (!g_init_error.empty()) throw CudaError(g_init_error);
}
int sm_count() { require_device(); return g_sms; }

DevFst upload(const CsrFst& h, cudaStream_t s) {
  if (!h.inf_finals.empty())
    throw FstError("final weights equal to +inf (TropicalWeight::zero) are not supported on the device path");
  DevFst d(s);
  size_t n = h.num_states(), a = h.arcs.size();
  d.offsets.reserve_discard(n + 1);
  d.arcs.reserve_discard(a);
  d.finals.reserve_discard(n);
  B200_CUDA(cudaMemcpyAsync(d.offsets.p, h.offsets.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
  if (a) B200_CUDA(cudaMemcpyAsync(d.arcs.p, h.arcs.data(), a * sizeof(Tr), cudaMemcpyHostToDevice, s));
  if (n) B200_CUDA(cudaMemcpyAsync(d.finals.p, h.finals.data(), n * 4, cudaMemcpyHostToDevice, s));
  d.num_states = (uint32_t)n; d.num_arcs = (uint32_t)a;
  d.has_start = h.has_start; d.start = h.start; d.props = h.props;
  B200_CUDA(cudaStreamSynchronize(s));
  return d;
}

CsrFst download(const DevFst& d, cudaStream_t s) {
  CsrFst h;
  size_t n = d.num_states, a = d.num_arcs;
  h.offsets.resize(n + 1);
  h.arcs.resize(a);
  h.finals.resize(n);
  B200_CUDA(cudaMemcpyAsync(h.offsets.data(), d.offsets.p, (n + 1) * 4, cudaMemcpyDeviceToHost, s));
  if (a) B200_CUDA(cudaMemcpyAsync(h.arcs.data(), d.arcs.p, a * sizeof(Tr), cudaMemcpyDeviceToHost, s));
  if (n) B200_CUDA(cudaMemcpyAsync(h.finals.data(), d.finals.p, n * 4, cudaMemcpyDeviceToHost, s));
  h.has_start = d.has_start; h.start = d.start; h.props = d.props & props::kTrinary;
  B200_CUDA(cudaStreamSynchronize(s));
  return h;
}

// ---- page-locked host pool ------------------------------------------------------------------------------------
namespace {
constexpr size_t kPinThreshold = 256 * 1024;          // smaller blocks: plain malloc
size_t pool_keep_bytes() {                             // cached free blocks above this are returned to the driver
  static const size_t v = [] {
    const char* e = std::getenv("B200_PINNED_POOL_MB");  // default 16 GiB; the C3 end-to-end path cycles ~1.4 GB
    return e ? (size_t)std::strtoull(e, nullptr, 10) << 20 : (size_t)16 << 30;
  }();
  return v;
}
std::mutex g_pool_mu;
std::map<size_t, std::vector<void*>> g_pool_free;     // size class -> cached blocks
size_t g_pool_cached = 0;
std::atomic<int> g_pin_state{0};                      // 0 unknown, 1 pinned allocations work, -1 no device
size_t size_class(size_t b) { size_t c = kPinThreshold; while (c < b) c <<= 1; return c; }
}  // namespace

void* host_pool_alloc(size_t bytes) {
  if (bytes == 0) bytes = 1;
  void* p = nullptr;
  if (bytes < kPinThreshold) {
    p = std::malloc(bytes);
    if (!p) throw std::bad_alloc();
    return p;
  }
  size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> g(g_pool_mu);
    auto it = g_pool_free.find(cls);
    if (it != g_pool_free.end() && !it->second.empty()) {
      p = it->second.back(); it->second.pop_back(); g_pool_cached -= cls;
      return p;
    }
    if (g_pin_state == 0) {
      int n = 0;
      g_pin_state = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : -1;
      cudaGetLastError();
    }
  }
  if (g_pin_state == 1 && cudaHostAlloc(&p, cls, cudaHostAllocPortable) == cudaSuccess) return p;
  cudaGetLastError();
  // no device (CPU-only box) or pinning refused: ordinary memory, tagged by a side table so free() knows
  p = std::malloc(cls);
  if (!p) throw std::bad_alloc();
  std::lock_guard<std::mutex> g(g_pool_mu);
  g_pool_free[0].push_back(p);  // class 0 = registry of malloc'ed large blocks
  return p;
}

void host_pool_free(void* p, size_t bytes) noexcept {
  if (!p) return;
  if (bytes == 0) bytes = 1;
  if (bytes < kPinThreshold) { std::free(p); return; }
  size_t cls = size_class(bytes);
  std::lock_guard<std::mutex> g(g_pool_mu);
  auto& reg = g_pool_free[0];
  for (size_t i = 0; i < reg.size(); i++)
    if (reg[i] == p) { reg[i] = reg.back(); reg.pop_back(); std::free(p); return; }
  g_pool_free[cls].push_back(p);
  g_pool_cached += cls;
  // Over the cap: give the LARGEST cached blocks back to the driver first.  (Returning whatever block happens to be freed
  // made a full cache thrash: every 20 MB result block of a batched call then cost a cudaFreeHost + cudaHostAlloc pair.)
  while (g_pool_cached > pool_keep_bytes()) {
    auto it = g_pool_free.end();
    bool evicted = false;
    while (it != g_pool_free.begin()) {
      --it;
      if (it->first != 0 && !it->second.empty()) {
        cudaFreeHost(it->second.back());
        it->second.pop_back();
        g_pool_cached -= it->first;
        evicted = true;
        break;
      }
    }
    if (!evicted) break;
  }
}

void exclusive_sum_u32(const uint32_t* in, uint32_t* out, size_t n, DevBuf<uint8_t>& temp, cudaStream_t s) {
  if (n == 0) return;
  size_t bytes = 0;
  B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n, s));
  temp.reserve_discard(bytes);
  B200_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, bytes, in, out, (int64_t)n, s));
}

// Stable LSD radix sort of (key, value) pairs (CUB DeviceRadixSort); results land in keys_out / vals_out.
void sort_pairs_u64_u32(const unsigned long long* keys_in, unsigned long long* keys_out, const uint32_t* vals_in,
                        uint32_t* vals_out, size_t n, int end_bit, DevBuf<uint8_t>& temp, cudaStream_t s) {
  if (n == 0) return;
  size_t bytes = 0;
  B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int64_t)n, 0, end_bit, s));
  temp.reserve_discard(bytes);
  B200_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, bytes, keys_in, keys_out, vals_in, vals_out, (int64_t)n, 0, end_bit, s));
}

}  // namespace b200
