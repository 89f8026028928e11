// device_common.cuh — CUDA plumbing shared by the compose / connect / shortest-path kernels:
// error checking, a per-call stream context with stream-ordered (cudaMallocAsync) buffers, device-resident CSR
// FSTs and thin wrappers over the CUB prefix sums used between the hand-written kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <utility>

#include "fst_types.h"
#include "host_fst.h"

namespace b200 {

struct CudaError : FstError {
  using FstError::FstError;
};

#define B200_CUDA(expr)                                                                                      \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess)                                                                                   \
      throw ::b200::CudaError(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                              std::to_string(__LINE__) + " (" #expr ")");                                    \
  } while (0)

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n, int threads = kThreads) { return (unsigned)((n + threads - 1) / threads); }

// Fails loudly when there is no usable GPU: there is no CPU fallback anywhere in this library.
void require_device();
int sm_count();
void configure_device_pool(int dev);

// Held around the kernel section of a call (see device_common.cu): the persistent kernels are exclusive users of the GPU.
std::recursive_mutex& device_exclusive();
using DeviceExclusive = std::lock_guard<std::recursive_mutex>;

// One stream per C-ABI call: concurrent calls from different host threads do not serialise on the legacy stream.
// One CUDA stream per call, taken from a per-host-thread cache: the stream-ordered allocator hands a freed block back
// to the SAME stream at once, whereas blocks freed on a stream that was destroyed after the call are only reused when
// the allocator happens to notice that the free has completed — otherwise the pool grows by hundreds of MB (50-100 ms
// of cuMemCreate / map in the middle of a 5 ms call; seen as outliers of the shortest-path order in a fresh process).
cudaStream_t acquire_thread_stream();
void release_thread_stream(cudaStream_t s);
struct Stream {
  cudaStream_t s = nullptr;
  Stream() {
    require_device();
    s = acquire_thread_stream();
  }
  ~Stream() { if (s) release_thread_stream(s); }
  Stream(const Stream&) = delete;
  Stream& operator=(const Stream&) = delete;
  void sync() const { B200_CUDA(cudaStreamSynchronize(s)); }
};

// Stream-ordered device buffer (cudaMallocAsync pool; growth preserves contents).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  explicit DevBuf(cudaStream_t st) : s(st) {}
  DevBuf(cudaStream_t st, size_t n) : s(st) { reserve_discard(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap), s(o.s) { o.p = nullptr; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; cap = o.cap; s = o.s; o.p = nullptr; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr; cap = 0;
  }
  // Grow to >= n elements, discarding contents.
  void reserve_discard(size_t n) {
    if (n <= cap) return;
    release();
    size_t c = n < 16 ? 16 : n;
    B200_CUDA(cudaMallocAsync((void**)&p, c * sizeof(T), s));
    cap = c;
  }
  // Grow to >= n elements keeping the first `keep` elements (geometric growth).
  void reserve_keep(size_t n, size_t keep) {
    if (n <= cap) return;
    size_t c = cap * 2 > n ? cap * 2 : n;
    if (c < 16) c = 16;
    T* q = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&q, c * sizeof(T), s));
    if (p && keep) B200_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) cudaFreeAsync(p, s);
    p = q; cap = c;
  }
};

// Derived columns of a device-resident machine that the composition matcher reads (dense label arrays: fst1 output
// labels / fst2 input labels, padded).  Built on first use on the machine's own stream and kept with it, so a machine
// that stays in HBM across calls pays for them once; consumers on other streams wait for `ready`.
struct DevLabelColumns {
  std::mutex mu;
  DevBuf<uint32_t> olab, ilab;
  bool has_olab = false, has_ilab = false;
  cudaEvent_t ready = nullptr;
  ~DevLabelColumns() { if (ready) cudaEventDestroy(ready); }
};

// Device-resident CSR FST (what the kernels read and produce).
struct DevFst {
  DevBuf<uint32_t> offsets;  // num_states + 1
  DevBuf<Tr> arcs;
  DevBuf<float> finals;
  uint32_t num_states = 0;
  uint32_t num_arcs = 0;
  bool has_start = false;
  StateId start = 0;
  uint64_t props = props::kNull;
  std::shared_ptr<DevLabelColumns> columns;  // lazily built by label_column(); never copied with the arrays
  DevFst() = default;
  explicit DevFst(cudaStream_t s) : offsets(s), arcs(s), finals(s) {}
  size_t bytes() const { return (size_t)(num_states + 1) * 4 + (size_t)num_arcs * 16 + (size_t)num_states * 4; }
};

// Dense label column of `f` (olabel = true: output labels, else input labels), padded with kNoLabel by `pad` entries;
// built once per machine, valid on stream `s` when the call returns.  *built = a kernel was launched for it.
const uint32_t* label_column(const DevFst& f, bool olabel, uint32_t pad, cudaStream_t s, bool* built);

DevFst upload(const CsrFst& h, cudaStream_t s);
CsrFst download(const DevFst& d, cudaStream_t s);

// Exclusive prefix sum of n uint32 values (CUB DeviceScan; compiled once in device_common.cu); out may alias in.
void exclusive_sum_u32(const uint32_t* in, uint32_t* out, size_t n, DevBuf<uint8_t>& temp, cudaStream_t s);

// Stable radix sort of (u64 key, u32 value) pairs on key bits [0, end_bit) (CUB; compiled once in device_common.cu).
void sort_pairs_u64_u32(const unsigned long long* keys_in, unsigned long long* keys_out, const uint32_t* vals_in,
                        uint32_t* vals_out, size_t n, int end_bit, DevBuf<uint8_t>& temp, cudaStream_t s);

inline uint32_t read_u32(const uint32_t* dptr, cudaStream_t s) {
  uint32_t v = 0;
  B200_CUDA(cudaMemcpyAsync(&v, dptr, 4, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  return v;
}

}  // namespace b200
