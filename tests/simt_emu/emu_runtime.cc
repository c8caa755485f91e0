// emu_runtime.cc - host SIMT emulator runtime + CUDA runtime stub (TEST ONLY, see simt.h / cuda_stub.h).
// One OS thread per CUDA thread of a CTA; CTAs run one after another.
#include <stdio.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <thread>
#include <vector>

#define PDSP_EMU 1
#include "../../pragma_dsp_b200/csrc/simt.h"
#include "cuda_stub.h"

namespace simt {
class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const long gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen_ != gen; });
    }
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  long gen_ = 0;
};

struct EmuBlock {
  int nthreads;
  std::vector<unsigned char> smem;
  std::unique_ptr<Barrier> block_bar;
  std::vector<std::unique_ptr<Barrier>> warp_bar;
  std::vector<unsigned char> mailbox;  // 16 B per thread
  std::mutex named_mu;
  std::map<int, std::unique_ptr<Barrier>> named;
};

thread_local EmuThread emu_self;
static EmuBlock* blk() { return static_cast<EmuBlock*>(emu_self.block); }
static std::atomic<long long> g_launches{0};

void emu_sync_warp() { blk()->warp_bar[emu_self.tid / 32]->wait(); }
void emu_sync_block() { blk()->block_bar->wait(); }
void emu_sync_named(int id, int nthreads) {
  Barrier* b;
  {
    std::lock_guard<std::mutex> lk(blk()->named_mu);
    auto& slot = blk()->named[id];
    if (!slot) slot.reset(new Barrier(nthreads));
    b = slot.get();
  }
  b->wait();
}
void emu_shfl(const void* in, void* out, int bytes, int src_lane, int width, bool is_xor) {
  EmuBlock* b = blk();
  const int tid = emu_self.tid, lane = tid & 31, warp_base = tid & ~31;
  memcpy(&b->mailbox[(size_t)tid * 16], in, (size_t)bytes);
  emu_sync_warp();
  const int src = is_xor ? (lane ^ src_lane) : ((lane & ~(width - 1)) | (src_lane & (width - 1)));
  memcpy(out, &b->mailbox[(size_t)(warp_base + src) * 16], (size_t)bytes);
  emu_sync_warp();
}

// mbarrier emulation: the 64-bit word in shared memory holds {phase bit 63, pending tx bytes 0..31, armed bit 32}.
// expect_tx arms the current phase with a byte count, complete_tx retires bytes and flips the phase at zero.
static std::mutex g_mbar_mu;
static std::condition_variable g_mbar_cv;
void emu_mbar_init(unsigned long long* bar) {
  std::lock_guard<std::mutex> lk(g_mbar_mu);
  *bar = 0;
}
void emu_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  std::lock_guard<std::mutex> lk(g_mbar_mu);
  const unsigned long long phase = *bar >> 63;
  const long long pending = (long long)(int)(*bar & 0xffffffffull) + (long long)bytes;
  if (pending == 0) {
    *bar = (phase ^ 1ull) << 63;  // transfers finished before the arrive: phase completes now
    g_mbar_cv.notify_all();
  } else {
    *bar = (phase << 63) | (1ull << 32) | (unsigned long long)(unsigned)(int)pending;
  }
}
void emu_mbar_complete_tx(unsigned long long* bar, unsigned bytes) {
  std::lock_guard<std::mutex> lk(g_mbar_mu);
  const unsigned long long phase = *bar >> 63;
  const bool armed = (*bar >> 32) & 1ull;
  const long long pending = (long long)(int)(*bar & 0xffffffffull) - (long long)bytes;
  if (armed && pending == 0) {
    *bar = (phase ^ 1ull) << 63;
    g_mbar_cv.notify_all();
  } else {
    *bar = (phase << 63) | ((armed ? 1ull : 0ull) << 32) | (unsigned long long)(unsigned)(int)pending;
  }
}
void emu_mbar_wait(unsigned long long* bar, unsigned parity) {
  std::unique_lock<std::mutex> lk(g_mbar_mu);
  g_mbar_cv.wait(lk, [&] { return (unsigned)(*bar >> 63) != (parity & 1u); });
}

void emu_launch(int nblocks, int nthreads, size_t smem_bytes, const std::function<void()>& body) {
  ++g_launches;
  for (int bid = 0; bid < nblocks; ++bid) {
    EmuBlock b;
    b.nthreads = nthreads;
    b.smem.assign(smem_bytes + 64, 0xCD);  // poison: uninitialised reads show up as garbage
    b.block_bar.reset(new Barrier(nthreads));
    for (int w = 0; w * 32 < nthreads; ++w) {
      const int n = nthreads - w * 32;
      b.warp_bar.emplace_back(new Barrier(n < 32 ? n : 32));
    }
    b.mailbox.assign((size_t)nthreads * 16, 0);
    unsigned char* sm = b.smem.data();
    sm += (16 - (reinterpret_cast<uintptr_t>(sm) & 15)) & 15;
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
      th.emplace_back([&, t] {
        emu_self = EmuThread{t, bid, nblocks, nthreads, sm, &b};
        body();
      });
    }
    for (auto& x : th) x.join();
  }
}
}  // namespace simt

// ---------------------------------------------------------------- CUDA runtime stub
namespace {
std::mutex g_mu;
std::map<const char*, size_t> g_pinned;  // base -> bytes
std::set<void*> g_device;
std::atomic<long long> g_h2d{0}, g_d2h{0};
}  // namespace

// PDSP_STUB_DEVICES=n pretends to have n devices (all of them this host's memory): lets the multi-device group API run
static int stub_devices() {
  const char* e = getenv("PDSP_STUB_DEVICES");
  const int n = e ? atoi(e) : 1;
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}
cudaError_t cudaGetDeviceCount(int* n) {
  *n = stub_devices();
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) { return d >= 0 && d < stub_devices() ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) {
  *can = 1;
  return cudaSuccess;
}
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaMemcpyPeerAsync(void* dst, int, const void* src, int, size_t bytes, cudaStream_t) {
  memcpy(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaDeviceGetPCIBusId(char* id, int len, int dev) {
  snprintf(id, (size_t)len, "0000:%02x:00.0", dev);
  return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->major = 10;
  p->minor = 0;
  p->multiProcessorCount = 2;
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "stub error"; }
cudaError_t cudaMalloc(void** p, size_t bytes) {
  *p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
  memset(*p, 0xAB, bytes);
  std::lock_guard<std::mutex> lk(g_mu);
  g_device.insert(*p);
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_device.erase(p)) return cudaErrorInvalidValue;
  free(p);
  return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) {
  *p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
  std::lock_guard<std::mutex> lk(g_mu);
  g_pinned[static_cast<const char*>(*p)] = bytes;
  return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
  if (!p) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_pinned.erase(static_cast<const char*>(p))) return cudaErrorInvalidValue;
  free(p);
  return cudaSuccess;
}
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) {
  *d = h;
  return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int value, size_t bytes) {
  memset(p, value, bytes);
  return cudaSuccess;
}
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  if (kind == cudaMemcpyHostToDevice) ++g_h2d;
  if (kind == cudaMemcpyDeviceToHost) ++g_d2h;
  memcpy(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t) {
  return cudaMemcpy(dst, src, bytes, kind);
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
                              cudaStream_t) {
  if (kind == cudaMemcpyHostToDevice) ++g_h2d;
  if (kind == cudaMemcpyDeviceToHost) ++g_d2h;
  for (size_t r = 0; r < height; ++r) memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = reinterpret_cast<cudaStream_t>(malloc(1));
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  free(s);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = reinterpret_cast<cudaEvent_t>(malloc(1));
  return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  free(e);
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  a->type = cudaMemoryTypeUnregistered;
  const char* c = static_cast<const char*>(p);
  auto it = g_pinned.upper_bound(c);
  if (it != g_pinned.begin()) {
    --it;
    if (c >= it->first && c < it->first + it->second) a->type = cudaMemoryTypeHost;
  }
  return cudaSuccess;
}
extern "C" __attribute__((visibility("default"))) long long pdsp_stub_counter(int which) {
  switch (which) {
    case 0: return g_h2d.load();
    case 1: return g_d2h.load();
    case 2: return simt::g_launches.load();
    default: {
      std::lock_guard<std::mutex> lk(g_mu);
      return (long long)g_device.size();
    }
  }
}
