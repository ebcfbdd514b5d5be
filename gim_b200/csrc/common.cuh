// common.cuh - shared host/device helpers for libgimb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

namespace gimb {

// ---------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI; failures set a thread-local message.
void set_error(const char* fmt, ...);
#define GIMB_CHECK(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ::gimb::set_error(__VA_ARGS__);  \
      return 1;                        \
    }                                  \
  } while (0)
#define GIMB_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::gimb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                        __LINE__);                                                            \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)
#define GIMB_TRY(expr)         \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per device (context), so the "done"
// state is a bit per device ordinal, set only after the call succeeded.
#define GIMB_SMEM_OPTIN(kernel, bytes)                                                                     \
  do {                                                                                                     \
    static unsigned long long _done_mask = 0ull;                                                           \
    int _dev = 0;                                                                                          \
    GIMB_CUDA(cudaGetDevice(&_dev));                                                                       \
    const unsigned long long _bit = 1ull << (_dev & 63);                                                   \
    if (!(_done_mask & _bit)) {                                                                            \
      GIMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
      _done_mask |= _bit;                                                                                  \
    }                                                                                                      \
  } while (0)

// Every C entry point runs on its handle's device and leaves the caller's current device untouched.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------------------------------------
// Stack (bump) allocator over the caller-provided workspace.  In dry mode nothing is dereferenced:
// the same forward code runs with launches skipped and `peak` is the workspace requirement.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, top = 0, peak = 0;
  bool dry = false;
  bool overflow = false;
  static constexpr size_t kAlign = 1024;
  void* alloc_bytes(size_t n) {
    size_t start = (top + kAlign - 1) / kAlign * kAlign;
    top = start + n;
    if (top > peak) peak = top;
    if (!dry && top > cap) {
      overflow = true;
      return nullptr;
    }
    return dry ? (void*)(uintptr_t)(0x1000 + start) : (void*)(base + start);
  }
  template <typename T>
  T* alloc(size_t n) {
    return (T*)alloc_bytes(n * sizeof(T));
  }
  size_t mark() const { return top; }
  void release(size_t m) { top = m; }
};

struct Marker {  // per-stage CUDA-event profiling hook (no-op when null)
  virtual void mark(const char* name) = 0;
  virtual ~Marker() {}
};

struct Ctx {
  cudaStream_t stream = nullptr;
  Marker* marker = nullptr;
  void mark(const char* name) {
    if (marker && !dry) marker->mark(name);
  }
  Arena arena;
  bool dry = false;         // plan only: no launches
  uint64_t launches = 0;    // kernels launched through this context
  int sm_count = 148;
};

#define GIMB_LAUNCH_CHECK()                                                                  \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ::gimb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                           \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// activation / epilogue codes of the GEMM-shaped kernels
enum Act : int {
  ACT_NONE = 0,
  ACT_RELU = 1,
  ACT_LEAKY = 2,   // LeakyReLU(0.01)
  ACT_ELU1 = 3,    // elu(x) + 1          (linear-attention feature map)
  ACT_DIVS = 4,    // x / div             (values / v_length)
};

__device__ __forceinline__ float apply_act(float v, int act, float div) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
    case ACT_ELU1: return v > 0.f ? v + 1.f : expm1f(v) + 1.f;
    case ACT_DIVS: return __fdiv_rn(v, div);
    default: return v;
  }
}

}  // namespace gimb
