// common.cuh — shared device helpers: tensor views, the fused epilogue, error plumbing.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/windsr.h"

namespace ws {

void set_error(const char* fmt, ...);
void count_launch(int n);

#define WS_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ws::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

// after every kernel launch: count it (ws_launch_count) and surface launch errors
#define WS_POST_LAUNCH(n)                 \
  do {                                    \
    ws::count_launch(n);                  \
    WS_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

#define WS_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      ws::set_error(__VA_ARGS__);  \
      return 2;                    \
    }                              \
  } while (0)

// Launch with programmatic stream serialization (and optionally as clusters of `cluster_x` CTAs).  ONLY for
// kernels that execute ptx::griddep_wait() before touching global memory a predecessor may still be writing.
// WS_DISABLE_PDL=1 falls back to plain stream order.
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("WS_DISABLE_PDL");
    return !(v && v[0] && v[0] != '0');
  }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Device-side copy of ws_tensor with typed accessors.
struct View {
  void* ptr;
  int dtype;
  long long ns, vs, cs;

  __host__ __device__ View() : ptr(nullptr), dtype(WS_F32), ns(0), vs(0), cs(0) {}
  __host__ __device__ explicit View(const ws_tensor& t)
      : ptr(t.ptr), dtype(t.dtype), ns(t.nstride), vs(t.vstride), cs(t.cstride) {}
  __host__ View(const ws_tensor* t) {
    if (t) { ptr = t->ptr; dtype = t->dtype; ns = t->nstride; vs = t->vstride; cs = t->cstride; }
    else { ptr = nullptr; dtype = WS_F32; ns = vs = cs = 0; }
  }
  __device__ __forceinline__ long long off(int n, int c, long long v) const {
    return (long long)n * ns + v * vs + (long long)c * cs;
  }
  __device__ __forceinline__ float ld(long long o) const {
    return dtype == WS_F32 ? ((const float*)ptr)[o] : __bfloat162float(((const __nv_bfloat16*)ptr)[o]);
  }
  __device__ __forceinline__ void st(long long o, float f) const {
    if (dtype == WS_F32) ((float*)ptr)[o] = f;
    else ((__nv_bfloat16*)ptr)[o] = __float2bfloat16_rn(f);
  }
  __device__ __forceinline__ float ld(int n, int c, long long v) const { return ld(off(n, c, v)); }
  __device__ __forceinline__ void st(int n, int c, long long v, float f) const { st(off(n, c, v), f); }
  __host__ __device__ bool valid() const { return ptr != nullptr; }
  __host__ __device__ int esize() const { return dtype == WS_F32 ? 4 : 2; }
};

// fp32 -> nearest TF32 value (10 mantissa bits), ties away from zero.  The tensor cores TRUNCATE fp32 operands to TF32;
// rounding the stored operand first halves that error and removes its bias.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Device-side epilogue (see ws_epilogue in windsr.h for the exact semantics).
struct Epi {
  const float* bias;
  const float* oscale;
  const float* chan_scale;
  float lrelu_slope, alpha, beta1, beta2;
  View res1, res2, mask, out2;
  int mask_c0, mask_c1;
  float mask_slope;
  float* stat_sum;
  float* stat_sqsum;
  View tail_out, tail_mask;  // see ws_epilogue: fused LeakyReLU-backward of the tail channels
  int tail_c0;
  float tail_slope;
  int round_out;  // ws_epilogue::flags bit 0: round the stored values to TF32 (pure activations in TF32 mode)
  int cout;  // channel count of the output (indexing chan_scale)

  __host__ Epi() {}
  __host__ Epi(const ws_epilogue* e, int cout_) {
    cout = cout_;
    if (e) {
      bias = e->bias; oscale = e->oscale; chan_scale = e->chan_scale;
      lrelu_slope = e->lrelu_slope; alpha = e->alpha; beta1 = e->beta1; beta2 = e->beta2;
      res1 = View(e->res1); res2 = View(e->res2); mask = View(e->mask); out2 = View(e->out2);
      mask_c0 = e->mask_c0; mask_c1 = e->mask_c1; mask_slope = e->mask_slope;
      stat_sum = e->stat_sum; stat_sqsum = e->stat_sqsum;
      tail_out = View(e->tail_out); tail_mask = View(e->tail_mask); tail_c0 = e->tail_c0; tail_slope = e->tail_slope;
      round_out = e->flags & 1;
    } else {
      round_out = 0;
      tail_c0 = 0; tail_slope = 1.f;
      bias = oscale = chan_scale = nullptr;
      lrelu_slope = 1.f; alpha = 1.f; beta1 = beta2 = 0.f;
      mask_c0 = mask_c1 = 0; mask_slope = 1.f;
      stat_sum = stat_sqsum = nullptr;
    }
  }

  // Returns the value to store; `pre` receives the pre-activation (for BN statistics).
  __device__ __forceinline__ float apply(float acc, int n, int c, long long v, float& pre) const {
    float t = acc;
    if (oscale) t *= oscale[c];
    if (bias) t += bias[c];
    pre = t;
    t = t > 0.f ? t : lrelu_slope * t;
    if (chan_scale) t *= chan_scale[(long long)n * cout + c];
    float y = alpha * t;
    if (res1.ptr) y += beta1 * res1.ld(n, c, v);
    if (res2.ptr) y += beta2 * res2.ld(n, c, v);
    if (mask.ptr && c >= mask_c0 && c < mask_c1) {
      float m = mask.ld(n, c, v);
      y *= (m > 0.f ? 1.f : mask_slope);
    }
    if (round_out) y = round_tf32(y);
    return y;
  }
};

struct ConvGeom {
  int n, x, y, z, cin, cout;
  int kx, ky, kz, sx, sy, sz, px, py, pz;
  int xo, yo, zo;
  __host__ ConvGeom() {}
  __host__ explicit ConvGeom(const ws_conv_shape& s) {
    n = s.n; x = s.x; y = s.y; z = s.z; cin = s.cin; cout = s.cout;
    kx = s.kx; ky = s.ky; kz = s.kz; sx = s.sx; sy = s.sy; sz = s.sz; px = s.px; py = s.py; pz = s.pz;
    xo = (x + 2 * px - kx) / sx + 1;
    yo = (y + 2 * py - ky) / sy + 1;
    zo = (z + 2 * pz - kz) / sz + 1;
  }
  __host__ __device__ int taps() const { return kx * ky * kz; }
  __host__ __device__ long long vin() const { return (long long)x * y * z; }
  __host__ __device__ long long vout() const { return (long long)xo * yo * zo; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

int validate_shape(const ws_conv_shape* s);

}  // namespace ws
