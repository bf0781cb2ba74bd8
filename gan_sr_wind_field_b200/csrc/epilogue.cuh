// epilogue.cuh — the fused epilogue of the tcgen05 conv kernels, one 16-channel chunk of one accumulator row at a
// time, with 16-byte vector access to every channels-last operand (output, second output, both residuals and the
// LeakyReLU-backward mask).  Semantics: ws_epilogue in include/windsr.h.
#pragma once
#include "common.cuh"

namespace ws {

__device__ __forceinline__ bool view_vec_ok(const View& t) {
  if (!t.ptr || t.cs != 1) return false;
  const int es = t.dtype == WS_F32 ? 4 : 2;
  return ((reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0) && ((t.vs * es) % 16 == 0) && ((t.ns * es) % 16 == 0);
}

struct EpiVec {
  bool dst, out2, res1, res2, mask;
};

__device__ __forceinline__ EpiVec make_epi_vec(const View& dst, const Epi& ep) {
  EpiVec v;
  v.dst = view_vec_ok(dst);
  v.out2 = view_vec_ok(ep.out2);
  v.res1 = view_vec_ok(ep.res1);
  v.res2 = view_vec_ok(ep.res2);
  v.mask = view_vec_ok(ep.mask);
  return v;
}

// 16 consecutive channels [c0, c0+16) of voxel (n, v); channels >= cn read as 0
__device__ __forceinline__ void load16(const View& t, bool vec, int n, int c0, long long v, int cn, float (&o)[16]) {
  const int es = t.dtype == WS_F32 ? 4 : 2;
  if (vec && c0 + 16 <= cn && ((c0 * es) % 16 == 0)) {
    const long long off = t.off(n, c0, v);
    if (t.dtype == WS_F32) {
      const float4* q = reinterpret_cast<const float4*>((const float*)t.ptr + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 f = q[j];
        o[4 * j] = f.x; o[4 * j + 1] = f.y; o[4 * j + 2] = f.z; o[4 * j + 3] = f.w;
      }
    } else {
      const uint4* q = reinterpret_cast<const uint4*>((const __nv_bfloat16*)t.ptr + off);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 u = q[j];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
          o[8 * j + 2 * k] = f.x; o[8 * j + 2 * k + 1] = f.y;
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = (c0 + j < cn) ? t.ld(n, c0 + j, v) : 0.f;
  }
}

__device__ __forceinline__ void store16(const View& t, bool vec, int n, int c0, long long v, int cn,
                                        const float (&y)[16]) {
  const int es = t.dtype == WS_F32 ? 4 : 2;
  if (vec && c0 + 16 <= cn && ((c0 * es) % 16 == 0)) {
    const long long off = t.off(n, c0, v);
    if (t.dtype == WS_BF16) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      uint4* q = reinterpret_cast<uint4*>((__nv_bfloat16*)t.ptr + off);
      q[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      q[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {
      float4* q = reinterpret_cast<float4*>((float*)t.ptr + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < cn) t.st(n, c0 + j, v, y[j]);
  }
}

// Fast path (warp-uniform choice): no per-channel scale / dropout scale / mask / statistics — the epilogues of the
// RDB dense convs, their dgrads (fp32 accumulate), LFF (+bias, two residuals) and the HR convs.  The generic
// version below spends most of its issue slots on predicated-off per-element branches (ncu: the dense-conv dgrad
// was ~80 % epilogue); this one is ~6 instructions per element.
__device__ __forceinline__ bool epi_is_simple(const Epi& ep) { return !ep.mask.ptr && !ep.stat_sum; }

// `r1`: the res1 values of this chunk when the caller prefetched them (software-pipelined epilogue), else null
__device__ __forceinline__ void epilogue16_simple(const Epi& ep, const EpiVec& ev, const View& dst, int n,
                                                  long long v, int cbase, int cn, bool row_ok,
                                                  const uint32_t (&rr)[16], const float* r1 = nullptr) {
  if (!row_ok) return;
  float y[16];
  const float slope = ep.lrelu_slope, alpha = ep.alpha;
#pragma unroll
  for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(rr[j]);
  if (ep.oscale) {  // eval-mode BatchNorm scale (uniform branch; the loads are warp-wide broadcasts)
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] *= (cbase + j < cn ? ep.oscale[cbase + j] : 0.f);
  }
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] += (cbase + j < cn ? ep.bias[cbase + j] : 0.f);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) y[j] = alpha * (y[j] > 0.f ? y[j] : slope * y[j]);
  if (ep.chan_scale) {  // Dropout3d channel scale; alpha commutes with it
    const float* cs = ep.chan_scale + (long long)n * ep.cout + cbase;
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] *= (cbase + j < cn ? cs[j] : 0.f);
  }
  if (ep.res1.ptr) {
    if (r1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = fmaf(ep.beta1, r1[j], y[j]);
    } else {
      float r[16];
      load16(ep.res1, ev.res1, n, cbase, v, cn, r);
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = fmaf(ep.beta1, r[j], y[j]);
    }
  }
  if (ep.res2.ptr) {
    float r[16];
    load16(ep.res2, ev.res2, n, cbase, v, cn, r);
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = fmaf(ep.beta2, r[j], y[j]);
  }
  if (ep.tail_out.ptr && cbase >= ep.tail_c0) {
    // tail channels: LeakyReLU-backward against the saved activation, written to the gradient buffer of the
    // previous conv instead of `dst` (warp-uniform branch: tail_c0 is a multiple of 16)
    float mk[16];
    const int ct = cbase - ep.tail_c0, cnt = cn - ep.tail_c0;
    load16(ep.tail_mask, view_vec_ok(ep.tail_mask), n, ct, v, cnt, mk);
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = mk[j] > 0.f ? y[j] : ep.tail_slope * y[j];
    store16(ep.tail_out, view_vec_ok(ep.tail_out), n, ct, v, cnt, y);
    return;
  }
  if (ep.round_out) {
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = round_tf32(y[j]);
  }
  store16(dst, ev.dst, n, cbase, v, cn, y);
  if (ep.out2.ptr) store16(ep.out2, ev.out2, n, cbase, v, cn, y);
}

// res1 values of one 16-channel chunk, fetched one chunk ahead of the TMEM load that needs them
__device__ __forceinline__ void prefetch_res16(const Epi& ep, const EpiVec& ev, int n, long long v, int cbase, int cn,
                                               int climit, bool row_ok, float (&r)[16]) {
  if (row_ok && cbase < climit) load16(ep.res1, ev.res1, n, cbase, v, cn, r);
}

// One 16-column chunk of fp32 accumulators `rr` (tcgen05.ld 32x32b.x16) for accumulator row = voxel (n, v).
// Must be called by all 32 lanes (the BN-statistics reduction is warp-wide); stores are predicated on row_ok.
// `stat_s` (optional): 512 floats of shared memory — BN sums of this CTA's column c (0..255, relative to
// `stat_c0`) at [c], square sums at [256 + c]; the caller flushes them with flush_bn_stats() after its last chunk.
// Without it every warp sends its partial sums straight to global memory (thousands of CTAs x 4 warps hammering 2 x
// Cout addresses: D's BN convs spent 3/4 of their time there).
__device__ __forceinline__ void epilogue16(const Epi& ep, const EpiVec& ev, const View& dst, int n, long long v,
                                           int cbase, int cn, bool row_ok, const uint32_t (&rr)[16], int lane,
                                           float* stat_s = nullptr, int stat_c0 = 0) {
  float y[16], pre[16];
  float r1[16], r2[16], mk[16];
  const bool has1 = ep.res1.ptr != nullptr, has2 = ep.res2.ptr != nullptr;
  const bool hasm = ep.mask.ptr != nullptr && cbase < ep.mask_c1 && cbase + 16 > ep.mask_c0;
  if (row_ok) {
    if (has1) load16(ep.res1, ev.res1, n, cbase, v, cn, r1);
    if (has2) load16(ep.res2, ev.res2, n, cbase, v, cn, r2);
    if (hasm) load16(ep.mask, ev.mask, n, cbase, v, cn, mk);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int c = cbase + j;
    float t = __uint_as_float(rr[j]);
    const bool ok = row_ok && c < cn;
    if (ok && ep.oscale) t *= ep.oscale[c];
    if (ok && ep.bias) t += ep.bias[c];
    pre[j] = ok ? t : 0.f;
    t = t > 0.f ? t : ep.lrelu_slope * t;
    if (ok && ep.chan_scale) t *= ep.chan_scale[(long long)n * ep.cout + c];
    float o = ep.alpha * t;
    if (ok && has1) o += ep.beta1 * r1[j];
    if (ok && has2) o += ep.beta2 * r2[j];
    if (ok && hasm && c >= ep.mask_c0 && c < ep.mask_c1) o *= (mk[j] > 0.f ? 1.f : ep.mask_slope);
    y[j] = o;
  }
  if (ep.stat_sum) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float s1 = warp_sum(pre[j]);
      const float s2 = warp_sum(pre[j] * pre[j]);
      const int c = cbase + j;
      if (lane == 0 && c < cn) {
        if (stat_s) {
          atomicAdd(&stat_s[c - stat_c0], s1);
          atomicAdd(&stat_s[256 + c - stat_c0], s2);
        } else {
          atomicAdd(&ep.stat_sum[c], s1);
          atomicAdd(&ep.stat_sqsum[c], s2);
        }
      }
    }
  }
  if (row_ok) {
    if (ep.round_out) {
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = round_tf32(y[j]);
    }
    store16(dst, ev.dst, n, cbase, v, cn, y);
    if (ep.out2.ptr) store16(ep.out2, ev.out2, n, cbase, v, cn, y);
  }
}

// one global atomic per channel and CTA: call from `nthreads` epilogue threads (tid 0..nthreads-1) after they have
// all finished their chunks and synchronised
__device__ __forceinline__ void flush_bn_stats(const Epi& ep, const float* stat_s, int c0, int ncols, int cn, int tid,
                                               int nthreads) {
  for (int i = tid; i < 2 * ncols; i += nthreads) {
    const int c = i < ncols ? i : i - ncols;
    if (c0 + c >= cn) continue;
    if (i < ncols) atomicAdd(&ep.stat_sum[c0 + c], stat_s[c]);
    else atomicAdd(&ep.stat_sqsum[c0 + c], stat_s[256 + c]);
  }
}

}  // namespace ws
