// elementwise.cu — HBM-bound helpers around the convolutions: weight packing, strided/dtype copies
// (the dense-concat and layout changes of torch_blocks.py:214,286 / Generator…py:228), axpby, LeakyReLU
// backward, the x/y nearest upsample and its 2x2 block-sum backward (torch_blocks.py:347), and the
// BatchNorm3d statistics finalize / apply / backward (torch_blocks.py:24-25).
#include <string.h>
#include "common.cuh"

namespace ws {
namespace {

constexpr int kBlock = 256;

inline int grid_for(long long total, int per_thread = 1) {
  long long b = (total + (long long)kBlock * per_thread - 1) / ((long long)kBlock * per_thread);
  long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- weight packing ---------------------------------------------------------------------------------
// w: (cout, cin, taps) fp32
__global__ void pack_simt(const float* __restrict__ w, float* __restrict__ p, int cout, int cin, int taps,
                          int dgrad) {
  long long total = (long long)cout * cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the packed array
    if (!dgrad) {  // [tap][cin][cout]
      int co = (int)(i % cout);
      long long r = i / cout;
      int ci = (int)(r % cin);
      int tap = (int)(r / cin);
      p[i] = w[((long long)co * cin + ci) * taps + tap];
    } else {  // [tap][cout][cin]
      int ci = (int)(i % cin);
      long long r = i / cin;
      int co = (int)(r % cout);
      int tap = (int)(r / cout);
      p[i] = w[((long long)co * cin + ci) * taps + tap];
    }
  }
}

// tcgen05 packings, bf16, zero padded.
// fwd:   p[tap][co_pad][ci_pad]          = w[co][ci][tap]
// dgrad: p[taps-1-tap][ci_pad16][co_pad8] = w[co][ci][tap]   (taps-1-tap == flip of all three axes)
__device__ __forceinline__ void store_packed(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_packed(float* p, long long i, float v) { p[i] = round_tf32(v); }  // TF32 operands

template <typename T>
__global__ void pack_tc(const float* __restrict__ w, T* __restrict__ p, int cout, int cin,
                        int taps, int rows_pad, int cols_pad, int dgrad) {
  long long total = (long long)taps * rows_pad * cols_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int col = (int)(i % cols_pad);
    long long r = i / cols_pad;
    int row = (int)(r % rows_pad);
    int tp = (int)(r / rows_pad);
    float v = 0.f;
    if (!dgrad) {
      if (row < cout && col < cin) v = w[((long long)row * cin + col) * taps + tp];
    } else {
      if (row < cin && col < cout) v = w[((long long)col * cin + row) * taps + (taps - 1 - tp)];
    }
    store_packed(p, i, v);
  }
}

// Several stride-1 tcgen05 packings in one launch (blockIdx.y = entry): the residual dense block executor repacks
// all of its convs after every optimizer step, and five 3-us launches per block direction added up to 1.4 ms.
struct PackEntry {
  const float* w;
  __nv_bfloat16* p;
  int cout, cin, taps, rows_pad, cols_pad, dgrad;
  int fold;  // > 0: x-fold forward packing — `fold` = kx taps side by side on the rows: p[(ky,kz)][dx*cout+co][ci]
  int fold_z;  // with fold > 0: fold the kz taps instead (persistent RDB kernel): p[(kx,ky)][dz*cout+co][ci]
  int tf32;    // fp32 output (TF32 mode) instead of bf16
};
struct PackTable {
  int n;
  PackEntry e[WS_RDB_MAX_CONVS + 1];
};
__device__ __forceinline__ void pack_tc_entry(const PackEntry& e, const float* __restrict__ w, __nv_bfloat16* __restrict__ p);
__global__ void pack_tc_multi(const PackTable t) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const PackEntry& e = t.e[blockIdx.y];
  pack_tc_entry(e, e.w, e.p);
}
// The same for a whole run of identical blocks (blockIdx.z = block): the shapes are shared, the pointers come per
// block — one launch instead of one per block and direction (96 launches of ~4.5 us in the training step).
constexpr int kTrunkPackMax = 64 * (WS_RDB_MAX_CONVS + 1);
struct TrunkPackTable {
  int per_block;
  PackEntry shape[WS_RDB_MAX_CONVS + 1];
  const float* w[kTrunkPackMax];
  void* p[kTrunkPackMax];
};
__global__ void pack_tc_trunk(const __grid_constant__ TrunkPackTable t) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int idx = (int)blockIdx.z * t.per_block + (int)blockIdx.y;
  pack_tc_entry(t.shape[blockIdx.y], t.w[idx], (__nv_bfloat16*)t.p[idx]);
}
__device__ __forceinline__ void pack_tc_entry(const PackEntry& e, const float* __restrict__ w, __nv_bfloat16* __restrict__ p) {
  const long long total = (long long)e.taps * e.rows_pad * e.cols_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % e.cols_pad);
    const long long r = i / e.cols_pad;
    const int row = (int)(r % e.rows_pad);
    const int tp = (int)(r / e.rows_pad);
    float v = 0.f;
    if (e.fold) {
      // e.taps = ky*kz here; the full kernel has fold*e.taps taps, ordered (kx, ky, kz)
      const int dx = row / e.cout, co = row - dx * e.cout;
      const int full_tap = e.fold_z ? tp * e.fold + dx : dx * e.taps + tp;
      if (dx < e.fold && col < e.cin) v = w[((long long)co * e.cin + col) * (e.fold * e.taps) + full_tap];
    } else if (!e.dgrad) {
      if (row < e.cout && col < e.cin) v = w[((long long)row * e.cin + col) * e.taps + tp];
    } else {
      if (row < e.cin && col < e.cout) v = w[((long long)col * e.cin + row) * e.taps + (e.taps - 1 - tp)];
    }
    if (e.tf32) reinterpret_cast<float*>(p)[i] = round_tf32(v);
    else p[i] = __float2bfloat16_rn(v);
  }
}

// One parity class of a STRIDED data-gradient (conv_tc*.cu + api.cu: tc_strided_dgrad): taps (tu,tv,tl) of the
// class map to original taps i = i0 + s*(nt-1-t) per axis.  p[t][ci_pad16][co_pad8] = w[co][ci][ix][iy][iz].
__global__ void pack_tc_class(const float* __restrict__ w, __nv_bfloat16* __restrict__ p, int cout, int cin, int kx,
                              int ky, int kz, int ntx, int nty, int ntz, int i0x, int i0y, int i0z, int sx, int sy,
                              int sz, int rows_pad, int cols_pad) {
  const int taps = ntx * nty * ntz;
  long long total = (long long)taps * rows_pad * cols_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int col = (int)(i % cols_pad);
    long long r = i / cols_pad;
    int row = (int)(r % rows_pad);
    int t = (int)(r / rows_pad);
    int tl = t % ntz, tv = (t / ntz) % nty, tu = t / (ntz * nty);
    int ix = i0x + sx * (ntx - 1 - tu), iy = i0y + sy * (nty - 1 - tv), iz = i0z + sz * (ntz - 1 - tl);
    float v = 0.f;
    if (row < cin && col < cout) v = w[(((long long)col * cin + row) * kx + ix) * ky * kz + iy * kz + iz];
    p[i] = __float2bfloat16_rn(v);
  }
}

// ---- copies ------------------------------------------------------------------------------------------
// generic strided copy; thread index runs over (n, v, c) with c fastest when dst is channels-last,
// and over (n, c, v) with v fastest when dst is NCXYZ, so the WRITE side is always coalesced.
__global__ void copy_kernel(View src, View dst, int n, int c, long long v, int c_fastest) {
  long long total = (long long)n * c * v;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % v;
      nn = (int)(r / v);
    } else {
      vv = i % v;
      long long r = i / v;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    dst.st(nn, cc, vv, src.ld(nn, cc, vv));
  }
}

// both channels-last, same dtype, 16-byte vectors
__global__ void copy_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long rows,
                                int vec_per_row, long long src_row_vecs, long long dst_row_vecs) {
  long long total = rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / vec_per_row;
    int q = (int)(i % vec_per_row);
    dst[r * dst_row_vecs + q] = src[r * src_row_vecs + q];
  }
}

// ---- 8-channel vector helpers for the channels-last elementwise fast paths -------------------------------------
__device__ __forceinline__ void ld8(const View& t, long long o, float (&f)[8]) {
  if (t.dtype == WS_F32) {
    const float4 a = *reinterpret_cast<const float4*>((const float*)t.ptr + o);
    const float4 b = *reinterpret_cast<const float4*>((const float*)t.ptr + o + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 r = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)t.ptr + o);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[2 * j] = __uint_as_float(w[j] << 16);
      f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ void st8(const View& t, long long o, const float (&f)[8]) {
  if (t.dtype == WS_F32) {
    *reinterpret_cast<float4*>((float*)t.ptr + o) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>((float*)t.ptr + o + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>((__nv_bfloat16*)t.ptr + o) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// host: can `t` be addressed as 8-channel vectors (channels contiguous, every row start 16/32-byte aligned)?
inline bool vec8_ok(const View& t, int c) {
  if (!t.ptr) return true;
  const int al = t.dtype == WS_F32 ? 4 : 8;  // elements per 16 bytes
  return t.cs == 1 && c % 8 == 0 && t.vs % al == 0 && t.ns % al == 0 && ((uintptr_t)t.ptr % 16) == 0;
}

// y = a*x1 (+ b*x2), all channels-last, 8 channels per thread
__global__ void axpby_cl8_kernel(View x1, float a, View x2, float b, View y, int c8, long long v, long long total) {
  // programmatic dependent launch (common.cuh launch_pdl): start early, wait for the predecessor's data here
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % c8);
    const long long r = i / c8;
    const long long vv = r % v;
    const int nn = (int)(r / v);
    float f[8];
    ld8(x1, x1.off(nn, q * 8, vv), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= a;
    if (x2.ptr) {
      float h[8];
      ld8(x2, x2.off(nn, q * 8, vv), h);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += b * h[j];
    }
    st8(y, y.off(nn, q * 8, vv), f);
  }
}

// g = dy * lrelu'(y) [* chan_scale[n][c]] [* oscale[c]], all channels-last, 8 channels per thread
__global__ void lrelu_bwd_cl8_kernel(View dy, View yv, float slope, const float* __restrict__ chan_scale,
                                     const float* __restrict__ oscale, View g, int c8, long long v, long long total) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % c8);
    const long long r = i / c8;
    const long long vv = r % v;
    const int nn = (int)(r / v);
    float d[8], o[8];
    ld8(dy, dy.off(nn, q * 8, vv), d);
    ld8(yv, yv.off(nn, q * 8, vv), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : slope * d[j];
    if (chan_scale) {
      const float* cs = chan_scale + (long long)nn * c8 * 8 + q * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] *= cs[j];
    }
    if (oscale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] *= oscale[q * 8 + j];
    }
    st8(g, g.off(nn, q * 8, vv), d);
  }
}

__global__ void axpby_kernel(View x1, float a, View x2, float b, View y, int n, int c, long long v,
                             int c_fastest) {
  long long total = (long long)n * c * v;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % v;
      nn = (int)(r / v);
    } else {
      vv = i % v;
      long long r = i / v;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    float t = a * x1.ld(nn, cc, vv);
    if (x2.ptr) t += b * x2.ld(nn, cc, vv);
    y.st(nn, cc, vv, t);
  }
}

__global__ void lrelu_bwd_kernel(View dy, View yv, float slope, const float* __restrict__ chan_scale,
                                 const float* __restrict__ oscale, View g, int n, int c, long long v,
                                 int c_fastest) {
  long long total = (long long)n * c * v;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % v;
      nn = (int)(r / v);
    } else {
      vv = i % v;
      long long r = i / v;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    float d = dy.ld(nn, cc, vv);
    float o = yv.ld(nn, cc, vv);
    float r = o > 0.f ? d : slope * d;
    if (chan_scale) r *= chan_scale[(long long)nn * c + cc];
    if (oscale) r *= oscale[cc];
    g.st(nn, cc, vv, r);
  }
}

// ---- nearest upsample x2 in (x, y) -------------------------------------------------------------------
// Channels-last fast path: one 16-byte vector of channels per thread, each input vector is read once and
// written to its 2x2 block: algorithmic traffic = 1 read + 4 writes per element.
__global__ void upsample_fwd_vec(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int x, int y,
                                 int z, int vec_per_vox, long long in_vs, long long out_vs, long long in_ns,
                                 long long out_ns) {
  long long total = (long long)n * x * y * z * vec_per_vox;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % vec_per_vox);
    long long r = i / vec_per_vox;
    int zz = (int)(r % z);
    r /= z;
    int yy = (int)(r % y);
    r /= y;
    int xx = (int)(r % x);
    int nn = (int)(r / x);
    uint4 val = in[nn * in_ns + (((long long)xx * y + yy) * z + zz) * in_vs + q];
    long long Y2 = 2LL * y;
    long long base = nn * out_ns;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
        out[base + (((2LL * xx + dx) * Y2 + 2 * yy + dy) * z + zz) * out_vs + q] = val;
  }
}

__global__ void upsample_fwd_generic(View in, View out, int n, int c, int x, int y, int z, int c_fastest) {
  long long vo = 4LL * x * y * z;
  long long total = (long long)n * c * vo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % vo;
      nn = (int)(r / vo);
    } else {
      vv = i % vo;
      long long r = i / vo;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    int zz = (int)(vv % z);
    long long r2 = vv / z;
    int yy = (int)(r2 % (2 * y));
    int xx = (int)(r2 / (2 * y));
    long long vi = (((long long)(xx >> 1)) * y + (yy >> 1)) * z + zz;
    // raw element move keeps the bits exact for equal dtypes
    if (in.dtype == out.dtype) {
      if (in.dtype == WS_F32) ((float*)out.ptr)[out.off(nn, cc, vv)] = ((const float*)in.ptr)[in.off(nn, cc, vi)];
      else ((uint16_t*)out.ptr)[out.off(nn, cc, vv)] = ((const uint16_t*)in.ptr)[in.off(nn, cc, vi)];
    } else {
      out.st(nn, cc, vv, in.ld(nn, cc, vi));
    }
  }
}

__global__ void upsample_bwd_generic(View dout, View din, int n, int c, int x, int y, int z, int c_fastest) {
  long long vi_total = (long long)x * y * z;
  long long total = (long long)n * c * vi_total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % vi_total;
      nn = (int)(r / vi_total);
    } else {
      vv = i % vi_total;
      long long r = i / vi_total;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    int zz = (int)(vv % z);
    long long r2 = vv / z;
    int yy = (int)(r2 % y);
    int xx = (int)(r2 / y);
    long long Y2 = 2LL * y;
    float s = 0.f;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
        s += dout.ld(nn, cc, ((2LL * xx + dx) * Y2 + 2 * yy + dy) * z + zz);
    din.st(nn, cc, vv, s);
  }
}

// channels-last both sides: one 8-channel vector of din per thread = sum of the 2x2 block of dout
__global__ void upsample_bwd_cl8(View dout, View din, int c8, int x, int y, int z, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % c8);
    long long r = i / c8;
    const int zz = (int)(r % z); r /= z;
    const int yy = (int)(r % y); r /= y;
    const int xx = (int)(r % x);
    const int nn = (int)(r / x);
    const long long Y2 = 2LL * y;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        float f[8];
        ld8(dout, dout.off(nn, q * 8, ((2LL * xx + dx) * Y2 + 2 * yy + dy) * z + zz), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
    st8(din, din.off(nn, q * 8, ((long long)xx * y + yy) * z + zz), s);
  }
}

// ---- x-fold (narrow-output conv, see windsr.h) -----------------------------------------------------------
__global__ void xfold_sum_kernel(View y, const float* __restrict__ bias, View out, int n, int co, int kx, int pad,
                                 int X, int Y, int Z) {
  const long long V = (long long)X * Y * Z;
  const long long total = (long long)n * co * V;
  const long long sx = (long long)Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const int c = (int)((i / V) % co);
    const int nn = (int)(i / (V * co));
    const int xx = (int)(v / sx);
    float acc = bias ? bias[c] : 0.f;
    for (int dx = 0; dx < kx; ++dx) {
      const int xs = xx + dx - pad;
      if (xs >= 0 && xs < X) acc += y.ld(nn, dx * co + c, v + (long long)(dx - pad) * sx);
    }
    out.st(nn, c, v, acc);
  }
}

__global__ void xunfold_kernel(View dout, View u, int n, int co, int kx, int pad, int cpad, int X, int Y, int Z) {
  const long long V = (long long)X * Y * Z;
  const long long total = (long long)n * V * cpad;
  const long long sx = (long long)Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cpad);
    const long long r = i / cpad;
    const long long v = r % V;
    const int nn = (int)(r / V);
    float val = 0.f;
    if (ch < kx * co) {
      const int dx = ch / co, c = ch % co;
      const int xx = (int)(v / sx);
      const int xs = xx - dx + pad;
      if (xs >= 0 && xs < X) val = dout.ld(nn, c, v + (long long)(pad - dx) * sx);
    }
    u.st(nn, ch, v, val);
  }
}

// y[x] = lrelu(sum_dx U[x + dx - pad][dx*co + c]) for a channels-last fp32 U and output, 8 channels per thread:
// the tail of the x-folded dense convs of a residual dense block (api.cu: ws_rdb_forward)
__global__ void xfold_sum_lrelu_cl8(View u, View out, float slope, int co8, int co, int kx, int pad, int X, int Y, int Z,
                                    long long total) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long V = (long long)X * Y * Z;
  const long long sx = (long long)Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % co8);
    const long long r = i / co8;
    const long long v = r % V;
    const int nn = (int)(r / V);
    const int xx = (int)(v / sx);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int dx = 0; dx < kx; ++dx) {
      const int xs = xx + dx - pad;
      if (xs < 0 || xs >= X) continue;
      float f[8];
      ld8(u, u.off(nn, dx * co + q * 8, v + (long long)(dx - pad) * sx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = acc[j] > 0.f ? acc[j] : slope * acc[j];
    st8(out, out.off(nn, q * 8, v), acc);
  }
}

// ---- xy-fold: BOTH lateral tap axes folded into the channel dimension (hr_convs.2: 5x5x5, 144 -> 3) ------------------
// Y[x', y', z, (dx*ky + dy)*co + c] = sum_{dz, ci} W[c, ci, dx, dy, dz] * in[x', y', z + dz - pz, ci] is a (1,1,kz) conv
// with kx*ky*co (75 -> 80) output channels — 25x fewer MMAs than the direct form, 5x fewer than the x-fold — and
//   out[x, y, z, c] = bias[c] + sum_{dx, dy} Y[x + dx - px, y + dy - py, z, (dx*ky + dy)*co + c].
// One thread per output voxel: every (dx, dy) neighbour contributes `co` consecutive floats of its Y row.
__global__ void xyfold_sum_kernel(View y, const float* __restrict__ bias, View out, int n, int co, int kx, int ky, int px,
                                  int py, int X, int Y, int Z) {
  const long long V = (long long)X * Y * Z;
  const long long total = (long long)n * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const int nn = (int)(i / V);
    const int zz = (int)(v % Z), yy = (int)((v / Z) % Y), xx = (int)(v / ((long long)Z * Y));
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (bias && c < co) ? bias[c] : 0.f;
    for (int dx = 0; dx < kx; ++dx) {
      const int xs = xx + dx - px;
      if (xs < 0 || xs >= X) continue;
      for (int dy = 0; dy < ky; ++dy) {
        const int ys = yy + dy - py;
        if (ys < 0 || ys >= Y) continue;
        const long long vs = ((long long)xs * Y + ys) * Z + zz;
        const long long o = y.off(nn, (dx * ky + dy) * co, vs);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < co) acc[c] += y.ld(o + c * y.cs);
      }
    }
    for (int c = 0; c < co && c < 8; ++c) out.st(nn, c, v, acc[c]);
  }
}
// U[x', y', z, (dx*ky + dy)*co + c] = dout[x' - dx + px, y' - dy + py, z, c] (zero outside, zero pad channels): the operand
// of the data- and weight-gradient of the folded conv.  One thread per 8 consecutive U channels (one 16-byte store).
__global__ void xyunfold_st8_kernel(View dout, View u, int co, int kx, int ky, int px, int py, int cpad8, int X, int Y,
                                    int Z, long long total) {
  const long long V = (long long)X * Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cpad8);
    const long long r = i / cpad8;
    const long long v = r % V;
    const int nn = (int)(r / V);
    const int zz = (int)(v % Z), yy = (int)((v / Z) % Y), xx = (int)(v / ((long long)Z * Y));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = q * 8 + j;
      const int tap = ch / co, c = ch - tap * co;
      float val = 0.f;
      if (tap < kx * ky) {
        const int dx = tap / ky, dy = tap - dx * ky;
        const int xs = xx - dx + px, ys = yy - dy + py;
        if (xs >= 0 && xs < X && ys >= 0 && ys < Y) val = dout.ld(nn, c, ((long long)xs * Y + ys) * Z + zz);
      }
      f[j] = val;
    }
    st8(u, u.off(nn, q * 8, v), f);
  }
}

// ---- im2col for the narrow first layers -------------------------------------------------------------------
// u[n, v, tap*cin + ci] = x[n, ci, v (+) tap] (zero outside the volume and in the pad columns), u channels-last bf16
// with cpad % 8 == 0 columns: turns a Cin <= 4 conv (D's features.0: 3 -> 32 at 128x128x10) into a 1x1x1 conv with
// K = taps*cin (81 -> 96) that runs on the tensor-core kernels; one 8-column vector store per thread.
__global__ void im2col_small_kernel(View x, View u, ConvGeom g, int cpad8, long long total) {
  // one thread per output voxel: walks the taps once (bounds and offsets resolved per (ti, tj) line), gathers the
  // taps*cin values into 8-wide groups and stores each group as one 16-byte vector of the voxel's U row
  const long long VO = (long long)g.xo * g.yo * g.zo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % VO;
    const int nn = (int)(i / VO);
    const int zo = (int)(v % g.zo);
    const int yo = (int)((v / g.zo) % g.yo);
    const int xo = (int)(v / ((long long)g.zo * g.yo));
    const long long urow = u.off(nn, 0, v);
    float f[8];
    int fill = 0, q = 0;
    for (int ti = 0; ti < g.kx; ++ti) {
      const int xi = xo * g.sx - g.px + ti;
      for (int tj = 0; tj < g.ky; ++tj) {
        const int yi = yo * g.sy - g.py + tj;
        const bool line_ok = xi >= 0 && xi < g.x && yi >= 0 && yi < g.y;
        const long long line = ((long long)xi * g.y + yi) * g.z;
        for (int tl = 0; tl < g.kz; ++tl) {
          const int zi = zo * g.sz - g.pz + tl;
          const bool ok = line_ok && zi >= 0 && zi < g.z;
          for (int ci = 0; ci < g.cin; ++ci) {
            f[fill++] = ok ? x.ld(nn, ci, line + zi) : 0.f;
            if (fill == 8) {
              st8(u, urow + q * 8, f);
              ++q;
              fill = 0;
            }
          }
        }
      }
    }
    // tail group and the zero pad columns
    for (; q < cpad8; ++q) {
      for (; fill < 8; ++fill) f[fill] = 0.f;
      st8(u, urow + q * 8, f);
      fill = 0;
    }
  }
}

// same, u channels-last with cpad % 8 == 0: one 8-channel vector store per thread
__global__ void xunfold_st8_kernel(View dout, View u, int co, int kx, int pad, int cpad8, int X, int Y, int Z,
                                   long long total, int src_vec) {
  const long long V = (long long)X * Y * Z;
  const long long sx = (long long)Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cpad8);
    const long long r = i / cpad8;
    const long long v = r % V;
    const int nn = (int)(r / V);
    const int xx = (int)(v / sx);
    float f[8];
    if (src_vec) {
      // co % 8 == 0 and a channels-last source: the 8 channels of this chunk come from ONE shifted voxel
      const int ch0 = q * 8;
      const int dx = ch0 / co, c0 = ch0 - dx * co;
      const int xs = xx - dx + pad;
      if (ch0 < kx * co && xs >= 0 && xs < X) {
        ld8(dout, dout.off(nn, c0, v + (long long)(pad - dx) * sx), f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      st8(u, u.off(nn, q * 8, v), f);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = q * 8 + j;
      float val = 0.f;
      if (ch < kx * co) {
        const int dx = ch / co, c = ch - dx * co;
        const int xs = xx - dx + pad;
        if (xs >= 0 && xs < X) val = dout.ld(nn, c, v + (long long)(pad - dx) * sx);
      }
      f[j] = val;
    }
    st8(u, u.off(nn, q * 8, v), f);
  }
}

// ---- BatchNorm --------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* sum, const float* sqsum, long long count, int c,
                                   const float* gamma, const float* beta, float eps, float momentum,
                                   float* running_mean, float* running_var, float* scale, float* shift,
                                   float* save_mean, float* save_invstd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  double cnt = (double)count;
  double mean = (double)sum[i] / cnt;
  double var = (double)sqsum[i] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  float invstd = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  scale[i] = g * invstd;
  shift[i] = b - (float)mean * g * invstd;
  if (save_mean) save_mean[i] = (float)mean;
  if (save_invstd) save_invstd[i] = invstd;
  if (running_mean) {
    double unbiased = count > 1 ? var * cnt / (cnt - 1.0) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)mean;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
}

__global__ void scale_shift_lrelu_kernel(View x, const float* __restrict__ scale,
                                         const float* __restrict__ shift, float slope, View y, int n, int c,
                                         long long v, int c_fastest) {
  long long total = (long long)n * c * v;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % v;
      nn = (int)(r / v);
    } else {
      vv = i % v;
      long long r = i / v;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    float t = x.ld(nn, cc, vv) * scale[cc] + shift[cc];
    y.st(nn, cc, vv, t > 0.f ? t : slope * t);
  }
}

// One block per (channel, slice): partial sums of g and g*xhat, then atomics.
__global__ void bn_bwd_reduce_kernel(View dy, View yv, View x, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, float slope, float* sum_g,
                                     float* sum_gx, int n, int c, long long v, int slices) {
  int ch = blockIdx.x;
  int sl = blockIdx.y;
  long long total = (long long)n * v;
  long long per = (total + slices - 1) / slices;
  long long beg = sl * per, end = beg + per < total ? beg + per : total;
  float m = mean[ch], is = invstd[ch];
  float sg = 0.f, sgx = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    int nn = (int)(i / v);
    long long vv = i % v;
    float d = dy.ld(nn, ch, vv);
    float o = yv.ld(nn, ch, vv);
    float g = o > 0.f ? d : slope * d;
    float xh = (x.ld(nn, ch, vv) - m) * is;
    sg += g;
    sgx += g * xh;
  }
  __shared__ float r1[32], r2[32];
  sg = warp_sum(sg);
  sgx = warp_sum(sgx);
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = sg; r2[threadIdx.x >> 5] = sgx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < (blockDim.x >> 5) ? r1[threadIdx.x] : 0.f;
    float b = threadIdx.x < (blockDim.x >> 5) ? r2[threadIdx.x] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) { atomicAdd(&sum_g[ch], a); atomicAdd(&sum_gx[ch], b); }
  }
}

__global__ void bn_bwd_apply_kernel(View dy, View yv, View x, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ sum_g, const float* __restrict__ sum_gx,
                                    float slope, float inv_count, View dx, int n, int c, long long v,
                                    int c_fastest) {
  long long total = (long long)n * c * v;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int nn, cc;
    long long vv;
    if (c_fastest) {
      cc = (int)(i % c);
      long long r = i / c;
      vv = r % v;
      nn = (int)(r / v);
    } else {
      vv = i % v;
      long long r = i / v;
      cc = (int)(r % c);
      nn = (int)(r / c);
    }
    float d = dy.ld(nn, cc, vv);
    float o = yv.ld(nn, cc, vv);
    float g = o > 0.f ? d : slope * d;
    float is = invstd[cc];
    float xh = (x.ld(nn, cc, vv) - mean[cc]) * is;
    float gm = gamma ? gamma[cc] : 1.f;
    dx.st(nn, cc, vv, gm * is * (g - sum_g[cc] * inv_count - xh * sum_gx[cc] * inv_count));
  }
}

}  // namespace

int pack_weights_launch(const float* w, const ConvGeom& g, int kind, void* packed, cudaStream_t st) {
  int taps = g.taps();
  if (kind == WS_PACK_SIMT_FWD || kind == WS_PACK_SIMT_DGRAD) {
    long long total = (long long)g.cout * g.cin * taps;
    pack_simt<<<grid_for(total), kBlock, 0, st>>>(w, (float*)packed, g.cout, g.cin, taps,
                                                 kind == WS_PACK_SIMT_DGRAD);
  } else if (kind == WS_PACK_TC_DGRAD && (g.sx != 1 || g.sy != 1 || g.sz != 1)) {
    // strided conv: one block of taps per parity class of the input grid, classes in (pa, pb, pc) order
    const int rows = (g.cin + 15) / 16 * 16, cols = (g.cout + 7) / 8 * 8;
    __nv_bfloat16* out = (__nv_bfloat16*)packed;
    for (int pa = 0; pa < g.sx; ++pa)
      for (int pb = 0; pb < g.sy; ++pb)
        for (int pc = 0; pc < g.sz; ++pc) {
          const int i0x = (pa + g.px) % g.sx, i0y = (pb + g.py) % g.sy, i0z = (pc + g.pz) % g.sz;
          const int ntx = i0x < g.kx ? (g.kx - i0x + g.sx - 1) / g.sx : 0;
          const int nty = i0y < g.ky ? (g.ky - i0y + g.sy - 1) / g.sy : 0;
          const int ntz = i0z < g.kz ? (g.kz - i0z + g.sz - 1) / g.sz : 0;
          const long long total = (long long)ntx * nty * ntz * rows * cols;
          if (total == 0) continue;
          pack_tc_class<<<grid_for(total), kBlock, 0, st>>>(w, out, g.cout, g.cin, g.kx, g.ky, g.kz, ntx, nty, ntz,
                                                           i0x, i0y, i0z, g.sx, g.sy, g.sz, rows, cols);
          WS_POST_LAUNCH(1);
          out += total;
        }
    return 0;
  } else if (kind == WS_PACK_TC_FWD_TF32 || kind == WS_PACK_TC_DGRAD_TF32) {
    int dgrad = kind == WS_PACK_TC_DGRAD_TF32;
    int rows = dgrad ? (g.cin + 15) / 16 * 16 : (g.cout + 15) / 16 * 16;
    int cols = dgrad ? (g.cout + 3) / 4 * 4 : (g.cin + 3) / 4 * 4;
    long long total = (long long)taps * rows * cols;
    pack_tc<float><<<grid_for(total), kBlock, 0, st>>>(w, (float*)packed, g.cout, g.cin, taps, rows, cols, dgrad);
  } else {
    int dgrad = kind == WS_PACK_TC_DGRAD;
    int rows = dgrad ? (g.cin + 15) / 16 * 16 : (g.cout + 15) / 16 * 16;
    int cols = dgrad ? (g.cout + 7) / 8 * 8 : (g.cin + 7) / 8 * 8;
    long long total = (long long)taps * rows * cols;
    pack_tc<__nv_bfloat16><<<grid_for(total), kBlock, 0, st>>>(w, (__nv_bfloat16*)packed, g.cout, g.cin, taps, rows,
                                                              cols, dgrad);
  }
  WS_POST_LAUNCH(1);
  return 0;
}

// stride-1 tensor-core packings of up to WS_RDB_MAX_CONVS + 1 convs in one launch
static long long fill_pack_entry(PackEntry& e, const ConvGeom& g, int dgrad, int fold, int tf32) {
  e.cout = g.cout; e.cin = g.cin; e.taps = g.taps(); e.dgrad = dgrad;
  e.rows_pad = dgrad ? (g.cin + 15) / 16 * 16 : (g.cout + 15) / 16 * 16;
  e.tf32 = tf32 ? 1 : 0;
  if (tf32) e.cols_pad = dgrad ? (g.cout + 3) / 4 * 4 : (g.cin + 3) / 4 * 4;
  else e.cols_pad = dgrad ? (g.cout + 7) / 8 * 8 : (g.cin + 7) / 8 * 8;
  if (fold == 1 && !dgrad) {
    e.fold = g.kx;
    e.taps = g.ky * g.kz;
    e.rows_pad = (g.kx * g.cout + 15) / 16 * 16;
  } else if (fold == 2 && !dgrad) {
    e.fold = g.kz; e.fold_z = 1;
    e.taps = g.kx * g.ky;
    e.rows_pad = (g.kz * g.cout + 15) / 16 * 16;
  }
  return (long long)e.taps * e.rows_pad * e.cols_pad;
}

// all blocks of a trunk in one launch: `per_block` convs of identical geometry per block, pointers [block][conv]
int pack_tc_trunk_launch(int nblocks, int per_block, const ConvGeom* g, const int* fold, int dgrad,
                         const float* const* w, void* const* packed, cudaStream_t st, int tf32) {
  if (nblocks <= 0 || per_block <= 0) return 0;
  WS_REQUIRE(per_block <= WS_RDB_MAX_CONVS + 1 && nblocks * per_block <= kTrunkPackMax, "pack_tc_trunk: too many entries");
  static TrunkPackTable t;  // 10 KB: not on the stack; launches copy it (single-threaded host use, like the rest of the API)
  memset(&t, 0, sizeof(t));
  t.per_block = per_block;
  long long most = 0;
  for (int i = 0; i < per_block; ++i) {
    const long long total = fill_pack_entry(t.shape[i], g[i], dgrad, fold ? fold[i] : 0, tf32);
    if (total > most) most = total;
  }
  for (int i = 0; i < nblocks * per_block; ++i) { t.w[i] = w[i]; t.p[i] = packed[i]; }
  int bx = (int)((most + kBlock - 1) / kBlock);
  if (bx > 16) bx = 16;  // nblocks * per_block blocks of work already fill the machine
  WS_CHECK_CUDA(launch_pdl(pack_tc_trunk, dim3((unsigned)bx, (unsigned)per_block, (unsigned)nblocks), dim3(kBlock), 0, st,
                           1, t));
  WS_POST_LAUNCH(1);
  return 0;
}

int pack_tc_batch_launch(int n, const float* const* w, const ConvGeom* g, int dgrad, void* const* packed,
                         cudaStream_t st, const int* fold, int tf32) {
  if (n <= 0) return 0;
  WS_REQUIRE(n <= WS_RDB_MAX_CONVS + 1, "pack_tc_batch: too many entries");
  PackTable t;
  memset(&t, 0, sizeof(t));
  t.n = n;
  long long most = 0;
  for (int i = 0; i < n; ++i) {
    PackEntry& e = t.e[i];
    e.w = w[i]; e.p = (__nv_bfloat16*)packed[i];
    const long long total = fill_pack_entry(e, g[i], dgrad, fold ? fold[i] : 0, tf32);
    if (total > most) most = total;
  }
  int bx = (int)((most + kBlock - 1) / kBlock);
  if (bx > 148 * 2) bx = 148 * 2;
  WS_CHECK_CUDA(launch_pdl(pack_tc_multi, dim3((unsigned)bx, (unsigned)n), dim3(kBlock), 0, st, 1, t));
  WS_POST_LAUNCH(1);
  return 0;
}

// in place: x = nearest TF32 value of x (fp32, contiguous channels per voxel)
__global__ void round_tf32_kernel(View x, int n, int c, long long v) {
  const long long total = (long long)n * v * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c);
    const long long r = i / c;
    const long long o = x.off((int)(r / v), cc, r % v);
    ((float*)x.ptr)[o] = round_tf32(((const float*)x.ptr)[o]);
  }
}
int round_tf32_launch(const View& x, int n, int c, long long v, cudaStream_t st) {
  if (x.dtype != WS_F32) return 0;
  round_tf32_kernel<<<grid_for((long long)n * v * c), kBlock, 0, st>>>(x, n, c, v);
  WS_POST_LAUNCH(1);
  return 0;
}

static inline int c_fastest_of(const View& dst) { return dst.cs == 1 ? 1 : 0; }

int copy_launch(const View& src, const View& dst, int n, int c, long long v, cudaStream_t st) {
  long long total = (long long)n * c * v;
  if (total <= 0) return 0;
  // vector fast path: channels-last both sides, same dtype, aligned, contiguous batch
  int es = src.esize();
  int per = 16 / es;
  if (src.dtype == dst.dtype && src.cs == 1 && dst.cs == 1 && c % per == 0 && src.vs % per == 0 &&
      dst.vs % per == 0 && ((uintptr_t)src.ptr % 16) == 0 && ((uintptr_t)dst.ptr % 16) == 0 &&
      src.ns == src.vs * v && dst.ns == dst.vs * v) {
    long long rows = (long long)n * v;
    int vpr = c / per;
    copy_vec_kernel<<<grid_for(rows * vpr), kBlock, 0, st>>>((const uint4*)src.ptr, (uint4*)dst.ptr, rows, vpr,
                                                            src.vs / per, dst.vs / per);
  } else if (vec8_ok(src, c) && vec8_ok(dst, c)) {
    // channels-last both sides, converting dtype: the 8-channel axpby kernel with a = 1 (exact)
    WS_CHECK_CUDA(launch_pdl(axpby_cl8_kernel, dim3(grid_for(total / 8)), dim3(kBlock), 0, st, 1, src, 1.f, View(), 0.f,
                             dst, c / 8, v, total / 8));
  } else {
    copy_kernel<<<grid_for(total), kBlock, 0, st>>>(src, dst, n, c, v, c_fastest_of(dst));
  }
  WS_POST_LAUNCH(1);
  return 0;
}

int axpby_launch(const View& x1, float a, const View& x2, float b, const View& y, int n, int c, long long v,
                 cudaStream_t st) {
  long long total = (long long)n * c * v;
  if (total <= 0) return 0;
  if (vec8_ok(x1, c) && vec8_ok(x2, c) && vec8_ok(y, c)) {
    WS_CHECK_CUDA(launch_pdl(axpby_cl8_kernel, dim3(grid_for(total / 8)), dim3(kBlock), 0, st, 1, x1, a, x2, b, y, c / 8,
                             v, total / 8));
    WS_POST_LAUNCH(1);
    return 0;
  }
  axpby_kernel<<<grid_for(total), kBlock, 0, st>>>(x1, a, x2, b, y, n, c, v, c_fastest_of(y));
  WS_POST_LAUNCH(1);
  return 0;
}

int lrelu_bwd_launch(const View& dy, const View& yv, float slope, const float* chan_scale, const float* oscale,
                     const View& g, int n, int c, long long v, cudaStream_t st) {
  long long total = (long long)n * c * v;
  if (total <= 0) return 0;
  if (vec8_ok(dy, c) && vec8_ok(yv, c) && vec8_ok(g, c)) {
    WS_CHECK_CUDA(launch_pdl(lrelu_bwd_cl8_kernel, dim3(grid_for(total / 8)), dim3(kBlock), 0, st, 1, dy, yv, slope,
                             chan_scale, oscale, g, c / 8, v, total / 8));
    WS_POST_LAUNCH(1);
    return 0;
  }
  lrelu_bwd_kernel<<<grid_for(total), kBlock, 0, st>>>(dy, yv, slope, chan_scale, oscale, g, n, c, v,
                                                      c_fastest_of(g));
  WS_POST_LAUNCH(1);
  return 0;
}

int upsample_fwd_launch(const View& in, const View& out, int n, int c, int x, int y, int z, cudaStream_t st) {
  long long vi = (long long)x * y * z;
  if ((long long)n * c * vi <= 0) return 0;
  int es = in.esize();
  int per = 16 / es;
  if (in.dtype == out.dtype && in.cs == 1 && out.cs == 1 && c % per == 0 && in.vs % per == 0 &&
      out.vs % per == 0 && in.ns % per == 0 && out.ns % per == 0 && ((uintptr_t)in.ptr % 16) == 0 &&
      ((uintptr_t)out.ptr % 16) == 0) {
    int vpv = c / per;
    upsample_fwd_vec<<<grid_for((long long)n * vi * vpv), kBlock, 0, st>>>(
        (const uint4*)in.ptr, (uint4*)out.ptr, n, x, y, z, vpv, in.vs / per, out.vs / per, in.ns / per,
        out.ns / per);
  } else {
    upsample_fwd_generic<<<grid_for((long long)n * c * vi * 4), kBlock, 0, st>>>(in, out, n, c, x, y, z,
                                                                                c_fastest_of(out));
  }
  WS_POST_LAUNCH(1);
  return 0;
}

int upsample_bwd_launch(const View& dout, const View& din, int n, int c, int x, int y, int z,
                        cudaStream_t st) {
  long long total = (long long)n * c * x * y * z;
  if (total <= 0) return 0;
  if (vec8_ok(dout, c) && vec8_ok(din, c)) {
    upsample_bwd_cl8<<<grid_for(total / 8), kBlock, 0, st>>>(dout, din, c / 8, x, y, z, total / 8);
    WS_POST_LAUNCH(1);
    return 0;
  }
  upsample_bwd_generic<<<grid_for(total), kBlock, 0, st>>>(dout, din, n, c, x, y, z, c_fastest_of(din));
  WS_POST_LAUNCH(1);
  return 0;
}

int xfold_sum_launch(const View& y, const float* bias, const View& out, int n, int co, int kx, int pad, int X,
                     int Y, int Z, cudaStream_t st) {
  long long total = (long long)n * co * X * Y * Z;
  if (total <= 0) return 0;
  xfold_sum_kernel<<<grid_for(total), kBlock, 0, st>>>(y, bias, out, n, co, kx, pad, X, Y, Z);
  WS_POST_LAUNCH(1);
  return 0;
}

int xfold_sum_lrelu_launch(const View& u, const View& out, float slope, int n, int co, int kx, int pad, int X, int Y,
                           int Z, cudaStream_t st) {
  const long long total = (long long)n * X * Y * Z * (co / 8);
  if (total <= 0) return 0;
  WS_REQUIRE(vec8_ok(u, co) && vec8_ok(out, co), "xfold_sum_lrelu: operands must be 8-channel vectorisable");
  WS_CHECK_CUDA(launch_pdl(xfold_sum_lrelu_cl8, dim3(grid_for(total)), dim3(kBlock), 0, st, 1, u, out, slope, co / 8, co,
                           kx, pad, X, Y, Z, total));
  WS_POST_LAUNCH(1);
  return 0;
}

// fully unrolled K^3 x CIN variant: the gather buffer stays in registers (static indices)
template <int K, int CIN>
__global__ void im2col_k_kernel(View x, View u, ConvGeom g, int cpad8, long long total) {
  constexpr int kCols = K * K * K * CIN;
  constexpr int kGroups = (kCols + 7) / 8;
  const long long VO = (long long)g.xo * g.yo * g.zo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % VO;
    const int nn = (int)(i / VO);
    const int zo = (int)(v % g.zo);
    const int yo = (int)((v / g.zo) % g.yo);
    const int xo = (int)(v / ((long long)g.zo * g.yo));
    const long long urow = u.off(nn, 0, v);
    float f[kGroups * 8];
#pragma unroll
    for (int ti = 0; ti < K; ++ti) {
      const int xi = xo * g.sx - g.px + ti;
#pragma unroll
      for (int tj = 0; tj < K; ++tj) {
        const int yi = yo * g.sy - g.py + tj;
        const bool line_ok = xi >= 0 && xi < g.x && yi >= 0 && yi < g.y;
        const long long line = ((long long)xi * g.y + yi) * g.z;
#pragma unroll
        for (int tl = 0; tl < K; ++tl) {
          const int zi = zo * g.sz - g.pz + tl;
          const bool ok = line_ok && zi >= 0 && zi < g.z;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
            f[((ti * K + tj) * K + tl) * CIN + ci] = ok ? x.ld(nn, ci, line + zi) : 0.f;
        }
      }
    }
#pragma unroll
    for (int c = kCols; c < kGroups * 8; ++c) f[c] = 0.f;
#pragma unroll
    for (int q = 0; q < kGroups; ++q) {
      float h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = f[q * 8 + j];
      st8(u, urow + q * 8, h);
    }
    const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = kGroups; q < cpad8; ++q) st8(u, urow + q * 8, zero);
  }
}

int im2col_small_launch(const View& x, const View& u, const ConvGeom& g, int cpad, cudaStream_t st) {
  WS_REQUIRE(cpad % 8 == 0 && cpad >= g.taps() * g.cin && vec8_ok(u, cpad) && u.dtype == WS_BF16,
             "im2col: u must be channels-last bf16 with a multiple of 8 columns >= taps*cin");
  const long long total = (long long)g.n * g.xo * g.yo * g.zo;
  if (total <= 0) return 0;
  const bool k3 = g.kx == 3 && g.ky == 3 && g.kz == 3;
  if (k3 && g.cin == 3) im2col_k_kernel<3, 3><<<grid_for(total), kBlock, 0, st>>>(x, u, g, cpad / 8, total);
  else if (k3 && g.cin == 4) im2col_k_kernel<3, 4><<<grid_for(total), kBlock, 0, st>>>(x, u, g, cpad / 8, total);
  else if (k3 && g.cin == 2) im2col_k_kernel<3, 2><<<grid_for(total), kBlock, 0, st>>>(x, u, g, cpad / 8, total);
  else im2col_small_kernel<<<grid_for(total), kBlock, 0, st>>>(x, u, g, cpad / 8, total);
  WS_POST_LAUNCH(1);
  return 0;
}

int xunfold_launch(const View& dout, const View& u, int n, int co, int kx, int pad, int cpad, int X, int Y, int Z,
                   cudaStream_t st) {
  long long total = (long long)n * cpad * X * Y * Z;
  if (total <= 0) return 0;
  if (vec8_ok(u, cpad)) {
    const int src_vec = (co % 8 == 0 && vec8_ok(dout, co)) ? 1 : 0;
    xunfold_st8_kernel<<<grid_for(total / 8), kBlock, 0, st>>>(dout, u, co, kx, pad, cpad / 8, X, Y, Z, total / 8,
                                                              src_vec);
    WS_POST_LAUNCH(1);
    return 0;
  }
  xunfold_kernel<<<grid_for(total), kBlock, 0, st>>>(dout, u, n, co, kx, pad, cpad, X, Y, Z);
  WS_POST_LAUNCH(1);
  return 0;
}

int xyfold_sum_launch(const View& y, const float* bias, const View& out, int n, int co, int kx, int ky, int px, int py,
                      int X, int Y, int Z, cudaStream_t st) {
  const long long total = (long long)n * X * Y * Z;
  if (total <= 0) return 0;
  WS_REQUIRE(co <= 8, "xyfold_sum: at most 8 output channels");
  xyfold_sum_kernel<<<grid_for(total), kBlock, 0, st>>>(y, bias, out, n, co, kx, ky, px, py, X, Y, Z);
  WS_POST_LAUNCH(1);
  return 0;
}
int xyunfold_launch(const View& dout, const View& u, int n, int co, int kx, int ky, int px, int py, int cpad, int X,
                    int Y, int Z, cudaStream_t st) {
  const long long total = (long long)n * X * Y * Z * (cpad / 8);
  if (total <= 0) return 0;
  WS_REQUIRE(cpad % 8 == 0 && vec8_ok(u, cpad), "xyunfold: the unfolded operand must be 8-channel vectorisable");
  xyunfold_st8_kernel<<<grid_for(total), kBlock, 0, st>>>(dout, u, co, kx, ky, px, py, cpad / 8, X, Y, Z, total);
  WS_POST_LAUNCH(1);
  return 0;
}

int bn_finalize_launch(const float* sum, const float* sqsum, long long count, int c, const float* gamma,
                       const float* beta, float eps, float momentum, float* rm, float* rv, float* scale,
                       float* shift, float* save_mean, float* save_invstd, cudaStream_t st) {
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(sum, sqsum, count, c, gamma, beta, eps, momentum, rm,
                                                      rv, scale, shift, save_mean, save_invstd);
  WS_POST_LAUNCH(1);
  return 0;
}

int scale_shift_lrelu_launch(const View& x, const float* scale, const float* shift, float slope,
                             const View& y, int n, int c, long long v, cudaStream_t st) {
  long long total = (long long)n * c * v;
  if (total <= 0) return 0;
  scale_shift_lrelu_kernel<<<grid_for(total), kBlock, 0, st>>>(x, scale, shift, slope, y, n, c, v,
                                                              c_fastest_of(y));
  WS_POST_LAUNCH(1);
  return 0;
}

int bn_bwd_reduce_launch(const View& dy, const View& yv, const View& x, const float* mean,
                         const float* invstd, float slope, float* sum_g, float* sum_gx, int n, int c,
                         long long v, cudaStream_t st) {
  WS_CHECK_CUDA(cudaMemsetAsync(sum_g, 0, sizeof(float) * c, st));
  WS_CHECK_CUDA(cudaMemsetAsync(sum_gx, 0, sizeof(float) * c, st));
  long long total = (long long)n * v;
  int slices = (int)((total + 16383) / 16384);
  if (slices < 1) slices = 1;
  if (slices > 64) slices = 64;
  dim3 grid(c, slices);
  bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(dy, yv, x, mean, invstd, slope, sum_g, sum_gx, n, c, v, slices);
  WS_POST_LAUNCH(1);
  return 0;
}

int bn_bwd_apply_launch(const View& dy, const View& yv, const View& x, const float* mean,
                        const float* invstd, const float* gamma, const float* sum_g, const float* sum_gx,
                        float slope, long long count, const View& dx, int n, int c, long long v,
                        cudaStream_t st) {
  long long total = (long long)n * c * v;
  if (total <= 0) return 0;
  bn_bwd_apply_kernel<<<grid_for(total), kBlock, 0, st>>>(dy, yv, x, mean, invstd, gamma, sum_g, sum_gx,
                                                         slope, 1.f / (float)count, dx, n, c, v,
                                                         c_fastest_of(dx));
  WS_POST_LAUNCH(1);
  return 0;
}

}  // namespace ws
