// windloss.cu — the wind-field loss stencils as fused, warp-reduced FP32 kernels.
//
// Reference arithmetic (paths relative to the reference repo):
//   process_data.py:273-298   calculate_div_z          non-uniform 3-point d/dz, spacing from raw altitude Z
//   process_data.py:301-313   calculate_gradient_of_wind_field   torch.gradient over x,y with coordinate
//                             spacing (edge_order=1) + calculate_div_z, concatenated to 9 channels
//   GAN_models/wind_field_GAN_3D.py:773-814  get_norm_factors_of_gradients   8 global maxes
//   GAN_models/wind_field_GAN_3D.py:377-424  pixel L1/L2 + 4 MSE terms on the normalised fields
//
// Because every normaliser is a scalar, MSE(a/m, b/m) = sum (a-b)^2 / (count m^2): one pass over HR, SR
// and Z yields every sum and max without materialising the two 9-channel Jacobians.
// Algorithmic HBM traffic: 3+3+1 fp32 = 28 B per HR voxel forward.
#include "common.cuh"

namespace ws {
namespace {

constexpr int kBlock = 256;

// ---- axis coefficients --------------------------------------------------------------------------------
// coef[6*i + 0..2]: forward form of torch.gradient (a, b, c) for interior points; for the two edge points
//                   slot 1 holds the spacing used as divisor ((f1-f0)/dx), slots 0,2 are 0.
// coef[6*i + 3..5]: the same row as plain linear coefficients on f[i-1], f[i], f[i+1] (for the transpose).
__global__ void axis_coeffs_kernel(const float* __restrict__ x, int len, float* __restrict__ coef) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  float* o = coef + 6 * i;
  if (len < 2) {
    for (int j = 0; j < 6; ++j) o[j] = 0.f;
    return;
  }
  if (i == 0) {
    float dx = x[1] - x[0];
    o[0] = 0.f; o[1] = dx; o[2] = 0.f;
    o[3] = 0.f; o[4] = -1.f / dx; o[5] = 1.f / dx;
  } else if (i == len - 1) {
    float dx = x[len - 1] - x[len - 2];
    o[0] = 0.f; o[1] = dx; o[2] = 0.f;
    o[3] = -1.f / dx; o[4] = 1.f / dx; o[5] = 0.f;
  } else {
    float dx1 = x[i] - x[i - 1];
    float dx2 = x[i + 1] - x[i];
    float a = __fdiv_rn(-dx2, __fmul_rn(dx1, __fadd_rn(dx1, dx2)));
    float b = __fdiv_rn(__fsub_rn(dx2, dx1), __fmul_rn(dx1, dx2));
    float c = __fdiv_rn(dx1, __fmul_rn(dx2, __fadd_rn(dx1, dx2)));
    o[0] = a; o[1] = b; o[2] = c;
    o[3] = a; o[4] = b; o[5] = c;
  }
}

__device__ __forceinline__ float axis_deriv(const float* __restrict__ coef, int i, int len, float fm, float f0,
                                            float fp) {
  const float* o = coef + 6 * i;
  if (i == 0) return __fdiv_rn(__fsub_rn(fp, f0), o[1]);
  if (i == len - 1) return __fdiv_rn(__fsub_rn(f0, fm), o[1]);
  return __fadd_rn(__fadd_rn(__fmul_rn(o[0], fm), __fmul_rn(o[1], f0)), __fmul_rn(o[2], fp));
}

// d/dz exactly as calculate_div_z orders the operations.
__device__ __forceinline__ float z_deriv(int k, int Z, float zm, float z0, float zp, float fm, float f0,
                                         float fp) {
  if (Z < 2) return 0.f;
  if (k == 0) return __fdiv_rn(__fsub_rn(fp, f0), __fsub_rn(zp, z0));
  if (k == Z - 1) return __fdiv_rn(__fsub_rn(f0, fm), __fsub_rn(z0, zm));
  float hl = __fsub_rn(z0, zm), hr = __fsub_rn(zp, z0);
  float hl2 = __fmul_rn(hl, hl), hr2 = __fmul_rn(hr, hr);
  float num = __fsub_rn(__fadd_rn(__fmul_rn(hl2, fp), __fmul_rn(__fsub_rn(hr2, hl2), f0)), __fmul_rn(hr2, fm));
  float den = __fmul_rn(__fmul_rn(hl, hr), __fadd_rn(hl, hr));
  return __fdiv_rn(num, den);
}

// linear coefficients (on f[k-1], f[k], f[k+1]) of row k of the z stencil
__device__ __forceinline__ void z_lin(int k, int Z, float zm, float z0, float zp, float& cm, float& c0,
                                      float& cp) {
  cm = c0 = cp = 0.f;
  if (Z < 2) return;
  if (k == 0) { float d = zp - z0; c0 = -1.f / d; cp = 1.f / d; return; }
  if (k == Z - 1) { float d = z0 - zm; cm = -1.f / d; c0 = 1.f / d; return; }
  float hl = z0 - zm, hr = zp - z0;
  float den = hl * hr * (hl + hr);
  cm = -(hr * hr) / den;
  c0 = (hr * hr - hl * hl) / den;
  cp = (hl * hl) / den;
}

struct Vox {
  int n, x, y, z;
  long long v;
};

__device__ __forceinline__ Vox decode(long long i, int X, int Y, int Z) {
  Vox p;
  long long V = (long long)X * Y * Z;
  p.n = (int)(i / V);
  p.v = i % V;
  p.z = (int)(p.v % Z);
  long long r = p.v / Z;
  p.y = (int)(r % Y);
  p.x = (int)(r / Y);
  return p;
}

// 9 derivatives of a 3-channel field at one voxel. `zl` = altitude at z-1, z, z+1.
__device__ __forceinline__ void jacobian(const View& f, const Vox& p, int X, int Y, int Z,
                                         const float* __restrict__ cx, const float* __restrict__ cy,
                                         const float zl[3], float d[9]) {
  const long long sx = (long long)Y * Z, sy = Z;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f0 = f.ld(p.n, c, p.v);
    float fxm = p.x > 0 ? f.ld(p.n, c, p.v - sx) : 0.f;
    float fxp = p.x < X - 1 ? f.ld(p.n, c, p.v + sx) : 0.f;
    float fym = p.y > 0 ? f.ld(p.n, c, p.v - sy) : 0.f;
    float fyp = p.y < Y - 1 ? f.ld(p.n, c, p.v + sy) : 0.f;
    float fzm = p.z > 0 ? f.ld(p.n, c, p.v - 1) : 0.f;
    float fzp = p.z < Z - 1 ? f.ld(p.n, c, p.v + 1) : 0.f;
    d[c] = X > 1 ? axis_deriv(cx, p.x, X, fxm, f0, fxp) : 0.f;
    d[3 + c] = Y > 1 ? axis_deriv(cy, p.y, Y, fym, f0, fyp) : 0.f;
    d[6 + c] = z_deriv(p.z, Z, zl[0], zl[1], zl[2], fzm, f0, fzp);
  }
}

__device__ __forceinline__ void load_zl(const View& zalt, const Vox& p, int Z, float zl[3]) {
  zl[1] = zalt.ld(p.n, 0, p.v);
  zl[0] = p.z > 0 ? zalt.ld(p.n, 0, p.v - 1) : 0.f;
  zl[2] = p.z < Z - 1 ? zalt.ld(p.n, 0, p.v + 1) : 0.f;
}

__global__ void wind_gradient_kernel(View f, View zalt, const float* __restrict__ cx,
                                     const float* __restrict__ cy, View out, int N, int X, int Y, int Z) {
  long long total = (long long)N * X * Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    Vox p = decode(i, X, Y, Z);
    float zl[3], d[9];
    load_zl(zalt, p, Z, zl);
    jacobian(f, p, X, Y, Z, cx, cy, zl, d);
#pragma unroll
    for (int c = 0; c < 9; ++c) out.st(p.n, c, p.v, d[c]);
  }
}

// ---- fused forward -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int order_bits(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(unsigned int o) {
  unsigned int b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long make_key(float val, unsigned int idx) {
  return ((unsigned long long)order_bits(val) << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  return k;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// scratch layout (device): double sums[6]; unsigned long long keys[8]
struct WlScratch {
  double sums[6];
  unsigned long long keys[8];
};

__global__ void windloss_init_kernel(WlScratch* s) {
  int t = threadIdx.x;
  if (t < 6) s->sums[t] = 0.0;
  if (t < 8) s->keys[t] = 0ull;  // below every real key (order_bits(-inf) = 0x007fffff > 0 ... keys of real data are > 0)
}

__global__ void __launch_bounds__(kBlock)
windloss_fwd_kernel(View hr, View sr, View zalt, const float* __restrict__ cx, const float* __restrict__ cy,
                    int N, int X, int Y, int Z, WlScratch* scratch) {
  const long long V = (long long)X * Y * Z;
  const long long total = (long long)N * V;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  unsigned long long key[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) key[k] = 0ull;

  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    Vox p = decode(i, X, Y, Z);
    float zl[3], dh[9], ds[9];
    load_zl(zalt, p, Z, zl);
    jacobian(hr, p, X, Y, Z, cx, cy, zl, dh);
    jacobian(sr, p, X, Y, Z, cx, cy, zl, ds);
    // flat index base into the virtual (n, 9, v) Jacobian (fits 32 bits for every shipped config)
    unsigned int base = (unsigned int)((long long)p.n * 9 * V + p.v);
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      float e = ds[c] - dh[c];
      unsigned int idx = base + (unsigned int)(c * V);
      if (c < 6) {
        s[0] += e * e;
        unsigned long long kh = make_key(fabsf(dh[c]), idx), ks = make_key(fabsf(ds[c]), idx);
        key[0] = kh > key[0] ? kh : key[0];
        key[1] = ks > key[1] ? ks : key[1];
      } else {
        s[1] += e * e;
        unsigned long long kh = make_key(dh[c], idx), ks = make_key(ds[c], idx);
        key[2] = kh > key[2] ? kh : key[2];
        key[3] = ks > key[3] ? ks : key[3];
      }
    }
    float divh = dh[0] + dh[4] + dh[8], divs = ds[0] + ds[4] + ds[8];
    float dxyh = dh[0] + dh[4], dxys = ds[0] + ds[4];
    float e1 = divs - divh, e2 = dxys - dxyh;
    s[2] += e1 * e1;
    s[3] += e2 * e2;
    unsigned int vidx = (unsigned int)i;
    unsigned long long k;
    k = make_key(fabsf(divh), vidx); key[4] = k > key[4] ? k : key[4];
    k = make_key(fabsf(divs), vidx); key[5] = k > key[5] ? k : key[5];
    k = make_key(fabsf(dxyh), vidx); key[6] = k > key[6] ? k : key[6];
    k = make_key(fabsf(dxys), vidx); key[7] = k > key[7] ? k : key[7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float e = sr.ld(p.n, c, p.v) - hr.ld(p.n, c, p.v);
      s[4] += fabsf(e);
      s[5] += e * e;
    }
  }

  __shared__ double sh_s[kBlock / 32][6];
  __shared__ unsigned long long sh_k[kBlock / 32][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = warp_sum_d((double)s[k]);
    if (lane == 0) sh_s[warp][k] = v;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    unsigned long long v = warp_max_u64(key[k]);
    if (lane == 0) sh_k[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) v += sh_s[w][threadIdx.x];
    atomicAdd(&scratch->sums[threadIdx.x], v);
  } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
    int k = threadIdx.x - 32;
    unsigned long long v = 0ull;
    for (int w = 0; w < kBlock / 32; ++w) v = sh_k[w][k] > v ? sh_k[w][k] : v;
    atomicMax(&scratch->keys[k], v);
  }
}

__global__ void windloss_finalize_kernel(const WlScratch* s, float* result, long long* argmax) {
  int t = threadIdx.x;
  if (t < 6) result[t] = (float)s->sums[t];
  if (t < 8) {
    unsigned long long k = s->keys[t];
    result[WS_WL_MAX_HR_XY + t] = unorder_bits((unsigned int)(k >> 32));
    if (argmax && (t & 1)) argmax[t >> 1] = (long long)(0xffffffffu - (unsigned int)(k & 0xffffffffu));
  }
  if (t == 14 || t == 15) result[t] = 0.f;
}

// ---- backward ------------------------------------------------------------------------------------------
// pass A: G(n, 9, v) = dL/d(SR Jacobian)
__global__ void windloss_bwd_G_kernel(View hr, View sr, View zalt, const float* __restrict__ cx,
                                      const float* __restrict__ cy, int N, int X, int Y, int Z,
                                      const float* __restrict__ coef, const long long* __restrict__ argmax,
                                      float* __restrict__ G) {
  const long long V = (long long)X * Y * Z;
  const long long total = (long long)N * V;
  const float c_xy = 2.f * coef[0], c_z = 2.f * coef[1], c_div = 2.f * coef[2], c_dxy = 2.f * coef[3];
  const float a_xy = coef[6], a_z = coef[7], a_div = coef[8], a_dxy = coef[9];
  const long long i_xy = argmax ? argmax[0] : -1, i_z = argmax ? argmax[1] : -1;
  const long long i_div = argmax ? argmax[2] : -1, i_dxy = argmax ? argmax[3] : -1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    Vox p = decode(i, X, Y, Z);
    float zl[3], dh[9], ds[9], g[9];
    load_zl(zalt, p, Z, zl);
    jacobian(hr, p, X, Y, Z, cx, cy, zl, dh);
    jacobian(sr, p, X, Y, Z, cx, cy, zl, ds);
    float e[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) e[c] = ds[c] - dh[c];
    float div = e[0] + e[4] + e[8], dxy = e[0] + e[4];
    // a zero coefficient means "this term is not in the loss" (e.g. dropped by the NaN guard): its field may be
    // Inf/NaN and must not be multiplied in (0 * inf = NaN)
    auto mul0 = [](float c, float v) { return c == 0.f ? 0.f : c * v; };
#pragma unroll
    for (int c = 0; c < 9; ++c) g[c] = mul0(c < 6 ? c_xy : c_z, e[c]);
    const float gd = mul0(c_div, div), gxy = mul0(c_dxy, dxy);
    g[0] += gd + gxy;
    g[4] += gd + gxy;
    g[8] += gd;
    // normaliser path: the element(s) that attain SR_max receive dL/dSR_max (sign for the |.| maxes)
    long long base = (long long)p.n * 9 * V + p.v;
    if (a_xy != 0.f || a_z != 0.f) {
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        long long idx = base + (long long)c * V;
        if (c < 6 && idx == i_xy) g[c] += a_xy * (ds[c] >= 0.f ? 1.f : -1.f);
        if (c >= 6 && idx == i_z) g[c] += a_z;
      }
    }
    if (a_div != 0.f && i == i_div) {
      float sg = (ds[0] + ds[4] + ds[8]) >= 0.f ? 1.f : -1.f;
      g[0] += a_div * sg; g[4] += a_div * sg; g[8] += a_div * sg;
    }
    if (a_dxy != 0.f && i == i_dxy) {
      float sg = (ds[0] + ds[4]) >= 0.f ? 1.f : -1.f;
      g[0] += a_dxy * sg; g[4] += a_dxy * sg;
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) G[base + (long long)c * V] = g[c];
  }
}

// pass B: dSR = stencil^T G + pixel terms
__global__ void windloss_bwd_apply_kernel(View hr, View sr, View zalt, const float* __restrict__ cx,
                                          const float* __restrict__ cy, int N, int X, int Y, int Z,
                                          const float* __restrict__ coef, const float* __restrict__ G,
                                          View dsr) {
  const long long V = (long long)X * Y * Z;
  const long long total = (long long)N * V;
  const long long sx = (long long)Y * Z, sy = Z;
  const float c_l1 = coef[4], c_l2 = coef[5];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    Vox p = decode(i, X, Y, Z);
    const float* Gn = G + (long long)p.n * 9 * V;
    // z-stencil rows z-1, z, z+1 need altitude at z-2 .. z+2
    float za[5];
#pragma unroll
    for (int j = -2; j <= 2; ++j) {
      int zz = p.z + j;
      za[j + 2] = (zz >= 0 && zz < Z) ? zalt.ld(p.n, 0, p.v + j) : 0.f;
    }
    // coefficient of f[z] in row z' = z + r (r = -1, 0, 1) is lin(z')[1 - r]
    float wz[3];
#pragma unroll
    for (int r = -1; r <= 1; ++r) {
      int zr = p.z + r;
      float w = 0.f;
      if (zr >= 0 && zr < Z) {
        float cm, c0, cp;
        z_lin(zr, Z, za[r + 1], za[r + 2], za[r + 3], cm, c0, cp);
        w = r == -1 ? cp : (r == 0 ? c0 : cm);
      }
      wz[r + 1] = w;
    }
    float wx[3], wy[3];
#pragma unroll
    for (int r = -1; r <= 1; ++r) {
      int xr = p.x + r, yr = p.y + r;
      wx[r + 1] = (X > 1 && xr >= 0 && xr < X) ? cx[6 * xr + 3 + (1 - r)] : 0.f;
      wy[r + 1] = (Y > 1 && yr >= 0 && yr < Y) ? cy[6 * yr + 3 + (1 - r)] : 0.f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int r = -1; r <= 1; ++r) {
        // G == 0 means "no loss term touches this derivative": skip it even if the stencil weight is Inf
        // (degenerate grid spacing), 0 * inf must not poison the gradient
        if (wx[r + 1] != 0.f) { float gv = Gn[(long long)c * V + p.v + r * sx]; if (gv != 0.f) acc += wx[r + 1] * gv; }
        if (wy[r + 1] != 0.f) { float gv = Gn[(long long)(3 + c) * V + p.v + r * sy]; if (gv != 0.f) acc += wy[r + 1] * gv; }
        if (wz[r + 1] != 0.f) { float gv = Gn[(long long)(6 + c) * V + p.v + r]; if (gv != 0.f) acc += wz[r + 1] * gv; }
      }
      float e = sr.ld(p.n, c, p.v) - hr.ld(p.n, c, p.v);
      float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
      acc += c_l1 * sg + c_l2 * 2.f * e;
      dsr.st(p.n, c, p.v, acc);
    }
  }
}

inline int grid_for(long long total) {
  long long b = (total + kBlock - 1) / kBlock;
  long long cap = 148LL * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int axis_coeffs_launch(const float* coords, int len, float* coef, cudaStream_t st) {
  if (len <= 0) return 0;
  axis_coeffs_kernel<<<(len + 127) / 128, 128, 0, st>>>(coords, len, coef);
  WS_POST_LAUNCH(1);
  return 0;
}

int wind_gradient_launch(const View& f, const View& zalt, const float* cx, const float* cy, const View& out,
                         int n, int x, int y, int z, cudaStream_t st) {
  long long total = (long long)n * x * y * z;
  if (total <= 0) return 0;
  wind_gradient_kernel<<<grid_for(total), kBlock, 0, st>>>(f, zalt, cx, cy, out, n, x, y, z);
  WS_POST_LAUNCH(1);
  return 0;
}

int windloss_fwd_launch(const View& hr, const View& sr, const View& zalt, const float* cx, const float* cy,
                        int n, int x, int y, int z, float* result, long long* argmax, cudaStream_t st) {
  long long total = (long long)n * x * y * z;
  WS_REQUIRE((long long)n * 9 * x * y * z < 0xffffffffLL, "windloss: volume too large for 32-bit argmax keys");
  static_assert(sizeof(WlScratch) <= 16 * sizeof(float) * 2 + 64, "scratch layout");
  // scratch lives in result[16..] (ws_windloss_fwd requires WS_WL_SLOTS*2 floats... see api)
  WlScratch* scratch = reinterpret_cast<WlScratch*>(result + WS_WL_SLOTS);
  windloss_init_kernel<<<1, 32, 0, st>>>(scratch);
  if (total > 0)
    windloss_fwd_kernel<<<grid_for(total), kBlock, 0, st>>>(hr, sr, zalt, cx, cy, n, x, y, z, scratch);
  windloss_finalize_kernel<<<1, 32, 0, st>>>(scratch, result, argmax);
  WS_POST_LAUNCH(3);
  return 0;
}

int windloss_bwd_launch(const View& hr, const View& sr, const View& zalt, const float* cx, const float* cy,
                        int n, int x, int y, int z, const float* coef, const long long* argmax,
                        const View& dsr, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  long long total = (long long)n * x * y * z;
  size_t need = (size_t)total * 9 * sizeof(float);
  WS_REQUIRE(workspace && workspace_bytes >= need, "windloss_bwd workspace too small: %zu < %zu",
             workspace_bytes, need);
  if (total <= 0) return 0;
  float* G = (float*)workspace;
  windloss_bwd_G_kernel<<<grid_for(total), kBlock, 0, st>>>(hr, sr, zalt, cx, cy, n, x, y, z, coef, argmax, G);
  windloss_bwd_apply_kernel<<<grid_for(total), kBlock, 0, st>>>(hr, sr, zalt, cx, cy, n, x, y, z, coef, G, dsr);
  WS_POST_LAUNCH(2);
  return 0;
}

}  // namespace ws
