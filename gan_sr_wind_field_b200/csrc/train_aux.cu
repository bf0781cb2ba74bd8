// train_aux.cu — the HBM-bound pieces of the training step around the convolutions (SURVEY §8-f "next" rows and
// row L7), each a single pass over its operands:
//   * multi-tensor Adam with the reference's hyper-parameters and skip-on-non-finite guard
//     (GAN_models/wind_field_GAN_3D.py:151-162, 457-460);
//   * instance noise  x + U[0,1) * scale  (tools/trainingtricks.py:49-58, used wind_field_GAN_3D.py:250-299) with a
//     counter-based Philox4x32-10 generator whose call counter lives on the device (CUDA-graph replays draw fresh noise);
//   * validation metrics: PSNR sums of SR and of the trilinear (align_corners) upsample of LR against HR in one pass
//     (wind_field_GAN_3D.py:730-770, 597-618);
//   * the input pipeline: crop + normalise + LR subsample + rot90 / flip with the wind-component sign fixes
//     (process_data.py:159-262, 420-494) as one bit-exact gather from the float64 source fields.
#include "common.cuh"

namespace ws {
namespace {

constexpr int kBlock = 256;

// ---- Adam -----------------------------------------------------------------------------------------------------
struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* step;
  long long n;
};
static_assert(sizeof(AdamTensor) == sizeof(ws_adam_tensor), "ws_adam_tensor layout");
constexpr int kAdamChunk = 8192;  // elements per block

// Same update as torch.optim.Adam (amsgrad=False, maximize=False):
//   g' = g * grad_scale + wd * p;  m = m + (g' - m)(1 - b1);  v = b2 v + (1 - b2) g'^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   t = step + 1
__global__ void __launch_bounds__(kBlock)
adam_multi_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ chunks, const float* __restrict__ lr_dev,
                  float lr, double b1d, double b2d, float eps, float wd, float grad_scale,
                  const float* __restrict__ found_inf) {
  if (found_inf && *found_inf != 0.f) return;
  const int2 ck = chunks[blockIdx.x];
  const AdamTensor t = tab[ck.x];
  if (lr_dev) lr = *lr_dev;
  // bias corrections and the (1 - beta) factors in double, like torch's host-side Python arithmetic: float(0.999) is
  // 1.3e-5 (relative to 1 - beta2) away from 0.999
  const double stepd = (double)*t.step + 1.0;
  const float b1 = (float)b1d, b2 = (float)b2d;
  const float omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
  const float bc2_sqrt = (float)sqrt(1.0 - pow(b2d, stepd));
  const float step_size = (float)((double)lr / (1.0 - pow(b1d, stepd)));
  const long long lo = (long long)ck.y * kAdamChunk;
  const long long hi = lo + kAdamChunk < t.n ? lo + kAdamChunk : t.n;
  const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g = g * grad_scale + wd * p;
    m = m + (g - m) * omb1;
    v = b2 * v + omb2 * g * g;
    p -= step_size * m / (sqrtf(v) / bc2_sqrt + eps);
  };
  if (vec) {
    const long long lo4 = lo / 4, hi4 = hi / 4;  // lo is a multiple of 4 (kAdamChunk % 4 == 0)
    for (long long i = lo4 + threadIdx.x; i < hi4; i += kBlock) {
      float4 p = reinterpret_cast<float4*>(t.p)[i], m = reinterpret_cast<float4*>(t.m)[i];
      float4 v = reinterpret_cast<float4*>(t.v)[i];
      const float4 g = reinterpret_cast<const float4*>(t.g)[i];
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      reinterpret_cast<float4*>(t.p)[i] = p;
      reinterpret_cast<float4*>(t.m)[i] = m;
      reinterpret_cast<float4*>(t.v)[i] = v;
    }
    for (long long i = hi4 * 4 + threadIdx.x; i < hi; i += kBlock) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kBlock) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
  }
}
__global__ void adam_bump_kernel(const AdamTensor* __restrict__ tab, int n, const float* __restrict__ found_inf) {
  if (found_inf && *found_inf != 0.f) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) *tab[i].step += 1.f;
}

// ---- Philox4x32-10 ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(unsigned int r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }  // [0, 1)

// out = x + U[0,1) * scale; element i uses word (i & 3) of philox(counter = (i / 4, call), key = seed).
// state[0] = call counter (incremented by the last block to finish), state[1] = blocks finished.
__global__ void __launch_bounds__(kBlock)
instance_noise_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, float scale,
                      const float* __restrict__ scale_dev, unsigned long long seed, unsigned long long* state) {
  const unsigned long long call = state[0];
  if (scale_dev) scale *= *scale_dev;
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  const long long n4 = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox(make_uint4((unsigned int)q, (unsigned int)(q >> 32), (unsigned int)call,
                                      (unsigned int)(call >> 32)), key);
    const unsigned int w[4] = {r.x, r.y, r.z, r.w};
    const long long i0 = q * 4;
    if (i0 + 4 <= n && ((((uintptr_t)x | (uintptr_t)out) & 15) == 0)) {
      const float4 a = reinterpret_cast<const float4*>(x)[q];
      reinterpret_cast<float4*>(out)[q] = make_float4(a.x + u01(w[0]) * scale, a.y + u01(w[1]) * scale,
                                                      a.z + u01(w[2]) * scale, a.w + u01(w[3]) * scale);
    } else {
      for (int j = 0; j < 4 && i0 + j < n; ++j) out[i0 + j] = x[i0 + j] + u01(w[j]) * scale;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&state[1], 1ull) == (unsigned long long)gridDim.x - 1) {  // every block has read state[0]
      state[1] = 0ull;
      state[0] = call + 1ull;
    }
  }
}

// ---- validation metrics ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sums[0] = sum (HR-SR)^2, [1] = sum (HR-tri)^2, [2] = sum |HR-tri|, [3] = sum |HR-SR| over (n, 3, v);
// tri = F.interpolate(LR[:, :3], scale_factor=(s,s,1), mode="trilinear", align_corners=True): source coordinate
// dst * (in-1)/(out-1) per axis (z: identity), corner weights as torch's upsample_trilinear3d.
__global__ void __launch_bounds__(kBlock)
metrics_kernel(View hr, View sr, View lr, int N, int X, int Y, int Z, int xl, int yl, double* sums) {
  const long long V = (long long)X * Y * Z, total = (long long)N * 3 * V;
  const float rx = X > 1 ? (float)(xl - 1) / (float)(X - 1) : 0.f;
  const float ry = Y > 1 ? (float)(yl - 1) / (float)(Y - 1) : 0.f;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i % V;
    const int c = (int)((i / V) % 3), n = (int)(i / (3 * V));
    const int z = (int)(v % Z), y = (int)((v / Z) % Y), x = (int)(v / ((long long)Z * Y));
    const float fx = rx * x, fy = ry * y;
    const int x0 = (int)fx, y0 = (int)fy;
    const int x1 = x0 + (x0 < xl - 1 ? 1 : 0), y1 = y0 + (y0 < yl - 1 ? 1 : 0);
    const float lx1 = fx - x0, lx0 = 1.f - lx1, ly1 = fy - y0, ly0 = 1.f - ly1;
    auto at = [&](int a, int b) { return lr.ld(n, c, ((long long)a * yl + b) * Z + z); };
    const float tri = lx0 * (ly0 * at(x0, y0) + ly1 * at(x0, y1)) + lx1 * (ly0 * at(x1, y0) + ly1 * at(x1, y1));
    const float h = hr.ld(n, c, v), q = sr.ptr ? sr.ld(n, c, v) : h;
    const float e0 = h - q, e1 = h - tri;
    s[0] += e0 * e0; s[1] += e1 * e1; s[2] += fabsf(e1); s[3] += fabsf(e0);
  }
  __shared__ double sh[kBlock / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double w = warp_sum_d((double)s[k]);
    if (lane == 0) sh[warp][k] = w;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) t += sh[w][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

// ---- input pipeline -------------------------------------------------------------------------------------------
struct PrepArgs {
  const double* u; const double* v; const double* w; const double* p; const double* z; const double* zag;
  long long sample_stride;  // elements between consecutive samples of each source field
  int SX, SY, SZ;           // source field extents
  int X, Y;                 // output HR extents (the crop: slice_size, or SX / SY)
  int cf;                   // coarseness factor
  int lr_c;                 // LR channels
  int pressure, zchan, above;
  double uvw_max, p_min, p_range, z_min, z_range, zag_max, z_range_above;
  const int* aug;           // per sample: x_start, y_start, rotations (0..3), flip_x, flip_y
};
// Source coordinate of output (a, b) in an (E0 x E1) rot90^k'd, then flipped array (torch.rot90(t, k, [1, 2]) followed
// by torch.flip(t, [1]) / torch.flip(t, [2]), process_data.py:198-262).  Returns the sign the wind components pick up.
__device__ __forceinline__ void aug_source(int a, int b, int E0, int E1, int k, int fx, int fy, int& sa, int& sb) {
  // after rotation the array is (R0 x R1): R = (E0,E1) for even k, (E1,E0) for odd k; the crop is square in practice
  const int R0 = (k & 1) ? E1 : E0, R1 = (k & 1) ? E0 : E1;
  if (fx) a = R0 - 1 - a;
  if (fy) b = R1 - 1 - b;
  switch (k & 3) {
    case 0: sa = a; sb = b; break;
    case 1: sa = b; sb = E1 - 1 - a; break;           // rot90: out[a][b] = in[b][E1-1-a]
    case 2: sa = E0 - 1 - a; sb = E1 - 1 - b; break;
    default: sa = E0 - 1 - b; sb = a; break;          // rot270: out[a][b] = in[E0-1-b][a]
  }
}
// channel c of the normalised (pre-augmentation) sample at source voxel (x, y, z) of the crop
__device__ __forceinline__ float prep_value(const PrepArgs& A, long long base, int c, int x, int y, int z) {
  const long long o = base + ((long long)x * A.SY + y) * A.SZ + z;
  if (c == 0) return (float)(A.u[o] / A.uvw_max);
  if (c == 1) return (float)(A.v[o] / A.uvw_max);
  if (c == 2) return (float)(A.w[o] / A.uvw_max);
  int k = 3;
  if (A.pressure) { if (c == k) return (float)((A.p[o] - A.p_min) / A.p_range); ++k; }
  if (A.zchan && A.above) {
    if (c == k) return (float)(A.zag[o] / A.zag_max);
    return (float)((A.z[o] - A.zag[o] - A.z_min) / A.z_range_above);
  }
  return (float)((A.z[o] - A.z_min) / A.z_range);
}
// The rotation maps (u, v) -> components of the rotated frame (process_data.py:203-244), each flip negates one.
__device__ __forceinline__ float aug_wind(const PrepArgs& A, long long base, int c, int sx, int sy, int z, int k, int fx,
                                          int fy) {
  if (c >= 2) return prep_value(A, base, c, sx, sy, z);
  int src = c;
  float sign = 1.f;
  switch (k & 3) {
    case 1: src = 1 - c; sign = c == 0 ? -1.f : 1.f; break;   // u' = -v, v' = u
    case 2: sign = -1.f; break;                                // u' = -u, v' = -v
    case 3: src = 1 - c; sign = c == 0 ? 1.f : -1.f; break;   // u' = v, v' = -u
    default: break;
  }
  if (fx && c == 0) sign = -sign;
  if (fy && c == 1) sign = -sign;
  return sign * prep_value(A, base, src, sx, sy, z);
}
__global__ void __launch_bounds__(kBlock)
prepare_batch_kernel(PrepArgs A, int N, float* __restrict__ LR, float* __restrict__ HR, float* __restrict__ Zo) {
  const int xl = (A.X + A.cf - 1) / A.cf, yl = (A.Y + A.cf - 1) / A.cf;
  const long long hr_per = 3LL * A.X * A.Y * A.SZ, z_per = (long long)A.X * A.Y * A.SZ;
  const long long lr_per = (long long)A.lr_c * xl * yl * A.SZ;
  const long long per = hr_per + z_per + lr_per, total = per * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    long long r = i - (long long)n * per;
    const int* g = A.aug + 5 * n;
    const int xs = g[0], ys = g[1], k = g[2], fx = g[3], fy = g[4];
    const long long base = (long long)n * A.sample_stride;
    int sa, sb;
    if (r < hr_per) {
      const int z = (int)(r % A.SZ), b = (int)((r / A.SZ) % A.Y), a = (int)((r / ((long long)A.SZ * A.Y)) % A.X);
      const int c = (int)(r / z_per);
      aug_source(a, b, A.X, A.Y, k, fx, fy, sa, sb);
      HR[(long long)n * hr_per + r] = aug_wind(A, base, c, xs + sa, ys + sb, z, k, fx, fy);
    } else if (r < hr_per + z_per) {
      r -= hr_per;
      const int z = (int)(r % A.SZ), b = (int)((r / A.SZ) % A.Y), a = (int)(r / ((long long)A.SZ * A.Y));
      aug_source(a, b, A.X, A.Y, k, fx, fy, sa, sb);
      Zo[(long long)n * z_per + r] = (float)A.z[base + ((long long)(xs + sa) * A.SY + (ys + sb)) * A.SZ + z];
    } else {
      r -= hr_per + z_per;
      const int z = (int)(r % A.SZ), b = (int)((r / A.SZ) % yl), a = (int)((r / ((long long)A.SZ * yl)) % xl);
      const int c = (int)(r / ((long long)xl * yl * A.SZ));
      aug_source(a, b, xl, yl, k, fx, fy, sa, sb);  // the LR grid is augmented on its own (subsampled) index space
      LR[(long long)n * lr_per + r] = aug_wind(A, base, c, xs + sa * A.cf, ys + sb * A.cf, z, k, fx, fy);
    }
  }
}

inline int grid_for(long long total) {
  long long b = (total + kBlock - 1) / kBlock;
  const long long cap = 148LL * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace
}  // namespace ws

using namespace ws;

extern "C" {

int ws_adam_chunk_elems(void) { return kAdamChunk; }

int ws_adam_step(const void* table, const void* chunks, int ntensors, int nchunks, const float* lr_dev, float lr,
                 double beta1, double beta2, float eps, float weight_decay, float grad_scale, const float* found_inf,
                 void* stream) {
  WS_REQUIRE(table && chunks && ntensors > 0 && nchunks > 0, "ws_adam_step: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  adam_multi_kernel<<<nchunks, kBlock, 0, st>>>((const AdamTensor*)table, (const int2*)chunks, lr_dev, lr, beta1,
                                                beta2, eps, weight_decay, grad_scale, found_inf);
  adam_bump_kernel<<<(ntensors + 255) / 256, 256, 0, st>>>((const AdamTensor*)table, ntensors, found_inf);
  WS_POST_LAUNCH(2);
  return 0;
}

int ws_instance_noise(const float* x, float* out, int64_t n, float scale, const float* scale_dev, uint64_t seed,
                      uint64_t* state, void* stream) {
  WS_REQUIRE(x && out && state && n > 0, "ws_instance_noise: bad arguments");
  instance_noise_kernel<<<grid_for((n + 3) / 4), kBlock, 0, (cudaStream_t)stream>>>(
      x, out, (long long)n, scale, scale_dev, (unsigned long long)seed, (unsigned long long*)state);
  WS_POST_LAUNCH(1);
  return 0;
}

int ws_validation_metrics(const ws_tensor* hr, const ws_tensor* sr, const ws_tensor* lr, int n, int x, int y, int z,
                          int xl, int yl, double* sums, void* stream) {
  WS_REQUIRE(hr && hr->ptr && lr && lr->ptr && sums, "ws_validation_metrics: null pointer");
  WS_REQUIRE(x > 0 && y > 0 && z > 0 && xl > 0 && yl > 0, "ws_validation_metrics: bad extents");
  cudaStream_t st = (cudaStream_t)stream;
  WS_CHECK_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), st));
  metrics_kernel<<<grid_for((long long)n * 3 * x * y * z), kBlock, 0, st>>>(View(hr), View(sr), View(lr), n, x, y, z,
                                                                            xl, yl, sums);
  WS_POST_LAUNCH(1);
  return 0;
}

int ws_prepare_batch(const ws_prepare_desc* d, const double* u, const double* v, const double* w, const double* p,
                     const double* z, const double* zag, const int32_t* aug, float* lr, float* hr, float* zout,
                     void* stream) {
  WS_REQUIRE(d && u && v && w && z && aug && lr && hr && zout, "ws_prepare_batch: null pointer");
  WS_REQUIRE(!d->include_pressure || p, "ws_prepare_batch: pressure channel requested without a pressure field");
  WS_REQUIRE(!(d->include_z_channel && d->include_above_ground_channel) || zag,
             "ws_prepare_batch: above-ground channel requested without z_above_ground");
  WS_REQUIRE(d->n > 0 && d->x > 0 && d->y > 0 && d->sz > 0 && d->coarseness > 0 && d->x <= d->sx && d->y <= d->sy,
             "ws_prepare_batch: bad extents");
  PrepArgs A;
  A.u = u; A.v = v; A.w = w; A.p = p; A.z = z; A.zag = zag;
  A.sample_stride = d->sample_stride;
  A.SX = d->sx; A.SY = d->sy; A.SZ = d->sz; A.X = d->x; A.Y = d->y; A.cf = d->coarseness;
  A.pressure = d->include_pressure; A.zchan = d->include_z_channel; A.above = d->include_above_ground_channel;
  A.lr_c = 3 + (A.pressure ? 1 : 0) + (A.zchan ? (A.above ? 2 : 1) : 0);
  A.uvw_max = d->uvw_max; A.p_min = d->p_min; A.p_range = d->p_max - d->p_min;
  A.z_min = d->z_min; A.z_range = d->z_max - d->z_min; A.zag_max = d->z_above_ground_max;
  A.z_range_above = d->z_max - d->z_min - d->z_above_ground_max;
  A.aug = aug;
  const int xl = (A.X + A.cf - 1) / A.cf, yl = (A.Y + A.cf - 1) / A.cf;
  const long long total = (long long)d->n * ((3LL + 1) * A.X * A.Y * A.SZ + (long long)A.lr_c * xl * yl * A.SZ);
  prepare_batch_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)stream>>>(A, d->n, lr, hr, zout);
  WS_POST_LAUNCH(1);
  return 0;
}

}  // extern "C"
