// conv_tc.cu — Conv3d forward / data-gradient as a tcgen05 implicit GEMM for sm_100a.
//
//   D[M = 128 voxel rows, N = Cout (padded to 16)] += A[M, K] * B[N, K]^T,   K = taps x Cin
//
// * A: the activation volume is channels-last (N,X,Y,Z,C) bf16.  For every (tap, 64-channel chunk) one TMA
//   box {64 ch, bz, by, bx, 1} is fetched straight from the 5-D tensor map at the tap-shifted coordinate;
//   the halo / zero padding of the convolution is TMA's out-of-bounds zero fill (coordinates are signed), and
//   a strided convolution is the tensor map's elementStrides.  Each box row is one voxel's 64 channels =
//   128 B, i.e. exactly the canonical K-major SWIZZLE_128B operand tile (8-row atoms, SBO = 1024 B).
// * B: weights pre-packed as bf16 [tap][cout_pad16][cin_pad8]; box {64 ch, N, 1}.
// * accumulators: fp32 in TMEM (N columns), read back with tcgen05.ld 32x32b by 4 epilogue warps which apply
//   the fused epilogue (common.cuh: bias/BN-affine, LeakyReLU, dropout scale, residuals, lrelu-backward
//   mask, BN statistics) and store bf16/fp32 into any channel slice of a channels-last buffer (dense-concat
//   write, torch_blocks.py:214) or an NCXYZ boundary tensor.
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 =
//   epilogue.  smem ring of kStages {A 16 KB, B N*128 B} slots with full/empty mbarriers; tcgen05.commit
//   releases slots and publishes the accumulator.
//
// Replaces nn.Conv3d forward + the dgrad half of convolution_backward at torch_blocks.py:17,278 and
// Generator_3D_Resnet_ESRGAN.py:105 for the stride-1 layers (dgrad = same kernel, taps flipped, channel
// roles swapped, pad' = k-1-p) and the strided discriminator forward (torch_blocks.py:467-506).
#include <cuda.h>
#include <array>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>
#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "tc_task.cuh"

namespace ws {

namespace {

constexpr int kTcThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 128 B
constexpr int kMaxStages = 8;

struct TcParams {
  // destination tile grid (dst = output for fwd, dx for dgrad)
  int N, DX, DY, DZ;
  int bx, by, bz;
  int tiles_x, tiles_y, tiles_z;
  // source coordinate of dst voxel d and tap i:  d*s - p + i
  int kx, ky, kz, sx, sy, sz, px, py, pz;
  int ck;          // reduction channels (cin for fwd, cout for dgrad)
  int kchunks;     // ceil(ck / 64)
  int last_k16;    // MMAs in the last chunk
  int cn;          // valid destination channels
  int n_umma;      // UMMA N of this CTA's tile (multiple of 16, <= 256)
  int n_tile;      // channels per grid.y tile (== n_umma except possibly the last one; kept equal)
  int stages;
  int stage_bytes;
  int a_rows;      // bz*by*bx
  int tap_base;    // first tap of the packed weights this launch uses
  int omx, oax, omy, oay, omz, oaz, ODY, ODZ;  // destination voxel transform (tc_task.cuh)
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kTcThreads, 2)  // two CTAs per SM when the ring is short enough: two MMA issuers
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const TcParams p, const View dst, const Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float stat_s[512];  // per-CTA BatchNorm partial sums (epilogue.cuh)
  if (ep.stat_sum)
    for (int i = threadIdx.x; i < 512; i += blockDim.x) stat_s[i] = 0.f;  // visible after the prologue barrier
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_base = smem_base + p.stages * p.stage_bytes;  // 8-byte aligned (stage_bytes % 128 == 0)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + p.stages * p.stage_bytes + 8 * (2 * kMaxStages + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode
  int t = blockIdx.x;
  const int tz = t % p.tiles_z; t /= p.tiles_z;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int n = t;
  const int x0 = tx * p.bx, y0 = ty * p.by, z0 = tz * p.bz;
  const int n0 = blockIdx.y * p.n_tile;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int taps = p.kx * p.ky * p.kz;
  const int iters = taps * p.kchunks;
  const uint32_t a_bytes = (uint32_t)p.a_rows * 128u;
  const uint32_t b_bytes = (uint32_t)p.n_umma * 128u;

  // Both loops run warp-uniformly and elect one lane only around UTMALDG / UTCHMMA (see conv_tc2.cu).
  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tap = 0; tap < taps; ++tap) {
      const int ti = tap / (p.ky * p.kz), tj = (tap / p.kz) % p.ky, tl = tap % p.kz;
      const int cx = x0 * p.sx - p.px + ti;
      const int cy = y0 * p.sy - p.py + tj;
      const int cz = z0 * p.sz - p.pz + tl;
      for (int ch = 0; ch < p.kchunks; ++ch) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t a_dst = smem_base + stage * p.stage_bytes;
          ptx::mbar_expect_tx(full_bar(stage), a_bytes + b_bytes);
          ptx::tma_load_5d(a_dst, &tmA, full_bar(stage), ch * 64, cz, cy, cx, n);
          ptx::tma_load_3d(a_dst + kABytes, &tmB, full_bar(stage), ch * 64, n0, p.tap_base + tap);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = ptx::make_idesc(/*bf16*/ 1u, 128u, (uint32_t)p.n_umma, 0u, 0u);
    const uint64_t desc_hi = ptx::make_smem_desc_sw128(0, 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      const int ch = it % p.kchunks;
      ptx::mbar_wait(full_bar(stage), phase);
      ptx::tc_fence_after();
      const uint32_t a_addr = smem_base + stage * p.stage_bytes;
      const uint64_t adesc = desc_hi | (uint64_t)((a_addr >> 4) & 0x3fffu);
      const uint64_t bdesc = desc_hi | (uint64_t)(((a_addr + kABytes) >> 4) & 0x3fffu);
      const int nk = (ch == p.kchunks - 1) ? p.last_k16 : 4;
      const uint32_t acc0 = it > 0 ? 1u : 0u;
      if (ptx::elect_one()) {
        ptx::mma_f16_ss(tmem_base, adesc, bdesc, idesc, acc0);
        if (nk > 1) ptx::mma_f16_ss(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
        if (nk > 2) ptx::mma_f16_ss(tmem_base, adesc + 4, bdesc + 4, idesc, 1u);
        if (nk > 3) ptx::mma_f16_ss(tmem_base, adesc + 6, bdesc + 6, idesc, 1u);
        ptx::mma_commit(empty_bar(stage));  // slot free once these MMAs retire
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    if (ptx::elect_one()) ptx::mma_commit(accum_bar);  // accumulator complete
    __syncwarp();
  } else {
    // ===== epilogue: TMEM -> registers -> fused epilogue -> global =====
    const int sub = warp & 3;              // TMEM sub-partition this warp may access
    const int row = sub * 32 + lane;       // accumulator row == voxel index inside the tile
    const int rz = row % p.bz;
    const int ry = (row / p.bz) % p.by;
    const int rx = row / (p.bz * p.by);
    const int gx = x0 + rx, gy = y0 + ry, gz = z0 + rz;
    const bool row_ok = row < p.a_rows && gx < p.DX && gy < p.DY && gz < p.DZ;
    const long long v = ((long long)(gx * p.omx + p.oax) * p.ODY + (gy * p.omy + p.oay)) * p.ODZ + (gz * p.omz + p.oaz);

    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after();

    const EpiVec ev = make_epi_vec(dst, ep);
    const bool simple = epi_is_simple(ep);
    for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
      if (n0 + c0 >= p.cn) break;  // warp-uniform
      uint32_t r[16];
      ptx::tmem_ld16(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (simple) epilogue16_simple(ep, ev, dst, n, v, n0 + c0, p.cn, row_ok, r);
      else epilogue16(ep, ev, dst, n, v, n0 + c0, p.cn, row_ok, r, lane, ep.stat_sum ? stat_s : nullptr, n0);
    }
    if (ep.stat_sum) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps
      flush_bn_stats(ep, stat_s, n0, p.n_umma, p.cn, (int)threadIdx.x - 64, 128);
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKeyEq {
  bool operator()(const MapKey& a, const MapKey& b) const { return memcmp(&a, &b, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return h;
  }
};
int get_tensor_map_impl(const MapKey& key, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash, MapKeyEq> cache;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode();
  WS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap m;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5];
  for (int i = 0; i < 5; ++i) { dims[i] = key.dims[i]; box[i] = key.box[i]; estr[i] = key.estr[i]; }
  for (int i = 0; i < 4; ++i) strides[i] = key.strides[i];
  // key.dtype: low byte = element type, bit 8 = SWIZZLE_64B instead of SWIZZLE_128B (tmap.cuh)
  CUtensorMapDataType dt =
      (key.dtype & 0xffu) == WS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle swz = (key.dtype & kMapSwizzle64)          ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : (key.dtype & kMapSwizzle128Atom32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                                      : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(&m, dt, key.rank, reinterpret_cast<void*>(key.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WS_REQUIRE(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled failed (%d): rank %u dims [%llu %llu %llu %llu %llu] strides [%llu %llu %llu "
             "%llu] box [%u %u %u %u %u]",
             (int)r, key.rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)dims[4],
             (unsigned long long)strides[0], (unsigned long long)strides[1], (unsigned long long)strides[2],
             (unsigned long long)strides[3], box[0], box[1], box[2], box[3], box[4]);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

// choose (bz, by, bx), bz*by*bx <= 128, maximising useful rows per 128-row MMA tile
void choose_tile_search(int DX, int DY, int DZ, int sx, int sy, int sz, int& bx, int& by, int& bz) {
  double best = -1.0;
  bx = by = bz = 1;
  for (int z = 1; z <= DZ && z <= 128; ++z) {
    if ((z - 1) * sz + 1 > 256) break;
    // only full-z or divisors-ish: try all, cost is negligible
    for (int y = 1; y <= DY && z * y <= 128; ++y) {
      if ((y - 1) * sy + 1 > 256) break;
      int xmax = 128 / (z * y);
      if (xmax > DX) xmax = DX;
      for (int x = 1; x <= xmax; ++x) {
        if ((x - 1) * sx + 1 > 256) break;
        long long tiles = (long long)((DX + x - 1) / x) * ((DY + y - 1) / y) * ((DZ + z - 1) / z);
        double eff = (double)DX * DY * DZ / ((double)tiles * 128.0);
        // prefer long z (innermost, contiguous) on ties
        double score = eff + 1e-6 * z + 1e-9 * y;
        if (score > best) { best = score; bx = x; by = y; bz = z; }
      }
    }
  }
}

void choose_tile(int DX, int DY, int DZ, int sx, int sy, int sz, int& bx, int& by, int& bz) {
  struct Hit { int bx, by, bz; };
  static std::mutex mu;
  static std::map<std::array<int, 6>, Hit> memo;  // keyed by the full geometry (no hash collisions)
  const std::array<int, 6> key = {DX, DY, DZ, sx, sy, sz};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(key);
    if (it != memo.end()) { bx = it->second.bx; by = it->second.by; bz = it->second.bz; return; }
  }
  choose_tile_search(DX, DY, DZ, sx, sy, sz, bx, by, bz);
  std::lock_guard<std::mutex> lk(mu);
  memo[key] = Hit{bx, by, bz};
}

}  // namespace

int get_tensor_map(const MapKey& key, CUtensorMap* out) { return get_tensor_map_impl(key, out); }

bool tc_view_ok(const View& v, int channels) {
  if (v.dtype != WS_BF16 || v.cs != 1) return false;
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) != 0) return false;
  if ((v.vs * 2) % 16 != 0 || (v.ns * 2) % 16 != 0) return false;
  return channels >= 16;
}

// mode 0: forward (src = in, dst = out); mode 1: stride-1 dgrad (src = dy, dst = dx).
int tc_conv_launch(const ConvGeom& g, int mode, const View& src, const void* packed_w, const View& dst,
                   const Epi& ep, cudaStream_t st, const TcOverride* ov) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  int SX, SY, SZ;
  int taps_total = g.taps();
  if (ov) {
    p.DX = ov->DX; p.DY = ov->DY; p.DZ = ov->DZ; SX = ov->SX; SY = ov->SY; SZ = ov->SZ;
    p.sx = p.sy = p.sz = 1; p.px = ov->px; p.py = ov->py; p.pz = ov->pz; p.ck = ov->ck; p.cn = ov->cn;
  } else if (mode == 0) {
    p.DX = g.xo; p.DY = g.yo; p.DZ = g.zo; SX = g.x; SY = g.y; SZ = g.z;
    p.sx = g.sx; p.sy = g.sy; p.sz = g.sz; p.px = g.px; p.py = g.py; p.pz = g.pz;
    p.ck = g.cin; p.cn = g.cout;
  } else {
    WS_REQUIRE(g.sx == 1 && g.sy == 1 && g.sz == 1, "tcgen05 dgrad requires stride 1");
    p.DX = g.x; p.DY = g.y; p.DZ = g.z; SX = g.xo; SY = g.yo; SZ = g.zo;
    p.sx = p.sy = p.sz = 1;
    p.px = g.kx - 1 - g.px; p.py = g.ky - 1 - g.py; p.pz = g.kz - 1 - g.pz;
    p.ck = g.cout; p.cn = g.cin;
  }
  p.N = g.n; p.kx = g.kx; p.ky = g.ky; p.kz = g.kz;
  p.omx = p.omy = p.omz = 1; p.ODY = p.DY; p.ODZ = p.DZ;
  if (ov) {
    p.kx = ov->kx; p.ky = ov->ky; p.kz = ov->kz; p.tap_base = ov->tap_base; taps_total = ov->taps_total;
    p.omx = ov->omx; p.oax = ov->oax; p.omy = ov->omy; p.oay = ov->oay; p.omz = ov->omz; p.oaz = ov->oaz;
    p.ODY = ov->ODY; p.ODZ = ov->ODZ;
  }
  choose_tile(p.DX, p.DY, p.DZ, p.sx, p.sy, p.sz, p.bx, p.by, p.bz);
  p.tiles_x = (p.DX + p.bx - 1) / p.bx;
  p.tiles_y = (p.DY + p.by - 1) / p.by;
  p.tiles_z = (p.DZ + p.bz - 1) / p.bz;
  p.a_rows = p.bx * p.by * p.bz;
  p.kchunks = (p.ck + 63) / 64;
  int last = p.ck - 64 * (p.kchunks - 1);
  p.last_k16 = (last + 15) / 16;
  const int cn_pad = (p.cn + 15) / 16 * 16;
  const int ck_pad = (p.ck + 7) / 8 * 8;
  const int n_tiles = (cn_pad + 255) / 256;
  p.n_tile = ((cn_pad + n_tiles - 1) / n_tiles + 15) / 16 * 16;
  p.n_umma = p.n_tile;
  p.stage_bytes = kABytes + p.n_umma * 128;
  p.stages = (int)((220 * 1024 - 1024 - 256) / p.stage_bytes);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  {
    // This kernel's single issuing thread pays a barrier wait + a commit per stage of 4 MMAs (scripts/micro/
    // mma_rate3.cu: ~500 cycles against 288 of tensor-pipe work): with a ring short enough for TWO CTAs per SM the second
    // CTA's issuer fills the gaps of the first (and one CTA's epilogue overlaps the other's main loop).
    static const bool env_single = getenv("WS_TC1_SINGLE") != nullptr;
    const int half = (int)(((227 * 1024) / 2 - 2048 - 1024 - 256) / p.stage_bytes);
    if (!env_single && half >= 3 && p.n_umma <= 256) p.stages = half < p.stages ? half : p.stages;
  }
  int iters = p.kx * p.ky * p.kz * p.kchunks;
  if (p.stages > iters) p.stages = iters < 2 ? 2 : iters;
  uint32_t cols = 32;
  while ((int)cols < p.n_umma) cols <<= 1;
  p.tmem_cols = cols;

  // activation map: dims {C, Z, Y, X, N}
  MapKey ka;
  memset(&ka, 0, sizeof(ka));
  ka.ptr = reinterpret_cast<uintptr_t>(src.ptr);
  ka.rank = 5; ka.dtype = WS_BF16;
  ka.dims[0] = (uint64_t)p.ck; ka.dims[1] = (uint64_t)SZ; ka.dims[2] = (uint64_t)SY; ka.dims[3] = (uint64_t)SX;
  ka.dims[4] = (uint64_t)g.n;
  ka.strides[0] = (uint64_t)src.vs * 2;
  ka.strides[1] = (uint64_t)src.vs * 2 * SZ;
  ka.strides[2] = (uint64_t)src.vs * 2 * SZ * SY;
  ka.strides[3] = (uint64_t)src.ns * 2;
  ka.box[0] = 64;
  ka.box[1] = (uint32_t)((p.bz - 1) * p.sz + 1);
  ka.box[2] = (uint32_t)((p.by - 1) * p.sy + 1);
  ka.box[3] = (uint32_t)((p.bx - 1) * p.sx + 1);
  ka.box[4] = 1;
  ka.estr[0] = 1; ka.estr[1] = (uint32_t)p.sz; ka.estr[2] = (uint32_t)p.sy; ka.estr[3] = (uint32_t)p.sx;
  ka.estr[4] = 1;
  CUtensorMap tmA, tmB;
  if (int e = get_tensor_map(ka, &tmA)) return e;

  // weight map: dims {ck_pad, cn_pad, taps}
  MapKey kb;
  memset(&kb, 0, sizeof(kb));
  kb.ptr = reinterpret_cast<uintptr_t>(packed_w);
  kb.rank = 3; kb.dtype = WS_BF16;
  kb.dims[0] = (uint64_t)ck_pad; kb.dims[1] = (uint64_t)cn_pad; kb.dims[2] = (uint64_t)taps_total;
  kb.strides[0] = (uint64_t)ck_pad * 2;
  kb.strides[1] = (uint64_t)ck_pad * 2 * cn_pad;
  kb.box[0] = 64; kb.box[1] = (uint32_t)p.n_umma; kb.box[2] = 1;
  kb.estr[0] = kb.estr[1] = kb.estr[2] = 1;
  if (int e = get_tensor_map(kb, &tmB)) return e;

  size_t smem = (size_t)p.stages * p.stage_bytes + 8 * (2 * kMaxStages + 2) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv3d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);  // + 2 KB static
  });
  WS_REQUIRE(attr_err == cudaSuccess, "cudaFuncSetAttribute(max dynamic smem) failed: %s",
             cudaGetErrorString(attr_err));
  WS_REQUIRE(smem <= 225 * 1024, "tcgen05 conv: smem request %zu too large", smem);
  dim3 grid((unsigned)(p.N * p.tiles_x * p.tiles_y * p.tiles_z), (unsigned)n_tiles);
  conv3d_tc_kernel<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p, dst, ep);
  WS_POST_LAUNCH(1);
  return 0;
}

}  // namespace ws
