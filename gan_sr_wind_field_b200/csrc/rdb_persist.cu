// rdb_persist.cu — a whole residual dense block (torch_blocks.py:192-290, 328-330) as ONE persistent cooperative kernel.
//
// Round-1 profile: the RRDB trunk (48 blocks x 5 convs on a 16x16x10 volume) ran as ~10 launches per block and
// direction at ~17 us each while its tensor-core work is ~5 us per conv: launch latency, prologue, pipeline ramp and
// drain of 500 tiny kernels were 25-40 % of the training step at 2.5 % tensor-pipe activity.  Here one CTA per
// (sample, x-slab) stays resident for the whole block:
//
//   phase -1  cast the fp32 block input into channels [0, F) of the bf16 concat buffer
//   phase i   dense conv i (3x3x3, cin = F + i*gc -> gc) + LeakyReLU -> channels [cin, cin + gc) of the concat buffer
//   phase n   LFF 1x1x1 (F + n*gc -> F) with bias, alpha, the block skip and the optional RRDB skip -> fp32 output
//
// with a grid-wide barrier between phases (every conv reads its neighbours' previous outputs through the 3^3 halo).
// TMEM, mbarriers, tensor-map prefetch and the role split (TMA producer / MMA issuer / 4 epilogue warps) are set up
// once; the weights of the next phase are prefetched into the ring while the CTA waits at the barrier.
//
// Dense conv formulation ("z-fold"): the kz taps go side by side on the UMMA N dimension (N = 3*gc = 96: an MMA costs
// the same 72 cycles for N = 32 and N = 96), so U[r][dz*gc + co] = sum_{kx,ky,c} A[r + kx*slab_p + ky*DZ][c] *
// W[kx,ky,dz][c][co] needs 9 MMAs per K step instead of 27, and out[z] = U[z-1][dz=0] + U[z][dz=1] + U[z+1][dz=2] is a
// neighbour-row sum in the epilogue (warp shuffles; rows are ordered (y, z) so z +- 1 are adjacent TMEM lanes).
// The activation operand of a 64-channel chunk is ONE TMA box {64 ch, DZ, DY + 2, 3 slabs} — the output slab with its
// x and y halo, zero padding by TMA out-of-bounds fill — and the (kx, ky) taps are UMMA descriptor row offsets into it.
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace ws {

namespace {

constexpr int kMaxPhases = WS_RDB_MAX_CONVS + 1;  // dense convs + LFF
constexpr int kWSlots = 4;
constexpr int kAStages = 2;
// warp 0: TMA producer, warps 1-4: MMA issuers (warp 1 owns the TMEM allocation), warps 5-8: epilogue.
// Several issuing warps because a tcgen05.mma occupies its issuing thread for about as long as it executes
// (scripts/micro/mma_rate3.cu): with one issuer every barrier wait (~280 cycles), commit (~220) and descriptor
// computation between two MMAs was tensor-pipe idle time — the dense-conv phases ran at 2-3x their MMA floor.
constexpr int kThreads = 288;
constexpr int kEpiThread0 = 160;  // first epilogue thread

struct RdbMaps {
  CUtensorMap a[kMaxPhases];
  CUtensorMap b[kMaxPhases];
};

struct RdbFwdParams {
  int N, DX, DY, DZ;
  int F, gc, nconv, ctot;
  int slab;    // output rows of a CTA: DY * DZ
  int slab_p;  // rows of one x-slab of the halo box: (DY + 2) * DZ
  int t_m;     // 128-row accumulator tiles per CTA
  int a_stage_bytes, w_slot_bytes;
  int w_group;  // dense-conv weight tiles (taps) per ring slot = per barrier round trip of an issuer
  int w_slots;  // ring slots (2..kWSlots)
  int n_iss;    // issuing warps: issuer q owns the accumulator tiles m = q, q + n_iss, ...
  int nb_sync;  // 1: neighbour flags instead of grid-wide barriers between the phases
  int early;    // 1: the chunks of a phase that do not depend on the previous phase start before its grid barrier
  int acc_stride;  // TMEM columns between the accumulator sets of even and odd phases (0: one set)
  int a_box_bytes_conv, a_box_bytes_lff;
  int kchunks[kMaxPhases], last_k16[kMaxPhases];
  int n_conv;  // UMMA N of a dense conv: 3 * gc
  int n_lff;   // UMMA N of the LFF: F
  float slope;
  uint32_t tmem_cols;
  int bar_slot;
  long long* dbg;  // optional (WS_RDB_DEBUG_TIMES=1): globaltimer stamps of CTA 0's epilogue at every phase edge
};

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- grid-wide barrier ------------------------------------------------------------------------------------------
// Sense-reversing counter barrier in global memory; reusable across launches without a host reset (the count returns
// to zero, the generation only grows).  One thread per CTA calls it.  The spin is bounded: a protocol bug traps
// instead of hanging the GPU.
__device__ unsigned int g_bar_count[4];
__device__ unsigned int g_bar_gen[4];

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(int slot, unsigned int nblocks) {
  const unsigned int gen = ld_acquire(&g_bar_gen[slot]);
  __threadfence();
  const unsigned int old = atomicAdd(&g_bar_count[slot], 1u);
  if (old == nblocks - 1) {
    g_bar_count[slot] = 0u;
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&g_bar_gen[slot]) : "memory");
  } else {
    const long long t0 = clock64();
    while (ld_acquire(&g_bar_gen[slot]) == gen) {
      if (clock64() - t0 > (1ll << 32)) __trap();  // ~2 s at 2 GHz
    }
  }
  __threadfence();
}
// ---- neighbour synchronisation ----------------------------------------------------------------------------------
// A CTA (sample n, x-slab x0) only ever reads what the CTAs (n, x0 - 1) and (n, x0 + 1) wrote in the previous phase
// (the 3^3 halo), so a grid-wide barrier (~3 us: 128 atomics on one counter plus the slowest CTA of the whole grid)
// is more than needed.  Each CTA publishes "phase k done" in its own flag word and waits for its two neighbours'.
// Flag values are launch_epoch * 16 + k; the epoch lives in device memory and is advanced by the last CTA of a launch
// to finish, so nothing has to be reset between launches and a CUDA-graph replay (frozen kernel arguments) just works.
// Stale flags of earlier launches are always behind the current targets (signed wrap-around comparison).
__device__ unsigned int g_rdb_epoch[2];
__device__ unsigned int g_rdb_done[2];
constexpr int kFlagStride = 32;  // one 128-byte line per flag: a poller must not share a line with another CTA's flag
__device__ unsigned int g_rdb_flag[2][256 * kFlagStride];

__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void flag_wait(const unsigned int* p, unsigned int target) {
  const long long t0 = clock64();
  while ((int)(ld_acquire(p) - target) < 0) {
    __nanosleep(32);
    if (clock64() - t0 > (1ll << 32)) __trap();  // ~2 s at 2 GHz
  }
}
// one thread per CTA: publish phase k and wait until both x-neighbours of the same sample have published it too
__device__ __forceinline__ void neighbour_sync(int kind, unsigned int base, int k, int me, bool has_lo, bool has_hi) {
  __threadfence();
  st_release(&g_rdb_flag[kind][me * kFlagStride], base + (unsigned int)k);
  if (has_lo) flag_wait(&g_rdb_flag[kind][(me - 1) * kFlagStride], base + (unsigned int)k);
  if (has_hi) flag_wait(&g_rdb_flag[kind][(me + 1) * kFlagStride], base + (unsigned int)k);
  __threadfence();
}
// one thread per CTA, after its last phase: the last CTA of the launch advances the epoch
__device__ __forceinline__ void launch_done(int kind, unsigned int nblocks) {
  __threadfence();
  const unsigned int old = atomicAdd(&g_rdb_done[kind], 1u);
  if (old == nblocks - 1) {
    g_rdb_done[kind] = 0u;
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&g_rdb_epoch[kind]) : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// fp32 row epilogue through shared memory: a warp's 32 accumulator rows x `ncols` columns go TMEM -> registers ->
// a padded staging tile, then every row leaves as ONE coalesced 16-byte-per-lane access together with its residual
// rows (the per-thread row walk of epilogue.cuh issues 32 scattered 16-byte accesses per instruction and was
// latency-bound here: 11 us for the LFF of one block).  y = alpha * (acc + bias) + beta1 * r1 + beta2 * r2.
constexpr int kStagePitch = 132;  // floats per staged row (128 + 4: conflict-free 16-byte row writes)
template <typename RowMap>  // staged row (0..31) -> output row (voxel index inside the sample) or -1
__device__ __forceinline__ void rows_epilogue_f32(uint32_t t_row, int ncols, float* stg, int lane, const float* bias,
                                                  float alpha, const float* r1, float beta1, const float* r2,
                                                  float beta2, float* out, long long row_stride, RowMap row_map) {
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t rr[16];
    ptx::tmem_ld16(t_row + (uint32_t)c0, rr);
    ptx::tmem_ld_wait();
    float4* d = reinterpret_cast<float4*>(stg + lane * kStagePitch + c0);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d[q] = make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]), __uint_as_float(rr[4 * q + 2]),
                         __uint_as_float(rr[4 * q + 3]));
  }
  __syncwarp();
  const int c = lane * 4;
  if (c < ncols) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b4 = *reinterpret_cast<const float4*>(bias + c);
    // 8 rows per batch: all residual loads of the batch are issued before the first is consumed (one L2 round trip per
    // batch instead of one per row — this loop was 8-15 us of latency per block)
    for (int rb = 0; rb < 32; rb += 8) {
      long long orow[8];
      float4 x1[8], x2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        orow[u] = row_map(rb + u);
        x1[u] = x2[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (orow[u] >= 0) {
          if (r1) x1[u] = *reinterpret_cast<const float4*>(r1 + orow[u] * row_stride + c);
          if (r2) x2[u] = *reinterpret_cast<const float4*>(r2 + orow[u] * row_stride + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (orow[u] < 0) continue;
        const float4 a = *reinterpret_cast<const float4*>(stg + (rb + u) * kStagePitch + c);
        float4 y = make_float4(alpha * (a.x + b4.x), alpha * (a.y + b4.y), alpha * (a.z + b4.z), alpha * (a.w + b4.w));
        y.x = fmaf(beta1, x1[u].x, y.x); y.y = fmaf(beta1, x1[u].y, y.y); y.z = fmaf(beta1, x1[u].z, y.z); y.w = fmaf(beta1, x1[u].w, y.w);
        y.x = fmaf(beta2, x2[u].x, y.x); y.y = fmaf(beta2, x2[u].y, y.y); y.z = fmaf(beta2, x2[u].z, y.z); y.w = fmaf(beta2, x2[u].w, y.w);
        *reinterpret_cast<float4*>(out + orow[u] * row_stride + c) = y;
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 1)
rdb_fwd_persist_kernel(const __grid_constant__ RdbMaps maps, const RdbFwdParams p, const View x, const View buf,
                       const View out, const Epi ep_lff) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float edge_s[2 * 4 * 2 * 32];  // [tile (t_m <= 2)][warp quarter][u0 of lane 31 | u2 of lane 0][32 channels]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t w_base = smem_base + (uint32_t)kAStages * p.a_stage_bytes;
  const uint32_t bar_off = (uint32_t)kAStages * p.a_stage_bytes + (uint32_t)p.w_slots * p.w_slot_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kAStages + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * kAStages + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * kAStages + kWSlots + s); };
  constexpr int kNumBars = 2 * kAStages + 2 * kWSlots;
  const uint32_t accum_bar = bar_base + 8u * kNumBars;
  const uint32_t phase_bar = bar_base + 8u * (kNumBars + 1);
  const uint32_t tmem_slot = bar_base + 8u * (kNumBars + 2);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * (kNumBars + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.DX, x0 = blockIdx.x % p.DX;
  const int nphases = p.nconv + 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < nphases; ++i) {
      ptx::prefetch_tmap(&maps.a[i]);
      ptx::prefetch_tmap(&maps.b[i]);
    }
    // every issuing warp commits its own MMAs: "empty" / "accumulators complete" take n_iss arrivals
    for (int s = 0; s < kAStages; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), (uint32_t)p.n_iss); }
    for (int s = 0; s < kWSlots; ++s) { ptx::mbar_init(w_full(s), 1); ptx::mbar_init(w_empty(s), (uint32_t)p.n_iss); }
    ptx::mbar_init(accum_bar, (uint32_t)p.n_iss);
    ptx::mbar_init(phase_bar, 1);
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

  if (warp == 0) {
    // ===== TMA producer =====
    int ab = 0, wsl = 0;
    uint32_t aph = 0, wph = 0;
    for (int ph = 0; ph < nphases; ++ph) {
      const bool lff = ph == p.nconv;
      const int ntaps = lff ? 1 : 9;
      const int wg = lff ? 1 : p.w_group;          // weight tiles per ring slot
      const int gpc = (ntaps + wg - 1) / wg;       // tile groups per 64-channel chunk
      const int kch = p.kchunks[ph];
      const int total_w = kch * gpc;
      const uint32_t w_bytes = (uint32_t)(lff ? p.n_lff : p.n_conv) * 128u;
      int wi = 0;
      auto issue_w = [&]() {
        const int ch = wi / gpc, t0 = (wi - ch * gpc) * wg;
        const int cnt = min(wg, ntaps - t0);
        ptx::mbar_wait(w_empty(wsl), wph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(w_full(wsl), (uint32_t)cnt * w_bytes);
          for (int j = 0; j < cnt; ++j)
            ptx::tma_load_3d(w_base + wsl * p.w_slot_bytes + j * w_bytes, &maps.b[ph], w_full(wsl), ch * 64, 0, t0 + j);
        }
        __syncwarp();
        if (++wsl == p.w_slots) { wsl = 0; wph ^= 1u; }
        ++wi;
      };
      // weights do not depend on the previous phase: fill the ring before waiting for the grid barrier
      while (wi < total_w && wi < p.w_slots) issue_w();
      // Only the 64-channel chunks that contain the PREVIOUS phase's output depend on it (the last chunk: conv i reads
      // channels [0, F + i*gc) and conv i-1 wrote the last gc of them).  The chunks before it were final one grid
      // barrier earlier, so their loads — and their MMAs, into the other half of tensor memory — overlap the previous
      // phase's epilogue and the grid barrier.  Before the first dependent chunk: wait until every CTA has published
      // (phase_bar is armed by this CTA's epilogue after the grid barrier), then order the generic-proxy writes before
      // our async-proxy reads.
      const int dep_ch = (ph == 0 || !p.early) ? 0 : (p.F + (ph - 1) * p.gc) / 64;
      for (int ch = 0; ch < kch; ++ch) {
        if (ch == dep_ch) {
          ptx::mbar_wait(phase_bar, (uint32_t)(ph & 1));
          fence_proxy_async_global();
        }
        ptx::mbar_wait(a_empty(ab), aph ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t d = smem_base + ab * p.a_stage_bytes;
          if (lff) {
            ptx::mbar_expect_tx(a_full(ab), (uint32_t)p.a_box_bytes_lff);
            ptx::tma_load_5d(d, &maps.a[ph], a_full(ab), ch * 64, 0, 0, x0, n);
          } else {
            ptx::mbar_expect_tx(a_full(ab), (uint32_t)p.a_box_bytes_conv);
            ptx::tma_load_5d(d, &maps.a[ph], a_full(ab), ch * 64, 0, -1, x0 - 1, n);
          }
        }
        __syncwarp();
        if (++ab == kAStages) { ab = 0; aph ^= 1u; }
        const int upto = (ch + 1) * gpc + p.w_slots;
        while (wi < total_w && wi < upto) issue_w();
      }
    }
  } else if (warp <= 4) {
    // ===== MMA issuers: warps 1 .. n_iss =====
    const int q = warp - 1;
    if (q < p.n_iss) {
    const uint64_t desc_hi = ptx::make_smem_desc_sw128(0, 16, 1024);
    int ab = 0, wsl = 0;
    uint32_t aph = 0, wph = 0;
    for (int ph = 0; ph < nphases; ++ph) {
      const bool lff = ph == p.nconv;
      const int ntaps = lff ? 1 : 9;
      const int wg = lff ? 1 : p.w_group;
      const uint32_t w_bytes = (uint32_t)(lff ? p.n_lff : p.n_conv) * 128u;
      const int n_umma = lff ? p.n_lff : p.n_conv;
      const uint32_t idesc = ptx::make_idesc(1u, 128u, (uint32_t)n_umma, 0u, 0u);
      const int kch = p.kchunks[ph];
      for (int ch = 0; ch < kch; ++ch) {
        const int nk = (ch == kch - 1) ? p.last_k16[ph] : 4;
        ptx::mbar_wait(a_full(ab), aph);
        const uint32_t a_addr = smem_base + ab * p.a_stage_bytes;
        int gpos = 0;  // position of the tap inside its weight group
        for (int tap = 0; tap < ntaps; ++tap) {
          if (gpos == 0) {
            ptx::mbar_wait(w_full(wsl), wph);
            ptx::tc_fence_after();
          }
          const uint64_t bdesc = desc_hi | (uint64_t)(((w_base + wsl * p.w_slot_bytes + gpos * w_bytes) >> 4) & 0x3fffu);
          const uint32_t acc0 = (ch > 0 || tap > 0) ? 1u : 0u;
          // tap (kx, ky) -> first operand row: the output rows start one y line into the centre slab of the box
          const int roff = lff ? 0 : (tap / 3) * p.slab_p + (tap % 3) * p.DZ;
          for (int m = q; m < p.t_m; m += p.n_iss) {
            const uint32_t am = a_addr + (uint32_t)(m * 128 + roff) * 128u;
            const uint64_t adesc = desc_hi | (uint64_t)((am >> 4) & 0x3fffu);
            const uint32_t d_tmem = tmem_base + (uint32_t)((ph & 1) * p.acc_stride + m * 128);
            if (ptx::elect_one()) {
              ptx::mma_f16_ss(d_tmem, adesc, bdesc, idesc, acc0);
              if (nk > 1) ptx::mma_f16_ss(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              if (nk > 2) ptx::mma_f16_ss(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              if (nk > 3) ptx::mma_f16_ss(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
            }
            __syncwarp();
          }
          if (++gpos == wg || tap == ntaps - 1) {
            gpos = 0;
            if (ptx::elect_one()) ptx::mma_commit(w_empty(wsl));
            __syncwarp();
            if (++wsl == p.w_slots) { wsl = 0; wph ^= 1u; }
          }
        }
        if (ptx::elect_one()) ptx::mma_commit(a_empty(ab));
        __syncwarp();
        if (++ab == kAStages) { ab = 0; aph ^= 1u; }
      }
      if (ptx::elect_one()) ptx::mma_commit(accum_bar);
      __syncwarp();
    }
    }
  } else {
    // ===== epilogue warps =====
    const int sub = warp & 3;            // TMEM lane quarter this warp may read
    const int et = threadIdx.x - kEpiThread0;  // 0..127
    int stamp_i = 0;
    auto stamp = [&]() {
      if (p.dbg && (int)blockIdx.x == p.DX / 2 && et == 0) p.dbg[stamp_i] = globaltimer_ns();
      ++stamp_i;
    };
    // (neighbour flags: this launch's epoch, read before anything is published)
    const unsigned int sync_base = (p.nb_sync && et == 0) ? (ld_acquire(&g_rdb_epoch[0]) << 4) : 0u;
    int sync_k = 0;
    auto publish = [&]() {
      // make this CTA's global writes visible grid-wide, wait for everybody else's, release the producer
      __threadfence();
      fence_proxy_async_global();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      stamp();  // epilogue stores done
      if (et == 0) {
        if (p.nb_sync) neighbour_sync(0, sync_base, ++sync_k, (int)blockIdx.x, x0 > 0, x0 < p.DX - 1);
        else grid_barrier(p.bar_slot, gridDim.x);
        ptx::mbar_arrive(phase_bar);
      }
      stamp();  // barrier passed
    };
    stamp();  // kernel start (after the prologue)
    // ---- phase -1: x (fp32) -> buf[:, 0:F) (bf16), this CTA's slab (4 independent 32-byte reads in flight per thread)
    {
      const int c8 = p.F / 8;
      const long long v0 = (long long)x0 * p.slab;
      const int total = p.slab * c8;
      for (int i0 = et; i0 < total; i0 += 8 * 128) {
        float4 a[8], b[8];  // 8 independent 32-byte reads in flight per thread
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 128;
          if (i < total) {
            const int r = i / c8, q = i - r * c8;
            const float* src = (const float*)x.ptr + x.off(n, q * 8, v0 + r);
            a[u] = reinterpret_cast<const float4*>(src)[0];
            b[u] = reinterpret_cast<const float4*>(src)[1];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 128;
          if (i < total) {
            const int r = i / c8, q = i - r * c8;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(a[u].x, a[u].y), h1 = __floats2bfloat162_rn(a[u].z, a[u].w);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(b[u].x, b[u].y), h3 = __floats2bfloat162_rn(b[u].z, b[u].w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>((__nv_bfloat16*)buf.ptr + buf.off(n, q * 8, v0 + r)) = o;
          }
        }
      }
      publish();
    }
    for (int ph = 0; ph < nphases; ++ph) {
      const bool lff = ph == p.nconv;
      ptx::mbar_wait(accum_bar, (uint32_t)(ph & 1));
      ptx::tc_fence_after();
      stamp();  // MMAs of this phase complete
      const uint32_t acc_off = (uint32_t)((ph & 1) * p.acc_stride);  // this phase's half of tensor memory
      if (!lff) {
        const int c_out0 = p.F + ph * p.gc;  // first channel of this conv's slice of the concat buffer
        const int gc = p.gc;
        // pass 1: the rows a neighbouring warp (or the other tile) needs: u0 of lane 31, u2 of lane 0
        for (int m = 0; m < p.t_m; ++m) {
          const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + acc_off + (uint32_t)(m * 128);
          float* e0 = edge_s + ((m * 4 + sub) * 2 + 0) * 32;
          float* e2 = edge_s + ((m * 4 + sub) * 2 + 1) * 32;
          for (int c0 = 0; c0 < gc; c0 += 16) {
            uint32_t r0[16], r2[16];
            ptx::tmem_ld16(t_row + (uint32_t)c0, r0);
            ptx::tmem_ld16(t_row + (uint32_t)(2 * gc + c0), r2);
            ptx::tmem_ld_wait();
            if (lane == 31) {
#pragma unroll
              for (int j = 0; j < 16; ++j) e0[c0 + j] = __uint_as_float(r0[j]);
            }
            if (lane == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) e2[c0 + j] = __uint_as_float(r2[j]);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // pass 2: out[z] = lrelu(U0[z-1] + U1[z] + U2[z+1])
        for (int m = 0; m < p.t_m; ++m) {
          const int r = m * 128 + sub * 32 + lane;
          const int z = r % p.DZ;
          const bool row_ok = r < p.slab;
          const bool has_lo = z > 0, has_hi = z < p.DZ - 1;
          const int g = m * 4 + sub;
          const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + acc_off + (uint32_t)(m * 128);
          const long long v = (long long)x0 * p.slab + r;
          __nv_bfloat16* dst = (__nv_bfloat16*)buf.ptr + buf.off(n, c_out0, v);
          for (int c0 = 0; c0 < gc; c0 += 16) {
            uint32_t r0[16], r1[16], r2[16];
            ptx::tmem_ld16(t_row + (uint32_t)c0, r0);
            ptx::tmem_ld16(t_row + (uint32_t)(gc + c0), r1);
            ptx::tmem_ld16(t_row + (uint32_t)(2 * gc + c0), r2);
            ptx::tmem_ld_wait();
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float lo = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);
              float hi = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);
              if (lane == 0 && g > 0) lo = edge_s[((g - 1) * 2 + 0) * 32 + c0 + j];
              if (lane == 31 && g < 4 * p.t_m - 1) hi = edge_s[((g + 1) * 2 + 1) * 32 + c0 + j];
              float t = __uint_as_float(r1[j]) + (has_lo ? lo : 0.f) + (has_hi ? hi : 0.f);
              y[j] = t > 0.f ? t : p.slope * t;
            }
            if (row_ok) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              uint4* q = reinterpret_cast<uint4*>(dst + c0);
              q[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              q[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        ptx::tc_fence_before();
        publish();
      } else {
        // LFF: out = alpha * (acc + bias) + beta1 * x + beta2 * outer (fp32): rows are contiguous voxels of this
        // sample's slab, staged through the (now idle) activation buffers
        float* stg = reinterpret_cast<float*>(smem) + sub * 32 * kStagePitch;
        const long long row_base = (long long)x0 * p.slab;
        const float* r1 = ep_lff.res1.ptr ? (const float*)ep_lff.res1.ptr + ep_lff.res1.off(n, 0, 0) : nullptr;
        const float* r2 = ep_lff.res2.ptr ? (const float*)ep_lff.res2.ptr + ep_lff.res2.off(n, 0, 0) : nullptr;
        float* o = (float*)out.ptr + out.off(n, 0, 0);
        for (int m = 0; m < p.t_m; ++m) {
          const int r0 = m * 128 + sub * 32;
          const int ok = p.slab - r0;  // valid rows of this warp's 32
          if (ok <= 0) break;
          const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + acc_off + (uint32_t)(m * 128);
          rows_epilogue_f32(t_row, p.n_lff, stg, lane, ep_lff.bias, ep_lff.alpha, r1, ep_lff.beta1, r2, ep_lff.beta2, o,
                            out.vs, [&](int r) -> long long { return r < ok ? row_base + r0 + r : -1; });
        }
        ptx::tc_fence_before();
        stamp();  // LFF epilogue done
      }
    }
    if (p.nb_sync && et == 0) launch_done(0, gridDim.x);
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---- backward: the data-gradient chain of the block ----------------------------------------------------------------
// dL/d(concat buffer) never leaves the SM: a CTA owns the same voxel rows in every phase and keeps their fp32
// gradient for ALL ctot channels in tensor memory (2 tiles x 256 columns = the whole 512-column TMEM), so the
// accumulate-into-gradient read-modify-write of the per-conv kernels (37 MB through L2 per conv) disappears:
//
//   phase -1  g_lff = alpha * dy (bf16) for this CTA's rows                                   (CTA-local)
//   phase L   acc[:, 0:ctot)  = g_lff x W_lff^T                                               (1x1x1: no halo)
//             g_{n-1} = acc[:, ctot-gc:ctot) * lrelu'(buf[..])  -> gbuf slice n-1             | grid barrier
//   phase i   acc[:, 0:cin_i) += dgrad_i(g_i)   (27 taps as row offsets into ONE halo box of g_i, K = gc = 32)
//             g_{i-1} = acc[:, cin_i-gc:cin_i) * lrelu'(buf[..]) -> gbuf slice i-1            | grid barrier   (i = n-1 .. 1)
//   phase 0   acc[:, 0:F) += dgrad_0(g_0);  dx = acc[:, 0:F) + beta1 * dy                     (fp32 out)
//
// Rows are ordered (y, z) with a z pitch of DZ + 2, so the kz taps are row offsets too (the halo box carries the z pad
// rows; TMA out-of-bounds fill supplies every zero).  gbuf (all g_i side by side) and g_lff feed the weight-gradient
// GEMMs afterwards (wgrad_tc.cu), exactly as in the per-conv path.
struct RdbBwdParams {
  int N, DX, DY, DZ;
  int F, gc, nconv, ctot;
  int pz;      // z pitch of the row order: DZ + 2
  int slab;    // voxels of an x-slab: DY * DZ
  int slab_p;  // rows of one x-slab of the halo box: (DY + 2) * pz
  int t_m;
  int ctot_pad;  // accumulator columns per tile
  int a_stage_bytes, w_slot_bytes;
  int w_group;    // dense-conv weight tiles (taps) per ring slot
  int w_slots;    // ring slots (2..kWSlots)
  long long* dbg;  // optional (WS_RDB_DEBUG_TIMES=1): globaltimer stamps of an interior CTA at every phase edge
  int nb_sync;    // 1: neighbour flags instead of grid-wide barriers between the phases
  int n_iss_m;    // issuers across accumulator tiles
  int tap_split;  // x issuers across the taps of a dense conv (they only ever ADD to the accumulators: any order works)
  int a_box_bytes_conv, a_box_bytes_lff;
  int kch_lff, last_k16_lff;
  int cin[kMaxPhases];
  float slope, alpha, beta1;
  uint32_t tmem_cols;
  int bar_slot;
};

__global__ void __launch_bounds__(kThreads, 1)
rdb_bwd_persist_kernel(const __grid_constant__ RdbMaps maps, const RdbBwdParams p, const View dy, const View buf,
                       const View g_lff, const View gbuf, const View dx) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t w_base = smem_base + (uint32_t)kAStages * p.a_stage_bytes;
  const uint32_t bar_off = (uint32_t)kAStages * p.a_stage_bytes + (uint32_t)p.w_slots * p.w_slot_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kAStages + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * kAStages + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * kAStages + kWSlots + s); };
  constexpr int kNumBars = 2 * kAStages + 2 * kWSlots;
  const uint32_t accum_bar = bar_base + 8u * kNumBars;
  const uint32_t phase_bar = bar_base + 8u * (kNumBars + 1);
  const uint32_t tmem_slot = bar_base + 8u * (kNumBars + 2);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * (kNumBars + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.DX, x0 = blockIdx.x % p.DX;
  const int nphases = p.nconv + 1;  // phase 0 = LFF, phase j >= 1 = dense conv (nconv - j)

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < nphases; ++i) {
      ptx::prefetch_tmap(&maps.a[i]);
      ptx::prefetch_tmap(&maps.b[i]);
    }
    const uint32_t n_iss = (uint32_t)(p.n_iss_m * p.tap_split);
    for (int s = 0; s < kAStages; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), n_iss); }
    for (int s = 0; s < kWSlots; ++s) { ptx::mbar_init(w_full(s), 1); ptx::mbar_init(w_empty(s), n_iss); }
    ptx::mbar_init(accum_bar, n_iss);
    ptx::mbar_init(phase_bar, 1);
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

  // maps.a[0] / maps.b[0]: LFF (g_lff, packed LFF dgrad weights); maps.a[1 + i] / maps.b[1 + i]: dense conv i
  if (warp == 0) {
    // ===== TMA producer =====
    int ab = 0, wsl = 0;
    uint32_t aph = 0, wph = 0;
    for (int ph = 0; ph < nphases; ++ph) {
      const bool lff = ph == 0;
      const int ci = lff ? 0 : p.nconv - ph;  // dense conv index
      const int mi = lff ? 0 : 1 + ci;
      const int ntaps = lff ? 1 : 27;
      const int wg = lff ? 1 : p.w_group;      // weight tiles per ring slot
      const int gpc = (ntaps + wg - 1) / wg;   // tile groups per K chunk
      const int kch = lff ? p.kch_lff : 1;
      const int total_w = kch * gpc;
      const uint32_t w_bytes = lff ? (uint32_t)p.ctot_pad * 128u : (uint32_t)p.cin[ci] * 64u;
      int wi = 0;
      auto issue_w = [&]() {
        const int ch = wi / gpc, t0 = (wi - ch * gpc) * wg;
        const int cnt = min(wg, ntaps - t0);
        ptx::mbar_wait(w_empty(wsl), wph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(w_full(wsl), (uint32_t)cnt * w_bytes);
          for (int j = 0; j < cnt; ++j)
            ptx::tma_load_3d(w_base + wsl * p.w_slot_bytes + j * w_bytes, &maps.b[mi], w_full(wsl), ch * 64, 0, t0 + j);
        }
        __syncwarp();
        if (++wsl == p.w_slots) { wsl = 0; wph ^= 1u; }
        ++wi;
      };
      while (wi < total_w && wi < p.w_slots) issue_w();
      ptx::mbar_wait(phase_bar, (uint32_t)(ph & 1));
      fence_proxy_async_global();
      for (int ch = 0; ch < kch; ++ch) {
        ptx::mbar_wait(a_empty(ab), aph ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t d = smem_base + ab * p.a_stage_bytes;
          if (lff) {
            ptx::mbar_expect_tx(a_full(ab), (uint32_t)p.a_box_bytes_lff);
            ptx::tma_load_5d(d, &maps.a[0], a_full(ab), ch * 64, 0, 0, x0, n);
          } else {
            ptx::mbar_expect_tx(a_full(ab), (uint32_t)p.a_box_bytes_conv);
            ptx::tma_load_5d(d, &maps.a[mi], a_full(ab), 0, -1, -1, x0 - 1, n);
          }
        }
        __syncwarp();
        if (++ab == kAStages) { ab = 0; aph ^= 1u; }
        const int upto = (ch + 1) * gpc + p.w_slots;
        while (wi < total_w && wi < upto) issue_w();
      }
    }
  } else if (warp <= 4) {
    // ===== MMA issuers: warps 1 .. n_iss_m * tap_split.  Issuer q owns the accumulator tiles m = qm, qm + n_iss_m, ...
    // and, in the dense-conv phases (which only ever add to the accumulators), the taps with tap % tap_split == qt.
    // Every issuer follows every barrier phase (waits and commits), whether or not it has MMAs in it. =====
    const int q = warp - 1;
    if (q < p.n_iss_m * p.tap_split) {
    const int qm = q % p.n_iss_m, qt = q / p.n_iss_m;
    const uint64_t desc128 = ptx::make_smem_desc_sw128(0, 16, 1024);
    // SWIZZLE_64B operand rows (K = gc = 32 channels = 64 B): layout type 4, SBO = 8 rows x 64 B (scripts/micro/sw64.cu)
    const uint64_t desc64 = ((uint64_t)1 << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
    int ab = 0, wsl = 0;
    uint32_t aph = 0, wph = 0;
    for (int ph = 0; ph < nphases; ++ph) {
      const bool lff = ph == 0;
      const int ci = lff ? 0 : p.nconv - ph;
      const int ntaps = lff ? 1 : 27;
      const int wg = lff ? 1 : p.w_group;
      const uint32_t w_bytes = lff ? (uint32_t)p.ctot_pad * 128u : (uint32_t)p.cin[ci] * 64u;
      const int kch = lff ? p.kch_lff : 1;
      const int n_umma = lff ? p.ctot_pad : p.cin[ci];
      const uint32_t idesc = ptx::make_idesc(1u, 128u, (uint32_t)n_umma, 0u, 0u);
      const uint64_t desc_hi = lff ? desc128 : desc64;
      const uint32_t row_bytes = lff ? 128u : 64u;
      for (int ch = 0; ch < kch; ++ch) {
        const int nk = lff ? ((ch == kch - 1) ? p.last_k16_lff : 4) : 2;
        ptx::mbar_wait(a_full(ab), aph);
        const uint32_t a_addr = smem_base + ab * p.a_stage_bytes;
        int gpos = 0;  // position of the tap inside its weight group
        for (int tap = 0; tap < ntaps; ++tap) {
          if (gpos == 0) {
            ptx::mbar_wait(w_full(wsl), wph);
            ptx::tc_fence_after();
          }
          const bool mine = lff ? qt == 0 : (tap & (p.tap_split - 1)) == qt;
          const uint64_t bdesc = desc_hi | (uint64_t)(((w_base + wsl * p.w_slot_bytes + gpos * w_bytes) >> 4) & 0x3fffu);
          // the LFF overwrites the accumulators, every dense conv adds to them
          const uint32_t acc0 = (!lff || ch > 0) ? 1u : 0u;
          const int ti = tap / 9, tj = (tap / 3) % 3, tl = tap % 3;
          const int roff = lff ? 0 : ti * p.slab_p + tj * p.pz + tl;
          for (int m = qm; mine && m < p.t_m; m += p.n_iss_m) {
            const uint32_t am = a_addr + (uint32_t)(m * 128 + roff) * row_bytes;
            const uint64_t adesc = desc_hi | (uint64_t)((am >> 4) & 0x3fffu);
            const uint32_t d_tmem = tmem_base + (uint32_t)(m * p.ctot_pad);
            if (ptx::elect_one()) {
              ptx::mma_f16_ss(d_tmem, adesc, bdesc, idesc, acc0);
              if (nk > 1) ptx::mma_f16_ss(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              if (nk > 2) ptx::mma_f16_ss(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              if (nk > 3) ptx::mma_f16_ss(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
            }
            __syncwarp();
          }
          if (++gpos == wg || tap == ntaps - 1) {
            gpos = 0;
            if (ptx::elect_one()) ptx::mma_commit(w_empty(wsl));
            __syncwarp();
            if (++wsl == p.w_slots) { wsl = 0; wph ^= 1u; }
          }
        }
        if (ptx::elect_one()) ptx::mma_commit(a_empty(ab));
        __syncwarp();
        if (++ab == kAStages) { ab = 0; aph ^= 1u; }
      }
      if (ptx::elect_one()) ptx::mma_commit(accum_bar);
      __syncwarp();
    }
    }
  } else {
    // ===== epilogue warps =====
    const int sub = warp & 3;
    const int et = threadIdx.x - kEpiThread0;
    const unsigned int sync_base = (p.nb_sync && et == 0) ? (ld_acquire(&g_rdb_epoch[1]) << 4) : 0u;
    int sync_k = 0;
    int stamp_i = 0;
    auto stamp = [&]() {
      if (p.dbg && (int)blockIdx.x == p.DX / 2 && et == 0) p.dbg[stamp_i] = globaltimer_ns();
      ++stamp_i;
    };
    stamp();  // kernel start (after the prologue)
    // ---- phase -1: g_lff = alpha * dy (bf16) for this CTA's rows; only this CTA reads them back (1x1x1 conv)
    {
      const int c8 = p.F / 8;
      const long long v0 = (long long)x0 * p.slab;
      const int total = p.slab * c8;
      for (int i0 = et; i0 < total; i0 += 8 * 128) {
        float4 a[8], b[8];  // 8 independent 32-byte reads in flight per thread
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 128;
          if (i < total) {
            const int r = i / c8, q = i - r * c8;
            const float* src = (const float*)dy.ptr + dy.off(n, q * 8, v0 + r);
            a[u] = reinterpret_cast<const float4*>(src)[0];
            b[u] = reinterpret_cast<const float4*>(src)[1];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 128;
          if (i < total) {
            const int r = i / c8, q = i - r * c8;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(p.alpha * a[u].x, p.alpha * a[u].y);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(p.alpha * a[u].z, p.alpha * a[u].w);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(p.alpha * b[u].x, p.alpha * b[u].y);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(p.alpha * b[u].z, p.alpha * b[u].w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>((__nv_bfloat16*)g_lff.ptr + g_lff.off(n, q * 8, v0 + r)) = o;
          }
        }
      }
      __threadfence();
      fence_proxy_async_global();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) ptx::mbar_arrive(phase_bar);
      stamp();  // g_lff written
    }
    for (int ph = 0; ph < nphases; ++ph) {
      ptx::mbar_wait(accum_bar, (uint32_t)(ph & 1));
      ptx::tc_fence_after();
      stamp();  // MMAs of this phase complete
      const bool last = ph == nphases - 1;
      // channels [c_hi - gc, c_hi) of the accumulated gradient are final now: the output of dense conv j
      const int j = p.nconv - 1 - ph;
      const int c_hi = p.F + (j + 1) * p.gc;
      for (int m = 0; m < p.t_m; ++m) {
        const int r = m * 128 + sub * 32 + lane;
        const int y = r / p.pz, z = r - y * p.pz;
        const bool row_ok = y < p.DY && z < p.DZ;
        const long long v = (long long)x0 * p.slab + (long long)y * p.DZ + z;
        const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(m * p.ctot_pad);
        if (!last) {
          const __nv_bfloat16* mk = (const __nv_bfloat16*)buf.ptr + buf.off(n, c_hi - p.gc, row_ok ? v : 0);
          __nv_bfloat16* dst = (__nv_bfloat16*)gbuf.ptr + gbuf.off(n, j * p.gc, row_ok ? v : 0);
          for (int c0 = 0; c0 < p.gc; c0 += 16) {
            uint32_t rr[16];
            ptx::tmem_ld16(t_row + (uint32_t)(c_hi - p.gc + c0), rr);
            uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
            if (row_ok) {
              m0 = reinterpret_cast<const uint4*>(mk + c0)[0];
              m1 = reinterpret_cast<const uint4*>(mk + c0)[1];
            }
            ptx::tmem_ld_wait();
            if (row_ok) {
              const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
              uint32_t pk[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mw[q]));
                float g0 = __uint_as_float(rr[2 * q]), g1 = __uint_as_float(rr[2 * q + 1]);
                g0 = a.x > 0.f ? g0 : p.slope * g0;
                g1 = a.y > 0.f ? g1 : p.slope * g1;
                __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
                pk[q] = *reinterpret_cast<uint32_t*>(&h);
              }
              reinterpret_cast<uint4*>(dst + c0)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              reinterpret_cast<uint4*>(dst + c0)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        } else {
          // dx = dL/d(concat)[:, 0:F) + beta1 * dy (the block's skip connection): staged through the idle operand
          // buffers so that every voxel row is one coalesced access (rows_epilogue_f32)
          float* stg = reinterpret_cast<float*>(smem) + sub * 32 * kStagePitch;
          const int r0 = m * 128 + sub * 32;
          const long long row_base = (long long)x0 * p.slab;
          rows_epilogue_f32(t_row, p.F, stg, lane, nullptr, 1.f, (const float*)dy.ptr + dy.off(n, 0, 0), p.beta1, nullptr,
                            0.f, (float*)dx.ptr + dx.off(n, 0, 0), dx.vs, [&](int r) -> long long {
                              const int rr = r0 + r, yy = rr / p.pz, zz = rr - yy * p.pz;
                              return (yy < p.DY && zz < p.DZ) ? row_base + (long long)yy * p.DZ + zz : -1;
                            });
        }
      }
      ptx::tc_fence_before();
      stamp();  // epilogue stores issued
      if (!last) {
        __threadfence();
        fence_proxy_async_global();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        stamp();  // ... and visible
        if (et == 0) {
          if (p.nb_sync) neighbour_sync(1, sync_base, ++sync_k, (int)blockIdx.x, x0 > 0, x0 < p.DX - 1);
          else grid_barrier(p.bar_slot, gridDim.x);
          ptx::mbar_arrive(phase_bar);
        }
        stamp();  // neighbours have published too
      }
    }
    if (p.nb_sync && et == 0) launch_done(1, gridDim.x);
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int make_act_map(const View& src, int channels, int DX, int DY, int DZ, int N, int box_y, int box_x, CUtensorMap* out,
                 int box_z = 0, int box_c = 64) {
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = reinterpret_cast<uintptr_t>(src.ptr);
  k.rank = 5; k.dtype = WS_BF16 | (box_c == 32 ? kMapSwizzle64 : 0u);
  k.dims[0] = (uint64_t)channels; k.dims[1] = (uint64_t)DZ; k.dims[2] = (uint64_t)DY; k.dims[3] = (uint64_t)DX;
  k.dims[4] = (uint64_t)N;
  k.strides[0] = (uint64_t)src.vs * 2;
  k.strides[1] = (uint64_t)src.vs * 2 * DZ;
  k.strides[2] = (uint64_t)src.vs * 2 * DZ * DY;
  k.strides[3] = (uint64_t)src.ns * 2;
  k.box[0] = (uint32_t)box_c; k.box[1] = (uint32_t)(box_z ? box_z : DZ); k.box[2] = (uint32_t)box_y;
  k.box[3] = (uint32_t)box_x; k.box[4] = 1;
  for (int i = 0; i < 5; ++i) k.estr[i] = 1;
  return get_tensor_map(k, out);
}
int make_w_map(const void* packed, int k_pad, int rows, int taps, CUtensorMap* out, int box_c = 64) {
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = reinterpret_cast<uintptr_t>(packed);
  k.rank = 3; k.dtype = WS_BF16 | (box_c == 32 ? kMapSwizzle64 : 0u);
  k.dims[0] = (uint64_t)k_pad; k.dims[1] = (uint64_t)rows; k.dims[2] = (uint64_t)taps;
  k.strides[0] = (uint64_t)k_pad * 2;
  k.strides[1] = (uint64_t)k_pad * 2 * rows;
  k.box[0] = (uint32_t)box_c; k.box[1] = (uint32_t)rows; k.box[2] = 1;
  k.estr[0] = k.estr[1] = k.estr[2] = 1;
  return get_tensor_map(k, out);
}

bool env_off(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}

}  // namespace

// Geometry the persistent kernels cover: bf16 tensor-core mode, 3^3 dense convs with gc % 16 == 0 and 3 * gc <= 128,
// 1^3 LFF with F % 16 == 0 and F <= 128, one x-slab of DY * DZ <= 256 rows per CTA and at most one CTA per SM.
bool rdb_persist_ok(const ws_rdb_desc* d, const View& x, const View& buf, int sm_count) {
  static const bool off = env_off("WS_DISABLE_RDB_PERSIST");
  if (off || d->math != WS_MATH_BF16) return false;
  if (d->nconv < 1 || d->nconv > WS_RDB_MAX_CONVS || d->k != 3 || d->k_lff != 1) return false;
  if (d->gc % 16 != 0 || 3 * d->gc > 128 || d->f % 16 != 0 || d->f > 128 || d->f % 8 != 0) return false;
  const int slab = d->y * d->z;
  if (slab > 256 || d->z > 128 || d->y + 2 > 256) return false;
  if ((long long)d->n * d->x > sm_count) return false;
  if (buf.dtype != WS_BF16 || buf.cs != 1 || x.dtype != WS_F32 || x.cs != 1) return false;
  if ((reinterpret_cast<uintptr_t>(buf.ptr) & 15) || (buf.vs * 2) % 16 || (buf.ns * 2) % 16) return false;
  if ((reinterpret_cast<uintptr_t>(x.ptr) & 15) || (x.vs * 4) % 16 || (x.ns * 4) % 16) return false;
  return true;
}

// packed[i], i < nconv: z-folded forward packing [tap (kx,ky)][dz * gc + co][cin_pad8] (pack_tc_multi, fold axis z);
// packed[nconv]: the LFF's ordinary forward packing.
int rdb_persist_forward(const ws_rdb_desc* d, const View& x, const View& buf, const View& out, void* const* packed,
                        const Epi& ep_lff, cudaStream_t st) {
  RdbFwdParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n; p.DX = d->x; p.DY = d->y; p.DZ = d->z;
  p.F = d->f; p.gc = d->gc; p.nconv = d->nconv; p.ctot = d->f + d->nconv * d->gc;
  p.slab = d->y * d->z;
  p.slab_p = (d->y + 2) * d->z;
  p.t_m = (p.slab + 127) / 128;
  p.n_conv = 3 * d->gc;
  p.n_lff = d->f;
  p.slope = d->slope;
  p.bar_slot = 0;
  static long long* dbg_buf = nullptr;
  static int dbg_calls = 0;
  static const bool dbg_on = env_off("WS_RDB_DEBUG_TIMES");  // (env_off: "set and not 0")
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    static const int dbg_at = atoi(getenv("WS_RDB_DEBUG_TIMES")) > 1 ? atoi(getenv("WS_RDB_DEBUG_TIMES")) : 40;
    if (++dbg_calls == dbg_at) {  // a warm call
      cudaStreamSynchronize(st);
      long long h[64];
      cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
      const int nst = 1 + 2 + 3 * d->nconv + 2;
      fprintf(stderr, "[rdb_fwd_persist] phase-edge stamps of an interior CTA (us since kernel start):");
      for (int i = 1; i < nst && i < 64; ++i) fprintf(stderr, " %.1f", (h[i] - h[0]) * 1e-3);
      fprintf(stderr, "\n  order: cast-done, barrier | per conv: mma-done, epilogue-done, barrier | lff mma-done, epilogue-done\n");
    }
    p.dbg = dbg_buf;
  }
  p.a_box_bytes_conv = 3 * p.slab_p * 128;
  p.a_box_bytes_lff = p.slab * 128;
  p.a_stage_bytes = (p.a_box_bytes_conv + 1023) / 1024 * 1024;
  // dense-conv weight tiles share a ring slot (one barrier round trip of the issuers per group) in threes when the
  // shared memory allows: 2 activation stages + 2 slots must fit
  static const int env_wg = getenv("WS_RDB_WGROUP") ? atoi(getenv("WS_RDB_WGROUP")) : 3;
  static const int env_niss = getenv("WS_RDB_NISS") ? atoi(getenv("WS_RDB_NISS")) : 4;
  p.n_iss = p.t_m < 4 ? p.t_m : 4;
  if (p.n_iss > env_niss) p.n_iss = env_niss < 1 ? 1 : env_niss;
  const int conv_tile = p.n_conv * 128, lff_tile = p.n_lff * 128;
  p.w_group = env_wg < 1 ? 1 : (env_wg > 9 ? 9 : env_wg);
  auto slot_bytes = [&](int wg) { int b = wg * conv_tile; if (lff_tile > b) b = lff_tile; return (b + 1023) / 1024 * 1024; };
  while (p.w_group > 1 && kAStages * p.a_stage_bytes + 2 * slot_bytes(p.w_group) + 2048 > 222 * 1024) --p.w_group;
  p.w_slot_bytes = slot_bytes(p.w_group);
  p.w_slots = (222 * 1024 - 2048 - kAStages * p.a_stage_bytes) / p.w_slot_bytes;
  if (p.w_slots > kWSlots) p.w_slots = kWSlots;
  WS_REQUIRE(p.w_slots >= 2, "rdb_persist: no room for two weight slots");
  RdbMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i <= d->nconv; ++i) {
    const bool lff = i == d->nconv;
    const int cin = lff ? p.ctot : d->f + i * d->gc;
    p.kchunks[i] = (cin + 63) / 64;
    p.last_k16[i] = (cin - 64 * (p.kchunks[i] - 1) + 15) / 16;
    if (int e = make_act_map(buf, cin, d->x, d->y, d->z, d->n, lff ? d->y : d->y + 2, lff ? 1 : 3, &maps.a[i])) return e;
    const int k_pad = (cin + 7) / 8 * 8;
    if (int e = make_w_map(packed[i], k_pad, lff ? (d->f + 15) / 16 * 16 : p.n_conv, lff ? 1 : 9, &maps.b[i])) return e;
  }
  static const bool grid_bar = env_off("WS_RDB_GRID_BARRIER");
  p.nb_sync = (!grid_bar && d->n * d->x <= 256) ? 1 : 0;
  static const bool no_early = env_off("WS_RDB_NO_EARLY");
  p.early = (!no_early && 2 * p.t_m * 128 <= 512) ? 1 : 0;
  p.acc_stride = p.early ? p.t_m * 128 : 0;
  uint32_t cols = 32;
  while ((int)cols < (p.early ? 2 : 1) * p.t_m * 128) cols <<= 1;
  p.tmem_cols = cols;
  // the MMA rows of the last tile / the largest tap offset read past the loaded box (rows that are never stored):
  // they must still lie inside this CTA's shared memory
  const size_t reach = (size_t)(p.t_m * 128 + 2 * p.slab_p + 2 * p.DZ) * 128;
  size_t smem = (size_t)kAStages * p.a_stage_bytes + (size_t)p.w_slots * p.w_slot_bytes + 8 * (2 * kAStages + 2 * kWSlots + 3) + 1024;
  if (smem < (size_t)p.a_stage_bytes + reach + 1024) smem = (size_t)p.a_stage_bytes + reach + 1024;
  WS_REQUIRE(smem <= 224 * 1024, "rdb_persist: shared memory request %zu too large", smem);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute((const void*)rdb_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    224 * 1024);  // + 2 KB static
  });
  WS_REQUIRE(attr_err == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(d->n * d->x));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rdb_fwd_persist_kernel, maps, p, x, buf, out, ep_lff));
  WS_POST_LAUNCH(1);
  return 0;
}

// Backward data-gradient chain.  packed[i], i < nconv: dgrad packing [flipped tap][cin_pad16][gc_pad8]; packed[nconv]:
// LFF dgrad packing [1][ctot_pad16][F_pad8].  Additional requirements: gc == 32 (one SWIZZLE_64B K chunk per tap) and
// 2 tiles x ctot_pad16 <= 512 TMEM columns.
bool rdb_persist_bwd_ok(const ws_rdb_desc* d, const View& dy, const View& buf, const View& g_lff, const View& gbuf,
                        const View& dx, int sm_count) {
  static const bool off = env_off("WS_DISABLE_RDB_PERSIST_BWD");
  if (off || !rdb_persist_ok(d, dy, buf, sm_count)) return false;
  if (d->gc != 32 || !dx.ptr) return false;
  const int ctot_pad = (d->f + d->nconv * d->gc + 15) / 16 * 16;
  const int pz = d->z + 2;
  const int span = (d->y - 1) * pz + d->z;
  const int t_m = (span + 127) / 128;
  if (ctot_pad > 256 || t_m * ctot_pad > 512 || pz > 256) return false;
  auto bf16_ok = [](const View& v) {
    return v.dtype == WS_BF16 && v.cs == 1 && !(reinterpret_cast<uintptr_t>(v.ptr) & 15) && (v.vs * 2) % 16 == 0 &&
           (v.ns * 2) % 16 == 0;
  };
  if (!bf16_ok(g_lff) || !bf16_ok(gbuf)) return false;
  if (dx.dtype != WS_F32 || dx.cs != 1 || (reinterpret_cast<uintptr_t>(dx.ptr) & 15) || (dx.vs * 4) % 16 ||
      (dx.ns * 4) % 16)
    return false;
  return true;
}

int rdb_persist_backward(const ws_rdb_desc* d, const View& dy, const View& buf, const View& g_lff, const View& gbuf,
                         const View& dx, void* const* packed, cudaStream_t st) {
  RdbBwdParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n; p.DX = d->x; p.DY = d->y; p.DZ = d->z;
  p.F = d->f; p.gc = d->gc; p.nconv = d->nconv; p.ctot = d->f + d->nconv * d->gc;
  p.ctot_pad = (p.ctot + 15) / 16 * 16;
  p.pz = d->z + 2;
  p.slab = d->y * d->z;
  p.slab_p = (d->y + 2) * p.pz;
  p.t_m = ((d->y - 1) * p.pz + d->z + 127) / 128;
  p.slope = d->slope; p.alpha = d->alpha; p.beta1 = d->beta1;
  p.bar_slot = 1;
  {
    static long long* dbg_buf = nullptr;
    static int dbg_calls = 0;
    static const bool dbg_on = env_off("WS_RDB_DEBUG_TIMES");
    if (dbg_on) {
      if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
      if (++dbg_calls == 40) {  // a warm call
        cudaStreamSynchronize(st);
        long long h[64];
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        const int nst = 2 + 4 * d->nconv + 2;
        fprintf(stderr, "[rdb_bwd_persist] phase-edge stamps of an interior CTA (us since kernel start):");
        for (int i = 1; i < nst && i < 64; ++i) fprintf(stderr, " %.1f", (h[i] - h[0]) * 1e-3);
        fprintf(stderr, "\n  order: g_lff written | per phase (LFF, conv n-1 .. 1): mma-done, stores issued, visible, neighbours | conv 0: mma-done, dx written\n");
      }
      p.dbg = dbg_buf;
    }
  }
  p.kch_lff = (d->f + 63) / 64;
  p.last_k16_lff = (d->f - 64 * (p.kch_lff - 1) + 15) / 16;
  p.a_box_bytes_conv = 3 * p.slab_p * 64;
  p.a_box_bytes_lff = d->y * p.pz * 128;
  int a_bytes = p.a_box_bytes_conv > p.a_box_bytes_lff ? p.a_box_bytes_conv : p.a_box_bytes_lff;
  p.a_stage_bytes = (a_bytes + 1023) / 1024 * 1024;
  int w_bytes = p.ctot_pad * 128;
  RdbMaps maps;
  memset(&maps, 0, sizeof(maps));
  // LFF: A = g_lff rows at the z pitch of the chain (pad rows zero-filled by TMA), B = [ctot_pad][F_pad8]
  if (int e = make_act_map(g_lff, d->f, d->x, d->y, d->z, d->n, d->y, 1, &maps.a[0], p.pz)) return e;
  if (int e = make_w_map(packed[d->nconv], (d->f + 7) / 8 * 8, p.ctot_pad, 1, &maps.b[0])) return e;
  for (int i = 0; i < d->nconv; ++i) {
    p.cin[i] = d->f + i * d->gc;
    View gi = gbuf;
    gi.ptr = (char*)gbuf.ptr + (size_t)i * d->gc * 2;
    if (int e = make_act_map(gi, d->gc, d->x, d->y, d->z, d->n, d->y + 2, 3, &maps.a[1 + i], p.pz, 32)) return e;
    if (int e = make_w_map(packed[i], (d->gc + 7) / 8 * 8, (p.cin[i] + 15) / 16 * 16, 27, &maps.b[1 + i], 32)) return e;
    if (p.cin[i] * 64 > w_bytes) w_bytes = p.cin[i] * 64;
  }
  {
    static const bool grid_bar = env_off("WS_RDB_GRID_BARRIER");
    p.nb_sync = (!grid_bar && d->n * d->x <= 256) ? 1 : 0;
    static const int env_wg = getenv("WS_RDB_WGROUP") ? atoi(getenv("WS_RDB_WGROUP")) : 3;
    static const int env_niss = getenv("WS_RDB_NISS") ? atoi(getenv("WS_RDB_NISS")) : 4;
    p.n_iss_m = p.t_m < 4 ? p.t_m : 4;
    if (p.n_iss_m > env_niss) p.n_iss_m = env_niss < 1 ? 1 : env_niss;
    // (two issuers adding to ONE accumulator do so in a timing-dependent order: run-to-run fp32 noise in dL/dx — opt-in)
    static const bool env_split = env_off("WS_RDB_TAP_SPLIT");
    p.tap_split = (env_split && p.n_iss_m * 2 <= 4 && p.n_iss_m * 2 <= env_niss) ? 2 : 1;
    const int lff_tile = p.ctot_pad * 128;
    int conv_tile = 0;
    for (int i = 0; i < d->nconv; ++i)
      if (p.cin[i] * 64 > conv_tile) conv_tile = p.cin[i] * 64;
    p.w_group = env_wg < 1 ? 1 : (env_wg > 27 ? 27 : env_wg);
    auto slot_bytes = [&](int wg) { int b = wg * conv_tile; if (lff_tile > b) b = lff_tile; return (b + 1023) / 1024 * 1024; };
    while (p.w_group > 1 && kAStages * p.a_stage_bytes + 2 * slot_bytes(p.w_group) + 2048 > 222 * 1024) --p.w_group;
    p.w_slot_bytes = slot_bytes(p.w_group);
    p.w_slots = (222 * 1024 - 2048 - kAStages * p.a_stage_bytes) / p.w_slot_bytes;
    if (p.w_slots > kWSlots) p.w_slots = kWSlots;
    WS_REQUIRE(p.w_slots >= 2, "rdb_persist backward: no room for two weight slots");
    (void)w_bytes;
  }
  uint32_t cols = 32;
  while ((int)cols < p.t_m * p.ctot_pad) cols <<= 1;
  p.tmem_cols = cols;
  const size_t reach = (size_t)(p.t_m * 128 + 2 * p.slab_p + 2 * p.pz + 2) * 128;  // rows x the wider (128 B) row
  size_t smem = (size_t)kAStages * p.a_stage_bytes + (size_t)p.w_slots * p.w_slot_bytes + 8 * (2 * kAStages + 2 * kWSlots + 3) + 1024;
  if (smem < (size_t)p.a_stage_bytes + reach + 1024) smem = (size_t)p.a_stage_bytes + reach + 1024;
  WS_REQUIRE(smem <= 225 * 1024, "rdb_persist backward: shared memory request %zu too large", smem);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute((const void*)rdb_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    225 * 1024);
  });
  WS_REQUIRE(attr_err == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(d->n * d->x));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rdb_bwd_persist_kernel, maps, p, dy, buf, g_lff, gbuf, dx));
  WS_POST_LAUNCH(1);
  return 0;
}

}  // namespace ws
