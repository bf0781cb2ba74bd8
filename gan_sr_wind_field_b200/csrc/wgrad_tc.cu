// wgrad_tc.cu — Conv3d weight gradient as a tcgen05 GEMM with MN-major (transposed) operands.
//
//   dW[tap][co][ci] = sum_{n, v} dy(n, v, co) * x(n, v (+) tap, ci)
//
// Both operands are channels-last activations, so the reduction index (voxels) is the strided one: the
// TMA boxes {64 ch, bz, by, bx, 1} land as R voxel rows x 128 B, which is exactly the canonical *MN-major*
// SWIZZLE_128B UMMA tile (64 contiguous M/N elements per row, 8-row K atoms, SBO = 1024 B, LBO = one
// 64-channel block = R*128 B).  One MMA consumes 16 voxel rows; the tap shift / zero padding / stride of
// the convolution is again the TMA coordinate, out-of-bounds fill and elementStrides of the x operand.
//
// A CTA owns one 128-row M tile, a group of T taps (T * N <= 512 TMEM columns: one fp32 accumulator per
// tap) and a slice of the voxel tiles (split-K).  The "unshifted" operand tile (dy) is loaded once per voxel
// tile into its own 2-slot ring and reused by all T taps; the shifted operand streams through a second ring.
// Partial results are reduced into an fp32 workspace with red.global.add, then a finalize kernel transposes
// into torch's (cout, cin, kx, ky, kz) layout.
//
// Roles may be swapped (M = cin, N = cout) so that the narrow side (e.g. Cout = 32 of the RDB convs,
// torch_blocks.py:256-267) sits on the UMMA N dimension where it costs proportionally less.
//
// Replaces the weight-gradient half of autograd's convolution_backward for torch_blocks.py:17,278 and
// Generator_3D_Resnet_ESRGAN.py:105.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include <mutex>
#include <array>
#include <map>

namespace ws {

namespace {

constexpr int kThreads = 288;  // warp 0: TMA producer, warps 1-4: MMA issuers (warp 1 owns TMEM), warps 5-8: epilogue
constexpr int kMaxBSlots = 6;

struct WgParams {
  int N;                 // batch
  int bx, by, bz, rows;  // voxel tile (in OUTPUT / dy coordinates); rows % 16 == 0
  int tiles_x, tiles_y, tiles_z;
  long long total_tiles;       // N * tiles
  long long items;             // ranges * tap_groups * range_tiles: the (tile range, tap group, tile) work items of one M block
  long long items_per_cta;     // each CTA takes a contiguous run of items -> one or more (tap group, tile range) segments
  long long range_tiles;       // voxel tiles per range (the last range may be short: tiles >= total_tiles are skipped)
  int kx, ky, kz, sx, sy, sz, px, py, pz;
  int taps, taps_per_cta, tap_groups;
  int m_total;           // channels of the M operand covered by this launch (grid.z blocks of 128)
  int n_blocks;          // 64-channel blocks of the N operand
  int m0, n0;            // channel offsets of this launch inside the full cout / cin ranges
  int m_valid, n_valid;  // valid rows / columns of the accumulator
  int n_umma;
  int shift_on_m;        // 1: the M operand is x (shifted), 0: the N operand is x
  int tf32;              // 1: fp32 operands, kind::tf32 — 32 channels per 128-byte row, 8 voxel rows per instruction
  int a_slot_bytes, b_slot_bytes, b_slots;
  uint32_t tmem_cols;
  // workspace addressing: wsp[range * range_stride + tap * tap_stride + m * m_stride + n * n_stride]
  long long tap_stride, m_stride, n_stride;
  long long range_stride;  // batched launches (one tile range per residual dense block): a result per range
  int xhalo;               // 1: a tap group = taps_per_cta consecutive kx taps of one (ky, kz); the shifted operand is ONE
                           // box with an x halo (bx + taps_per_cta - 1 slabs) per voxel tile and the taps are row offsets
                           // into it (rows are ordered (x, y, z): a shift by one x slab = by * bz rows) — the L2 -> SMEM
                           // traffic of the shifted operand drops by ~taps_per_cta
  int halo_rows;           // rows of the halo box, halo_blk_bytes = its padded size per channel block
  int halo_blk_bytes;
  int kx_groups;           // xhalo: tap groups per (ky, kz) = ceil(kx / taps_per_cta)
  int n_iss;               // MMA-issuing warps (1..4, <= taps_per_cta): issuer q owns the accumulators of taps q, q + n_iss, ...
                           // (a tcgen05.mma occupies its issuing thread about as long as it executes, so barrier waits,
                           // commits and descriptor arithmetic of ONE issuer are tensor-pipe idle time — conv_tc2.cu)
  int plain_store;         // 1: every output element has exactly one segment -> st.global instead of red.global.add
  int mz;                  // > 0: the M block index is blockIdx.x % mz (CTAs of one range adjacent in launch order)
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
                const __grid_constant__ CUtensorMap tmH, const WgParams p, float* __restrict__ wsp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = ptx::smem_u32(smem);
  // layout: [A slot 0][A slot 1][B slots...][barriers]
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + 2u * p.a_slot_bytes;
  const uint32_t bar_off = 2u * p.a_slot_bytes + (uint32_t)p.b_slots * p.b_slot_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (4 + kMaxBSlots + s); };
  const uint32_t accum_bar = bar_base + 8u * (4 + 2 * kMaxBSlots);
  const uint32_t tmem_free = bar_base + 8u * (5 + 2 * kMaxBSlots);
  const uint32_t tmem_slot = bar_base + 8u * (6 + 2 * kMaxBSlots);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * (6 + 2 * kMaxBSlots));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mblk = p.mz > 0 ? (int)(blockIdx.x % (unsigned)p.mz) : (int)blockIdx.z;
  const long long bx_lin = p.mz > 0 ? (long long)(blockIdx.x / (unsigned)p.mz) : (long long)blockIdx.x;
  const int m0 = p.m0 + mblk * 128;                                  // this CTA's 128-row M block
  const int m_valid = min(128, p.m_total - mblk * 128);
  const int cblk = p.tf32 ? 32 : 64;  // channels per 128-byte operand row
  const int m_blocks = (m_valid + cblk - 1) / cblk;
  // Work = (tile range, tap group, tile) items in that order; CTA b owns items [b*per, (b+1)*per): the same load for
  // every CTA however taps and tiles divide (hr_convs.0: 63 groups x 2 splits left 22 of 148 SMs idle), and the CTAs
  // of a wave that share a tile range walk the same tiles at the same time, so a tile is fetched from DRAM once per
  // pass and re-used out of L2 by the other tap groups (ncu: 27 GB of DRAM reads per launch in group-major order).
  // A run that crosses a (range, group) boundary is processed as consecutive SEGMENTS, each with its own accumulate
  // / reduce phase.
  const long long item_lo = bx_lin * p.items_per_cta;
  const long long item_hi = min(p.items, item_lo + p.items_per_cta);

  if (threadIdx.x == 0) ptx::griddep_launch();  // programmatic dependent launch, see conv_tc2.cu
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmM);
    ptx::prefetch_tmap(&tmN);
    if (p.xhalo) ptx::prefetch_tmap(&tmH);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), (uint32_t)p.n_iss); }
    // (xhalo: every issuer reads every halo slot; otherwise a slot belongs to one issuer's private ring)
    for (int s = 0; s < p.b_slots; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), p.xhalo ? (uint32_t)p.n_iss : 1u); }
    ptx::mbar_init(accum_bar, (uint32_t)p.n_iss);
    ptx::mbar_init(tmem_free, 4);  // one arrival per epilogue warp
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
  const bool has_work = item_hi > item_lo;

  const uint32_t blk_bytes = (uint32_t)p.rows * 128u;  // one 64-channel block of one voxel tile
  const int tiles_per_n = p.tiles_x * p.tiles_y * p.tiles_z;

  // Both loops run warp-uniformly and elect one lane only around UTMALDG / UTCHMMA (see conv_tc2.cu).
  if (warp == 0) {
    ptx::griddep_wait();
    if (has_work) {
      // ===== TMA producer =====
      int as = 0;
      uint32_t aph = 0;
      // the ring of shifted-operand slots is split into one private ring per issuing warp (slots_per slots each): a
      // barrier then has exactly one consumer, which sees every one of its phases in order
      const int slots_per = p.b_slots / p.n_iss;
      int hs = 0;           // xhalo: position / phase in the shared ring of halo slots
      uint32_t hph = 0;
      uint32_t bstate = 0;  // per issuer q a nibble: position in its private ring (3 bits) | phase << 3
      // the unshifted operand (dy) is the one that is NOT x
      const CUtensorMap* tm_fix = p.shift_on_m ? &tmN : &tmM;
      const CUtensorMap* tm_sh = p.shift_on_m ? &tmM : &tmN;
      const int fix_blocks = p.shift_on_m ? p.n_blocks : m_blocks;
      const int sh_blocks = p.shift_on_m ? m_blocks : p.n_blocks;
      const int fix_c0 = p.shift_on_m ? p.n0 : m0;
      const int sh_c0 = p.shift_on_m ? m0 : p.n0;
      for (long long item = item_lo; item < item_hi; ++item) {
        const long long sg = item / p.range_tiles;
        const long long tile = (sg / p.tap_groups) * p.range_tiles + item % p.range_tiles;
        if (tile >= p.total_tiles) continue;  // tail of the last range
        const int tap_lo = (int)(sg % p.tap_groups) * p.taps_per_cta;
        const int tap_hi = min(p.taps, tap_lo + p.taps_per_cta);
        int t = (int)(tile % tiles_per_n);
        const int n = (int)(tile / tiles_per_n);
        const int tz = t % p.tiles_z; t /= p.tiles_z;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int tx = t;
        const int x0 = tx * p.bx, y0 = ty * p.by, z0 = tz * p.bz;
        ptx::mbar_wait(a_empty(as), aph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(a_full(as), blk_bytes * fix_blocks);
          for (int b = 0; b < fix_blocks; ++b)
            ptx::tma_load_5d(a_base + as * p.a_slot_bytes + b * blk_bytes, tm_fix, a_full(as), fix_c0 + b * cblk, z0,
                             y0, x0, n);
        }
        __syncwarp();
        if (++as == 2) { as = 0; aph ^= 1u; }
        if (p.xhalo) {
          // one halo box for all taps of the group: x from the group's first kx tap, (ky, kz) fixed
          const int g = (int)(sg % p.tap_groups);
          const int yz = g / p.kx_groups, ti0 = (g - yz * p.kx_groups) * p.taps_per_cta;
          const int tj = yz / p.kz, tl = yz - tj * p.kz;
          const int cx = x0 - p.px + ti0, cy = y0 * p.sy - p.py + tj, cz = z0 * p.sz - p.pz + tl;
          const int bs = hs;
          ptx::mbar_wait(b_empty(bs), hph ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(b_full(bs), (uint32_t)p.halo_rows * 128u * sh_blocks);
            for (int b = 0; b < sh_blocks; ++b)
              ptx::tma_load_5d(b_base + bs * p.b_slot_bytes + b * p.halo_blk_bytes, &tmH, b_full(bs), sh_c0 + b * cblk, cz,
                               cy, cx, n);
          }
          __syncwarp();
          if (++hs == p.b_slots) { hs = 0; hph ^= 1u; }
          continue;
        }
        for (int tap = tap_lo; tap < tap_hi; ++tap) {
          const int ti = tap / (p.ky * p.kz), tj = (tap / p.kz) % p.ky, tl = tap % p.kz;
          const int cx = x0 * p.sx - p.px + ti, cy = y0 * p.sy - p.py + tj, cz = z0 * p.sz - p.pz + tl;
          const int q = (tap - tap_lo) % p.n_iss;
          const uint32_t nib = (bstate >> (4 * q)) & 0xfu;
          const int pos = (int)(nib & 7u);
          const uint32_t bph = nib >> 3;
          const int bs = q * slots_per + pos;
          const uint32_t nxt = pos + 1 == slots_per ? ((bph ^ 1u) << 3) : (nib + 1u);
          bstate = (bstate & ~(0xfu << (4 * q))) | (nxt << (4 * q));
          ptx::mbar_wait(b_empty(bs), bph ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(b_full(bs), blk_bytes * sh_blocks);
            for (int b = 0; b < sh_blocks; ++b)
              ptx::tma_load_5d(b_base + bs * p.b_slot_bytes + b * blk_bytes, tm_sh, b_full(bs), sh_c0 + b * cblk, cz,
                               cy, cx, n);
          }
          __syncwarp();
        }
      }
    }
  }
  if (warp >= 1 && warp - 1 < p.n_iss) {
    if (has_work) {
      // ===== MMA issuers: warps 1 .. n_iss =====
      const int q = warp - 1;
      const uint32_t idesc = ptx::make_idesc(p.tf32 ? 2u : 1u, 128u, (uint32_t)p.n_umma, 1u, 1u);  // both MN-major
      // bf16: SWIZZLE_128B, K atoms of 8 rows (SBO = 1024 B).  tf32: the MN-major operand must use the 32-byte-atom
      // flavour of the 128-byte swizzle (layout type 1), whose K atoms are 4 rows (SBO = 512 B); LBO = one channel block
      const uint64_t desc_hi =
          p.tf32 ? (((uint64_t)((blk_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) |
                    ((uint64_t)1 << 61))
                 : ptx::make_smem_desc_sw128(0, blk_bytes, 1024);
      int as = 0;
      uint32_t aph = 0;
      const int slots_per = p.b_slots / p.n_iss;
      int bpos = 0;  // position / phase in this issuer's private ring (xhalo: in the shared ring of halo slots)
      uint32_t bph = 0;
      // the shifted operand's tile has its own row count (LBO = bytes of one channel block)
      const uint64_t desc_hi_sh = (p.xhalo && !p.tf32) ? ptx::make_smem_desc_sw128(0, (uint32_t)p.halo_blk_bytes, 1024) : desc_hi;
      const uint32_t slab_bytes = (uint32_t)(p.by * p.bz) * 128u;  // xhalo: one x slab of the halo box
      // one instruction consumes 16 voxel rows of bf16 (2048 B of the MN-major tile) or 8 rows of tf32 (1024 B)
      const int krows = p.tf32 ? 8 : 16;
      const int k16 = p.rows / krows;
      const uint64_t kadv = (uint64_t)(krows * 128 / 16);
      uint32_t acc_tile = 0;
      int seg = 0;
      for (long long item = item_lo; item < item_hi; ++item) {
        const long long sg = item / p.range_tiles;
        const long long tile = (sg / p.tap_groups) * p.range_tiles + item % p.range_tiles;
        int ntap;
        if (p.xhalo) {
          const int g = (int)(sg % p.tap_groups);
          const int ti0 = (g % p.kx_groups) * p.taps_per_cta;
          ntap = min(p.kx, ti0 + p.taps_per_cta) - ti0;
        } else {
          const int tap_lo = (int)(sg % p.tap_groups) * p.taps_per_cta;
          ntap = min(p.taps, tap_lo + p.taps_per_cta) - tap_lo;
        }
        if (item > item_lo && item % p.range_tiles == 0) {
          // group boundary: hand the finished accumulators to the epilogue, then wait until it has drained them
          if (ptx::elect_one()) ptx::mma_commit(accum_bar);
          __syncwarp();
          ptx::mbar_wait(tmem_free, (uint32_t)(seg & 1));
          ptx::tc_fence_after();
          ++seg;
          acc_tile = 0;
        }
        if (tile >= p.total_tiles) continue;  // tail of the last range (never the first item of a segment)
        ptx::mbar_wait(a_full(as), aph);
        const uint32_t fix_addr = a_base + as * p.a_slot_bytes;
        if (p.xhalo) {
          // every issuer follows every halo slot; its taps are row offsets into the box
          const int bs = bpos;
          ptx::mbar_wait(b_full(bs), bph);
          ptx::tc_fence_after();
          for (int tp = q; tp < ntap; tp += p.n_iss) {
            const uint32_t sh_addr = b_base + bs * p.b_slot_bytes + (uint32_t)tp * slab_bytes;
            const uint64_t sh_desc = desc_hi_sh | (uint64_t)((sh_addr >> 4) & 0x3fffu);
            const uint64_t fx_desc = desc_hi | (uint64_t)((fix_addr >> 4) & 0x3fffu);
            const uint64_t adesc = p.shift_on_m ? sh_desc : fx_desc;
            const uint64_t bdesc = p.shift_on_m ? fx_desc : sh_desc;
            const uint32_t d_tmem = tmem_base + (uint32_t)(tp * p.n_umma);
            if (ptx::elect_one()) {
              for (int k = 0; k < k16; ++k)
                ptx::mma_f16_ss(d_tmem, adesc + kadv * k, bdesc + kadv * k, idesc, (acc_tile | (uint32_t)k) ? 1u : 0u);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx::mma_commit(b_empty(bs));
          __syncwarp();
          if (++bpos == p.b_slots) { bpos = 0; bph ^= 1u; }
        } else
        for (int tp = q; tp < ntap; tp += p.n_iss) {  // the other taps belong to other issuers (and their rings)
          const int bs = q * slots_per + bpos;
          ptx::mbar_wait(b_full(bs), bph);
          ptx::tc_fence_after();
          const uint32_t sh_addr = b_base + bs * p.b_slot_bytes;
          const uint32_t m_addr = p.shift_on_m ? sh_addr : fix_addr;
          const uint32_t n_addr = p.shift_on_m ? fix_addr : sh_addr;
          const uint64_t adesc = desc_hi | (uint64_t)((m_addr >> 4) & 0x3fffu);
          const uint64_t bdesc = desc_hi | (uint64_t)((n_addr >> 4) & 0x3fffu);
          const uint32_t d_tmem = tmem_base + (uint32_t)(tp * p.n_umma);
          if (ptx::elect_one()) {
            // one MMA per `krows` voxel rows: the descriptor start advances by krows * 128 B
            if (p.tf32) {
              for (int k = 0; k < k16; ++k)
                ptx::mma_tf32_ss(d_tmem, adesc + kadv * k, bdesc + kadv * k, idesc, (acc_tile | (uint32_t)k) ? 1u : 0u);
            } else {
              for (int k = 0; k < k16; ++k)
                ptx::mma_f16_ss(d_tmem, adesc + kadv * k, bdesc + kadv * k, idesc, (acc_tile | (uint32_t)k) ? 1u : 0u);
            }
            ptx::mma_commit(b_empty(bs));
          }
          __syncwarp();
          if (++bpos == slots_per) { bpos = 0; bph ^= 1u; }
        }
        if (ptx::elect_one()) ptx::mma_commit(a_empty(as));
        __syncwarp();
        if (++as == 2) { as = 0; aph ^= 1u; }
        acc_tile = 1;
      }
      if (ptx::elect_one()) ptx::mma_commit(accum_bar);
      __syncwarp();
    }
  }
  if (warp >= 5 && has_work) {
    // ===== epilogue: TMEM -> red.global.add into the fp32 workspace, once per segment =====
    const int sub = warp & 3;
    const int m = sub * 32 + lane;
    ptx::griddep_wait();  // the workspace memset / earlier reductions precede the atomics
    int seg = 0;
    for (long long item = item_lo; item < item_hi; ++seg) {
      // accumulator tp of this segment belongs to tap  tap_lo + tp * tap_step
      int tap_lo, ntap, tap_step = 1;
      if (p.xhalo) {
        const int g = (int)((item / p.range_tiles) % p.tap_groups);
        const int yz = g / p.kx_groups, ti0 = (g - yz * p.kx_groups) * p.taps_per_cta;
        ntap = min(p.kx, ti0 + p.taps_per_cta) - ti0;
        tap_step = p.ky * p.kz;
        tap_lo = ti0 * tap_step + yz;  // (ti * ky + tj) * kz + tl with yz = tj * kz + tl
      } else {
        tap_lo = (int)((item / p.range_tiles) % p.tap_groups) * p.taps_per_cta;
        ntap = min(p.taps, tap_lo + p.taps_per_cta) - tap_lo;
      }
      const long long seg_end = min(item_hi, (item / p.range_tiles + 1) * p.range_tiles);
      float* wrange = wsp + ((item / p.range_tiles) / p.tap_groups) * p.range_stride;
      ptx::mbar_wait(accum_bar, (uint32_t)(seg & 1));
      ptx::tc_fence_after();
      for (int tp = 0; tp < ntap; ++tp) {
        const int tap = tap_lo + tp * tap_step;
        for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
          if (c0 >= p.n_valid) break;
          uint32_t r[16];
          ptx::tmem_ld16(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(tp * p.n_umma + c0), r);
          ptx::tmem_ld_wait();
          if (m < m_valid) {
            float* dst = wrange + (long long)tap * p.tap_stride + (long long)(m0 + m) * p.m_stride +
                         (long long)(p.n0 + c0) * p.n_stride;
            if (p.plain_store) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.n_valid) dst[(long long)j * p.n_stride] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.n_valid) atomicAdd(dst + (long long)j * p.n_stride, __uint_as_float(r[j]));
            }
          }
        }
      }
      item = seg_end;
      if (item < item_hi) {  // more segments follow: release the accumulators
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tmem_free);
      }
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

void choose_wgrad_tile_search(int DX, int DY, int DZ, int sx, int sy, int sz, int max_rows, int& bx, int& by, int& bz) {
  double best = -1.0;
  bx = by = 1; bz = 16;
  for (int z = 1; z <= max_rows; ++z) {
    if (z > DZ && z != 16) continue;  // allow a padded z only as the fallback
    if ((z - 1) * sz + 1 > 256) break;
    for (int y = 1; z * y <= max_rows; ++y) {
      if ((y - 1) * sy + 1 > 256) break;
      if (y > 2 * DY) break;
      for (int x = 1; z * y * x <= max_rows; ++x) {
        if ((x - 1) * sx + 1 > 256) break;
        if (x > 2 * DX) break;
        int rows = z * y * x;
        if (rows % 16) continue;
        long long tiles = (long long)((DX + x - 1) / x) * ((DY + y - 1) / y) * ((DZ + z - 1) / z);
        double eff = (double)DX * DY * DZ / ((double)tiles * rows);
        double score = eff + 1e-4 * rows / (double)max_rows;
        if (score > best) { best = score; bx = x; by = y; bz = z; }
      }
    }
  }
}

// max_rows: voxel rows of one operand tile (128; 64 in TF32 mode, whose 128-byte rows hold half as many channels)
void choose_wgrad_tile(int DX, int DY, int DZ, int sx, int sy, int sz, int max_rows, int& bx, int& by, int& bz) {
  struct Hit { int bx, by, bz; };
  static std::mutex mu;
  static std::map<std::array<int, 7>, Hit> memo;  // keyed by the full geometry (no hash collisions)
  const std::array<int, 7> key = {DX, DY, DZ, sx, sy, sz, max_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(key);
    if (it != memo.end()) { bx = it->second.bx; by = it->second.by; bz = it->second.bz; return; }
  }
  choose_wgrad_tile_search(DX, DY, DZ, sx, sy, sz, max_rows, bx, by, bz);
  std::lock_guard<std::mutex> lk(mu);
  memo[key] = Hit{bx, by, bz};
}

}  // namespace

size_t tc_wgrad_workspace_bytes(const ConvGeom& g) {
  return (size_t)g.taps() * g.cin * g.cout * sizeof(float);
}

// One launch: M rows = channels [m0, m0+m_count) of the M operand, N cols = channels [n0, n0+n_count) of the
// N operand.  swap == 0: M = cout (dy), N = cin (x).  swap == 1: M = cin (x), N = cout (dy).
// batch: the launch covers `ranges` independent problems of identical geometry laid out back to back along the batch
// dimension (g.n = ranges * samples per problem): one tile range per problem, results `range_stride` floats apart,
// written with plain stores (every output element belongs to exactly one segment).
struct WgBatch {
  int ranges;
  long long range_stride;
};
static int launch_one(const ConvGeom& g_in, const View& x, const View& dy, float* wsp, int swap, int m_total,
                      int n0, int n_count, cudaStream_t st, const WgBatch* batch = nullptr) {
  // A conv without taps along x (kx == 1: the x-folded remainder of hr_convs.0's gradient is a (1,5,5) conv) is
  // handled with the roles of x and y exchanged — a pure relabelling (tap index (ti*ky + tj)*kz + tl is the same
  // number either way when one of the two extents is 1; the tensor maps swap the two strides) — so that its ky taps
  // get the halo treatment of WgParams::xhalo.
  ConvGeom g = g_in;
  const bool swapxy = g_in.kx == 1 && g_in.ky > 1 && g_in.sx == 1 && g_in.sy == 1 && !getenv("WS_WGRAD_NO_SWAPXY");
  if (swapxy) {
    g.x = g_in.y; g.y = g_in.x; g.xo = g_in.yo; g.yo = g_in.xo;
    g.kx = g_in.ky; g.ky = g_in.kx; g.px = g_in.py; g.py = g_in.px;
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.N = g.n;
  choose_wgrad_tile(g.xo, g.yo, g.zo, g.sx, g.sy, g.sz, x.dtype == WS_F32 ? 64 : 128, p.bx, p.by, p.bz);
  p.rows = p.bx * p.by * p.bz;
  p.tiles_x = (g.xo + p.bx - 1) / p.bx;
  p.tiles_y = (g.yo + p.by - 1) / p.by;
  p.tiles_z = (g.zo + p.bz - 1) / p.bz;
  p.total_tiles = (long long)g.n * p.tiles_x * p.tiles_y * p.tiles_z;
  p.kx = g.kx; p.ky = g.ky; p.kz = g.kz; p.sx = g.sx; p.sy = g.sy; p.sz = g.sz;
  p.px = g.px; p.py = g.py; p.pz = g.pz;
  p.taps = g.taps();
  p.m0 = 0; p.m_total = m_total; p.n0 = n0; p.n_valid = n_count;
  p.n_umma = (n_count + 15) / 16 * 16;
  WS_REQUIRE(p.n_umma <= 256, "wgrad N tile too large (%d)", p.n_umma);
  const int mz = (m_total + 127) / 128;
  const bool tf32 = x.dtype == WS_F32;
  const int esize = tf32 ? 4 : 2, cblk = 128 / esize;
  p.tf32 = tf32 ? 1 : 0;
  p.n_blocks = (p.n_umma + cblk - 1) / cblk;
  p.shift_on_m = swap;
  const int blk = p.rows * 128;
  // the M operand tile must always present 128 channels worth of address space (UMMA M = 128 reads all of its blocks)
  const int fix_alloc = swap ? p.n_blocks : 128 / cblk;
  const int sh_alloc = swap ? 128 / cblk : p.n_blocks;
  p.a_slot_bytes = fix_alloc * blk;
  p.b_slot_bytes = sh_alloc * blk;
  const int budget = 220 * 1024 - 1024 - 512 - 2 * p.a_slot_bytes;
  // x-halo mode (WgParams::xhalo): stride-1 bf16 layers with kx >= 2 whose tap groups are runs of kx taps
  static const bool env_no_xhalo = getenv("WS_WGRAD_XHALO") && atoi(getenv("WS_WGRAD_XHALO")) == 0;
  const bool xh_ok = !tf32 && g.sx == 1 && g.kx >= 2 && !env_no_xhalo;
  auto halo_blk = [&](int tpc) { return ((p.bx + tpc - 1) * p.by * p.bz * 128 + 1023) / 1024 * 1024; };
  auto xh_for = [&](int tpc) {
    return xh_ok && tpc >= 2 && tpc <= g.kx && (p.bx + tpc - 1) <= 256 && 2 * sh_alloc * halo_blk(tpc) <= budget;
  };

  // Work split: (tap group, voxel tile) items of an M block dealt out in equal contiguous runs to `ncta` CTAs
  // (grid.x), M blocks in grid.z.  Modelled cost per CTA =
  //   items * (taps_per_cta * max(MMA, L2->SMEM) + unshifted tile load) + segments * (reduce epilogue) + fixed,
  // MMA at the measured max(72, N/2) cycles per K=16 step.  Few taps per CTA keep the atomic epilogue short.
  const double per_mma = p.n_umma / 2.0 > 72.0 ? p.n_umma / 2.0 : 72.0;
  const double unit_mma = (p.rows / 16) * per_mma;
  const double unit_load = (double)sh_alloc * blk / 33.0;
  const double unit = unit_mma > unit_load ? unit_mma : unit_load;
  const double atom = 128.0 * p.n_umma * 0.35;  // coalesced red.global.add epilogue of one tap (measured)
  const int max_tpc = 512 / p.n_umma;
  double best = 1e30;
  int best_tpc = 1, best_gpc = 1;
  long long best_range = p.total_tiles;
  bool best_xh = false;
  for (int tpc = 1; tpc <= max_tpc && tpc <= p.taps; ++tpc) {
    const bool xh = xh_for(tpc);
    if (xh_ok && tpc > g.kx) break;  // (with an x halo available, groups that mix (ky, kz) are never better)
    const int tg = xh ? g.ky * g.kz * ((g.kx + tpc - 1) / tpc) : (p.taps + tpc - 1) / tpc;
    if (!xh && budget / p.b_slot_bytes < 2) continue;
    for (int gpc : {1, 2, 3, 4, 6, 8}) {          // tap groups (= segments) per CTA
      if (gpc > tg) break;
      // candidates: every range count up to 32, a coarser ladder above, and the counts that fill exactly 1 / 2 / 3
      // waves of 148 CTAs
      long long cand[64];
      int ncand = 0;
      for (int s_try = 1; s_try <= 160; s_try += (s_try < 32 ? 1 : (s_try < 64 ? 4 : 16))) cand[ncand++] = s_try;
      for (int k = 1; k <= 3; ++k) {
        const long long fill = 148LL * k * gpc / ((long long)tg * mz);
        if (fill >= 1) cand[ncand++] = fill;
      }
      if (batch) { ncand = 0; cand[ncand++] = batch->ranges; }
      for (int ci = 0; ci < ncand; ++ci) {
        long long S = cand[ci];                    // tile ranges
        if (S > p.total_tiles) continue;
        const long long R = (p.total_tiles + S - 1) / S;
        S = (p.total_tiles + R - 1) / R;
        const long long ncta = (S * tg + gpc - 1) / gpc;
        const long long ctas = ncta * mz;
        if (ctas > 444 && !batch) continue;
        if (batch && S != batch->ranges) continue;  // one tile range per problem
        const long long waves = (ctas + 147) / 148;
        const long long per = (long long)gpc * R;
        // CTAs that walk the same tiles at the same time: L2 serves their re-reads (good), but past ~2 dozen the
        // same lines are hammered by too many SMs at once (measured on hr_convs.0: 63 sharers 8.3 ms, 21 sharers
        // 5.9 ms, none — group-major order — 6.5 ms with 27 GB of DRAM re-reads)
        const double sharers = (double)(tg + gpc - 1) / gpc;
        const double hot = sharers > 24.0 ? 1.0 + (sharers - 24.0) / 40.0 : 1.0;
        // red.global.add also has a chip-wide rate (a few hundred lanes/clk when coalesced): with many CTAs on a
        // short K loop the sum over a wave, not one CTA's epilogue, is what is waited for
        const double wave_ctas = ctas < 148 ? (double)ctas : 148.0;
        const double atom_chip = wave_ctas * gpc * tpc * 128.0 * p.n_umma / 512.0;
        const double atom_t = gpc * tpc * atom > atom_chip ? gpc * tpc * atom : atom_chip;
        const double sh_load = xh ? (double)sh_alloc * halo_blk(tpc) / 33.0 : tpc * unit_load;
        const double item_load = hot * (sh_load + (double)fix_alloc * blk / 33.0);
        const double item_mma = tpc * unit_mma;
        const double t =
            (double)waves * (per * (item_mma > item_load ? item_mma : item_load) + atom_t + 8000.0);
        if (t < best) { best = t; best_tpc = tpc; best_gpc = gpc; best_range = R; best_xh = xh; }
      }
    }
  }
  // experiment hook: WS_WGRAD_FORCE="taps_per_cta,groups_per_cta,tile_ranges" (read at every launch)
  if (batch) {
    // experiment hook of the batched launches: WS_TRUNK_WGRAD_FORCE="taps_per_cta,groups_per_cta"
    if (const char* f = getenv("WS_TRUNK_WGRAD_FORCE")) {
      int ftpc = 0, fgpc = 0;
      if (sscanf(f, "%d,%d", &ftpc, &fgpc) == 2 && ftpc >= 1 && ftpc <= max_tpc && ftpc <= p.taps && fgpc >= 1) {
        best_tpc = ftpc;
        best_gpc = fgpc;
        best_xh = xh_for(ftpc);
      }
    }
    best_range = p.total_tiles / batch->ranges;
  } else if (const char* f = getenv("WS_WGRAD_FORCE")) {
    int ftpc = 0, fgpc = 0, fs = 0;
    if (sscanf(f, "%d,%d,%d", &ftpc, &fgpc, &fs) == 3 && ftpc >= 1 && ftpc <= max_tpc && fgpc >= 1 && fs >= 1) {
      best_xh = xh_for(ftpc);
      best_tpc = ftpc;
      best_gpc = fgpc;
      best_range = (p.total_tiles + fs - 1) / fs;
    }
  }
  p.taps_per_cta = best_tpc;
  p.tap_groups = (p.taps + p.taps_per_cta - 1) / p.taps_per_cta;
  p.xhalo = best_xh ? 1 : 0;
  if (p.xhalo) {
    p.kx_groups = (g.kx + best_tpc - 1) / best_tpc;
    p.tap_groups = g.ky * g.kz * p.kx_groups;
    p.halo_rows = (p.bx + best_tpc - 1) * p.by * p.bz;
    p.halo_blk_bytes = halo_blk(best_tpc);
    p.b_slot_bytes = sh_alloc * p.halo_blk_bytes;
  }
  p.b_slots = budget / p.b_slot_bytes;
  if (p.b_slots > kMaxBSlots) p.b_slots = kMaxBSlots;
  WS_REQUIRE(p.b_slots >= 2, "wgrad: shared memory budget too small for 2 operand slots");
  p.range_tiles = best_range < 1 ? 1 : best_range;
  const long long ranges = (p.total_tiles + p.range_tiles - 1) / p.range_tiles;
  p.items = ranges * p.tap_groups * p.range_tiles;
  // runs are whole segments: every segment a CTA sees then starts at a valid tile (only range tails are skipped)
  p.items_per_cta = (long long)best_gpc * p.range_tiles;
  const long long ncta = (p.items + p.items_per_cta - 1) / p.items_per_cta;
  WS_REQUIRE(ncta <= 2147483647LL, "wgrad: too many CTAs");
  uint32_t cols = 32;
  while ((int)cols < p.taps_per_cta * p.n_umma) cols <<= 1;
  p.tmem_cols = cols;
  static const int env_niss = getenv("WS_WGRAD_NISS") ? atoi(getenv("WS_WGRAD_NISS")) : 4;
  p.n_iss = p.taps_per_cta < 4 ? p.taps_per_cta : 4;
  if (p.n_iss > env_niss) p.n_iss = env_niss < 1 ? 1 : env_niss;
  if (!p.xhalo)
    while (p.n_iss > 1 && p.b_slots / p.n_iss < 2) --p.n_iss;  // two slots of its private ring per issuer

  // workspace layout: the M index is always the contiguous one, so that the 32 lanes of an epilogue warp (one
  // accumulator row each) hit consecutive floats with every red.global.add — [tap][cout][cin] when M = cin
  // (swap), [tap][cin][cout] when M = cout.  (Lane stride = cin made the 1x1x1 LFF gradient atomic-bound: 39 us.)
  p.tap_stride = (long long)g.cout * g.cin;
  p.m_stride = 1;
  p.n_stride = swap ? g.cin : g.cout;
  if (batch) {
    WS_REQUIRE(p.total_tiles % batch->ranges == 0 && p.range_tiles == p.total_tiles / batch->ranges,
               "wgrad: batched launch needs whole tile ranges (%lld tiles, %d problems)", p.total_tiles, batch->ranges);
    p.range_stride = batch->range_stride;
    p.plain_store = 1;
    p.mz = mz;
  }

  auto make_map = [&](const View& v, int channels, int X, int Y, int Z, bool strided, CUtensorMap* out,
                      int x_halo) -> int {
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.ptr = reinterpret_cast<uintptr_t>(v.ptr);
    k.rank = 5; k.dtype = tf32 ? (WS_F32 | kMapSwizzle128Atom32) : WS_BF16;
    k.dims[0] = (uint64_t)channels; k.dims[1] = (uint64_t)Z; k.dims[2] = (uint64_t)Y; k.dims[3] = (uint64_t)X;
    k.dims[4] = (uint64_t)g.n;
    k.strides[0] = (uint64_t)v.vs * esize;
    k.strides[1] = (uint64_t)v.vs * esize * Z;
    k.strides[2] = (uint64_t)v.vs * esize * Z * Y;
    k.strides[3] = (uint64_t)v.ns * esize;
    if (swapxy) {
      // memory order is (x_mem = Y here, y_mem = X here, z): dimension 2 (our "y") strides over whole y_mem-z planes
      k.strides[1] = (uint64_t)v.vs * esize * Z * X;
      k.strides[2] = (uint64_t)v.vs * esize * Z;
    }
    int ssx = strided ? g.sx : 1, ssy = strided ? g.sy : 1, ssz = strided ? g.sz : 1;
    k.box[0] = (uint32_t)cblk;
    k.box[1] = (uint32_t)((p.bz - 1) * ssz + 1);
    k.box[2] = (uint32_t)((p.by - 1) * ssy + 1);
    k.box[3] = (uint32_t)((p.bx - 1) * ssx + 1 + x_halo);
    k.box[4] = 1;
    k.estr[0] = 1; k.estr[1] = (uint32_t)ssz; k.estr[2] = (uint32_t)ssy; k.estr[3] = (uint32_t)ssx; k.estr[4] = 1;
    return get_tensor_map(k, out);
  };
  CUtensorMap tm_x, tm_dy, tm_halo;
  if (int e = make_map(x, g.cin, g.x, g.y, g.z, true, &tm_x, 0)) return e;
  if (int e = make_map(dy, g.cout, g.xo, g.yo, g.zo, false, &tm_dy, 0)) return e;
  tm_halo = tm_x;
  if (p.xhalo)
    if (int e = make_map(x, g.cin, g.x, g.y, g.z, true, &tm_halo, p.taps_per_cta - 1)) return e;

  size_t smem = 2 * (size_t)p.a_slot_bytes + (size_t)p.b_slots * p.b_slot_bytes + 8 * (7 + 2 * kMaxBSlots) + 1024;
  static std::once_flag* once = new std::once_flag;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(*once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  WS_REQUIRE(attr_err == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  WS_REQUIRE(smem <= 227 * 1024, "wgrad smem %zu too large", smem);
  dim3 grid((unsigned)ncta, 1u, (unsigned)mz);
  if (batch) grid = dim3((unsigned)(ncta * mz), 1u, 1u);
  if (!swap) WS_CHECK_CUDA(launch_pdl(wgrad_tc_kernel, grid, dim3(kThreads), smem, st, 1, tm_dy, tm_x, tm_halo, p, wsp));
  else WS_CHECK_CUDA(launch_pdl(wgrad_tc_kernel, grid, dim3(kThreads), smem, st, 1, tm_x, tm_dy, tm_halo, p, wsp));
  WS_POST_LAUNCH(1);
  return 0;
}

__global__ void wgrad_tc_finalize(const float* __restrict__ wsp, float* __restrict__ dw, int taps, int cin,
                                  int cout, int accumulate, int cin_major) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // wsp [tap][co][ci] (cin_major == 0) or [tap][ci][co] (cin_major == 1) -> dw [co][ci][tap]
  long long total = (long long)taps * cin * cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps);
    long long r = i / taps;  // co*cin + ci
    if (cin_major) r = (r % cin) * cout + r / cin;
    float v = wsp[(long long)tap * cin * cout + r];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

int tc_conv_wgrad(const ConvGeom& g, const View& x, const View& dy, float* dw, int accumulate, void* workspace,
                  size_t workspace_bytes, cudaStream_t st) {
  size_t need = tc_wgrad_workspace_bytes(g);
  WS_REQUIRE(workspace && workspace_bytes >= need, "wgrad workspace too small: %zu < %zu", workspace_bytes, need);
  float* wsp = (float*)workspace;
  WS_CHECK_CUDA(cudaMemsetAsync(wsp, 0, need, st));
  // Role choice: put the narrower channel count on N when it is much smaller than 128.
  const bool swap = g.cout < 64 && g.cin >= g.cout;
  const int mtot = swap ? g.cin : g.cout;
  const int ntot = swap ? g.cout : g.cin;
  for (int n0 = 0; n0 < ntot; n0 += 256) {
    int nc = ntot - n0 < 256 ? ntot - n0 : 256;
    if (int e = launch_one(g, x, dy, wsp, swap ? 1 : 0, mtot, n0, nc, st)) return e;
  }
  long long total = (long long)g.taps() * g.cin * g.cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  WS_CHECK_CUDA(launch_pdl(wgrad_tc_finalize, dim3(blocks), dim3(256), 0, st, 1, wsp, dw, g.taps(), g.cin, g.cout,
                           accumulate, swap ? 0 : 1));
  WS_POST_LAUNCH(1);
  return 0;
}


// ---- merged weight gradient of the dense convs of one residual dense block ---------------------------------------
// The k dense convs of an RDB (torch_blocks.py:256-267) all read a prefix of the same concat buffer and each
// produces gc (=32) channels, so their weight gradients are one GEMM: M = buffer channels [0, cin_max), N = the
// k*gc gradient channels side by side, K = voxels, per tap.  N = 160 instead of five launches at N = 32 (the MMA
// costs max(72, N/2) cycles either way).  The (ci >= cin_i) corner of conv i is computed and dropped.
namespace {
struct RdbSlices {
  int nconv;
  int cin[8];
  float* dw[8];
};
__global__ void wgrad_tc_finalize_rdb(const float* __restrict__ wsp, const RdbSlices t, int taps, int gc, int wcin,
                                      int wcout) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // wsp [tap][i*gc + co][ci] -> dw_i [co][ci][tap]
  const int i = blockIdx.y;
  if (i >= t.nconv || !t.dw[i]) return;
  const int cin = t.cin[i];
  const long long total = (long long)taps * cin * gc;
  float* __restrict__ dw = t.dw[i];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(e % taps);
    const long long r = e / taps;
    const int ci = (int)(r % cin), co = (int)(r / cin);
    dw[e] = wsp[((long long)tap * wcout + (i * gc + co)) * wcin + ci];
  }
}
}  // namespace

size_t tc_rdb_wgrad_workspace_bytes(int taps, int cin_max, int nconv, int gc) {
  return (size_t)taps * cin_max * nconv * gc * sizeof(float);
}

// gm: pseudo conv (cin = cin_max, cout = nconv*gc); x = concat buffer, g = the nconv*gc gradient channels.
int tc_rdb_wgrad(const ConvGeom& gm, const View& x, const View& g, float* const* dw, const int* cin, int nconv,
                 int gc, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const size_t need = tc_rdb_wgrad_workspace_bytes(gm.taps(), gm.cin, nconv, gc);
  WS_REQUIRE(workspace && workspace_bytes >= need, "rdb wgrad workspace too small: %zu < %zu", workspace_bytes, need);
  WS_REQUIRE(nconv <= 8 && gm.cout == nconv * gc && gm.cout <= 256, "rdb wgrad: bad channel layout");
  float* wsp = (float*)workspace;
  WS_CHECK_CUDA(cudaMemsetAsync(wsp, 0, need, st));
  if (int e = launch_one(gm, x, g, wsp, 1, gm.cin, 0, gm.cout, st)) return e;
  RdbSlices t;
  memset(&t, 0, sizeof(t));
  t.nconv = nconv;
  long long most = 0;
  for (int i = 0; i < nconv; ++i) {
    t.cin[i] = cin[i];
    t.dw[i] = dw[i];
    const long long e = (long long)gm.taps() * cin[i] * gc;
    if (e > most) most = e;
  }
  int blocks = (int)((most + 255) / 256);
  if (blocks > 148 * 2) blocks = 148 * 2;
  WS_CHECK_CUDA(launch_pdl(wgrad_tc_finalize_rdb, dim3((unsigned)blocks, (unsigned)nconv), dim3(256), 0, st, 1, wsp, t,
                           gm.taps(), gc, gm.cin, gm.cout));
  WS_POST_LAUNCH(1);
  return 0;
}


// ---- weight gradients of a whole run of identical residual dense blocks as batched launches ----------------------
// Per block the voxel count (20 480 at the shipped size) is too small for a split-K GEMM to amortise its reduce
// epilogue: 48 x (merged dense GEMM 63 us + LFF GEMM 61 us + 2 finalize + bias sums) was 7 ms of a 45 ms step.  The
// blocks' operands (concat buffers, g, g_lff) are kept in slabs with the block index outermost, so the batch dimension
// of the tensor maps runs over (block, sample) and ONE launch does every block: a CTA owns (block, tap group, M block)
// with the full voxel range as its K loop — no split-K, no atomics, plain stores.
namespace {
struct TrunkLayout {
  int nconv, gc, taps, wcin, wcout;
  int cin[8];
  long long off[8];        // offset of dw_i inside a block's gradient record (floats)
  long long block_stride;  // floats between the gradient records of consecutive blocks
  long long ws_stride;     // floats between the workspace results of consecutive blocks
};
// one CTA per (co, conv, block): the 27 taps' rows [ci] are contiguous in the workspace, the output row
// dw[co][ci][tap] is contiguous too — transpose through shared memory
__global__ void __launch_bounds__(256)
wgrad_tc_finalize_trunk(const float* __restrict__ wsp, float* __restrict__ grads, const TrunkLayout t) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ float tile[];  // [taps][cin + 1]
  const int co = blockIdx.x, i = blockIdx.y, r = blockIdx.z;
  const int cin = t.cin[i], pitch = cin + 1;
  const float* src = wsp + (long long)r * t.ws_stride;
  for (int e = threadIdx.x; e < t.taps * cin; e += blockDim.x) {
    const int tap = e / cin, ci = e % cin;
    tile[tap * pitch + ci] = src[((long long)tap * t.wcout + (i * t.gc + co)) * t.wcin + ci];
  }
  __syncthreads();
  float* dst = grads + (long long)r * t.block_stride + t.off[i] + (long long)co * cin * t.taps;
  for (int e = threadIdx.x; e < t.taps * cin; e += blockDim.x) {
    const int ci = e / t.taps, tap = e % t.taps;
    dst[e] = tile[tap * pitch + ci];
  }
}
}  // namespace

size_t tc_trunk_wgrad_workspace_bytes(int nblocks, int taps, int cin_max, int nconv, int gc) {
  return (size_t)nblocks * tc_rdb_wgrad_workspace_bytes(taps, cin_max, nconv, gc);
}

// gm / gl: the merged dense pseudo conv and the LFF conv of ONE block, with n = nblocks * samples.  buf / g / g_lff:
// views of block 0; block r starts n_samples * nstride elements later.  grads: per block a record
// [dw_0 | ... | dw_{nconv-1} | dw_lff | db_lff] (torch layouts), block_stride floats apart.
int tc_trunk_wgrad(const ConvGeom& gm, const ConvGeom& gl, int nblocks, const View& buf, const View& g,
                   const View& g_lff, const int* cin, int nconv, int gc, float* grads, long long block_stride,
                   void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const size_t need = tc_trunk_wgrad_workspace_bytes(nblocks, gm.taps(), gm.cin, nconv, gc);
  WS_REQUIRE(workspace && workspace_bytes >= need, "trunk wgrad workspace too small: %zu < %zu", workspace_bytes, need);
  WS_REQUIRE(nconv >= 1 && nconv <= 8 && gm.cout == nconv * gc && gm.cout <= 256, "trunk wgrad: bad channel layout");
  WS_REQUIRE(gl.taps() == 1, "trunk wgrad: the batched LFF gradient is the 1x1x1 case");
  TrunkLayout t;
  memset(&t, 0, sizeof(t));
  t.nconv = nconv; t.gc = gc; t.taps = gm.taps(); t.wcin = gm.cin; t.wcout = gm.cout;
  long long off = 0;
  int cmax = 0;
  for (int i = 0; i < nconv; ++i) {
    t.cin[i] = cin[i];
    t.off[i] = off;
    off += (long long)gc * cin[i] * gm.taps();
    if (cin[i] > cmax) cmax = cin[i];
  }
  const long long off_lff = off;
  t.block_stride = block_stride;
  t.ws_stride = (long long)gm.taps() * gm.cin * gm.cout;
  float* wsp = (float*)workspace;
  // dense convs: M = concat-buffer channels (shifted operand), N = the nconv*gc gradient channels
  WgBatch bd = {nblocks, t.ws_stride};
  if (int e = launch_one(gm, buf, g, wsp, 1, gm.cin, 0, gm.cout, st, &bd)) return e;
  // LFF (1x1x1): with M = cin the M-contiguous result [co][ci] IS torch's layout — straight into the record
  WgBatch bl = {nblocks, block_stride};
  for (int n0 = 0; n0 < gl.cout; n0 += 256) {
    const int nc = gl.cout - n0 < 256 ? gl.cout - n0 : 256;
    if (int e = launch_one(gl, buf, g_lff, grads + off_lff, 1, gl.cin, n0, nc, st, &bl)) return e;
  }
  const size_t smem = (size_t)gm.taps() * (cmax + 1) * sizeof(float);
  WS_REQUIRE(smem <= 48 * 1024, "trunk wgrad finalize: tile too large");
  WS_CHECK_CUDA(launch_pdl(wgrad_tc_finalize_trunk, dim3((unsigned)gc, (unsigned)nconv, (unsigned)nblocks), dim3(256),
                           smem, st, 1, (const float*)wsp, grads, t));
  WS_POST_LAUNCH(1);
  return 0;
}

}  // namespace ws
