// conv_tc2.cu — second-generation tcgen05 implicit-GEMM Conv3d (stride 1, forward and data-gradient) built
// around a TMA-staged **halo tile** of the voxel grid.
//
// Round-1 profiling showed conv_tc.cu (one TMA box per tap) to be bound by the L2 -> SMEM fabric: hr_convs.0
// moved 130 GB per launch at 9.7 TB/s for 36 % tensor utilisation.  Here a CTA owns `tx` consecutive x-slabs of
// by*Z voxels (rows ordered x slowest, then y, then z) and, per (ky, kz, 64-channel chunk), loads ONE activation
// box {64 ch, Z, by, tx+kx-1, 1} — the output slabs plus the x halo — with TMA zero fill supplying the conv
// padding.  Inside that box the kx taps along x are just UMMA descriptor start offsets of ti*by*Z rows
// (by*Z % 8 == 0, so SWIZZLE_128B 8-row atoms stay aligned and need no base offset).  The CTA keeps up to four
// 128-row accumulators in TMEM (t_m * N <= 512 columns), so every weight tile fetched serves all of them.
// Per MMA this moves ~3x fewer bytes than v1 (hr_convs.0: 158 KB per 60 MMAs instead of 34 KB per 4).
//
// Pipeline: warp 0 = TMA producer (2 halo buffers + a ring of weight slots), warp 1 = MMA issuer / TMEM owner,
// warps 2-5 = epilogue (same fused epilogue as v1, common.cuh).
//
// Replaces nn.Conv3d forward + dgrad for the stride-1 layers: torch_blocks.py:17 (RDB / trunk / UpConv convs),
// Generator_3D_Resnet_ESRGAN.py:95-111 (hr_convs).
#include <cuda.h>
#include <mutex>
#include <vector>
#include <array>
#include <map>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "tc_task.cuh"

namespace ws {

namespace {

constexpr int kMaxWSlots = 6;
constexpr int kMaxABufs = 6;

struct Tc2Params {
  int N, DX, DY, DZ;
  int by, bz, tx, slabrows;
  int tiles_x, tiles_y;
  int kx, ky, kz, px, py, pz;
  int ck, kchunks, last_k16, cn, n_umma, n_tile;
  int t_m, out_rows, halo_rows;
  int a_buf_bytes, w_bytes, w_slots;
  int w_group;       // weight tiles (taps) per ring slot: the slot's barrier pair is waited / committed once per group.
                     // A completed mbarrier wait + tcgen05 fence costs the issuing thread ~280 cycles and a tcgen05.commit
                     // ~220 (scripts/micro/mma_rate3.cu) — per TAP that is more than the tap's MMAs (t_m x 288 cycles)
  int merged;        // 1: w_group == taps per halo load and the weight tiles ride on the halo buffer's barriers
                     // (ONE wait + ONE commit per K iteration)
  int a_bufs;        // halo buffers in the ring (2..kMaxABufs)
  int a_sub_slabs;   // x-slabs per TMA instruction of the halo load
  int a_ops;         // TMA instructions per halo load
  int vol;           // 1: volume mode — ONE padded halo box per 64-channel chunk, all kx*ky*kz taps = row offsets
  int pitch_y;       // rows between consecutive y lines of the tile (DZ, or DZ + kz - 1 in volume mode)
  int box_y, box_z;  // TMA box extents in y and z (by, DZ, plus the ky-1 / kz-1 halo in volume mode)
  int row_bytes;     // bytes of one operand row: 128 (SWIZZLE_128B, 64 channels per chunk) or 64 (SWIZZLE_64B, 32 channels:
                     // K <= 32 layers — the dense-conv data-gradients — issue 2 K-steps per tap instead of 4 half-empty ones)
  int kelems;        // channels per K chunk = row_bytes / element size
  int tf32;          // 1: fp32 activations and weights, tcgen05 kind::tf32 (K = 8 per instruction instead of 16)
  int tap_base;      // first tap of the packed weights this launch uses
  int omx, oax, omy, oay, omz, oaz, ODY, ODZ;  // destination voxel transform (tc_task.cuh)
  uint32_t tmem_cols;
  int n_iss;         // MMA-issuing warps (1..4, <= t_m): warp 1 + q owns the accumulators m = q, q + n_iss, ...  A tcgen05.mma
                     // occupies its issuing thread for about as long as it executes (scripts/micro/mma_rate3.cu: a loop
                     // with +150 cycles of scalar work per 4 MMAs ran 110 instead of 73 cycles per MMA): whatever the
                     // thread does between two MMAs — descriptor arithmetic, elect / reconvergence, barrier waits, commits —
                     // is tensor-pipe idle time unless ANOTHER warp has MMAs queued meanwhile
  int lean;          // 1: the streamlined issue loop (WS_TC2_LEAN=0 keeps the general one)
  int chunk_major;  // 1: K loop ordered (chunk, ky, kz) instead of (ky, kz, chunk): the short tail chunk of a channel count
                    // that is not a multiple of 64 (144 = 64 + 64 + 16) runs as one block at the end — its quarter-length
                    // MMA bursts no longer have to hide the load of a full-size halo tile
  long long* dbg;  // optional (WS_TC2_DEBUG_TIMES=1): per-CTA cycle counts of the pipeline stages, 8 slots per CTA
};

// kPair: launched as clusters of two CTAs (cta_group::2).  Each CTA keeps its own halo tile, accumulators and
// epilogue, but only HALF of every weight tile; the leader issues one M=256 MMA per step that multiplies both CTAs'
// A tiles with the pair's combined B tile.  Per SM this halves the weight traffic from L2 and the B-operand reads
// from shared memory (measured: a cta_group::1 N=144 MMA already saturates the 128 B/clk SMEM port).
// kEpiWarps: 4 or 8 epilogue warps.  A warp can only read the TMEM lane quarter (warp % 4); with 8 warps two of them
// share a quarter and take alternate 16-column chunks — twice the loads / stores in flight for the wide fp32
// accumulate epilogues, used when the CTA has the SM to itself anyway (ncu: the dense-conv dgrad spent ~60 % of its
// cycles in a latency-bound 4-warp epilogue).
template <bool kPair, int kEpiWarps>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, (kPair || kEpiWarps == 8) ? 1 : 2)  // 2 CTAs/SM: <= 168 regs
conv3d_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const Tc2Params p, const View dst, const Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float stat_s[512];  // per-CTA BatchNorm partial sums (train-mode discriminator convs)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t w_base = smem_base + (uint32_t)p.a_bufs * p.a_buf_bytes;
  if (ep.stat_sum)
    for (int i = threadIdx.x; i < 512; i += blockDim.x) stat_s[i] = 0.f;  // visible after the prologue barrier
  const uint32_t w_slot_bytes = (uint32_t)p.w_group * (uint32_t)p.w_bytes;
  const uint32_t bar_off = (uint32_t)p.a_bufs * p.a_buf_bytes + (uint32_t)p.w_slots * w_slot_bytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kMaxABufs + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * kMaxABufs + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * kMaxABufs + kMaxWSlots + s); };
  constexpr int kNumBars = 2 * kMaxABufs + 2 * kMaxWSlots;
  const uint32_t accum_bar = bar_base + 8u * kNumBars;
  const uint32_t tmem_slot = bar_base + 8u * (kNumBars + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * (kNumBars + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  int t = blockIdx.x;
  // a padding CTA (odd tile count in pair mode) works on an out-of-range tile: zero-filled loads, no stores
  const bool tile_ok = t < p.N * p.tiles_x * p.tiles_y;
  const int tyi = t % p.tiles_y; t /= p.tiles_y;
  const int txi = t % p.tiles_x; t /= p.tiles_x;
  const int n = tile_ok ? t : p.N - 1;
  const int x0 = tile_ok ? txi * p.tx : p.DX + p.kx, y0 = tyi * p.by;
  const int n0 = blockIdx.y * p.n_tile;

  // programmatic dependent launch: let the next kernel of the stream start its prologue as soon as every CTA of
  // this grid is running; this kernel's own global-memory traffic starts after griddep_wait() below
  if (threadIdx.x == 0) ptx::griddep_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    // every issuing warp commits its own MMAs: the "empty" and "accumulators complete" barriers take n_iss arrivals
    for (int s = 0; s < p.a_bufs; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), (uint32_t)p.n_iss); }
    if (!p.merged)
      for (int s = 0; s < p.w_slots; ++s) { ptx::mbar_init(w_full(s), 1); ptx::mbar_init(w_empty(s), (uint32_t)p.n_iss); }
    ptx::mbar_init(accum_bar, (uint32_t)p.n_iss);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) { ptx::tmem_alloc2(tmem_slot, p.tmem_cols); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(tmem_slot, p.tmem_cols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();  // the peer's TMA and the leader's commits target barriers of the other CTA
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int nyz = p.ky * p.kz;
  // pair mode: per-CTA weight slice = rows [rank * n_umma/2, +n_umma/2) of the tile; the leader's full barriers
  // count the bytes of both CTAs
  const uint32_t w_cta_bytes = (uint32_t)p.w_bytes;  // bytes this CTA loads per weight tile (half a tile in pair mode)
  const int w_row0 = n0 + (kPair ? (int)rank * (p.n_umma / 2) : 0);
  const uint32_t tx_mult = kPair ? 2u : 1u;

  // Producer and MMA warps run their loops warp-uniformly (all 32 lanes) and elect one lane only around the
  // UTMALDG / UTCHMMA instructions: operands then live in uniform registers.  (Round-1 ncu finding: issuing from
  // inside an `if (lane == 0)` region made ptxas wrap every tcgen05.mma in an R2UR/ELECT waterfall loop and the
  // single issuing thread, not the tensor pipe, bounded the kernel.)
  if (warp == 0) {
    // ===== TMA producer =====
    ptx::griddep_wait();  // the predecessor's activations / packed weights must be complete before the first load
    int ab = 0, wsl = 0;
    uint32_t aph = 0, wph = 0;
    long long dbg_wa = 0, dbg_ww = 0;
    const uint32_t a_op_bytes = (uint32_t)(p.a_sub_slabs * p.slabrows) * (uint32_t)p.row_bytes;
    const int ngroups = p.vol ? 1 : nyz;             // activation loads per 64-channel chunk
    const int ntaps = p.vol ? p.kx * nyz : p.kx;     // weight tiles per activation load
    for (int it = 0; it < ngroups * p.kchunks; ++it) {
      const int yz = p.chunk_major ? it % ngroups : it / p.kchunks;
      const int ch = p.chunk_major ? it / ngroups : it % p.kchunks;
      const int tj = p.vol ? 0 : yz / p.kz, tl = p.vol ? 0 : yz % p.kz;
      {
        if (p.dbg) { const long long c0 = clock64(); ptx::mbar_wait(a_empty(ab), aph ^ 1u); dbg_wa += clock64() - c0; }
        else ptx::mbar_wait(a_empty(ab), aph ^ 1u);
        const uint32_t a_tx = a_op_bytes * (uint32_t)p.a_ops;
        if (ptx::elect_one()) {
          // the halo box is fetched as a_ops independent TMA instructions (more requests in flight)
          if (leader) ptx::mbar_expect_tx(a_full(ab), (a_tx + (p.merged ? (uint32_t)ntaps * w_cta_bytes : 0u)) * tx_mult);
          for (int o = 0; o < p.a_ops; ++o) {
            const uint32_t d = smem_base + ab * p.a_buf_bytes + o * a_op_bytes;
            if (kPair)
              ptx::tma_load_5d_2sm(d, &tmA, a_full(ab), ch * p.kelems, tl - p.pz, y0 - p.py + tj,
                                   x0 - p.px + o * p.a_sub_slabs, n);
            else
              ptx::tma_load_5d(d, &tmA, a_full(ab), ch * p.kelems, tl - p.pz, y0 - p.py + tj,
                               x0 - p.px + o * p.a_sub_slabs, n);
          }
        }
        __syncwarp();
        for (int t0 = 0; t0 < ntaps; t0 += p.w_group) {
          const int cnt = min(p.w_group, ntaps - t0);
          // merged: the weight tiles of this K iteration live in slot `ab` and complete the halo buffer's barrier
          const int slot = p.merged ? ab : wsl;
          const uint32_t bar = p.merged ? a_full(ab) : w_full(wsl);
          if (!p.merged) {
            if (p.dbg) { const long long c0 = clock64(); ptx::mbar_wait(w_empty(wsl), wph ^ 1u); dbg_ww += clock64() - c0; }
            else ptx::mbar_wait(w_empty(wsl), wph ^ 1u);
          }
          if (ptx::elect_one()) {
            if (leader && !p.merged) ptx::mbar_expect_tx(bar, (uint32_t)cnt * w_cta_bytes * tx_mult);
            for (int j = 0; j < cnt; ++j) {
              // volume mode walks the packed taps in storage order (ti, tj, tl); otherwise the kx taps of this (tj, tl)
              const int tt = t0 + j;
              const int tap = p.tap_base + (p.vol ? tt : (tt * p.ky + tj) * p.kz + tl);
              const uint32_t d = w_base + slot * w_slot_bytes + j * p.w_bytes;
              if (kPair) ptx::tma_load_3d_2sm(d, &tmB, bar, ch * p.kelems, w_row0, tap);
              else ptx::tma_load_3d(d, &tmB, bar, ch * p.kelems, w_row0, tap);
            }
          }
          __syncwarp();
          if (!p.merged && ++wsl == p.w_slots) { wsl = 0; wph ^= 1u; }
        }
        if (++ab == p.a_bufs) { ab = 0; aph ^= 1u; }
      }
    }
    if (p.dbg && lane == 0) { p.dbg[8 * blockIdx.x + 5] = dbg_wa; p.dbg[8 * blockIdx.x + 6] = dbg_ww; }
  }
  if (warp >= 1 && warp - 1 < p.n_iss && leader) {
    // ===== MMA issuers (pair mode: the leader CTA only): warps 1 .. n_iss; warps >= 2 go on to the epilogue =====
    const int q = warp - 1;
    const uint32_t idesc = ptx::make_idesc(p.tf32 ? 2u : 1u, kPair ? 256u : 128u, (uint32_t)p.n_umma, 0u, 0u);
    const int k_per_row = p.row_bytes / 32;  // K steps per operand row: one instruction consumes 32 bytes of K
    // everything but the start address: SWIZZLE_128B (SBO = 8 rows x 128 B) or SWIZZLE_64B (layout type 4, SBO = 512;
    // scripts/micro/sw64.cu)
    const uint64_t desc_hi = p.row_bytes == 128
                                 ? ptx::make_smem_desc_sw128(0, 16, 1024)
                                 : (((uint64_t)1 << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) |
                                    ((uint64_t)4 << 61));
    int ab = 0, wsl = 0;
    uint32_t aph = 0, wph = 0;
    const int ntaps = p.vol ? p.kx * nyz : p.kx;
    const int iters = (p.vol ? 1 : nyz) * p.kchunks;
    long long dbg_wa = 0, dbg_ww = 0;
    const bool dbg = p.dbg && q == 0;
    const long long dbg_t0 = dbg ? clock64() : 0;
    if (!p.vol && !p.tf32 && !p.dbg && p.lean) {
      // ---- lean issue loop (the common case).  The issuing warp is a single latency chain: the ncu source page of the
      // general loop below showed ~90 dependent instructions per tap, a dozen of them re-loads of kernel parameters
      // from the constant bank, against 4 MMAs.  Here every loop invariant is pinned in a register (keep(): the
      // compiler cannot rematerialise it from the constant bank) and the descriptors advance by additions.
      const int t_m = ptx::keep(p.t_m), n_iss = ptx::keep(p.n_iss), w_group = ptx::keep(p.w_group), merged = ptx::keep(p.merged);
      const int a_bufs = ptx::keep(p.a_bufs), w_slots = ptx::keep(p.w_slots), kchunks = ptx::keep(p.kchunks);
      const int kfull = ptx::keep(k_per_row), klast = ptx::keep(p.last_k16), cmaj = ptx::keep(p.chunk_major), nyz_k = ptx::keep(nyz);
      const uint32_t a_buf_bytes = (uint32_t)ptx::keep(p.a_buf_bytes), w_bytes16 = (uint32_t)ptx::keep(p.w_bytes >> 4);
      const uint32_t slab16 = (uint32_t)ptx::keep((p.slabrows * p.row_bytes) >> 4);   // descriptor units (16 B) per kx tap
      const uint32_t tile16 = (uint32_t)ptx::keep((128 * p.row_bytes) >> 4);          // ... per accumulator tile
      const uint32_t acc_cols = (uint32_t)ptx::keep(p.n_umma);
      const uint32_t hi32 = (uint32_t)(desc_hi >> 32), lo_const = (uint32_t)desc_hi;
      const uint32_t wslot16 = (uint32_t)ptx::keep((int)(w_slot_bytes >> 4)), wbase16 = (w_base >> 4);
      for (int it = 0; it < iters; ++it) {
        const int ch = cmaj ? it / nyz_k : it % kchunks;
        const int nk = (ch == kchunks - 1) ? klast : kfull;
        ptx::mbar_wait(a_full(ab), aph);
        ptx::tc_fence_after();
        const uint32_t a16 = (smem_base + (uint32_t)ab * a_buf_bytes) >> 4;
        uint32_t a_tap16 = a16 + (uint32_t)q * tile16;  // this issuer's first accumulator tile, tap 0
        int gpos = 0;
        uint32_t b16 = wbase16 + (uint32_t)(merged ? ab : wsl) * wslot16;
        for (int tt = 0; tt < ntaps; ++tt) {
          if (!merged && gpos == 0) {
            ptx::mbar_wait(w_full(wsl), wph);
            ptx::tc_fence_after();
            b16 = wbase16 + (uint32_t)wsl * wslot16;
          }
          const uint32_t acc0 = (it > 0 || tt > 0) ? 1u : 0u;
          const uint32_t b_lo = lo_const | (b16 & 0x3fffu);
          uint32_t am16 = a_tap16;
          for (int m = q; m < t_m; m += n_iss) {
            const uint32_t a_lo = lo_const | (am16 & 0x3fffu);
            const uint32_t d_tmem = tmem_base + (uint32_t)m * acc_cols;
            if (ptx::elect_one()) {
              if (kPair) {
                ptx::mma_f16_ss2_lohi(d_tmem, a_lo, b_lo, hi32, idesc, acc0);
                if (nk > 1) ptx::mma_f16_ss2_lohi(d_tmem, a_lo + 2, b_lo + 2, hi32, idesc, 1u);
                if (nk > 2) ptx::mma_f16_ss2_lohi(d_tmem, a_lo + 4, b_lo + 4, hi32, idesc, 1u);
                if (nk > 3) ptx::mma_f16_ss2_lohi(d_tmem, a_lo + 6, b_lo + 6, hi32, idesc, 1u);
              } else {
                ptx::mma_f16_ss_lohi(d_tmem, a_lo, b_lo, hi32, idesc, acc0);
                if (nk > 1) ptx::mma_f16_ss_lohi(d_tmem, a_lo + 2, b_lo + 2, hi32, idesc, 1u);
                if (nk > 2) ptx::mma_f16_ss_lohi(d_tmem, a_lo + 4, b_lo + 4, hi32, idesc, 1u);
                if (nk > 3) ptx::mma_f16_ss_lohi(d_tmem, a_lo + 6, b_lo + 6, hi32, idesc, 1u);
              }
            }
            __syncwarp();
            am16 += (uint32_t)n_iss * tile16;
          }
          a_tap16 += slab16;
          b16 += w_bytes16;
          if (++gpos == w_group) gpos = 0;
          if (!merged && (gpos == 0 || tt == ntaps - 1)) {
            gpos = 0;
            if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(w_empty(wsl)); else ptx::mma_commit(w_empty(wsl)); }
            __syncwarp();
            if (++wsl == w_slots) { wsl = 0; wph ^= 1u; }
          }
        }
        if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(a_empty(ab)); else ptx::mma_commit(a_empty(ab)); }
        __syncwarp();
        if (++ab == a_bufs) { ab = 0; aph ^= 1u; }
      }
    } else
    for (int it = 0; it < iters; ++it) {
      const int ch = p.chunk_major ? it / (p.vol ? 1 : nyz) : it % p.kchunks;
      const int nk = (ch == p.kchunks - 1) ? p.last_k16 : k_per_row;
      if (dbg) { const long long c0 = clock64(); ptx::mbar_wait(a_full(ab), aph); dbg_wa += clock64() - c0; }
      else ptx::mbar_wait(a_full(ab), aph);
      const uint32_t a_addr = smem_base + ab * p.a_buf_bytes;
      if (p.merged) ptx::tc_fence_after();
      int gpos = 0;  // position inside the weight group
      for (int tt = 0; tt < ntaps; ++tt) {
        const int tj_ = gpos;
        const int slot = p.merged ? ab : wsl;
        if (!p.merged && tj_ == 0) {
          if (dbg) { const long long c0 = clock64(); ptx::mbar_wait(w_full(wsl), wph); dbg_ww += clock64() - c0; }
          else ptx::mbar_wait(w_full(wsl), wph);
          ptx::tc_fence_after();
        }
        const uint64_t bdesc = desc_hi | (uint64_t)(((w_base + slot * w_slot_bytes + tj_ * p.w_bytes) >> 4) & 0x3fffu);
        const uint32_t acc0 = (it > 0 || tt > 0) ? 1u : 0u;
        // tap -> first A row.  Any 128-byte row of the 1024-byte-aligned tile is a valid SWIZZLE_128B operand
        // start with base_offset 0: the swizzle is a function of the absolute address bits
        // (scripts/micro/row_offset.cu), so ky / kz shifts need no 8-row alignment.
        int roff = tt * p.slabrows;
        if (p.vol) {
          const int ti = tt / nyz, tyz = tt - ti * nyz;
          const int tj = tyz / p.kz, tl = tyz - tj * p.kz;
          roff = ti * p.slabrows + tj * p.pitch_y + tl;
        }
        for (int m = q; m < p.t_m; m += p.n_iss) {
          const uint32_t am = a_addr + (uint32_t)(m * 128 + roff) * (uint32_t)p.row_bytes;
          const uint64_t adesc = desc_hi | (uint64_t)((am >> 4) & 0x3fffu);
          const uint32_t d_tmem = tmem_base + (uint32_t)(m * p.n_umma);
          if (ptx::elect_one()) {
            // K advance inside the 128-byte swizzle row: +32 B = +2 in the (addr >> 4) field
            if (p.tf32) {
              if (kPair) {
                ptx::mma_tf32_ss2(d_tmem, adesc, bdesc, idesc, acc0);
                if (nk > 1) ptx::mma_tf32_ss2(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                if (nk > 2) ptx::mma_tf32_ss2(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                if (nk > 3) ptx::mma_tf32_ss2(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
              } else {
                ptx::mma_tf32_ss(d_tmem, adesc, bdesc, idesc, acc0);
                if (nk > 1) ptx::mma_tf32_ss(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                if (nk > 2) ptx::mma_tf32_ss(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                if (nk > 3) ptx::mma_tf32_ss(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
              }
            } else if (kPair) {
              ptx::mma_f16_ss2(d_tmem, adesc, bdesc, idesc, acc0);
              if (nk > 1) ptx::mma_f16_ss2(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              if (nk > 2) ptx::mma_f16_ss2(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              if (nk > 3) ptx::mma_f16_ss2(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
            } else {
              ptx::mma_f16_ss(d_tmem, adesc, bdesc, idesc, acc0);
              if (nk > 1) ptx::mma_f16_ss(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              if (nk > 2) ptx::mma_f16_ss(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              if (nk > 3) ptx::mma_f16_ss(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
            }
          }
          __syncwarp();
        }
        if (++gpos == p.w_group) gpos = 0;
        if (!p.merged && (gpos == 0 || tt == ntaps - 1)) {
          if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(w_empty(wsl)); else ptx::mma_commit(w_empty(wsl)); }
          __syncwarp();
          if (++wsl == p.w_slots) { wsl = 0; wph ^= 1u; }
        }
      }
      if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(a_empty(ab)); else ptx::mma_commit(a_empty(ab)); }
      __syncwarp();
      if (++ab == p.a_bufs) { ab = 0; aph ^= 1u; }
    }
    if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(accum_bar); else ptx::mma_commit(accum_bar); }
    __syncwarp();
    if (dbg && lane == 0) {
      p.dbg[8 * blockIdx.x + 0] = clock64() - dbg_t0;  // issue loop
      p.dbg[8 * blockIdx.x + 1] = dbg_wa;              // ... of which waiting for halo tiles
      p.dbg[8 * blockIdx.x + 2] = dbg_ww;              // ... and for weight tiles
    }
  }
  if (warp >= 2) {
    // ===== epilogue =====
    const int sub = warp & 3;
    constexpr int kParts = kEpiWarps / 4;       // warps sharing one TMEM lane quarter
    const int part = (warp - 2) >> 2;           // this warp takes chunks part, part + kParts, ...
    ptx::griddep_wait();  // residual reads / output writes also order after the predecessor
    const long long dbg_e0 = p.dbg ? clock64() : 0;
    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after();
    const long long dbg_e1 = p.dbg ? clock64() : 0;
    long long dbg_ld = 0;  // cycles inside tcgen05.ld + wait
    const EpiVec ev = make_epi_vec(dst, ep);
    const bool simple = epi_is_simple(ep);
    for (int m = 0; m < p.t_m; ++m) {
      const int r = m * 128 + sub * 32 + lane;  // row inside the CTA's output region
      const int rx = r / p.slabrows;
      const int rem = r - rx * p.slabrows;
      const int ry = rem / p.pitch_y, rz = rem - ry * p.pitch_y;
      const int gx = x0 + rx, gy = y0 + ry;
      const bool row_ok = tile_ok && r < p.out_rows && ry < p.by && rz < p.bz && gx < p.DX && gy < p.DY;
      const long long v = ((long long)(gx * p.omx + p.oax) * p.ODY + (gy * p.omy + p.oay)) * p.ODZ + (rz * p.omz + p.oaz);
      if (simple && ep.res1.ptr) {
        // accumulate-into-residual epilogue (dense-conv dgrads, LFF): the res1 reads of the NEXT 16 channels are in
        // flight while the current 16 are converted and stored, so the L2 round trips overlap instead of
        // serialising one per chunk (the dgrad chain was bound by exactly that latency).
        const int climit = p.cn - n0 < p.n_umma ? p.cn - n0 : p.n_umma;  // columns of this tile that exist
        const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(m * p.n_umma);
        float cur[16], nxt[16];
        prefetch_res16(ep, ev, n, v, n0 + 16 * part, p.cn, n0 + climit, row_ok, cur);
        for (int c0 = 16 * part; c0 < climit; c0 += 16 * kParts) {
          uint32_t rr[16];
          ptx::tmem_ld16(t_row + (uint32_t)c0, rr);
          prefetch_res16(ep, ev, n, v, n0 + c0 + 16 * kParts, p.cn, n0 + climit, row_ok, nxt);
          ptx::tmem_ld_wait();
          epilogue16_simple(ep, ev, dst, n, v, n0 + c0, p.cn, row_ok, rr, cur);
#pragma unroll
          for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
        }
      } else {
        for (int c0 = 16 * part; c0 < p.n_umma; c0 += 16 * kParts) {
          if (n0 + c0 >= p.cn) break;
          uint32_t rr[16];
          const long long dbg_c0 = p.dbg ? clock64() : 0;
          ptx::tmem_ld16(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(m * p.n_umma + c0), rr);
          ptx::tmem_ld_wait();
          if (p.dbg) dbg_ld += clock64() - dbg_c0;
          if (simple) epilogue16_simple(ep, ev, dst, n, v, n0 + c0, p.cn, row_ok, rr);
          else epilogue16(ep, ev, dst, n, v, n0 + c0, p.cn, row_ok, rr, lane, ep.stat_sum ? stat_s : nullptr, n0);
        }
      }
    }
    if (ep.stat_sum) {
      // all epilogue warps have added their partial sums: one global atomic per channel for the whole CTA
      asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
      flush_bn_stats(ep, stat_s, n0, p.n_umma, p.cn, (int)threadIdx.x - 64, 32 * kEpiWarps);
    }
    ptx::tc_fence_before();
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[8 * blockIdx.x + 3] = dbg_e1 - dbg_e0;    // prologue end -> accumulators complete
      p.dbg[8 * blockIdx.x + 4] = clock64() - dbg_e1;  // epilogue
      p.dbg[8 * blockIdx.x + 7] = dbg_ld;              // ... of which TMEM loads
    }
  }

  if (kPair) ptx::cluster_sync();
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc2(tmem_base, p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

constexpr int kSmemBudget = 222 * 1024;  // dynamic SMEM; 2 KB of static SMEM (BN partial sums) sit beside it

// How many weight tiles share one barrier pair (Tc2Params::w_group / merged) under an SMEM budget: all `ntaps` of a K
// iteration on the halo buffer's own barriers when two such stages fit, else the largest even split that leaves room
// for `need_bufs` halo buffers and two weight-group slots.  Returns the barrier round trips per K iteration.
int weight_grouping(int a_buf_bytes, int w_bytes, int ntaps, int budget, int need_bufs, int& w_group, int& merged) {
  static const bool no_merge = getenv("WS_TC2_NO_MERGE") != nullptr;
  static const int env_group = getenv("WS_TC2_WGROUP") ? atoi(getenv("WS_TC2_WGROUP")) : 0;
  merged = 0;
  if (!no_merge && env_group <= 0 && need_bufs * (a_buf_bytes + ntaps * w_bytes) + 2048 <= budget) {
    merged = 1;
    w_group = ntaps;
    return 1;
  }
  int g = (budget - 2048 - need_bufs * a_buf_bytes) / (2 * w_bytes);
  if (g > ntaps) g = ntaps;
  if (g < 1) g = 1;
  if (env_group > 0 && env_group < g) g = env_group;
  const int ngr = (ntaps + g - 1) / g;
  w_group = (ntaps + ngr - 1) / ngr;
  return 1 + ngr;
}

// Pick (by, tx): minimise estimated SM-time = waves * max(MMA cycles, L2->SMEM cycles) per K iteration.
bool choose_cfg_search(int N, int DX, int DY, int DZ, int kx, int n_umma, bool pair, Tc2Params& p) {
  const int tmax = 512 / n_umma < 4 ? 512 / n_umma : 4;
  if (tmax < 1 || DZ > 128) return false;
  double best = 1e30;
  bool found = false;
  // experiment hook: WS_TC2_FORCE="by,tx" pins the tile shape
  static int force_by = 0, force_tx = 0;
  static const bool forced = [] {
    const char* v = getenv("WS_TC2_FORCE");
    return v && sscanf(v, "%d,%d", &force_by, &force_tx) == 2;
  }();
  for (int by = 1; by <= DY && by * DZ <= 128 * tmax; ++by) {
    const int slab = by * DZ;
    // (any slab size works: a SWIZZLE_128B operand may start at any 128-byte row, scripts/micro/row_offset.cu)
    if (slab % 8 && !getenv("WS_TC2_ANY_SLAB")) continue;
    if (forced && by != force_by) continue;
    for (int tx = 1; tx <= DX && tx * slab <= 128 * tmax; ++tx) {
      if (forced && tx != force_tx) continue;
      if (tx + kx - 1 > 256) break;
      const int t_m = (tx * slab + 127) / 128;
      const int halo = (tx + kx - 1) * slab;
      int a_rows = t_m * 128 + (kx - 1) * slab;
      // the halo is fetched in up to 4 TMA instructions of `sub` slabs each: the last one may overrun the halo
      for (int want = 1; want <= 4; ++want) {
        const int sub = (tx + kx - 1 + want - 1) / want;
        const int ops_rows = ((tx + kx - 1 + sub - 1) / sub) * sub * slab;
        if (a_rows < ops_rows) a_rows = ops_rows;
      }
      const int a_bytes = (a_rows * 128 + 1023) / 1024 * 1024;
      const int w_bytes = n_umma * 128 / (pair ? 2 : 1);  // per-CTA share of a weight tile
      if (2 * a_bytes + 2 * w_bytes + 2048 > kSmemBudget) continue;
      // measured on B200 (scripts/micro/mma_rate.cu): a cta_group::1 M=128 MMA costs max(72, N/2) cycles — the
      // 4 KB A-operand read from SMEM floors it at ~72 regardless of N
      const double per_mma = n_umma / 2.0 > 72.0 ? n_umma / 2.0 : 72.0;
      const double mma = (double)t_m * kx * 4 * per_mma;
      const double load = (halo * 128.0 + (double)kx * w_bytes) / 34.0;
      const long long ctas = (long long)N * ((DX + tx - 1) / tx) * ((DY + by - 1) / by);
      // co-residency: small tiles let 2 CTAs share an SM (SMEM and TMEM permitting), which hides one CTA's
      // barrier / TMA latency behind the other's MMAs (measured: trunk conv 73 -> 55 us)
      const int smem_min = 2 * a_bytes + (kx + 1 < 3 ? 3 : kx + 1) * w_bytes + 2048;
      int cols = 32;
      while (cols < t_m * n_umma) cols <<= 1;
      int cps = (227 * 1024) / (smem_min + 2048) >= 2 && cols <= 256 && !pair ? 2 : 1;  // + static SMEM
      const long long per_sm = (ctas + 147) / 148;
      const int r = per_sm < cps ? (int)per_sm : cps;  // CTAs actually sharing an SM
      const long long waves = (ctas + 148LL * r - 1) / (148LL * r);
      // the issuing thread: ~170 cycles per group of 4 MMAs, ~500 per barrier round trip (wait + fence + commit);
      // co-resident CTAs issue independently
      int wg, mg;
      const int syncs = weight_grouping(a_bytes, w_bytes, kx, r >= 2 ? (227 * 1024) / 2 - 1024 - 2048 : kSmemBudget, 2, wg, mg);
      // (an issuing warp is busy for its MMAs' execution time plus ~200 cycles of scalar work per group of 4; up to 4
      // warps issue, one accumulator each)
      const int n_iss = t_m < 4 ? t_m : 4;
      const double issue = (double)((t_m + n_iss - 1) / n_iss) * kx * (4.0 * per_mma + 200.0) + 500.0 * syncs;
      double iter = (mma > load ? mma : load) * r;
      if (issue > iter) iter = issue;
      iter += 100.0;
      // fixed per-CTA cost (prologue, pipeline ramp, epilogue of t_m tiles) in units of iterations' cycles
      const double cost = (double)waves * (iter + (6000.0 + 2500.0 * t_m) / 30.0);
      if (cost < best) {
        best = cost;
        found = true;
        p.by = by; p.tx = tx; p.t_m = t_m; p.slabrows = slab; p.halo_rows = halo; p.a_buf_bytes = a_bytes;
        p.w_bytes = w_bytes;
        p.a_bufs = r;  // (ab)used as "CTAs per SM this config was costed with"; finalised by the caller
      }
    }
  }
  return found;
}

// Volume mode: the CTA's tile is tx x by x DZ outputs and its activation operand is the (tx+kx-1) x (by+ky-1) x
// (DZ+kz-1) padded halo, fetched ONCE per 64-channel chunk; rows are ordered (x, y, z) with the padded pitches, so
// tap (ti, tj, tl) is the row offset ti*PY*PZ + tj*PZ + tl.  Pad positions inside the row span are computed and
// dropped by the epilogue.  Against the per-(ky,kz) halo this cuts the L2 -> SMEM activation traffic and the number
// of pipeline stages by ~ky*kz: it is for the small-volume trunk convs, which are bound by exactly that.
bool choose_cfg_vol_search(int N, int DX, int DY, int DZ, int kx, int ky, int kz, int n_umma, int kchunks,
                           Tc2Params& p) {
  const int tmax = 512 / n_umma < 4 ? 512 / n_umma : 4;
  if (tmax < 1) return false;
  const int PZ = DZ + kz - 1;
  const int taps = kx * ky * kz;
  double best = 1e30;
  bool found = false;
  static int force_by = 0, force_tx = 0;
  static const bool forced = [] {
    const char* v = getenv("WS_TC2_FORCE");
    return v && sscanf(v, "%d,%d", &force_by, &force_tx) == 2;
  }();
  for (int by = 1; by <= DY; ++by) {
    if (forced && by != force_by) continue;
    const int PY = by + ky - 1;
    if (PY > 256 || PZ > 256) break;
    const int slab = PY * PZ;
    for (int tx = 1; tx <= DX; ++tx) {
      if (forced && tx != force_tx) continue;
      if (tx + kx - 1 > 256) break;
      const int span = (tx - 1) * slab + (by - 1) * PZ + DZ;
      const int t_m = (span + 127) / 128;
      if (t_m > tmax) break;
      const int halo = (tx + kx - 1) * slab;
      int a_rows = t_m * 128 + (kx - 1) * slab + (ky - 1) * PZ + kz - 1;
      for (int want = 1; want <= 4; ++want) {
        const int sub = (tx + kx - 1 + want - 1) / want;
        const int ops_rows = ((tx + kx - 1 + sub - 1) / sub) * sub * slab;
        if (a_rows < ops_rows) a_rows = ops_rows;
      }
      const int a_bytes = (a_rows * 128 + 1023) / 1024 * 1024;
      const int w_bytes = n_umma * 128;
      const int nbuf = kchunks > 1 ? 2 : 1;
      if (nbuf * a_bytes + 3 * w_bytes + 2048 > kSmemBudget) continue;
      const double per_mma = n_umma / 2.0 > 72.0 ? n_umma / 2.0 : 72.0;
      const double mma = (double)t_m * taps * 4 * per_mma;
      const double load = (halo * 128.0 + (double)taps * w_bytes) / 34.0;
      const long long ctas = (long long)N * ((DX + tx - 1) / tx) * ((DY + by - 1) / by);
      const int smem_min = nbuf * a_bytes + 4 * w_bytes + 2048;
      int cols = 32;
      while (cols < t_m * n_umma) cols <<= 1;
      const int cps = (227 * 1024) / (smem_min + 2048) >= 2 && cols <= 256 ? 2 : 1;
      const long long per_sm = (ctas + 147) / 148;
      const int r = per_sm < cps ? (int)per_sm : cps;
      const long long waves = (ctas + 148LL * r - 1) / (148LL * r);
      const double chunk = (mma > load ? mma : load) * r + (150.0 * taps + 300.0) / r;
      const double cost = (double)waves * (chunk * kchunks + 6000.0 + 2500.0 * t_m);
      if (cost < best) {
        best = cost;
        found = true;
        p.by = by; p.tx = tx; p.t_m = t_m; p.slabrows = slab; p.halo_rows = halo; p.a_buf_bytes = a_bytes;
        p.w_bytes = w_bytes;
        p.a_bufs = r;  // CTAs per SM this config was costed with (see choose_cfg_search)
      }
    }
  }
  return found;
}

// The search above costs ~10 us of host time; the trunk launches ~500 convs per step, so memoise it per geometry.
bool choose_cfg(int N, int DX, int DY, int DZ, int kx, int n_umma, bool pair, Tc2Params& p, int vol = 0, int ky = 1,
                int kz = 1, int kchunks = 1) {
  struct Hit { bool ok; int by, tx, t_m, slabrows, halo_rows, a_buf_bytes, w_bytes, a_bufs; };
  static std::mutex mu;
  static std::map<std::array<int, 11>, Hit> memo;  // keyed by the full geometry (no hash collisions)
  const std::array<int, 11> key = {N, DX, DY, DZ, kx, n_umma, (int)pair, vol, vol ? ky : 1, vol ? kz : 1,
                                   vol ? kchunks : 1};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(key);
    if (it != memo.end()) {
      const Hit& h = it->second;
      p.by = h.by; p.tx = h.tx; p.t_m = h.t_m; p.slabrows = h.slabrows; p.halo_rows = h.halo_rows;
      p.a_buf_bytes = h.a_buf_bytes; p.w_bytes = h.w_bytes; p.a_bufs = h.a_bufs;
      return h.ok;
    }
  }
  const bool ok = vol ? choose_cfg_vol_search(N, DX, DY, DZ, kx, ky, kz, n_umma, kchunks, p)
                      : choose_cfg_search(N, DX, DY, DZ, kx, n_umma, pair, p);
  Hit h = {ok, p.by, p.tx, p.t_m, p.slabrows, p.halo_rows, p.a_buf_bytes, p.w_bytes, p.a_bufs};
  std::lock_guard<std::mutex> lk(mu);
  memo[key] = h;
  return ok;
}

}  // namespace

bool tc2_enabled() {
  static const bool off = [] {
    const char* v = getenv("WS_DISABLE_TC_V2");
    return v && v[0] && v[0] != '0';
  }();
  return !off;
}

// mode 0: forward; mode 1: stride-1 dgrad. Returns -1 when this geometry is not covered (caller uses v1).
// src.dtype == WS_F32 selects the TF32 flavour (fp32 tensor maps, weights packed as fp32, kind::tf32).
// dry != 0: only decide whether this geometry is covered (0) or not (-1); nothing is launched.
int tc2_conv_launch(const ConvGeom& g, int mode, const View& src, const void* packed_w, const View& dst,
                    const Epi& ep, cudaStream_t st, const TcOverride* ov, int dry) {
  if (!ov && (g.sx != 1 || g.sy != 1 || g.sz != 1)) return -1;
  const bool tf32 = src.dtype == WS_F32;
  const int esize = tf32 ? 4 : 2;
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  int SX, SY, SZ;
  int taps_total = g.taps();
  if (ov) {
    p.DX = ov->DX; p.DY = ov->DY; p.DZ = ov->DZ; SX = ov->SX; SY = ov->SY; SZ = ov->SZ;
    p.px = ov->px; p.py = ov->py; p.pz = ov->pz; p.ck = ov->ck; p.cn = ov->cn;
  } else if (mode == 0) {
    p.DX = g.xo; p.DY = g.yo; p.DZ = g.zo; SX = g.x; SY = g.y; SZ = g.z;
    p.px = g.px; p.py = g.py; p.pz = g.pz; p.ck = g.cin; p.cn = g.cout;
  } else {
    p.DX = g.x; p.DY = g.y; p.DZ = g.z; SX = g.xo; SY = g.yo; SZ = g.zo;
    p.px = g.kx - 1 - g.px; p.py = g.ky - 1 - g.py; p.pz = g.kz - 1 - g.pz; p.ck = g.cout; p.cn = g.cin;
  }
  p.N = g.n; p.kx = g.kx; p.ky = g.ky; p.kz = g.kz; p.bz = p.DZ;
  p.omx = p.omy = p.omz = 1; p.ODY = p.DY; p.ODZ = p.DZ;
  if (ov) {
    p.kx = ov->kx; p.ky = ov->ky; p.kz = ov->kz; p.tap_base = ov->tap_base; taps_total = ov->taps_total;
    p.omx = ov->omx; p.oax = ov->oax; p.omy = ov->omy; p.oay = ov->oay; p.omz = ov->omz; p.oaz = ov->oaz;
    p.ODY = ov->ODY; p.ODZ = ov->ODZ;
  }
  const int cn_pad = (p.cn + 15) / 16 * 16;
  const int ck_pad = tf32 ? (p.ck + 3) / 4 * 4 : (p.ck + 7) / 8 * 8;  // 16-byte rows of the packed weights
  p.tf32 = tf32 ? 1 : 0;
  const int n_tiles = (cn_pad + 255) / 256;
  p.n_tile = ((cn_pad + n_tiles - 1) / n_tiles + 15) / 16 * 16;
  p.n_umma = p.n_tile;
  // CTA-pair mode (cta_group::2): worth it when the weight tile is wide and the grid is large
  static const int env_pair = getenv("WS_TC2_PAIR") ? atoi(getenv("WS_TC2_PAIR")) : -1;
  const long long vox = (long long)p.N * p.DX * p.DY * p.DZ;
  bool pair = env_pair >= 0 ? env_pair != 0 : (p.n_umma >= 128 && vox >= 148LL * 2 * 384);
  if (n_tiles != 1) pair = false;
  // K <= 32 (the gc = 32 data-gradients of the RRDB trunk, D's 32-channel layers): 64-byte operand rows
  static const bool env_no_sw64 = getenv("WS_DISABLE_SW64") != nullptr;
  const bool sw64 = ck_pad <= 32 && !pair && !env_no_sw64 && !tf32;
  p.row_bytes = sw64 ? 64 : 128;
  p.kelems = p.row_bytes / esize;
  p.kchunks = (p.ck + p.kelems - 1) / p.kelems;
  const int kper = 32 / esize;  // K elements of one instruction: 16 (bf16) or 8 (tf32)
  p.last_k16 = (p.ck - p.kelems * (p.kchunks - 1) + kper - 1) / kper;
  // volume mode is opt-in (WS_TC2_VOL=1): measured on the RRDB trunk it is parity-clean but not faster (dense
  // conv 54 vs 48 us, dgrad 46 vs 45 us) — those layers are bound by the N=32 MMA floor and by the weight
  // stream, not by the activation re-reads this mode removes.
  static const int env_vol = getenv("WS_TC2_VOL") ? atoi(getenv("WS_TC2_VOL")) : 0;
  bool vol = !pair && p.ky * p.kz > 1 && env_vol != 0;
  if (vol && !choose_cfg(p.N, p.DX, p.DY, p.DZ, p.kx, p.n_umma, false, p, 1, p.ky, p.kz, p.kchunks)) vol = false;
  if (!vol && !choose_cfg(p.N, p.DX, p.DY, p.DZ, p.kx, p.n_umma, pair, p)) return -1;
  if (sw64) {  // the searches size their buffers for 128-byte rows
    p.a_buf_bytes = (p.a_buf_bytes / 2 + 1023) / 1024 * 1024;
    p.w_bytes /= 2;
  }
  p.vol = vol ? 1 : 0;
  p.pitch_y = vol ? p.DZ + p.kz - 1 : p.DZ;
  p.box_z = p.pitch_y;
  p.box_y = vol ? p.by + p.ky - 1 : p.by;
  p.out_rows = vol ? (p.tx - 1) * p.slabrows + (p.by - 1) * p.pitch_y + p.DZ : p.tx * p.slabrows;
  p.tiles_x = (p.DX + p.tx - 1) / p.tx;
  p.tiles_y = (p.DY + p.by - 1) / p.by;
  static const int env_split = getenv("WS_TC2_ASPLIT") ? atoi(getenv("WS_TC2_ASPLIT")) : 4;
  static const int env_bufs = getenv("WS_TC2_ABUFS") ? atoi(getenv("WS_TC2_ABUFS")) : kMaxABufs;
  {
    const int slabs = p.tx + p.kx - 1;
    int want_ops = env_split < 1 ? 1 : (env_split > 4 ? 4 : env_split);
    p.a_sub_slabs = (slabs + want_ops - 1) / want_ops;
    p.a_ops = (slabs + p.a_sub_slabs - 1) / p.a_sub_slabs;
  }
  // SMEM budget: the whole SM for one resident CTA, half of it when the config was costed with two
  const int budget = p.a_bufs >= 2 ? (227 * 1024) / 2 - 1024 - 2048 : kSmemBudget;  // 2 KB static SMEM per CTA
  const int ntaps_it = vol ? p.kx * p.ky * p.kz : p.kx;  // weight tiles per halo load
  const int loads = (vol ? 1 : p.ky * p.kz) * p.kchunks;  // activation loads of the whole kernel
  const int need_bufs = loads > 1 ? 2 : 1;
  weight_grouping(p.a_buf_bytes, p.w_bytes, ntaps_it, budget, need_bufs, p.w_group, p.merged);
  const int w_slot = p.w_group * p.w_bytes;
  if (p.merged) {
    p.a_bufs = (budget - 2048) / (p.a_buf_bytes + w_slot);
    if (p.a_bufs > kMaxABufs) p.a_bufs = kMaxABufs;
    if (p.a_bufs > env_bufs) p.a_bufs = env_bufs < 2 ? 2 : env_bufs;
    if (p.a_bufs > loads) p.a_bufs = loads;
    if (p.a_bufs < need_bufs) return -1;
    p.w_slots = p.a_bufs;
  } else {
    // two weight-group slots, as many halo buffers as the rest allows, then spare room back to the weight ring
    p.w_slots = 2;
    if (need_bufs * p.a_buf_bytes + p.w_slots * w_slot + 2048 > budget) return -1;
    p.a_bufs = (budget - 2048 - p.w_slots * w_slot) / p.a_buf_bytes;
    if (p.a_bufs > kMaxABufs) p.a_bufs = kMaxABufs;
    if (p.a_bufs > env_bufs) p.a_bufs = env_bufs < 2 ? 2 : env_bufs;
    if (p.a_bufs > loads) p.a_bufs = loads;
    if (p.a_bufs < need_bufs) return -1;
    int spare = (budget - 2048 - p.a_bufs * p.a_buf_bytes) / w_slot;
    if (spare > kMaxWSlots) spare = kMaxWSlots;
    if (spare > p.w_slots) p.w_slots = spare;
  }
  uint32_t cols = 32;
  while ((int)cols < p.t_m * p.n_umma) cols <<= 1;
  p.tmem_cols = cols;
  static const bool env_no_lean = getenv("WS_TC2_LEAN") && atoi(getenv("WS_TC2_LEAN")) == 0;
  p.lean = env_no_lean ? 0 : 1;
  static const int env_niss = getenv("WS_TC2_NISS") ? atoi(getenv("WS_TC2_NISS")) : 4;
  p.n_iss = p.t_m < 4 ? p.t_m : 4;
  if (p.n_iss > env_niss) p.n_iss = env_niss < 1 ? 1 : env_niss;
  static const int env_cm = getenv("WS_TC2_CHUNK_MAJOR") ? atoi(getenv("WS_TC2_CHUNK_MAJOR")) : -1;
  p.chunk_major = env_cm >= 0 ? (env_cm != 0) : (p.kchunks > 1 && p.last_k16 < p.row_bytes / 32);
  if (dry) return 0;

  MapKey ka;
  memset(&ka, 0, sizeof(ka));
  ka.ptr = reinterpret_cast<uintptr_t>(src.ptr);
  ka.rank = 5; ka.dtype = (tf32 ? WS_F32 : WS_BF16) | (sw64 ? kMapSwizzle64 : 0u);
  ka.dims[0] = (uint64_t)p.ck; ka.dims[1] = (uint64_t)SZ; ka.dims[2] = (uint64_t)SY; ka.dims[3] = (uint64_t)SX;
  ka.dims[4] = (uint64_t)g.n;
  ka.strides[0] = (uint64_t)src.vs * esize;
  ka.strides[1] = (uint64_t)src.vs * esize * SZ;
  ka.strides[2] = (uint64_t)src.vs * esize * SZ * SY;
  ka.strides[3] = (uint64_t)src.ns * esize;
  ka.box[0] = (uint32_t)p.kelems; ka.box[1] = (uint32_t)p.box_z; ka.box[2] = (uint32_t)p.box_y; ka.box[3] = (uint32_t)p.a_sub_slabs;
  ka.box[4] = 1;
  for (int i = 0; i < 5; ++i) ka.estr[i] = 1;
  CUtensorMap tmA, tmB;
  if (int e = get_tensor_map(ka, &tmA)) return e;
  MapKey kb;
  memset(&kb, 0, sizeof(kb));
  kb.ptr = reinterpret_cast<uintptr_t>(packed_w);
  kb.rank = 3; kb.dtype = (tf32 ? WS_F32 : WS_BF16) | (sw64 ? kMapSwizzle64 : 0u);
  kb.dims[0] = (uint64_t)ck_pad; kb.dims[1] = (uint64_t)cn_pad; kb.dims[2] = (uint64_t)taps_total;
  kb.strides[0] = (uint64_t)ck_pad * esize;
  kb.strides[1] = (uint64_t)ck_pad * esize * cn_pad;
  kb.box[0] = (uint32_t)p.kelems; kb.box[1] = (uint32_t)(pair ? p.n_umma / 2 : p.n_umma); kb.box[2] = 1;
  kb.estr[0] = kb.estr[1] = kb.estr[2] = 1;
  if (int e = get_tensor_map(kb, &tmB)) return e;

  size_t smem = (size_t)p.a_bufs * p.a_buf_bytes + (size_t)p.w_slots * p.w_group * p.w_bytes +
                8 * (2 * kMaxABufs + 2 * kMaxWSlots + 2) + 1024;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    auto set = [](const void* f) {
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);  // + 2 KB static
    };
    set((const void*)conv3d_tc2_kernel<false, 4>);
    set((const void*)conv3d_tc2_kernel<false, 8>);
    set((const void*)conv3d_tc2_kernel<true, 4>);
    set((const void*)conv3d_tc2_kernel<true, 8>);
  });
  WS_REQUIRE(attr_err == cudaSuccess, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  WS_REQUIRE(smem <= 225 * 1024, "conv_tc2: smem request %zu too large", smem);
  const unsigned tiles = (unsigned)(p.N * p.tiles_x * p.tiles_y);
  // WS_TC2_DEBUG_TIMES=1: per-CTA stage cycle counts (clock64), printed as means after a blocking sync — profiling only
  static const bool dbg_on = getenv("WS_TC2_DEBUG_TIMES") && atoi(getenv("WS_TC2_DEBUG_TIMES")) != 0;
  static long long* dbg_buf = nullptr;
  const size_t dbg_ctas = (size_t)(tiles + 1) * (size_t)n_tiles;
  if (dbg_on && dbg_ctas <= 16384) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 16384 * 8 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, dbg_ctas * 8 * sizeof(long long), st);
    p.dbg = dbg_buf;
  }
  // 8 epilogue warps when the CTA cannot share its SM anyway (SMEM) and there is more than one chunk per warp
  static const int env_epi = getenv("WS_TC2_EPI8") ? atoi(getenv("WS_TC2_EPI8")) : -1;
  const bool single = smem > 113 * 1024 || pair;
  const bool epi8 = env_epi >= 0 ? (env_epi != 0 && single) : (single && p.n_umma >= 32);
  if (!pair) {
    dim3 grid(tiles, (unsigned)n_tiles);
    if (epi8)
      WS_CHECK_CUDA(launch_pdl(conv3d_tc2_kernel<false, 8>, grid, dim3(320), smem, st, 1, tmA, tmB, p, dst, ep));
    else
      WS_CHECK_CUDA(launch_pdl(conv3d_tc2_kernel<false, 4>, grid, dim3(192), smem, st, 1, tmA, tmB, p, dst, ep));
  } else {
    // pairs of adjacent tiles; an odd tail gets a padding CTA
    const dim3 grid((tiles + 1) / 2 * 2, 1, 1);
    if (epi8)
      WS_CHECK_CUDA(launch_pdl(conv3d_tc2_kernel<true, 8>, grid, dim3(320), smem, st, 2, tmA, tmB, p, dst, ep));
    else
      WS_CHECK_CUDA(launch_pdl(conv3d_tc2_kernel<true, 4>, grid, dim3(192), smem, st, 2, tmA, tmB, p, dst, ep));
  }
  WS_POST_LAUNCH(1);
  if (p.dbg) {
    cudaStreamSynchronize(st);
    std::vector<long long> h(dbg_ctas * 8);
    cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum[8] = {0};
    long long n_mma = 0, n_all = 0;
    for (size_t c = 0; c < dbg_ctas; ++c) {
      if (h[8 * c + 0]) { ++n_mma; for (int j = 0; j < 3; ++j) sum[j] += (double)h[8 * c + j]; }
      if (h[8 * c + 4]) { ++n_all; for (int j = 3; j < 8; ++j) sum[j] += (double)h[8 * c + j]; }
    }
    if (n_mma && n_all)
      fprintf(stderr,
              "[tc2 dbg] pair=%d by=%d tx=%d t_m=%d a_bufs=%d w_slots=%d w_group=%d merged=%d n_iss=%d kchunks=%d grid=%u | issue loop %.0f cyc (wait halo %.0f, "
              "wait weights %.0f) | accumulators ready after %.0f, epilogue %.0f (TMEM loads %.0f) | producer waits: halo slot %.0f, weight "
              "slot %.0f | MMA floor %.0f\n",
              (int)pair, p.by, p.tx, p.t_m, p.a_bufs, p.w_slots, p.w_group, p.merged, p.n_iss, p.kchunks, tiles, sum[0] / n_mma, sum[1] / n_mma,
              sum[2] / n_mma, sum[3] / n_all, sum[4] / n_all, sum[7] / n_all, sum[5] / n_all, sum[6] / n_all,
              (double)p.t_m * p.kx * p.ky * p.kz * (4.0 * (p.kchunks - 1) + p.last_k16) * (p.n_umma / 2.0 > 72.0 ? p.n_umma / 2.0 : 72.0));
  }
  return 0;
}

}  // namespace ws
