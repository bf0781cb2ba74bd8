/*
 * windsr.h — C-ABI of the B200-native (sm_100a) hot path for GAN_SR_wind_field.
 *
 * The reference (jacobwulffwold/GAN_SR_wind_field) is pure Python and reaches the GPU only through
 * stock torch ops, so it has no FFI of its own.  Every entry point below therefore cites the reference
 * *call site* whose arithmetic it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success, non-zero on failure (ws_last_error() has the message);
 *     nothing throws or exits across the ABI;
 *   - no allocation inside: workspaces are passed in (ws_*_workspace_bytes gives the size);
 *   - every launch goes to the cudaStream_t passed as `void* stream` (re-entrant per stream);
 *   - activations are addressed through ws_tensor views, so both the reference's logical
 *     (N,C,X,Y,Z) layout and the internal channels-last (N,X,Y,Z,C) layout (and channel slices of
 *     a wider dense-concat buffer) are expressible without copies.
 */
#ifndef WINDSR_H_
#define WINDSR_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WS_VERSION 1

/* element types of activation tensors */
enum { WS_F32 = 0, WS_BF16 = 1 };

/* arithmetic mode of a convolution pass (north-star tolerances 1e-5 / 2e-3 / 2e-2) */
enum {
  WS_MATH_FP32 = 0, /* CUDA-core FFMA, fp32 accumulate: the 1e-5 parity mode and the narrow layers     */
  WS_MATH_TF32 = 1, /* tcgen05 kind::tf32 on fp32 activations                                            */
  WS_MATH_BF16 = 2  /* tcgen05 kind::f16 (bf16 operands, fp32 TMEM accumulators)                          */
};

/* which kernel family a call resolved to (ws_conv3d_*_path); tests assert on it */
enum { WS_PATH_NONE = 0, WS_PATH_SIMT = 1, WS_PATH_TCGEN05 = 2 };

/* weight packings (ws_pack_weights) */
enum {
  WS_PACK_SIMT_FWD = 0,   /* fp32 [tap][cin][cout]                 B operand of fwd on CUDA cores        */
  WS_PACK_SIMT_DGRAD = 1, /* fp32 [tap][cout][cin]                 B operand of dgrad on CUDA cores       */
  WS_PACK_TC_FWD = 2,     /* bf16 [tap][cout_pad16][cin_pad8]      K-major B operand for tcgen05 fwd      */
  WS_PACK_TC_DGRAD = 3,   /* bf16 [flipped tap][cin_pad16][cout_pad8]  K-major B operand for tcgen05 dgrad */
  WS_PACK_TC_FWD_TF32 = 4,   /* fp32 [tap][cout_pad16][cin_pad4]            the same for WS_MATH_TF32          */
  WS_PACK_TC_DGRAD_TF32 = 5  /* fp32 [flipped tap][cin_pad16][cout_pad4]                                       */
};

/* A view of a 5-D activation: element (n, c, v) with v = (x*Y + y)*Z + z lives at
 *   ptr + n*nstride + v*vstride + c*cstride   (strides in ELEMENTS of dtype).
 * channels-last:  vstride = C_total, cstride = 1, nstride = V*C_total   (ptr may point at a channel offset)
 * reference NCXYZ: vstride = 1, cstride = V, nstride = C*V */
typedef struct ws_tensor {
  void* ptr;
  int32_t dtype;
  int32_t _pad;
  int64_t nstride;
  int64_t vstride;
  int64_t cstride;
} ws_tensor;

/* Geometry of one Conv3d layer as torch.nn.Conv3d defines it (torch_blocks.py:17,278;
 * Generator_3D_Resnet_ESRGAN.py:105): input (n, cin, x, y, z) -> output (n, cout, xo, yo, zo),
 * xo = (x + 2*px - kx)/sx + 1 etc. */
typedef struct ws_conv_shape {
  int32_t n, x, y, z;
  int32_t cin, cout;
  int32_t kx, ky, kz;
  int32_t sx, sy, sz;
  int32_t px, py, pz;
} ws_conv_shape;

/* Fused epilogue applied to each accumulator element acc(n, c, v):
 *   t = acc * (oscale ? oscale[c] : 1) + (bias ? bias[c] : 0)
 *   t = t > 0 ? t : lrelu_slope * t                     (lrelu_slope == 1 -> identity)
 *   t = t * (chan_scale ? chan_scale[n*C + c] : 1)      (Dropout3d mask * 1/(1-p), Generator…py:70-74)
 *   y = alpha * t + beta1 * res1(n,c,v) + beta2 * res2(n,c,v)
 *   if (mask.ptr && mask_c0 <= c < mask_c1)  y *= (mask(n,c,v) > 0 ? 1 : mask_slope)   (LeakyReLU backward
 *                                               of the producer layer, fused into this dgrad)
 *   out(n,c,v) = y;  out2(n,c,v) = y (optional second dtype/layout)
 *   stat_sum[c] += t_preact, stat_sqsum[c] += t_preact^2 over (n,v)   (BatchNorm3d batch statistics of the
 *                                               raw conv output, torch_blocks.py:24-25; taken before lrelu)
 * This covers: conv+LeakyReLU (torch_blocks.py:35), the RDB dense-concat slice write (:214, out = channel
 * slice of the concat buffer), LFF bias + 0.2*residual + x (:278-290), RRDB 0.2*residual + x (:328-330),
 * the trunk skip x + module(x) (:46), and hr_convs.2 bias (Generator…py:105-110). */
typedef struct ws_epilogue {
  const float* bias;
  const float* oscale;
  const float* chan_scale;
  float lrelu_slope;
  float alpha;
  float beta1;
  float beta2;
  ws_tensor res1;
  ws_tensor res2;
  ws_tensor mask;
  int32_t mask_c0, mask_c1;
  float mask_slope;
  int32_t flags; /* bit 0: round the stored values to the nearest TF32 value (outputs that only feed WS_MATH_TF32 convs:
                    the tensor cores truncate fp32 operands, rounding first halves the error and removes its bias) */
  ws_tensor out2;
  float* stat_sum;
  float* stat_sqsum;
  /* "tail" channels [tail_c0, cout) (tail_c0 % 16 == 0; tensor-core kernels only): instead of going to `out`, value t
   * is written as  t * (tail_mask > 0 ? 1 : tail_slope)  to channel (c - tail_c0) of tail_out — the LeakyReLU-backward
   * of the previous dense conv fused into a data-gradient of the residual dense block (torch_blocks.py:192-214). */
  ws_tensor tail_out;
  ws_tensor tail_mask;
  int32_t tail_c0;
  float tail_slope;
} ws_epilogue;

/* ---- library ---------------------------------------------------------------------------------- */
int ws_version(void);
const char* ws_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05 path usable) */
int ws_device_supports_tcgen05(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t ws_launch_count(void);

/* ---- weights ---------------------------------------------------------------------------------- */
/* bytes of a packed weight buffer of the given kind */
size_t ws_packed_weight_bytes(const ws_conv_shape* s, int pack_kind);
/* w: torch layout (cout, cin, kx, ky, kz) fp32 (state_dict layout) -> packed */
int ws_pack_weights(const float* w, const ws_conv_shape* s, int pack_kind, void* packed, void* stream);

/* ---- Conv3d passes (replace nn.Conv3d forward + autograd convolution_backward; torch_blocks.py:17) */
int ws_conv3d_fwd_path(const ws_conv_shape* s, const ws_tensor* in, const ws_tensor* out, int math);
int ws_conv3d_dgrad_path(const ws_conv_shape* s, const ws_tensor* dy, const ws_tensor* dx, int math);
int ws_conv3d_wgrad_path(const ws_conv_shape* s, const ws_tensor* in, const ws_tensor* dy, int math);

/* out = epilogue(conv(in, W)).  `packed_w` must be the packing matching ws_conv3d_fwd_path(). */
int ws_conv3d_fwd(const ws_conv_shape* s, const ws_tensor* in, const void* packed_w, const ws_tensor* out,
                  const ws_epilogue* ep, int math, void* stream);
/* dx = epilogue(conv_transpose(dy, W)) : gradient w.r.t. the conv input.  `dx` channel count = cin.
 * Use ep->res1 = dx, beta1 = 1 to accumulate into an existing gradient (dense-concat backward). */
int ws_conv3d_dgrad(const ws_conv_shape* s, const ws_tensor* dy, const void* packed_w, const ws_tensor* dx,
                    const ws_epilogue* ep, int math, void* stream);
/* dw (torch layout (cout,cin,kx,ky,kz) fp32) = [accumulate ? dw : 0] + sum_{n,v} dy(n,co,v) * in(n,ci,v (+) tap)
 * db (optional, [cout]) likewise gets sum dy. workspace: ws_conv3d_wgrad_workspace_bytes(). */
size_t ws_conv3d_wgrad_workspace_bytes(const ws_conv_shape* s, int math);
int ws_conv3d_wgrad(const ws_conv_shape* s, const ws_tensor* in, const ws_tensor* dy, float* dw, float* db,
                    int accumulate, int math, void* workspace, size_t workspace_bytes, void* stream);

/* ---- im2col for the narrow first layers (Cin <= 4) ------------------------------------------------------ */
/* u (n, cpad, xo, yo, zo) channels-last bf16, cpad % 8 == 0, cpad >= taps*cin:
 *   u[n, v, tap*cin + ci] = x[n, ci, v (+) tap]   (zero padding, zero pad columns)
 * so that conv(x, w) == conv1x1x1(u, w') with w'[co][tap*cin+ci] = w[co][ci][tap]: the 3 -> 32 first conv of the
 * discriminator (torch_blocks.py:372-521 via Discriminator_3D.py:67-136) then runs on the tensor-core kernels. */
int ws_im2col(const ws_conv_shape* s, const ws_tensor* x, const ws_tensor* u, int cpad, void* stream);

/* ---- residual dense block executor (torch_blocks.py:217-290, 328-330) ------------------------------- */
/* One call runs a whole RDB: (re)pack its weights, 4 dense conv + LeakyReLU launches writing channel slices of the
 * concat buffer, then the LFF conv whose epilogue applies  out = alpha*(LFF(buf)+bias) + beta1*x + beta2*outer.
 * The launch-bound trunk (240 of 249 generator convs) spends its time in launch overhead otherwise. */
#define WS_RDB_MAX_CONVS 8
typedef struct ws_rdb_desc {
  int32_t n, x, y, z;       /* LR volume                                   */
  int32_t f, gc, nconv;     /* in/out channels, growth channels, dense convs */
  int32_t k, k_lff;         /* dense kernel size (3), LFF kernel size (1)    */
  float slope, alpha, beta1, beta2;
  int32_t math;
  int32_t repack;           /* 1: weights changed since the last call -> pack again */
} ws_rdb_desc;
/* bytes each per-conv packed-weight buffer must have (max over the packings this path may choose) */
size_t ws_rdb_packed_bytes(const ws_rdb_desc* d, int conv_index /* 0..nconv-1 dense, nconv = LFF */, int dgrad);
/* x: fp32 trunk state (n,f,..); outer: optional fp32 (RRDB input); buf: (n, f+nconv*gc, ..) activation-dtype
 * concat buffer (filled here, saved for backward); out: fp32 (n,f,..).
 * w[i]: torch-layout fp32 weights, packed[i]: device buffers of ws_rdb_packed_bytes(i, 0). */
/* workspace (optional, ws_rdb_forward_workspace_bytes): fp32 scratch of the x-folded dense convs — with it the
 * tensor-core path runs each k^3 dense conv as a (1,k,k) conv with the kx taps side by side on N (3x fewer MMAs)
 * plus a shifted sum + LeakyReLU pass; without it (NULL) the convs run in their direct form. */
size_t ws_rdb_forward_workspace_bytes(const ws_rdb_desc* d);
int ws_rdb_forward(const ws_rdb_desc* d, const ws_tensor* x, const ws_tensor* outer, const ws_tensor* buf,
                   const ws_tensor* out, const float* const* w, void* const* packed, const float* lff_bias,
                   void* workspace, size_t workspace_bytes, void* stream);
/* bytes of the `workspace` ws_rdb_backward needs */
size_t ws_rdb_backward_workspace_bytes(const ws_rdb_desc* d);
/* aux_stream (optional, with its own aux_workspace of ws_rdb_backward_workspace_bytes): the weight gradients run on
 * it beside the data-gradient chain; the caller must make the consumer of dw / db_lff wait for aux_stream and keep
 * buf, g_lff, g alive until then.  NULL: everything on `stream`. */
/* dy: fp32 gradient of `out`.  Scratch: dbuf fp32 (n, f+nconv*gc, ..), g_lff activation-dtype (n,f,..),
 * g activation-dtype (n,nconv*gc,..) — the output gradients of all dense convs side by side (their weight
 * gradients are one merged GEMM on the tensor-core path).  dx (optional) = dL/dx incl. the beta1 skip.  dw[i] (optional, torch layout)
 * and db_lff receive the parameter gradients (overwritten).  packed[i]: buffers of ws_rdb_packed_bytes(i, 1). */
int ws_rdb_backward(const ws_rdb_desc* d, const ws_tensor* dy, const ws_tensor* buf, const ws_tensor* dbuf,
                    const ws_tensor* g_lff, const ws_tensor* g, const ws_tensor* dx, const float* const* w,
                    void* const* packed, float* const* dw, float* db_lff, void* workspace,
                    size_t workspace_bytes, void* stream, void* aux_stream, void* aux_workspace,
                    size_t aux_workspace_bytes);

/* ---- weight gradients of a run of identical residual dense blocks, batched (torch_blocks.py:256-290 x 48) ------- */
/* The weight-gradient half of convolution_backward for every conv of `nblocks` consecutive RDBs in ONE set of launches.
 * The caller ran ws_rdb_backward with dw = NULL for each block, with buf / g / g_lff of block r living in slabs whose
 * block index is the outermost dimension: the tensors passed here describe block 0 and block r starts
 * r * d->n * nstride elements later.  grads: per block one record  [dw_0 | ... | dw_{nconv-1} | dw_lff | db_lff]
 * (each in torch layout, ws_trunk_wgrad_record_floats floats, consecutive records block_stride floats apart; overwritten).
 * Tensor-core BF16 path only (returns an error otherwise — the per-block path of ws_rdb_backward covers the rest). */
int ws_trunk_wgrad_supported(const ws_rdb_desc* d); /* 1: this geometry / precision runs on the batched path */
size_t ws_trunk_wgrad_record_floats(const ws_rdb_desc* d);
size_t ws_trunk_wgrad_workspace_bytes(const ws_rdb_desc* d, int nblocks);
int ws_trunk_wgrad(const ws_rdb_desc* d, int nblocks, const ws_tensor* buf, const ws_tensor* g, const ws_tensor* g_lff,
                   float* grads, long long block_stride, void* workspace, size_t workspace_bytes, void* stream);

/* Weight (re)packing for ALL blocks of such a run in one launch per direction (the persistent per-block kernels'
 * layouts: z-folded forward packing / data-gradient packing); the caller then runs ws_rdb_forward / ws_rdb_backward with
 * repack = 0.  The tensors describe block 0 (only layouts are examined).  w / packed: [block][conv 0..nconv-1, LFF],
 * packed[i] of ws_rdb_packed_bytes.  Returns 0 = enqueued, 1 = this geometry does not take the persistent kernels (repack
 * block by block instead), anything else = error. */
int ws_trunk_repack_fwd(const ws_rdb_desc* d, int nblocks, const ws_tensor* x, const ws_tensor* buf, const ws_tensor* out,
                        const float* const* w, void* const* packed, void* stream);
int ws_trunk_repack_bwd(const ws_rdb_desc* d, int nblocks, const ws_tensor* dy, const ws_tensor* buf,
                        const ws_tensor* g_lff, const ws_tensor* g, const ws_tensor* dx, const float* const* w,
                        void* const* packed, void* stream);

/* ---- nearest upsample (x2 in x and y, z untouched): nn.Upsample(scale_factor=(2,2,1)) torch_blocks.py:347 */
/* in: (n, c, x, y, z) view, out: (n, c, 2x, 2y, z) view; bit-exact gather out[x,y,z] = in[x/2, y/2, z] */
int ws_upsample_nearest_xy_fwd(const ws_tensor* in, const ws_tensor* out, int n, int c, int x, int y, int z,
                               void* stream);
/* din[x,y,z] = sum of the 2x2 block of dout */
int ws_upsample_nearest_xy_bwd(const ws_tensor* dout, const ws_tensor* din, int n, int c, int x, int y, int z,
                               void* stream);

/* ---- x-fold helpers for very narrow outputs (hr_convs.2: 144 -> 3, Generator…py:105-110) ----------------
 * A tcgen05 MMA costs the same for N = 16 as for N = 144, so a Cout = 3 conv wastes the tensor pipe.  Folding the
 * kx taps along x into the output-channel dimension — Y[x', (dx,co)] = sum_{dy,dz,ci} W[co,ci,dx,dy,dz] *
 * in[x', y+dy, z+dz, ci], out[x, co] = bias[co] + sum_dx Y[x + dx - pad, (dx,co)] — turns the 5x5x5 conv into a
 * 1x5x5 conv with 15 (->16) output channels (5x fewer MMAs) plus this shifted sum; the same unfolding of the
 * output gradient, U[x', (dx,co)] = dout[x' - dx + pad, co], serves dgrad and wgrad. */
int ws_xfold_sum(const ws_tensor* y, const float* bias, const ws_tensor* out, int n, int co, int kx, int pad,
                 int x, int yy, int z, void* stream);
int ws_xunfold(const ws_tensor* dout, const ws_tensor* u, int n, int co, int kx, int pad, int cpad, int x, int yy,
               int z, void* stream);
/* The same with BOTH lateral axes folded (round 2): Y[x', y', z, (dx*ky+dy)*co + c] from a (1,1,kz) conv with kx*ky*co
 * (75 -> 80) output channels — 25x fewer MMAs than the direct 5x5x5 form —
 *   out[x, y, z, c] = bias[c] + sum_{dx,dy} Y[x + dx - px, y + dy - py, z, (dx*ky+dy)*co + c]          (co <= 8)
 *   U[x', y', z, (dx*ky+dy)*co + c] = dout[x' - dx + px, y' - dy + py, z, c]   (cpad % 8 == 0 channels, pad = 0) */
int ws_xyfold_sum(const ws_tensor* y, const float* bias, const ws_tensor* out, int n, int co, int kx, int ky, int px,
                  int py, int x, int yy, int z, void* stream);
int ws_xyunfold(const ws_tensor* dout, const ws_tensor* u, int n, int co, int kx, int ky, int px, int py, int cpad,
                int x, int yy, int z, void* stream);

/* ---- elementwise helpers -------------------------------------------------------------------------- */
/* dst(n,c,v) = src(n,c,v) with dtype/layout conversion (the torch.cat / clone / layout changes of
 * torch_blocks.py:214,286 and Generator…py:228 expressed as strided copies) */
int ws_copy(const ws_tensor* src, const ws_tensor* dst, int n, int c, int64_t v, void* stream);
/* y = a*x1 + b*x2 (x2 optional) */
int ws_axpby(const ws_tensor* x1, float a, const ws_tensor* x2, float b, const ws_tensor* y, int n, int c,
             int64_t v, void* stream);
/* g = dy * (y > 0 ? 1 : slope) * (chan_scale ? chan_scale[n*C+c] : 1) * (oscale ? oscale[c] : 1):
 * LeakyReLU backward from the saved OUTPUT (torch_blocks.py:35), with the Dropout3d channel scale
 * (Generator…py:70-74) and an eval-mode BatchNorm scale folded in. */
int ws_lrelu_bwd(const ws_tensor* dy, const ws_tensor* y, float slope, const float* chan_scale,
                 const float* oscale, const ws_tensor* g, int n, int c, int64_t v, void* stream);

/* ---- BatchNorm3d (torch_blocks.py:24-25; defaults eps 1e-5, momentum 0.1, affine) ------------------ */
/* From per-channel sum / sum-of-squares of the raw conv output over `count` elements:
 * scale[c] = gamma/sqrt(var_biased+eps), shift[c] = beta - mean*scale; save_mean/save_invstd for backward;
 * running stats updated with the unbiased variance when running_mean != NULL. */
int ws_bn_finalize(const float* sum, const float* sqsum, int64_t count, int c, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
/* y = lrelu(x*scale[c] + shift[c]) */
int ws_scale_shift_lrelu(const ws_tensor* x, const float* scale, const float* shift, float slope,
                         const ws_tensor* y, int n, int c, int64_t v, void* stream);
/* Backward of y = lrelu(bn(x)) in training mode. pass 1 accumulates per-channel sums
 * sum_g = sum g, sum_gx = sum g*xhat with g = dy*(y>0?1:slope);  pass 2 writes
 * dx = gamma*invstd*(g - sum_g/count - xhat*sum_gx/count).  dgamma = sum_gx, dbeta = sum_g. */
int ws_bn_lrelu_bwd_reduce(const ws_tensor* dy, const ws_tensor* y, const ws_tensor* x, const float* mean,
                           const float* invstd, float slope, float* sum_g, float* sum_gx, int n, int c,
                           int64_t v, void* stream);
int ws_bn_lrelu_bwd_apply(const ws_tensor* dy, const ws_tensor* y, const ws_tensor* x, const float* mean,
                          const float* invstd, const float* gamma, const float* sum_g, const float* sum_gx,
                          float slope, int64_t count, const ws_tensor* dx, int n, int c, int64_t v,
                          void* stream);

/* ---- wind-field loss stencils (process_data.py:273-313; wind_field_GAN_3D.py:377-424,773-814) ----- */
/* Result slots of ws_windloss_fwd (float[WS_WL_SLOTS]) */
enum {
  WS_WL_SUM_XY = 0,    /* sum over ch 0-5 of (dSR - dHR)^2                      (count 6*N*V) */
  WS_WL_SUM_Z = 1,     /* sum over ch 6-8                                        (count 3*N*V) */
  WS_WL_SUM_DIV = 2,   /* sum (divSR - divHR)^2, div = ch0+ch4+ch8              (count N*V)   */
  WS_WL_SUM_DIVXY = 3, /* sum (ch0+ch4 difference)^2                            (count N*V)   */
  WS_WL_SUM_PIX_L1 = 4,/* sum |HR - SR| over the 3 wind channels                (count 3*N*V) */
  WS_WL_SUM_PIX_L2 = 5,/* sum (HR - SR)^2                                                     */
  WS_WL_MAX_HR_XY = 6, WS_WL_MAX_SR_XY = 7,     /* max |.| over ch 0-5                          */
  WS_WL_MAX_HR_Z = 8, WS_WL_MAX_SR_Z = 9,       /* max (no abs!) over ch 6-8 (wind_field_GAN_3D.py:780-781) */
  WS_WL_MAX_HR_DIV = 10, WS_WL_MAX_SR_DIV = 11, /* max |ch0+ch4+ch8|                            */
  WS_WL_MAX_HR_DIVXY = 12, WS_WL_MAX_SR_DIVXY = 13,
  WS_WL_SLOTS = 16,        /* floats holding results                                                   */
  WS_WL_RESULT_FLOATS = 64 /* size of the `result` buffer: slots + internal fp64/u64 reduction scratch */
};
/* Coefficient slots of ws_windloss_bwd (device float[WS_WLB_SLOTS]) */
enum {
  WS_WLB_DSUM_XY = 0, WS_WLB_DSUM_Z = 1, WS_WLB_DSUM_DIV = 2, WS_WLB_DSUM_DIVXY = 3, /* dL/d(sum_k)      */
  WS_WLB_DSUM_PIX_L1 = 4, WS_WLB_DSUM_PIX_L2 = 5,
  WS_WLB_DMAX_SR_XY = 6, WS_WLB_DMAX_SR_Z = 7, WS_WLB_DMAX_SR_DIV = 8, WS_WLB_DMAX_SR_DIVXY = 9, /* dL/d(SR max_k),
                                   non-zero only when SR_max/100 > HR_max (wind_field_GAN_3D.py:805-814) */
  WS_WLB_SLOTS = 12
};
/* Per-axis 3-point coefficients of torch.gradient(spacing=coords, edge_order=1) (process_data.py:303).
 * coef has 6*len floats: [6i+0..2] the forward form (a,b,c on f[i-1],f[i],f[i+1]; at the two edge points
 * slot 1 holds the spacing used as divisor), [6i+3..5] the same row as plain linear coefficients. */
int ws_axis_coeffs(const float* coords, int len, float* coef, void* stream);
/* 9-channel Jacobian of a wind field (calculate_gradient_of_wind_field, process_data.py:301-313):
 * field (n,3,x,y,z), zalt (n,1,x,y,z) raw altitude, out (n,9,x,y,z): ch 0-2 d/dx, 3-5 d/dy, 6-8 d/dz. */
int ws_wind_gradient(const ws_tensor* field, const ws_tensor* zalt, const float* coef_x, const float* coef_y,
                     const ws_tensor* out, int n, int x, int y, int z, void* stream);
/* One fused pass over HR, SR, Z: all sums and maxes above, nothing materialised.  `result` is a device
 * buffer of WS_WL_RESULT_FLOATS floats, 8-byte aligned; the call initialises it itself on `stream`.
 * argmax[4] (device int64, optional): flat index of the element attaining each SR max (xy, z: index into
 * the virtual (n,9,v) SR Jacobian; div, divxy: voxel index n*V+v), needed when the normaliser's gradient
 * flows through SR_max/100. */
int ws_windloss_fwd(const ws_tensor* hr, const ws_tensor* sr, const ws_tensor* zalt, const float* coef_x,
                    const float* coef_y, int n, int x, int y, int z, float* result, int64_t* argmax,
                    void* stream);
/* dSR(n,3,v) = sum_k coef[DSUM_k] * d(sum_k)/dSR + coef[DSUM_PIX_L1] * sign(SR-HR)
 *             + coef[DSUM_PIX_L2] * 2(SR-HR) + sum_k coef[DMAX_SR_k] * d(SR max_k)/dSR.
 * coef: device float[WS_WLB_SLOTS].  workspace: ws_windloss_bwd_workspace_bytes() = 9*N*V floats. */
size_t ws_windloss_bwd_workspace_bytes(int n, int x, int y, int z);
int ws_windloss_bwd(const ws_tensor* hr, const ws_tensor* sr, const ws_tensor* zalt, const float* coef_x,
                    const float* coef_y, int n, int x, int y, int z, const float* coef,
                    const int64_t* argmax, const ws_tensor* dsr, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- optimiser: multi-tensor Adam (GAN_models/wind_field_GAN_3D.py:151-162; guard :457-460) ---------------------- */
/* One entry per parameter tensor (all fp32, device pointers).  `step` is the tensor's own step counter (a float, like
 * torch.optim.Adam's state["step"]), read as t = step + 1 and incremented by the call. */
typedef struct ws_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* step;
  int64_t numel;
} ws_adam_tensor;
/* elements of one work chunk: the caller lists (tensor index, chunk index) int32 pairs covering every tensor */
int ws_adam_chunk_elems(void);
/* torch.optim.Adam's update (amsgrad off) for all tensors in ONE launch (+ a step-counter bump):
 *   g' = grad * grad_scale + weight_decay * p;  m += (g' - m)(1 - beta1);  v = beta2 v + (1 - beta2) g'^2
 *   p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
 * table: device array of ws_adam_tensor[ntensors]; chunks: device int32[nchunks][2].  lr_dev (optional device scalar)
 * overrides lr — a CUDA-graph replay then follows the MultiStepLR schedule (wind_field_GAN_3D.py:163-174).
 * found_inf (optional device float): non-zero -> the whole call is a no-op, the reference's "skip the step when the
 * loss is NaN/Inf" guard without a host read. */
int ws_adam_step(const void* table, const void* chunks, int ntensors, int nchunks, const float* lr_dev, float lr,
                 double beta1, double beta2, float eps, float weight_decay, float grad_scale, const float* found_inf,
                 void* stream);

/* ---- instance noise (tools/trainingtricks.py:49-58; used wind_field_GAN_3D.py:250-299) --------------------------- */
/* out[i] = x[i] + U[0,1) * scale * (scale_dev ? *scale_dev : 1), n contiguous fp32 elements (out may alias x).
 * Counter-based Philox4x32-10 keyed by `seed`; state: device uint64[2] (zero-initialised by the caller once) holding
 * the call counter, advanced by the kernel itself so that CUDA-graph replays draw fresh noise. */
int ws_instance_noise(const float* x, float* out, int64_t n, float scale, const float* scale_dev, uint64_t seed,
                      uint64_t* state, void* stream);

/* ---- validation metrics (wind_field_GAN_3D.py:730-770, 597-618) ---------------------------------------------- */
/* One pass over HR (n,3,x,y,z), SR (same; optional) and LR (n,>=3,xl,yl,z):
 *   sums[0] = sum (HR-SR)^2, sums[1] = sum (HR-tri)^2, sums[2] = sum |HR-tri|, sums[3] = sum |HR-SR|
 * with tri = F.interpolate(LR[:, :3], scale_factor=(s,s,1), mode="trilinear", align_corners=True) evaluated on the
 * fly.  sums: device double[4] (zeroed by the call).  PSNR = 10 log10(4 / (sums[k]/(n*x*y*z) + 1e-8)). */
int ws_validation_metrics(const ws_tensor* hr, const ws_tensor* sr, const ws_tensor* lr, int n, int x, int y, int z,
                          int xl, int yl, double* sums, void* stream);

/* ---- input pipeline (process_data.py:159-262, 420-494) ---------------------------------------------------------- */
/* Builds a training batch on the device from the float64 per-hour fields of the on-disk format
 * (download_data.py:456-467: z, z_above_ground, u, v, w, pressure, each (sx, sy, sz)):
 * crop [x_start, x_start+x) x [y_start, y_start+y), normalise (reformat_to_torch), LR = every `coarseness`-th point in
 * x and y, then torch.rot90(k, [1,2]) and the two flips with the wind-component sign fixes of
 * CustomizedDataset.__getitem__.  Bit-exact against the reference's numpy/torch CPU code (float64 arithmetic, one
 * rounding to float32). */
typedef struct ws_prepare_desc {
  int32_t n, sx, sy, sz;   /* samples, source field extents                                       */
  int32_t x, y;            /* crop extents (slice_size, or sx / sy without slicing)               */
  int32_t coarseness;      /* cfg.scale                                                           */
  int32_t include_pressure, include_z_channel, include_above_ground_channel;
  int64_t sample_stride;   /* elements between consecutive samples in each source array           */
  double uvw_max, p_min, p_max, z_min, z_max, z_above_ground_max;
} ws_prepare_desc;
/* aug: device int32[n][5] = x_start, y_start, rotations (0..3), flip_x, flip_y.
 * lr: (n, C, ceil(x/c), ceil(y/c), sz), hr: (n, 3, x, y, sz), zout: (n, 1, x, y, sz), contiguous fp32. */
int ws_prepare_batch(const ws_prepare_desc* d, const double* u, const double* v, const double* w, const double* p,
                     const double* z, const double* zag, const int32_t* aug, float* lr, float* hr, float* zout,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WINDSR_H_ */
