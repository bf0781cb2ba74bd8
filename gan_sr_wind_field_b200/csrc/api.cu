// api.cu — the extern "C" boundary declared in include/windsr.h: argument validation, path selection
// (CUDA-core FFMA vs tcgen05) and launch.  No torch types, no allocation, no exceptions cross this file.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "common.cuh"
#include "tc_task.cuh"

namespace ws {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int validate_shape(const ws_conv_shape* s) {
  WS_REQUIRE(s != nullptr, "null conv shape");
  WS_REQUIRE(s->n > 0 && s->x > 0 && s->y > 0 && s->z > 0, "conv shape: non-positive volume");
  WS_REQUIRE(s->cin > 0 && s->cout > 0, "conv shape: non-positive channels");
  WS_REQUIRE(s->kx > 0 && s->ky > 0 && s->kz > 0 && s->sx > 0 && s->sy > 0 && s->sz > 0,
             "conv shape: bad kernel/stride");
  WS_REQUIRE(s->px >= 0 && s->py >= 0 && s->pz >= 0, "conv shape: negative padding");
  WS_REQUIRE(s->x + 2 * s->px >= s->kx && s->y + 2 * s->py >= s->ky && s->z + 2 * s->pz >= s->kz,
             "conv shape: kernel larger than padded input");
  return 0;
}

// implemented in the other translation units
int simt_conv_fwd(const ConvGeom&, const View&, const float*, const View&, const Epi&, cudaStream_t);
int simt_conv_dgrad(const ConvGeom&, const View&, const float*, const View&, const Epi&, cudaStream_t);
int simt_conv_wgrad(const ConvGeom&, const View&, const View&, float*, int, void*, size_t, cudaStream_t, bool);
size_t simt_wgrad_workspace_bytes(const ConvGeom&, bool);
int bias_grad(const View&, float*, int, int, long long, int, cudaStream_t);
bool tc_view_ok(const View&, int);
int tc_conv_launch(const ConvGeom&, int, const View&, const void*, const View&, const Epi&, cudaStream_t,
                   const TcOverride* ov = nullptr);
int tc2_conv_launch(const ConvGeom&, int, const View&, const void*, const View&, const Epi&, cudaStream_t,
                    const TcOverride* ov = nullptr, int dry = 0);
bool tc2_enabled();
int tc_conv_wgrad(const ConvGeom&, const View&, const View&, float*, int, void*, size_t, cudaStream_t);
size_t tc_wgrad_workspace_bytes(const ConvGeom&);
int pack_tc_batch_launch(int, const float* const*, const ConvGeom*, int, void* const*, cudaStream_t, const int*,
                         int tf32 = 0);
int im2col_small_launch(const View&, const View&, const ConvGeom&, int, cudaStream_t);
int xfold_sum_lrelu_launch(const View&, const View&, float, int, int, int, int, int, int, int, cudaStream_t);
int tc_rdb_wgrad(const ConvGeom&, const View&, const View&, float* const*, const int*, int, int, void*, size_t,
                 cudaStream_t);
size_t tc_rdb_wgrad_workspace_bytes(int, int, int, int);
size_t tc_trunk_wgrad_workspace_bytes(int, int, int, int, int);
int tc_trunk_wgrad(const ConvGeom&, const ConvGeom&, int, const View&, const View&, const View&, const int*, int, int,
                   float*, long long, void*, size_t, cudaStream_t);
int bias_grad_batched(const View&, float*, int, long long, int, int, long long, cudaStream_t);
int pack_tc_trunk_launch(int, int, const ConvGeom*, const int*, int, const float* const*, void* const*, cudaStream_t, int);
int pack_weights_launch(const float*, const ConvGeom&, int, void*, cudaStream_t);
bool rdb_persist_ok(const ws_rdb_desc*, const View&, const View&, int);
int rdb_persist_forward(const ws_rdb_desc*, const View&, const View&, const View&, void* const*, const Epi&,
                        cudaStream_t);
bool rdb_persist_bwd_ok(const ws_rdb_desc*, const View&, const View&, const View&, const View&, const View&, int);
int rdb_persist_backward(const ws_rdb_desc*, const View&, const View&, const View&, const View&, const View&,
                         void* const*, cudaStream_t);
int copy_launch(const View&, const View&, int, int, long long, cudaStream_t);
int round_tf32_launch(const View&, int, int, long long, cudaStream_t);
int axpby_launch(const View&, float, const View&, float, const View&, int, int, long long, cudaStream_t);
int lrelu_bwd_launch(const View&, const View&, float, const float*, const float*, const View&, int, int,
                     long long, cudaStream_t);
int upsample_fwd_launch(const View&, const View&, int, int, int, int, int, cudaStream_t);
int upsample_bwd_launch(const View&, const View&, int, int, int, int, int, cudaStream_t);
int bn_finalize_launch(const float*, const float*, long long, int, const float*, const float*, float, float,
                       float*, float*, float*, float*, float*, float*, cudaStream_t);
int scale_shift_lrelu_launch(const View&, const float*, const float*, float, const View&, int, int, long long,
                             cudaStream_t);
int bn_bwd_reduce_launch(const View&, const View&, const View&, const float*, const float*, float, float*,
                         float*, int, int, long long, cudaStream_t);
int bn_bwd_apply_launch(const View&, const View&, const View&, const float*, const float*, const float*,
                        const float*, const float*, float, long long, const View&, int, int, long long,
                        cudaStream_t);
int xfold_sum_launch(const View&, const float*, const View&, int, int, int, int, int, int, int, cudaStream_t);
int xunfold_launch(const View&, const View&, int, int, int, int, int, int, int, int, cudaStream_t);
int xyfold_sum_launch(const View&, const float*, const View&, int, int, int, int, int, int, int, int, int, cudaStream_t);
int xyunfold_launch(const View&, const View&, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int axis_coeffs_launch(const float*, int, float*, cudaStream_t);
int wind_gradient_launch(const View&, const View&, const float*, const float*, const View&, int, int, int, int,
                         cudaStream_t);
int windloss_fwd_launch(const View&, const View&, const View&, const float*, const float*, int, int, int, int,
                        float*, long long*, cudaStream_t);
int windloss_bwd_launch(const View&, const View&, const View&, const float*, const float*, int, int, int, int,
                        const float*, const long long*, const View&, void*, size_t, cudaStream_t);

static int device_sm_count() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  }
  return sms;
}

static int device_cc_major() {
  static int major = -1;
  if (major < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    major = prop.major;
  }
  return major;
}

// Bring-up / bisection switches (read once): WS_DISABLE_TC=1 forces the CUDA-core family everywhere,
// WS_DISABLE_TC_{FWD,DGRAD,WGRAD,STRIDED}=1 do so for one kernel class.
static bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}
static bool tc_disabled(const char* which) {
  static const bool all = env_flag("WS_DISABLE_TC");
  static const bool fwd = env_flag("WS_DISABLE_TC_FWD"), dgrad = env_flag("WS_DISABLE_TC_DGRAD");
  static const bool wgrad = env_flag("WS_DISABLE_TC_WGRAD"), strided = env_flag("WS_DISABLE_TC_STRIDED");
  if (all) return true;
  switch (which[0]) {
    case 'f': return fwd;
    case 'd': return dgrad;
    case 'w': return wgrad;
    case 's': return strided;
  }
  return false;
}
static bool is_strided(const ConvGeom& g) { return g.sx != 1 || g.sy != 1 || g.sz != 1; }

// TF32 mode (tcgen05 kind::tf32 on fp32 channels-last activations): the halo-tile kernel for stride-1 forward /
// data-gradient convs with at least 8 reduction channels, the MN-major GEMM for every weight gradient; the strided
// discriminator convs and the narrow first layers stay on the fp32 CUDA-core family (more accurate, and a few
// percent of the FLOPs).
static bool tf32_view_ok(const View& v, int channels) {
  if (v.dtype != WS_F32 || v.cs != 1) return false;
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) != 0) return false;
  if ((v.vs * 4) % 16 != 0 || (v.ns * 4) % 16 != 0) return false;
  return channels >= 8;
}
static int tf32_conv_path(const ConvGeom& g, int mode, const View& src, const char* which) {
  if (tc_disabled(which) || device_cc_major() != 10 || is_strided(g) || !tc2_enabled()) return WS_PATH_SIMT;
  if (!tf32_view_ok(src, mode == 0 ? g.cin : g.cout)) return WS_PATH_SIMT;
  Epi none(nullptr, mode == 0 ? g.cout : g.cin);
  return tc2_conv_launch(g, mode, src, nullptr, View(), none, nullptr, nullptr, 1) == 0 ? WS_PATH_TCGEN05
                                                                                       : WS_PATH_SIMT;
}

static int fwd_path(const ConvGeom& g, const View& in, int math) {
  if (math == WS_MATH_TF32) return tf32_conv_path(g, 0, in, "fwd");
  if (math != WS_MATH_BF16) return WS_PATH_SIMT;
  if (tc_disabled("fwd") || (is_strided(g) && tc_disabled("strided"))) return WS_PATH_SIMT;
  if (device_cc_major() != 10) return WS_PATH_SIMT;
  if (!tc_view_ok(in, g.cin)) return WS_PATH_SIMT;
  if (g.zo > 128) return WS_PATH_SIMT;
  return WS_PATH_TCGEN05;
}
static int dgrad_path(const ConvGeom& g, const View& dy, int math) {
  if (math == WS_MATH_TF32) return tf32_conv_path(g, 1, dy, "dgrad");
  if (math != WS_MATH_BF16) return WS_PATH_SIMT;
  if (tc_disabled("dgrad") || (is_strided(g) && tc_disabled("strided"))) return WS_PATH_SIMT;
  if (device_cc_major() != 10) return WS_PATH_SIMT;
  if (!tc_view_ok(dy, g.cout)) return WS_PATH_SIMT;
  if (g.z > 128) return WS_PATH_SIMT;
  return WS_PATH_TCGEN05;
}
static int wgrad_path(const ConvGeom& g, const View& in, const View& dy, int math) {
  if (math == WS_MATH_TF32) {
    if (tc_disabled("wgrad") || (is_strided(g) && tc_disabled("strided")) || device_cc_major() != 10)
      return WS_PATH_SIMT;
    if (!tf32_view_ok(in, g.cin) || !tf32_view_ok(dy, g.cout) || g.cin < 16 || g.cout < 16) return WS_PATH_SIMT;
    if (g.cin > 256 && g.cout > 256) return WS_PATH_SIMT;
    return WS_PATH_TCGEN05;
  }
  if (math != WS_MATH_BF16) return WS_PATH_SIMT;
  if (tc_disabled("wgrad") || (is_strided(g) && tc_disabled("strided"))) return WS_PATH_SIMT;
  if (device_cc_major() != 10) return WS_PATH_SIMT;
  if (!tc_view_ok(in, g.cin) || !tc_view_ok(dy, g.cout)) return WS_PATH_SIMT;
  if (g.cin > 256 && g.cout > 256) return WS_PATH_SIMT;
  return WS_PATH_TCGEN05;
}

// Data-gradient of a STRIDED conv on the tensor cores: the input grid splits into sx*sy*sz parity classes; within
// a class the gradient is a stride-1 correlation of dy with the sub-set of taps of matching parity, written to the
// (d*s + parity) voxels of dx.  (Replaces the dgrad of the (4,4,3)/(2,2,1|2) discriminator convs,
// torch_blocks.py:467-506, which the CUDA-core family computed by testing every tap for divisibility.)
static int tc_strided_dgrad(const ConvGeom& g, const View& dy, const void* packed_w, const View& dx, const Epi& ep,
                            cudaStream_t st) {
  const int rows = (g.cin + 15) / 16 * 16, cols = (g.cout + 7) / 8 * 8;
  (void)rows; (void)cols;
  int tap_base = 0;
  for (int pa = 0; pa < g.sx; ++pa)
    for (int pb = 0; pb < g.sy; ++pb)
      for (int pc = 0; pc < g.sz; ++pc) {
        const int par[3] = {pa, pb, pc}, k[3] = {g.kx, g.ky, g.kz}, s[3] = {g.sx, g.sy, g.sz};
        const int pad[3] = {g.px, g.py, g.pz}, ext[3] = {g.x, g.y, g.z};
        int nt[3], pp[3], dd[3];
        bool empty = false;
        for (int a = 0; a < 3; ++a) {
          const int i0 = (par[a] + pad[a]) % s[a];
          nt[a] = i0 < k[a] ? (k[a] - i0 + s[a] - 1) / s[a] : 0;
          const int s0 = (par[a] + pad[a] - i0) / s[a];
          pp[a] = nt[a] - 1 - s0;
          dd[a] = (ext[a] - par[a] + s[a] - 1) / s[a];
          if (nt[a] == 0 || dd[a] <= 0) empty = true;
        }
        const int class_taps = nt[0] * nt[1] * nt[2];
        if (dd[0] <= 0 || dd[1] <= 0 || dd[2] <= 0) { tap_base += class_taps; continue; }
        WS_REQUIRE(!empty, "strided dgrad: a parity class without taps is not supported on the tcgen05 path");
        TcOverride ov;
        ov.DX = dd[0]; ov.DY = dd[1]; ov.DZ = dd[2];
        ov.SX = g.xo; ov.SY = g.yo; ov.SZ = g.zo;
        ov.kx = nt[0]; ov.ky = nt[1]; ov.kz = nt[2];
        ov.px = pp[0]; ov.py = pp[1]; ov.pz = pp[2];
        ov.ck = g.cout; ov.cn = g.cin;
        ov.tap_base = tap_base; ov.taps_total = g.taps();
        ov.omx = g.sx; ov.oax = pa; ov.omy = g.sy; ov.oay = pb; ov.omz = g.sz; ov.oaz = pc;
        ov.ODY = g.y; ov.ODZ = g.z;
        int r = tc2_enabled() ? tc2_conv_launch(g, 1, dy, packed_w, dx, ep, st, &ov) : -1;
        if (r < 0) r = tc_conv_launch(g, 1, dy, packed_w, dx, ep, st, &ov);
        if (r) return r;
        tap_base += class_taps;
      }
  return 0;
}

}  // namespace ws

using namespace ws;

extern "C" {

int ws_version(void) { return WS_VERSION; }
const char* ws_last_error(void) { return g_err; }
int ws_device_supports_tcgen05(void) { return device_cc_major() == 10 ? 1 : 0; }
int64_t ws_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

size_t ws_packed_weight_bytes(const ws_conv_shape* s, int kind) {
  if (!s) return 0;
  size_t taps = (size_t)s->kx * s->ky * s->kz;
  switch (kind) {
    case WS_PACK_SIMT_FWD:
    case WS_PACK_SIMT_DGRAD:
      return taps * s->cin * s->cout * sizeof(float);
    case WS_PACK_TC_FWD:
      return taps * (size_t)((s->cout + 15) / 16 * 16) * (size_t)((s->cin + 7) / 8 * 8) * 2;
    case WS_PACK_TC_DGRAD:
      return taps * (size_t)((s->cin + 15) / 16 * 16) * (size_t)((s->cout + 7) / 8 * 8) * 2;
    case WS_PACK_TC_FWD_TF32:
      return taps * (size_t)((s->cout + 15) / 16 * 16) * (size_t)((s->cin + 3) / 4 * 4) * 4;
    case WS_PACK_TC_DGRAD_TF32:
      return taps * (size_t)((s->cin + 15) / 16 * 16) * (size_t)((s->cout + 3) / 4 * 4) * 4;
  }
  return 0;
}

int ws_pack_weights(const float* w, const ws_conv_shape* s, int kind, void* packed, void* stream) {
  if (int e = validate_shape(s)) return e;
  WS_REQUIRE(w && packed, "ws_pack_weights: null pointer");
  WS_REQUIRE(kind >= WS_PACK_SIMT_FWD && kind <= WS_PACK_TC_DGRAD_TF32, "ws_pack_weights: bad kind %d", kind);
  WS_REQUIRE(kind < WS_PACK_TC_FWD_TF32 || (s->sx == 1 && s->sy == 1 && s->sz == 1),
             "ws_pack_weights: the TF32 tensor-core packings are stride-1 only");
  return pack_weights_launch(w, ConvGeom(*s), kind, packed, (cudaStream_t)stream);
}

int ws_conv3d_fwd_path(const ws_conv_shape* s, const ws_tensor* in, const ws_tensor* out, int math) {
  if (validate_shape(s) || !in) return WS_PATH_NONE;
  (void)out;
  return fwd_path(ConvGeom(*s), View(in), math);
}
int ws_conv3d_dgrad_path(const ws_conv_shape* s, const ws_tensor* dy, const ws_tensor* dx, int math) {
  if (validate_shape(s) || !dy) return WS_PATH_NONE;
  (void)dx;
  return dgrad_path(ConvGeom(*s), View(dy), math);
}
int ws_conv3d_wgrad_path(const ws_conv_shape* s, const ws_tensor* in, const ws_tensor* dy, int math) {
  if (validate_shape(s) || !in || !dy) return WS_PATH_NONE;
  return wgrad_path(ConvGeom(*s), View(in), View(dy), math);
}

int ws_conv3d_fwd(const ws_conv_shape* s, const ws_tensor* in, const void* packed_w, const ws_tensor* out,
                  const ws_epilogue* ep, int math, void* stream) {
  if (int e = validate_shape(s)) return e;
  WS_REQUIRE(in && in->ptr && out && out->ptr && packed_w, "ws_conv3d_fwd: null pointer");
  ConvGeom g(*s);
  View vin(in), vout(out);
  Epi e(ep, g.cout);
  if (fwd_path(g, vin, math) == WS_PATH_TCGEN05) {
    if (tc2_enabled()) {
      int r = tc2_conv_launch(g, 0, vin, packed_w, vout, e, (cudaStream_t)stream);
      if (r >= 0) return r;  // -1: geometry not covered by the halo-tile kernel
    }
    WS_REQUIRE(math == WS_MATH_BF16, "ws_conv3d_fwd: geometry not covered by the TF32 tensor-core kernel");
    return tc_conv_launch(g, 0, vin, packed_w, vout, e, (cudaStream_t)stream);
  }
  return simt_conv_fwd(g, vin, (const float*)packed_w, vout, e, (cudaStream_t)stream);
}

int ws_conv3d_dgrad(const ws_conv_shape* s, const ws_tensor* dy, const void* packed_w, const ws_tensor* dx,
                    const ws_epilogue* ep, int math, void* stream) {
  if (int e = validate_shape(s)) return e;
  WS_REQUIRE(dy && dy->ptr && dx && dx->ptr && packed_w, "ws_conv3d_dgrad: null pointer");
  ConvGeom g(*s);
  View vdy(dy), vdx(dx);
  Epi e(ep, g.cin);
  if (dgrad_path(g, vdy, math) == WS_PATH_TCGEN05) {
    if (is_strided(g)) return tc_strided_dgrad(g, vdy, packed_w, vdx, e, (cudaStream_t)stream);
    if (tc2_enabled()) {
      int r = tc2_conv_launch(g, 1, vdy, packed_w, vdx, e, (cudaStream_t)stream);
      if (r >= 0) return r;
    }
    WS_REQUIRE(math == WS_MATH_BF16, "ws_conv3d_dgrad: geometry not covered by the TF32 tensor-core kernel");
    return tc_conv_launch(g, 1, vdy, packed_w, vdx, e, (cudaStream_t)stream);
  }
  return simt_conv_dgrad(g, vdy, (const float*)packed_w, vdx, e, (cudaStream_t)stream);
}

size_t ws_conv3d_wgrad_workspace_bytes(const ws_conv_shape* s, int math) {
  if (!s || validate_shape(s)) return 0;
  // FP32 mode: deterministic split-K (one accumulator array per split, summed in order)
  if (math == WS_MATH_FP32) return simt_wgrad_workspace_bytes(ConvGeom(*s), true);
  ws_conv_shape t = *s;
  return (size_t)t.kx * t.ky * t.kz * t.cin * t.cout * sizeof(float);
}

int ws_conv3d_wgrad(const ws_conv_shape* s, const ws_tensor* in, const ws_tensor* dy, float* dw, float* db,
                    int accumulate, int math, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = validate_shape(s)) return e;
  WS_REQUIRE(in && in->ptr && dy && dy->ptr, "ws_conv3d_wgrad: null pointer");
  ConvGeom g(*s);
  View vin(in), vdy(dy);
  cudaStream_t st = (cudaStream_t)stream;
  if (db) {
    if (int e = bias_grad(vdy, db, g.n, g.cout, g.vout(), accumulate, st)) return e;
  }
  if (!dw) return 0;
  if (wgrad_path(g, vin, vdy, math) == WS_PATH_TCGEN05)
    return tc_conv_wgrad(g, vin, vdy, dw, accumulate, workspace, workspace_bytes, st);
  return simt_conv_wgrad(g, vin, vdy, dw, accumulate, workspace, workspace_bytes, st, math == WS_MATH_FP32);
}

int ws_upsample_nearest_xy_fwd(const ws_tensor* in, const ws_tensor* out, int n, int c, int x, int y, int z,
                               void* stream) {
  WS_REQUIRE(in && in->ptr && out && out->ptr, "ws_upsample_nearest_xy_fwd: null pointer");
  return upsample_fwd_launch(View(in), View(out), n, c, x, y, z, (cudaStream_t)stream);
}
int ws_upsample_nearest_xy_bwd(const ws_tensor* dout, const ws_tensor* din, int n, int c, int x, int y, int z,
                               void* stream) {
  WS_REQUIRE(dout && dout->ptr && din && din->ptr, "ws_upsample_nearest_xy_bwd: null pointer");
  return upsample_bwd_launch(View(dout), View(din), n, c, x, y, z, (cudaStream_t)stream);
}

int ws_xfold_sum(const ws_tensor* y, const float* bias, const ws_tensor* out, int n, int co, int kx, int pad,
                 int x, int yy, int z, void* stream) {
  WS_REQUIRE(y && y->ptr && out && out->ptr && co > 0 && kx > 0, "ws_xfold_sum: bad arguments");
  return xfold_sum_launch(View(y), bias, View(out), n, co, kx, pad, x, yy, z, (cudaStream_t)stream);
}
int ws_xunfold(const ws_tensor* dout, const ws_tensor* u, int n, int co, int kx, int pad, int cpad, int x, int yy,
               int z, void* stream) {
  WS_REQUIRE(dout && dout->ptr && u && u->ptr && cpad >= kx * co, "ws_xunfold: bad arguments");
  return xunfold_launch(View(dout), View(u), n, co, kx, pad, cpad, x, yy, z, (cudaStream_t)stream);
}

int ws_xyfold_sum(const ws_tensor* y, const float* bias, const ws_tensor* out, int n, int co, int kx, int ky, int px,
                  int py, int x, int yy, int z, void* stream) {
  WS_REQUIRE(y && y->ptr && out && out->ptr && co > 0 && kx > 0 && ky > 0, "ws_xyfold_sum: bad arguments");
  return xyfold_sum_launch(View(y), bias, View(out), n, co, kx, ky, px, py, x, yy, z, (cudaStream_t)stream);
}
int ws_xyunfold(const ws_tensor* dout, const ws_tensor* u, int n, int co, int kx, int ky, int px, int py, int cpad,
                int x, int yy, int z, void* stream) {
  WS_REQUIRE(dout && dout->ptr && u && u->ptr && cpad >= kx * ky * co, "ws_xyunfold: bad arguments");
  return xyunfold_launch(View(dout), View(u), n, co, kx, ky, px, py, cpad, x, yy, z, (cudaStream_t)stream);
}

int ws_im2col(const ws_conv_shape* s, const ws_tensor* x, const ws_tensor* u, int cpad, void* stream) {
  if (int e = validate_shape(s)) return e;
  WS_REQUIRE(x && x->ptr && u && u->ptr, "ws_im2col: null pointer");
  return im2col_small_launch(View(x), View(u), ConvGeom(*s), cpad, (cudaStream_t)stream);
}

int ws_copy(const ws_tensor* src, const ws_tensor* dst, int n, int c, int64_t v, void* stream) {
  WS_REQUIRE(src && src->ptr && dst && dst->ptr, "ws_copy: null pointer");
  return copy_launch(View(src), View(dst), n, c, v, (cudaStream_t)stream);
}
int ws_axpby(const ws_tensor* x1, float a, const ws_tensor* x2, float b, const ws_tensor* y, int n, int c,
             int64_t v, void* stream) {
  WS_REQUIRE(x1 && x1->ptr && y && y->ptr, "ws_axpby: null pointer");
  return axpby_launch(View(x1), a, View(x2), b, View(y), n, c, v, (cudaStream_t)stream);
}
int ws_lrelu_bwd(const ws_tensor* dy, const ws_tensor* y, float slope, const float* chan_scale,
                 const float* oscale, const ws_tensor* g, int n, int c, int64_t v, void* stream) {
  WS_REQUIRE(dy && dy->ptr && y && y->ptr && g && g->ptr, "ws_lrelu_bwd: null pointer");
  return lrelu_bwd_launch(View(dy), View(y), slope, chan_scale, oscale, View(g), n, c, v, (cudaStream_t)stream);
}

int ws_bn_finalize(const float* sum, const float* sqsum, int64_t count, int c, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  WS_REQUIRE(sum && sqsum && scale && shift && c > 0 && count > 0, "ws_bn_finalize: bad arguments");
  return bn_finalize_launch(sum, sqsum, count, c, gamma, beta, eps, momentum, running_mean, running_var, scale,
                            shift, save_mean, save_invstd, (cudaStream_t)stream);
}
int ws_scale_shift_lrelu(const ws_tensor* x, const float* scale, const float* shift, float slope,
                         const ws_tensor* y, int n, int c, int64_t v, void* stream) {
  WS_REQUIRE(x && x->ptr && y && y->ptr && scale && shift, "ws_scale_shift_lrelu: null pointer");
  return scale_shift_lrelu_launch(View(x), scale, shift, slope, View(y), n, c, v, (cudaStream_t)stream);
}
int ws_bn_lrelu_bwd_reduce(const ws_tensor* dy, const ws_tensor* y, const ws_tensor* x, const float* mean,
                           const float* invstd, float slope, float* sum_g, float* sum_gx, int n, int c,
                           int64_t v, void* stream) {
  WS_REQUIRE(dy && y && x && mean && invstd && sum_g && sum_gx, "ws_bn_lrelu_bwd_reduce: null pointer");
  return bn_bwd_reduce_launch(View(dy), View(y), View(x), mean, invstd, slope, sum_g, sum_gx, n, c, v,
                              (cudaStream_t)stream);
}
int ws_bn_lrelu_bwd_apply(const ws_tensor* dy, const ws_tensor* y, const ws_tensor* x, const float* mean,
                          const float* invstd, const float* gamma, const float* sum_g, const float* sum_gx,
                          float slope, int64_t count, const ws_tensor* dx, int n, int c, int64_t v,
                          void* stream) {
  WS_REQUIRE(dy && y && x && dx && mean && invstd && sum_g && sum_gx, "ws_bn_lrelu_bwd_apply: null pointer");
  return bn_bwd_apply_launch(View(dy), View(y), View(x), mean, invstd, gamma, sum_g, sum_gx, slope, count,
                             View(dx), n, c, v, (cudaStream_t)stream);
}

int ws_axis_coeffs(const float* coords, int len, float* coef, void* stream) {
  WS_REQUIRE(coords && coef && len > 0, "ws_axis_coeffs: bad arguments");
  return axis_coeffs_launch(coords, len, coef, (cudaStream_t)stream);
}
int ws_wind_gradient(const ws_tensor* field, const ws_tensor* zalt, const float* coef_x, const float* coef_y,
                     const ws_tensor* out, int n, int x, int y, int z, void* stream) {
  WS_REQUIRE(field && field->ptr && zalt && zalt->ptr && out && out->ptr && coef_x && coef_y,
             "ws_wind_gradient: null pointer");
  return wind_gradient_launch(View(field), View(zalt), coef_x, coef_y, View(out), n, x, y, z,
                              (cudaStream_t)stream);
}
int ws_windloss_fwd(const ws_tensor* hr, const ws_tensor* sr, const ws_tensor* zalt, const float* coef_x,
                    const float* coef_y, int n, int x, int y, int z, float* result, int64_t* argmax,
                    void* stream) {
  WS_REQUIRE(hr && hr->ptr && sr && sr->ptr && zalt && zalt->ptr && coef_x && coef_y && result,
             "ws_windloss_fwd: null pointer");
  WS_REQUIRE((reinterpret_cast<uintptr_t>(result) & 7) == 0, "ws_windloss_fwd: result must be 8-byte aligned");
  return windloss_fwd_launch(View(hr), View(sr), View(zalt), coef_x, coef_y, n, x, y, z, result,
                             (long long*)argmax, (cudaStream_t)stream);
}
size_t ws_windloss_bwd_workspace_bytes(int n, int x, int y, int z) {
  return (size_t)n * x * y * z * 9 * sizeof(float);
}
int ws_windloss_bwd(const ws_tensor* hr, const ws_tensor* sr, const ws_tensor* zalt, const float* coef_x,
                    const float* coef_y, int n, int x, int y, int z, const float* coef, const int64_t* argmax,
                    const ws_tensor* dsr, void* workspace, size_t workspace_bytes, void* stream) {
  WS_REQUIRE(hr && hr->ptr && sr && sr->ptr && zalt && zalt->ptr && coef_x && coef_y && coef && dsr && dsr->ptr,
             "ws_windloss_bwd: null pointer");
  return windloss_bwd_launch(View(hr), View(sr), View(zalt), coef_x, coef_y, n, x, y, z, coef,
                             (const long long*)argmax, View(dsr), workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

}  // extern "C"

// ---- residual dense block executor ---------------------------------------------------------------------------
namespace {
struct RdbGeom {
  ws_conv_shape dense[WS_RDB_MAX_CONVS];
  ws_conv_shape lff;
  int ctot;
};
int rdb_geom(const ws_rdb_desc* d, RdbGeom& r) {
  WS_REQUIRE(d && d->nconv >= 0 && d->nconv <= WS_RDB_MAX_CONVS, "rdb: bad descriptor");
  WS_REQUIRE(d->k % 2 == 1 && d->k_lff % 2 == 1, "rdb: kernel sizes must be odd");
  for (int i = 0; i < d->nconv; ++i) {
    ws_conv_shape s = {d->n, d->x, d->y, d->z, d->f + i * d->gc, d->gc, d->k, d->k, d->k, 1, 1, 1,
                       (d->k - 1) / 2, (d->k - 1) / 2, (d->k - 1) / 2};
    r.dense[i] = s;
  }
  r.ctot = d->f + d->nconv * d->gc;
  ws_conv_shape l = {d->n, d->x, d->y, d->z, r.ctot, d->f, d->k_lff, d->k_lff, d->k_lff, 1, 1, 1,
                     (d->k_lff - 1) / 2, (d->k_lff - 1) / 2, (d->k_lff - 1) / 2};
  r.lff = l;
  return 0;
}
ws_tensor slice(const ws_tensor& t, int c0) {
  ws_tensor v = t;
  v.ptr = (char*)t.ptr + (size_t)c0 * t.cstride * (t.dtype == WS_F32 ? 4 : 2);
  return v;
}
ws_epilogue plain_epilogue() {
  ws_epilogue e;
  memset(&e, 0, sizeof(e));
  e.lrelu_slope = 1.f; e.alpha = 1.f; e.mask_slope = 1.f;
  return e;
}
}  // namespace

namespace {
// Pseudo conv of the merged dense-conv weight gradient: M = buffer channels read by the widest dense conv,
// N = all nconv*gc gradient channels (wgrad_tc.cu, tc_rdb_wgrad).
ws_conv_shape rdb_merged_shape(const ws_rdb_desc* d) {
  ws_conv_shape s = {d->n, d->x, d->y, d->z, d->f + (d->nconv - 1) * d->gc, d->nconv * d->gc, d->k, d->k, d->k, 1, 1,
                     1, (d->k - 1) / 2, (d->k - 1) / 2, (d->k - 1) / 2};
  return s;
}
// (Re)pack all weights of the block for one direction: the tensor-core packings go out as ONE launch.
// act: the operand view whose layout decides the path of the dense convs' / LFF's input (forward: the concat buffer;
// backward: g_lff for the LFF and a gc-channel slice of gbuf for the dense convs).
int rdb_repack(const ws_rdb_desc* d, const RdbGeom& r, const ws_tensor* lff_act, const ws_tensor* dense_act,
               const float* const* w, void* const* packed, int dgrad, cudaStream_t st, int fold_fwd = 0) {
  const float* bw[WS_RDB_MAX_CONVS + 1];
  void* bp[WS_RDB_MAX_CONVS + 1];
  ConvGeom bg[WS_RDB_MAX_CONVS + 1] = {};
  int bf[WS_RDB_MAX_CONVS + 1] = {};
  int nb = 0;
  for (int i = 0; i <= d->nconv; ++i) {
    const ws_conv_shape* s = i < d->nconv ? &r.dense[i] : &r.lff;
    ConvGeom g(*s);
    bool tc;
    if (!dgrad) {
      tc = fwd_path(g, View(lff_act), d->math) == WS_PATH_TCGEN05;  // every conv reads the concat buffer
    } else if (i < d->nconv) {
      ws_tensor gs = slice(*dense_act, i * d->gc);
      tc = dgrad_path(g, View(&gs), d->math) == WS_PATH_TCGEN05;
    } else {
      tc = dgrad_path(g, View(lff_act), d->math) == WS_PATH_TCGEN05;
    }
    if (tc) {
      bw[nb] = w[i]; bp[nb] = packed[i]; bg[nb] = g; bf[nb] = (fold_fwd && !dgrad && i < d->nconv) ? fold_fwd : 0; ++nb;
    } else if (int e = pack_weights_launch(w[i], g, dgrad ? WS_PACK_SIMT_DGRAD : WS_PACK_SIMT_FWD, packed[i], st)) {
      return e;
    }
  }
  return pack_tc_batch_launch(nb, bw, bg, dgrad, bp, st, bf, d->math == WS_MATH_TF32 ? 1 : 0);
}
// x-folded forward of the dense convs: the kx taps go side by side on the UMMA N dimension (N = kx*gc = 96 instead of
// 32 — an MMA costs the same 72 cycles either way), the conv runs as a (1,k,k) conv into an fp32 scratch U and a small
// vector kernel forms y[x] = lrelu(sum_dx U[x+dx-p][dx]) into the concat-buffer slice.
ws_conv_shape rdb_fold_shape(const ws_rdb_desc* d, int i) {
  ws_conv_shape s = {d->n, d->x, d->y, d->z, d->f + i * d->gc, d->k * d->gc, 1, d->k, d->k, 1, 1, 1,
                     0, (d->k - 1) / 2, (d->k - 1) / 2};
  return s;
}
bool rdb_fold_ok(const ws_rdb_desc* d, const ws_tensor* buf) {
  if (d->nconv < 1 || d->k < 2 || d->k * d->gc > 256 || d->gc % 8 != 0 || d->math != WS_MATH_BF16 ||
      getenv("WS_DISABLE_RDB_XFOLD"))
    return false;
  for (int i = 0; i < d->nconv; ++i) {
    ws_conv_shape s = rdb_fold_shape(d, i);
    if (fwd_path(ConvGeom(s), View(buf), d->math) != WS_PATH_TCGEN05) return false;
  }
  return true;
}
bool rdb_merged_wgrad_ok(const ws_rdb_desc* d, const ws_tensor* buf, const ws_tensor* gbuf) {
  if (d->math != WS_MATH_BF16) return false;  // the merged GEMM is a bf16 kernel
  if (d->nconv < 2 || d->nconv * d->gc > 256 || d->gc % 16 != 0 || getenv("WS_DISABLE_RDB_MERGED_WGRAD")) return false;
  ws_conv_shape s = rdb_merged_shape(d);
  return wgrad_path(ConvGeom(s), View(buf), View(gbuf), d->math) == WS_PATH_TCGEN05;
}
}  // namespace

extern "C" size_t ws_rdb_forward_workspace_bytes(const ws_rdb_desc* d) {
  if (!d || d->nconv < 1) return 0;
  return (size_t)d->n * d->x * d->y * d->z * d->k * d->gc * sizeof(float);
}

extern "C" size_t ws_rdb_backward_workspace_bytes(const ws_rdb_desc* d) {
  RdbGeom r;
  if (!d || rdb_geom(d, r)) return 0;
  size_t need = 0;
  for (int i = 0; i <= d->nconv; ++i) {
    size_t b = ws_conv3d_wgrad_workspace_bytes(i < d->nconv ? &r.dense[i] : &r.lff, d->math);
    if (b > need) need = b;
  }
  if (d->nconv >= 2 && d->math == WS_MATH_BF16) {
    size_t b = tc_rdb_wgrad_workspace_bytes(d->k * d->k * d->k, d->f + (d->nconv - 1) * d->gc, d->nconv, d->gc);
    if (b > need) need = b;
  }
  return need;
}

extern "C" size_t ws_rdb_packed_bytes(const ws_rdb_desc* d, int i, int dgrad) {
  RdbGeom r;
  if (!d || rdb_geom(d, r) || i < 0 || i > d->nconv) return 0;
  const ws_conv_shape* s = i < d->nconv ? &r.dense[i] : &r.lff;
  size_t a = ws_packed_weight_bytes(s, dgrad ? WS_PACK_SIMT_DGRAD : WS_PACK_SIMT_FWD);
  size_t b = ws_packed_weight_bytes(s, dgrad ? WS_PACK_TC_DGRAD : WS_PACK_TC_FWD);
  size_t c = ws_packed_weight_bytes(s, dgrad ? WS_PACK_TC_DGRAD_TF32 : WS_PACK_TC_FWD_TF32);
  if (c > b) b = c;
  return a > b ? a : b;
}

extern "C" int ws_rdb_forward(const ws_rdb_desc* d, const ws_tensor* x, const ws_tensor* outer,
                              const ws_tensor* buf, const ws_tensor* out, const float* const* w,
                              void* const* packed, const float* lff_bias, void* workspace, size_t workspace_bytes,
                              void* stream) {
  RdbGeom r;
  if (int e = rdb_geom(d, r)) return e;
  WS_REQUIRE(x && x->ptr && buf && buf->ptr && out && out->ptr && w && packed, "ws_rdb_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long v = (long long)d->x * d->y * d->z;
  // (the coalesced fp32 row epilogue of the persistent kernel walks x / outer / out with one row stride)
  const bool rows_ok = out->dtype == WS_F32 && out->cstride == 1 && x->vstride == out->vstride &&
                       x->nstride == out->nstride && !(reinterpret_cast<uintptr_t>(out->ptr) & 15) &&
                       (!outer || !outer->ptr ||
                        (outer->dtype == WS_F32 && outer->cstride == 1 && outer->vstride == out->vstride &&
                         outer->nstride == out->nstride && !(reinterpret_cast<uintptr_t>(outer->ptr) & 15)));
  if (device_cc_major() == 10 && rows_ok && rdb_persist_ok(d, View(x), View(buf), device_sm_count()) &&
      fwd_path(ConvGeom(r.lff), View(buf), d->math) == WS_PATH_TCGEN05) {
    // the whole block as one persistent cooperative kernel (rdb_persist.cu): z-folded dense-conv weights
    if (d->repack)
      if (int e = rdb_repack(d, r, buf, nullptr, w, packed, 0, st, 2)) return e;
    ws_epilogue ep = plain_epilogue();
    ep.bias = lff_bias;
    ep.alpha = d->alpha;
    ep.res1 = *x; ep.beta1 = d->beta1;
    if (outer && outer->ptr) { ep.res2 = *outer; ep.beta2 = d->beta2; }
    return rdb_persist_forward(d, View(x), View(buf), View(out), packed, Epi(&ep, d->f), st);
  }
  // buf[:, :f] = x (cast to the activation dtype)
  ws_tensor b0 = *buf;
  if (int e = copy_launch(View(x), View(&b0), d->n, d->f, v, st)) return e;
  if (d->math == WS_MATH_TF32)  // the conv-operand copy of the block input (the fp32 residual stream stays exact)
    if (int e = round_tf32_launch(View(&b0), d->n, d->f, v, st)) return e;
  const bool fold = workspace && workspace_bytes >= ws_rdb_forward_workspace_bytes(d) && rdb_fold_ok(d, buf);
  if (d->repack)
    if (int e = rdb_repack(d, r, buf, nullptr, w, packed, 0, st, fold ? 1 : 0)) return e;
  for (int i = 0; i < d->nconv; ++i) {
    const ws_conv_shape* s = &r.dense[i];
    ws_tensor in = *buf, o = slice(*buf, s->cin);
    if (fold) {
      ws_conv_shape fs = rdb_fold_shape(d, i);
      ws_tensor u = {workspace, WS_F32, 0, v * fs.cout, fs.cout, 1};
      ws_epilogue ep = plain_epilogue();
      if (int e = ws_conv3d_fwd(&fs, &in, packed[i], &u, &ep, d->math, stream)) return e;
      if (int e = xfold_sum_lrelu_launch(View(&u), View(&o), d->slope, d->n, d->gc, d->k, (d->k - 1) / 2, d->x, d->y,
                                         d->z, st))
        return e;
      continue;
    }
    ws_epilogue ep = plain_epilogue();
    ep.lrelu_slope = d->slope;
    ep.flags = d->math == WS_MATH_TF32 ? 1 : 0;
    if (int e = ws_conv3d_fwd(s, &in, packed[i], &o, &ep, d->math, stream)) return e;
  }
  {
    ws_epilogue ep = plain_epilogue();
    ep.bias = lff_bias;
    ep.alpha = d->alpha;
    ep.res1 = *x; ep.beta1 = d->beta1;
    if (outer && outer->ptr) { ep.res2 = *outer; ep.beta2 = d->beta2; }
    if (int e = ws_conv3d_fwd(&r.lff, buf, packed[d->nconv], out, &ep, d->math, stream)) return e;
  }
  return 0;
}

namespace {
// a few reusable events for the main -> auxiliary stream hand-offs of ws_rdb_backward (cudaStreamWaitEvent captures
// the record that is current when it is called, so re-recording an event later does not disturb earlier waits)
cudaEvent_t rdb_event() {
  static std::mutex mu;
  static cudaEvent_t ring[16];
  static int next = -1;
  std::lock_guard<std::mutex> lk(mu);
  if (next < 0) {
    for (auto& e : ring)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    next = 0;
  }
  cudaEvent_t e = ring[next];
  next = (next + 1) % 16;
  return e;
}
}  // namespace

extern "C" int ws_rdb_backward(const ws_rdb_desc* d, const ws_tensor* dy, const ws_tensor* buf,
                               const ws_tensor* dbuf, const ws_tensor* g_lff, const ws_tensor* gbuf,
                               const ws_tensor* dx, const float* const* w, void* const* packed,
                               float* const* dw, float* db_lff, void* workspace, size_t workspace_bytes,
                               void* stream, void* aux_stream, void* aux_workspace, size_t aux_workspace_bytes) {
  RdbGeom r;
  if (int e = rdb_geom(d, r)) return e;
  WS_REQUIRE(dy && dy->ptr && buf && buf->ptr && dbuf && dbuf->ptr && g_lff && g_lff->ptr && w && packed,
             "ws_rdb_backward: null pointer");
  WS_REQUIRE(d->nconv == 0 || (gbuf && gbuf->ptr), "ws_rdb_backward: null g scratch");
  cudaStream_t st = (cudaStream_t)stream;
  const long long v = (long long)d->x * d->y * d->z;
  const bool want_w = dw != nullptr;
  if (device_cc_major() == 10 && d->nconv > 0 && dx && dx->ptr && dx->vstride == dy->vstride &&
      dx->nstride == dy->nstride &&
      rdb_persist_bwd_ok(d, View(dy), View(buf), View(g_lff), View(gbuf), View(dx), device_sm_count()) &&
      dgrad_path(ConvGeom(r.lff), View(g_lff), d->math) == WS_PATH_TCGEN05) {
    // the whole data-gradient chain as one persistent cooperative kernel (rdb_persist.cu); the weight-gradient GEMMs
    // follow on the auxiliary stream (or behind it)
    if (d->repack)
      if (int e = rdb_repack(d, r, g_lff, gbuf, w, packed, 1, st)) return e;
    if (int e = rdb_persist_backward(d, View(dy), View(buf), View(g_lff), View(gbuf), View(dx), packed, st)) return e;
    if (!want_w) return 0;
    const bool merged = rdb_merged_wgrad_ok(d, buf, gbuf);
    const bool aux = aux_stream && aux_stream != stream && aux_workspace &&
                     aux_workspace_bytes >= ws_rdb_backward_workspace_bytes(d);
    cudaStream_t wst = aux ? (cudaStream_t)aux_stream : st;
    void* wws = aux ? aux_workspace : workspace;
    const size_t wws_bytes = aux ? aux_workspace_bytes : workspace_bytes;
    if (aux) {
      cudaEvent_t ev = rdb_event();
      WS_REQUIRE(ev != nullptr, "ws_rdb_backward: cannot create events");
      WS_CHECK_CUDA(cudaEventRecord(ev, st));  // g_lff and every g_i are complete
      WS_CHECK_CUDA(cudaStreamWaitEvent(wst, ev, 0));
    }
    if (dw[d->nconv] || db_lff)
      if (int e = ws_conv3d_wgrad(&r.lff, buf, g_lff, dw[d->nconv], db_lff, 0, d->math, wws, wws_bytes, (void*)wst))
        return e;
    if (merged) {
      ws_conv_shape ms = rdb_merged_shape(d);
      int cin[WS_RDB_MAX_CONVS];
      for (int j = 0; j < d->nconv; ++j) cin[j] = r.dense[j].cin;
      return tc_rdb_wgrad(ConvGeom(ms), View(buf), View(gbuf), dw, cin, d->nconv, d->gc, wws, wws_bytes, wst);
    }
    for (int i = 0; i < d->nconv; ++i) {
      if (!dw[i]) continue;
      ws_tensor gslice = slice(*gbuf, i * d->gc);
      if (int e = ws_conv3d_wgrad(&r.dense[i], buf, &gslice, dw[i], nullptr, 0, d->math, wws, wws_bytes, (void*)wst))
        return e;
    }
    return 0;
  }
  // gradient entering the LFF accumulator: alpha * dy, in the activation dtype
  if (int e = axpby_launch(View(dy), d->alpha, View((const ws_tensor*)nullptr), 0.f, View(g_lff), d->n, d->f, v, st))
    return e;
  bool dx_done = false;
  // gbuf holds the gradients of ALL dense conv outputs side by side (n, nconv*gc, ..): the dgrad chain consumes
  // slice i as it goes, the weight gradients of the dense convs are then one merged GEMM at the end.
  const bool merged = want_w && rdb_merged_wgrad_ok(d, buf, gbuf);
  // Weight gradients are off the critical path (nothing later in this block, or in the blocks before it, reads them):
  // with an auxiliary stream they run beside the latency-bound data-gradient chain.  The caller joins the auxiliary
  // stream before anything consumes dw / db_lff and keeps buf, g_lff, gbuf alive until then.
  const bool aux = merged && aux_stream && aux_stream != stream && aux_workspace &&
                   aux_workspace_bytes >= ws_rdb_backward_workspace_bytes(d);
  cudaStream_t wst = aux ? (cudaStream_t)aux_stream : st;
  void* wws = aux ? aux_workspace : workspace;
  const size_t wws_bytes = aux ? aux_workspace_bytes : workspace_bytes;
  if (want_w && (dw[d->nconv] || db_lff)) {
    if (aux) {
      cudaEvent_t ev = rdb_event();
      WS_REQUIRE(ev != nullptr, "ws_rdb_backward: cannot create events");
      WS_CHECK_CUDA(cudaEventRecord(ev, st));  // g_lff is complete
      WS_CHECK_CUDA(cudaStreamWaitEvent(wst, ev, 0));
    }
    if (int e = ws_conv3d_wgrad(&r.lff, buf, g_lff, dw[d->nconv], db_lff, 0, d->math, wws, wws_bytes, (void*)wst))
      return e;
  }
  if (d->repack)
    if (int e = rdb_repack(d, r, g_lff, gbuf, w, packed, 1, st)) return e;
  // The gradient of dense conv j's output is final once the data-gradients of the LFF and of the convs i > j have
  // been accumulated: those are the LAST gc channels of the data-gradient that runs just before conv j's turn, so on
  // the tensor-core path that epilogue applies the LeakyReLU-backward and writes g_j itself ("tail", windsr.h).
  auto tail_ok = [&](const ws_conv_shape* s, const ws_tensor* act) {
    return !getenv("WS_DISABLE_RDB_TAIL") && d->gc % 16 == 0 &&
           dgrad_path(ConvGeom(*s), View(act), d->math) == WS_PATH_TCGEN05;
  };
  auto set_tail = [&](ws_epilogue& ep, int j, int total_c) {  // j: the conv whose output gradient is the tail
    ep.tail_out = slice(*gbuf, j * d->gc);
    ep.tail_mask = slice(*buf, r.dense[j].cin);
    ep.tail_c0 = total_c - d->gc;
    ep.tail_slope = d->slope;
  };
  bool g_ready = false;  // g of the conv about to be processed was produced by the previous epilogue
  {
    ws_epilogue ep = plain_epilogue();
    if (d->nconv > 0 && tail_ok(&r.lff, g_lff)) {
      set_tail(ep, d->nconv - 1, r.ctot);
      g_ready = true;
    }
    if (int e = ws_conv3d_dgrad(&r.lff, g_lff, packed[d->nconv], dbuf, &ep, d->math, stream)) return e;
  }
  for (int i = d->nconv - 1; i >= 0; --i) {
    const ws_conv_shape* s = &r.dense[i];
    ConvGeom g(*s);
    ws_tensor dslice = slice(*dbuf, s->cin), yslice = slice(*buf, s->cin), gslice = slice(*gbuf, i * d->gc);
    const ws_tensor* gi = &gslice;
    // g_i = dbuf[:, cin:cin+gc] * lrelu'(buf[:, cin:cin+gc])
    if (!g_ready)
      if (int e = lrelu_bwd_launch(View(&dslice), View(&yslice), d->slope, nullptr, nullptr, View(gi), d->n, d->gc,
                                   v, st))
        return e;
    g_ready = false;
    if (want_w && dw[i] && !merged) {
      if (int e = ws_conv3d_wgrad(s, buf, gi, dw[i], nullptr, 0, d->math, workspace, workspace_bytes, stream))
        return e;
    }
    ws_epilogue ep = plain_epilogue();
    ep.res1 = *dbuf; ep.beta1 = 1.f;  // accumulate into dbuf[:, :cin]
    if (i > 0 && tail_ok(s, gi)) {
      set_tail(ep, i - 1, s->cin);  // channels [cin - gc, cin) of this gradient = output of conv i-1
      g_ready = true;
    }
    if (i == 0 && merged) {
      // every g_j is complete (the tail epilogues / LeakyReLU-backward passes above): the merged weight gradient of
      // the dense convs can go — beside conv0's data-gradient when there is an auxiliary stream
      if (aux) {
        cudaEvent_t ev = rdb_event();
        WS_REQUIRE(ev != nullptr, "ws_rdb_backward: cannot create events");
        WS_CHECK_CUDA(cudaEventRecord(ev, st));
        WS_CHECK_CUDA(cudaStreamWaitEvent(wst, ev, 0));
      }
      ws_conv_shape ms = rdb_merged_shape(d);
      int cin[WS_RDB_MAX_CONVS];
      for (int j = 0; j < d->nconv; ++j) cin[j] = r.dense[j].cin;
      if (int e = tc_rdb_wgrad(ConvGeom(ms), View(buf), View(gbuf), dw, cin, d->nconv, d->gc, wws, wws_bytes, wst))
        return e;
    }
    if (i == 0 && dx && dx->ptr) {
      // conv0 reads exactly the block input (cin = f): its data-gradient epilogue also adds the skip term and
      // writes dL/dx directly — dx = dbuf[:, :f] + dgrad_0 + beta1 * dy — instead of a separate axpby pass
      ep.res2 = *dy; ep.beta2 = d->beta1;
      if (int e = ws_conv3d_dgrad(s, gi, packed[i], dx, &ep, d->math, stream)) return e;
      dx_done = true;
      continue;
    }
    if (int e = ws_conv3d_dgrad(s, gi, packed[i], dbuf, &ep, d->math, stream)) return e;
  }
  if (dx && dx->ptr && !dx_done) {
    if (int e = axpby_launch(View(dbuf), 1.f, View(dy), d->beta1, View(dx), d->n, d->f, v, st)) return e;
  }
  return 0;
}

// ---- batched weight gradients of a run of identical residual dense blocks (wgrad_tc.cu: tc_trunk_wgrad) -------------
extern "C" size_t ws_trunk_wgrad_record_floats(const ws_rdb_desc* d) {
  RdbGeom r;
  if (!d || rdb_geom(d, r)) return 0;
  size_t total = 0;
  for (int i = 0; i < d->nconv; ++i)
    total += (size_t)r.dense[i].cout * r.dense[i].cin * r.dense[i].kx * r.dense[i].ky * r.dense[i].kz;
  total += (size_t)r.lff.cout * r.lff.cin * r.lff.kx * r.lff.ky * r.lff.kz;
  return total + (size_t)r.lff.cout;
}

extern "C" size_t ws_trunk_wgrad_workspace_bytes(const ws_rdb_desc* d, int nblocks) {
  if (!d || nblocks < 1 || d->nconv < 1) return 0;
  return tc_trunk_wgrad_workspace_bytes(nblocks, d->k * d->k * d->k, d->f + (d->nconv - 1) * d->gc, d->nconv, d->gc);
}

namespace {
bool trunk_wgrad_ok(const ws_rdb_desc* d, const RdbGeom& r, const ws_tensor* buf, const ws_tensor* g,
                    const ws_tensor* g_lff) {
  return d->k_lff == 1 && d->nconv >= 2 && device_cc_major() == 10 && rdb_merged_wgrad_ok(d, buf, g) &&
         wgrad_path(ConvGeom(r.lff), View(buf), View(g_lff), d->math) == WS_PATH_TCGEN05 &&
         View(g_lff).cs == 1 && r.lff.cout % 8 == 0 && 256 % (r.lff.cout / 8) == 0;
}
}  // namespace

/* 1 when ws_trunk_wgrad covers blocks of this geometry / precision with densely packed channels-last slabs */
extern "C" int ws_trunk_wgrad_supported(const ws_rdb_desc* d) {
  RdbGeom r;
  if (!d || d->nconv < 1 || d->nconv > WS_RDB_MAX_CONVS || d->k % 2 != 1 || d->k_lff % 2 != 1) return 0;
  rdb_geom(d, r);
  const int dt = d->math == WS_MATH_BF16 ? WS_BF16 : WS_F32;
  const long long v = (long long)d->x * d->y * d->z;
  void* fake = reinterpret_cast<void*>(uintptr_t(1) << 20);  // aligned stand-in: only layouts are examined
  ws_tensor buf = {fake, dt, 0, v * r.ctot, r.ctot, 1};
  ws_tensor g = {fake, dt, 0, v * d->nconv * d->gc, (long long)d->nconv * d->gc, 1};
  ws_tensor gl = {fake, dt, 0, v * d->f, d->f, 1};
  return trunk_wgrad_ok(d, r, &buf, &g, &gl) ? 1 : 0;
}

extern "C" int ws_trunk_wgrad(const ws_rdb_desc* d, int nblocks, const ws_tensor* buf, const ws_tensor* g,
                              const ws_tensor* g_lff, float* grads, long long block_stride, void* workspace,
                              size_t workspace_bytes, void* stream) {
  RdbGeom r;
  if (int e = rdb_geom(d, r)) return e;
  WS_REQUIRE(nblocks >= 1 && buf && buf->ptr && g && g->ptr && g_lff && g_lff->ptr && grads,
             "ws_trunk_wgrad: null pointer");
  WS_REQUIRE(trunk_wgrad_ok(d, r, buf, g, g_lff),
             "ws_trunk_wgrad: geometry / precision outside the batched tensor-core path");
  WS_REQUIRE((size_t)block_stride >= ws_trunk_wgrad_record_floats(d), "ws_trunk_wgrad: block_stride too small");
  cudaStream_t st = (cudaStream_t)stream;
  ws_conv_shape ms = rdb_merged_shape(d), ls = r.lff;
  ms.n = ls.n = d->n * nblocks;  // the batch dimension of the tensor maps runs over (block, sample)
  int cin[WS_RDB_MAX_CONVS];
  for (int j = 0; j < d->nconv; ++j) cin[j] = r.dense[j].cin;
  if (int e = tc_trunk_wgrad(ConvGeom(ms), ConvGeom(ls), nblocks, View(buf), View(g), View(g_lff), cin, d->nconv, d->gc,
                             grads, block_stride, workspace, workspace_bytes, st))
    return e;
  const long long off_bias = (long long)ws_trunk_wgrad_record_floats(d) - r.lff.cout;
  return bias_grad_batched(View(g_lff), grads + off_bias, nblocks, block_stride, d->n, r.lff.cout,
                           (long long)d->x * d->y * d->z, st);
}

// ---- weight packing for a whole run of identical residual dense blocks (one launch per direction) ------------------
// Returns 0 when the packings of every block have been enqueued, 1 when this geometry / layout does not take the
// persistent per-block kernels (the caller then lets ws_rdb_forward / ws_rdb_backward repack block by block), another
// value on error.  w / packed: [block][conv 0 .. nconv-1, LFF].
extern "C" int ws_trunk_repack_fwd(const ws_rdb_desc* d, int nblocks, const ws_tensor* x, const ws_tensor* buf,
                                   const ws_tensor* out, const float* const* w, void* const* packed, void* stream) {
  RdbGeom r;
  if (int e = rdb_geom(d, r)) return e;
  WS_REQUIRE(nblocks >= 1 && x && buf && out && w && packed, "ws_trunk_repack_fwd: null pointer");
  const bool rows_ok = out->dtype == WS_F32 && out->cstride == 1 && x->vstride == out->vstride &&
                       x->nstride == out->nstride && !(reinterpret_cast<uintptr_t>(out->ptr) & 15);
  if (!(device_cc_major() == 10 && rows_ok && rdb_persist_ok(d, View(x), View(buf), device_sm_count()) &&
        fwd_path(ConvGeom(r.lff), View(buf), d->math) == WS_PATH_TCGEN05))
    return 1;
  if (nblocks * (d->nconv + 1) > 64 * (WS_RDB_MAX_CONVS + 1)) return 1;
  ConvGeom g[WS_RDB_MAX_CONVS + 1];
  int fold[WS_RDB_MAX_CONVS + 1];
  for (int i = 0; i <= d->nconv; ++i) {
    g[i] = ConvGeom(i < d->nconv ? r.dense[i] : r.lff);
    if (fwd_path(g[i], View(buf), d->math) != WS_PATH_TCGEN05) return 1;
    fold[i] = i < d->nconv ? 2 : 0;  // dense convs: z-folded (rdb_persist.cu)
  }
  return pack_tc_trunk_launch(nblocks, d->nconv + 1, g, fold, 0, w, packed, (cudaStream_t)stream, 0);
}

extern "C" int ws_trunk_repack_bwd(const ws_rdb_desc* d, int nblocks, const ws_tensor* dy, const ws_tensor* buf,
                                   const ws_tensor* g_lff, const ws_tensor* gbuf, const ws_tensor* dx,
                                   const float* const* w, void* const* packed, void* stream) {
  RdbGeom r;
  if (int e = rdb_geom(d, r)) return e;
  WS_REQUIRE(nblocks >= 1 && dy && buf && g_lff && gbuf && dx && w && packed, "ws_trunk_repack_bwd: null pointer");
  if (!(device_cc_major() == 10 && d->nconv > 0 && dx->ptr && dx->vstride == dy->vstride && dx->nstride == dy->nstride &&
        rdb_persist_bwd_ok(d, View(dy), View(buf), View(g_lff), View(gbuf), View(dx), device_sm_count()) &&
        dgrad_path(ConvGeom(r.lff), View(g_lff), d->math) == WS_PATH_TCGEN05))
    return 1;
  if (nblocks * (d->nconv + 1) > 64 * (WS_RDB_MAX_CONVS + 1)) return 1;
  ConvGeom g[WS_RDB_MAX_CONVS + 1];
  for (int i = 0; i <= d->nconv; ++i) {
    g[i] = ConvGeom(i < d->nconv ? r.dense[i] : r.lff);
    if (i < d->nconv) {
      ws_tensor gs = slice(*gbuf, i * d->gc);
      if (dgrad_path(g[i], View(&gs), d->math) != WS_PATH_TCGEN05) return 1;
    }
  }
  return pack_tc_trunk_launch(nblocks, d->nconv + 1, g, nullptr, 1, w, packed, (cudaStream_t)stream, 0);
}
