// C-ABI entry point for the fused LoRA Conv2d (NHWC implicit GEMM, stride 1, dilation 1,
// groups 1; 1x1 and 3x3 in the SD1.5 ResnetBlock2D sites). Forward and, on flipped/transposed
// frozen weights, the input gradient. Replaces LoraInjectedConv2d.forward
// (/root/reference/lora_diffusion/lora.py:130-135) and the dX part of its autograd backward.
// Kernel: fused_core.cuh.
#include "fused_core.cuh"
#include "lora_b200.h"
#include "tmap.h"

namespace lb {

template <int BLOCK_N, int STAGES, typename OutT, int G>
static int launch_conv(const void* X, const void* W, const void* Dn, void* Y, const FusedParams& p,
                       int n_img, int down_cols, int out_dtype, cudaStream_t stream) {
  using S = Smem<BLOCK_N, STAGES, OutT, G>;
  auto kern = fused_lora_kernel<BLOCK_N, STAGES, OutT, true, G>;
  static unsigned long long attr_mask = 0;
  if (!ensure_dyn_smem(reinterpret_cast<const void*>(kern), S::DYN_BYTES, attr_mask)) return LB_ERR_CUDA;
  const CUtensorMapDataType in_dt = p.fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapDataType out_dt = out_dtype == LB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : in_dt;
  CUtensorMap tmX, tmW, tmD, tmY;
  if (!tmap_nhwc(&tmX, X, in_dt, 2, p.C, p.W, p.H, n_img, BLOCK_K, p.TW, p.TH)) return LB_ERR_TMAP;
  if (!tmap_2d(&tmW, W, in_dt, 2, p.K, p.N, BLOCK_K, BLOCK_N, true)) return LB_ERR_TMAP;
  if (!tmap_2d(&tmD, Dn, in_dt, 2, down_cols, R_PAD, BLOCK_K, R_PAD, true)) return LB_ERR_TMAP;
  if (!tmap_nhwc(&tmY, Y, out_dt, sizeof(OutT), p.N, p.W, p.H, n_img, S::BOX_COLS, p.TW, p.TH)) return LB_ERR_TMAP;
  dim3 grid((p.N + BLOCK_N - 1) / BLOCK_N, n_img * p.tiles_h * p.tiles_w, 1);
  return launch_ex(kern, grid, dim3(NUM_THREADS), S::DYN_BYTES, stream, 1, tmX, tmW, tmD, tmY, p) == cudaSuccess
             ? LB_OK : LB_ERR_CUDA;
}

}  // namespace lb

extern "C" int lb_lora_conv2d_fwd(const void* X, const void* W, const float* bias,
                                  const void* down16, const float* up, long long up_rs,
                                  long long up_cs, long long up_gs, const float* diag, float scale,
                                  void* Y, float* T_out, const float* T_in, int n_img, int H, int Wd, int Cin,
                                  int Cout, int kh, int kw, int pad_h, int pad_w, int r,
                                  int per_tap_T, int in_dtype, int out_dtype, void* stream) {
  using namespace lb;
  if (n_img <= 0 || H <= 0 || Wd <= 0 || Cin <= 0 || Cout <= 0) return LB_ERR_SHAPE;
  if (!((kh == 1 && kw == 1) || (kh == 3 && kw == 3))) return LB_ERR_SHAPE;
  // "same" geometry only (output extent == input extent), which is what stride-1 SD convs use
  if (2 * pad_h != kh - 1 || 2 * pad_w != kw - 1) return LB_ERR_SHAPE;
  if (r < 1 || r > R_PAD) return LB_ERR_RANK;
  if (in_dtype != LB_BF16 && in_dtype != LB_F16) return LB_ERR_DTYPE;
  if (out_dtype != in_dtype && out_dtype != LB_F32) return LB_ERR_DTYPE;
  if ((Cin % 8) != 0 || (Cout % 8) != 0) return LB_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) |
       reinterpret_cast<uintptr_t>(down16) | reinterpret_cast<uintptr_t>(Y) |
       reinterpret_cast<uintptr_t>(T_out) | reinterpret_cast<uintptr_t>(T_in)) & 15)
    return LB_ERR_ALIGN;

  const int taps = kh * kw;
  FusedParams p = {};
  p.bias = bias; p.up = up; p.up_rs = up_rs; p.up_cs = up_cs; p.up_gs = up_gs; p.diag = diag;
  p.t_out = T_out; p.t_in = T_in; p.scale = scale;
  p.M = n_img * H * Wd; p.N = Cout; p.K = taps * Cin; p.r = r;
  p.fmt = (in_dtype == LB_BF16) ? 1 : 0;
  p.H = H; p.W = Wd; p.C = Cin; p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w;
  p.TW = Wd >= 16 ? 16 : 8; p.TH = BLOCK_M / p.TW;
  p.tiles_h = (H + p.TH - 1) / p.TH; p.tiles_w = (Wd + p.TW - 1) / p.TW;
  const bool groups = per_tap_T != 0 && taps > 1;
  p.down_per_tap = groups ? 0 : 1;
  p.t_group = groups ? pad_h * kw + pad_w : 0;   // the tap with zero shift
  const int down_cols = groups ? Cin : taps * Cin;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  const long long tiles128 = static_cast<long long>(n_img) * p.tiles_h * p.tiles_w * ((Cout + 127) / 128);
  const bool narrow = tiles128 < 120;
#define LB_CONV(BN, OT, GG) launch_conv<BN, 4, OT, GG>(X, W, down16, Y, p, n_img, down_cols, out_dtype, st)
  if (groups) {
    if (out_dtype == LB_F32) return narrow ? LB_CONV(64, float, 9) : LB_CONV(128, float, 9);
    return narrow ? LB_CONV(64, uint16_t, 9) : LB_CONV(128, uint16_t, 9);
  }
  if (out_dtype == LB_F32) return narrow ? LB_CONV(64, float, 1) : LB_CONV(128, float, 1);
  return narrow ? LB_CONV(64, uint16_t, 1) : LB_CONV(128, uint16_t, 1);
#undef LB_CONV
}
