// C-ABI entry for the 3x3 implicit-GEMM convolution used by SPADE (forward convs
// and input gradients).  Dispatch: the tcgen05/TMEM/TMA kernel (k3_conv_tc.cu)
// for the shapes it supports, otherwise the generic mma.sync kernel
// (k3_conv_mma.cu).  Both are GPU kernels of this library; there is no CPU or
// vendor-library fallback.
#include "k3_common.cuh"

namespace ag2v {
int conv3x3_check(const ConvParams& p, int epi);
int conv3x3_mma(const ConvParams& p, int epi, int round_out, int precise, cudaStream_t stream);
int conv3x3_tc(const ConvParams& p, int epi, int round_out, cudaStream_t stream);   // may return AG2V_ERR_UNSUPPORTED
bool conv3x3_tc_supported(const ConvParams& p, int epi);
bool conv3x3_tc_stat_geometry(const ConvParams& p, int groups, int* mtiles, int* tiles_per_group);
int conv_stats_reduce(const float* part, int tiles_per_group, int Nout, int groups, double* sums, cudaStream_t stream);
}  // namespace ag2v

using namespace ag2v;

// impl: 0 = auto (tcgen05 when supported), 1 = force mma.sync, 2 = force tcgen05 (error if unsupported),
//       3 = mma.sync with 3xTF32 products (fp32-class accuracy; validation mode)
extern "C" int ag2v_conv3x3(const float* in, long long in_sb, long long in_sy, long long in_sx, int B, int Hh,
                            int Ww, int Cin, const float* wpk, const float* bias, int Nout, float* out,
                            long long out_sb, long long out_sy, long long out_sx, int epilogue, int round_out,
                            const float* x, const float* mean, const float* rstd, float* gamma_out,
                            float slope, int C, long long group_pixels, int x_up, const float* scale, const float* res,
                            const float* gate, float* splitk_ws, size_t splitk_ws_floats, int impl,
                            cudaStream_t stream) {
  ConvParams p{};
  p.in = in; p.in_sb = in_sb; p.in_sy = in_sy; p.in_sx = in_sx;
  p.B = B; p.Hh = Hh; p.Ww = Ww; p.Cin = Cin;
  p.wpk = wpk; p.bias = bias; p.Nout = Nout;
  p.out = out; p.out_sb = out_sb; p.out_sy = out_sy; p.out_sx = out_sx;
  p.x = x; p.mean = mean; p.rstd = rstd; p.gamma_out = gamma_out; p.slope = slope; p.C = C;
  p.group_pixels = group_pixels; p.x_up = x_up; p.scale = scale; p.res = res;
  p.gate = gate;
  p.splitk_ws = splitk_ws; p.splitk_ws_floats = splitk_ws_floats;
  int rc = conv3x3_check(p, epilogue);
  if (rc) return rc;
  if (impl == 1) return conv3x3_mma(p, epilogue, round_out, 0, stream);
  if (impl == 3) return conv3x3_mma(p, epilogue, 0, 1, stream);
  if (impl == 2) {
    if (!conv3x3_tc_supported(p, epilogue))
      return fail(AG2V_ERR_UNSUPPORTED, "conv3x3: shape not supported by the tcgen05 kernel (Cin=%d Nout=%d %dx%d)", Cin, Nout, Hh, Ww);
    return conv3x3_tc(p, epilogue, round_out, stream);
  }
  if (conv3x3_tc_supported(p, epilogue)) return conv3x3_tc(p, epilogue, round_out, stream);
  return conv3x3_mma(p, epilogue, round_out, 0, stream);
}

// Floats of split-K scratch worth passing to ag2v_conv3x3 for this shape (0: not needed).
extern "C" size_t ag2v_conv3x3_splitk_floats(int B, int Hh, int Ww, int Cin, int Nout) {
  const long long tiles = ceil_div_ll((long long)B * Hh * Ww, 128) * ceil_div(Nout, 128);
  const int total = 9 * (Cin / 32);
  if (tiles * 2 > sm_count() || total < 24 || Cin % 32) return 0;
  long long ks = ceil_div_ll(sm_count(), tiles);
  if (ks > total / 8) ks = total / 8;
  return ks <= 1 ? 0 : (size_t)ks * B * Hh * Ww * Nout;
}

extern "C" int ag2v_conv3x3_tc_supported(int B, int Hh, int Ww, int Cin, int Nout, int epilogue) {
  ConvParams p{};
  p.B = B; p.Hh = Hh; p.Ww = Ww; p.Cin = Cin; p.Nout = Nout; p.C = Nout / 2;
  return conv3x3_tc_supported(p, epilogue) ? 1 : 0;
}

// ---- BN statistics fused into the producing convolution's epilogue (north_star; normalization.py:99 consumes them) ----
// 1 if y = conv3x3(x) * scale + bias (+ res) on a dense NHWC [B,Hh,Ww,*] pair can also emit the per-channel sums of y per
// statistics group (B / groups consecutive images): tcgen05 kernel, no split-K, tiles not straddling groups.
// mtiles / tiles_per_group size the scratch: stat_part = mtiles * 2 * Nout floats.
extern "C" int ag2v_conv3x3_stats_info(int B, int Hh, int Ww, int Cin, int Nout, int groups, int* mtiles, int* tiles_per_group) {
  ConvParams p{};
  p.B = B; p.Hh = Hh; p.Ww = Ww; p.Cin = Cin; p.Nout = Nout;
  int mt = 0, tpg = 0;
  // Cin < 128: the main loop of a work item (9 * Cin / 32 steps) is shorter than the epilogue with the column sums in
  // it, and the kernel becomes epilogue-bound (measured on B200, r=256, 64 -> 64: 94 us fused against 51 us + 30 us
  // for the separate statistics pass; 128 -> 64: 79 us against 90 us) - those layers keep the separate pass.
  if (Cin < 128 || !conv3x3_tc_supported(p, EPI_BIAS) || ag2v_conv3x3_splitk_floats(B, Hh, Ww, Cin, Nout) != 0 ||
      !conv3x3_tc_stat_geometry(p, groups, &mt, &tpg))
    return 0;
  if (mtiles) *mtiles = mt;
  if (tiles_per_group) *tiles_per_group = tpg;
  return 1;
}

// The EPI_BIAS convolution of ag2v_conv3x3 (dense NHWC input and output) that also writes stat_part; then
// sums[g][0][c] = sum y, sums[g][1][c] = sum y^2 over group g in double (the layout ag2v_bn_finalize reads).
extern "C" int ag2v_conv3x3_bias_stats(const float* in, int B, int Hh, int Ww, int Cin, const float* wpk, const float* bias,
                                       int Nout, float* out, int round_out, long long group_pixels, const float* scale,
                                       const float* res, int groups, float* stat_part, double* sums, cudaStream_t stream) {
  ConvParams p{};
  p.in = in; p.in_sb = (long long)Hh * Ww * Cin; p.in_sy = (long long)Ww * Cin; p.in_sx = Cin;
  p.B = B; p.Hh = Hh; p.Ww = Ww; p.Cin = Cin; p.wpk = wpk; p.bias = bias; p.Nout = Nout;
  p.out = out; p.out_sb = (long long)Hh * Ww * Nout; p.out_sy = (long long)Ww * Nout; p.out_sx = Nout;
  p.slope = 1.f; p.group_pixels = group_pixels; p.scale = scale; p.res = res;
  p.stat_part = stat_part;
  AG2V_REQUIRE(stat_part && sums, "conv3x3_bias_stats: null statistics buffers");
  int rc = conv3x3_check(p, EPI_BIAS);
  if (rc) return rc;
  int mt = 0, tpg = 0;
  if (!ag2v_conv3x3_stats_info(B, Hh, Ww, Cin, Nout, groups, &mt, &tpg))
    return fail(AG2V_ERR_UNSUPPORTED, "conv3x3_bias_stats: shape %dx%dx%d %d->%d (groups %d) has no fused statistics", B, Hh, Ww, Cin, Nout, groups);
  rc = conv3x3_tc(p, EPI_BIAS, round_out, stream);
  if (rc) return rc;
  return conv_stats_reduce(stat_part, tpg, Nout, groups, sums, stream);
}
