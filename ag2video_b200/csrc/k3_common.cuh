// Shared definitions for the SPADE kernels (reference:
// models/spade_models/networks/normalization.py:66-110, architecture.py:50-68).
//
// Internal layout: every activation the K3 kernels touch is NHWC fp32 (logical
// NCHW tensors in torch.channels_last memory format), so the implicit-GEMM
// reduction index (ky, kx, c) is contiguous in c.
//
// gamma/beta packing ("gb8"): the two modulation convolutions run as ONE GEMM
// whose output columns interleave gamma and beta in groups of eight channels:
//     n = 16*(c/8) + 8*is_beta + (c%8)
// so that the thread that owns gamma[p,c] in the accumulator also owns beta[p,c]
// and the epilogue can apply  out = xhat*(1+gamma)+beta  without a round trip.
#pragma once
#include "common.cuh"

namespace ag2v {

enum ConvEpilogue : int {
  EPI_BIAS = 0,        // out = acc * scale[group] + bias + res   (scale / bias / res optional)
  EPI_BIAS_RELU = 1,   // out = relu(acc + bias)                    (mlp_shared, normalization.py:103)
  EPI_SPADE = 2,       // out = act((x-mean)*rstd*(1+g)+b), g/b = acc + bias (normalization.py:104-108)
  EPI_GATE = 3,        // out = gate > 0 ? acc : 0                  (backward through the ReLU)
  EPI_ACCUM = 4,       // out += acc   (gradient w.r.t. the shared full-resolution segmap)
  EPI_RAW = 5,         // internal: split-K partial sums, finished by conv_finish_kernel
};

struct ConvParams {
  // input view [B, R, R(w), Cin]: element strides for b / y / x, channels contiguous
  const float* in; long long in_sb, in_sy, in_sx;
  int B, Hh, Ww, Cin;
  const float* wpk;          // packed weights [9][Nout][Cin]
  const float* bias;         // [Nout] in packed column order, may be null
  int Nout;
  float* out; long long out_sb, out_sy, out_sx;   // output view [B, Hh, Ww, Nout]
  // EPI_SPADE: Nout == 2*C in gb8 order; out is [.., C]
  const float* x; const float* mean; const float* rstd; float* gamma_out; float slope; int C;
  long long group_pixels;    // > 0: mean/rstd are [groups][C], group = pixel index / group_pixels
  int x_up;                  // EPI_SPADE: x is [B, Hh/2, Ww/2, C] and is read through a nearest 2x up-sampling
  // EPI_BIAS extras (the spectrally normalised convolutions of SPADEResnetBlock run on weight_orig):
  // scale [groups] = 1/sigma per frame group (group = pixel / group_pixels), res = dense [P, Nout] residual
  const float* scale; const float* res;
  // EPI_GATE
  const float* gate;
  // split-K workspace (caller-provided, may be null): ksplit * P * Nout floats
  float* splitk_ws; size_t splitk_ws_floats;
  // EPI_BIAS, tcgen05 kernel, no split-K: per-M-tile column sums of the OUTPUT for the batch norm that consumes it
  // (normalization.py:99): stat_part[mtile][0][n] = sum over the tile's pixels of out[p, n], [mtile][1][n] = sum of squares
  float* stat_part;
  // tcgen05 kernel: {next work index, CTAs done} of the dynamic tile scheduler (set by the launcher)
  unsigned* sched;
};

__host__ __device__ inline int gb8_col(int c, int is_beta) { return 16 * (c >> 3) + 8 * is_beta + (c & 7); }

}  // namespace ag2v
