// Tensor-core (tcgen05) path: chunk-plane activations and the conv entry points (conv_tc.cu).
#pragma once
#include "common.h"

namespace orca {

// (nb, C, n) activation as two bf16 planes hi/lo[nb][C/8][npad][8]; data row l lives at row l + 4,
// rows [0,4) and [n+4, npad) are zero (they are the convolution's zero padding).
struct TcAct {
  void* hi = nullptr;
  void* lo = nullptr;
  int nb = 0, C = 0;
  int64_t n = 0, npad = 0;
};

inline int64_t tc_npad(int64_t n) { return ((n + 127) / 128) * 128 + 8; }
inline size_t tc_plane_bytes(int nb, int C, int64_t n) { return (size_t)nb * (C / 8) * tc_npad(n) * 16; }  // one of hi/lo

bool tc_layer_eligible(const ConvLayer& L);
// out = act(conv(in) + b) [+ res], optionally max-pooled by `pool` (1, 2 or 4) along n, written either
// as chunk planes (out_planes) or as fp32 channel-last [nb][n/pool][C_out] (out_f32).
int tc_conv1d(const ConvLayer& L, const TcAct& in, const TcAct* res, TcAct* out_planes, float* out_f32, int pool,
              int relu, cudaStream_t s);
int tc_conv_first(const ConvLayer& L, const float* x, int64_t sB, int64_t sC, int64_t sL, int nb, int64_t Ltot,
                  int64_t l_begin, int64_t n, TcAct* out, cudaStream_t s);
int tc_pool_planes(const TcAct& in, TcAct* out, int p, cudaStream_t s);

}  // namespace orca
