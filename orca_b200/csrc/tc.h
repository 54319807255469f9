// Tensor-core (tcgen05) path: chunk-plane activations and the conv entry points
// (conv_tc.cu: Conv1d k=9).  The dilated 3x3 Conv2d path of the decoders lives in dec_stream.h / conv2d_stream.cu.
#pragma once
#include <vector>
#include "common.h"

namespace orca {

// ---- 1D: (nb, C, n) activation as two bf16 planes hi/lo[nb][C/8][npad][8]; data row l lives at row
// l + 4, rows [0,4) and [n+4, npad) are zero (they are the convolution's zero padding).
// fmt 0: two bf16 planes (x = hi + lo, three tensor-core products per conv); fmt 1: ONE fp16 plane in `hi`
// (lo unused; one tensor-core product per conv) -- the early encoder stages, see modules.cu encoder_window_tc.
struct TcAct {
  void* hi = nullptr;
  void* lo = nullptr;
  int nb = 0, C = 0;
  int64_t n = 0, npad = 0;
  int fmt = 0;
  // fmt 1 only: device word that kernels OR 1 into when a value they round to fp16 exceeds the fp16 range guard
  // (|x| > 60000); nullptr = no check.  Read through orca_b200_module_status (modules.cu).
  unsigned int* sat = nullptr;
};

inline int64_t tc_npad(int64_t n) { return ((n + 127) / 128) * 128 + 8; }
inline size_t tc_plane_bytes(int nb, int C, int64_t n) { return (size_t)nb * (C / 8) * tc_npad(n) * 16; }  // one of hi/lo

bool tc_layer_eligible(const ConvLayer& L);
int tc_pack_layer(ConvLayer& L, const float* w_folded /*[tap][c_in][c_out]*/, std::vector<void*>& allocs);
// out = act(conv(in) + b) [+ res], optionally max-pooled by `pool` (1, 2 or 4) along n, written either
// as chunk planes (out_planes) or as fp32 channel-last [nb][n/pool][C_out] (out_f32).
// Both outputs may be requested at once; res2 is a second residual with the geometry of res.
int tc_conv1d(const ConvLayer& L, const TcAct& in, const TcAct* res, TcAct* out_planes, float* out_f32, int pool,
              int relu, cudaStream_t s, const TcAct* res2 = nullptr);
// nearest x2 upsample / channel-last fp32 -> planes, on chunk planes (U-net glue)
int tc_upsample2_planes(const TcAct& in, TcAct* out, cudaStream_t s);
int tc_from_channel_last(const float* xcl /*[nb][n][C]*/, TcAct* out, cudaStream_t s);
int tc_conv_first(const ConvLayer& L, const SeqIn& in, int nb, int64_t Ltot, int64_t l_begin, int64_t n, TcAct* out,
                  cudaStream_t s);
int tc_pool_planes(const TcAct& in, TcAct* out, int p, cudaStream_t s);
// lconv1 (Conv 4->64, BN, Conv 64->64, BN: no nonlinearity) composed into ONE k=17 tensor-core conv
// (conv_first_tc.cu); L0/L1 = lconv1[0], lconv1[1] (their fp32 weights are used for the sequence-end fix).
int tc_pack_lconv1(ConvLayer& L0, const float* w1, const float* b1, const float* w2, const float* b2,
                   std::vector<void*>& allocs);
int tc_lconv1(const ConvLayer& L0, const ConvLayer& L1, const SeqIn& in, int nb, int64_t Ltot, int64_t l_begin, int64_t n,
              TcAct* out, cudaStream_t s);

// 2D maps and the decoder kernels: dec_stream.h

// glue.cu
int symmetrise(const float* tmp, float* out, int B, int S, cudaStream_t s);

}  // namespace orca
