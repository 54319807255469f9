// Shared declarations of liborca_b200 (internal; the public ABI is include/orca_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>

#include "../../include/orca_b200.h"

namespace orca {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define ORCA_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      orca::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                        \
      return ORCA_B200_ECUDA;                                                           \
    }                                                                                   \
  } while (0)

#define ORCA_LAUNCH_OK()                                                            \
  do {                                                                              \
    orca::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    cudaError_t _e = cudaPeekAtLastError();                                         \
    if (_e != cudaSuccess) {                                                        \
      orca::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                      __FILE__, __LINE__);                                          \
      return ORCA_B200_ECUDA;                                                       \
    }                                                                               \
  } while (0)

#define ORCA_TRY(expr)             \
  do {                             \
    int _s = (expr);               \
    if (_s != ORCA_B200_OK) return _s; \
  } while (0)

// ---- one folded convolution layer on the device -----------------------------------------
// SIMT layout: w[tap][c_in][c_out] fp32 (BN scale folded in), b[c_out] fp32 (BN shift folded in).
// tcgen05 layout (when has_tc): see conv_tc.cu.
struct ConvLayer {
  int c_in = 0, c_out = 0, kh = 1, kw = 1, dil = 1;
  float* w = nullptr;
  float* b = nullptr;
  // tensor-core images (bf16 hi/lo split, UMMA-swizzled), optional
  void* tc_w = nullptr;
  size_t tc_w_bytes = 0;
  void* tc_w16 = nullptr;  // Conv1d k=9 only: fp16 image for the single-pass encoder stages (conv_tc.cu, FMT = 1)
  // extra input channels of the Decoder combiners (the 129th / 65th channel in orca_modules, num_2d of them in
  // orca_leukemia), kept out of the aligned implicit GEMM: w_extra[n_extra][tap][c_out]
  float* w_extra = nullptr;
  int n_extra = 0;
  // bias that goes with tc_w when it differs from b (the composed lconv1 of the Encoder, conv_first_tc.cu)
  float* tc_bias = nullptr;
};

// ---- generic conv (channel-last activations) -----------------------------------------
// out[b][y][x][:] = act(bias + sum_taps sum_ci w * in[b][y+dy][x+dx][ci]) + res + res2
struct ConvCall {
  const float* in = nullptr;   // [B][H][W][in_ld], first c_in channels used
  float* out = nullptr;        // [B][H][W][out_ld]
  const float* res = nullptr;  // optional, [B][H][W][res_ld]
  const float* res2 = nullptr; // optional, same geometry as res
  int B = 1, H = 1, W = 1;
  int in_ld = 0, out_ld = 0, res_ld = 0;
  int relu = 0;
};

int conv_simt(const ConvLayer& L, const ConvCall& c, cudaStream_t s);

// first encoder layer: Conv1d(4 -> 64, k9) on the caller's input window (strided fp32 (B,4,L) view or
// packed bases, seq_in.cuh).  `in` addresses sample 0 / position 0; positions [l_begin, l_begin + n) are
// produced, reading [l_begin-4, l_begin+n+4) clipped to [0, L) (zero outside = the conv's own padding).
struct SeqIn;
int conv_first_simt(const ConvLayer& L, const SeqIn& in, int B, int64_t Ltot, int64_t l_begin, int64_t n,
                    float* out /*[B][n][64]*/, cudaStream_t s);

// ---- glue kernels -------------------------------------------------------------------------
// out[b][l][c] = max_{i<p} (a[b][l*p+i][c] + (bb ? bb[b][l*p+i][c] : 0))
int add_maxpool1d(const float* a, const float* bb, float* out, int B, int64_t L_in, int C, int p,
                  cudaStream_t s);
// nearest x2: out[b][l][c] = in[b][l/2][c]
int upsample2_1d(const float* in, float* out, int B, int64_t L_in, int C, cudaStream_t s);
// strided (B,C,L) -> channel-last [B][L][C]
int to_channel_last(const float* x, int64_t sB, int64_t sC, int64_t sL, float* out, int B, int C,
                    int64_t L, cudaStream_t s);
// mat[b][i][j][c] = xcl[b][i][c] + xcl[b][j][c]   (xcl channel-last [B][S][C])
int outer_sum(const float* xcl, float* mat, int B, int C, int S, cudaStream_t s);
// n_extra-channel 3x3 conv feeding c_out channels (the extra input channels of the Decoder
// combiners: 1 in orca_modules, num_2d in orca_leukemia).  mode 0: src is (B,n_extra,S,S);
// mode 1/2: src is (B,n_extra,S/2,S/2) and is upsampled x2 on the fly (1 = nearest,
// 2 = bilinear align_corners=False).  w_extra is [n_extra][9][c_out].
int extra_channel_conv(const float* src, int64_t sB, int64_t sC, int64_t sH, int64_t sW, int n_extra,
                       const float* w_extra, float* out, int B, int S, int c_out, int mode, cudaStream_t s);
// final head: tmp[b][o][i][j] = W2[o] . relu(W1 v + b1) + b2[o] ; out = 0.5*(tmp + tmp^T), (B, O, S, S)
bool final_head_ok(const ConvLayer& f0, const ConvLayer& f1);
int final_head(const float* in /*[B][S][S][64]*/, const ConvLayer& f0, const ConvLayer& f1, float* tmp,
               float* out, int B, int S, cudaStream_t s);
// Net.final_1d second conv + sigmoid: out[b][k][l] = sigmoid(b[k] + sum_c w[c][k] * in[b][l][c])
int head_1d_sigmoid(const float* in, const ConvLayer& L, float* out, int B, int S, cudaStream_t s);
int copy_f32(const float* src, float* dst, int64_t n, cudaStream_t s);

int background_level(const double* normmat, int64_t n, int64_t r0, int64_t f, int64_t S, int flip,
                     float* out, double* out_mean, cudaStream_t s);
int background_assemble(const double* coord, const int* chrom, const double* cis, double trans, double binsize,
                        double* out, int64_t n, cudaStream_t s);

}  // namespace orca
