// fp32 CUDA-core implicit-GEMM convolution (channel-last), the exact-arithmetic path of
// liborca_b200: used for every layer when ORCA_B200_IMPL_SIMT is selected, for the layers
// the tcgen05 path does not cover, and as the on-device checker of the tensor-core kernels.
//
// One kernel covers Conv1d(k=9) (orca_modules.py:811-927, :991-1149), dilated 3x3 Conv2d
// (:22-459) and 1x1 convs: a tile of 128 output pixels x C_out channels, K loop over
// (tap, 16-channel slice), register-prefetch double buffering through shared memory.
// Epilogue: folded-BN bias, ReLU, up to two residual adds  --  out = act(conv + b) + res + res2.
#include "common.h"
#include "seq_in.cuh"

namespace orca {

struct ConvKArgs {
  const float* in;
  const float* w;
  const float* bias;
  const float* res;
  const float* res2;
  float* out;
  long long NP;  // B*H*W
  int H, W, c_in, kh, kw, dil, in_ld, out_ld, res_ld, relu;
};

template <int V>
struct VecT;
template <>
struct VecT<4> { using T = float4; };
template <>
struct VecT<2> { using T = float2; };

template <int BN, int V>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvKArgs a) {
  constexpr int BM = 128, BK = 16, J = BN / (16 * V), NB = J * V;
  constexpr int BLOADS = (4 * BN + 255) / 256;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long p0 = (long long)blockIdx.x * BM;
  const int HW = a.H * a.W;

  // A loader: this thread fetches channels [kv*4, kv*4+4) of the K-slice for two pixels
  const int kv = tid & 3;
  int lp[2], py[2], px[2];
  long long pg[2];
  bool pv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    lp[i] = (tid >> 2) + 64 * i;
    pg[i] = p0 + lp[i];
    pv[i] = pg[i] < a.NP;
    long long rem = pv[i] ? (pg[i] % HW) : 0;
    py[i] = (int)(rem / a.W);
    px[i] = (int)(rem - (long long)py[i] * a.W);
  }
  const int kcs = a.c_in >> 4;
  const int nk = a.kh * a.kw * kcs;

  float4 ra[2];
  float4 rb[BLOADS];
  float acc[8][NB];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[i][n] = 0.f;

  auto load_tiles = [&](int ks) {
    const int tap = ks / kcs, kc = ks - tap * kcs;
    const int th = tap / a.kw, tw = tap - th * a.kw;
    const int dy = (th - a.kh / 2) * a.dil, dx = (tw - a.kw / 2) * a.dil;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int yy = py[i] + dy, xx = px[i] + dx;
      const bool ok = pv[i] && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float* src = a.in + (pg[i] + (long long)dy * a.W + dx) * a.in_ld + kc * 16 + kv * 4;
        ra[i] = __ldg(reinterpret_cast<const float4*>(src));
      }
    }
#pragma unroll
    for (int j = 0; j < BLOADS; ++j) {
      const int idx = tid + j * 256;
      if (idx < 4 * BN) {
        const int k = idx / (BN / 4), n4 = idx - k * (BN / 4);
        const float* src = a.w + ((long long)(tap * a.c_in + kc * 16 + k)) * BN + n4 * 4;
        rb[j] = __ldg(reinterpret_cast<const float4*>(src));
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][kv * 4 + 0][lp[i]] = ra[i].x;
      As[buf][kv * 4 + 1][lp[i]] = ra[i].y;
      As[buf][kv * 4 + 2][lp[i]] = ra[i].z;
      As[buf][kv * 4 + 3][lp[i]] = ra[i].w;
    }
#pragma unroll
    for (int j = 0; j < BLOADS; ++j) {
      const int idx = tid + j * 256;
      if (idx < 4 * BN) {
        const int k = idx / (BN / 4), n4 = idx - k * (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][k][n4 * 4]) = rb[j];
      }
    }
  };

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int ks = 0; ks < nk; ++ks) {
    const int cur = ks & 1;
    if (ks + 1 < nk) load_tiles(ks + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[NB];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const typename VecT<V>::T t =
            *reinterpret_cast<const typename VecT<V>::T*>(&Bs[cur][k][tx * V + j * 16 * V]);
        const float* tp = reinterpret_cast<const float*>(&t);
#pragma unroll
        for (int v = 0; v < V; ++v) bv[j * V + v] = tp[v];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[i][n] = fmaf(av[i], bv[n], acc[i][n]);
    }
    if (ks + 1 < nk) store_tiles(cur ^ 1);
    __syncthreads();
  }

  // epilogue
  float bias[NB];
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int v = 0; v < V; ++v) bias[j * V + v] = __ldg(a.bias + tx * V + j * 16 * V + v);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long p = p0 + ty * 8 + i;
    if (p >= a.NP) continue;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = tx * V + j * 16 * V;
      typename VecT<V>::T o;
      float* op = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float t = acc[i][j * V + v] + bias[j * V + v];
        op[v] = a.relu ? fmaxf(t, 0.f) : t;
      }
      if (a.res) {
        const typename VecT<V>::T r = *reinterpret_cast<const typename VecT<V>::T*>(a.res + p * a.res_ld + c);
        const float* rp = reinterpret_cast<const float*>(&r);
#pragma unroll
        for (int v = 0; v < V; ++v) op[v] += rp[v];
      }
      if (a.res2) {
        const typename VecT<V>::T r = *reinterpret_cast<const typename VecT<V>::T*>(a.res2 + p * a.res_ld + c);
        const float* rp = reinterpret_cast<const float*>(&r);
#pragma unroll
        for (int v = 0; v < V; ++v) op[v] += rp[v];
      }
      *reinterpret_cast<typename VecT<V>::T*>(a.out + p * a.out_ld + c) = o;
    }
  }
}

int conv_simt(const ConvLayer& L, const ConvCall& c, cudaStream_t s) {
  if (L.c_in % 16 != 0 || c.in_ld % 4 != 0 || c.out_ld % 4 != 0 || (c.res && c.res_ld % 4 != 0)) {
    set_error("conv_simt: unsupported channel geometry c_in=%d in_ld=%d out_ld=%d", L.c_in, c.in_ld, c.out_ld);
    return ORCA_B200_EUNSUPPORTED;
  }
  ConvKArgs a;
  a.in = c.in; a.w = L.w; a.bias = L.b; a.res = c.res; a.res2 = c.res2; a.out = c.out;
  a.NP = (long long)c.B * c.H * c.W;
  a.H = c.H; a.W = c.W; a.c_in = L.c_in; a.kh = L.kh; a.kw = L.kw; a.dil = L.dil;
  a.in_ld = c.in_ld; a.out_ld = c.out_ld; a.res_ld = c.res_ld; a.relu = c.relu;
  if (a.NP <= 0) return ORCA_B200_OK;
  const long long tiles = (a.NP + 127) / 128;
  if (tiles > 0x7fffffffLL) { set_error("conv_simt: too many tiles"); return ORCA_B200_EUNSUPPORTED; }
  dim3 grid((unsigned)tiles), block(256);
  switch (L.c_out) {
    case 32: conv_igemm_kernel<32, 2><<<grid, block, 0, s>>>(a); break;
    case 64: conv_igemm_kernel<64, 4><<<grid, block, 0, s>>>(a); break;
    case 96: conv_igemm_kernel<96, 2><<<grid, block, 0, s>>>(a); break;
    case 128: conv_igemm_kernel<128, 4><<<grid, block, 0, s>>>(a); break;
    default:
      set_error("conv_simt: unsupported c_out=%d", L.c_out);
      return ORCA_B200_EUNSUPPORTED;
  }
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ------------------------------------------------------------------------------------------
// First encoder layer: Conv1d(4 -> 64, k=9, pad=4) + folded BN, straight from the caller's
// strided (B, 4, L) tensor (orca_modules.py:812-813).  Generic fp32 inputs (N bases are 0.25
// in all four channels, selene_utils2.py:216-230), so this is an FMA kernel, not a LUT.
// Block = 128 positions; thread = 8 consecutive positions x 4 output channels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_first_kernel(const SeqIn in, long long Ltot,
                                                         long long l_begin, long long n,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         float* __restrict__ out) {
  constexpr int TP = 128, HALO = 4, NX = TP + 2 * HALO;
  __shared__ __align__(16) float Xs[NX][4];
  __shared__ __align__(16) float Ws[9 * 4 * 64];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * TP;  // first output position of the tile (relative)

  for (int i = tid; i < 9 * 4 * 64; i += 256) Ws[i] = __ldg(w + i);
  for (int j = tid; j < NX; j += 256) {  // one position per thread: fp32 view or packed bases (seq_in.cuh)
    const long long l = l_begin + t0 - HALO + j;
    *reinterpret_cast<float4*>(&Xs[j][0]) = (l >= 0 && l < Ltot) ? seq_load(in, b, l) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  const int cg = tid & 15, pgp = tid >> 4;
  float4 xw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) xw[i] = *reinterpret_cast<const float4*>(&Xs[pgp * 8 + i][0]);
  float4 acc[8];
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cg * 4));
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = bv;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[(t * 4 + ci) * 64 + cg * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xv = reinterpret_cast<const float*>(&xw[i + t])[ci];
        acc[i].x = fmaf(xv, wv.x, acc[i].x);
        acc[i].y = fmaf(xv, wv.y, acc[i].y);
        acc[i].z = fmaf(xv, wv.z, acc[i].z);
        acc[i].w = fmaf(xv, wv.w, acc[i].w);
      }
    }
  }
  float* ob = out + (long long)b * n * 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = t0 + pgp * 8 + i;
    if (r < n) *reinterpret_cast<float4*>(ob + r * 64 + cg * 4) = acc[i];
  }
}

int conv_first_simt(const ConvLayer& L, const SeqIn& in, int B, int64_t Ltot, int64_t l_begin, int64_t n, float* out,
                    cudaStream_t s) {
  if (L.c_in != 4 || L.c_out != 64 || L.kh != 1 || L.kw != 9) {
    set_error("conv_first_simt: layer is not Conv1d(4,64,k=9)");
    return ORCA_B200_EINVAL;
  }
  if (n <= 0) return ORCA_B200_OK;
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)B), block(256);
  conv_first_kernel<<<grid, block, 0, s>>>(in, Ltot, l_begin, n, L.w, L.b, out);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
