// The encoder's input as the kernels see it: either the caller's strided fp32 (B, 4, L) one-hot view
// (orca_predict.py:333-337 hands over a transposed view, strides (4L, 1, 4)) or PACKED bases, one byte per
// position (1 B/bp instead of 16 B/bp over PCIe and out of HBM).
//
// Packed bytes are either codes 0..4 (A, C, G, T, N) or raw ASCII as it sits in a FASTA record: 'A','C','G','T'
// in either case -> one-hot, anything else -> 0.25 in all four channels, which is what the reference's feeder
// produces (selene_utils2.py:125-128 via selene_sdk's sequence_to_encoding; pads are 0.25 too, :216-222).
// The reverse-complement strand (orca_predict.py:324-329: sequence[:, ::-1, ::-1]) is the same buffer walked
// with a negative position stride and the code complemented (A<->T, C<->G, N stays).
#pragma once
#include <stdint.h>

namespace orca {

struct SeqIn {
  const float* x = nullptr;        // fp32 view: element strides sB, sC, sL (may be negative)
  const uint8_t* bases = nullptr;  // packed: byte strides sB, sL (sL may be negative); sC unused
  long long sB = 0, sC = 0, sL = 0;
  int complement = 0;              // packed only
};

#ifdef __CUDACC__
__device__ __forceinline__ int seq_base_code(unsigned c) {
  if (c < 5u) return (int)c;
  c &= 0xDFu;  // fold ASCII lower case onto upper case
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

// the four channel values of position p (virtual position: the base pointers address position 0) of sample b
__device__ __forceinline__ float4 seq_load(const SeqIn& in, int b, long long p) {
  if (in.bases) {
    int code = seq_base_code(__ldg(in.bases + (long long)b * in.sB + p * in.sL));
    if (code > 3) return make_float4(0.25f, 0.25f, 0.25f, 0.25f);
    if (in.complement) code = 3 - code;
    return make_float4(code == 0 ? 1.f : 0.f, code == 1 ? 1.f : 0.f, code == 2 ? 1.f : 0.f, code == 3 ? 1.f : 0.f);
  }
  const float* xb = in.x + (long long)b * in.sB;
  const long long sC = in.sC, sL = in.sL;
  const bool vec_ok = ((sL & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb + (sC == -1 ? -3 : 0)) & 15) == 0);
  if (sC == 1 && vec_ok)  // channel-last memory: one 16-byte load per position
    return __ldg(reinterpret_cast<const float4*>(xb + p * sL));
  if (sC == -1 && vec_ok) {  // reverse-complement walk of channel-last memory: channels stored descending
    const float4 r = __ldg(reinterpret_cast<const float4*>(xb + p * sL - 3));
    return make_float4(r.w, r.z, r.y, r.x);
  }
  return make_float4(__ldg(xb + p * sL), __ldg(xb + p * sL + sC), __ldg(xb + p * sL + 2 * sC), __ldg(xb + p * sL + 3 * sC));
}
#endif

}  // namespace orca
