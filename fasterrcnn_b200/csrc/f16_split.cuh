// fp16 engine operand format (FRCNN_ENGINE_TC_3XF16), shared by the producers of operand splits (conv_tc.cu, elementwise.cu).
//
//   x * 2^e = hi + lo / 2048,   hi = fp16(x * 2^e),   lo = fp16((x * 2^e - hi) * 2048)
//
// Buffer = [header 4096 B | hi (count fp16, padded to 1024 B) | lo (same)].  Header words (32-bit):
//   [0] bit pattern of the tensor's max |x| the exponent was derived from     [1] e
//   [16 .. 16 + G) per-block partial maxima of the amax pass (G <= kF16MaxPartials), reduced by every block of the consumer
// e puts max |x| * 2^e in [2^13, 2^14): both halves keep 11 significant bits in fp16's normal range for every element within
// 2^-27 of the maximum; values are saturated to +-65504 so that a producer working from a stale exponent cannot emit inf.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace frcnn {

constexpr int kF16Header = 4096;
constexpr int kF16PartialsAt = 16;           // first header word of the partial maxima
constexpr int kF16MaxPartials = 592;         // 4 x 148 blocks of 512 threads
constexpr int kF16LoShift = 11;              // lo is stored multiplied by 2^11

inline size_t f16_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline size_t f16_half_bytes(size_t count) { return f16_align_up(count * 2, 1024); }

// e such that amax * 2^e lies in [2^13, 2^14); clamped so that 2^-(e_a + e_b) stays a normal float
__device__ __forceinline__ int f16_exponent(unsigned amax_bits)
{
  if (amax_bits == 0u) return 0;
  int e = 140 - (int)((amax_bits >> 23) & 0xffu);
  return e > 60 ? 60 : (e < -60 ? -60 : e);
}

__device__ __forceinline__ float pow2i(int e) { return __int_as_float((e + 127) << 23); }

__device__ __forceinline__ void split16(float x, float s, __half &hi, __half &lo)
{
  float xs = __fmul_rn(x, s);                                         // exact (power of two) unless it underflows
  xs = xs > 65504.0f ? 65504.0f : (xs < -65504.0f ? -65504.0f : xs);   // comparisons, not fmin / fmax: a NaN stays a NaN and reaches the output
  hi = __float2half_rn(xs);
  lo = __float2half_rn(__fmul_rn(__fsub_rn(xs, __half2float(hi)), 2048.0f));
}

__device__ __forceinline__ void split16x4(const float4 &v, float s, uint2 &hi, uint2 &lo)
{
  __half h[4], l[4];
  split16(v.x, s, h[0], l[0]); split16(v.y, s, h[1], l[1]); split16(v.z, s, h[2], l[2]); split16(v.w, s, h[3], l[3]);
  hi = *reinterpret_cast<const uint2 *>(h);
  lo = *reinterpret_cast<const uint2 *>(l);
}

// every thread of the block gets the maximum of the G partial maxima (bit patterns of non-negative floats: integer order);
// scratch = 32 words of shared memory
__device__ __forceinline__ unsigned f16_reduce_partials(const unsigned *__restrict__ header, int G, unsigned *scratch)   // header: words [16, 16 + G) are read
{
  unsigned m = 0u;
  for (int i = threadIdx.x; i < G; i += blockDim.x) m = max(m, __ldg(header + kF16PartialsAt + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = m;
  __syncthreads();
  const int warps = (blockDim.x + 31) >> 5;
  unsigned r = 0u;
  for (int i = 0; i < warps; i++) r = max(r, scratch[i]);
  __syncthreads();
  return r;
}

// amax pass: block b writes the maximum |x| of its grid-stride share to header[kF16PartialsAt + b] (plain store: no initialisation)
__global__ void f16_amax_partials_kernel(const float *__restrict__ x, size_t count, unsigned *__restrict__ header);
int f16_amax_grid(size_t count);
// launches the amax pass; returns G
int f16_launch_amax(const float *x, size_t count, void *header, cudaStream_t st);

}  // namespace frcnn
