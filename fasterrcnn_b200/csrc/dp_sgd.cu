// Data-parallel optimizer step as ONE kernel per rank over NVLink / NVSwitch (EXPERIMENT, opt-in -- optim.NvlsShardedSGD):
// reduce-scatter of the weight gradients + torch.optim.SGD update + all-gather of the updated weights, fused.
//
//   every rank holds the optimizer's tensors in two flat symmetric-memory arenas (same layout on every rank): weights W, gradients G.
//   rank r owns the contiguous shard [r * S, (r + 1) * S) of the flat index space and, for every element i of its shard:
//     g     = sum over ranks of G_rank[i]          multimem.ld_reduce.add on the multicast address: the sum is formed INSIDE the NVSwitch
//                                                   (or, without multicast support, peer-to-peer loads from every rank's arena)
//     p, m  = SGD(W[i], g / world, M[i - r * S])    the momentum buffer exists only for the own shard: 1 / world of its memory and traffic
//     W_rank[i] = p for every rank                  multimem.st on the multicast address (or peer-to-peer stores)
//   so the gradients cross the wire once (reduced on the way), the weights once, and no rank ever runs an optimizer pass over more than
//   its shard.  The caller brackets the launch with two cross-rank barriers (all gradients written | all weights delivered and all
//   gradients consumed); both are stream-ordered.
#include "common.cuh"

namespace frcnn {

constexpr int kMaxPeers = 8;

struct DpPeers {
  const float *grad[kMaxPeers];      // every rank's gradient arena (index = rank), device pointers valid on this rank
  float *weight[kMaxPeers];          // every rank's weight arena
  int world;
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float *mc)
{
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}

__device__ __forceinline__ void multimem_st(float *mc, float4 v)
{
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// torch.optim.SGD, one element (momentum, dampening 0, no nesterov, L2 weight decay folded into the gradient): the roundings of
// elementwise.cu:sgd_one -- separate multiplies and adds, no FMA contraction
__device__ __forceinline__ float dp_sgd_one(float p, float g, float &buf, float lr, float mom, float wd, float gs, int first)
{
  float gg = __fmul_rn(g, gs);
  if (wd != 0.f) gg = __fadd_rn(gg, __fmul_rn(wd, p));
  const float b = first ? gg : __fadd_rn(__fmul_rn(buf, mom), gg);
  buf = b;
  return __fadd_rn(p, __fmul_rn(-lr, b));
}

// MC: reduce / broadcast through the multicast mapping; else through the peer pointers.  All offsets in float4 units.
// A thread keeps UNROLL independent float4 reductions in flight (a multimem.ld_reduce over NVLink has microseconds of latency: bytes in
// flight, not threads, set the rate).  Launch shapes: many CTAs per SM when the kernel has the GPU to itself, ONE 256-thread CTA per SM
// (~80 registers per thread) when it runs UNDER the convolution backward -- that CTA fits beside a resident GEMM CTA, which is launched
// with 152 registers per thread for exactly this purpose (conv_tc.cu), so neither kernel waits for the other's SM slots.
template <bool MC, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
dp_sgd_kernel(const float *__restrict__ grad_mc, float *__restrict__ weight_mc, DpPeers peers, const float *__restrict__ weight_local,
              float *__restrict__ momentum_shard, size_t shard_begin4, size_t shard_count4, float lr, float mom, float wd, float gs, int first)
{
  pdl_enter();
  const size_t stride = (size_t)gridDim.x * THREADS;
  for (size_t k0 = blockIdx.x * (size_t)THREADS + threadIdx.x; k0 < shard_count4; k0 += stride * UNROLL) {
    float4 g[UNROLL], p[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t k = k0 + (size_t)u * stride;
      if (k < shard_count4) {
        const size_t i = shard_begin4 + k;
        if (MC) {
          g[u] = multimem_ld_reduce_add(grad_mc + 4 * i);
        } else {
          g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int r = 0; r < peers.world; r++) {                    // fixed rank order: every rank would form the same sum
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(peers.grad[r]) + i);
            g[u].x += v.x; g[u].y += v.y; g[u].z += v.z; g[u].w += v.w;
          }
        }
        p[u] = __ldcs(reinterpret_cast<const float4 *>(weight_local) + i);
        b[u] = first ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcs(reinterpret_cast<const float4 *>(momentum_shard) + k);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t k = k0 + (size_t)u * stride;
      if (k < shard_count4) {
        const size_t i = shard_begin4 + k;
        p[u].x = dp_sgd_one(p[u].x, g[u].x, b[u].x, lr, mom, wd, gs, first);
        p[u].y = dp_sgd_one(p[u].y, g[u].y, b[u].y, lr, mom, wd, gs, first);
        p[u].z = dp_sgd_one(p[u].z, g[u].z, b[u].z, lr, mom, wd, gs, first);
        p[u].w = dp_sgd_one(p[u].w, g[u].w, b[u].w, lr, mom, wd, gs, first);
        __stcs(reinterpret_cast<float4 *>(momentum_shard) + k, b[u]);
        if (MC) {
          multimem_st(weight_mc + 4 * i, p[u]);
        } else {
          for (int r = 0; r < peers.world; r++) __stcg(reinterpret_cast<float4 *>(peers.weight[r]) + i, p[u]);
        }
      }
    }
  }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" int frcnn_dp_sgd_fused(const float *grad_multicast, float *weight_multicast, const void *const *grad_peers, void *const *weight_peers, int world,
                                  const float *weight_local, float *momentum_shard, size_t shard_begin, size_t shard_count,
                                  float lr, float momentum, float weight_decay, float grad_scale, int first_step, int ctas_per_sm, void *stream)
{
  FRCNN_REQUIRE(weight_local && momentum_shard && world >= 1 && world <= kMaxPeers, "dp_sgd_fused: bad argument");
  FRCNN_REQUIRE((shard_begin % 4) == 0 && (shard_count % 4) == 0, "dp_sgd_fused: shard bounds must be multiples of 4 elements");
  const bool mc = grad_multicast != nullptr && weight_multicast != nullptr;
  FRCNN_REQUIRE(mc || (grad_peers && weight_peers), "dp_sgd_fused: need the multicast pointers or the peer pointer tables");
  FRCNN_REQUIRE(((reinterpret_cast<uintptr_t>(grad_multicast) | reinterpret_cast<uintptr_t>(weight_multicast) | reinterpret_cast<uintptr_t>(weight_local) |
                  reinterpret_cast<uintptr_t>(momentum_shard)) & 15) == 0, "dp_sgd_fused: pointers must be 16-byte aligned");
  if (shard_count == 0) return FRCNN_OK;
  DpPeers peers;
  memset(&peers, 0, sizeof(peers));
  peers.world = world;
  if (!mc) {
    for (int r = 0; r < world; r++) {
      peers.grad[r] = reinterpret_cast<const float *>(grad_peers[r]);
      peers.weight[r] = reinterpret_cast<float *>(weight_peers[r]);
      FRCNN_REQUIRE(peers.grad[r] && peers.weight[r] && ((reinterpret_cast<uintptr_t>(peers.grad[r]) | reinterpret_cast<uintptr_t>(peers.weight[r])) & 15) == 0,
                    "dp_sgd_fused: null or misaligned peer pointer");
    }
  }
  const size_t n4 = shard_count / 4;
  cudaStream_t st = as_stream(stream);
  // 256 threads x 4 float4 in flight per thread; ctas_per_sm = 1 is the shape that runs UNDER the convolution backward (one such CTA,
  // <= 104 registers per thread, fits beside a resident GEMM CTA: conv_tc.cu kTcLaunchRegs), 0 = 8 per SM when the kernel runs alone
  const int grid = elementwise_grid(ceil_div<size_t>(n4, 4), 256, ctas_per_sm > 0 ? ctas_per_sm : 8);
  if (mc) launch(dp_sgd_kernel<true, 256, 4>, grid, 256, 0, st, grad_multicast, weight_multicast, peers, weight_local, momentum_shard, shard_begin / 4, n4, lr, momentum, weight_decay, grad_scale, first_step);
  else launch(dp_sgd_kernel<false, 256, 4>, grid, 256, 0, st, grad_multicast, weight_multicast, peers, weight_local, momentum_shard, shard_begin / 4, n4, lr, momentum, weight_decay, grad_scale, first_step);
  FRCNN_CHECK_LAUNCH("dp_sgd_kernel");
  return FRCNN_OK;
}
