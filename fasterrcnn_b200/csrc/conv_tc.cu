// tcgen05 engine (FRCNN_ENGINE_TC_3XTF32): implicit-GEMM convolution -- forward, data gradient and
// filter gradient -- on the 5th-gen tensor cores.
//
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) straight from the NHWC fp32
//     tensors.  For filter tap (kh,kw) the activation tile of an output patch is the input patch
//     shifted by (kh-pad, kw-pad); the TMA unit zero-fills the part outside the image, so im2col
//     and padding are folded into the shared-memory staging and never exist in HBM;
//   * tcgen05.mma kind::tf32 issued by one thread, fp32 accumulators in TMEM (128 lanes x BN cols);
//   * fp32-grade accuracy through the error-compensated split x = hi + lo (hi = the tf32 truncation
//     the tensor core applies itself, lo = x - hi): D += A_hi*B_hi + (A_hi*B_lo + A_lo*B_hi).  The B_hi and
//     B_lo tiles sit back to back in shared memory, so A_hi * [B_hi | B_lo] is ONE N = 2*BN instruction (A_hi
//     is read once) into accumulator columns [main | corr]; A_lo * B_hi accumulates into the corr columns;
//   * the tensor core's own fp32 accumulation is the remaining error source on long K chains, so
//     chains are kept short: every kChunkKB k-blocks (48 MMAs) the accumulator -- double-buffered in
//     TMEM -- is drained by the epilogue warps into fp32 registers (round-to-nearest adds on the CUDA
//     cores) while the issuer already fills the other buffer;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 =
//     accumulate/epilogue (tcgen05.ld -> register sums -> scale/bias/residual/activation -> 128-bit
//     stores); STAGES-deep mbarrier ring (full/empty) for operands, 2-deep ring (acc_full/acc_empty)
//     for accumulators, tcgen05.commit signals both.
//
// GEMM views (M x N x K), all with 128 x BN x 32 tiles:
//   FWD    pixels x Cout x (tap,ci) : A = x patch   (K-major, 4-D map)   B = w[co][(tap,ci)] (K-major, 2-D map)
//   DGRAD  pixels x Cin  x (tap,co) : A = dy patch  (K-major, flipped tap) B = w[co][tap][ci] (N-major, 3-D map)
//   WGRAD  Cout   x Cin  x pixels   : A = dy patch  (M-major)            B = shifted x patch (N-major), one tap per CTA
//
// Second engine, FRCNN_ENGINE_TC_3XF16 (template parameter F16): the same kernel on kind::f16 -- twice the tensor-pipe rate and half
// the operand bytes per multiply-accumulate.  Operands are the per-tensor power-of-two scaled split  x * 2^e = hi + lo / 2048  with
// hi = fp16(x * 2^e), lo = fp16((x * 2^e - hi) * 2048) (e chosen from the tensor's absolute maximum so that |hi| < 2^14: both halves
// keep 11 significant bits in the fp16 normal range); the three products and the [main | corr] accumulator columns are those of the
// tf32 scheme, the drain folds corr / 2048 and the 2^-(e_a + e_b) rescale into its two FMAs (powers of two: exact).  A k-block is
// 64 elements (one 128-byte swizzle row of fp16), MN-major operands use the plain SWIZZLE_128B layout with 64-channel atoms.
#include <stdlib.h>
#include <atomic>
#include "common.cuh"
#include "tc_common.cuh"
#include "f16_split.cuh"

namespace frcnn {

using namespace tc;

// conv_simt.cu
struct Epilogue {
  const float *scale;
  const float *bias;
  const float *residual;
  int act;
};
int launch_splitk_reduce(const float *partial, float *out, int M, int Nn, int splits, const Epilogue &epi, cudaStream_t st);

// ---- x = hi + lo: hi = x rounded to tf32 (exactly representable, so the tensor core's own fp32->tf32
// conversion -- whatever its rounding -- is the identity), lo = x - hi (exact in fp32, <= 13 bits) ----
__device__ __forceinline__ float tf32_rna(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void split_hi_lo_kernel(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, size_t count)
{
  pdl_enter();
  size_t n4 = count / 4;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
    float4 h, r;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    r.x = v.x - h.x; r.y = v.y - h.y; r.z = v.z - h.z; r.w = v.w - h.w;
    reinterpret_cast<float4 *>(hi)[i] = h;
    reinterpret_cast<float4 *>(lo)[i] = r;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) {
    float h = tf32_rna(x[i]);
    hi[i] = h;
    lo[i] = x[i] - h;
  }
}

// ---- fp16 engine operand split (format: f16_split.cuh): amax pass (per-block partial maxima, no initialisation needed), then the
// split pass, whose every block reduces the partials to the tensor's exponent ----
__global__ void __launch_bounds__(512)
f16_amax_partials_kernel(const float *__restrict__ x, size_t count, unsigned *__restrict__ header)
{
  pdl_enter();
  __shared__ float red[16];
  const size_t n4 = count / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 16; i++) m = fmaxf(m, red[i]);
    header[kF16PartialsAt + blockIdx.x] = __float_as_uint(m);
  }
}

int f16_amax_grid(size_t count)
{
  size_t want = ceil_div<size_t>(count / 4 + 1, 512 * 4);             // >= 4 float4 per thread before adding blocks
  if (want > (size_t)kF16MaxPartials) want = kF16MaxPartials;
  return want < 1 ? 1 : (int)want;
}

int f16_launch_amax(const float *x, size_t count, void *header, cudaStream_t st)
{
  const int G = f16_amax_grid(count);
  launch(f16_amax_partials_kernel, G, 512, 0, st, x, count, reinterpret_cast<unsigned *>(header));
  return G;
}

__global__ void __launch_bounds__(256)
split_f16_kernel(const float *__restrict__ x, unsigned *__restrict__ header, const unsigned *__restrict__ partials, int G, __half *__restrict__ hi, __half *__restrict__ lo,
                 size_t count)
{
  pdl_enter();
  __shared__ unsigned scratch[32];
  const unsigned amax = f16_reduce_partials(partials, G, scratch);
  const int e = f16_exponent(amax);
  if (blockIdx.x == 0 && threadIdx.x == 0) { header[0] = amax; header[1] = (unsigned)e; }
  const float s = pow2i(e);
  const size_t n4 = count / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    uint2 h, l;
    split16x4(__ldg(reinterpret_cast<const float4 *>(x) + i), s, h, l);
    reinterpret_cast<uint2 *>(hi)[i] = h;
    reinterpret_cast<uint2 *>(lo)[i] = l;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) split16(x[i], s, hi[i], lo[i]);
}

// split with the exponent the buffer already carries (header word 1): ONE pass, for tensors that change slowly (weights after an optimizer
// step whose update kernel could not write the split itself: the sharded data-parallel step).  Values are saturated, see f16_split.cuh.
__global__ void __launch_bounds__(256)
split_f16_carried_kernel(const float *__restrict__ x, const unsigned *__restrict__ header, __half *__restrict__ hi, __half *__restrict__ lo, size_t count)
{
  pdl_enter();
  const float s = pow2i((int)__ldcg(header + 1));
  const size_t n4 = count / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    uint2 h, l;
    split16x4(__ldcs(reinterpret_cast<const float4 *>(x) + i), s, h, l);
    reinterpret_cast<uint2 *>(hi)[i] = h;
    reinterpret_cast<uint2 *>(lo)[i] = l;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) split16(x[i], s, hi[i], lo[i]);
}

int f16_split_carried(const float *x, size_t count, void *out, int ctas_per_sm, cudaStream_t st)
{
  uint8_t *o = reinterpret_cast<uint8_t *>(out);
  __half *hi = reinterpret_cast<__half *>(o + kF16Header);
  __half *lo = reinterpret_cast<__half *>(o + kF16Header + f16_half_bytes(count));
  launch(split_f16_carried_kernel, elementwise_grid(count / 4 + 1, 256, ctas_per_sm > 0 ? ctas_per_sm : 8), 256, 0, st, x, reinterpret_cast<const unsigned *>(o), hi, lo, count);
  FRCNN_CHECK_LAUNCH("split_f16_carried_kernel");
  return FRCNN_OK;
}

// ---- kernel ---------------------------------------------------------------------------------------
enum { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2 };

struct TcGeom {
  int Cin, Cout, KH, KW, pad;
  int H, W, nimg;             // spatial size (input == output: stride-1 "same" convs; linear: 1 x rows)
  int tile_w, tile_h, tile_n; // FWD/DGRAD: output patch of one CTA (tile_w * tile_h * tile_n == 128; tile_n images)
  int tiles_w, tiles_h;       //            patches per image group
  int pw, ph, pn;             // WGRAD: pixel patch of one k-block (pw * ph * pn == 32)
  int patches_w, patches_h;
  int kb_per_split, total_kb, splits;
  int mn_lbo, mn_sbo, mn_kstep;   // MN-major operand descriptor strides in bytes (4096 / 512 / 1024)
  int m_tiles, n_tiles, items;    // work items = m_tiles * n_tiles * (taps for WGRAD) * splits; CTAs loop over them (persistent grid)
  int m_pairs;                    // CTA-pair kernels: ceil(m_tiles / 2) -- a work item is then TWO vertically adjacent 128-row tiles (rank 0 / 1)
  int pair_late_trigger;          // pair + stream-K: griddepcontrol.launch_dependents after the work loop instead of at the top
  int paired_stores;              // epilogue: lane pairs write whole 32-byte sectors (see the store loop)
  unsigned long long *trace;      // debug: per-CTA %globaltimer stamps (frcnn_debug_tc_trace), NULL in production
  // stream-K (streamk != 0): the launch's (tile, k-block) units are cut into gridDim.x equal contiguous ranges, one per CTA, so
  // every SM gets the same number of k-blocks whatever the tile count.  A CTA whose range starts inside a tile parks its raw
  // partial sums in sk_slots[cta] and publishes sk_flags[cta] = sk_tag; the CTA that owns the tile's first k-block adds them.
  int streamk;
  long long units;                // tiles * total_kb
  float *sk_slots;                // gridDim.x x (128 x BN) fp32
  unsigned long long *sk_flags;   // gridDim.x
  unsigned long long sk_tag;      // unique per launch (stale workspace contents can never match)
  const int *a_exp, *b_exp;       // fp16 engine: device words holding the operands' scale exponents (NULL for tf32)
  unsigned *amax_out;             // optional: CTA b stores the bit pattern of max |final output| over its tiles at amax_out[kF16PartialsAt + b]
  int amax_slots;                 // slots the consumer reads (the grid of an unreserved launch); CTA 0 zero-fills [gridDim.x, amax_slots)
};

// one work item of the persistent loop: an output tile (or one split-K slice of it)
enum { TC_ITEM_FULL = 0, TC_ITEM_HEAD = 1, TC_ITEM_PART = 2 };

struct TcItem {
  int img, oh0, ow0, m0, n0, tap_w, split, kb_begin, nkb;
  int kind;                       // FULL: whole K range here | HEAD: first k-blocks, finishes the tile | PART: later k-blocks, parked
  long long tile_end;             // stream-K: first unit after this tile
};

template <int MODE, int BN, bool PAIR>
__device__ __forceinline__ TcItem tc_decode_item(const TcGeom &g, int w, int rank)
{
  TcItem it;
  const int mdiv = PAIR ? g.m_pairs : g.m_tiles;
  const int mt = PAIR ? 2 * (w % mdiv) + rank : w % mdiv;       // (a pair's second tile may lie past the last one: TMA zero-fills, stores are masked)
  int r = w / mdiv;
  const int nt = r % g.n_tiles;
  r /= g.n_tiles;                                    // FWD/DGRAD: split; WGRAD: tap * splits + split
  it.n0 = nt * BN;
  it.img = it.oh0 = it.ow0 = it.m0 = it.tap_w = 0;
  if (MODE == TC_WGRAD) {
    it.m0 = mt * 128;
    it.tap_w = r / g.splits;
    it.split = r - it.tap_w * g.splits;
  } else {
    const int tiles_per_img = g.tiles_w * g.tiles_h;
    it.img = (mt / tiles_per_img) * g.tile_n;        // first image of this tile's image group
    const int trem = mt % tiles_per_img;
    it.oh0 = (trem / g.tiles_w) * g.tile_h;
    it.ow0 = (trem % g.tiles_w) * g.tile_w;
    it.split = r;
  }
  it.kb_begin = it.split * g.kb_per_split;
  int kb_end = it.kb_begin + g.kb_per_split;
  if (kb_end > g.total_kb) kb_end = g.total_kb;
  it.nkb = kb_end - it.kb_begin;
  it.kind = TC_ITEM_FULL;
  it.tile_end = 0;
  return it;
}

// walks the work of one CTA: split-K mode -> items blockIdx.x, +gridDim.x, ...; stream-K mode -> units [cta*U/G, (cta+1)*U/G)
struct TcCursor {
  long long pos, end;
};

// work groups: CTAs (gid = blockIdx.x of G = gridDim.x) or, for the CTA-pair kernels, clusters (gid = %clusterid.x of G = gridDim.x / 2);
// both CTAs of a pair walk the same range
__device__ __forceinline__ long long tc_sk_start(const TcGeom &g, int gid, int G) { return (long long)gid * g.units / (long long)G; }

__device__ __forceinline__ TcCursor tc_cursor(const TcGeom &g, int gid, int G)
{
  TcCursor c;
  if (g.streamk) { c.pos = tc_sk_start(g, gid, G); c.end = tc_sk_start(g, gid + 1, G); }
  else { c.pos = gid; c.end = g.items; }
  return c;
}

template <int MODE, int BN, bool PAIR>
__device__ __forceinline__ bool tc_next(const TcGeom &g, TcCursor &c, TcItem &t, int rank, int G)
{
  if (c.pos >= c.end) return false;
  if (!g.streamk) {
    t = tc_decode_item<MODE, BN, PAIR>(g, (int)c.pos, rank);
    c.pos += G;
    return true;
  }
  const int tile = (int)(c.pos / g.total_kb);
  t = tc_decode_item<MODE, BN, PAIR>(g, tile, rank);   // splits == 1 in this mode: the tile's coordinates (and tap)
  t.kb_begin = (int)(c.pos - (long long)tile * g.total_kb);
  t.tile_end = (long long)(tile + 1) * g.total_kb;
  const long long stop = t.tile_end < c.end ? t.tile_end : c.end;
  t.nkb = (int)(stop - c.pos);
  t.kind = t.kb_begin > 0 ? TC_ITEM_PART : (t.nkb == g.total_kb ? TC_ITEM_FULL : TC_ITEM_HEAD);
  c.pos = stop;
  return true;
}

__device__ __forceinline__ unsigned long long tc_globaltimer()
{
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
#define TC_TRACE(slot) do { if (g.trace) g.trace[(size_t)blockIdx.x * 16 + (slot)] = tc_globaltimer(); } while (0)

// Two warpgroups: warps 0-3 = accumulate / epilogue (warp w owns TMEM lane quarter w), warps 4-7 = TMA producer (A operand), MMA issuer
// + TMEM allocator, TMA producer (B operand), spare.  The kernel is LAUNCHED with kTcLaunchRegs registers per thread; the producer
// warpgroup then shrinks to kTcProducerRegs and the epilogue warpgroup (128 fp32 partial sums per thread) grows to kTcEpilogueRegs
// (setmaxnreg: 128 * 248 + 128 * 56 = 256 * 152).  A launch that reserved 255 registers for every thread would fill the register file
// of every SM sub-partition (2 warps x 8192 of 16384) and nothing else could become resident next to a GEMM CTA; at 152 each
// sub-partition keeps 6656 registers free -- room for the 256-thread CTAs of the optimizer / gradient-exchange kernels that run UNDER the
// convolution backward on a side stream (optim.FusedSGD(eager), optim.NvlsShardedSGD).
constexpr int kTcThreads = 256;
constexpr int kTcLaunchRegs = 152, kTcProducerRegs = 56, kTcEpilogueRegs = 248;
constexpr int kWarpTmaA = 4, kWarpMma = 5, kWarpTmaB = 6;
constexpr int kBK = 32;                       // fp32 elements per 128-byte swizzle row
constexpr int kABytes = 128 * kBK * 4;        // 16 KB: one 128 x 32 A tile
constexpr int kAtomBytes = 32 * kBK * 4;      // 4 KB: 32 x 32 fp32 block (one MN-major 32-column atom x 32 k-rows)
constexpr int kChunkKB = 8;                   // k-blocks (32 accumulations per column) inside the tensor core before a register drain
constexpr int kChunkKB16 = 4;                 // fp16 engine: a k-block holds 64 accumulations per column -> same chain length

__device__ __forceinline__ float tc_act(float v, int act)
{
  if (act == FRCNN_ACT_RELU) return v > 0.0f ? v : 0.0f;
  if (act == FRCNN_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

// PAIR (fp16 engine only): the kernel runs as clusters of two CTAs on cta_group::2 -- one M = 256 MMA per pair.  Each CTA stages its own
// 128 rows of A (hi, lo) and HALF of the B tile's columns (hi, lo: BN / 2 rows each); per k-step the leader issues three M = 256, N = BN
// MMAs:  main += A_hi * B_hi,  corr += A_hi * B_lo,  corr += A_lo * B_hi  (the single-CTA kernel's N = 2 BN instruction cannot be kept:
// a pair's B operand is split by COLUMNS across the CTAs, so [B_hi | B_lo] would put the main and corr columns in different CTAs' halves).
// Shared-memory traffic per k-block and CTA drops from 144 KB (64 written by TMA + 80 read by the tensor core) to 120 KB (48 + 72) for the
// same tensor work, and a stage is 48 KB instead of 64 KB (one more stage in flight).
template <int MODE, int BN, int STAGES, bool F16, bool PAIR = false>
__global__ void __maxnreg__(kTcLaunchRegs)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               TcGeom g, float *__restrict__ out, float *__restrict__ partial, Epilogue epi)
{
  static_assert(!PAIR || F16, "the CTA-pair variant exists for the fp16 engine");
  constexpr int kBRows = PAIR ? BN / 2 : BN;   // B columns (= rows of the staged tile) per CTA
  constexpr int kBBytes = kBRows * kBK * 4;
  constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  constexpr bool kAMajorMN = (MODE == TC_WGRAD);
  constexpr bool kBMajorMN = (MODE != TC_FWD);
  constexpr int kElems = F16 ? 64 : 32;        // operand elements per 128-byte row = K extent of a k-block = channels of an MN-major atom
  constexpr int kChunk = F16 ? kChunkKB16 : kChunkKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * kStageBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;        // [2]
  uint64_t *acc_empty = acc_full + 2;         // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KHW = g.KH * g.KW;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;               // 0 = leader: owns the `full` / `acc_empty` barriers and issues the MMAs
  const int gid = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;     // work-group index (see tc_cursor)
  const int G = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  // The next kernel's CTAs may become resident behind this one.  A pair kernel whose CTAs spin on each other (stream-K) triggers LATE, after
  // its work loop: with the early trigger the bench hung in round 2 whenever 2-CTA clusters, the stream-K spin protocol and programmatic
  // dependent launch met (neither alone, nor any two of them); FRCNN_TC_PAIR_TRIGGER=early restores it for experiments.
  const bool late_trigger = PAIR && g.streamk && g.pair_late_trigger;
  if (!late_trigger) pdl_trigger();
  const unsigned long long t_entry = (g.trace && threadIdx.x == 0) ? tc_globaltimer() : 0ull;   // stored after the wait: no global access before it
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }      // full: one arrive.expect_tx per producer
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], PAIR ? 8 : 4); }   // pair: the epilogue warps of both CTAs
    fence_barrier_init();
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi); tma_prefetch_desc(&map_b_lo);
  }
  if (warp == kWarpMma) {
    __syncwarp();
    if (PAIR) tmem_alloc_pair(tmem_slot, 4 * BN);                    // (issued by warp 1 of BOTH CTAs: same columns in both TMEMs)
    else tmem_alloc(tmem_slot, 4 * BN);                              // 2 accumulator buffers x [main | corr] columns
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                                      // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                                        // set-up above touched no global memory: it overlaps the previous kernel's tail

  if (threadIdx.x == 0 && g.trace) { g.trace[(size_t)blockIdx.x * 16] = t_entry; TC_TRACE(1); }
  if (warp >= 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTcProducerRegs));      // whole warpgroup, before its roles diverge
  if (warp == kWarpTmaA || warp == kWarpTmaB) {
    if (lane == 0) {
      // ===== TMA producers: warp 0 feeds the A operand, warp 6 the B operand (up to 8 box loads each per k-block: issuing them
      // from one thread was the limit of the filter-gradient mainloop).  Both run ahead across work items. =====
      const bool feed_a = (warp == kWarpTmaA);
      const int kblocks_c = (MODE == TC_FWD ? g.Cin : g.Cout) / kElems;       // channel blocks per tap (FWD/DGRAD)
      const int patches_per_img = g.patches_w * g.patches_h;
      int it = 0;                                                            // k-blocks issued by this CTA so far (ring position)
      TcCursor cur = tc_cursor(g, gid, G);
      TcItem t;
      while (tc_next<MODE, BN, PAIR>(g, cur, t, rank, G)) {
        const int nb0 = t.n0 + rank * kBRows;                                  // first B column staged by this CTA (pair: its half)
        for (int i = 0; i < t.nkb; i++, it++) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          const int kb = t.kb_begin + i;
          uint8_t *a_hi = smem + s * kStageBytes;
          uint8_t *a_lo = a_hi + kABytes;
          uint8_t *b_hi = a_hi + 2 * kABytes;
          uint8_t *b_lo = b_hi + kBBytes;
          // pair: every load of both CTAs is counted on the LEADER's barrier, which its two producers arm with the bytes of both CTAs
          const uint32_t fbar = PAIR ? mapa_u32(smem_u32(&full[s]), 0) : 0u;
          if (!PAIR) mbar_expect_tx(&full[s], feed_a ? 2 * kABytes : 2 * kBBytes);
          else if (rank == 0) mbar_expect_tx(&full[s], feed_a ? 4 * kABytes : 4 * kBBytes);
          if (MODE == TC_WGRAD) {
            const int im = (kb / patches_per_img) * g.pn;            // first image of the patch's image group
            const int prem = kb % patches_per_img;
            const int py = (prem / g.patches_w) * g.ph, px = (prem % g.patches_w) * g.pw;
            const int kh = t.tap_w / g.KW, kw = t.tap_w - kh * g.KW;
            if (feed_a) {                                                      // A: dy, 128 output channels = 4 atoms in one box
              if (PAIR) {
                tma_load_5d_pair(a_hi, &map_a_hi, fbar, 0, px, py, im, t.m0 / kElems);
                tma_load_5d_pair(a_lo, &map_a_lo, fbar, 0, px, py, im, t.m0 / kElems);
              } else {
                tma_load_5d(a_hi, &map_a_hi, &full[s], 0, px, py, im, t.m0 / kElems);
                tma_load_5d(a_lo, &map_a_lo, &full[s], 0, px, py, im, t.m0 / kElems);
              }
            } else {                                                           // B: x shifted by the tap, BN / 32 atoms in one box
              if (PAIR) {
                tma_load_5d_pair(b_hi, &map_b_hi, fbar, 0, px + kw - g.pad, py + kh - g.pad, im, nb0 / kElems);
                tma_load_5d_pair(b_lo, &map_b_lo, fbar, 0, px + kw - g.pad, py + kh - g.pad, im, nb0 / kElems);
              } else {
                tma_load_5d(b_hi, &map_b_hi, &full[s], 0, px + kw - g.pad, py + kh - g.pad, im, t.n0 / kElems);
                tma_load_5d(b_lo, &map_b_lo, &full[s], 0, px + kw - g.pad, py + kh - g.pad, im, t.n0 / kElems);
              }
            }
          } else {
            const int tap = kb / kblocks_c, c0 = (kb - tap * kblocks_c) * kElems;
            const int kh = tap / g.KW, kw = tap - kh * g.KW;
            const int dx = (MODE == TC_FWD) ? (kw - g.pad) : (g.pad - kw);
            const int dy = (MODE == TC_FWD) ? (kh - g.pad) : (g.pad - kh);
            if (feed_a) {
              if (PAIR) {
                tma_load_4d_pair(a_hi, &map_a_hi, fbar, c0, t.ow0 + dx, t.oh0 + dy, t.img);
                tma_load_4d_pair(a_lo, &map_a_lo, fbar, c0, t.ow0 + dx, t.oh0 + dy, t.img);
              } else {
                tma_load_4d(a_hi, &map_a_hi, &full[s], c0, t.ow0 + dx, t.oh0 + dy, t.img);
                tma_load_4d(a_lo, &map_a_lo, &full[s], c0, t.ow0 + dx, t.oh0 + dy, t.img);
              }
            } else if (MODE == TC_FWD) {
              if (PAIR) {
                tma_load_2d_pair(b_hi, &map_b_hi, fbar, tap * g.Cin + c0, nb0);
                tma_load_2d_pair(b_lo, &map_b_lo, fbar, tap * g.Cin + c0, nb0);
              } else {
                tma_load_2d(b_hi, &map_b_hi, &full[s], tap * g.Cin + c0, t.n0);
                tma_load_2d(b_lo, &map_b_lo, &full[s], tap * g.Cin + c0, t.n0);
              }
            } else {                                                           // B: w[co0..+32][tap][n0 .. n0+BN) as BN / 32 atoms in one box
              if (PAIR) {
                tma_load_4d_pair(b_hi, &map_b_hi, fbar, 0, tap, c0, nb0 / kElems);
                tma_load_4d_pair(b_lo, &map_b_lo, fbar, 0, tap, c0, nb0 / kElems);
              } else {
                tma_load_4d(b_hi, &map_b_hi, &full[s], 0, tap, c0, t.n0 / kElems);
                tma_load_4d(b_lo, &map_b_lo, &full[s], 0, tap, c0, t.n0 / kElems);
              }
            }
          }
          if (it == 0 && feed_a) TC_TRACE(2);
        }
      }
      if (PAIR && feed_a) {
        // producer tail: the leader's commits arrive on THIS CTA's `empty` barriers asynchronously; wait for the last one of every slot
        // so that no arrival can land in the shared memory of a CTA that has already exited
        for (int i = it > STAGES ? it - STAGES : 0; i < it; i++) mbar_wait(&empty[i % STAGES], (i / STAGES) & 1);
      }
    }
  } else if (warp == kWarpMma) {
    if (rank == 0) {
      // ===== MMA issuer (pair: the leader CTA's, for both).  The whole warp walks the loop (waits included); one elected lane issues. =====
      constexpr uint32_t idesc_main = F16 ? make_idesc_f16(128, 2 * BN, kAMajorMN ? 1 : 0, kBMajorMN ? 1 : 0)
                                          : make_idesc_tf32(128, 2 * BN, kAMajorMN ? 1 : 0, kBMajorMN ? 1 : 0);   // A_hi x [B_hi | B_lo]
      constexpr uint32_t idesc_corr = F16 ? make_idesc_f16(128, BN, kAMajorMN ? 1 : 0, kBMajorMN ? 1 : 0)
                                          : make_idesc_tf32(128, BN, kAMajorMN ? 1 : 0, kBMajorMN ? 1 : 0);       // A_lo x B_hi
      const uint32_t a_kstep = kAMajorMN ? g.mn_kstep : 32, a_lbo = kAMajorMN ? g.mn_lbo : 16, a_sbo = kAMajorMN ? g.mn_sbo : 1024;
      const uint32_t b_kstep = kBMajorMN ? g.mn_kstep : 32, b_lbo = kBMajorMN ? g.mn_lbo : 16, b_sbo = kBMajorMN ? g.mn_sbo : 1024;
      constexpr uint32_t mn_lt = F16 ? kLayoutSW128 : kLayoutSW128Base32B;
      const uint64_t a_desc = make_smem_desc_base(a_lbo, a_sbo, kAMajorMN ? mn_lt : kLayoutSW128);
      const uint64_t b_desc = make_smem_desc_base(b_lbo, b_sbo, kBMajorMN ? mn_lt : kLayoutSW128);
      uint32_t accumulate = 0;
      int it = 0, chunk = 0;                                         // ring position / accumulation chains started, across all items
      uint32_t tmem_acc = tmem_base;
      constexpr uint32_t idesc_pair = make_idesc_f16(256, BN, kAMajorMN ? 1 : 0, kBMajorMN ? 1 : 0);                // pair: M = 256, N = BN
      TcCursor cur = tc_cursor(g, gid, G);
      TcItem t;
      bool first_item = true;
      while (tc_next<MODE, BN, PAIR>(g, cur, t, rank, G)) {
        for (int i = 0; i < t.nkb; i++, it++) {
          if (i % kChunk == 0) {                                     // new accumulation chain in the other TMEM buffer
            const int b = chunk & 1;
            mbar_wait(&acc_empty[b], ((chunk >> 1) & 1) ^ 1);        // drained by the epilogue warps (first use passes)
            tc_fence_after();
            tmem_acc = tmem_base + b * 2 * BN;
            accumulate = 0;
          }
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (it == 0 && lane == 0) TC_TRACE(3);
          const uint32_t a_hi = smem_u32(smem + s * kStageBytes);
          const uint32_t a_lo = a_hi + kABytes;
          const uint32_t b_hi = a_hi + 2 * kABytes;                  // b_lo follows at + kBBytes
          const bool chain_end = (i % kChunk == kChunk - 1 || i == t.nkb - 1);
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; k++) {                              // four MMAs per k-block: K = 8 (tf32) / 16 (fp16) = 32 bytes of a K-major row each
            const uint64_t da_hi = smem_desc_at(a_desc, a_hi + k * a_kstep);
            const uint64_t da_lo = smem_desc_at(a_desc, a_lo + k * a_kstep);
            const uint64_t db = smem_desc_at(b_desc, b_hi + k * b_kstep);                 // covers b_hi then b_lo (contiguous)
            if (PAIR) {
              const uint64_t db_lo = smem_desc_at(b_desc, b_hi + kBBytes + k * b_kstep);
              umma_f16_pair(tmem_acc, da_hi, db, idesc_pair, k ? 1u : accumulate);         // main (+)= A_hi * B_hi     (each CTA: its rows x all BN columns)
              umma_f16_pair(tmem_acc + BN, da_hi, db_lo, idesc_pair, k ? 1u : accumulate); // corr (+)= A_hi * B_lo
              umma_f16_pair(tmem_acc + BN, da_lo, db, idesc_pair, 1);             // corr  += A_lo * B_hi
            } else if (F16) {
              umma_f16(tmem_acc, da_hi, db, idesc_main, k ? 1u : accumulate);
              umma_f16(tmem_acc + BN, da_lo, db, idesc_corr, 1);
            } else {
              umma_tf32(tmem_acc, da_hi, db, idesc_main, k ? 1u : accumulate); // [main | corr] (+)= A_hi * [B_hi | B_lo]
              umma_tf32(tmem_acc + BN, da_lo, db, idesc_corr, 1);     // corr += A_lo * B_hi
            }
          }
          if (PAIR) umma_commit_pair(&empty[s]); else umma_commit(&empty[s]);   // frees the operand slot (pair: in both CTAs) when these MMAs retire
          if (chain_end) {
            if (PAIR) umma_commit_pair(&acc_full[chunk & 1]); else umma_commit(&acc_full[chunk & 1]);   // chain complete -> epilogue warps may drain it
          }
          }
          __syncwarp();
          accumulate = 1;
          if (chain_end) chunk++;
        }
        if (first_item && lane == 0) TC_TRACE(4);
        first_item = false;
      }
      if (lane == 0) TC_TRACE(5);
    }
  }
  } else {
    // ===== accumulate + epilogue: warps 0..3 own TMEM lane quarters (warp % 4) =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTcEpilogueRegs));
    // While these warps finish and store item i, the MMA warp is already up to two chains into item i+1.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool raw = g.splits > 1;
    // fp16 engine: main columns hold sum(hi_a hi_b) of operands scaled by 2^e_a, 2^e_b, corr columns additionally by 2^11; both
    // factors are powers of two, so the FMAs below are exact rescales followed by the same round-to-nearest add as the tf32 path
    float s_main = 1.0f, s_corr = 1.0f;
    if (F16) {
      const int e = __ldg(g.a_exp) + __ldg(g.b_exp);
      s_main = pow2i(-e);
      s_corr = pow2i(-e - kF16LoShift);
    }
    int cg = 0;                                                      // accumulation chains drained so far, across all items
    float out_max = 0.f;                                             // max |final output| written by this thread (amax_out)
    TcCursor cur = tc_cursor(g, gid, G);
    TcItem t;
    bool first_item = true;
    while (tc_next<MODE, BN, PAIR>(g, cur, t, rank, G)) {
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; j++) acc[j] = 0.f;
      const int nchunks = (t.nkb + kChunk - 1) / kChunk;
      for (int c = 0; c < nchunks; c++, cg++) {
        const int b = cg & 1;
        mbar_wait(&acc_full[b], (cg >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < BN / 32; cc++) {
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + b * 2 * BN + BN + cc * 32, v);    // corr: A_hi*B_lo + A_lo*B_hi
#pragma unroll
          for (int j = 0; j < 32; j++) acc[cc * 32 + j] = F16 ? fmaf(v[j], s_corr, acc[cc * 32 + j]) : acc[cc * 32 + j] + v[j];   // fp32 round-to-nearest, outside the tensor core
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + b * 2 * BN + cc * 32, v);         // main: A_hi*B_hi
#pragma unroll
          for (int j = 0; j < 32; j++) acc[cc * 32 + j] = F16 ? fmaf(v[j], s_main, acc[cc * 32 + j]) : acc[cc * 32 + j] + v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(&acc_empty[b]), 0));     // the leader's barrier counts the warps of both CTAs
          else mbar_arrive(&acc_empty[b]);
        }
      }
      if (first_item && threadIdx.x == 0) TC_TRACE(6);
      if (t.kind == TC_ITEM_PART) {
        // park the raw partial sums of this (tile, k-range) and publish them; the tile's HEAD owner folds them in
        float *slot = g.sk_slots + ((size_t)blockIdx.x * 128 + row) * BN;
#pragma unroll
        for (int j = 0; j < BN; j += 4) __stcg(reinterpret_cast<float4 *>(slot + j), make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps
        if (threadIdx.x == 0) {
          asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(g.sk_flags + blockIdx.x), "l"(g.sk_tag) : "memory");
          if (first_item) TC_TRACE(7);
        }
        first_item = false;
        continue;
      }
      if (t.kind == TC_ITEM_HEAD) {
        // the CTAs after this one hold the rest of the tile's K range as the FIRST item of their ranges: long done, or about to be
        for (int j = gid + 1; j < G && tc_sk_start(g, j, G) < t.tile_end; j++) {
          const int pj = PAIR ? 2 * j + rank : j;                    // the CTA of work group j that holds the same 128 rows
          if (threadIdx.x == 0) {
            const long long t0 = clock64();
            unsigned long long seen;
            do {
              asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(g.sk_flags + pj) : "memory");
              if (seen != g.sk_tag && clock64() - t0 > 20000000000ll) __trap();    // ~10 s: a partner that never ran
            } while (seen != g.sk_tag);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const float *slot = g.sk_slots + ((size_t)pj * 128 + row) * BN;
#pragma unroll
          for (int q = 0; q < BN; q += 4) {
            const float4 p = __ldcg(reinterpret_cast<const float4 *>(slot + q));
            acc[q] += p.x; acc[q + 1] += p.y; acc[q + 2] += p.z; acc[q + 3] += p.w;
          }
        }
      }
      bool valid;
      size_t row_off;           // element offset of this row's first column (col = n0)
      size_t res_off = 0;
      int ntot;
      if (MODE == TC_WGRAD) {
        ntot = KHW * g.Cin;
        valid = (t.m0 + row) < g.Cout;
        row_off = ((size_t)(t.m0 + row) * KHW + t.tap_w) * g.Cin + t.n0;
        if (raw) row_off += (size_t)t.split * g.Cout * ntot;
      } else {
        ntot = (MODE == TC_FWD) ? g.Cout : g.Cin;
        const int per_img = g.tile_w * g.tile_h;
        const int nn = row / per_img, rrem = row - nn * per_img;
        const int oh = t.oh0 + rrem / g.tile_w, ow = t.ow0 + rrem % g.tile_w;
        valid = (t.img + nn) < g.nimg && oh < g.H && ow < g.W;
        const size_t pix = ((size_t)(t.img + nn) * g.H + oh) * g.W + ow;
        row_off = pix * ntot + t.n0;
        res_off = row_off;
        if (raw) row_off += (size_t)t.split * ((size_t)g.nimg * g.H * g.W) * ntot;
      }
      if (valid) {
        float *dst = (raw ? partial : out) + row_off;
        if (!raw && MODE != TC_WGRAD) {
          // scale / bias / residual as 128-bit loads with no control flow between them (the loads of a whole row batch up), the
          // activation dispatched ONCE outside the element loop: a per-element `switch (act)` serialises every load behind a branch
          const bool has_scale = epi.scale != nullptr, has_bias = epi.bias != nullptr, has_res = epi.residual != nullptr;
#pragma unroll
          for (int j = 0; j < BN; j += 4) {
            if (has_scale) {
              const float4 sc = __ldg(reinterpret_cast<const float4 *>(epi.scale + t.n0 + j));
              acc[j] *= sc.x; acc[j + 1] *= sc.y; acc[j + 2] *= sc.z; acc[j + 3] *= sc.w;
            }
            if (has_bias) {
              const float4 bi = __ldg(reinterpret_cast<const float4 *>(epi.bias + t.n0 + j));
              acc[j] += bi.x; acc[j + 1] += bi.y; acc[j + 2] += bi.z; acc[j + 3] += bi.w;
            }
            if (has_res) {
              const float4 re = __ldg(reinterpret_cast<const float4 *>(epi.residual + res_off + j));
              acc[j] += re.x; acc[j + 1] += re.y; acc[j + 2] += re.z; acc[j + 3] += re.w;
            }
          }
          if (epi.act == FRCNN_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < BN; j++) acc[j] = acc[j] > 0.0f ? acc[j] : 0.0f;
          } else if (epi.act == FRCNN_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < BN; j++) acc[j] = 1.0f / (1.0f + expf(-acc[j]));
          }
          if (g.amax_out) {
#pragma unroll
            for (int j = 0; j < BN; j++) out_max = fmaxf(out_max, fabsf(acc[j]));
          }
        }
        if (!g.paired_stores) {
#pragma unroll
          for (int j = 0; j < BN; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
      }
      if (g.paired_stores) {
        // Full-sector stores.  A thread owns one output row, so a plain 128-bit store instruction touches 32 rows x 16 bytes: 32 half
        // sectors.  Lane pairs swap every other 4-column group (one shuffle per element) and then write 2 x 16 adjacent bytes of ONE row
        // per instruction: every 32-byte sector is written whole, by one instruction.  (Largest where the epilogue is not hidden -- fc1's
        // filter gradient, 411 MB of output for 128 k-elements per tile: 0.188 -> 0.120 ms -- and never slower: profiles/r02_pair_ab.md.)
        const bool odd = lane & 1;
        float *base = raw ? partial : out;
        const unsigned long long own_off = valid ? (unsigned long long)row_off : ~0ull;
        const unsigned long long peer_off = __shfl_xor_sync(0xffffffffu, own_off, 1);
#pragma unroll
        for (int j = 0; j < BN; j += 8) {                              // groups (j .. j+3) and (j+4 .. j+7)
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float send = odd ? acc[j + e] : acc[j + 4 + e];      // the even lane gives away its odd group, the odd lane its even group
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            if (odd) acc[j + e] = recv; else acc[j + 4 + e] = recv;
          }
          // slot j: even lane = own row, group j | odd lane = the even lane's row, group j + 4  -> 32 adjacent bytes of the even lane's row
          const unsigned long long off0 = odd ? peer_off : own_off;
          if (off0 != ~0ull) *reinterpret_cast<float4 *>(base + off0 + j + (odd ? 4 : 0)) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          // slot j + 4: even lane = the odd lane's row, group j | odd lane = own row, group j + 4   -> 32 adjacent bytes of the odd lane's row
          const unsigned long long off1 = odd ? own_off : peer_off;
          if (off1 != ~0ull) *reinterpret_cast<float4 *>(base + off1 + j + (odd ? 4 : 0)) = make_float4(acc[j + 4], acc[j + 5], acc[j + 6], acc[j + 7]);
        }
      }
      if (first_item && threadIdx.x == 0) TC_TRACE(7);
      first_item = false;
    }
    if (g.amax_out) {
      // the operand split of this output needs its absolute maximum: one partial per CTA, so the consumer needs no amax pass
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) out_max = fmaxf(out_max, __shfl_xor_sync(0xffffffffu, out_max, o));
      float *wmax = reinterpret_cast<float *>(tmem_slot + 1);        // 4 words behind the TMEM slot (inside the 256-byte tail)
      if (lane == 0) wmax[q] = out_max;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) g.amax_out[kF16PartialsAt + blockIdx.x] = __float_as_uint(fmaxf(fmaxf(wmax[0], wmax[1]), fmaxf(wmax[2], wmax[3])));
      // a launch on fewer CTAs than the consumer expects partials from (frcnn_set_sm_reserve): the missing ones are zero
      if (blockIdx.x == 0)
        for (int sl = (int)gridDim.x + threadIdx.x; sl < g.amax_slots; sl += 128) g.amax_out[kF16PartialsAt + sl] = 0u;
    }
    if (threadIdx.x == 0) TC_TRACE(8);
  }
  if (late_trigger) pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                                      // both CTAs are done with each other's barriers and tensor memory
  if (warp == kWarpMma) {
    __syncwarp();
    if (PAIR) tmem_dealloc_pair(tmem_base, 4 * BN); else tmem_dealloc(tmem_base, 4 * BN);
  }
  // tcgen05.dealloc.cta_group::2 is issued by one warp of EACH CTA for the columns of BOTH: neither CTA may exit (and hand its SM -- with
  // programmatic dependent launch, instantly -- to a CTA of the next kernel that allocates the same columns) before the peer has issued its own
  if (PAIR) cluster_sync_all();
  if (threadIdx.x == 0 && g.trace) {
    TC_TRACE(9);
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g.trace[(size_t)blockIdx.x * 16 + 10] = smid;
  }
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// mn_major: the tile feeds an MN-major (transposed) operand; 32-bit elements then need the 32-byte-atom swizzle, fp16 the plain one.
// Dimensions are in elements, strides in bytes; elem = 4 (tf32 engine) or 2 (fp16 engine).
static bool encode(CUtensorMap *m, const void *base, int elem, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box, bool mn_major)
{
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn f = encode_fn();
  if (!f) return false;
  const bool f16 = elem == 2;
  return f(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           (mn_major && !f16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// activation (N,H,W,C) as a 4-D tensor (C, W, H, N); box {one 128-byte row of channels, box_w, box_h, box_n}
static bool make_act_map(CUtensorMap *m, const void *base, int elem, int N, int H, int W, int C, int box_w, int box_h, int box_n, bool mn_major = false)
{
  const cuuint64_t e = (cuuint64_t)elem;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * e, (cuuint64_t)W * C * e, (cuuint64_t)H * W * C * e};
  cuuint32_t box[4] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  return encode(m, base, elem, 4, dims, strides, box, mn_major);
}

// matrix (rows, K) row-major as a 2-D tensor (K, rows); box {one 128-byte row, box_rows}
static bool make_mat_map(CUtensorMap *m, const void *base, int elem, int rows, int K, int box_rows)
{
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * elem};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows};
  return encode(m, base, elem, 2, dims, strides, box, false);
}

// MN-major operand tiles are stacks of atoms of A = 128 / elem channels ([atom][k-rows][A channels]: 32 x 32 fp32 = 4 KB, 64 x 64 fp16 =
// 8 KB).  Splitting the channel axis into (A, C/A) and putting the atom index LAST in the tensor map lets ONE box load deliver the whole
// stack in exactly that order (the producer thread's issue rate of small boxes was the limit of the filter-gradient mainloop):
// activation (N,H,W,C) as the 5-D tensor (A, W, H, N, C/A); box {A, box_w, box_h, box_n, atoms}
static bool make_act_map_atoms(CUtensorMap *m, const void *base, int elem, int N, int H, int W, int C, int box_w, int box_h, int box_n, int atoms)
{
  const cuuint64_t e = (cuuint64_t)elem, A = 128 / elem;
  cuuint64_t dims[5] = {A, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)C / A};
  cuuint64_t strides[4] = {(cuuint64_t)C * e, (cuuint64_t)W * C * e, (cuuint64_t)H * W * C * e, 128};
  cuuint32_t box[5] = {(cuuint32_t)A, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n, (cuuint32_t)atoms};
  return encode(m, base, elem, 5, dims, strides, box, true);
}

// filter (Cout, taps, Cin) as the 4-D tensor (A, taps, Cout, Cin/A); box {A ci, 1 tap, A co (the k-block), atoms}
static bool make_filter_map_atoms(CUtensorMap *m, const void *base, int elem, int Cout, int taps, int Cin, int atoms)
{
  const cuuint64_t e = (cuuint64_t)elem, A = 128 / elem;
  cuuint64_t dims[4] = {A, (cuuint64_t)taps, (cuuint64_t)Cout, (cuuint64_t)Cin / A};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * e, (cuuint64_t)taps * Cin * e, 128};
  cuuint32_t box[4] = {(cuuint32_t)A, 1, (cuuint32_t)A, (cuuint32_t)atoms};
  return encode(m, base, elem, 4, dims, strides, box, true);
}

struct TcPlan {
  int N, H, W;                       // after folding nn.Linear (H=W=1) into a 1 x rows "image"
  int BN, stages;
  int tile_w, tile_h, tile_n, tiles_w, tiles_h, groups;
  int pw, ph, pn, patches_w, patches_h, pgroups;
  int total_kb, splits, kb_per_split;
  int m_tiles, n_tiles, items;       // persistent-loop work items (see TcGeom)
  int pair, m_pairs;                 // CTA-pair kernel (fp16 engine): work items are pairs of 128-row tiles, the grid is made of 2-CTA clusters
  int streamk, grid;                 // stream-K decomposition (default) and its CTA count
  int grid_max;                      // the CTA count with no SMs reserved: sizes the workspace layout and the amax slots, whatever `grid` is
  long long units;
  size_t flags_off;
  size_t a_hi_off, a_lo_off, b_hi_off, b_lo_off, partial_off, total_bytes;
  size_t a_count, b_count;           // element counts of the two operands that need a lo part
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// patch (pw x ph x pn, product == total, powers of two) with the least padded volume; ties -> wider, then taller
static void best_patch(int total, int N, int H, int W, int *pw, int *ph, int *pn)
{
  long long best = -1;
  for (int w = total; w >= 1; w >>= 1)
    for (int h = total / w; h >= 1; h >>= 1) {
      int n = total / (w * h);
      if (w > 256 || h > 256 || n > 256) continue;
      long long vol = (long long)ceil_div(W, w) * w * ceil_div(H, h) * h * ceil_div(N, n) * n;
      if (best < 0 || vol < best) { best = vol; *pw = w; *ph = h; *pn = n; }
    }
}

static bool make_tc_plan(int mode, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, TcPlan *p, bool f16 = false)
{
  const int kel = f16 ? 64 : 32;                                          // elements per k-block row / channels per MN-major atom
  const int chunk = f16 ? kChunkKB16 : kChunkKB;
  const size_t esz = f16 ? 2 : 4;
  if (stride != 1 || KH != KW || 2 * pad != KH - 1) return false;        // stride-1 "same" convs, 1x1, linear
  if (KH == 1 && H == 1 && W == 1) { p->N = 1; p->H = 1; p->W = N; }      // linear: rows become the W axis
  else { p->N = N; p->H = H; p->W = W; }
  const long long pixels = (long long)p->N * p->H * p->W;
  if (pixels < 64) return false;                                         // tiny problems stay on the CUDA-core engine
  const int ntot = (mode == TC_FWD) ? Cout : Cin;                         // GEMM N extent
  if (ntot % 64 != 0) return false;
  if (mode == TC_FWD && Cin % kel != 0) return false;
  if (mode == TC_DGRAD && Cout % kel != 0) return false;
  if (mode == TC_WGRAD && (Cout % 128 != 0 || Cin % 64 != 0)) return false;
  p->BN = (ntot % 128 == 0) ? 128 : 64;
  p->stages = p->BN == 128 ? 3 : 4;
  // CTA pairs (cta_group::2): fp16 engine; the forward pass takes both tile widths (its B operand is K-major: any row count splits in
  // two), dgrad / wgrad need BN = 128 (an MN-major half must be a whole 64-column swizzle atom).  Default on (FRCNN_TC_PAIR=0 = single CTAs):
  // +10-15 % per layer on the 75x125 .. 300x500 convolutions once the MMA issue loop ran on uniform registers (profiles/r02_pair_ab.md).
  static const bool use_pair = !(getenv("FRCNN_TC_PAIR") && atoi(getenv("FRCNN_TC_PAIR")) == 0);
  p->pair = (f16 && use_pair && (mode == TC_FWD || p->BN == 128)) ? 1 : 0;                  // (+ an even / large tile count, below)
  if (p->pair) p->stages = p->BN == 128 ? 4 : 5;                          // 48 KB / 40 KB per stage
  p->tile_w = p->tile_h = p->tile_n = p->tiles_w = p->tiles_h = p->groups = 1;
  p->pw = p->ph = p->pn = p->patches_w = p->patches_h = p->pgroups = 1;
  const int taps = KH * KW;
  int ctas;
  if (mode == TC_WGRAD) {
    best_patch(kel, p->N, p->H, p->W, &p->pw, &p->ph, &p->pn);
    p->patches_w = ceil_div(p->W, p->pw);
    p->patches_h = ceil_div(p->H, p->ph);
    p->pgroups = ceil_div(p->N, p->pn);
    p->total_kb = p->pgroups * p->patches_w * p->patches_h;
    p->m_tiles = Cout / 128;
    p->n_tiles = Cin / p->BN;
    p->m_pairs = ceil_div(p->m_tiles, 2);
    ctas = (p->pair ? p->m_pairs : p->m_tiles) * p->n_tiles * taps;
  } else {
    best_patch(128, p->N, p->H, p->W, &p->tile_w, &p->tile_h, &p->tile_n);
    p->tiles_w = ceil_div(p->W, p->tile_w);
    p->tiles_h = ceil_div(p->H, p->tile_h);
    p->groups = ceil_div(p->N, p->tile_n);
    p->total_kb = taps * ((mode == TC_FWD ? Cin : Cout) / kel);
    p->m_tiles = p->groups * p->tiles_w * p->tiles_h;
    p->n_tiles = ntot / p->BN;
    p->m_pairs = ceil_div(p->m_tiles, 2);
    ctas = (p->pair ? p->m_pairs : p->m_tiles) * p->n_tiles;
  }
  if (p->pair && (p->m_tiles & 1) && p->m_tiles < 16) {
    // an odd tile count leaves one CTA of the last pair idle: nn.Linear on 128 RoIs (ONE row tile) would do twice the tensor work
    p->pair = 0;
    p->stages = p->BN == 128 ? 3 : 4;
    ctas = (mode == TC_WGRAD) ? p->m_tiles * p->n_tiles * taps : p->m_tiles * p->n_tiles;
  }
  // from here on `ctas` counts WORK GROUPS' tiles: 128-row tiles for single CTAs, 256-row tile pairs for CTA pairs; `wg_max` is the number
  // of work groups one wave holds (SMs, or 2-SM clusters)
  const int wg_max = p->pair ? kNumSMs / 2 : kNumSMs;
  // Work decomposition.  Default = stream-K: the tiles * total_kb k-block units of the launch are cut into one equal contiguous
  // range per CTA (persistent grid <= one CTA per SM), so the SMs finish together whatever the tile count; a tile that straddles
  // two ranges is summed through a per-CTA slot in the workspace by the CTA owning its first k-block (no reduce kernel).
  // FRCNN_TC_STREAMK=0 selects the older scheme: whole-tile items, optional split-K chosen by a wave-quantisation cost model,
  // partial sums reduced by a second kernel.
  static const bool use_streamk = !(getenv("FRCNN_TC_STREAMK") && atoi(getenv("FRCNN_TC_STREAMK")) == 0);
  const size_t out_elems_plan = (mode == TC_WGRAD) ? (size_t)Cout * taps * Cin : (size_t)pixels * ntot;
  int splits = 1;
  // stream-K pays when a launch has between ~half a wave and a few waves of tiles: fewer tiles mean many CTAs share one tile and
  // the owner's serial fix-up (64 KB per partner) outweighs the balance -- the split-K + parallel reduce path is better there --
  // and with many waves of tiles the quantisation loss is below 1/8 of a tile per SM anyway.
  p->streamk = (use_streamk && ctas >= wg_max / 2 && ctas < 8 * wg_max) ? 1 : 0;
  p->units = (long long)ctas * p->total_kb;
  // SMs this launch may occupy: all of them, minus the ones set aside for a concurrent collective (frcnn_set_sm_reserve).  Only the
  // CTA count follows it; the decomposition (stream-K or split-K, number of splits) and the workspace layout never change.
  const int sms = p->pair ? sm_budget() / 2 : sm_budget();           // work groups this launch may occupy
  const int per_wg = p->pair ? 2 : 1;                                // CTAs per work group
  if (p->streamk) {
    long long gsz = p->units / chunk;                                // at least one accumulation chain per work group
    if (gsz < 1) gsz = 1;
    p->grid_max = per_wg * (int)(gsz > wg_max ? wg_max : gsz);
    p->grid = per_wg * (int)(gsz > sms ? sms : gsz);
  } else {
    const double t_kb = p->pair ? ((p->BN == 128) ? 0.48 : 0.36) : ((p->BN == 128) ? 0.56 : 0.42);   // us per k-block (smem-bandwidth bound mainloop)
    const double t_item = 2.5;                                       // us of per-item pipeline refill / final drain not overlapped
    double best_t = -1.0;
    const int max_splits = p->total_kb / chunk > 32 ? 32 : p->total_kb / chunk;
    for (int sp = 1; sp <= (max_splits < 1 ? 1 : max_splits); sp++) {
      const int kbs = ceil_div(p->total_kb, sp);
      if (ceil_div(p->total_kb, kbs) != sp) continue;                // slice lengths that do not produce exactly sp splits
      const long long rounds = ((long long)ctas * sp + wg_max - 1) / wg_max;
      double tt = rounds * (kbs * t_kb + t_item);
      if (sp > 1) tt += 4.0 + (double)(sp + 2) * out_elems_plan * 4.0 / 4.0e6;   // reduce pass at ~4 TB/s effective + launch
      if (best_t < 0 || tt < best_t) { best_t = tt; splits = sp; }
    }
  }
  p->kb_per_split = ceil_div(p->total_kb, splits);
  p->splits = ceil_div(p->total_kb, p->kb_per_split);
  p->items = ctas * p->splits;
  if (!p->streamk) {
    p->grid_max = per_wg * (p->items > wg_max ? wg_max : p->items);
    p->grid = per_wg * (p->items > sms ? sms : p->items);
  }
  const size_t act_in = (size_t)pixels * Cin, act_out = (size_t)pixels * Cout, filt = (size_t)Cout * taps * Cin;
  size_t out_elems;
  if (mode == TC_FWD) { p->a_count = act_in; p->b_count = filt; out_elems = act_out; }
  else if (mode == TC_DGRAD) { p->a_count = act_out; p->b_count = filt; out_elems = act_in; }
  else { p->a_count = act_out; p->b_count = act_in; out_elems = filt; }
  // internal operand splits (used when the caller passes none): tf32 = [hi | lo]; fp16 = [header | hi | lo] per operand
  const size_t hdr = f16 ? kF16Header : 0;
  p->a_hi_off = hdr;
  p->a_lo_off = p->a_hi_off + align_up(p->a_count * esz, 1024);
  p->b_hi_off = p->a_lo_off + align_up(p->a_count * esz, 1024) + hdr;
  p->b_lo_off = p->b_hi_off + align_up(p->b_count * esz, 1024);
  p->partial_off = p->b_lo_off + align_up(p->b_count * esz, 1024);
  if (p->streamk) {
    p->flags_off = p->partial_off + align_up((size_t)p->grid_max * 128 * p->BN * 4, 1024);
    p->total_bytes = p->flags_off + align_up((size_t)p->grid_max * 8, 1024);
  } else {
    p->flags_off = 0;
    p->total_bytes = p->partial_off + (p->splits > 1 ? (size_t)p->splits * out_elems * 4 : 0);
  }
  return true;
}

template <int MODE, int BN, int STAGES, bool F16, bool PAIR = false>
static int launch_tc(const CUtensorMap *maps, const TcGeom &g, dim3 grid, float *out, float *partial, const Epilogue &epi, cudaStream_t st)
{
  constexpr int smem = STAGES * (2 * kABytes + 2 * (PAIR ? BN / 2 : BN) * kBK * 4) + 1024 + 256;
  static const cudaError_t attr = cudaFuncSetAttribute(tc_conv_kernel<MODE, BN, STAGES, F16, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (attr != cudaSuccess) return cuda_fail(attr, "tc_conv_kernel: smem attribute");
  // The pair kernels themselves are launched WITHOUT the programmatic-serialization attribute (they start only after the previous kernel
  // has completed; the kernel behind them still overlaps their tail): a 2-CTA cluster grid launched as a programmatic dependent hung or
  // trapped within ~30 train steps in round 2 (profiles/r02_pair_ab.md), every other combination ran clean.  FRCNN_TC_PAIR_PDL=1 restores it.
  static const bool pair_pdl = getenv("FRCNN_TC_PAIR_PDL") && atoi(getenv("FRCNN_TC_PAIR_PDL")) != 0;
  if (PAIR) launch_cluster(tc_conv_kernel<MODE, BN, STAGES, F16, PAIR>, grid, kTcThreads, smem, st, 2, pair_pdl, maps[0], maps[1], maps[2], maps[3], g, out, partial, epi);
  else launch(tc_conv_kernel<MODE, BN, STAGES, F16, PAIR>, grid, kTcThreads, smem, st, maps[0], maps[1], maps[2], maps[3], g, out, partial, epi);
  FRCNN_CHECK_LAUNCH("tc_conv_kernel");
  return FRCNN_OK;
}

// how many 2-CTA clusters of the pair kernel (forward, BN = 128) the device can hold at once: the persistent stream-K grid must not exceed
// it -- a cluster that is not resident cannot publish the partial sums its neighbour spins on
int tc_pair_max_active_clusters()
{
  auto kernel = tc_conv_kernel<TC_FWD, 128, 4, true, true>;
  constexpr int smem = 4 * (2 * kABytes + 2 * 64 * kBK * 4) + 1024 + 256;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kNumSMs, 1, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  return n;
}

static unsigned long long *g_tc_trace = nullptr;      // debug only (frcnn_debug_tc_trace); never set on the product path
void tc_set_trace(void *buf) { g_tc_trace = reinterpret_cast<unsigned long long *>(buf); }

// a: the activation-side operand of the mode (x | dy | dy), b: the other one (w | w | x)
size_t tf32_split_bytes(size_t count) { return 2 * align_up(count * 4, 1024); }

int tf32_split(const float *x, size_t count, void *out, cudaStream_t st)
{
  float *hi = reinterpret_cast<float *>(out);
  float *lo = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(out) + align_up(count * 4, 1024));
  launch(split_hi_lo_kernel, elementwise_grid(count / 4 + 1, 256), 256, 0, st, x, hi, lo, count);
  FRCNN_CHECK_LAUNCH("split_hi_lo_kernel");
  return FRCNN_OK;
}

size_t f16_split_bytes(size_t count) { return kF16Header + 2 * f16_half_bytes(count); }

// partials != NULL: G partial maxima (bit patterns, laid out like a header: first one at word kF16PartialsAt) already produced by the
// kernel that wrote x (the GEMM epilogue) -- or valid for a superset of x (an un-pooled map) -- so the amax pass is skipped
int f16_split(const float *x, size_t count, void *out, cudaStream_t st, const void *partials, int G)
{
  uint8_t *o = reinterpret_cast<uint8_t *>(out);
  if (partials == nullptr) {
    G = f16_launch_amax(x, count, o, st);
    FRCNN_CHECK_LAUNCH("f16_amax_partials_kernel");
    partials = o;
  }
  __half *hi = reinterpret_cast<__half *>(o + kF16Header);
  __half *lo = reinterpret_cast<__half *>(o + kF16Header + f16_half_bytes(count));
  launch(split_f16_kernel, elementwise_grid(count / 4 + 1, 256), 256, 0, st, x, reinterpret_cast<unsigned *>(o), reinterpret_cast<const unsigned *>(partials), G, hi, lo, count);
  FRCNN_CHECK_LAUNCH("split_f16_kernel");
  return FRCNN_OK;
}

// a_split / b_split: optional buffers produced by tf32_split / f16_split for the two operands (NULL = split here)
static int run_tc(int mode, const float *a, const float *b, float *out, const Epilogue &epi,
                  int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                  void *workspace, size_t workspace_bytes, cudaStream_t st, const void *a_split = nullptr, const void *b_split = nullptr, bool f16 = false,
                  void *amax_out = nullptr)
{
  TcPlan p;
  if (!make_tc_plan(mode, N, H, W, Cin, Cout, KH, KW, stride, pad, &p, f16)) return fail(FRCNN_E_UNSUPPORTED, "tcgen05 engine: unsupported shape");
  if (workspace == nullptr || workspace_bytes < p.total_bytes) return fail(FRCNN_E_WORKSPACE, "tcgen05 engine: workspace too small");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return fail(FRCNN_E_BADARG, "tcgen05 engine: operands and workspace must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(epi.scale) & 15) || (reinterpret_cast<uintptr_t>(epi.bias) & 15) ||
      (reinterpret_cast<uintptr_t>(epi.residual) & 15))
    return fail(FRCNN_E_BADARG, "tcgen05 engine: output, scale, bias and residual must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(a_split) & 127) || (reinterpret_cast<uintptr_t>(b_split) & 127))
    return fail(FRCNN_E_BADARG, "tcgen05 engine: operand split buffers must be 128-byte aligned");
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  const size_t esz = f16 ? 2 : 4, hdr = f16 ? kF16Header : 0;
  const void *a_hi = ws + p.a_hi_off, *a_lo = ws + p.a_lo_off, *b_hi = ws + p.b_hi_off, *b_lo = ws + p.b_lo_off;
  const int *a_exp = nullptr, *b_exp = nullptr;
  float *partial = reinterpret_cast<float *>(ws + p.partial_off);
  if (a_split) {
    a_hi = reinterpret_cast<const uint8_t *>(a_split) + hdr;
    a_lo = reinterpret_cast<const uint8_t *>(a_split) + hdr + align_up(p.a_count * esz, 1024);
  } else if (f16) {
    int rc = f16_split(a, p.a_count, ws + p.a_hi_off - hdr, st, nullptr, 0);
    if (rc != FRCNN_OK) return rc;
  } else {
    launch(split_hi_lo_kernel, elementwise_grid(p.a_count / 4 + 1, 256), 256, 0, st, a, (float *)a_hi, (float *)a_lo, p.a_count);
    FRCNN_CHECK_LAUNCH("split_hi_lo_kernel(a)");
  }
  if (b_split) {
    b_hi = reinterpret_cast<const uint8_t *>(b_split) + hdr;
    b_lo = reinterpret_cast<const uint8_t *>(b_split) + hdr + align_up(p.b_count * esz, 1024);
  } else if (f16) {
    int rc = f16_split(b, p.b_count, ws + p.b_hi_off - hdr, st, nullptr, 0);
    if (rc != FRCNN_OK) return rc;
  } else {
    launch(split_hi_lo_kernel, elementwise_grid(p.b_count / 4 + 1, 256), 256, 0, st, b, (float *)b_hi, (float *)b_lo, p.b_count);
    FRCNN_CHECK_LAUNCH("split_hi_lo_kernel(b)");
  }
  if (f16) {                                                          // scale exponents: header word 1 of each operand's split
    a_exp = reinterpret_cast<const int *>(reinterpret_cast<const uint8_t *>(a_hi) - hdr) + 1;
    b_exp = reinterpret_cast<const int *>(reinterpret_cast<const uint8_t *>(b_hi) - hdr) + 1;
  }

  const int taps = KH * KW, el = (int)esz, atom = 128 / el;
  const int bn_cta = p.pair ? p.BN / 2 : p.BN;                       // B columns one CTA stages (a pair's CTA: its half)
  CUtensorMap maps[4];
  bool ok;
  dim3 grid;
  if (mode == TC_FWD) {
    ok = make_act_map(&maps[0], a_hi, el, p.N, p.H, p.W, Cin, p.tile_w, p.tile_h, p.tile_n) && make_act_map(&maps[1], a_lo, el, p.N, p.H, p.W, Cin, p.tile_w, p.tile_h, p.tile_n) &&
         make_mat_map(&maps[2], b_hi, el, Cout, taps * Cin, bn_cta) && make_mat_map(&maps[3], b_lo, el, Cout, taps * Cin, bn_cta);
  } else if (mode == TC_DGRAD) {
    ok = make_act_map(&maps[0], a_hi, el, p.N, p.H, p.W, Cout, p.tile_w, p.tile_h, p.tile_n) && make_act_map(&maps[1], a_lo, el, p.N, p.H, p.W, Cout, p.tile_w, p.tile_h, p.tile_n) &&
         make_filter_map_atoms(&maps[2], b_hi, el, Cout, taps, Cin, bn_cta / atom) && make_filter_map_atoms(&maps[3], b_lo, el, Cout, taps, Cin, bn_cta / atom);
  } else {
    ok = make_act_map_atoms(&maps[0], a_hi, el, p.N, p.H, p.W, Cout, p.pw, p.ph, p.pn, 128 / atom) && make_act_map_atoms(&maps[1], a_lo, el, p.N, p.H, p.W, Cout, p.pw, p.ph, p.pn, 128 / atom) &&
         make_act_map_atoms(&maps[2], b_hi, el, p.N, p.H, p.W, Cin, p.pw, p.ph, p.pn, bn_cta / atom) && make_act_map_atoms(&maps[3], b_lo, el, p.N, p.H, p.W, Cin, p.pw, p.ph, p.pn, bn_cta / atom);
  }
  if (!ok) return fail(FRCNN_E_BADARG, "tcgen05 engine: cuTensorMapEncodeTiled failed");

  // MN-major descriptor strides: tf32 = 32-channel atoms of 32 k-rows (4 KB), 4-row K atoms (512 B), 8 k-rows per MMA (1 KB);
  // fp16 = 64-channel atoms of 64 k-rows (8 KB), 8-row K atoms (1 KB), 16 k-rows per MMA (2 KB)
  static const int dbg_lbo = getenv("FRCNN_TC_MN_LBO") ? atoi(getenv("FRCNN_TC_MN_LBO")) : 0;
  static const int dbg_sbo = getenv("FRCNN_TC_MN_SBO") ? atoi(getenv("FRCNN_TC_MN_SBO")) : 0;
  static const int dbg_kstep = getenv("FRCNN_TC_MN_KSTEP") ? atoi(getenv("FRCNN_TC_MN_KSTEP")) : 0;
  const int mn_lbo = dbg_lbo ? dbg_lbo : (f16 ? 8192 : kAtomBytes), mn_sbo = dbg_sbo ? dbg_sbo : (f16 ? 1024 : 512), mn_kstep = dbg_kstep ? dbg_kstep : (f16 ? 2048 : 1024);
  static std::atomic<unsigned long long> launch_serial{0};
  static const int pair_late = !(getenv("FRCNN_TC_PAIR_TRIGGER") && getenv("FRCNN_TC_PAIR_TRIGGER")[0] == 'e');
  // FRCNN_TC_PAIRED_STORES=0: plain per-row stores (the A/B switch of profiles/r02_pair_ab.md; paired stores are equal or faster on every layer)
  static const int paired_stores = getenv("FRCNN_TC_PAIRED_STORES") ? (atoi(getenv("FRCNN_TC_PAIRED_STORES")) != 0) : 1;
  grid = dim3(p.grid, 1, 1);
  TcGeom g{Cin, Cout, KH, KW, pad, p.H, p.W, p.N, p.tile_w, p.tile_h, p.tile_n, p.tiles_w, p.tiles_h, p.pw, p.ph, p.pn, p.patches_w, p.patches_h,
           p.kb_per_split, p.total_kb, p.splits, mn_lbo, mn_sbo, mn_kstep, p.m_tiles, p.n_tiles, p.items, p.m_pairs, pair_late, paired_stores, g_tc_trace,
           p.streamk, p.units, partial, reinterpret_cast<unsigned long long *>(ws + p.flags_off),
           0xF1A6000000000000ull | (++launch_serial & 0xFFFFFFFFFFFFull), a_exp, b_exp,
           (p.splits == 1 && mode != TC_WGRAD) ? reinterpret_cast<unsigned *>(amax_out) : nullptr, p.grid_max};
  int rc;
#define TC_LAUNCH(M)                                                                                                                      \
  (f16 ? (p.BN == 128 ? launch_tc<M, 128, 3, true>(maps, g, grid, out, partial, epi, st) : launch_tc<M, 64, 4, true>(maps, g, grid, out, partial, epi, st)) \
       : (p.BN == 128 ? launch_tc<M, 128, 3, false>(maps, g, grid, out, partial, epi, st) : launch_tc<M, 64, 4, false>(maps, g, grid, out, partial, epi, st)))
  if (p.pair) {
    if (mode == TC_FWD) rc = p.BN == 128 ? launch_tc<TC_FWD, 128, 4, true, true>(maps, g, grid, out, partial, epi, st) : launch_tc<TC_FWD, 64, 5, true, true>(maps, g, grid, out, partial, epi, st);
    else if (mode == TC_DGRAD) rc = launch_tc<TC_DGRAD, 128, 4, true, true>(maps, g, grid, out, partial, epi, st);
    else rc = launch_tc<TC_WGRAD, 128, 4, true, true>(maps, g, grid, out, partial, epi, st);
  }
  else if (mode == TC_FWD) rc = TC_LAUNCH(TC_FWD);
  else if (mode == TC_DGRAD) rc = TC_LAUNCH(TC_DGRAD);
  else rc = TC_LAUNCH(TC_WGRAD);
#undef TC_LAUNCH
  if (rc != FRCNN_OK) return rc;
  if (p.splits > 1) {
    const long long pixels = (long long)p.N * p.H * p.W;
    if (mode == TC_FWD) return launch_splitk_reduce(partial, out, (int)pixels, Cout, p.splits, epi, st);
    if (mode == TC_DGRAD) return launch_splitk_reduce(partial, out, (int)pixels, Cin, p.splits, epi, st);
    Epilogue none{nullptr, nullptr, nullptr, FRCNN_ACT_NONE};
    return launch_splitk_reduce(partial, out, Cout, taps * Cin, p.splits, none, st);
  }
  return FRCNN_OK;
}

#define GEOM_PARAMS int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad
#define GEOM_ARGS N, H, W, Cin, Cout, KH, KW, stride, pad

// mode: TC_FWD / TC_DGRAD / TC_WGRAD; f16: the fp16 engine (FRCNN_ENGINE_TC_3XF16) instead of tf32
bool tc_supported(int mode, GEOM_PARAMS, bool f16) { TcPlan p; return make_tc_plan(mode, GEOM_ARGS, &p, f16); }
size_t tc_workspace(int mode, GEOM_PARAMS, bool f16) { TcPlan p; return make_tc_plan(mode, GEOM_ARGS, &p, f16) ? p.total_bytes : 0; }
// number of per-CTA output maxima a fwd / dgrad launch of this shape writes when given an amax buffer (0: none -- split-K shapes, wgrad)
int tc_amax_slots(int mode, GEOM_PARAMS, bool f16) { TcPlan p; return (mode != TC_WGRAD && make_tc_plan(mode, GEOM_ARGS, &p, f16) && p.splits == 1) ? p.grid_max : 0; }
bool tc_fwd_supported(GEOM_PARAMS) { return tc_supported(TC_FWD, GEOM_ARGS, false); }
bool tc_dgrad_supported(GEOM_PARAMS) { return tc_supported(TC_DGRAD, GEOM_ARGS, false); }
bool tc_wgrad_supported(GEOM_PARAMS) { return tc_supported(TC_WGRAD, GEOM_ARGS, false); }
size_t tc_fwd_workspace(GEOM_PARAMS) { return tc_workspace(TC_FWD, GEOM_ARGS, false); }
size_t tc_dgrad_workspace(GEOM_PARAMS) { return tc_workspace(TC_DGRAD, GEOM_ARGS, false); }
size_t tc_wgrad_workspace(GEOM_PARAMS) { return tc_workspace(TC_WGRAD, GEOM_ARGS, false); }

int tc_conv2d_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual, float *y,
                  GEOM_PARAMS, int act, void *workspace, size_t workspace_bytes, cudaStream_t st, const void *x_split, const void *w_split, bool f16, void *amax_out)
{
  Epilogue epi{scale, bias, residual, act};
  return run_tc(TC_FWD, x, w, y, epi, GEOM_ARGS, workspace, workspace_bytes, st, x_split, w_split, f16, amax_out);
}

int tc_conv2d_dgrad(const float *dy, const float *w, const float *addend, float *dx, GEOM_PARAMS, void *workspace, size_t workspace_bytes, cudaStream_t st,
                    const void *dy_split, const void *w_split, bool f16, void *amax_out)
{
  Epilogue epi{nullptr, nullptr, addend, FRCNN_ACT_NONE};
  return run_tc(TC_DGRAD, dy, w, dx, epi, GEOM_ARGS, workspace, workspace_bytes, st, dy_split, w_split, f16, amax_out);
}

int tc_conv2d_wgrad(const float *dy, const float *x, float *dw, GEOM_PARAMS, void *workspace, size_t workspace_bytes, cudaStream_t st,
                    const void *dy_split, const void *x_split, bool f16)
{
  Epilogue none{nullptr, nullptr, nullptr, FRCNN_ACT_NONE};
  return run_tc(TC_WGRAD, dy, x, dw, none, GEOM_ARGS, workspace, workspace_bytes, st, dy_split, x_split, f16);
}

}  // namespace frcnn
