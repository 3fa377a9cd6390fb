// Persistent recurrent kernels of the PPO update: the T-step LSTM recurrence of a minibatch of whole
// world sequences - what sb3-contrib's RecurrentActorCriticPolicy._process_sequence does step by step with torch.nn.LSTM
// (reached from /root/reference/src/train/trainer.py:67-71 through RecurrentPPO.train) - without leaving the SM.
//
// Round 1 ran the recurrence as 2 x T dependent launch pairs per network (cuBLAS GEMM B x 4H x H, then the cell kernel;
// ~16 us per step, launch-latency bound). Here one thread-block CLUSTER owns a tile of 128 worlds for all T steps:
//   * CTA q of the cluster owns hidden units [64 q, 64 q + 64): its W_hh rows (4 gates x 64 units = 256 rows, bf16, packed into
//     the UMMA K-major core-matrix image once per minibatch) stay RESIDENT in shared memory (128 KB at H = 256);
//   * the A operand keep_t * h_{t-1} [128 x H] (bf16, same layout) sits in every CTA's shared memory; each step one thread issues
//     H / 16 tcgen05.mma (M = 128, N = 256, K = 16) into a 256-column TMEM accumulator;
//   * 8 epilogue warps (a thread = one world row x 32 units) read the gates with tcgen05.ld, add the input projection
//     x W_ih^T (precomputed for all steps by one library GEMM) and the biases, run the cell with c_t in REGISTERS for the whole
//     sequence, write gates / c_t / h_t for the backward pass, and store keep_{t+1} * h_t straight into the A tiles of ALL
//     CTAs of the cluster through distributed shared memory (st.shared::cluster);
//   * two cluster barriers per step order the exchange (all MMAs done reading A -> remote writes -> next MMA).
// The backward kernel mirrors it (see lstm_seq_bwd_kernel below).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include "../../include/myo_b200.h"
#include "myo_lstm_seq.hpp"

namespace myo { void set_error(const std::string& msg); }

namespace {

typedef __nv_bfloat16 bf16;
constexpr int TILE_M = 128;
constexpr int UNITS = 64;                 // hidden units per CTA
constexpr int NCOL = 4 * UNITS;           // gate columns per CTA = MMA N
constexpr int CHUNK_K = 32;
constexpr int CHUNK_BYTES = NCOL * CHUNK_K * 2;   // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int THREADS = EPI_WARPS * 32;           // 8 warps = 2 per scheduler: 255 registers per thread (a 9th warp would cap them at 168);
                                                  // thread 0 also allocates TMEM, loads the weights and issues the MMAs
constexpr int REC_BYTES = myo::kLstmRecBytes;     // activation record of (row, 8 units): 4 gates x 8 bf16 | c_t 8 fp32
constexpr int REC_CHUNKS = REC_BYTES / 16;
constexpr int STAGE_BYTES = 32 * REC_BYTES;       // per epilogue warp

// ---- PTX wrappers (same conventions as myo_policy.cu) ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16: D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 2.f / (1.f + __expf(-2.f * x)) - 1.f; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// element (n, k) of a 256 x K operand, K-major core matrices, chunks of CHUNK_K columns (the image cp.async.bulk copies verbatim)
__device__ __forceinline__ size_t pack_off(int n, int k) {
  const int chunk = k / CHUNK_K, kk = k % CHUNK_K;
  return (size_t)chunk * CHUNK_BYTES + (size_t)(kk >> 3) * NCOL * 16 + (size_t)(n >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
}

// forward weights: CTA q, column n = gate * 64 + j  <-  W_hh[gate * H + 64 q + j][k];  bias[q][n] = b_ih + b_hh of that row
__global__ void pack_whh_fwd_kernel(uint8_t* __restrict__ dst, float* __restrict__ bias, const float* __restrict__ whh, const float* __restrict__ bih,
                                    const float* __restrict__ bhh, int H) {
  const int NU = H / UNITS;
  const int total = NU * NCOL * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx % H, n = (idx / H) % NCOL, q = idx / (H * NCOL);
    const int srow = (n / UNITS) * H + q * UNITS + (n % UNITS);
    *reinterpret_cast<bf16*>(dst + (size_t)q * NCOL * H * 2 + pack_off(n, k)) = __float2bfloat16_rn(whh[(size_t)srow * H + k]);
    if (k == 0) bias[q * NCOL + n] = bih[srow] + bhh[srow];
  }
}

struct SeqFwdArgs {
  int T, B, H;
  const uint8_t* wpack;   // [H/64][256 x H] packed bf16
  const float* bias;      // [H/64][256]
  const float* G;         // [T][B][4H]: x W_ih^T (input projection, no bias)
  const float* keep;      // [T][B]
  const float* C0;        // [B][H] (already multiplied by keep_0)
  uint8_t* Rec;           // [T][B][H/8] records of REC_BYTES: activated gates i f g o (8 x bf16 each) | c_t (8 x fp32)
  bf16* Hs;               // [T][B][H]: h_t (unmasked)
  const bf16* HP0;        // [B][H]: keep_0 h0
  long long* prof;        // optional [8] cycle counters of CTA 0 / thread 0 (development)
  int dbg;                // development: bit 0 skip the x W_ih^T loads, bit 1 skip the global stores
};

// ordered, unpredicated 16-byte global loads / stores: issued where they are written (the compiler may not sink a load to its
// use, which serialised the x W_ih^T loads of a step into dependent round trips in the first version of this kernel)
__device__ __forceinline__ float4 ldg_v4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }
// shared::cta -> shared::cluster bulk copy (async proxy), completing bytes on an mbarrier of the destination CTA
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta),
               "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2_128(const void* p) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], 128;" ::"l"(p) : "memory"); }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return __fdividef(2.f, 1.f + __expf(-2.f * x)) - 1.f; }

__global__ void __launch_bounds__(THREADS, 1) lstm_seq_fwd_kernel(const __grid_constant__ SeqFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, NU = H / UNITS;
  const uint32_t q = cluster_ctarank();
  const int tile = blockIdx.x / NU;
  const int row0 = tile * TILE_M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sW = smem;
  uint8_t* sA = smem + (size_t)NCOL * H * 2;
  float* sBias = reinterpret_cast<float*>(sA + (size_t)TILE_M * H * 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + NCOL);
  // w: weights landed; acc: this step's MMAs complete; a: next step's A tile complete (own slice written + peers' slices landed)
  const uint32_t bar_w = smem_u32(bars), bar_acc = smem_u32(bars + 1), bar_a = smem_u32(bars + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  uint8_t* sStage = reinterpret_cast<uint8_t*>(bars + 8) + (size_t)warp * STAGE_BYTES;      // per epilogue warp: 32 rows x one record

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_acc, 1); mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < NCOL; i += THREADS) sBias[i] = a.bias[q * NCOL + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(NCOL * H * 2);
    mbar_expect_tx(bar_w, bytes);
    const uint8_t* src = a.wpack + (size_t)q * NCOL * H * 2;
    for (int c = 0; c < H / CHUNK_K; c++) bulk_g2s(smem_u32(sW + (size_t)c * CHUNK_BYTES), src + (size_t)c * CHUNK_BYTES, CHUNK_BYTES, bar_w);
  }
  // A(0) = keep_0 h0 tile, row-major global -> core-matrix layout: (k / 8) * 2048 + row * 16
  for (int idx = threadIdx.x; idx < TILE_M * (H / 8); idx += THREADS) {
    const int row = idx % TILE_M, kg = idx / TILE_M;
    const int b = row0 + row;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b < a.B) v = *reinterpret_cast<const uint4*>(a.HP0 + (size_t)b * H + kg * 8);
    *reinterpret_cast<uint4*>(sA + (size_t)kg * (TILE_M * 16) + row * 16) = v;
  }
  fence_proxy_async();
  cluster_arrive();
  cluster_wait();

  // epilogue role: a thread = world row x 32 units, walked in four blocks of 8 units (one 16-byte k-group of the A tile each)
  const int row = (warp & 3) * 32 + lane, half = (warp >> 2) & 1;
  const int b = row0 + row;
  const bool live = b < a.B;
  const int bl = live ? b : row0;                 // rows past the batch read a valid row and store nothing
  const int ucta = half * 32, ug = (int)q * UNITS + ucta;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float c[32];
#pragma unroll
  for (int u = 0; u < 32; u += 4) {
    float4 v = live ? *reinterpret_cast<const float4*>(a.C0 + (size_t)b * H + ug + u) : make_float4(0.f, 0.f, 0.f, 0.f);
    c[u] = v.x; c[u + 1] = v.y; c[u + 2] = v.z; c[u + 3] = v.w;
  }
  const uint32_t idesc = make_idesc(NCOL);
  const uint32_t a_lbo = TILE_M * 16, b_lbo = NCOL * 16;
  const uint32_t sA_u32 = smem_u32(sA);
  const uint32_t slice_off = (uint32_t)(8 * q) * (TILE_M * 16), slice_bytes = 8u * TILE_M * 16;     // this CTA's 64 units of the A tile
  // record copy-out: the warp's staging buffer (32 rows x REC_BYTES) leaves as REC_CHUNKS coalesced 16-byte chunks per lane;
  // chunk id = i * 32 + lane sits at staging byte 16 id and belongs to row id / REC_CHUNKS of the warp
  int rec_off[REC_CHUNKS];
  uint32_t rec_live = 0;
#pragma unroll
  for (int i = 0; i < REC_CHUNKS; i++) {
    const int id = i * 32 + lane, r = id / REC_CHUNKS;
    rec_off[i] = r * (H / 8) * REC_BYTES + (id % REC_CHUNKS) * 16;
    if (row0 + (warp & 3) * 32 + r < a.B) rec_live |= 1u << i;
  }
  // input projection x W_ih^T of this row: 4 blocks x (4 gates x 8 units) = 32 float4 in registers; block blk of step t + 1 is
  // requested as soon as block blk of step t has been consumed, so every load has a whole step to arrive
  float4 x[4][8];
  auto load_x = [&](float4 (&xb)[8], int t, int blk) {
    const float* g = a.G + ((int64_t)t * a.B + bl) * 4 * H + ug + 8 * blk;
    if (a.dbg & 1) { for (int gi = 0; gi < 8; gi++) xb[gi] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
#pragma unroll
    for (int gi = 0; gi < 4; gi++) { xb[2 * gi] = ldg_v4(g + gi * H); xb[2 * gi + 1] = ldg_v4(g + gi * H + 4); }
  };
#pragma unroll
  for (int blk = 0; blk < 4; blk++) load_x(x[blk], 0, blk);
  float kt_next = a.keep[bl];
  // this CTA's slice of h_t (unmasked, bf16, in the A tile) goes out to Hs[t]: a warp pass moves 4 rows x 128 bytes
  auto copy_out_h = [&](int t) {
#pragma unroll
    for (int i = 0; i < TILE_M / 4 / EPI_WARPS; i++) {
      const int r = 4 * (i * EPI_WARPS + warp) + (lane & 3), kg = lane >> 2;
      const uint4 v = *reinterpret_cast<const uint4*>(sA + slice_off + (uint32_t)kg * (TILE_M * 16) + r * 16);
      if (row0 + r < a.B) *reinterpret_cast<uint4*>(a.Hs + ((int64_t)t * a.B + row0 + r) * H + (int)q * UNITS + kg * 8) = v;
    }
  };

  const bool prof = a.prof && blockIdx.x == 0 && threadIdx.x == 0;
  long long pc[6] = {0, 0, 0, 0, 0, 0}, p0 = 0;
#define SEQ_PROF(i) if (prof) { const long long now = clock64(); pc[i] += now - p0; p0 = now; }
  for (int t = 0; t < a.T; t++) {
    if (prof) p0 = clock64();
    const bool more = t + 1 < a.T;
    if (threadIdx.x == 0) {
      if (t == 0) mbar_wait(bar_w, 0); else mbar_wait(bar_a, (uint32_t)((t - 1) & 1));
      tc_fence_after();
      for (int k0 = 0; k0 < H; k0 += 16) {
        const uint32_t a_addr = sA_u32 + (uint32_t)(k0 >> 3) * a_lbo;
        const uint32_t b_addr = smem_u32(sW) + (uint32_t)(k0 / CHUNK_K) * CHUNK_BYTES + (uint32_t)((k0 % CHUNK_K) >> 3) * b_lbo;
        umma_bf16(tmem_base, make_desc(a_addr, a_lbo, 128), make_desc(b_addr, b_lbo, 128), idesc, k0 > 0 ? 1u : 0u);
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
    if (t > 0 && !(a.dbg & 8)) copy_out_h(t - 1);   // overlaps the MMAs; the slice is rewritten only after cluster barrier #1 below
    mbar_wait(bar_acc, (uint32_t)(t & 1));     // this CTA's MMAs of step t are complete: its A tile may be overwritten
    tc_fence_after();
    SEQ_PROF(0)
    cluster_arrive();                           // #1 (waited for below, before anything is written into an A tile)
    bool waited = false;
    {
      const int64_t m = (int64_t)t * a.B + bl;
      const float kt = kt_next;                 // keep_t scales the recurrent term (the A tile holds the unmasked h_{t-1})
      if (more) kt_next = a.keep[m + a.B];
      uint8_t* rec_base = a.Rec + (((int64_t)t * a.B + row0 + (warp & 3) * 32) * (H / 8) + (8 * (int)q + 4 * half)) * REC_BYTES;
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        const int col = ucta + 8 * blk;
        uint32_t ri[8], rf[8], rg[8], ro[8];
        tmem_ld8(lane_base + 0 * UNITS + col, ri);
        tmem_ld8(lane_base + 1 * UNITS + col, rf);
        tmem_ld8(lane_base + 2 * UNITS + col, rg);
        tmem_ld8(lane_base + 3 * UNITS + col, ro);
        tmem_ld_wait();
        const float* xf = reinterpret_cast<const float*>(x[blk]);      // [gate][8]
        float gi_[8], gf_[8], gg_[8], go_[8], hn[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float ig = fast_sigmoid(kt * __uint_as_float(ri[u]) + xf[u] + sBias[0 * UNITS + col + u]);
          const float fg = fast_sigmoid(kt * __uint_as_float(rf[u]) + xf[8 + u] + sBias[1 * UNITS + col + u]);
          const float gv = fast_tanh(kt * __uint_as_float(rg[u]) + xf[16 + u] + sBias[2 * UNITS + col + u]);
          const float og = fast_sigmoid(kt * __uint_as_float(ro[u]) + xf[24 + u] + sBias[3 * UNITS + col + u]);
          const float cn = fg * (c[8 * blk + u] * kt) + ig * gv;
          c[8 * blk + u] = cn;
          hn[u] = og * fast_tanh(cn);
          gi_[u] = ig; gf_[u] = fg; gg_[u] = gv; go_[u] = og;
        }
        if (more) load_x(x[blk], t + 1, blk);
        if (!waited) { cluster_wait(); waited = true; }          // #1: every CTA of the cluster is done reading its A tile
        *reinterpret_cast<uint4*>(sA + slice_off + (uint32_t)(4 * half + blk) * (TILE_M * 16) + row * 16) =
            live ? make_uint4(pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]), pack_bf16(hn[4], hn[5]), pack_bf16(hn[6], hn[7])) : make_uint4(0u, 0u, 0u, 0u);
        if (!(a.dbg & 2)) {
          // record of (row, block) -> the warp's staging buffer, then out as coalesced 16-byte chunks
          __syncwarp();
          uint4* st = reinterpret_cast<uint4*>(sStage + lane * REC_BYTES);
          st[0] = make_uint4(pack_bf16(gi_[0], gi_[1]), pack_bf16(gi_[2], gi_[3]), pack_bf16(gi_[4], gi_[5]), pack_bf16(gi_[6], gi_[7]));
          st[1] = make_uint4(pack_bf16(gf_[0], gf_[1]), pack_bf16(gf_[2], gf_[3]), pack_bf16(gf_[4], gf_[5]), pack_bf16(gf_[6], gf_[7]));
          st[2] = make_uint4(pack_bf16(gg_[0], gg_[1]), pack_bf16(gg_[2], gg_[3]), pack_bf16(gg_[4], gg_[5]), pack_bf16(gg_[6], gg_[7]));
          st[3] = make_uint4(pack_bf16(go_[0], go_[1]), pack_bf16(go_[2], go_[3]), pack_bf16(go_[4], go_[5]), pack_bf16(go_[6], go_[7]));
          reinterpret_cast<float4*>(st)[4] = make_float4(c[8 * blk], c[8 * blk + 1], c[8 * blk + 2], c[8 * blk + 3]);
          reinterpret_cast<float4*>(st)[5] = make_float4(c[8 * blk + 4], c[8 * blk + 5], c[8 * blk + 6], c[8 * blk + 7]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < REC_CHUNKS; i++) {
            const uint4 v = *reinterpret_cast<const uint4*>(sStage + (i * 32 + lane) * 16);
            if (((rec_live >> i) & 1u) && !(a.dbg & 4)) *reinterpret_cast<uint4*>(rec_base + blk * REC_BYTES + rec_off[i]) = v;
          }
        }
      }
      SEQ_PROF(1)
      fence_proxy_async_smem();                 // h_t in this CTA's slice: visible to the bulk copies and the next MMAs
      tc_fence_before();
    }
    __syncthreads();                            // slice complete, TMEM reads of step t done
    SEQ_PROF(2)
    if (threadIdx.x == 0 && more) {
      // hand the slice to the peers with bulk copies that complete on THEIR mbarrier; own slice: this arrival
      mbar_expect_tx(bar_a, (uint32_t)(NU - 1) * slice_bytes);
      for (int d = 1; d < NU; d++) {
        const uint32_t dst = (q + (uint32_t)d) % (uint32_t)NU;
        bulk_s2s(mapa(sA_u32 + slice_off, dst), sA_u32 + slice_off, slice_bytes, mapa(bar_a, dst));
      }
    }
    SEQ_PROF(3)
  }
  if (prof) for (int i = 0; i < 6; i++) a.prof[i] = pc[i];
  copy_out_h(a.T - 1);
  // nobody leaves while a peer may still copy out of / into its shared memory
  cluster_arrive();
  cluster_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
}

// HP[t] = keep_t * Hs[t - 1] for t >= 1 (operand of the dW_hh GEMM; HP[0] comes from the stored states): 8 elements per thread
__global__ void make_hp_kernel(const bf16* __restrict__ Hs, const float* __restrict__ keep, bf16* __restrict__ HP, int64_t rows, int B, int H) {
  const int H8 = H / 8;
  const int64_t total = rows * H8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / H8 + B;              // destination row (t >= 1)
    const int j = (int)(i % H8) * 8;
    uint4 v = *reinterpret_cast<const uint4*>(Hs + (m - B) * H + j);
    if (keep[m] == 0.f) v = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(HP + m * H + j) = v;
  }
}

size_t fwd_smem_bytes(int H) { return (size_t)NCOL * H * 2 + (size_t)TILE_M * H * 2 + NCOL * sizeof(float) + 64 + (size_t)EPI_WARPS * STAGE_BYTES; }

}  // namespace

namespace myo {

bool lstm_seq_supported(int H) { return H >= UNITS && H % UNITS == 0 && H <= 256; }
size_t lstm_seq_wpack_bytes(int H) { return (size_t)4 * H * H * 2; }

int lstm_seq_forward(const LstmSeqFwd& f, cudaStream_t st) {
  const int H = f.H, NU = H / UNITS;
  if (!lstm_seq_supported(H)) { set_error("lstm_seq_forward: unsupported hidden size"); return MYO_E_LIMIT; }
  pack_whh_fwd_kernel<<<148, 256, 0, st>>>(f.wpack, f.bias, f.whh, f.bih, f.bhh, H);
  static bool attr_set = false;
  const size_t smem = fwd_smem_bytes(256);
  if (!attr_set) {
    if (cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error(std::string("lstm_seq_fwd_kernel: cudaFuncSetAttribute failed: ") + cudaGetErrorString(cudaGetLastError()));
      return MYO_E_CUDA;
    }
    attr_set = true;
  }
  SeqFwdArgs a{f.T, f.B, H, f.wpack, f.bias, f.G, f.keep, f.C0, f.Rec, reinterpret_cast<bf16*>(f.Hs), reinterpret_cast<const bf16*>(f.HP), f.prof, getenv("MYO_SEQ_DBG") ? atoi(getenv("MYO_SEQ_DBG")) : 0};
  const int tiles = (f.B + TILE_M - 1) / TILE_M;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(tiles * NU));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = fwd_smem_bytes(H);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)NU; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_seq_fwd_kernel, a);
  if (e != cudaSuccess) { set_error(std::string("lstm_seq_fwd_kernel launch: ") + cudaGetErrorString(e)); cudaGetLastError(); return MYO_E_CUDA; }
  if (f.T > 1) make_hp_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const bf16*>(f.Hs), f.keep, reinterpret_cast<bf16*>(f.HP), (int64_t)(f.T - 1) * f.B, f.B, H);
  return MYO_OK;
}

}  // namespace myo
