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
constexpr int THREADS = (EPI_WARPS + 1) * 32;     // + control warp (TMEM allocation, weight load, MMA issue)

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
  float* G;               // [T][B][4H]: x W_ih^T in, activated gates out
  const float* keep;      // [T][B]
  const float* C0;        // [B][H] (already multiplied by keep_0)
  float* Cs;              // [T][B][H]
  bf16* Hs;               // [T][B][H]
  bf16* HP;               // [T][B][H]: HP[0] in (keep_0 h0), HP[t+1] = keep_{t+1} h_t out
  long long* prof;        // optional [8] cycle counters of CTA 0 / thread 0 (development)
};

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
  const uint32_t bar_w = smem_u32(bars), bar_acc = smem_u32(bars + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < NCOL; i += THREADS) sBias[i] = a.bias[q * NCOL + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == EPI_WARPS && lane == 0) {
    const uint32_t bytes = (uint32_t)(NCOL * H * 2);
    mbar_expect_tx(bar_w, bytes);
    const uint8_t* src = a.wpack + (size_t)q * NCOL * H * 2;
    for (int c = 0; c < H / CHUNK_K; c++) bulk_g2s(smem_u32(sW + (size_t)c * CHUNK_BYTES), src + (size_t)c * CHUNK_BYTES, CHUNK_BYTES, bar_w);
  }
  // A(0) = HP[0] tile, row-major global -> core-matrix layout: (k / 8) * 2048 + row * 16
  for (int idx = threadIdx.x; idx < TILE_M * (H / 8); idx += THREADS) {
    const int row = idx % TILE_M, kg = idx / TILE_M;
    const int b = row0 + row;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b < a.B) v = *reinterpret_cast<const uint4*>(a.HP + (size_t)b * H + kg * 8);
    *reinterpret_cast<uint4*>(sA + (size_t)kg * (TILE_M * 16) + row * 16) = v;
  }
  fence_proxy_async();
  cluster_arrive();
  cluster_wait();

  // epilogue role: a thread = world row x 32 units
  const int row = (warp & 3) * 32 + lane, half = (warp >> 2) & 1;
  const int b = row0 + row;
  const bool epi = warp < EPI_WARPS, live = epi && b < a.B;
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

  const bool prof = a.prof && blockIdx.x == 0 && threadIdx.x == 0;
  long long pc[6] = {0, 0, 0, 0, 0, 0}, p0 = 0;
#define SEQ_PROF(i) if (prof) { const long long now = clock64(); pc[i] += now - p0; p0 = now; }
  for (int t = 0; t < a.T; t++) {
    if (prof) p0 = clock64();
    if (warp == EPI_WARPS) {
      if (lane == 0) {
        if (t == 0) mbar_wait(bar_w, 0);
        fence_proxy_async();
        tc_fence_after();
        for (int k0 = 0; k0 < H; k0 += 16) {
          const uint32_t a_addr = sA_u32 + (uint32_t)(k0 >> 3) * a_lbo;
          const uint32_t b_addr = smem_u32(sW) + (uint32_t)(k0 / CHUNK_K) * CHUNK_BYTES + (uint32_t)((k0 % CHUNK_K) >> 3) * b_lbo;
          umma_bf16(tmem_base, make_desc(a_addr, a_lbo, 128), make_desc(b_addr, b_lbo, 128), idesc, k0 > 0 ? 1u : 0u);
        }
        umma_commit(bar_acc);
      }
      __syncwarp();
    }
    mbar_wait(bar_acc, (uint32_t)(t & 1));     // this CTA's MMAs of step t are complete: its A tile may be overwritten
    tc_fence_after();
    SEQ_PROF(0)
    cluster_arrive();                           // #1
    uint32_t hp[16];
    if (epi) {
      const int64_t m = (int64_t)t * a.B + (live ? b : 0);
      const float kt = live ? a.keep[m] : 0.f;
      const bool more = t + 1 < a.T;
      const float kn = (live && more) ? a.keep[m + a.B] : 0.f;
#pragma unroll
      for (int jb = 0; jb < 2; jb++) {
        const int col = ucta + jb * 16;
        uint32_t ri[16], rf[16], rg[16], ro[16];
        tmem_ld16(lane_base + 0 * UNITS + col, ri);
        tmem_ld16(lane_base + 1 * UNITS + col, rf);
        tmem_ld16(lane_base + 2 * UNITS + col, rg);
        tmem_ld16(lane_base + 3 * UNITS + col, ro);
        float* g = a.G + m * 4 * H + ug + jb * 16;
        float xi[16], xf[16], xg[16], xo[16];
#pragma unroll
        for (int v4 = 0; v4 < 4; v4++) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 vi = live ? *reinterpret_cast<const float4*>(g + 4 * v4) : z;
          const float4 vf = live ? *reinterpret_cast<const float4*>(g + H + 4 * v4) : z;
          const float4 vg = live ? *reinterpret_cast<const float4*>(g + 2 * H + 4 * v4) : z;
          const float4 vo = live ? *reinterpret_cast<const float4*>(g + 3 * H + 4 * v4) : z;
          xi[4 * v4] = vi.x; xi[4 * v4 + 1] = vi.y; xi[4 * v4 + 2] = vi.z; xi[4 * v4 + 3] = vi.w;
          xf[4 * v4] = vf.x; xf[4 * v4 + 1] = vf.y; xf[4 * v4 + 2] = vf.z; xf[4 * v4 + 3] = vf.w;
          xg[4 * v4] = vg.x; xg[4 * v4 + 1] = vg.y; xg[4 * v4 + 2] = vg.z; xg[4 * v4 + 3] = vg.w;
          xo[4 * v4] = vo.x; xo[4 * v4 + 1] = vo.y; xo[4 * v4 + 2] = vo.z; xo[4 * v4 + 3] = vo.w;
        }
        tmem_ld_wait();
        float hn[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
          const float ig = sigmoidf_(__uint_as_float(ri[u]) + xi[u] + sBias[0 * UNITS + col + u]);
          const float fg = sigmoidf_(__uint_as_float(rf[u]) + xf[u] + sBias[1 * UNITS + col + u]);
          const float gv = tanhf_(__uint_as_float(rg[u]) + xg[u] + sBias[2 * UNITS + col + u]);
          const float og = sigmoidf_(__uint_as_float(ro[u]) + xo[u] + sBias[3 * UNITS + col + u]);
          const float cn = fg * (c[jb * 16 + u] * kt) + ig * gv;
          c[jb * 16 + u] = cn;
          hn[u] = og * tanhf_(cn);
          xi[u] = ig; xf[u] = fg; xg[u] = gv; xo[u] = og;
        }
        if (live) {
#pragma unroll
          for (int v4 = 0; v4 < 4; v4++) {
            *reinterpret_cast<float4*>(g + 4 * v4) = make_float4(xi[4 * v4], xi[4 * v4 + 1], xi[4 * v4 + 2], xi[4 * v4 + 3]);
            *reinterpret_cast<float4*>(g + H + 4 * v4) = make_float4(xf[4 * v4], xf[4 * v4 + 1], xf[4 * v4 + 2], xf[4 * v4 + 3]);
            *reinterpret_cast<float4*>(g + 2 * H + 4 * v4) = make_float4(xg[4 * v4], xg[4 * v4 + 1], xg[4 * v4 + 2], xg[4 * v4 + 3]);
            *reinterpret_cast<float4*>(g + 3 * H + 4 * v4) = make_float4(xo[4 * v4], xo[4 * v4 + 1], xo[4 * v4 + 2], xo[4 * v4 + 3]);
            *reinterpret_cast<float4*>(a.Cs + m * H + ug + jb * 16 + 4 * v4) =
                make_float4(c[jb * 16 + 4 * v4], c[jb * 16 + 4 * v4 + 1], c[jb * 16 + 4 * v4 + 2], c[jb * 16 + 4 * v4 + 3]);
          }
          uint4 p0, p1;
          p0.x = pack_bf16(hn[0], hn[1]); p0.y = pack_bf16(hn[2], hn[3]); p0.z = pack_bf16(hn[4], hn[5]); p0.w = pack_bf16(hn[6], hn[7]);
          p1.x = pack_bf16(hn[8], hn[9]); p1.y = pack_bf16(hn[10], hn[11]); p1.z = pack_bf16(hn[12], hn[13]); p1.w = pack_bf16(hn[14], hn[15]);
          bf16* hs = a.Hs + m * H + ug + jb * 16;
          *reinterpret_cast<uint4*>(hs) = p0; *reinterpret_cast<uint4*>(hs + 8) = p1;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) hp[jb * 8 + u] = pack_bf16(kn * hn[2 * u], kn * hn[2 * u + 1]);
        if (live && more) {
          bf16* hpn = a.HP + (m + a.B) * H + ug + jb * 16;
          *reinterpret_cast<uint4*>(hpn) = make_uint4(hp[jb * 8], hp[jb * 8 + 1], hp[jb * 8 + 2], hp[jb * 8 + 3]);
          *reinterpret_cast<uint4*>(hpn + 8) = make_uint4(hp[jb * 8 + 4], hp[jb * 8 + 5], hp[jb * 8 + 6], hp[jb * 8 + 7]);
        }
      }
    }
    SEQ_PROF(1)
    cluster_wait();                             // #1: every CTA of the cluster is done reading its A tile
    SEQ_PROF(2)
    if (epi && t + 1 < a.T) {
      // keep_{t+1} h_t of this thread's 32 units -> k-groups 8 q + 4 half + j of the A tile of every CTA of the cluster
      for (int dst = 0; dst < NU; dst++) {
        const uint32_t base = mapa(sA_u32, (uint32_t)dst) + (uint32_t)(8 * q + 4 * half) * (TILE_M * 16) + (uint32_t)row * 16;
#pragma unroll
        for (int j = 0; j < 4; j++) st_cluster_v4(base + (uint32_t)j * (TILE_M * 16), hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
      }
    }
    SEQ_PROF(3)
    fence_proxy_async();
    tc_fence_before();
    SEQ_PROF(4)
    cluster_arrive();                           // #2: A(t + 1) complete everywhere, TMEM reads of step t done
    cluster_wait();
    SEQ_PROF(5)
  }
  if (prof) for (int i = 0; i < 6; i++) a.prof[i] = pc[i];
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
}

size_t fwd_smem_bytes(int H) { return (size_t)NCOL * H * 2 + (size_t)TILE_M * H * 2 + NCOL * sizeof(float) + 64; }

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
  SeqFwdArgs a{f.T, f.B, H, f.wpack, f.bias, f.G, f.keep, f.C0, f.Cs, reinterpret_cast<bf16*>(f.Hs), reinterpret_cast<bf16*>(f.HP), f.prof};
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
  return MYO_OK;
}

}  // namespace myo
