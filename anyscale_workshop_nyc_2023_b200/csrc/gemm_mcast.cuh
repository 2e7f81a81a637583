// Skinny decode GEMM without a reduction:  D[M,N] = A[M,K] * W[N,K]^T  for N, K <= 768.
//
// gemm_splitk.cuh keeps the per-SM operand traffic low by cutting K across a cluster, and pays for
// it with a DSMEM reduce-scatter (stores + release fence + cluster barrier: ~30 % of that kernel,
// profiles/decode_ncu_r1.md). Here the cluster is laid over N instead: 4 CTAs compute 4 adjacent
// 128 x 16 output tiles of the same 128 rows, and the A operand they share is fetched ONCE per
// cluster - each CTA requests a quarter of A's k-blocks with cp.async.bulk.tensor...multicast::cluster,
// which deposits the tile at the same shared-memory offset of all four CTAs and completes on each
// CTA's own mbarrier. Every CTA then holds the full-K A tile (<= 12 k-blocks x 16 KB) plus its own
// 16 weight rows (2 KB per k-block): 216 KB, no stage is ever reused, so there is no empty-barrier
// protocol at all; per-CTA L2 ingest is 49 + 24 KB and the accumulator (128 x 16 fp32) is final -
// the epilogue writes HBM directly.
//
// MEASURED (B200, FLAN-T5-base decode): slower than split-K, 201.4 vs 191.0 ms per batch. Multicast removes L2
// reads but every SM still has to take in the whole 196 KB A tile; the per-SM ingest rate (~50-60 B/clk), not
// the L2, is what bounds these kernels, and split-K divides exactly that. Kept as an opt-in
// (B200T5_MCAST=1) with its parity tests.
//
// Roles (192 threads): warp 0 = TMA (own weight rows before griddepcontrol.wait; its share of A,
// multicast, after), warp 1 = TMEM + tcgen05.mma (N = 16), warps 2..5 = epilogue.
#pragma once
#include "gemm.cuh"
#include "gemm_splitk.cuh"  // cluster PTX helpers

namespace b200 {

constexpr int kMcBN = 16;
constexpr int kMcCluster = 4;
constexpr int kMcMaxKb = 12;                                   // k-blocks resident at once (K <= 768)
constexpr int kMcStageBytes = kBM * kBK * 2 + kMcBN * kBK * 2;  // 16 KB + 2 KB
constexpr int kMcThreads = 192;
constexpr int kMcSmemBytes = kMcMaxKb * kMcStageBytes + 1024 /*align*/ + 256 /*barriers*/;

DEVINL void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// 16-column epilogues (same rounding contract as the 32-column chunk functors of gemm.cuh)
struct McStore {
  typedef EpiStore::Params Params;
  static DEVINL void apply(const Params& p, const uint32_t (&acc)[16], int m, int n0, int N) {
    if (n0 + 16 > N) return;
    uint4* d4 = reinterpret_cast<uint4*>(p.C + static_cast<size_t>(m) * p.ldc + n0);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = pack_act2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
    d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
  struct Pre {};
  static DEVINL void pre(const Params&, int, int, Pre&) {}
};
struct McResidual {
  typedef EpiResidual::Params Params;
  struct Pre {
    uint4 r[2];
  };
  static DEVINL void pre(const Params& p, int m, int n0, Pre& pr) {
    const uint4* r4 = reinterpret_cast<const uint4*>(p.R + static_cast<size_t>(m) * p.ld + n0);
    pr.r[0] = r4[0];
    pr.r[1] = r4[1];
  }
  static DEVINL void apply(const Params& p, const uint32_t (&acc)[16], int m, int n0, int N, const Pre& pr) {
    if (n0 + 16 > N) return;
    const uint32_t rw[8] = {pr.r[0].x, pr.r[0].y, pr.r[0].z, pr.r[0].w, pr.r[1].x, pr.r[1].y, pr.r[1].z, pr.r[1].w};
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y0 = act_round(__uint_as_float(acc[2 * i]));
      const float y1 = act_round(__uint_as_float(acc[2 * i + 1]));
      o[i] = pack_act2(act_lo(rw[i]) + y0, act_hi(rw[i]) + y1);
    }
    uint4* d4 = reinterpret_cast<uint4*>(p.C + static_cast<size_t>(m) * p.ld + n0);
    d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
};

// grid = (4, tiles_n / 4, tiles_m), cluster (4,1,1); tmA box 64 x 128 rows, tmB box 64 x 16 rows.
// Requires K <= 768 (12 k-blocks), N % 64 == 0.
template <bool kResidual>
__global__ void __cluster_dims__(kMcCluster, 1, 1) __launch_bounds__(kMcThreads, 1)
gemm_mcast_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                  typename EpiResidual::Params ep_res, typename EpiStore::Params ep_store) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMcMaxKb * kMcStageBytes);
  uint64_t* full = bars;  // [kMcMaxKb]
  uint64_t* tfull = bars + kMcMaxKb;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int n_tile = blockIdx.y * kMcCluster + rank, m_tile = blockIdx.z;
  const int kblocks = (K + kBK - 1) / kBK;  // <= kMcMaxKb (host)
  const int m0 = m_tile * kBM, n0 = n_tile * kMcBN;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kMcMaxKb; ++i) mbar_init(&full[i], 1);
      mbar_init(tfull, 1);
      mbar_fence_init();
      // arm every k-block's barrier now: A tile (from whichever CTA multicasts it) + own weight rows
      for (int i = 0; i < kblocks; ++i) mbar_arrive_expect_tx(&full[i], kMcStageBytes);
    }
    __syncwarp();
    tmem_alloc<32>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  cluster_arrive_release();  // barriers of all four CTAs are initialised and armed before any multicast is issued
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb)  // own weight rows: independent of the previous kernel
        tma_load_2d(smem + kb * kMcStageBytes + kBM * kBK * 2, &tmB, &full[kb], kb * kBK, n0);
    }
    __syncwarp();
    cluster_wait_acquire();
    if (lane == 0) {
      pdl_wait();
      for (int kb = rank; kb < kblocks; kb += kMcCluster)  // this CTA's share of A, delivered to all four
        tma_load_2d_mcast(smem + kb * kMcStageBytes, &tmA, &full[kb], kb * kBK, m0, static_cast<uint16_t>((1u << kMcCluster) - 1));
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    cluster_wait_acquire();
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_act(kBM, kMcBN, 0, 0);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[kb], 0);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + kb * kMcStageBytes);
        const uint64_t a_desc = make_desc_sw128_kmajor(a_addr);
        const uint64_t b_desc = make_desc_sw128_kmajor(a_addr + kBM * kBK * 2);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_f16_ss(tmem_base, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                       (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(tfull);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    cluster_wait_acquire();
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < M;
    pdl_wait();
    McResidual::Pre pre;
    if (kResidual && m_ok && n0 + 16 <= N) McResidual::pre(ep_res, m, n0, pre);
    mbar_wait(tfull, 0);
    tc_fence_after_sync();
    uint32_t acc[16];
    tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16), acc);
    tmem_ld_wait();
    if (m_ok) {
      if (kResidual) McResidual::apply(ep_res, acc, m, n0, N, pre);
      else McStore::apply(ep_store, acc, m, n0, N);
    }
    tc_fence_before_sync();
  }
  // a CTA's shared memory is a multicast target of its peers until every peer's loads have landed: all four
  // have passed their last full-barrier wait (MMA) before anyone leaves
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<32>(tmem_base);
  }
}

inline cudaError_t prepare_gemm_mcast() {
  cudaError_t e = cudaFuncSetAttribute(gemm_mcast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMcSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_mcast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMcSmemBytes);
}

inline bool gemm_mcast_supports(int N, int K) { return K % 8 == 0 && (K + kBK - 1) / kBK <= kMcMaxKb && N % (kMcBN * kMcCluster) == 0; }

template <bool kResidual>
cudaError_t launch_gemm_mcast(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                              const EpiResidual::Params& ep_res, const EpiStore::Params& ep_store, cudaStream_t stream,
                              bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kMcCluster, N / (kMcBN * kMcCluster), (M + kBM - 1) / kBM);
  cfg.blockDim = dim3(kMcThreads);
  cfg.dynamicSmemBytes = kMcSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (launch_priority() != 0) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = launch_priority();
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_mcast_kernel<kResidual>, tmA, tmB, M, N, K, ep_res, ep_store);
}

}  // namespace b200
