// Persistent warp-specialised bf16 GEMM for sm_100a:
//   D[M,N] = A[M,K] * W[N,K]^T      (both operands K-major, i.e. nn.Linear layout)
// TMA (128-B swizzle) -> smem ring -> tcgen05.mma (one elected thread) -> fp32
// accumulators double-buffered in TMEM -> epilogue warps (tcgen05.ld) apply a
// fused epilogue functor and write HBM directly.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 = tile rows).
//
// The epilogue functors are where the T5 rounding contract lives: HF eager bf16
// rounds every Linear output to bf16 before anything else touches it
// (transformers/models/t5/modeling_t5.py:277,298-299,338),
// so each functor first rounds the fp32 accumulator to bf16 and only then fuses
// the residual add / GeGLU / KV scatter / arg-max.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;      // 2 control warps + 4 epilogue warps
constexpr int kGemmThreadsWide = 320;  // 2 control warps + 8 epilogue warps (big tiles: the epilogue is the bottleneck)
constexpr int kEpiSmemBytes = 8192;  // per-CTA scratch the epilogue functor may stage tables in

template <int BN>
struct GemmCfg {
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN");
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // Big tiles (encoder, M = B*S) take the whole SM; the small-N decode tiles keep <= ~100 KB so that
  // two CTAs of concurrently running decode chains can share an SM.
  static constexpr int kStagesRaw = (196 * 1024) / kStageBytes;
  static constexpr int kStages = BN <= 64 ? ((100 * 1024) / kStageBytes) : (kStagesRaw > 8 ? 8 : kStagesRaw);
  static constexpr int kTmemCols = (2 * BN) < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiSmemBytes;
};

struct TileCoord {
  int m_tile, n_tile;
};
DEVINL TileCoord tile_coord(int tile, int tiles_m, int tiles_n, int m_fastest) {
  TileCoord c;
  if (m_fastest) {
    c.m_tile = tile % tiles_m;
    c.n_tile = tile / tiles_m;
  } else {
    c.n_tile = tile % tiles_n;
    c.m_tile = tile / tiles_n;
  }
  return c;
}

// Epilogue warps: a warp may only touch the TMEM lane quarter 32*(warp%4), so 4 warps cover a tile's
// 128 rows; the big tiles (BN = 256) run 8 epilogue warps - two per lane quarter, each taking half of
// the tile's 32-column chunks - because with 4 the epilogue (not the MMA) bounds the kernel: measured
// on B200, tensor pipe active 27 % for the GeGLU tile and 34 % for the K = 768 residual tile
// (profiles/encoder_ncu_r1.md).
template <int BN>
struct EpiWarps {
  static constexpr int kCount = BN >= 256 ? 8 : 4;
  static constexpr int kThreads = 64 + 32 * kCount;
};

template <int BN, class Epi>
__global__ void __launch_bounds__(EpiWarps<BN>::kThreads, (BN <= 64 ? 2 : 1))
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                    int K, int m_fastest, typename Epi::Params ep) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* epi_smem = smem + Cfg::kStages * Cfg::kStageBytes + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + kBM - 1) / kBM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kblocks = (K + kBK - 1) / kBK;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::kStages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], EpiWarps<BN>::kCount);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail. Each role waits for the previous
  // kernel (griddepcontrol.wait) only right before it first touches memory that kernel may have
  // written; the weights (B operand) never depend on it and are prefetched into L2 before the wait.

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (static_cast<int>(blockIdx.x) < num_tiles) {
        const TileCoord tc0 = tile_coord(blockIdx.x, tiles_m, tiles_n, m_fastest);
        for (int kb = 0; kb < kblocks; ++kb) tma_prefetch_l2_2d(&tmB, kb * kBK, tc0.n_tile * BN);
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(tile, tiles_m, tiles_n, m_fastest);
        const int m0 = tc.m_tile * kBM, n0 = tc.n_tile * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kABytes;
          mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          tma_load_2d(sA, &tmA, &full[stage], kb * kBK, m0);
          tma_load_2d(sB, &tmB, &full[stage], kb * kBK, n0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_act(kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t a_desc = make_desc_sw128_kmajor(a_addr);
          const uint64_t b_desc = make_desc_sw128_kmajor(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr>>4) field
            umma_f16_ss(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    constexpr int kParts = EpiWarps<BN>::kCount / 4;  // column parts of a tile row
    const int part = (warp - 2) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    Epi::prologue(ep, epi_smem, static_cast<int>(threadIdx.x) - 64, 32 * EpiWarps<BN>::kCount);  // constant tables; overlaps the main loop
    pdl_wait();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = tile_coord(tile, tiles_m, tiles_n, m_fastest);
      const int m = tc.m_tile * kBM + q * 32 + lane;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      Epi::template run<BN>(ep, taddr, m, m < M, tc.n_tile, N, epi_smem, part, kParts);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ======================================================================== epilogues
// Every functor works on CHUNKS: 32 consecutive fp32 accumulator columns of one output row
// (acc[] holds their bit patterns). `chunk_pre` fetches whatever the chunk needs that does not
// depend on the accumulator (so callers can issue it early); `chunk` applies the T5 rounding
// contract and writes HBM. Paired functors (GeGLU) consume two chunks: gate and up.
// The persistent kernel above feeds chunks straight from TMEM (`run`); the split-K kernel
// (gemm_splitk.cuh) feeds them from the cluster-reduced partial sums.

DEVINL void round_pack_32(const uint32_t (&acc)[32], uint32_t (&out)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = pack_act2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
}

// Store 32 bf16 (64 B) to dst; columns [n0, n0+32) clipped to N in groups of 8.
DEVINL void store_row_chunk(act_t* dst, const uint32_t (&p)[16], int n0, int N) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (n0 + g * 8 + 8 <= N) d4[g] = make_uint4(p[4 * g], p[4 * g + 1], p[4 * g + 2], p[4 * g + 3]);
  }
}

// Store 32 fp32 (128 B) to dst, clipped to N in groups of 4.
DEVINL void store_row_chunk_f32(float* dst, const float (&v)[32], int n0, int N) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (n0 + g * 4 + 4 <= N) d4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  }
}

struct NoPre {};

// Drives an unpaired functor over the BN columns of this thread's TMEM row, fetching the
// pre-operands of chunk c+1 before chunk c is processed.
// `part` of `parts`: the epilogue warp handles chunks [part*C/parts, (part+1)*C/parts) of the C = BN/32 chunks.
template <int BN, class Epi>
DEVINL void run_chunks_from_tmem(const typename Epi::Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N,
                                 const uint8_t* epi_smem, int part, int parts) {
  constexpr int C = BN / 32;
  const int c_lo = part * C / parts, c_hi = (part + 1) * C / parts;
  typename Epi::ChunkPre pre[2];
  if (m_ok && n_tile * BN + c_lo * 32 < N) Epi::chunk_pre(p, m, n_tile * BN + c_lo * 32, N, pre[0]);
#pragma unroll 1
  for (int c = c_lo; c < c_hi; c += 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (c + u < c_hi) {
        uint32_t acc[32];
        tmem_ld_32x32(taddr + (c + u) * 32, acc);
        const int n0 = n_tile * BN + (c + u) * 32;
        if (m_ok && c + u + 1 < c_hi && n0 + 32 < N) Epi::chunk_pre(p, m, n0 + 32, N, pre[(u + 1) & 1]);
        tmem_ld_wait();
        if (m_ok && n0 < N) Epi::chunk(p, acc, m, n0, N, epi_smem, pre[u & 1]);
      }
    }
  }
}

// ---- plain store: C = bf16(acc)
struct EpiStore {
  struct Params {
    act_t* C;
    int ldc;
  };
  static constexpr bool kPaired = false;
  typedef NoPre ChunkPre;
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  static DEVINL void chunk_pre(const Params&, int, int, int, ChunkPre&) {}
  static DEVINL void chunk(const Params& p, const uint32_t (&acc)[32], int m, int n0, int N, const uint8_t*,
                           const ChunkPre&) {
    uint32_t o[16];
    round_pack_32(acc, o);
    store_row_chunk(p.C + static_cast<size_t>(m) * p.ldc + n0, o, n0, N);
  }
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t* es,
                         int part, int parts) {
    run_chunks_from_tmem<BN, EpiStore>(p, taddr, m, m_ok, n_tile, N, es, part, parts);
  }
};

#if B200T5_F16
// ---- residual, fp16 build: the stream is fp32 (res_t = float) and torch's type promotion decides the arithmetic
// (modeling_t5.py T5LayerSelfAttention / T5LayerCrossAttention / T5LayerFF.forward):
//   attention output projection (fp16 Linear):  C = R + float(fp16(acc))            round_acc = 1
//     ... while the stream is still fp16, i.e. before the first feed-forward block:  C = fp16(R + fp16(acc))   round_out = 1
//   feed-forward `wo` (fp32 weight, fp32 output; _keep_in_fp32_modules): C = R + acc   round_acc = 0
struct EpiResidual {
  struct Params {
    res_t* C;
    const res_t* R;
    int ld;
    float* ss = nullptr;  // (fused RMSNorm is a bf16-build experiment)
    int ss_ld = 0;
    int round_acc = 1;
    int round_out = 0;
  };
  static constexpr bool kPaired = false;
  struct ChunkPre {
    float4 r[8];
  };
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  static DEVINL void chunk_pre(const Params& p, int m, int n0, int N, ChunkPre& pre) {
    const float4* r4 = reinterpret_cast<const float4*>(p.R + static_cast<size_t>(m) * p.ld + n0);
#pragma unroll
    for (int g = 0; g < 8; ++g) pre.r[g] = (n0 + g * 4 + 4 <= N) ? r4[g] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  static DEVINL void chunk(const Params& p, const uint32_t (&acc)[32], int m, int n0, int N, const uint8_t*,
                           const ChunkPre& pre) {
    float o[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float rw[4] = {pre.r[g].x, pre.r[g].y, pre.r[g].z, pre.r[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y = __uint_as_float(acc[g * 4 + j]);
        if (p.round_acc) y = act_round(y);
        float v = rw[j] + y;
        if (p.round_out) v = act_round(v);
        o[g * 4 + j] = v;
      }
    }
    store_row_chunk_f32(p.C + static_cast<size_t>(m) * p.ld + n0, o, n0, N);
  }
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t* es,
                         int part, int parts) {
    run_chunks_from_tmem<BN, EpiResidual>(p, taddr, m, m_ok, n_tile, N, es, part, parts);
  }
};
#else
// ---- residual: C = bf16( float(R) + float(bf16(acc)) )   (modeling_t5.py:375,406,149)
struct EpiResidual {
  struct Params {
    act_t* C;
    const act_t* R;
    int ld;
    // optional: sum of squares of the 32 outputs of each (row, chunk) -> ss[m * ss_ld + n0 / 32], for a
    // consumer GEMM that applies the following RMSNorm to its A operand (gemm_splitk.cuh, NormA)
    float* ss = nullptr;
    int ss_ld = 0;
    int round_acc = 1, round_out = 1;  // (fp16 build only; this build always rounds both)
  };
  static constexpr bool kPaired = false;
  struct ChunkPre {
    uint4 r[4];
  };
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  // The residual operand does not depend on the accumulator: it is fetched while the main loop /
  // the previous chunk is still in flight.
  static DEVINL void chunk_pre(const Params& p, int m, int n0, int N, ChunkPre& pre) {
    const uint4* r4 = reinterpret_cast<const uint4*>(p.R + static_cast<size_t>(m) * p.ld + n0);
#pragma unroll
    for (int g = 0; g < 4; ++g) pre.r[g] = (n0 + g * 8 + 8 <= N) ? r4[g] : make_uint4(0, 0, 0, 0);
  }
  static DEVINL void chunk(const Params& p, const uint32_t (&acc)[32], int m, int n0, int N, const uint8_t*,
                           const ChunkPre& pre) {
    uint32_t o[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t rw[4] = {pre.r[g].x, pre.r[g].y, pre.r[g].z, pre.r[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y0 = act_round(__uint_as_float(acc[g * 8 + 2 * j]));
        const float y1 = act_round(__uint_as_float(acc[g * 8 + 2 * j + 1]));
        o[g * 4 + j] = pack_act2(act_lo(rw[j]) + y0, act_hi(rw[j]) + y1);
      }
    }
    store_row_chunk(p.C + static_cast<size_t>(m) * p.ld + n0, o, n0, N);
    if (p.ss) {
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (n0 + 2 * i + 2 <= N) {
          const float a = act_lo(o[i]), b = act_hi(o[i]);
          sq = fmaf(a, a, sq);
          sq = fmaf(b, b, sq);
        }
      }
      p.ss[static_cast<size_t>(m) * p.ss_ld + (n0 >> 5)] = sq;
    }
  }
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t* es,
                         int part, int parts) {
    run_chunks_from_tmem<BN, EpiResidual>(p, taddr, m, m_ok, n_tile, N, es, part, parts);
  }
};

#endif

// gelu_new exactly as HF eager evaluates it on bf16 tensors: every elementwise op rounds its
// result to bf16 (transformers/activations.py:59-66). torch.pow(x, 3.0) on
// a bf16 tensor is x*x*x in bf16 arithmetic (two roundings, pow_mode 0; verified exhaustively
// against torch on the GPU); pow_mode 1 keeps the single-rounding variant selectable.
DEVINL float gelu_new_act_exact(float x, int pow_mode) {
  const float half_x = act_round(0.5f * x);
  // (torch.pow on an fp16 tensor computes in fp32 and rounds once: oracle/t5_oracle.py gelu_new)
  const float x3 = (pow_mode == 0 && !B200T5_F16) ? act_round(act_round(x * x) * x) : act_round(x * x * x);
  const float t1 = act_round(0.044715f * x3);
  const float t2 = act_round(x + t1);
  const float t3 = act_round(0.7978845608028654f * t2);
  const float t4 = act_round(tanhf(t3));
  const float t5 = act_round(1.0f + t4);
  return act_round(half_x * t5);
}

// The input of gelu_new is itself a bf16 value, so the function has only 65536 possible
// arguments: it is tabulated once per device with the exact arithmetic above. Only magnitudes in
// [lo, hi) need the table (a few thousand entries, staged in shared memory by the epilogue);
// below lo the result is bf16(0.5*x) and above hi it is x (positive) or -0 (negative), which the
// host verifies entry by entry when it derives lo/hi from the full table (b200t5.cu).
struct GeluLut {
  const uint16_t* table;  // device: [2][hi - lo] bf16 bits, sign-major
  int lo, hi;             // magnitude bit patterns (bf16 bits & 0x7fff)
};

// Same function without divergent branches (the GEMM epilogue evaluates it 128 times per thread and tile):
// the three cases are computed side by side and selected.
DEVINL float gelu_from_lut_sel(float x, const uint16_t* lut, int lo, int n /* = hi - lo */) {
  const uint32_t bits = __float_as_uint(x) >> 16;
  const int mag = static_cast<int>(bits & 0x7FFFu);
  const bool neg = (bits >> 15) != 0;
  const int rel = mag - lo;
  const int idx = min(max(rel, 0), n - 1) + (neg ? n : 0);
  const float tab = __uint_as_float(static_cast<uint32_t>(lut[idx]) << 16);
  const float half = act_round(0.5f * x);
  float sat = neg ? (mag == 0x7F80 ? __int_as_float(0x7FC00000) : -0.0f) : x;
  sat = mag > 0x7F80 ? x : sat;
  return rel < 0 ? half : (rel >= n ? sat : tab);
}

DEVINL float gelu_from_lut(float x, const uint16_t* lut, int lo, int hi) {
  const uint32_t bits = __float_as_uint(x) >> 16;
  const int mag = static_cast<int>(bits & 0x7FFFu);
  const int neg = static_cast<int>(bits >> 15);
  if (mag < lo) return act_round(0.5f * x);  // tanh term rounds away: gelu_new(x) == bf16(0.5*x)
  if (mag >= hi) {
    if (mag > 0x7F80) return x;                     // NaN propagates
    if (!neg) return x;                             // tanh saturated: 0.5x * 2
    return mag == 0x7F80 ? __int_as_float(0x7FC00000) : -0.0f;  // 0.5x * (1 + -1): -inf*0 = NaN, else -0
  }
  return __uint_as_float(static_cast<uint32_t>(lut[neg * (hi - lo) + (mag - lo)]) << 16);
}

// ---- GeGLU: tile columns [0,BN/2) are wi_0 (gate) features, [BN/2,BN) the matching
// wi_1 features (weights are interleaved per tile at finalize).
//   out = bf16( gelu_new(bf16(gate)) * bf16(up) )          (modeling_t5.py:115-118)
struct EpiGeglu {
  struct Params {
    ffh_t* out;  // [M, F]
    int F;
    GeluLut lut;
  };
  static constexpr bool kPaired = true;
  typedef NoPre ChunkPre;
  // stage the gelu table (a few KB) with 16-byte loads; `nthreads` epilogue threads take part
  // (named barrier 2 is reserved for them)
  static DEVINL void prologue(const Params& p, uint8_t* epi_smem, int tid, int nthreads = 128) {
#if B200T5_F16
    return;  // fp16 build: gelu_new is evaluated directly (the table trick below indexes bf16 bit patterns)
#endif
    const int nvec = (2 * (p.lut.hi - p.lut.lo) * 2 + 15) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(p.lut.table);
    uint4* dst = reinterpret_cast<uint4*>(epi_smem);
    for (int i = tid; i < nvec; i += nthreads) dst[i] = src[i];
    asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory");
  }
  // g/u: gate and up accumulators of features [f0, f0+32)
  static DEVINL void chunk2(const Params& p, const uint32_t (&g)[32], const uint32_t (&u)[32], int m, int f0,
                            const uint8_t* epi_smem) {
#if B200T5_F16
    // out = float(fp16(gelu_new(fp16 gate) * fp16 up)): the fp16 product cast up for the fp32 `wo` GEMM
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = act_round(__uint_as_float(g[i]));
      const float lin = act_round(__uint_as_float(u[i]));
      o[i] = act_round(gelu_new_act_exact(x, 1) * lin);
    }
    store_row_chunk_f32(p.out + static_cast<size_t>(m) * p.F + f0, o, f0, p.F);
#else
    const uint16_t* lut = reinterpret_cast<const uint16_t*>(epi_smem);
    const int lo = p.lut.lo, n = p.lut.hi - p.lut.lo;
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float x = act_round(__uint_as_float(g[2 * i + e]));
        const float lin = act_round(__uint_as_float(u[2 * i + e]));
        r[e] = (n > 0 ? gelu_from_lut_sel(x, lut, lo, n) : gelu_from_lut(x, lut, lo, p.lut.hi)) * lin;
      }
      o[i] = pack_act2(r[0], r[1]);
    }
    store_row_chunk(p.out + static_cast<size_t>(m) * p.F + f0, o, f0, p.F);
#endif
  }
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int /*N*/, const uint8_t* epi_smem,
                         int part, int parts) {
    constexpr int HALF = BN / 2;
    constexpr int C = HALF / 32;
#pragma unroll 1
    for (int c = part * C / parts; c < (part + 1) * C / parts; ++c) {
      uint32_t g[32], u[32];
      tmem_ld_32x32(taddr + c * 32, g);
      tmem_ld_32x32(taddr + HALF + c * 32, u);
      tmem_ld_wait();
      const int f0 = n_tile * HALF + c * 32;
      if (m_ok && f0 < p.F) chunk2(p, g, u, m, f0, epi_smem);
    }
  }
};

// ---- cross-attention K/V projection for all decoder layers at once (X1):
// row m = (b, s) of the encoder output; column n = ((layer*2 + kv)*H + h)*64 + d.
// Written straight into the decode arena  [layer][kv][B][H][S][64].
struct EpiCrossKV {
  struct Params {
    act_t* arena;
    int B, H, S;
    const int* row_b = nullptr;  // packed encoder rows: row m is position row_s[m] of prompt row_b[m]
    const int* row_s = nullptr;
  };
  static constexpr bool kPaired = false;
  typedef NoPre ChunkPre;
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  static DEVINL void chunk_pre(const Params&, int, int, int, ChunkPre&) {}
  static DEVINL void chunk(const Params& p, const uint32_t (&acc)[32], int m, int n0, int N, const uint8_t*,
                           const ChunkPre&) {
    const int b = p.row_b ? p.row_b[m] : m / p.S;
    const int s = p.row_s ? p.row_s[m] : m - b * p.S;
    const int hd = p.H * 64;
    const int lkv = n0 / hd;
    const int rem = n0 - lkv * hd;
    const int h = rem >> 6, d0 = rem & 63;
    uint32_t o[16];
    round_pack_32(acc, o);
    act_t* dst = p.arena + ((((static_cast<size_t>(lkv) * p.B + b) * p.H + h) * p.S + s) << 6) + d0;
    store_row_chunk(dst, o, n0, N);
  }
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t* es,
                         int part, int parts) {
    run_chunks_from_tmem<BN, EpiCrossKV>(p, taddr, m, m_ok, n_tile, N, es, part, parts);
  }
};

// ---- decoder self-attention QKV for one new token: q -> q buffer, k/v appended in
// place at row *step of the preallocated cache [kv][B][H][Tmax][64] (replaces the
// torch.cat regrowth of transformers/cache_utils.py:119-120).
struct EpiQkvDecode {
  struct Params {
    act_t* q;      // [B, I]
    act_t* cache;  // this layer: [2][B][H][Tmax][64]
    const int* step;       // device scalar: current decode position t
    int B, H, Tmax;
    int step_stride = 0;   // 1 = slot pool: row m appends at its own position step[m]
  };
  static constexpr bool kPaired = false;
  typedef NoPre ChunkPre;
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  static DEVINL void chunk_pre(const Params&, int, int, int, ChunkPre&) {}
  static DEVINL void chunk(const Params& p, const uint32_t (&acc)[32], int m, int n0, int N, const uint8_t*,
                           const ChunkPre&) {
    const int I = p.H * 64;
    uint32_t o[16];
    round_pack_32(acc, o);
    act_t* dst;
    if (n0 < I) {
      dst = p.q + static_cast<size_t>(m) * I + n0;
    } else {
      const int t = p.step[m * p.step_stride];
      const int r = n0 - I;
      const int kv = r / I;
      const int rem = r - kv * I;
      const int h = rem >> 6, d0 = rem & 63;
      dst = p.cache + ((((static_cast<size_t>(kv) * p.B + m) * p.H + h) * p.Tmax + t) << 6) + d0;
    }
    store_row_chunk(dst, o, n0, N);
  }
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t* es,
                         int part, int parts) {
    run_chunks_from_tmem<BN, EpiQkvDecode>(p, taddr, m, m_ok, n_tile, N, es, part, parts);
  }
};

// ---- lm_head + greedy arg-max: logits never reach HBM. Each (row, n_tile) emits the
// max bf16-rounded logit of its BN columns and the lowest column index attaining it
// (torch.argmax first-index contract; generation/utils.py:2762,2793). EOS is masked
// to -inf while step < min_new_tokens (generation/logits_process.py:225-233).
struct EpiArgmax {
  struct Params {
    float* pval;  // [M, n_tiles]
    int* pidx;    // [M, n_tiles]
    int n_tiles;
    const int* step;
    int eos, min_new;
    int step_stride = 0;  // 1 = slot pool: per-row positions
  };
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t*, int, int) {
    float best = -INFINITY;
    int bidx = n_tile * BN;  // all -inf (cannot happen with finite logits) -> first column, like torch
    const bool block_eos = m_ok && p.step[m * p.step_stride] < p.min_new;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(taddr + c * 32, acc);
      tmem_ld_wait();
      const int n0 = n_tile * BN + c * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + j;
        float v = act_round(__uint_as_float(acc[j]));
        if (n >= N || (block_eos && n == p.eos)) v = -INFINITY;
        if (v > best) {  // ascending scan + strict '>' keeps the lowest index among equal maxima
          best = v;
          bidx = n;
        }
      }
    }
    if (m_ok) {
      p.pval[static_cast<size_t>(m) * p.n_tiles + n_tile] = best;
      p.pidx[static_cast<size_t>(m) * p.n_tiles + n_tile] = bidx;
    }
  }
};

// ---- fp32 logits store (test hook / teacher-forced parity): C = float(bf16(acc))
struct EpiStoreF32 {
  struct Params {
    float* C;
    int ldc;
  };
  static DEVINL void prologue(const Params&, uint8_t*, int, int = 128) {}
  template <int BN>
  static DEVINL void run(const Params& p, uint32_t taddr, int m, bool m_ok, int n_tile, int N, const uint8_t*, int, int) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(taddr + c * 32, acc);
      tmem_ld_wait();
      const int n0 = n_tile * BN + c * 32;
      if (m_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < N) p.C[static_cast<size_t>(m) * p.ldc + n0 + j] = act_round(__uint_as_float(acc[j]));
      }
    }
  }
};

// ======================================================================== host launch
// Opt in to the large dynamic shared memory carve-out once per process/device
// (done at b200t5_create so it never happens inside a stream capture).
template <int BN, class Epi>
cudaError_t prepare_gemm() {
  return cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              GemmCfg<BN>::kSmemBytes);
}

// Scheduling priority attached to every kernel launched through launch_kernel / launch_gemm_splitk
// (0 = default). The decode step raises it for the short latency-bound kernels so that, when
// row-chains run concurrently, their CTAs are placed ahead of the queued CTAs of another chain's
// HBM-streaming cross-attention kernel instead of behind them.
inline int& launch_priority() {
  static thread_local int prio = 0;
  return prio;
}

// Launch with (optionally) the programmatic-stream-serialization attribute (PDL).
template <class... KArgs, class... Args>
cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                          Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
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
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int BN, class Epi>
cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int m_fastest,
                        const typename Epi::Params& ep, int num_sms, cudaStream_t stream, bool pdl = false) {
  using Cfg = GemmCfg<BN>;
  const int tiles = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_kernel(gemm_bf16_tn_kernel<BN, Epi>, dim3(grid), dim3(EpiWarps<BN>::kThreads), Cfg::kSmemBytes, stream, pdl,
                       tmA, tmB, M, N, K, m_fastest, ep);
}

}  // namespace b200
