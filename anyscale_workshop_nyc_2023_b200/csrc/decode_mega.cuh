// The whole greedy decode loop as ONE persistent cooperative kernel.
//
// Why: a decode step of FLAN-T5-base is ~135 dependent kernels whose useful work is a few
// microseconds each (M = batch rows GEMMs, per-row norms, single-query attention). Launched one
// by one - even inside a CUDA graph with programmatic dependent launch - every kernel costs
// 5-10 us of launch, fill and drain, so the step takes ~1.6 ms while its HBM traffic needs 0.87 ms
// (profiles/decode_trace_r1.md). Here one CTA per SM stays resident for the whole generate()
// call; the ops of a step become PHASES separated by a grid-wide barrier (one atomic + one
// polled load, ~0.5 us), pipelines/barriers/TMEM are set up once, and nothing returns to the
// host until every row has finished (the EOS early exit is decided on the device).
//
// Phases of one decoder layer (reference: T5Block.forward, transformers modeling_t5.py:411-498):
//   G_qkv   xn  -> q, K/V appended to the self cache      tcgen05 tiles, direct epilogue
//   A_self  single-query attention over t+1 cached keys    one warp per (row, head)
//   G_o     ctx -> fp32 split-K partials                   tcgen05 tiles
//   E       x += bf16(sum partials); xn = RMSNorm(x)       one warp per row
//   G_cq    xn  -> q                                        direct epilogue
//   A_cross single-query attention over the encoder keys   4 warps per (row, head); HBM streaming
//   G_co, E, G_wi (GeGLU epilogue), G_ffo, E               as above
// then the lm_head tiles with the fused arg-max epilogue and the greedy bookkeeping (HF
// GenerationMixin._sample, generation/utils.py:2762-2805), which also gathers the next token's
// embedding and applies the first RMSNorm of the next step.
//
// GEMM phases reuse the warp roles of gemm.cuh inside the resident CTA: thread 0 = TMA producer,
// warp 1 lane 0 = tcgen05.mma issuer, warps 4..7 = epilogue (TMEM lane quarters); the smem ring,
// its mbarriers and the two TMEM accumulators persist across phases. The rounding contract is
// the shared chunk functors of gemm.cuh; split-K partial sums are added in a fixed order, so the
// output is run-to-run deterministic.
#pragma once
#include "attention_decode.cuh"
#include "attention_encoder.cuh"  // cp_async_commit / cp_async_wait
#include "elementwise.cuh"
#include "gemm.cuh"

namespace b200 {

constexpr int kMegaWarps = 12;                       // compute warps (attention, norms, GEMM MMA + epilogue roles)
constexpr int kMegaThreads = kMegaWarps * 32 + 32;   // + one control warp: TMA / bulk-copy producer and grid-barrier master
constexpr int kMegaMaster = kMegaWarps * 32;         // thread index of the control warp's lane 0
constexpr int kMegaStages = 3;
constexpr int kMegaBnMax = 128;
constexpr int kMegaStageBytes = kBM * kBK * 2 + kMegaBnMax * kBK * 2;  // 32 KB
constexpr int kMegaAccCols = 128;                                      // two accumulators -> 256 TMEM columns
constexpr int kMegaGroups = 3;                                         // cross-attention: 4-warp groups per CTA
constexpr int kXaSlots = 4;                                            // bulk-copy chunks in flight per group
constexpr int kXaChunkKeys = 64;
constexpr int kXaChunkBytes = kXaChunkKeys * 128;                      // 8 KB: 64 keys x 64 bf16
constexpr int kMegaXaRingBytes = kMegaGroups * kXaSlots * kXaChunkBytes;  // 96 KB
constexpr int kMegaScratchBytes = 24 * 1024;                           // scores / gelu table / norm partials (phases are disjoint)
constexpr int kMegaSmemBytes = kMegaStages * kMegaStageBytes + kMegaXaRingBytes + 1024 /*align*/ + 512 /*barriers*/ + kMegaScratchBytes;

struct MegaLayer {
  const CUtensorMap *tm_qkv, *tm_o, *tm_cq, *tm_co, *tm_wi, *tm_ffo;  // device-resident tensor maps
  const act_t *ln0, *ln1, *ln2;
  act_t* self_kv;         // [2][B][H][T][64]
  const act_t* cross_kv;  // [2][B][H][S][64]
  int wi_rows;                    // padded N of the interleaved wi weight
};

struct MegaParams {
  int B, S, T, d, I, F, H, V, Ld;
  float eps;
  act_t *dx, *dxn, *dq, *dctx, *dh;
  float* ws;  // [ks][B][d] fp32 split-K partials
  const CUtensorMap *tm_dxn, *tm_dctx, *tm_dh, *tm_lm;
  const MegaLayer* layers;
  const act_t *final_ln, *E;
  const int* extent;
  const unsigned char* key_ok;
  const float* dec_bias;
  DecodeState* st;
  int* unfinished;
  long long* out_ids;
  int* out_len;
  float* pval;
  int* pidx;
  int n_vtiles;
  long long eos, pad;
  int min_new, nsteps;
  GeluLut lut;
  unsigned int* bar;  // grid barrier counter, zeroed by the host before the launch
  long long* prof;    // optional: SM-clock stamps of CTA 0 at every phase boundary of step `prof_step`
  int prof_step;
  int bn_qkv, bn_proj, ks_proj, bn_cq, bn_wi, bn_ffo, ks_ffo, bn_lm;
};

enum MegaEpi { ME_PARTIAL = 0, ME_QKV = 1, ME_STORE = 2, ME_GEGLU = 3, ME_ARGMAX = 4 };

struct MegaShared {
  uint8_t* ring;     // GEMM operand ring (TMA tensor tiles)
  uint8_t* xa_ring;  // cross-attention K/V chunk ring (bulk copies), [group][slot][8 KB]
  uint64_t *full, *empty, *tfull, *tempty;
  uint64_t *xa_full, *xa_empty;  // [group][slot]
  uint32_t* tmem_slot;
  int* s_step;
  uint8_t* scratch;
};

// Pipeline positions that persist across phases (every thread carries a copy; each role uses its fields).
struct MegaPipe {
  int stage = 0;
  uint32_t phase = 0;
  int as = 0;
  uint32_t aphase = 0;
  int pref = 0;                       // producer: k-blocks of the next GEMM phase whose weight tile is already requested
  unsigned int xa_use = 0;            // cross-attention consumer: chunks consumed by this thread's group
  unsigned int xa_issue[kMegaGroups] = {0, 0, 0};  // producer: chunks issued per group
};

DEVINL unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (cooperative) grid: one release-add and a polled acquire-load by the control
// warp's lane 0 (SASS: RED.STRONG.GPU, LDG.STRONG.GPU + CCTL.IVALL - the acquire also drops this
// SM's stale L1 lines, so the other threads' plain loads after the bar.sync see the peers' writes).
// The same thread is the TMA producer: the proxy fence orders the peers' generic-proxy global writes, and this
// CTA's generic-proxy use of the shared-memory ring, before its next async-proxy (TMA) accesses.
DEVINL void grid_sync(unsigned int* bar, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == kMegaMaster) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int spins = 0;
    uint64_t t0 = 0;
    while (ld_acquire_u32(bar) < target) {
      if ((++spins & 0xFFFu) == 0) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();  // a protocol bug traps instead of hanging the GPU
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
}

// ---------------------------------------------------------------- GEMM phase
// D[M,N] = A[M,K] W[N,K]^T as 128 x bn tiles, K cut in `ksplit` slices; jobs round-robin over CTAs.
struct MegaGemm {
  const CUtensorMap *tmA, *tmB;
  int M, N, K, bn, ksplit, epi;
  // epilogue operands
  EpiQkvDecode::Params qkv;
  EpiStore::Params store;
  EpiGeglu::Params geglu;
  EpiArgmax::Params amax;
  float* ws;
  int ws_ld;  // row stride of ws (= N)
};

// Producer only: request the WEIGHT tiles of this CTA's first job of an upcoming GEMM phase. They never
// depend on the phases in between, so their HBM latency is hidden behind those phases and the barrier;
// the matching activation tiles are requested by mega_gemm_phase once the barrier has passed.
DEVINL void mega_gemm_prefetch_b(const MegaShared& sh, MegaPipe& ps, const MegaGemm& g) {
  if (threadIdx.x != kMegaMaster) return;
  const int tiles_m = (g.M + kBM - 1) / kBM;
  const int tiles_n = (g.N + g.bn - 1) / g.bn;
  const int kblocks = (g.K + kBK - 1) / kBK;
  const int kb_per = (kblocks + g.ksplit - 1) / g.ksplit;
  const int njobs = tiles_m * tiles_n * g.ksplit;
  ps.pref = 0;
  const int job = blockIdx.x;
  if (job >= njobs) return;
  const int ks = job % g.ksplit, t = job / g.ksplit;
  const int n_tile = t % tiles_n;
  const int kb0 = ks * kb_per;
  const int kb1 = kb0 + kb_per < kblocks ? kb0 + kb_per : kblocks;
  const int n = (kb1 - kb0) < kMegaStages ? (kb1 - kb0) : kMegaStages;
  const uint32_t stage_tx = static_cast<uint32_t>(kBM * kBK * 2 + g.bn * kBK * 2);
  int st = ps.stage;
  uint32_t ph = ps.phase;
  for (int i = 0; i < n; ++i) {
    mbar_wait(&sh.empty[st], ph ^ 1u);
    mbar_arrive_expect_tx(&sh.full[st], stage_tx);
    tma_load_2d(sh.ring + st * kMegaStageBytes + kBM * kBK * 2, g.tmB, &sh.full[st], (kb0 + i) * kBK, n_tile * g.bn);
    if (++st == kMegaStages) {
      st = 0;
      ph ^= 1u;
    }
  }
  ps.pref = n;
}

DEVINL void mega_gemm_phase(const MegaShared& sh, MegaPipe& ps, uint32_t tmem_base, const MegaGemm& g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (g.M + kBM - 1) / kBM;
  const int tiles_n = (g.N + g.bn - 1) / g.bn;
  const int kblocks = (g.K + kBK - 1) / kBK;
  const int kb_per = (kblocks + g.ksplit - 1) / g.ksplit;
  const int njobs = tiles_m * tiles_n * g.ksplit;
  const uint32_t stage_tx = static_cast<uint32_t>(kBM * kBK * 2 + g.bn * kBK * 2);

  if (threadIdx.x == kMegaMaster) {
    // ------------------------------------------------------------ TMA producer
    int pref = ps.pref;  // leading k-blocks of the first job whose weight tile was requested ahead of the barrier
    ps.pref = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int ks = job % g.ksplit, t = job / g.ksplit;
      const int n_tile = t % tiles_n, m_tile = t / tiles_n;
      const int kb0 = ks * kb_per;
      const int kb1 = kb0 + kb_per < kblocks ? kb0 + kb_per : kblocks;
      for (int kb = kb0; kb < kb1; ++kb) {
        uint8_t* sA = sh.ring + ps.stage * kMegaStageBytes;
        if (pref > 0) {
          --pref;  // slot reserved, transaction count armed and B in flight already
        } else {
          mbar_wait(&sh.empty[ps.stage], ps.phase ^ 1u);
          mbar_arrive_expect_tx(&sh.full[ps.stage], stage_tx);
          tma_load_2d(sA + kBM * kBK * 2, g.tmB, &sh.full[ps.stage], kb * kBK, n_tile * g.bn);
        }
        tma_load_2d(sA, g.tmA, &sh.full[ps.stage], kb * kBK, m_tile * kBM);
        if (++ps.stage == kMegaStages) {
          ps.stage = 0;
          ps.phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_act(kBM, g.bn, 0, 0);
      for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const int ks = job % g.ksplit;
        const int kb0 = ks * kb_per;
        const int kb1 = kb0 + kb_per < kblocks ? kb0 + kb_per : kblocks;
        mbar_wait(&sh.tempty[ps.as], ps.aphase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ps.as * kMegaAccCols);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sh.full[ps.stage], ps.phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sh.ring + ps.stage * kMegaStageBytes);
          const uint64_t a_desc = make_desc_sw128_kmajor(a_addr);
          const uint64_t b_desc = make_desc_sw128_kmajor(a_addr + kBM * kBK * 2);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16_ss(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&sh.empty[ps.stage]);
          if (++ps.stage == kMegaStages) {
            ps.stage = 0;
            ps.phase ^= 1u;
          }
        }
        umma_commit(&sh.tfull[ps.as]);
        ps.as ^= 1;
        if (ps.as == 0) ps.aphase ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------ epilogue warps (threads 128..255)
    const int q = warp & 3;
    if (g.epi == ME_GEGLU) EpiGeglu::prologue(g.geglu, sh.scratch, static_cast<int>(threadIdx.x) - 128, 128);
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int ks = job % g.ksplit, t = job / g.ksplit;
      const int n_tile = t % tiles_n, m_tile = t / tiles_n;
      const int m = m_tile * kBM + q * 32 + lane;
      const bool m_ok = m < g.M;
      mbar_wait(&sh.tfull[ps.as], ps.aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ps.as * kMegaAccCols);
      if (g.epi == ME_GEGLU) {
        const int half = g.bn >> 1;
        for (int c = 0; c < half / 32; ++c) {
          uint32_t gt[32], up[32];
          tmem_ld_32x32(taddr + c * 32, gt);
          tmem_ld_32x32(taddr + half + c * 32, up);
          tmem_ld_wait();
          const int f0 = n_tile * half + c * 32;
          if (m_ok && f0 < g.geglu.F) EpiGeglu::chunk2(g.geglu, gt, up, m, f0, sh.scratch);
        }
      } else if (g.epi == ME_ARGMAX) {
        float best = -INFINITY;
        int bidx = n_tile * g.bn;
        const bool block_eos = *g.amax.step < g.amax.min_new;
        for (int c = 0; c < g.bn / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          const int n0 = n_tile * g.bn + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            float v = act_round(__uint_as_float(acc[j]));
            if (n >= g.N || (block_eos && n == g.amax.eos)) v = -INFINITY;
            if (v > best) {
              best = v;
              bidx = n;
            }
          }
        }
        if (m_ok) {
          g.amax.pval[static_cast<size_t>(m) * g.amax.n_tiles + n_tile] = best;
          g.amax.pidx[static_cast<size_t>(m) * g.amax.n_tiles + n_tile] = bidx;
        }
      } else {
        for (int c = 0; c < g.bn / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          const int n0 = n_tile * g.bn + c * 32;
          if (m_ok && n0 < g.N) {
            if (g.epi == ME_PARTIAL) {
              float4* dst = reinterpret_cast<float4*>(g.ws + (static_cast<size_t>(ks) * g.M + m) * g.ws_ld + n0);
#pragma unroll
              for (int v = 0; v < 8; ++v)
                if (n0 + 4 * v + 4 <= g.N)
                  dst[v] = make_float4(__uint_as_float(acc[4 * v]), __uint_as_float(acc[4 * v + 1]),
                                       __uint_as_float(acc[4 * v + 2]), __uint_as_float(acc[4 * v + 3]));
            } else if (g.epi == ME_QKV) {
              EpiQkvDecode::chunk(g.qkv, acc, m, n0, g.N, nullptr, NoPre());
            } else {
              EpiStore::chunk(g.store, acc, m, n0, g.N, nullptr, NoPre());
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.tempty[ps.as]);
      ps.as ^= 1;
      if (ps.as == 0) ps.aphase ^= 1u;
    }
  }
}

// ---------------------------------------------------------------- E phase: residual + RMSNorm
// x[row] = bf16(x[row] + bf16(sum_ks ws[ks][row]))   (modeling_t5.py:375,406,149; skipped when ksplit == 0)
// xn[row] = T5LayerNorm(x[row]) * w                   (modeling_t5.py:55-68)
// The phase is pure latency (256 rows x 1.5 KB): a row is spread over W = ceil(d/256) warps so
// that every lane owns ONE 16-byte vector and all of its loads (x, the ksplit partials, w) are
// issued back to back - one L2 round trip instead of ksplit * d/256 dependent ones.

__device__ __noinline__ void mega_resnorm_phase(act_t* x, const float* ws, int ksplit, const act_t* w, act_t* xn,
                               int rows, int d, float eps, float* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;  // the control warp (12) falls outside every row group
  const int nvec = d >> 3;
  int W = (nvec + 31) >> 5;          // warps per row
  if (W > kMegaWarps) W = kMegaWarps;  // wider rows: lanes loop over vectors
  const int groups = kMegaWarps / W;
  const int grp = warp / W, wl = warp - grp * W;  // warps beyond groups*W idle
  const int gl = wl * 32 + lane;                  // lane index inside the row group
  float* red = scratch + grp * 16;                // per-group partial sums of squares
  const bool active = grp < groups;
  const int passes = (rows + gridDim.x * groups - 1) / (gridDim.x * groups);
  for (int ps = 0; ps < passes; ++ps) {
    const int row = (ps * gridDim.x + blockIdx.x) * groups + grp;
    const bool row_ok = active && row < rows;
    float ss = 0.f;
    if (row_ok) {
      for (int idx = gl; idx < nvec; idx += W * 32) {
        uint4* xr = reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * d) + idx;
        const uint4 v = *xr;
        uint32_t xs[4] = {v.x, v.y, v.z, v.w};
        if (ksplit > 0) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = 0.f;
          for (int k0 = 0; k0 < ksplit; k0 += 4) {  // 4 partials (8 loads) in flight per batch
            float4 pa[4], pb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (k0 + u < ksplit) {
                const float4* p4 = reinterpret_cast<const float4*>(ws + (static_cast<size_t>(k0 + u) * rows + row) * d + idx * 8);
                pa[u] = p4[0];
                pb[u] = p4[1];
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (k0 + u < ksplit) {
                y[0] += pa[u].x; y[1] += pa[u].y; y[2] += pa[u].z; y[3] += pa[u].w;
                y[4] += pb[u].x; y[5] += pb[u].y; y[6] += pb[u].z; y[7] += pb[u].w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            xs[j] = pack_act2(act_lo(xs[j]) + act_round(y[2 * j]), act_hi(xs[j]) + act_round(y[2 * j + 1]));
          *xr = make_uint4(xs[0], xs[1], xs[2], xs[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = act_lo(xs[j]), b = act_hi(xs[j]);
          ss = fmaf(a, a, ss);
          ss = fmaf(b, b, ss);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (active && lane == 0) red[wl] = ss;
    __syncthreads();  // uniform: every thread of the CTA runs the same number of passes
    if (row_ok) {
      float tot = 0.f;
      for (int i = 0; i < W; ++i) tot += red[i];  // fixed order
      const float inv = rsqrtf(tot * (1.0f / static_cast<float>(d)) + eps);
      for (int idx = gl; idx < nvec; idx += W * 32) {
        const uint4 v = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * d)[idx];
        const uint4 wv = reinterpret_cast<const uint4*>(w)[idx];
        const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
        const uint32_t wsv[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = act_round(act_lo(xs[j]) * inv);
          const float b = act_round(act_hi(xs[j]) * inv);
          o[j] = pack_act2(act_lo(wsv[j]) * a, act_hi(wsv[j]) * b);
        }
        reinterpret_cast<uint4*>(xn + static_cast<size_t>(row) * d)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();  // red[] is reused by the next pass
  }
}

// ---------------------------------------------------------------- self-attention phase
// One warp per (row, head): the shared per-item routine of attention_decode.cuh with plain (coherent) loads -
// the cache rows of this step were written by other CTAs of the same launch.
__device__ __noinline__ void mega_self_attn_phase(const act_t* q, const act_t* Kc, const act_t* Vc,
                                                  act_t* ctx, int BH, int H, int Tk, int t, const float* dist_bias,
                                                  float* scratch) {
  const int warp = threadIdx.x >> 5;
  if (warp >= kMegaWarps) return;  // control warp
  for (int bh = blockIdx.x * kMegaWarps + warp; bh < BH; bh += gridDim.x * kMegaWarps)
    self_attn_warp_item<false>(q, Kc, Vc, ctx, bh, H, Tk, t, dist_bias, scratch + warp * Tk);
}

// ---------------------------------------------------------------- cross-attention phase
// A group of 4 warps per (row, head), kMegaGroups groups per CTA; the arithmetic is
// attn_decode_kernel<false>'s (exact two-pass softmax, K and V each streamed once).
//
// The phase is the HBM roofline of the step, and with one resident CTA per SM the bytes in flight
// must come from depth, not from occupancy. The control warp streams every group's K and V slabs
// as 8 KB bulk copies (cp.async.bulk, 64 keys x 128 B, L2 evict-first) into a ring of kXaSlots
// chunks per group: 96 KB in flight per SM for ~one instruction per 8 KB, completion on per-slot
// mbarriers. The request stream runs ahead across the K -> softmax -> V -> next (row, head)
// boundaries (V and the next item's K do not depend on the scores), so the memory pipe never
// drains inside the phase; the compute warps only read shared memory.
DEVINL void group_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 4) : "memory"); }

DEVINL void bulk_load_1d_evict_first(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __noinline__ void mega_cross_attn_phase(const MegaShared& sh, MegaPipe& ps, const act_t* q,
                                                   const act_t* Kc, const act_t* Vc, act_t* ctx,
                                                   int BH, int H, int Tk, const int* extent, const unsigned char* key_ok) {
  const int stride = gridDim.x * kMegaGroups;
  if (threadIdx.x >= kMegaWarps * 32) {
    // ------------------------------------------------------------ bulk-copy producer (control warp, lane 0)
    if (threadIdx.x != kMegaMaster) return;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    int bh[kMegaGroups], e[kMegaGroups], nck[kMegaGroups], nkeys[kMegaGroups];
#pragma unroll
    for (int g = 0; g < kMegaGroups; ++g) {
      bh[g] = blockIdx.x * kMegaGroups + g;
      e[g] = 0;
      nkeys[g] = bh[g] < BH ? extent[bh[g] / H] : 0;
      nck[g] = (nkeys[g] + kXaChunkKeys - 1) / kXaChunkKeys;
    }
    bool any = true;
    while (any) {
      any = false;
#pragma unroll
      for (int g = 0; g < kMegaGroups; ++g) {
        if (bh[g] >= BH) continue;
        any = true;
        const bool isv = e[g] >= nck[g];
        const int c = isv ? e[g] - nck[g] : e[g];
        const int rows = nkeys[g] - c * kXaChunkKeys < kXaChunkKeys ? nkeys[g] - c * kXaChunkKeys : kXaChunkKeys;
        const unsigned int seq = ps.xa_issue[g]++;
        const int slot = seq % kXaSlots;
        uint64_t* full = &sh.xa_full[g * kXaSlots + slot];
        mbar_wait(&sh.xa_empty[g * kXaSlots + slot], ((seq / kXaSlots) & 1u) ^ 1u);
        mbar_arrive_expect_tx(full, static_cast<uint32_t>(rows) * 128u);
        const act_t* src = (isv ? Vc : Kc) + (static_cast<size_t>(bh[g]) * Tk + static_cast<size_t>(c) * kXaChunkKeys) * 64;
        bulk_load_1d_evict_first(sh.xa_ring + (g * kXaSlots + slot) * kXaChunkBytes, src, static_cast<uint32_t>(rows) * 128u, full, policy);
        if (++e[g] == 2 * nck[g]) {
          bh[g] += stride;
          e[g] = 0;
          nkeys[g] = bh[g] < BH ? extent[bh[g] / H] : 0;
          nck[g] = (nkeys[g] + kXaChunkKeys - 1) / kXaChunkKeys;
        }
      }
    }
    return;
  }
  // ------------------------------------------------------------ compute groups
  const int group = threadIdx.x >> 7;
  const int gt = threadIdx.x & 127;
  const int warp = gt >> 5, lane = gt & 31;
  const int ks = lane >> 3, dg = lane & 7;
  float* s_scores = reinterpret_cast<float*>(sh.scratch) + group * (Tk + 4 * 64 + 8);
  float* s_red = s_scores + Tk;    // [4][64]
  float* s_stat = s_red + 4 * 64;  // [8]
  const uint32_t ring0 = smem_u32(sh.xa_ring + group * kXaSlots * kXaChunkBytes) + (warp * 4 + ks) * 128 + dg * 16;
  uint64_t* full0 = sh.xa_full + group * kXaSlots;
  uint64_t* empty0 = sh.xa_empty + group * kXaSlots;
  for (int bh = blockIdx.x * kMegaGroups + group; bh < BH; bh += stride) {
    const int b = bh / H;
    const int nkeys = extent[b];
    const int nck = (nkeys + kXaChunkKeys - 1) / kXaChunkKeys;
    float qf[8];
    {
      const uint4 qv = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(bh) * 64 + dg * 8);
      qf[0] = act_lo(qv.x); qf[1] = act_hi(qv.x); qf[2] = act_lo(qv.y); qf[3] = act_hi(qv.y);
      qf[4] = act_lo(qv.z); qf[5] = act_hi(qv.z); qf[6] = act_lo(qv.w); qf[7] = act_hi(qv.w);
    }
    // ---- scores: chunk c holds keys [64c, 64c+64); this thread reads keys 64c + 16r + 4*warp + ks, r = 0..3
    for (int c = 0; c < nck; ++c) {
      const unsigned int seq = ps.xa_use++;
      const int slot = seq % kXaSlots;
      mbar_wait(&full0[slot], (seq / kXaSlots) & 1u);
      const uint32_t base = ring0 + slot * kXaChunkBytes;
      uint4 kv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = c * kXaChunkKeys + r * 16 + warp * 4 + ks;
        kv[r] = make_uint4(0, 0, 0, 0);
        if (j < nkeys)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(kv[r].x), "=r"(kv[r].y), "=r"(kv[r].z), "=r"(kv[r].w) : "r"(base + r * 16 * 128));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty0[slot]);  // this warp has copied its pieces out of the slot
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = c * kXaChunkKeys + r * 16 + warp * 4 + ks;
        float sc = dot8(kv[r], qf);
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);
        sc += __shfl_xor_sync(0xffffffffu, sc, 2);
        sc += __shfl_xor_sync(0xffffffffu, sc, 4);
        if (dg == 0 && j < nkeys) {
          sc = act_round(sc);
          if (!key_ok[static_cast<size_t>(b) * Tk + j]) sc = kActMin;
          s_scores[j] = sc;
        }
      }
    }
    group_sync(group);
    float mx = -INFINITY;
    for (int j = gt; j < nkeys; j += 128) mx = fmaxf(mx, s_scores[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_stat[warp] = mx;
    group_sync(group);
    mx = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
    float sum = 0.f;
    for (int j = gt; j < nkeys; j += 128) {
      const float ev = expf(s_scores[j] - mx);
      s_scores[j] = ev;
      sum += ev;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_stat[4 + warp] = sum;
    group_sync(group);
    sum = (s_stat[4] + s_stat[5]) + (s_stat[6] + s_stat[7]);
    for (int j = gt; j < nkeys; j += 128) s_scores[j] = act_round(s_scores[j] / sum);
    group_sync(group);
    // ---- out = P . V
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int c = 0; c < nck; ++c) {
      const unsigned int seq = ps.xa_use++;
      const int slot = seq % kXaSlots;
      mbar_wait(&full0[slot], (seq / kXaSlots) & 1u);
      const uint32_t base = ring0 + slot * kXaChunkBytes;
      uint4 vv[4];
      float pj[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = c * kXaChunkKeys + r * 16 + warp * 4 + ks;
        vv[r] = make_uint4(0, 0, 0, 0);
        pj[r] = 0.f;
        if (j < nkeys) {
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(vv[r].x), "=r"(vv[r].y), "=r"(vv[r].z), "=r"(vv[r].w) : "r"(base + r * 16 * 128));
          pj[r] = s_scores[j];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty0[slot]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[0] = fmaf(pj[r], act_lo(vv[r].x), acc[0]);
        acc[1] = fmaf(pj[r], act_hi(vv[r].x), acc[1]);
        acc[2] = fmaf(pj[r], act_lo(vv[r].y), acc[2]);
        acc[3] = fmaf(pj[r], act_hi(vv[r].y), acc[3]);
        acc[4] = fmaf(pj[r], act_lo(vv[r].z), acc[4]);
        acc[5] = fmaf(pj[r], act_hi(vv[r].z), acc[5]);
        acc[6] = fmaf(pj[r], act_lo(vv[r].w), acc[6]);
        acc[7] = fmaf(pj[r], act_hi(vv[r].w), acc[7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (ks == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_red[warp * 64 + dg * 8 + i] = acc[i];
    }
    group_sync(group);
    if (gt < 32) {
      const int d0 = gt * 2;
      const float o0 = (s_red[d0] + s_red[64 + d0]) + (s_red[128 + d0] + s_red[192 + d0]);
      const float o1 = (s_red[d0 + 1] + s_red[64 + d0 + 1]) + (s_red[128 + d0 + 1] + s_red[192 + d0 + 1]);
      *reinterpret_cast<uint32_t*>(ctx + static_cast<size_t>(bh) * 64 + d0) = pack_act2(o0, o1);
    }
    group_sync(group);  // s_scores / s_red are reused by the next item
  }
}

// ---------------------------------------------------------------- greedy bookkeeping phase
// One warp per row: reduce the per-tile arg-max partials (lowest index wins ties), apply HF's
// pad/EOS logic (generation/utils.py:2793-2805), gather the next token's embedding into x and
// apply the first RMSNorm of the next step.
__device__ __noinline__ void mega_finalize_phase(const MegaParams& P, int t) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= kMegaWarps) return;  // control warp
  const int d = P.d, nvec = d >> 3;
  for (int b = blockIdx.x * kMegaWarps + warp; b < P.B; b += gridDim.x * kMegaWarps) {
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int i = lane; i < P.n_vtiles; i += 32) {
      const float v = P.pval[static_cast<size_t>(b) * P.n_vtiles + i];
      const int ix = P.pidx[static_cast<size_t>(b) * P.n_vtiles + i];
      if (v > best || (v == best && ix < bidx)) {
        best = v;
        bidx = ix;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) {
        best = ov;
        bidx = oi;
      }
    }
    long long tok = 0;
    if (lane == 0) {
      const int unf = P.unfinished[b];
      tok = unf ? static_cast<long long>(bidx) : P.pad;
      P.out_ids[static_cast<size_t>(b) * (P.T + 1) + t + 1] = tok;
      if (unf) {
        P.out_len[b] = t + 1;
        if (tok == P.eos) {
          P.unfinished[b] = 0;
          atomicAdd(&P.st->finished_rows, 1);
        }
      }
    }
    tok = __shfl_sync(0xffffffffu, tok, 0);
    const uint4* src = reinterpret_cast<const uint4*>(P.E + static_cast<size_t>(tok) * d);
    uint4* dst = reinterpret_cast<uint4*>(P.dx + static_cast<size_t>(b) * d);
    for (int i = lane; i < nvec; i += 32) dst[i] = src[i];
    __syncwarp();
  }
}

// ================================================================== the kernel
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const __grid_constant__ MegaParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  MegaShared sh;
  sh.ring = smem;
  sh.xa_ring = smem + kMegaStages * kMegaStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh.xa_ring + kMegaXaRingBytes);
  sh.full = bars;
  sh.empty = bars + kMegaStages;
  sh.tfull = bars + 2 * kMegaStages;
  sh.tempty = sh.tfull + 2;
  sh.xa_full = sh.tempty + 2;
  sh.xa_empty = sh.xa_full + kMegaGroups * kXaSlots;
  sh.tmem_slot = reinterpret_cast<uint32_t*>(sh.xa_empty + kMegaGroups * kXaSlots);
  sh.s_step = reinterpret_cast<int*>(sh.tmem_slot + 1);
  sh.scratch = reinterpret_cast<uint8_t*>(bars) + 512;
  static_assert((2 * kMegaStages + 4 + 2 * kMegaGroups * kXaSlots) * 8 + 8 <= 512, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kMegaStages; ++i) {
        mbar_init(&sh.full[i], 1);
        mbar_init(&sh.empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sh.tfull[i], 1);
        mbar_init(&sh.tempty[i], 4);
      }
      for (int i = 0; i < kMegaGroups * kXaSlots; ++i) {
        mbar_init(&sh.xa_full[i], 1);
        mbar_init(&sh.xa_empty[i], 4);  // the group's four warps
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<2 * kMegaAccCols>(sh.tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *sh.tmem_slot;

  MegaPipe ps;
  unsigned int bar_target = 0;
  int prof_n = 0;
  int prof_t = -1;
#define GSYNC()                                                                                        \
  do {                                                                                                 \
    const bool _pr = P.prof != nullptr && prof_t == P.prof_step && blockIdx.x == 0 && threadIdx.x == kMegaMaster; \
    __syncthreads();                                                                                   \
    if (_pr) P.prof[prof_n++] = clock64(); /* this CTA's work of the phase is complete */            \
    grid_sync(P.bar, bar_target);                                                                      \
    if (_pr) P.prof[prof_n++] = clock64();                                                             \
  } while (0)
  const int B = P.B, d = P.d, I = P.I, F = P.F, H = P.H;
  float* fscratch = reinterpret_cast<float*>(sh.scratch);

  // GEMM phase descriptors
  auto g_qkv = [&](const MegaLayer& L) {
    MegaGemm g;
    g.tmA = P.tm_dxn; g.tmB = L.tm_qkv; g.M = B; g.N = 3 * I; g.K = d; g.bn = P.bn_qkv; g.ksplit = 1; g.epi = ME_QKV;
    g.qkv = EpiQkvDecode::Params{P.dq, L.self_kv, sh.s_step, B, H, P.T};
    return g;
  };
  auto g_part = [&](const CUtensorMap* tmA, const CUtensorMap* tmB, int K, int bn, int ksplit) {
    MegaGemm g;
    g.tmA = tmA; g.tmB = tmB; g.M = B; g.N = d; g.K = K; g.bn = bn; g.ksplit = ksplit; g.epi = ME_PARTIAL;
    g.ws = P.ws; g.ws_ld = d;
    return g;
  };
  auto g_cq = [&](const MegaLayer& L) {
    MegaGemm g;
    g.tmA = P.tm_dxn; g.tmB = L.tm_cq; g.M = B; g.N = I; g.K = d; g.bn = P.bn_cq; g.ksplit = 1; g.epi = ME_STORE;
    g.store = EpiStore::Params{P.dq, I};
    return g;
  };
  auto g_wi = [&](const MegaLayer& L) {
    MegaGemm g;
    g.tmA = P.tm_dxn; g.tmB = L.tm_wi; g.M = B; g.N = L.wi_rows; g.K = d; g.bn = P.bn_wi; g.ksplit = 1; g.epi = ME_GEGLU;
    g.geglu = EpiGeglu::Params{reinterpret_cast<ffh_t*>(P.dh), F, P.lut};  // (the fp16 build never launches this kernel)
    return g;
  };
  auto g_lm = [&]() {
    MegaGemm g;
    g.tmA = P.tm_dxn; g.tmB = P.tm_lm; g.M = B; g.N = P.V; g.K = d; g.bn = P.bn_lm; g.ksplit = 1; g.epi = ME_ARGMAX;
    g.amax = EpiArgmax::Params{P.pval, P.pidx, P.n_vtiles, sh.s_step, static_cast<int>(P.eos), P.min_new};
    return g;
  };

  // first RMSNorm of the first step (x = embedding of decoder_start, set by decode_init_kernel)
  mega_gemm_prefetch_b(sh, ps, g_qkv(P.layers[0]));
  mega_resnorm_phase(P.dx, nullptr, 0, P.layers[0].ln0, P.dxn, B, d, P.eps, fscratch);
  GSYNC();

  for (int t = 0; t < P.nsteps; ++t) {
    if (threadIdx.x == 0) *sh.s_step = t;
    prof_t = t;
    __syncthreads();
    for (int l = 0; l < P.Ld; ++l) {
      const MegaLayer& L = P.layers[l];
      const size_t self_plane = static_cast<size_t>(B) * I * P.T;
      const size_t cross_plane = static_cast<size_t>(B) * I * P.S;
      // ---- self-attention block
      mega_gemm_phase(sh, ps, tmem_base, g_qkv(L));
      mega_gemm_prefetch_b(sh, ps, g_part(P.tm_dctx, L.tm_o, I, P.bn_proj, P.ks_proj));
      GSYNC();
      mega_self_attn_phase(P.dq, L.self_kv, L.self_kv + self_plane, P.dctx, B * H, H, P.T, t, P.dec_bias, fscratch);
      GSYNC();
      mega_gemm_phase(sh, ps, tmem_base, g_part(P.tm_dctx, L.tm_o, I, P.bn_proj, P.ks_proj));
      mega_gemm_prefetch_b(sh, ps, g_cq(L));
      GSYNC();
      mega_resnorm_phase(P.dx, P.ws, P.ks_proj, L.ln1, P.dxn, B, d, P.eps, fscratch);
      GSYNC();
      // ---- cross-attention block
      mega_gemm_phase(sh, ps, tmem_base, g_cq(L));
      mega_gemm_prefetch_b(sh, ps, g_part(P.tm_dctx, L.tm_co, I, P.bn_proj, P.ks_proj));
      GSYNC();
      mega_cross_attn_phase(sh, ps, P.dq, L.cross_kv, L.cross_kv + cross_plane, P.dctx, B * H, H, P.S, P.extent, P.key_ok);
      GSYNC();
      mega_gemm_phase(sh, ps, tmem_base, g_part(P.tm_dctx, L.tm_co, I, P.bn_proj, P.ks_proj));
      mega_gemm_prefetch_b(sh, ps, g_wi(L));
      GSYNC();
      mega_resnorm_phase(P.dx, P.ws, P.ks_proj, L.ln2, P.dxn, B, d, P.eps, fscratch);
      GSYNC();
      // ---- feed-forward block
      mega_gemm_phase(sh, ps, tmem_base, g_wi(L));
      mega_gemm_prefetch_b(sh, ps, g_part(P.tm_dh, L.tm_ffo, F, P.bn_ffo, P.ks_ffo));
      GSYNC();
      mega_gemm_phase(sh, ps, tmem_base, g_part(P.tm_dh, L.tm_ffo, F, P.bn_ffo, P.ks_ffo));
      if (l + 1 < P.Ld) mega_gemm_prefetch_b(sh, ps, g_qkv(P.layers[l + 1]));
      else mega_gemm_prefetch_b(sh, ps, g_lm());
      GSYNC();
      mega_resnorm_phase(P.dx, P.ws, P.ks_ffo, l + 1 < P.Ld ? P.layers[l + 1].ln0 : P.final_ln, P.dxn, B, d, P.eps, fscratch);
      GSYNC();
    }
    // ---- lm_head + arg-max, then the greedy bookkeeping
    mega_gemm_phase(sh, ps, tmem_base, g_lm());
    mega_gemm_prefetch_b(sh, ps, g_qkv(P.layers[0]));  // next step (harmless after the last one: never consumed)
    GSYNC();
    mega_finalize_phase(P, t);
    GSYNC();
    // every row has emitted EOS: the remaining steps would only append pad tokens (already there)
    const int finished = *reinterpret_cast<volatile int*>(&P.st->finished_rows);
    if (blockIdx.x == 0 && threadIdx.x == 0) P.st->step = t + 1;
    if (finished >= B) break;
    mega_resnorm_phase(P.dx, nullptr, 0, P.layers[0].ln0, P.dxn, B, d, P.eps, fscratch);
    GSYNC();
  }
#undef GSYNC
  // a weight tile requested for a step that never ran must land before the shared memory is released
  if (threadIdx.x == kMegaMaster && ps.pref > 0) {
    // complete the armed transactions with the matching activation tiles (any valid tile) and wait for them
    const MegaGemm g = g_qkv(P.layers[0]);
    int st = ps.stage;
    uint32_t ph = ps.phase;
    for (int i = 0; i < ps.pref; ++i) {
      tma_load_2d(sh.ring + st * kMegaStageBytes, g.tmA, &sh.full[st], i * kBK, 0);
      mbar_wait(&sh.full[st], ph);
      if (++st == kMegaStages) {
        st = 0;
        ph ^= 1u;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<2 * kMegaAccCols>(tmem_base);
  }
}

}  // namespace b200
