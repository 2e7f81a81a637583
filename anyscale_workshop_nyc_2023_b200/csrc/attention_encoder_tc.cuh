// Encoder self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), S <= 512.
//
// T5Attention.forward (modeling_t5.py:253-344) for one (batch row b, head h), 128 queries at a time:
//   scores = bf16(Q K^T)                      tcgen05.mma, fp32 accumulators in TMEM (128 x 512)
//   scores = bf16(scores + bias[h, j-i])      bias = relative-position buckets + key-padding mask
//   p      = bf16(softmax_fp32(scores))       exact two-pass softmax (row max, row sum, then p)
//   out    = bf16(p V)                        tcgen05.mma, A = p staged in smem, B = V (MN-major)
// The score tile never leaves the SM: S lives in TMEM, is rewritten in place with the rounded,
// biased scores, and is read back twice more (sum, then p) - HF's exact rounding contract
//  without materialising [B,H,S,S] in HBM.
//
// Warp roles (32 + 512 threads): warp 0 = TMEM owner + TMA + MMA issue (one elected thread);
// warps 1..16 = softmax/epilogue: warp W owns TMEM lanes 32*(W%4).. (= query rows) and, of every
// 128-key chunk, the 32 keys selected by (W-1)/4 (four warps per scheduler hide the MUFU/TMEM latency). P is staged per 128-key chunk through a
// double-buffered 32 KB smem tile in the canonical 128-B-swizzled K-major layout (the same
// layout TMA writes), so the P.V MMA of chunk c overlaps the exp/normalise work of chunk c+1.
//
// Shared memory: Q 16 KB | K 4 x 16 KB | V 4 x 16 KB | P 2 x 32 KB (216 KB with the tables: one CTA per SM, as the
// 512 TMEM columns of the S tile dictate anyway).
// TMEM: 512 columns = the S tile; O (64 columns) aliases S[:, 0:64) after chunk 0 was consumed.
#pragma once
#include "attention_decode.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kEncTcParts = 4;                        // column parts per 128-key chunk (warps per lane quarter)
constexpr int kEncTcCompute = 128 * kEncTcParts;     // softmax/epilogue threads
constexpr int kEncTcThreads = 32 + kEncTcCompute;
constexpr int kEncTcQ = 128;      // queries per CTA
constexpr int kEncTcChunk = 128;  // keys per chunk
constexpr int kEncTcMaxS = 512;

DEVINL void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// exp(t) for t <= 0 as ex2.approx(t * log2(e)) with the product carried in two pieces (hi by FMA, lo by a multiply):
// 4 instructions per score where expf costs ~10, max error ~2-3 ulp like expf's documented 2 ulp (the naive
// ex2(t * log2e) would lose |t| * 2^-24 relative accuracy in the single rounding of the product). Results below
// 2^-126 flush to zero; such a probability is < 1e-38 of the row maximum's and vanishes in the fp32 P.V sum anyway.
DEVINL float exp_fast(float t) {
  const float y = fmaf(t, 1.4426950216293335f, t * 1.925963033500258e-8f);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return r;
}
DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct EncTcSmem {
  static constexpr int kQ = 0;
  static constexpr int kK = 16384;               // 4 x 16 KB
  static constexpr int kV = kK + 65536;          // 4 x 16 KB
  static constexpr int kP = kV + 65536;          // 2 x 32 KB: P chunks staged for the P.V MMA
  static constexpr int kBars = kP + 65536;       // mbarriers + tmem slot + flags
  static constexpr int kStat = kBars + 128;      // float[2][parts][128]: max, sum per column part
  // two packed bf16x2 bias tables, T0[k] = (bias[2k], bias[2k+1]) and T1[k] = (bias[2k+1], bias[2k+2]),
  // (S + 128) / 2 + 16 words each (the +16 staggers T1 by half the banks)
  static constexpr int kBias = kStat + 2 * kEncTcParts * 128 * 4;
  __host__ __device__ static int table_words(int S) { return (S + 128) / 2 + 16; }
  // the same rounded up to whole 16-byte vectors (layout of the precomputed tables and of sT0 / sT1 in shared memory)
  __host__ __device__ static int table_words_padded(int S) { return (table_words(S) + 3) & ~3; }
  static size_t bytes(int S) { return static_cast<size_t>(kBias) + 2 * table_words_padded(S) * 4 + S + 16 + 1024; }
};

// One CTA per (batch row, head), PERSISTENT over that row's 128-query tiles (round 2). K and V (up to 128 KB) are
// loaded once instead of once per tile, the barriers / TMEM allocation / key_ok scan happen once, the next tile's Q
// is requested as soon as the current Q K^T has completed and its bias tables are copied in while the current tile is
// in pass C; P has its own staging buffers (K's shared memory is no longer recycled), so nothing of a tile's
// prologue is left on the critical path of the following tiles. Round 1 ran one CTA per tile: ~3.2 k clocks of
// prologue + ~2.9 k of load latency in front of ~13 k of work, 83 times per SM and layer (profiles/encoder_ncu_r1.md).
__global__ void __launch_bounds__(kEncTcThreads, 1)
encoder_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,  // [B*S, 3I] bf16, box 64 x 128
                       act_t* __restrict__ ctx,            // [B*S, I]
                       const float* __restrict__ rel_bias,         // [H][2S-1], index j - i + S - 1
                       const unsigned char* __restrict__ key_ok,   // [B][S]
                       const int* __restrict__ extent,             // [B]
                       const int* __restrict__ cu,                 // packed rows: prompt b starts at row cu[b] (NULL: b * S)
                       int S, int H,
                       long long* __restrict__ prof = nullptr,     // diagnostic: SM-clock stamps of CTA 0's first tile, B200T5_ENC_PROF
                       const uint32_t* __restrict__ packed_bias = nullptr) {  // [H][q tiles][2][table_words_padded(S)]: sT0|sT1 ready-made
  extern __shared__ uint8_t enc_tc_raw[];
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh - b * H;
  const int ext = extent[b];
  // a row that is not part of this pass (slot-pool admission) or has nothing to attend: nothing downstream reads it
  if (ext <= 0) return;
  const bool pr0 = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 32;
  if (pr0) prof[0] = clock64();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(enc_tc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + EncTcSmem::kQ;
  uint8_t* sK = smem + EncTcSmem::kK;
  uint8_t* sV = smem + EncTcSmem::kV;
  uint8_t* sP = smem + EncTcSmem::kP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + EncTcSmem::kBars);
  uint64_t* bar_kv = bars;         // K and V of this (b, h) landed (once)
  uint64_t* bar_q = bars + 1;      // Q tile landed (per tile)
  uint64_t* bar_s = bars + 2;      // S = Q K^T complete (per tile)
  uint64_t* bar_pfull = bars + 3;  // [2] P chunk staged (512 arrivals)
  uint64_t* bar_pfree = bars + 5;  // [2] P.V MMA of that buffer complete
  uint64_t* bar_o = bars + 7;      // all P.V of the tile complete
  uint64_t* bar_tfree = bars + 8;  // the tile's O has been read out of TMEM: the next Q K^T may overwrite S (512 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  int* sHoles = reinterpret_cast<int*>(bars + 11);
  float* sStat = reinterpret_cast<float*>(smem + EncTcSmem::kStat);
  uint32_t* sT0 = reinterpret_cast<uint32_t*>(smem + EncTcSmem::kBias);
  const int tw = EncTcSmem::table_words_padded(S);
  uint32_t* sT1 = sT0 + tw;
  unsigned char* sOk = reinterpret_cast<unsigned char*>(sT1 + tw);
  if (threadIdx.x == 0) *sHoles = 0;
  __syncthreads();

  const int I = H * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (ext + kEncTcChunk - 1) / kEncTcChunk;   // key chunks
  const int ntiles = (ext + kEncTcQ - 1) / kEncTcQ;            // query tiles with at least one real row
  const int ntiles_all = (S + kEncTcQ - 1) / kEncTcQ;          // (layout of packed_bias)
  const int row0 = cu ? cu[b] : b * S;

  // bias tables of query tile ti -> sT0 / sT1 (x = j - i_local + 127  <->  rel index j - i + S - 1)
  auto load_tables = [&](int ti) {
    const int t = threadIdx.x - 32;
    if (packed_bias != nullptr) {
      // built once per plan on the host (b200t5.cu): a 16-byte-vector copy instead of ~1k dependent global reads
      const uint4* src = reinterpret_cast<const uint4*>(packed_bias + (static_cast<size_t>(h) * ntiles_all + ti) * 2 * tw);
      uint4* d0 = reinterpret_cast<uint4*>(sT0);
      for (int k = t; k < 2 * tw / 4; k += kEncTcCompute) d0[k] = src[k];  // sT1 = sT0 + tw follows contiguously
    } else {
      const int lo = S - 1 - ti * kEncTcQ - 127;
      // the bias values are bf16 embedding entries widened to fp32: packing them back is lossless
      auto bias_at = [&](int x) -> float {
        const int idx = lo + x;
        return (x < S + 127 && idx >= 0 && idx < 2 * S - 1) ? rel_bias[static_cast<size_t>(h) * (2 * S - 1) + idx] : 0.f;
      };
      for (int k = t; k < (S + 128) / 2; k += kEncTcCompute) {
        const float b0 = bias_at(2 * k), b1 = bias_at(2 * k + 1), b2 = bias_at(2 * k + 2);
        sT0[k] = pack_act2(b0, b1);
        sT1[k] = pack_act2(b1, b2);
      }
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      mbar_init(bar_kv, 1);
      mbar_init(bar_q, 1);
      mbar_init(bar_s, 1);
      mbar_init(&bar_pfull[0], kEncTcCompute);
      mbar_init(&bar_pfull[1], kEncTcCompute);
      mbar_init(&bar_pfree[0], 1);
      mbar_init(&bar_pfree[1], 1);
      mbar_init(bar_o, 1);
      mbar_init(bar_tfree, kEncTcCompute);
      mbar_fence_init();
      // first Q tile, then K and V of this (b, h): issued before anything else so that the TMA latency runs under the
      // bias-table copy and the TMEM allocation (the issuing thread initialised and fenced the barriers itself)
      mbar_arrive_expect_tx(bar_q, 16384u);
      tma_load_2d(sQ, &tmQKV, bar_q, h * 64, row0);
      mbar_arrive_expect_tx(bar_kv, 16384u * 2 * nchunks);
      for (int c = 0; c < nchunks; ++c) {
        tma_load_2d(sK + c * 16384, &tmQKV, bar_kv, I + h * 64, row0 + c * kEncTcChunk);
        tma_load_2d(sV + c * 16384, &tmQKV, bar_kv, 2 * I + h * 64, row0 + c * kEncTcChunk);
      }
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  } else {
    load_tables(0);
    const int t = threadIdx.x - 32;
    int holes = 0;
    for (int x = t; x < S; x += kEncTcCompute) {
      const unsigned char ok = key_ok[static_cast<size_t>(b) * S + x];
      sOk[x] = ok;
      holes |= (x < ext && !ok) ? 1 : 0;
    }
    if (holes) *sHoles = 1;  // rare: a non-prefix attention mask
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (pr0) prof[1] = clock64();  // prologue done (bias tables, barriers, TMEM)

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_act(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_act(128, 64, 0, 1);  // B = V is MN-major (d contiguous)
      const uint64_t dq = make_desc_sw128_kmajor(smem_u32(sQ));
      mbar_wait(bar_kv, 0);
      int g = 0;  // running P-chunk counter: buffer g & 1, barrier phase (g >> 1) & 1
      for (int ti = 0; ti < ntiles; ++ti) {
        mbar_wait(bar_q, ti & 1);
        if (ti > 0) mbar_wait(bar_tfree, (ti - 1) & 1);  // the previous tile's O has left TMEM
        tc_fence_after_sync();
        // ---------------- S[:, 128c : 128c+128] = Q K_c^T
        for (int c = 0; c < nchunks; ++c) {
          const uint64_t dk = make_desc_sw128_kmajor(smem_u32(sK + c * 16384));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ss(tmem_base + c * 128, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
        if (ti + 1 < ntiles) {
          // the next Q tile: its shared memory is free once this tile's Q K^T has completed
          mbar_wait(bar_s, ti & 1);
          mbar_arrive_expect_tx(bar_q, 16384u);
          tma_load_2d(sQ, &tmQKV, bar_q, h * 64, row0 + (ti + 1) * kEncTcQ);
        }
        // ---------------- O += P_c V_c as the softmax warps hand chunks over
        for (int c = 0; c < nchunks; ++c, ++g) {
          const int buf = g & 1;
          mbar_wait(&bar_pfull[buf], (g >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t pbase = smem_u32(sP + buf * 32768);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t dp = make_desc_sw128_kmajor(pbase + (kk >> 2) * 16384) + 2 * (kk & 3);
            const uint64_t dv = make_desc_sw128_mnmajor(smem_u32(sV + c * 16384 + kk * 2048), 1024, 1024);
            umma_f16_ss(tmem_base, dp, dv, idesc_o, (c | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_pfree[buf]);
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // ================================================================ softmax / epilogue warps
    const int q = warp & 3;
    const int part = (warp - 1) >> 2;  // which 32 keys of every 128-key chunk
    const int il = q * 32 + lane;      // local query row
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    constexpr int P = kEncTcParts;
    const bool holes = *sHoles != 0;
    const uint32_t* tab = ((127 - il) & 1) ? sT1 : sT0;
    int g = 0;
#pragma unroll 1
    for (int ti = 0; ti < ntiles; ++ti) {
      const int i0 = ti * kEncTcQ;
      const bool pr = pr0 && ti == 0;
      mbar_wait(bar_s, ti & 1);
      tc_fence_after_sync();
      if (pr) prof[2] = clock64();  // TMA loads + Q K^T done

      // ---- pass A: s = bf16(bf16(acc) + bias) (+ mask), two keys per instruction: the fp32 accumulators are
      // packed to bf16x2 (one F2FP), the bias pair comes from the packed table, add.rn.bf16x2 rounds
      // exactly like bf16(float(a) + float(b)) (the fp32 sum of two bf16 values is exact whenever it can
      // affect the bf16 rounding), max.bf16x2 keeps the running row max. The packed scores are written
      // back over the first 16 columns of the 32-column block they came from.
      act2_t mx2 = floats2act2(-INFINITY, -INFINITY);
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int jb = c * kEncTcChunk + part * 32;
        uint32_t v[32];
        tmem_ld_32x32(trow + jb, v);
        const int k0 = (jb - il + 127) >> 1;
        uint32_t pk[16];
        tmem_ld_wait();
        if (jb + 32 <= ext && !holes) {
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const uint32_t sb = pack_act2(__uint_as_float(v[2 * t]), __uint_as_float(v[2 * t + 1]));
            const uint32_t bb = tab[k0 + t];
            const act2_t r = __hadd2(*reinterpret_cast<const act2_t*>(&sb), *reinterpret_cast<const act2_t*>(&bb));
            mx2 = __hmax2(mx2, r);
            pk[t] = *reinterpret_cast<const uint32_t*>(&r);
          }
        } else {  // the chunk that contains the end of the row / a non-prefix mask: element by element
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            float r2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = jb + 2 * t + e;
              float sc = act_round(__uint_as_float(v[2 * t + e]));
              if (j < ext) {
                const uint32_t bb = tab[k0 + t];
                sc = act_round(sc + (e ? act_hi(bb) : act_lo(bb)));
                if (holes && !sOk[j]) sc = kActMin;
              } else {
                sc = -INFINITY;
              }
              r2[e] = sc;
            }
            const act2_t r = floats2act2(r2[0], r2[1]);  // exact: both are bf16 values
            mx2 = __hmax2(mx2, r);
            pk[t] = *reinterpret_cast<const uint32_t*>(&r);
          }
        }
        tmem_st_32x16(trow + jb, pk);
      }
      tmem_st_wait();
      float mx = fmaxf(__low2float(mx2), __high2float(mx2));
      sStat[part * 128 + il] = mx;
      named_bar_sync(1, kEncTcCompute);
      if (pr) prof[3] = clock64();  // pass A (bias, max) done
#pragma unroll
      for (int k = 0; k < P; ++k) mx = fmaxf(mx, sStat[k * 128 + il]);

      // ---- pass B: e = exp(s - max) kept in place as fp32; row sum
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int jb = c * kEncTcChunk + part * 32;
        uint32_t pk[16];
        tmem_ld_32x16(trow + jb, pk);
        tmem_ld_wait();
        uint32_t v[32];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float e0 = exp_fast(act_lo(pk[t]) - mx);
          const float e1 = exp_fast(act_hi(pk[t]) - mx);
          sum += e0;
          sum += e1;
          v[2 * t] = __float_as_uint(e0);
          v[2 * t + 1] = __float_as_uint(e1);
        }
        tmem_st_32x32(trow + jb, v);  // pass C only has to normalise
      }
      tmem_st_wait();
      sStat[(P + part) * 128 + il] = sum;
      named_bar_sync(1, kEncTcCompute);
      if (pr) prof[4] = clock64();  // pass B (exp, sum) done
      sum = 0.f;
#pragma unroll
      for (int k = 0; k < P; ++k) sum += sStat[(P + k) * 128 + il];  // fixed order: deterministic
      // the bias tables of this tile are dead (every warp is past pass A): bring in the next tile's under pass C
      if (ti + 1 < ntiles) load_tables(ti + 1);
      // p = exp(s - max) * (1 / sum): torch divides; the two differ by at most one fp32 ulp before the
      // rounding to bf16, i.e. in ~1e-5 of the elements by one bf16 ulp (far below the accumulation-order
      // noise between any two implementations), and the reciprocal removes a ~10-instruction IEEE divide
      // from the inner loop.
      const float inv_sum = 1.0f / sum;

      // ---- pass C: p = bf16(e * (1/sum)) staged per chunk for the P.V MMA
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c, ++g) {
        const int buf = g & 1;
        const int jb = c * kEncTcChunk + part * 32;
        uint32_t v[32];
        tmem_ld_32x32(trow + jb, v);
        tmem_ld_wait();
        uint32_t pk[16];  // 32 keys of this thread's row, packed bf16x2
#pragma unroll
        for (int t = 0; t < 16; ++t)
          pk[t] = pack_act2(__uint_as_float(v[2 * t]) * inv_sum, __uint_as_float(v[2 * t + 1]) * inv_sum);
        if (g >= 2) mbar_wait(&bar_pfree[buf], ((g - 2) >> 1) & 1);  // the P.V MMA that read this buffer two chunks ago
        // keys part*32.. of the chunk = sub-tile part/2 (64 keys each), 16-B groups (part&1)*4 .. +3 of the row
        uint8_t* tile = sP + buf * 32768 + (part >> 1) * 16384 + il * 128;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int gg = (part & 1) * 4 + gq;
          *reinterpret_cast<uint4*>(tile + ((gg ^ (il & 7)) << 4)) = make_uint4(pk[4 * gq], pk[4 * gq + 1], pk[4 * gq + 2], pk[4 * gq + 3]);
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&bar_pfull[buf]);
      }

      // ---- epilogue: O (TMEM cols 0..63) -> bf16 -> ctx[row0 + i, h*64 + d]; 16 columns per warp
      if (pr) prof[5] = clock64();  // pass C (normalise, stage P) done
      mbar_wait(bar_o, ti & 1);
      tc_fence_after_sync();
      if (pr) prof[6] = clock64();  // last P.V MMA done
      {
        uint32_t o[16];
        tmem_ld_32x16(trow + part * 16, o);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(bar_tfree);  // this thread's part of O is in registers: S may be overwritten by the next tile
        const int i = i0 + il;
        if (i < ext) {  // rows beyond the prompt's extent are padding (in the packed layout: another prompt's rows)
          uint32_t pkd[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) pkd[t] = pack_act2(__uint_as_float(o[2 * t]), __uint_as_float(o[2 * t + 1]));
          uint4* dst = reinterpret_cast<uint4*>(ctx + (static_cast<size_t>(row0) + i) * I + h * 64 + part * 16);
          dst[0] = make_uint4(pkd[0], pkd[1], pkd[2], pkd[3]);
          dst[1] = make_uint4(pkd[4], pkd[5], pkd[6], pkd[7]);
        }
      }
      if (pr) prof[7] = clock64();  // output written
      // the next tile's pass A reads the tables copied above and rewrites sStat
      named_bar_sync(1, kEncTcCompute);
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
