// Single-query cross-attention of the decode step on the tensor cores (tcgen05 + TMEM + TMA), S <= 512.
//
// attn_decode_kernel<false> (attention_decode.cuh) reads K and V with ~27 CUDA-core instructions per 16
// bytes and needs ~32 resident warps per SM to keep HBM busy. Here the CUDA cores only do the 512-element
// softmax; everything that touches K/V bytes is TMA + tcgen05:
//   scores  S_c[128 keys x 16] = K_chunk[128 x 64] . q^T[64 x 16]      (M = 128 keys, the query padded to N = 16:
//                                                                         column 0 is the real one)
//   output  O[128 x 64 d]     += P_t[128 x 64 keys] . V_t[64 keys x 64 d]       (only ROW 0 of P - hence of O - is real:
//                                                                         the A tile's other 127 rows alias whatever
//                                                                         follows the 128-byte probability row in
//                                                                         shared memory; their products land in
//                                                                         accumulator rows nobody reads)
// K and V chunks (128 keys x 128 B = 16 KB, contiguous in the arena) are streamed by one producer thread through a
// ring of kXtcStages TMA tiles in exactly the swizzled layout the MMA consumes; a persistent CTA per SM works
// through its (row, head) items with the scores of item i+1 computed while item i is in its softmax.
// Rounding contract = attn_decode_kernel<false>: s = bf16(q.k) (mask -> finfo.min, keys >= extent excluded),
// p = bf16(exp(s - max) / sum) in fp32, out = bf16(sum_j p_j v_j) accumulated in fp32.
//
// Warps: 0 = TMA producer, 1 = TMEM + MMA issuer, 2..5 = softmax / epilogue (thread <-> key lane of a chunk).
// TMEM: S[2] (64 columns each: 4 chunks x 16) and O[2] (64 columns each), double-buffered across items. All
// probabilities of an item are staged at once - eight 128-byte rows at a 1 KB pitch, each the first row of a
// (mostly aliased) 128 x 64 A tile - so an item needs one hand-over per stage. Measured per-MMA cost is what
// shapes this: ~90 clk for an M = 64 / N = 16 product regardless of its size, so the P.V product is issued as
// 32 MMAs of 128 x 64 x 16 (32 clk each) rather than as small ones.
// The K/V arena must hold finite values everywhere (it is zero-filled at allocation): keys beyond a prompt's
// extent are multiplied by p = 0.
#pragma once
#include "attention_decode.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kXtcKStages = 4;   // K chunks are consumed by the score MMAs as soon as they land
constexpr int kXtcVStages = 7;   // V chunks wait for the item's softmax: their ring also has to hold the NEXT item's loads
constexpr int kXtcStages = kXtcKStages + kXtcVStages;
constexpr int kXtcChunkKeys = 128;
constexpr int kXtcChunkBytes = kXtcChunkKeys * 128;  // 16 KB
constexpr int kXtcThreads = 192;
constexpr int kXtcMaxS = 512;
constexpr int kXtcQTileBytes = 16 * 128;             // B operand of the score MMA: 16 rows x 64 bf16
constexpr int kXtcPTileBytes = 8 * 1024;             // 8 probability rows at a 1 KB pitch; the 16 KB an A tile spans runs on
                                                     // into the next buffer / the rings (garbage rows, never read back)
constexpr int kXtcSmemBytes = 2 * kXtcPTileBytes + kXtcStages * kXtcChunkBytes + 2 * kXtcQTileBytes + 1024 /*align*/ + 512;

DEVINL void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr) : "memory");
}
DEVINL void xtc_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// Items bh in [bh0, bh0 + nitems): q, ctx, extent/key_ok rows are indexed with the ABSOLUTE bh / b;
// tmK/tmV: [rows, 64] bf16 maps (box 64 x 128, 128-B swizzle) whose row k_row0 + bh * Tk + j is key j of item bh.
__global__ void __launch_bounds__(kXtcThreads, 1)
attn_decode_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, int k_row0,
                      int v_row0, const act_t* __restrict__ q, act_t* __restrict__ ctx, int bh0, int nitems,
                      int H, int Tk, const int* __restrict__ extent, const unsigned char* __restrict__ key_ok,
                      long long* prof = nullptr) {  // diagnostic: SM-clock stamps of CTA 0's softmax thread 0, 8 per item
  extern __shared__ uint8_t xtc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xtc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ptile = smem;                                   // [2][8 KB], followed by >= 16 KB of other shared memory
  uint8_t* kring = smem + 2 * kXtcPTileBytes;              // [kXtcKStages][16 KB]
  uint8_t* vring = kring + kXtcKStages * kXtcChunkBytes;   // [kXtcVStages][16 KB]
  uint8_t* qtile = vring + kXtcVStages * kXtcChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(qtile + 2 * kXtcQTileBytes);
  uint64_t* full = bars;                  // [stages]: K ring first, then V ring
  uint64_t* empty = full + kXtcStages;    // [stages]
  uint64_t* q_ready = empty + kXtcStages; // [2] the query row of the item is staged
  uint64_t* s_full = q_ready + 2;         // [2] scores complete in TMEM
  uint64_t* s_free = s_full + 2;          // [2] scores read out
  uint64_t* p_full = s_free + 2;          // [2] the item's probabilities are staged
  uint64_t* p_free = p_full + 2;          // [2] ... and have been consumed by its P.V MMAs
  uint64_t* o_full = p_free + 2;          // [2] output complete in TMEM
  uint64_t* o_free = o_full + 2;          // [2] output read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 2);  // [8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kXtcStages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_ready[i], 8);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 128);
        mbar_init(&p_full[i], 128);
        mbar_init(&p_free[i], 1);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_free[i], 128);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  } else if (warp >= 2) {
    // zero the operand tiles once: only row 0 of each is ever rewritten
    uint4* z = reinterpret_cast<uint4*>(qtile);
    for (int i = threadIdx.x - 64; i < 2 * kXtcQTileBytes / 16; i += 128) z[i] = make_uint4(0, 0, 0, 0);
    z = reinterpret_cast<uint4*>(ptile);
    for (int i = threadIdx.x - 64; i < 2 * kXtcPTileBytes / 16; i += 128) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int my_items = blockIdx.x < static_cast<unsigned>(nitems) ? (nitems - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto item_bh = [&](int i) { return bh0 + static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x); };
  auto item_chunks = [&](int i) { return (extent[item_bh(i) / H] + kXtcChunkKeys - 1) / kXtcChunkKeys; };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Two independent rings so that neither stream blocks the other: K chunks (consumed as soon as they land) and
    // V chunks (resident until the item's softmax is done). Whichever has a free stage gets the next request.
    if (lane == 0 && my_items > 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      int ki = 0, kc = 0, vi = 0, vc = 0;  // next K / V chunk to request: (item, chunk)
      int knc = item_chunks(0), vnc = knc;
      while (ki < my_items || vi < my_items) {
        bool progressed = false;
        if (ki < my_items && mbar_try_wait(&empty[ks], kph ^ 1u)) {
          mbar_arrive_expect_tx(&full[ks], kXtcChunkBytes);
          tma_load_2d(kring + ks * kXtcChunkBytes, &tmK, &full[ks], 0, k_row0 + item_bh(ki) * Tk + kc * kXtcChunkKeys);
          if (++ks == kXtcKStages) {
            ks = 0;
            kph ^= 1u;
          }
          if (++kc == knc) {
            kc = 0;
            if (++ki < my_items) knc = item_chunks(ki);
          }
          progressed = true;
        }
        // V of item i is never requested before K of item i (its scores come first)
        if (vi < my_items && (vi < ki || ki >= my_items) && mbar_try_wait(&empty[kXtcKStages + vs], vph ^ 1u)) {
          mbar_arrive_expect_tx(&full[kXtcKStages + vs], kXtcChunkBytes);
          tma_load_2d(vring + vs * kXtcChunkBytes, &tmV, &full[kXtcKStages + vs], 0, v_row0 + item_bh(vi) * Tk + vc * kXtcChunkKeys);
          if (++vs == kXtcVStages) {
            vs = 0;
            vph ^= 1u;
          }
          if (++vc == vnc) {
            vc = 0;
            if (++vi < my_items) vnc = item_chunks(vi);
          }
          progressed = true;
        }
        if (!progressed) __nanosleep(32);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && my_items > 0) {
      constexpr uint32_t idesc_s = make_idesc_act(128, 16, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_act(128, 64, 0, 1);  // B = V is MN-major (d contiguous)
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      auto scores = [&](int i) {  // S[i&1][:, 16c .. 16c+16) = K_c . q^T
        const int ib = i & 1;
        const uint32_t par = static_cast<uint32_t>(i >> 1) & 1u;
        const int nc = item_chunks(i);
        mbar_wait(&q_ready[ib], par);
        mbar_wait(&s_free[ib], par ^ 1u);
        tc_fence_after_sync();
        const uint64_t dq = make_desc_sw128_kmajor(smem_u32(qtile + ib * kXtcQTileBytes));
        for (int c = 0; c < nc; ++c) {
          mbar_wait(&full[ks], kph);
          tc_fence_after_sync();
          const uint64_t dk = make_desc_sw128_kmajor(smem_u32(kring + ks * kXtcChunkBytes));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ss(tmem_base + ib * 64 + c * 16, dk + 2 * kk, dq + 2 * kk, idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(&empty[ks]);
          if (++ks == kXtcKStages) {
            ks = 0;
            kph ^= 1u;
          }
        }
        umma_commit(&s_full[ib]);
      };
      scores(0);
      for (int i = 0; i < my_items; ++i) {
        if (i + 1 < my_items) scores(i + 1);
        const int ib = i & 1;
        const uint32_t par = static_cast<uint32_t>(i >> 1) & 1u;
        const int nc = item_chunks(i);
        mbar_wait(&o_free[ib], par ^ 1u);
        mbar_wait(&p_full[ib], par);
        tc_fence_after_sync();
        for (int c = 0; c < nc; ++c) {
          mbar_wait(&full[kXtcKStages + vs], vph);
          tc_fence_after_sync();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA
            const uint64_t dv = make_desc_sw128_mnmajor(smem_u32(vring + vs * kXtcChunkBytes + kk * 2048), 1024, 1024);
            const uint64_t dp = make_desc_sw128_kmajor(smem_u32(ptile + ib * kXtcPTileBytes + (c * 2 + (kk >> 2)) * 1024)) + 2 * (kk & 3);
            umma_f16_ss(tmem_base + 128 + ib * 64, dp, dv, idesc_o, (c | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[kXtcKStages + vs]);
          if (++vs == kXtcVStages) {
            vs = 0;
            vph ^= 1u;
          }
        }
        umma_commit(&p_free[ib]);
        umma_commit(&o_full[ib]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / epilogue (128 threads)
    const int tid = threadIdx.x - 64;
    const int qd = warp & 3;            // TMEM lane quarter this warp may read
    const int j = qd * 32 + lane;       // key lane inside a chunk
    const int w4 = warp - 2;            // 0..3: slot in s_stat
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    pdl_wait();
    // the query row of item k (8 x 16 B) goes to row 0 of q tile k&1 (row 0 of the swizzled tile is unswizzled).
    // It is staged one item AHEAD: the MMA warp computes the scores of item i+1 while item i is in its softmax.
    // The score MMAs that last read tile k&1 belong to item k-2; they completed before s_full of that item,
    // which this warp group has waited for by the time it stages item k.
    // Nothing an item needs from global memory may sit on the per-item critical path (each L2 round trip is
    // ~0.5 us against a ~2 us item): the query row, the extent and the key mask of the NEXT item are fetched into
    // registers while the current item is being processed.
    auto load_q = [&](int k) -> uint4 {
      return (tid < 8 && k < my_items) ? *reinterpret_cast<const uint4*>(q + static_cast<size_t>(item_bh(k)) * 64 + tid * 8)
                                       : make_uint4(0, 0, 0, 0);
    };
    auto stage_q = [&](int k, const uint4& qv) {
      if (tid < 8 && k < my_items) {
        *reinterpret_cast<uint4*>(qtile + (k & 1) * kXtcQTileBytes + tid * 16) = qv;
        fence_proxy_async_smem();
        mbar_arrive(&q_ready[k & 1]);
      }
    };
    auto load_meta = [&](int k, int& nk, unsigned& okbits) {
      nk = 0;
      okbits = 0;
      if (k < my_items) {
        const int b = item_bh(k) / H;
        nk = extent[b];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jg = c * kXtcChunkKeys + j;
          if (jg < Tk && key_ok[static_cast<size_t>(b) * Tk + jg]) okbits |= 1u << c;
        }
      }
    };
    // output of item k: row 0 (TMEM lane 0) of O[k&1], 64 columns
    auto epilogue = [&](int k) {
      const int kb = k & 1;
      mbar_wait(&o_full[kb], static_cast<uint32_t>(k >> 1) & 1u);
      tc_fence_after_sync();
      if (qd == 0) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(tmem_base + 128 + kb * 64, o0);
        tmem_ld_32x32(tmem_base + 128 + kb * 64 + 32, o1);
        tmem_ld_wait();
        if (lane == 0) {
          uint4* dst = reinterpret_cast<uint4*>(ctx + static_cast<size_t>(item_bh(k)) * 64);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[g] = make_uint4(pack_act2(__uint_as_float(o0[8 * g]), __uint_as_float(o0[8 * g + 1])),
                                pack_act2(__uint_as_float(o0[8 * g + 2]), __uint_as_float(o0[8 * g + 3])),
                                pack_act2(__uint_as_float(o0[8 * g + 4]), __uint_as_float(o0[8 * g + 5])),
                                pack_act2(__uint_as_float(o0[8 * g + 6]), __uint_as_float(o0[8 * g + 7])));
#pragma unroll
          for (int g = 0; g < 4; ++g)
            dst[4 + g] = make_uint4(pack_act2(__uint_as_float(o1[8 * g]), __uint_as_float(o1[8 * g + 1])),
                                    pack_act2(__uint_as_float(o1[8 * g + 2]), __uint_as_float(o1[8 * g + 3])),
                                    pack_act2(__uint_as_float(o1[8 * g + 4]), __uint_as_float(o1[8 * g + 5])),
                                    pack_act2(__uint_as_float(o1[8 * g + 6]), __uint_as_float(o1[8 * g + 7])));
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&o_free[kb]);
    };
    stage_q(0, load_q(0));
    uint4 q_next = load_q(1);
    int nk_next;
    unsigned ok_next;
    load_meta(0, nk_next, ok_next);
    for (int i = 0; i < my_items; ++i) {
      const int ib = i & 1;
      const uint32_t par = static_cast<uint32_t>(i >> 1) & 1u;
      const int nk = nk_next;
      const unsigned okbits = ok_next;
      const int nc = (nk + kXtcChunkKeys - 1) / kXtcChunkKeys;
      const uint4 q_stage = q_next;
      q_next = load_q(i + 2);              // in flight during this item
      load_meta(i + 1, nk_next, ok_next);  // likewise
      const bool pr = prof != nullptr && blockIdx.x == 0 && tid == 0 && i < 32;
      if (pr) prof[i * 8 + 0] = clock64();
      // ---- scores of this thread's key lane, one per chunk
      mbar_wait(&s_full[ib], par);
      if (pr) prof[i * 8 + 1] = clock64();
      tc_fence_after_sync();
      uint32_t raw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < nc) tmem_ld_32x1(tlane + ib * 64 + c * 16, raw[c]);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&s_free[ib]);
      stage_q(i + 1, q_stage);
      if (pr) prof[i * 8 + 2] = clock64();
      float sc[4];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jg = c * kXtcChunkKeys + j;
        float v = -INFINITY;
        if (c < nc && jg < nk) {
          v = act_round(__uint_as_float(raw[c]));
          if (!((okbits >> c) & 1u)) v = kActMin;
        }
        sc[c] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) s_stat[w4] = mx;
      xtc_bar(128);
      mx = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sc[c] = sc[c] == -INFINITY ? 0.f : expf(sc[c] - mx);
        sum += sc[c];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) s_stat[4 + w4] = sum;
      xtc_bar(128);
      sum = (s_stat[4] + s_stat[5]) + (s_stat[6] + s_stat[7]);
      if (pr) prof[i * 8 + 3] = clock64();
      // ---- probabilities of the whole item: key j of chunk c -> probability row 2c + (j >> 6), element j & 63
      mbar_wait(&p_free[ib], par ^ 1u);  // the P.V MMAs of item i-2 have consumed this buffer
      if (pr) prof[i * 8 + 4] = clock64();
      for (int c = 0; c < nc; ++c)
        *reinterpret_cast<act_t*>(ptile + ib * kXtcPTileBytes + (c * 2 + (j >> 6)) * 1024 + (j & 63) * 2) =
            float2act(sc[c] / sum);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[ib]);
      if (pr) prof[i * 8 + 5] = clock64();
      // ---- the previous item's output is read while this item's P.V runs
      if (i > 0) epilogue(i - 1);
      if (pr) prof[i * 8 + 6] = clock64();
    }
    if (my_items > 0) epilogue(my_items - 1);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace b200
