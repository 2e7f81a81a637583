// Cross-attention of the decode step (D7) as a bulk-copy stream: the HBM roofline kernel, second design.
//
// attention_decode.cuh reads K/V with per-thread 16-byte loads, so its bandwidth is proportional to the warps
// resident per SM (it needs ~28 of them for 42 GB/s per SM). That is fine when it has the GPU to itself, but the
// decode step overlaps it with the latency-bound split-K GEMMs of the other row-chain, whose CTAs (29 K registers,
// ~130 KB of shared memory each) evict more than half of the attention CTAs from the SMs they land on: measured in
// round 1, a 64-row launch took 25 us instead of 16 under that contention and the step gained nothing from the
// overlap (profiles/decode_trace_r1.md).
//
// Here the bytes in flight do not depend on how many warps are resident: one producer lane per CTA streams the
// K and V slabs of the CTA's (row, head) items through a ring of 8 KB shared-memory stages with cp.async.bulk
// (complete_tx on an mbarrier, L2 evict-first), and four consumer warps do the arithmetic out of shared memory.
// Two CTAs per SM x `stages` x 8 KB are in flight whatever else is resident, the CTAs are persistent (grid sized
// so that every CTA gets the same number of items), and their footprint (2 x ~45 KB, 2 x 160 threads x <= 64
// registers) leaves room for a split-K GEMM CTA of the other chain on the same SM.
//
// The arithmetic, the rounding points AND the order of every fp32 accumulation are those of
// attn_decode_kernel<false>: within each 128-key block warp w / lane group ks / unroll slot u owns key
// 4w + ks + 16u exactly as there, so the two kernels return bit-identical results (tests/test_kernels_gpu.py) and
// the model-level parity evidence carries over unchanged.
#pragma once
#include "attention_decode.cuh"

namespace b200 {

constexpr int kXsConsumerWarps = 4;
constexpr int kXsThreads = (kXsConsumerWarps + 1) * 32;  // + the producer warp
constexpr int kXsChunkKeys = 64;                         // 8 KB of K or V rows per ring stage
constexpr int kXsChunkBytes = kXsChunkKeys * 128;
constexpr int kXsStageBytes = kXsChunkBytes + 128;       // + the chunk's 64 key_ok bytes (K phase), 128-byte aligned
constexpr int kXsMaxStages = 12;

struct XsSmem {
  // [stages][8 KB + 128 B] ring | scores [Tk] | red [4][64] | stat [8] | full[stages] empty[stages]
  static __host__ __device__ size_t bytes(int stages, int Tk) {
    return static_cast<size_t>(stages) * kXsStageBytes + static_cast<size_t>((Tk + 3) & ~3) * 4 + 4 * 64 * 4 + 8 * 4 +
           2 * kXsMaxStages * 8 + 128 /* alignment slack */;
  }
};

DEVINL void bulk_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
DEVINL void xs_bar_sync() { asm volatile("bar.sync 1, %0;" ::"r"(kXsConsumerWarps * 32) : "memory"); }  // the consumer warps only

// <= 88 registers: two of these CTAs (28 K registers) and one split-K GEMM CTA (192 x 160) share an SM's 64 K
__global__ void __maxnreg__(88)
attn_cross_stream_kernel(const act_t* __restrict__ q,    // [B, H*64]
                         const act_t* __restrict__ Kc,   // [B][H][Tk][64]
                         const act_t* __restrict__ Vc,   // [B][H][Tk][64]
                         act_t* __restrict__ ctx,        // [B, H*64]
                         int n_items,                    // B * H
                         int H, int Tk,
                         const int* __restrict__ extent,            // [B] keys to visit (0 = retired row)
                         const unsigned char* __restrict__ key_ok,  // [B][Tk] 1 = attended
                         int stages, int late_pdl, XsStamps stamps) {
  extern __shared__ uint8_t xs_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xs_raw) + 127) & ~uintptr_t(127));
  float* s_scores = reinterpret_cast<float*>(ring + static_cast<size_t>(stages) * kXsStageBytes);
  float* s_red = s_scores + ((Tk + 3) & ~3);  // [4][64]
  float* s_stat = s_red + 4 * 64;             // [8]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_stat + 8);
  uint64_t* empty = full + kXsMaxStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kXsConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  // late_pdl = 0: the dependent kernel (this chain's cross-attention output projection) may start its prologue at
  // once, as everywhere else in the step. late_pdl = 1: its CTAs would only sit on their SMs (~130 KB of shared
  // memory each) while this kernel streams, in the way of the OTHER chain's GEMMs: release them when this CTA
  // starts its LAST item, so that the prologue overlaps the tail of the stream only.
  if (!late_pdl) pdl_launch_dependents();
  // everything above overlapped the previous kernel's tail; q / extent are that kernel's (or the previous step's) output
  pdl_wait();
  unsigned long long t_start = 0;
  if (stamps.slots != nullptr && threadIdx.x == 0) t_start = global_timer_ns();
  const int first = blockIdx.x, stride = gridDim.x;
  // The key_ok bytes of a K chunk travel with it (a second, 64-byte bulk copy on the same barrier) when the rows
  // are 16-byte aligned; otherwise the consumers read them from global memory. They must not be fetched with
  // ordinary loads on the consumers' critical path: with four consumer warps per CTA there is nothing to hide an
  // L2 round trip per chunk behind (first version of this kernel: 11 us per item instead of 6, 3.3 TB/s).
  const bool mask_bulk = (Tk & 15) == 0;
  const int last_item = first + ((n_items - 1 - first) / stride) * stride;

  if (warp == kXsConsumerWarps) {
    // ------------------------------------------------------------ producer: K chunks then V chunks of every item
    if (lane == 0) {
      const uint64_t policy = l2_policy_evict_first();
      int stage = 0;
      uint32_t phase = 0;
      int n_next = first < n_items ? extent[first / H] : 0;
      for (int it = first; it < n_items; it += stride) {
        const int b = it / H;
        const int n = n_next;
        // (the next item's extent is fetched a whole item ahead: the ring holds < 2 us of stream, an L2 round trip
        // under load is of that order)
        if (it + stride < n_items) n_next = extent[(it + stride) / H];
        const size_t slab = static_cast<size_t>(it) * Tk * 64;
#pragma unroll 1
        for (int kv = 0; kv < 2; ++kv) {
          const act_t* src = (kv ? Vc : Kc) + slab;
          for (int k0 = 0; k0 < n; k0 += kXsChunkKeys) {
            const int keys = n - k0 < kXsChunkKeys ? n - k0 : kXsChunkKeys;
            const uint32_t mbytes = (kv == 0 && mask_bulk) ? static_cast<uint32_t>((keys + 15) & ~15) : 0u;
            uint8_t* dst = ring + static_cast<size_t>(stage) * kXsStageBytes;
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(keys) * 128u + mbytes);
            bulk_load_1d_hint(dst, src + static_cast<size_t>(k0) * 64, static_cast<uint32_t>(keys) * 128u, &full[stage], policy);
            if (mbytes) bulk_load_1d(dst + kXsChunkBytes, key_ok + static_cast<size_t>(b) * Tk + k0, mbytes, &full[stage]);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    return;  // (the bulk copies complete on barriers the consumer warps wait on: the CTA outlives them)
  }

  // -------------------------------------------------------------- consumers (4 warps = attn_decode_kernel's CTA)
  const int ks = lane >> 3, dg = lane & 7;
  int stage = 0;
  uint32_t phase = 0;
  // software pipeline over items: the next item's query and extent are fetched while this one streams
  float qf[8];
  int n = 0;
  auto load_q = [&](int it, float (&dst)[8], int& nn) {
    const uint4 qv = *reinterpret_cast<const uint4*>(q + static_cast<size_t>(it) * 64 + dg * 8);
    dst[0] = act_lo(qv.x); dst[1] = act_hi(qv.x); dst[2] = act_lo(qv.y); dst[3] = act_hi(qv.y);
    dst[4] = act_lo(qv.z); dst[5] = act_hi(qv.z); dst[6] = act_lo(qv.w); dst[7] = act_hi(qv.w);
    nn = extent[it / H];
  };
  if (first < n_items) load_q(first, qf, n);
  for (int it = first; it < n_items; it += stride) {
    const int b = it / H;
    const int nkeys = n;
    float qn[8];
    int nn = 0;
    const int nxt = it + stride;
    if (nxt < n_items) load_q(nxt, qn, nn);
    if (late_pdl && it == last_item) pdl_launch_dependents();
    const unsigned char* ok_row = key_ok + static_cast<size_t>(b) * Tk;

    // ---------------- phase 1: scores (keys 4*warp + ks + 16u of every 128-key block, u = 0..7, as in attn_decode_kernel)
    for (int k0 = 0; k0 < nkeys; k0 += kXsChunkKeys) {
      // this lane group's four keys of the chunk: chunk-local key 4*warp + ks + 16*uu
      unsigned char okv[4] = {0, 0, 0, 0};
      if (!mask_bulk && dg == 0) {
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int j = k0 + warp * 4 + ks + 16 * uu;
          okv[uu] = j < nkeys ? ok_row[j] : 0;
        }
      }
      mbar_wait(&full[stage], phase);
      const uint8_t* base = ring + static_cast<size_t>(stage) * kXsStageBytes + dg * 16;
      if (mask_bulk && dg == 0) {
        const uint8_t* okb = ring + static_cast<size_t>(stage) * kXsStageBytes + kXsChunkBytes;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) okv[uu] = okb[warp * 4 + ks + 16 * uu];
      }
      uint4 kv[4];
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int jl = warp * 4 + ks + 16 * uu;
        kv[uu] = k0 + jl < nkeys ? *reinterpret_cast<const uint4*>(base + jl * 128) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int j = k0 + warp * 4 + ks + 16 * uu;
        float s = dot8(kv[uu], qf);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (dg == 0 && j < nkeys) s_scores[j] = okv[uu] ? act_round(s) : kActMin;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    xs_bar_sync();

    // ---------------- softmax statistics over the rounded scores (fp32, exact two-pass; same order as attn_decode_kernel)
    const int tid = threadIdx.x;  // 0..127
    float mx = -INFINITY;
    for (int j = tid; j < nkeys; j += kXsConsumerWarps * 32) mx = fmaxf(mx, s_scores[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_stat[warp] = mx;
    xs_bar_sync();
    mx = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
    float sum = 0.f;
    for (int j = tid; j < nkeys; j += kXsConsumerWarps * 32) {
      const float e = expf(s_scores[j] - mx);
      s_scores[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_stat[4 + warp] = sum;
    xs_bar_sync();
    sum = (s_stat[4] + s_stat[5]) + (s_stat[6] + s_stat[7]);
    for (int j = tid; j < nkeys; j += kXsConsumerWarps * 32) s_scores[j] = act_round(s_scores[j] / sum);
    xs_bar_sync();

    // ---------------- phase 2: out = P . V
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int k0 = 0; k0 < nkeys; k0 += kXsChunkKeys) {
      mbar_wait(&full[stage], phase);
      const uint8_t* base = ring + static_cast<size_t>(stage) * kXsStageBytes + dg * 16;
      uint4 vv[4];
      float p[4];
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int jl = warp * 4 + ks + 16 * uu;
        const bool ok = k0 + jl < nkeys;
        vv[uu] = ok ? *reinterpret_cast<const uint4*>(base + jl * 128) : make_uint4(0, 0, 0, 0);
        p[uu] = ok ? s_scores[k0 + jl] : 0.f;
      }
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        acc[0] = fmaf(p[uu], act_lo(vv[uu].x), acc[0]);
        acc[1] = fmaf(p[uu], act_hi(vv[uu].x), acc[1]);
        acc[2] = fmaf(p[uu], act_lo(vv[uu].y), acc[2]);
        acc[3] = fmaf(p[uu], act_hi(vv[uu].y), acc[3]);
        acc[4] = fmaf(p[uu], act_lo(vv[uu].z), acc[4]);
        acc[5] = fmaf(p[uu], act_hi(vv[uu].z), acc[5]);
        acc[6] = fmaf(p[uu], act_lo(vv[uu].w), acc[6]);
        acc[7] = fmaf(p[uu], act_hi(vv[uu].w), acc[7]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (ks == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s_red[warp * 64 + dg * 8 + e] = acc[e];
    }
    xs_bar_sync();
    if (tid < 32) {
      const int d0 = tid * 2;
      const float o0 = (s_red[d0] + s_red[64 + d0]) + (s_red[128 + d0] + s_red[192 + d0]);
      const float o1 = (s_red[d0 + 1] + s_red[64 + d0 + 1]) + (s_red[128 + d0 + 1] + s_red[192 + d0 + 1]);
      *reinterpret_cast<uint32_t*>(ctx + static_cast<size_t>(it) * 64 + d0) = pack_act2(o0, o1);
    }
    // (a warp may run ahead into the next item's phase 1 and overwrite s_scores while warp 0 still reads s_red:
    // different arrays; every warp's phase-2 reads of s_scores completed before the barrier above, and s_red is
    // next written four barriers later)
#pragma unroll
    for (int e = 0; e < 8; ++e) qf[e] = qn[e];
    n = nn;
  }
  if (late_pdl && first >= n_items) pdl_launch_dependents();
  if (stamps.slots != nullptr && threadIdx.x == 0) {
    atomicMin(&stamps.slots[2 * stamps.slot], t_start);
    atomicMax(&stamps.slots[2 * stamps.slot + 1], static_cast<unsigned long long>(global_timer_ns()));
  }
}

// grid: every CTA gets the same number of items (rounds = ceil(items / (2 * SMs)), CTAs = ceil(items / rounds))
inline int xs_grid(int n_items, int num_sms) {
  const int max_cta = 2 * num_sms;
  if (n_items <= max_cta) return n_items;
  const int rounds = (n_items + max_cta - 1) / max_cta;
  return (n_items + rounds - 1) / rounds;
}

}  // namespace b200
