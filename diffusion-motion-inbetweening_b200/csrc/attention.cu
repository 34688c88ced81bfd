// Self-attention core of nn.MultiheadAttention for the MDM encoder (reference: model/mdm.py:107-114, :284;
// no key-padding mask -- `src_key_padding_mask` is commented out there), head dim 128, S <= 208 keys:
//
//   O[s, :] = softmax_j( Q[s,:] . K[j,:] / sqrt(128) ) V[j, :]
//
// One CTA per (sequence, head, 128-query tile).  Both contractions run on tcgen05 tensor cores:
//   S = Q K^T   : A = Q tile (smem, K-major), B = K tile (smem, K-major), D = 128 x 208 fp32 in TMEM
//   O = P V     : A = P (bf16, written into TMEM by the softmax warps, aliasing S), B = V (smem, MN-major),
//                 D = 128 x 128 fp32 in TMEM
// Q, K and V are read straight out of the QKV projection's [tokens, 3*H*128] bf16 planes with TMA
// (no head-split or transpose kernels).  With nsplit = 3 every product uses the hi/lo bf16 split
// (X_lo*Y_hi + X_hi*Y_lo + X_hi*Y_hi), P included, so the result is fp32-accurate.
//
// warp 0: TMA producer   warp 1: MMA issuer (+TMEM alloc)
// warps 2..9: softmax + output.  Two threads per query row (TMEM lane group = warp % 4, key half = (warp-2)/4):
//   each keeps its ~104 logits in registers (one TMEM read), the halves exchange row max / row sum through
//   shared memory, P is written back to TMEM, and O leaves through coalesced 16-byte stores staged in the
//   (by then free) Q tile region.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace cmdi {

namespace {

constexpr int kHeadDim = 128;
constexpr int kQTile = 128;
constexpr int kKeyPad = kAttnKeyPad;            // 208 = 13 * 16
constexpr int kQBlockBytes = kQTile * 128;      // 64 dh-columns of 128 queries
constexpr int kKVBlockBytes = kKeyPad * 128;    // 64 dh-columns of 208 keys  (26 swizzle atoms)
constexpr int kQPlane = 2 * kQBlockBytes;       // 32768
constexpr int kKVPlane = 2 * kKVBlockBytes;     // 53248
constexpr int kOffQHi = 0, kOffQLo = kQPlane;
constexpr int kOffKHi = 2 * kQPlane, kOffKLo = kOffKHi + kKVPlane;
constexpr int kOffVHi = kOffKLo + kKVPlane;
constexpr int kOffVLo = kOffKHi;                // V_lo reuses the K region once S is complete
constexpr int kSmemTiles = kOffVHi + kKVPlane;  // 225280
constexpr int kNumSoftmaxWarps = 8;
constexpr int kThreads = 64 + kNumSoftmaxWarps * 32;
// TMEM columns
constexpr uint32_t kColS = 0, kColPHi = 0, kColPLo = 208, kColO = 320, kTmemCols = 512;
// key chunks (16 keys each) per half: half 0 -> chunks [0,7), half 1 -> chunks [7,13)
constexpr int kChunks0 = 7, kChunks1 = 6;

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// logits of this thread's key range -> registers, row max, exchange, probabilities -> TMEM, row sum
template <int CHUNK0, int NCHUNKS>
__device__ __forceinline__ float softmax_half(uint32_t trow, int S, bool split, float c_scale, float (*red_max)[128], int half,
                                              int row, bool trunc) {
  float s[NCHUNKS * 16];
  {
    // all TMEM reads in flight at once, one wait
    uint32_t v[NCHUNKS][16];
#pragma unroll
    for (int c = 0; c < NCHUNKS; ++c) tmem_ld16(trow + kColS + (CHUNK0 + c) * 16, v[c]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < NCHUNKS; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        s[c * 16 + j] = ((CHUNK0 + c) * 16 + j < S) ? __uint_as_float(v[c][j]) : -INFINITY;  // padded keys: exp2 -> 0
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NCHUNKS * 16; ++i) mx = fmaxf(mx, s[i]);
  red_max[half][row] = mx;
  // every S value is in registers now: after this barrier P may overwrite the S columns
  tc_fence_before();
  named_barrier_sync(1, kNumSoftmaxWarps * 32);
  tc_fence_after();
  mx = fmaxf(red_max[0][row], red_max[1][row]);
  const float mc = mx * c_scale;
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NCHUNKS; ++c) {
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = fast_exp2(fmaf(s[c * 16 + j * 2], c_scale, -mc));
      const float p1 = fast_exp2(fmaf(s[c * 16 + j * 2 + 1], c_scale, -mc));
      sum += p0 + p1;
      if (trunc) split_bf16x2_trunc(p0, p1, ph[j], pl[j]); else split_bf16x2(p0, p1, ph[j], pl[j]);
    }
    tmem_st8(trow + kColPHi + (CHUNK0 + c) * 8, ph);
    if (split) tmem_st8(trow + kColPLo + (CHUNK0 + c) * 8, pl);
  }
  return sum;
}

// ------------------------------------------------------------------------------------------------
// Persistent version: one CTA per SM walks over (sequence, head, query-tile) items w = blockIdx.x, +gridDim.x, ...
// With 225 KB of operand tiles only one CTA fits an SM, so the one-shot kernel above serialises load -> S ->
// softmax -> PV -> store per CTA and every wave starts with all 148 CTAs pulling 33 MB through L2 at once.
// Here the loads of item i+1 are issued as soon as the buffers of item i drain:
//   K region   : K_i            -> free at s_done_i  -> K_{i+1}
//   Q region   : Q_i, V_lo_i    -> free at o_done_i  -> Q_{i+1}        (V_lo_i enters at s_done_i)
//   V_hi region: V_hi_i, then the store staging of item i -> free at stage_free_i -> V_hi_{i+1}
// and S_{i+1} = Q K^T is issued while the softmax warps still write out O_i.  Every barrier completes exactly once
// per item, so the wait parity of item number i is i & 1.
// ------------------------------------------------------------------------------------------------
struct __align__(8) AttnBarriers2 {
  uint64_t q_full, k_full, vhi_full, vlo_full, s_done, p_full, o_done, stage_free;
  uint32_t tmem_base;
  uint32_t pad;
  float red_max[2][128];
  float red_sum[2][128];
};

__global__ void __launch_bounds__(kThreads, 1)
attention_persistent_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                            const __grid_constant__ CUtensorMap map_kv_hi, const __grid_constant__ CUtensorMap map_kv_lo,
                            const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                            const AttnParams p, const int num_items, const int q_tiles) {
  const long long t_entry = clock64();
  unsigned long long ns_entry;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_entry));
  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBarriers2* bars = reinterpret_cast<AttnBarriers2*>(smem + kSmemTiles);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.seq_len;
  const bool split = (p.nsplit == 3);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_q_hi);
    tma_prefetch_desc(&map_kv_hi);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->k_full, 1);
    mbar_init(&bars->vhi_full, 1);
    mbar_init(&bars->vlo_full, 1);
    mbar_init(&bars->s_done, 1);
    mbar_init(&bars->p_full, kNumSoftmaxWarps * 32);
    mbar_init(&bars->o_done, 1);
    mbar_init(&bars->stage_free, kNumSoftmaxWarps);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_wait();

  // item -> (sequence, head, query tile)
  auto decode = [&](int w, int& seq, int& head, int& qtile) {
    qtile = w % q_tiles;
    const int sh = w / q_tiles;
    head = sh % p.num_heads;
    seq = sh / p.num_heads;
  };

  if (warp_idx == 0) {
    if (lane == 0) {
      const uint32_t planes = split ? 2 : 1;
      auto load_q = [&](int w) {
        int seq, head, qtile;
        decode(w, seq, head, qtile);
        const int row = seq * S + qtile * kQTile, col = head * kHeadDim;
        mbar_arrive_expect_tx(&bars->q_full, planes * kQPlane);
        for (int j = 0; j < 2; ++j) {
          tma_load_2d(smem + kOffQHi + j * kQBlockBytes, &map_q_hi, &bars->q_full, col + j * 64, row);
          if (split) tma_load_2d(smem + kOffQLo + j * kQBlockBytes, &map_q_lo, &bars->q_full, col + j * 64, row);
        }
      };
      auto load_k = [&](int w) {
        int seq, head, qtile;
        decode(w, seq, head, qtile);
        const int col = p.num_heads * kHeadDim + head * kHeadDim;
        mbar_arrive_expect_tx(&bars->k_full, planes * kKVPlane);
        for (int j = 0; j < 2; ++j) {
          tma_load_2d(smem + kOffKHi + j * kKVBlockBytes, &map_kv_hi, &bars->k_full, col + j * 64, seq * S);
          if (split) tma_load_2d(smem + kOffKLo + j * kKVBlockBytes, &map_kv_lo, &bars->k_full, col + j * 64, seq * S);
        }
      };
      auto load_v = [&](int w, bool lo) {
        int seq, head, qtile;
        decode(w, seq, head, qtile);
        const int col = 2 * p.num_heads * kHeadDim + head * kHeadDim;
        uint64_t* bar = lo ? &bars->vlo_full : &bars->vhi_full;
        mbar_arrive_expect_tx(bar, kKVPlane);
        // V_lo lives in the Q region (free once S is complete; 53 KB of its 64 KB)
        uint8_t* dst = smem + (lo ? kOffQHi : kOffVHi);
        for (int j = 0; j < 2; ++j) tma_load_2d(dst + j * kKVBlockBytes, lo ? &map_kv_lo : &map_kv_hi, bar, col + j * 64, seq * S);
      };
      int w = blockIdx.x;
      if (w < num_items) {
        load_q(w);
        load_k(w);
        load_v(w, false);
      }
      for (uint32_t it = 0; w < num_items; w += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int wn = w + gridDim.x;
        mbar_wait(&bars->s_done, ph);      // Q_i and K_i consumed
        if (split) load_v(w, true);
        if (wn < num_items) load_k(wn);
        if (wn < num_items && p.prefetch_q) {
          // Q_{i+1} can only land once V_lo_i has left the Q region; pull it into L2 meanwhile
          int seq, head, qtile;
          decode(wn, seq, head, qtile);
          for (int j = 0; j < 2; ++j) {
            tma_prefetch_2d(&map_q_hi, head * kHeadDim + j * 64, seq * S + qtile * kQTile);
            if (split) tma_prefetch_2d(&map_q_lo, head * kHeadDim + j * 64, seq * S + qtile * kQTile);
          }
        }
        mbar_wait(&bars->o_done, ph);      // V_lo_i and V_hi_i consumed
        if (wn < num_items) load_q(wn);
        mbar_wait(&bars->stage_free, ph);  // the epilogue of item i no longer reads its staging tiles (V_hi region)
        if (wn < num_items) load_v(wn, false);
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    if (lane == 0) {
      const uint32_t sbase = smem_u32(smem);
      constexpr uint32_t idesc_s = make_idesc_bf16(kQTile, kKeyPad, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(kQTile, kHeadDim, 1);
      const int nterms = split ? 3 : 1;
      uint32_t it = 0;
      long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_loop = clock64();
      for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        long long t0 = clock64(), t1;
        // ---------------- S = Q K^T ----------------
        mbar_wait(&bars->q_full, ph);
        t1 = clock64(); acc[0] += t1 - t0; t0 = t1;
        mbar_wait(&bars->k_full, ph);
        t1 = clock64(); acc[1] += t1 - t0; t0 = t1;
        tc_fence_after();
        uint32_t accum = 0;
        for (int j = 0; j < 2; ++j) {
          for (int term = 0; term < nterms; ++term) {
            const uint32_t qo = (split && term == 0) ? kOffQLo : kOffQHi;
            const uint32_t ko = (split && term == 1) ? kOffKLo : kOffKHi;
            const uint64_t da = make_desc_kmajor_sw128(sbase + qo + j * kQBlockBytes);
            const uint64_t db = make_desc_kmajor_sw128(sbase + ko + j * kKVBlockBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_ss(tmem_base + kColS, desc_advance(da, kk * 32), desc_advance(db, kk * 32), idesc_s, accum);
              accum = 1;
            }
          }
        }
        umma_commit(&bars->s_done);
        t1 = clock64(); acc[2] += t1 - t0; t0 = t1;
        // ---------------- O = P V ----------------
        // (p_full of this item also implies the softmax warps have finished reading O of the previous item)
        mbar_wait(&bars->p_full, ph);
        t1 = clock64(); acc[3] += t1 - t0; t0 = t1;
        mbar_wait(&bars->vhi_full, ph);
        t1 = clock64(); acc[4] += t1 - t0; t0 = t1;
        tc_fence_after();
        const uint64_t dv_hi = make_desc_mnmajor_sw128(sbase + kOffVHi, kKVBlockBytes);
        accum = 0;
        if (split) {
#pragma unroll 1
          for (int ks = 0; ks < kKeyPad / 16; ++ks) {
            umma_ts(tmem_base + kColO, tmem_base + kColPLo + ks * 8, desc_advance(dv_hi, ks * 2048), idesc_o, accum);
            accum = 1;
          }
        }
#pragma unroll 1
        for (int ks = 0; ks < kKeyPad / 16; ++ks) {
          umma_ts(tmem_base + kColO, tmem_base + kColPHi + ks * 8, desc_advance(dv_hi, ks * 2048), idesc_o, accum);
          accum = 1;
        }
        t1 = clock64(); acc[5] += t1 - t0; t0 = t1;
        if (split) {
          mbar_wait(&bars->vlo_full, ph);
          t1 = clock64(); acc[6] += t1 - t0; t0 = t1;
          tc_fence_after();
          const uint64_t dv_lo = make_desc_mnmajor_sw128(sbase + kOffQHi, kKVBlockBytes);
#pragma unroll 1
          for (int ks = 0; ks < kKeyPad / 16; ++ks)
            umma_ts(tmem_base + kColO, tmem_base + kColPHi + ks * 8, desc_advance(dv_lo, ks * 2048), idesc_o, 1u);
        }
        umma_commit(&bars->o_done);
        acc[7] += 1;
      }
      if (p.dbg_cycles) {
        for (int i = 0; i < 8; ++i) p.dbg_cycles[(size_t)blockIdx.x * 16 + i] = acc[i];
        unsigned long long ns_now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_now));
        p.dbg_cycles[(size_t)blockIdx.x * 16 + 12] = t_loop - t_entry;
        p.dbg_cycles[(size_t)blockIdx.x * 16 + 13] = clock64() - t_entry;
        p.dbg_cycles[(size_t)blockIdx.x * 16 + 14] = (long long)(ns_now - ns_entry);
        p.dbg_cycles[(size_t)blockIdx.x * 16 + 15] = (long long)ns_entry;
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax + output: two threads per query row ----------------
    const int sw = warp_idx - 2;
    const int lane_group = warp_idx & 3;
    const int half = sw >> 2;
    const int row = lane_group * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(lane_group * 32) << 16);
    const float c_scale = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
    const uint32_t stage = smem_u32(smem + kOffVHi + sw * kEpiStageBytes);
    const long long pitch = (long long)p.ld_out * 2;
    uint32_t it = 0;
    long long sacc[4] = {0, 0, 0, 0};
    for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      int seq, head, qtile;
      decode(w, seq, head, qtile);
      const int row0 = seq * S;
      long long u0 = clock64(), u1;
      mbar_wait(&bars->s_done, ph);
      u1 = clock64(); sacc[0] += u1 - u0; u0 = u1;
      tc_fence_after();
      float sum;
      if (half == 0) {
        sum = softmax_half<0, kChunks0>(trow, S, split, c_scale, bars->red_max, 0, row, p.trunc_split != 0);
      } else {
        sum = softmax_half<kChunks0, kChunks1>(trow, S, split, c_scale, bars->red_max, 1, row, p.trunc_split != 0);
      }
      bars->red_sum[half][row] = sum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars->p_full);
      named_barrier_sync(1, kNumSoftmaxWarps * 32);  // red_sum of both halves visible
      const float inv = 1.0f / (bars->red_sum[0][row] + bars->red_sum[1][row]);

      u1 = clock64(); sacc[1] += u1 - u0; u0 = u1;
      mbar_wait(&bars->o_done, ph);
      u1 = clock64(); sacc[2] += u1 - u0; u0 = u1;
      tc_fence_after();
      // all MMAs of this item are complete: the V_hi region is free and becomes the store-staging area (8 x 4 KB)
      uint32_t v0[32], v1[32];
      tmem_ld32(trow + kColO + half * 64, v0);
      tmem_ld32(trow + kColO + half * 64 + 32, v1);
      tmem_ld_wait();
      uint32_t hw[32], lw[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (p.trunc_split) {
          split_bf16x2_trunc(__uint_as_float(v0[2 * j]) * inv, __uint_as_float(v0[2 * j + 1]) * inv, hw[j], lw[j]);
          split_bf16x2_trunc(__uint_as_float(v1[2 * j]) * inv, __uint_as_float(v1[2 * j + 1]) * inv, hw[16 + j], lw[16 + j]);
        } else {
          split_bf16x2(__uint_as_float(v0[2 * j]) * inv, __uint_as_float(v0[2 * j + 1]) * inv, hw[j], lw[j]);
          split_bf16x2(__uint_as_float(v1[2 * j]) * inv, __uint_as_float(v1[2 * j + 1]) * inv, hw[16 + j], lw[16 + j]);
        }
      }
      const int group_row0 = qtile * kQTile + lane_group * 32;  // first query of this warp's 32 rows
      if (group_row0 + 32 <= S) {
        store_block_tma(stage, lane, hw, &map_o_hi, head * kHeadDim + half * 64, row0 + group_row0);
        if (p.nsplit_out == 3) store_block_tma(stage, lane, lw, &map_o_lo, head * kHeadDim + half * 64, row0 + group_row0);
        if (lane == 0) tma_store_wait_read();
      } else if (group_row0 < S) {
        RowSlots rows;
        rows.ok = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int qp = group_row0 + r * 4 + (lane >> 3);
          rows.row[r] = (qp < S) ? (row0 + qp) : 0;
          if (qp < S) rows.ok |= 1u << r;
        }
        char* dst_hi = reinterpret_cast<char*>(p.out_hi + head * kHeadDim + half * 64);
        store_block_coalesced(stage, lane, hw, dst_hi, rows, pitch, 8, 1, 0);
        if (p.nsplit_out == 3) {
          char* dst_lo = reinterpret_cast<char*>(p.out_lo + head * kHeadDim + half * 64);
          store_block_coalesced(stage, lane, lw, dst_lo, rows, pitch, 8, 1, 0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->stage_free);
      u1 = clock64(); sacc[3] += u1 - u0;
    }
    if (p.dbg_cycles && sw == 0 && lane == 0)
      for (int i = 0; i < 4; ++i) p.dbg_cycles[(size_t)blockIdx.x * 16 + 8 + i] = sacc[i];
    if (lane == 0) tma_store_wait_all();  // bulk stores of the last item complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

cudaError_t configure_attention_kernel() {
  return cudaFuncSetAttribute(attention_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(1024 + kSmemTiles + sizeof(AttnBarriers2)));
}

cudaError_t launch_attention(const CUtensorMap& q_hi, const CUtensorMap& q_lo, const CUtensorMap& kv_hi,
                             const CUtensorMap& kv_lo, const CUtensorMap& o_hi, const CUtensorMap& o_lo, const AttnParams& p,
                             cudaStream_t stream) {
  if (p.seq_len > kKeyPad || p.seq_len < 1 || (p.nsplit != 1 && p.nsplit != 3)) {
    set_last_error("launch_attention: unsupported seq_len=%d nsplit=%d", p.seq_len, p.nsplit);
    return cudaErrorInvalidValue;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int q_tiles = (p.seq_len + kQTile - 1) / kQTile;
  const int num_items = q_tiles * p.num_heads * p.num_seqs;
  const size_t smem2 = 1024 + kSmemTiles + sizeof(AttnBarriers2);
  return launch_kernel(attention_persistent_kernel, dim3(num_items < num_sms ? num_items : num_sms), dim3(kThreads), smem2,
                       stream, q_hi, q_lo, kv_hi, kv_lo, o_hi, o_lo, p, num_items, q_tiles);
}

}  // namespace cmdi
