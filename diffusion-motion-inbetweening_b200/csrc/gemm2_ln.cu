// Linear layer fused with the post-norm residual LayerNorm of nn.TransformerEncoderLayer (mdm.py:107-114):
//
//     y = LayerNorm( residual + A W^T + bias ) * gamma + beta          (N = 512 = d_model, one full row per tile)
//
// used for the attention out-projection (+norm1) and FFN linear2 (+norm2).  Unfused, the pre-norm sum made a round
// trip through HBM/L2 (25.8 MB written by the GEMM, read by the LayerNorm kernel) and the LayerNorm kernel itself
// cost 12.6 us x 16 per step; with BLOCK_N = 128 tiles the GEMM also ran at < 1 PFLOP/s.  Here a CTA pair owns a
// 256 x 512 tile (cta_group::2, two N = 256 MMAs per k-step into the full 512 TMEM columns), so each CTA holds 128
// COMPLETE rows and the epilogue can normalise them in place:
//
//   pass 1  TMEM -> +bias +residual -> written back to TMEM; shifted sum and sum of squares of each thread's 256
//           columns, the two threads of a row combine (mean, M2) by Chan's formula
//   pass 2  normalise, scale/shift, fp32 + bf16 hi/lo planes -> TMA stores
//
// Protocol identical to gemm2.cu (leader-owned full / tmem_empty barriers, multicast commits), single accumulator.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace cmdi {

namespace {

constexpr int kBlockM = 128;  // per CTA; 256 per pair
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kN = 512;       // full rows
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + kNumEpiWarps * 32;
constexpr int kMaxStages = 4;
constexpr int kSmemLimit = 232448;
constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
constexpr int kBHalfBytes = 128 * kBlockK * 2;        // this CTA's 128 rows of one 256-row W block
constexpr uint32_t kTmemCols = 512;

struct __align__(8) LnBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
  float red[2][128];  // row statistics exchanged between the two warps that share a row
};

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kNumEpiWarps * 32) : "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
linear2_ln_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                  const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                  const __grid_constant__ CUtensorMap map_o_f32, const LinearLnParams p, const int num_stages,
                  const int num_m_pairs) {
  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int nplanes = (p.nsplit == 3) ? 2 : 1;
  const uint32_t stage_bytes = nplanes * (kABytes + 2 * kBHalfBytes);  // A + both 256-row W blocks (this CTA's halves)
  uint8_t* epi_stage = smem + (size_t)num_stages * stage_bytes;
  LnBarriers* bars = reinterpret_cast<LnBarriers*>(epi_stage + kNumEpiWarps * kEpiStageBytes);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_k_blocks = (p.K + kBlockK - 1) / kBlockK;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_w_hi);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->tmem_full, 1);
    mbar_init(&bars->tmem_empty, 2 * kNumEpiWarps);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc_2sm(&bars->tmem_base, kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_wait();

  if (warp_idx == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_m_pairs; tile += num_clusters) {
        const int m_blk = 2 * tile + (int)cta_rank;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + nplanes * kABytes;
          if (leader) mbar_arrive_expect_tx(&bars->full[stage], 2 * stage_bytes);
          for (int pl = 0; pl < nplanes; ++pl) {
            const CUtensorMap* ma = pl ? &map_a_lo : &map_a_hi;
            const CUtensorMap* mw = pl ? &map_w_lo : &map_w_hi;
            tma_load_2d_2sm(sa + pl * kABytes, ma, &bars->full[stage], kb * kBlockK, m_blk * kBlockM);
            // W block h covers output columns [256h, 256h + 256); this CTA stages rows [256h + 128*rank, +128)
            for (int h = 0; h < 2; ++h)
              tma_load_2d_2sm(sb + (pl * 2 + h) * kBHalfBytes, mw, &bars->full[stage], kb * kBlockK,
                              h * 256 + (int)cta_rank * 128);
          }
          if (++stage == num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ====================================== MMA issuer (leader CTA only) ======================================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, 256, 0);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < num_m_pairs; tile += num_clusters) {
        mbar_wait(&bars->tmem_empty, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + nplanes * kABytes;
          // terms: (A_lo, W_hi), (A_hi, W_lo), (A_hi, W_hi); plain bf16: (A_hi, W_hi)
          const int nterms = (nplanes == 2) ? 3 : 1;
          for (int term = 0; term < nterms; ++term) {
            const int apl = (nplanes == 2 && term == 0) ? 1 : 0;
            const int wpl = (nplanes == 2 && term == 1) ? 1 : 0;
            const uint64_t da = make_desc_kmajor_sw128(sa + apl * kABytes);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t db = make_desc_kmajor_sw128(sb + (wpl * 2 + h) * kBHalfBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                umma_ss_2sm(tmem_base + h * 256, desc_advance(da, k * kUmmaK * 2), desc_advance(db, k * kUmmaK * 2), idesc,
                            (kb > 0 || term > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit_2sm(&bars->empty[stage], 0x3);
          if (++stage == num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&bars->tmem_full, 0x3);
        acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue: residual + bias, LayerNorm over the full 512-wide row =======================
    const int epi = warp_idx - 2;
    const int lane_group = warp_idx & 3;
    const int half = epi >> 2;  // columns [256*half, +256)
    const int row = lane_group * 32 + lane;
    const uint32_t stage = smem_u32(epi_stage + epi * kEpiStageBytes);
    const uint32_t trow = tmem_base + ((uint32_t)(lane_group * 32) << 16) + half * 256;
    uint32_t acc_phase = 0;
    // per-column vectors of this warp's 256 columns: lane l holds columns 4l..4l+3 and 128+4l..
    float4 bias4[2], gamma4[2], beta4[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = half * 256 + q * 128 + lane * 4;
      bias4[q] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
      gamma4[q] = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
      beta4[q] = __ldg(reinterpret_cast<const float4*>(p.beta + n));
    }
    for (int tile = cluster_id; tile < num_m_pairs; tile += num_clusters) {
      const int m_blk = 2 * tile + (int)cta_rank;
      const int warp_row0 = m_blk * kBlockM + lane_group * 32;
      RowSlots rows;
      rows.ok = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = warp_row0 + it * 4 + (lane >> 3);
        rows.row[it] = r < p.M ? r : 0;
        if (r < p.M) rows.ok |= 1u << it;
      }
      mbar_wait(&bars->tmem_full, acc_phase);
      tc_fence_after();

      // ---- pass 1: v = acc + bias + residual -> back into TMEM; shifted sums of this thread's 256 columns
      //      (shift = the first value, so the single-pass variance does not cancel) ----
      float s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + c * 64, v0);
        tmem_ld32(trow + c * 64 + 32, v1);
        uint32_t r0[32], r1[32];
        load_block_coalesced(stage, lane, r0, reinterpret_cast<const char*>(p.residual + half * 256 + c * 64), rows,
                             (long long)kN * 4, 8);
        load_block_coalesced(stage, lane, r1, reinterpret_cast<const char*>(p.residual + half * 256 + c * 64 + 32), rows,
                             (long long)kN * 4, 8);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const int off = c * 64 + g * 4;         // column offset inside this warp's 256
          const int q = off >> 7, src = (off & 127) >> 2;
          const float b0 = __shfl_sync(0xffffffffu, q ? bias4[1].x : bias4[0].x, src);
          const float b1 = __shfl_sync(0xffffffffu, q ? bias4[1].y : bias4[0].y, src);
          const float b2 = __shfl_sync(0xffffffffu, q ? bias4[1].z : bias4[0].z, src);
          const float b3 = __shfl_sync(0xffffffffu, q ? bias4[1].w : bias4[0].w, src);
          uint32_t* vv = (g < 8) ? &v0[g * 4] : &v1[(g - 8) * 4];
          const uint32_t* rr = (g < 8) ? &r0[g * 4] : &r1[(g - 8) * 4];
          const float x0 = __uint_as_float(vv[0]) + b0 + __uint_as_float(rr[0]);
          const float x1 = __uint_as_float(vv[1]) + b1 + __uint_as_float(rr[1]);
          const float x2 = __uint_as_float(vv[2]) + b2 + __uint_as_float(rr[2]);
          const float x3 = __uint_as_float(vv[3]) + b3 + __uint_as_float(rr[3]);
          if (c == 0 && g == 0) shift = x0;
          const float d0 = x0 - shift, d1 = x1 - shift, d2 = x2 - shift, d3 = x3 - shift;
          s1 += (d0 + d1) + (d2 + d3);
          s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
          vv[0] = __float_as_uint(x0); vv[1] = __float_as_uint(x1); vv[2] = __float_as_uint(x2); vv[3] = __float_as_uint(x3);
        }
        tmem_st32(trow + c * 64, v0);
        tmem_st32(trow + c * 64 + 32, v1);
      }
      tmem_st_wait();
      // the two threads of a row combine (count, mean, M2) of their halves (Chan et al.): 256 columns each
      // (one 1 KB exchange array, used twice: the smem budget leaves no room for a second one next to two stages)
      bars->red[half][row] = shift + s1 * (1.0f / 256.0f);
      tc_fence_before();
      epi_barrier();
      tc_fence_after();
      const float m0 = bars->red[0][row], m1 = bars->red[1][row];
      const float mean = 0.5f * (m0 + m1);
      epi_barrier();
      bars->red[half][row] = s2 - s1 * s1 * (1.0f / 256.0f);
      epi_barrier();
      const float m2 = bars->red[0][row] + bars->red[1][row] + (m0 - m1) * (m0 - m1) * 128.0f;
      const float rstd = rsqrtf(m2 * (1.0f / kN) + p.eps);

      // ---- pass 2: normalise, affine, store fp32 + bf16 planes ----
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v0[32], v1[32];
        tmem_ld32(trow + c * 64, v0);
        tmem_ld32(trow + c * 64 + 32, v1);
        tmem_ld_wait();
        float f[64];
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const int off = c * 64 + g * 4;
          const int q = off >> 7, src = (off & 127) >> 2;
          const float4 gq = q ? gamma4[1] : gamma4[0];
          const float4 bq = q ? beta4[1] : beta4[0];
          const uint32_t* vv = (g < 8) ? &v0[g * 4] : &v1[(g - 8) * 4];
          f[g * 4 + 0] = (__uint_as_float(vv[0]) - mean) * rstd * __shfl_sync(0xffffffffu, gq.x, src) + __shfl_sync(0xffffffffu, bq.x, src);
          f[g * 4 + 1] = (__uint_as_float(vv[1]) - mean) * rstd * __shfl_sync(0xffffffffu, gq.y, src) + __shfl_sync(0xffffffffu, bq.y, src);
          f[g * 4 + 2] = (__uint_as_float(vv[2]) - mean) * rstd * __shfl_sync(0xffffffffu, gq.z, src) + __shfl_sync(0xffffffffu, bq.z, src);
          f[g * 4 + 3] = (__uint_as_float(vv[3]) - mean) * rstd * __shfl_sync(0xffffffffu, gq.w, src) + __shfl_sync(0xffffffffu, bq.w, src);
        }
        const int n0 = half * 256 + c * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(f[h * 32 + j]);
          store_block_tma(stage, lane, w, &map_o_f32, n0 + h * 32, warp_row0);
        }
        uint32_t hw[32], lw[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) split_bf16x2(f[2 * j], f[2 * j + 1], hw[j], lw[j]);
        store_block_tma(stage, lane, hw, &map_o_hi, n0, warp_row0);
        if (p.nsplit_out == 3) store_block_tma(stage, lane, lw, &map_o_lo, n0, warp_row0);
      }
      // accumulator drained
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_on_leader(&bars->tmem_empty);
      acc_phase ^= 1;
      epi_barrier();  // red[] may be rewritten by the next tile
    }
    if (lane == 0) tma_store_wait_read();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

}  // namespace

cudaError_t configure_linear_ln_kernel() {
  return cudaFuncSetAttribute(linear2_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
}

// Tensor maps: A box {64, 128}; W box {64, 128} (W is [512, K]); outputs as TMA-store targets (bf16 {64, 32}, fp32 {32, 32}).
cudaError_t launch_linear_ln(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                             const CUtensorMap& o_hi, const CUtensorMap& o_lo, const CUtensorMap& o_f32,
                             const LinearLnParams& p, int num_sms, cudaStream_t stream) {
  if (p.nsplit != 1 && p.nsplit != 3) {
    set_last_error("launch_linear_ln: nsplit must be 1 or 3");
    return cudaErrorInvalidValue;
  }
  const int nplanes = (p.nsplit == 3) ? 2 : 1;
  const int stage_bytes = nplanes * (kABytes + 2 * kBHalfBytes);
  int num_stages = (kSmemLimit - 1024 - kNumEpiWarps * kEpiStageBytes - (int)sizeof(LnBarriers)) / stage_bytes;
  if (num_stages > kMaxStages) num_stages = kMaxStages;
  if (num_stages < 2) {
    set_last_error("launch_linear_ln: not enough shared memory for a 2-stage pipeline");
    return cudaErrorInvalidValue;
  }
  const size_t smem = 1024 + (size_t)num_stages * stage_bytes + kNumEpiWarps * kEpiStageBytes + sizeof(LnBarriers);
  const int num_m_pairs = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);
  int clusters = num_sms / 2;
  if (clusters > num_m_pairs) clusters = num_m_pairs;
  return launch_kernel(linear2_ln_kernel, dim3(2 * clusters), dim3(kNumThreads), smem, stream, a_hi, a_lo, w_hi, w_lo, o_hi, o_lo,
                       o_f32, p, num_stages, num_m_pairs);
}

}  // namespace cmdi
