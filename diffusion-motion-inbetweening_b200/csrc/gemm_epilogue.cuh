// Epilogue shared by the 1-CTA and 2-CTA linear kernels: one 128 x BLOCK_N accumulator tile, TMEM -> registers ->
// bias / positional encoding / residual / GELU -> fp32 and bf16 hi/lo planes in global memory.
//
// Thread mapping: epilogue warp w owns TMEM lanes 32*(w%4).. (thread = output row) and half of the tile's columns,
// processed 64 columns at a time.  A thread therefore holds a ROW fragment, which is the wrong shape for global
// stores (32 lanes x 16 B with a multi-KB stride ran at < 1 TB/s; r01 time decomposition: stores were ~50 % of the
// QKV kernel).  Every 32-row x 128-byte block is therefore transposed through a per-warp 4 KB shared-memory
// staging tile (16-byte chunks XOR-swizzled by row & 7: conflict-free both ways) and written with fully coalesced
// 16-byte stores, 4 complete 128-byte row segments per instruction.
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace cmdi {

constexpr int kEpiBlockM = 128;
constexpr int kEpiStageBytes = 32 * 128;  // per epilogue warp

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Store a 32-row x 32-word block (thread `lane` holds row `lane` in w[0..31]) to global memory.
// Row r goes to dst_base + row_off_bytes(r) if row_ok(r); the 16-byte chunk c of a row is written if c < valid_chunks.
// `row_off` / `ok` are this thread's own row offset (bytes) and validity; other rows' are fetched by shuffle.
__device__ __forceinline__ void store_block_coalesced(uint32_t stage, int lane, const uint32_t (&w)[32], char* dst_base,
                                                      long long row_off, bool ok, int valid_chunks, int ncopies,
                                                      long long dup_bytes) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st_shared_v4(stage + lane * 128 + ((j ^ (lane & 7)) << 4), w[j * 4], w[j * 4 + 1], w[j * 4 + 2], w[j * 4 + 3]);
  __syncwarp();
  const int c = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3);
    const uint4 v = ld_shared_v4(stage + rr * 128 + ((c ^ (rr & 7)) << 4));
    const long long off = __shfl_sync(0xffffffffu, row_off, rr);
    const int rok = __shfl_sync(0xffffffffu, (int)ok, rr);
    if (rok && c < valid_chunks) {
      char* d = dst_base + off + c * 16;
      st_global_v4(d, v.x, v.y, v.z, v.w);
      if (ncopies == 2) st_global_v4(d + dup_bytes, v.x, v.y, v.z, v.w);
    }
  }
}

// Inverse of store_block_coalesced: fetch a 32-row x 128-byte block with coalesced 16-byte loads and hand every
// thread its own row (w[0..31]).
__device__ __forceinline__ void load_block_coalesced(uint32_t stage, int lane, uint32_t (&w)[32], const char* src_base,
                                                     long long row_off, bool ok, int valid_chunks) {
  __syncwarp();
  const int c = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3);
    const long long off = __shfl_sync(0xffffffffu, row_off, rr);
    const int rok = __shfl_sync(0xffffffffu, (int)ok, rr);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (rok && c < valid_chunks) v = *reinterpret_cast<const uint4*>(src_base + off + c * 16);
    st_shared_v4(stage + rr * 128 + ((c ^ (rr & 7)) << 4), v.x, v.y, v.z, v.w);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 v = ld_shared_v4(stage + lane * 128 + ((j ^ (lane & 7)) << 4));
    w[j * 4] = v.x; w[j * 4 + 1] = v.y; w[j * 4 + 2] = v.z; w[j * 4 + 3] = v.w;
  }
}

template <int BLOCK_N>
__device__ __forceinline__ void epilogue_tile(const LinearParams& p, uint32_t tmem_base, int acc, int m_blk, int n_blk,
                                              int lane_group, int col_part, int lane, uint32_t stage) {
  constexpr int kBlockM = kEpiBlockM;
  constexpr int kColsPerPart = BLOCK_N / 2;
  const int a_row = m_blk * kBlockM + lane_group * 32 + lane;
  bool valid = a_row < p.M;
  long long out_row = a_row;
  int pos = 0;  // sequence position for the positional-encoding add
  if (p.rowmap == ROWMAP_FRAMES_TO_SEQ) {
    const int b = a_row / p.frames;
    const int l = a_row - b * p.frames;
    out_row = (long long)b * (p.frames + 1) + l + 1;
    pos = l + 1;
  } else if (p.rowmap == ROWMAP_SEQ_TO_FRAMES) {
    const int S = p.frames + 1;
    const int b = a_row / S;
    const int s = a_row - b * S;
    valid = valid && (s > 0);
    out_row = (long long)b * p.frames + (s - 1);
  }
  if (!valid) out_row = 0;
  const int ncopies = (p.debug & 1) ? 0 : ((p.dup_row_offset > 0) ? 2 : 1);

#pragma unroll 1
  for (int c64 = 0; c64 < kColsPerPart / 64; ++c64) {
    const int col_in_tile = col_part * kColsPerPart + c64 * 64;
    const int n0 = n_blk * BLOCK_N + col_in_tile;
    uint32_t v0[32], v1[32];
    tmem_ld32(tmem_addr(tmem_base, lane_group * 32, acc * BLOCK_N + col_in_tile), v0);
    tmem_ld32(tmem_addr(tmem_base, lane_group * 32, acc * BLOCK_N + col_in_tile + 32), v1);
    tmem_ld_wait();
    if (n0 >= p.N) continue;  // warp-uniform
    float f[64];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      f[j] = __uint_as_float(v0[j]);
      f[32 + j] = __uint_as_float(v1[j]);
    }
    if (p.residual) {
      // x + sublayer(x): the fp32 residual rows are fetched coalesced (two 32-column blocks)
      const long long roff = out_row * (long long)p.ld_res * 4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int nb = n0 + h * 32;
        load_block_coalesced(stage, lane, h == 0 ? v0 : v1, reinterpret_cast<const char*>(p.residual + nb), roff, valid,
                             (p.N - nb) / 4);
      }
    }
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      const int n = n0 + g * 4;
      if (n < p.N) {
        if (p.bias) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          f[g * 4 + 0] += bv.x; f[g * 4 + 1] += bv.y; f[g * 4 + 2] += bv.z; f[g * 4 + 3] += bv.w;
        }
        if (p.pos_enc) {
          const float4 pv = __ldg(reinterpret_cast<const float4*>(p.pos_enc + (size_t)pos * p.N + n));
          f[g * 4 + 0] += pv.x; f[g * 4 + 1] += pv.y; f[g * 4 + 2] += pv.z; f[g * 4 + 3] += pv.w;
        }
        if (p.residual) {
          const uint32_t* rv = (g < 8) ? &v0[g * 4] : &v1[(g - 8) * 4];
          f[g * 4 + 0] += __uint_as_float(rv[0]); f[g * 4 + 1] += __uint_as_float(rv[1]);
          f[g * 4 + 2] += __uint_as_float(rv[2]); f[g * 4 + 3] += __uint_as_float(rv[3]);
        }
        if (p.act == 1) {
          f[g * 4 + 0] = gelu_erf(f[g * 4 + 0]); f[g * 4 + 1] = gelu_erf(f[g * 4 + 1]);
          f[g * 4 + 2] = gelu_erf(f[g * 4 + 2]); f[g * 4 + 3] = gelu_erf(f[g * 4 + 3]);
        }
      }
    }
    if (ncopies == 0) continue;
    if (p.out_f32) {
      // two 32-column fp32 blocks: 128 B per row each
      const long long roff = out_row * (long long)p.ld_f32 * 4;
      const long long dup = (long long)p.dup_row_offset * p.ld_f32 * 4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(f[h * 32 + j]);
        const int nb = n0 + h * 32;
        const int chunks = (p.N - nb) / 4;  // 16-byte chunks (4 floats) still inside N; may be <= 0 or >= 8
        store_block_coalesced(stage, lane, w, reinterpret_cast<char*>(p.out_f32 + nb), roff, valid, chunks, ncopies, dup);
      }
    }
    if (p.out_hi) {
      // one 64-column bf16 block per plane: 128 B per row
      uint32_t hw[32], lw[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) split_bf16x2(f[2 * j], f[2 * j + 1], hw[j], lw[j]);
      const long long roff = out_row * (long long)p.ld_bf * 2;
      const long long dup = (long long)p.dup_row_offset * p.ld_bf * 2;
      const int chunks = (p.N - n0) / 8;  // 16-byte chunks (8 bf16) still inside N
      store_block_coalesced(stage, lane, hw, reinterpret_cast<char*>(p.out_hi + n0), roff, valid, chunks, ncopies, dup);
      if (p.nsplit_out == 3)
        store_block_coalesced(stage, lane, lw, reinterpret_cast<char*>(p.out_lo + n0), roff, valid, chunks, ncopies, dup);
    }
  }
}

}  // namespace cmdi
