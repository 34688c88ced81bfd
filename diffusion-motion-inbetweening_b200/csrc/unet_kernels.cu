// Row / elementwise kernels of the MDM_UNET denoiser (reference: model/mdm_unet.py).
//
// Activations are channel-last [rows, C] with a HALO layout: sequence b of a level with L positions occupies rows
// b * Lp .. b * Lp + Lp - 1 with Lp = 256 >> level (L = 224 >> level), position l at row b * Lp + 2 + l; the rows before
// and after a sequence's positions are zero.  A Conv1d(k = 5, padding = 2) is then a GEMM whose reduction walks five
// row-shifted views of the same matrix (gemm2.cu, LinearParams::num_taps), the zero rows are the convolution's padding,
// and with Lp halving per level the stride-2 convolutions become stride-1 over row PAIRS (engine_unet.cu).
//
//   unet_input_kernel        x <- obs_x0 * M + x * ~M ; cat([x, M])      mdm_unet.py:778-783 (+ the 224-frame padding :817)
//   unet_emb_kernel          emb = time_embed(pe[t]) (+ embed_text(cond))  mdm_unet.py:794-803
//   groupnorm_mish_kernel    GroupNorm(8) -> [AdaGN scale/shift] -> Mish -> [+ residual]   mdm_unet.py:33-100, :159-218
//   conv / conv-transpose weight re-layouts (once per weight load)
#include "common.cuh"
#include "kernels.h"

namespace cmdi {

namespace {

__device__ __forceinline__ float mish_dev(float x) { return x * tanhf(x > 20.0f ? x : log1pf(expf(x))); }

// ---------------------------------------------------------------------------------------------
// network input: planes [rows0, ld] (hi, lo), channels [0, D) = keyframe-blended x_t, [D, 2D) = mask (keyframe-
// conditioned models), zero elsewhere.  One thread per (sequence copy, frame, channel pair).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unet_input_kernel(const UnetInputParams p) {
  const int pairs = p.ld / 2;
  const size_t total = (size_t)p.B * p.L * pairs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cp = (int)(i % pairs);
    const size_t fl = i / pairs;
    const int l = (int)(fl % p.L), b = (int)(fl / p.L);
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = cp * 2 + j;
      float val = 0.f;
      const size_t src = ((size_t)b * p.L + l) * p.D_pad;
      if (c < p.D) {
        val = p.x_t[src + c];
        if (p.obs_mask && p.obs_mask[src + c]) val = p.obs[src + c];
      } else if (p.obs_mask && c < 2 * p.D) {
        val = p.obs_mask[src + c - p.D] ? 1.0f : 0.0f;
      }
      v[j] = val;
    }
    uint32_t hw, lw;
    split_bf16x2(v[0], v[1], hw, lw);
    for (int copy = 0; copy < p.copies; ++copy) {
      const size_t row = (size_t)(b + copy * p.B) * p.row_period + p.row_lo + l;
      *reinterpret_cast<uint32_t*>(p.out_hi + row * p.ld + cp * 2) = hw;
      if (p.out_lo) *reinterpret_cast<uint32_t*>(p.out_lo + row * p.ld + cp * 2) = lw;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// emb[seq, :] = temb_table[tmap[t], :] + (cond_proj[seq % B] if seq < n_cond else uncond_proj)   -> bf16 planes [rows, 512]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) unet_emb_kernel(const TokenParams p) {
  const int seq = blockIdx.x;
  int t = *p.step_ptr;
  if (p.timestep_map) t = p.timestep_map[t];
  const int col = threadIdx.x * 4;
  float4 e = *reinterpret_cast<const float4*>(p.temb_table + (size_t)t * 512 + col);
  if (p.cond_proj) {
    const float* c = (seq < p.n_cond_seqs) ? p.cond_proj + (size_t)(seq % p.seq_len) * 512 : p.uncond_proj;  // seq_len carries B here
    const float4 cv = *reinterpret_cast<const float4*>(c + col);
    e.x += cv.x; e.y += cv.y; e.z += cv.z; e.w += cv.w;
  }
  uint32_t h01, l01, h23, l23;
  split_bf16x2(e.x, e.y, h01, l01);
  split_bf16x2(e.z, e.w, h23, l23);
  *reinterpret_cast<uint2*>(p.x_hi + (size_t)seq * 512 + col) = make_uint2(h01, h23);
  if (p.x_lo) *reinterpret_cast<uint2*>(p.x_lo + (size_t)seq * 512 + col) = make_uint2(l01, l23);
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(groups of Cg channels over the L positions of one sequence) + optional AdaGN + Mish + optional residual.
// One CTA per (group, sequence): the [L, Cg] fp32 slab is staged in shared memory (<= 224 x 128 x 4 = 112 KB), the
// statistics are two-pass over that copy (mean, then centred squares: torch's group_norm numerics to fp32 rounding).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) groupnorm_mish_kernel(const GroupNormParams p) {
  extern __shared__ float slab[];  // [L][Cg]
  __shared__ float red[16];
  __shared__ float s_mean, s_rstd;
  const int g = blockIdx.x, b = blockIdx.y;
  const int Cg = p.C / p.groups, L = p.L;
  const int c0 = g * Cg;
  const size_t row0 = (size_t)b * p.row_period + p.row_lo;
  const int vec = Cg / 4;  // float4 per row
  const int n4 = L * vec;
  float sum = 0.f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int l = i / vec, q = i - l * vec;
    const float4 v = *reinterpret_cast<const float4*>(p.y + (row0 + l) * p.ld_y + c0 + q * 4);
    reinterpret_cast<float4*>(slab)[i] = v;
    sum += (v.x + v.y) + (v.z + v.w);
  }
  auto block_sum = [&](float v) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
  };
  const float inv_n = 1.0f / (float)(L * Cg);
  const float mean = block_sum(sum) * inv_n;
  float sq = 0.f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = reinterpret_cast<float4*>(slab)[i];
    const float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
    sq += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = rsqrtf(block_sum(sq) * inv_n + p.eps);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int l = i / vec, q = i - l * vec;
    const int c = c0 + q * 4;
    const float4 v = reinterpret_cast<float4*>(slab)[i];
    const float4 gm = *reinterpret_cast<const float4*>(p.gamma + c);
    const float4 bt = *reinterpret_cast<const float4*>(p.beta + c);
    float o[4] = {(v.x - mean) * rstd * gm.x + bt.x, (v.y - mean) * rstd * gm.y + bt.y, (v.z - mean) * rstd * gm.z + bt.z,
                  (v.w - mean) * rstd * gm.w + bt.w};
    if (p.ada) {
      // cond = time_mlp(t) as [scale | shift] (mdm_unet.py:95-99): x * (1 + scale) + shift
      const float4 sc = *reinterpret_cast<const float4*>(p.ada + (size_t)b * p.ld_ada + c);
      const float4 sh = *reinterpret_cast<const float4*>(p.ada + (size_t)b * p.ld_ada + p.C + c);
      o[0] = o[0] * (1.0f + sc.x) + sh.x; o[1] = o[1] * (1.0f + sc.y) + sh.y;
      o[2] = o[2] * (1.0f + sc.z) + sh.z; o[3] = o[3] * (1.0f + sc.w) + sh.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = mish_dev(o[j]);
    const size_t row = row0 + l;
    if (p.res_f32) {
      const float4 r = *reinterpret_cast<const float4*>(p.res_f32 + row * p.ld_res + c);
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    } else if (p.res_hi) {
      const uint2 rh = *reinterpret_cast<const uint2*>(p.res_hi + row * p.ld_res + c);
      const uint2 rl = *reinterpret_cast<const uint2*>(p.res_lo + row * p.ld_res + c);
      o[0] += __uint_as_float(rh.x << 16) + __uint_as_float(rl.x << 16);
      o[1] += __uint_as_float(rh.x & 0xffff0000u) + __uint_as_float(rl.x & 0xffff0000u);
      o[2] += __uint_as_float(rh.y << 16) + __uint_as_float(rl.y << 16);
      o[3] += __uint_as_float(rh.y & 0xffff0000u) + __uint_as_float(rl.y & 0xffff0000u);
    }
    uint32_t h01, l01, h23, l23;
    split_bf16x2(o[0], o[1], h01, l01);
    split_bf16x2(o[2], o[3], h23, l23);
    *reinterpret_cast<uint2*>(p.out_hi + row * p.ld_out + c) = make_uint2(h01, h23);
    if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + row * p.ld_out + c) = make_uint2(l01, l23);
  }
}

// Conv1d weight [Co, Ci, k] fp32 -> tap-major planes [Co, k * Cp] (Cp >= Ci, zero padded):  W2[o, j * Cp + c] = W[o, c, j]
__global__ void conv_weight_planes_kernel(const float* __restrict__ w, int Co, int Ci, int k, int Cp, __nv_bfloat16* hi,
                                          __nv_bfloat16* lo, int ld) {
  const size_t total = (size_t)Co * k * Cp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int j = (int)((i / Cp) % k);
    const int o = (int)(i / ((size_t)Cp * k));
    const float v = c < Ci ? w[((size_t)o * Ci + c) * k + j] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[(size_t)o * ld + (size_t)j * Cp + c] = h;
    lo[(size_t)o * ld + (size_t)j * Cp + c] = l;
  }
}

// ConvTranspose1d(k = 4, s = 2, p = 1) weight [Ci, Co, 4] -> planes [2 Co, 3 Ci]: even outputs (rows [0, Co)) use
// in[m - 1] W[..3] + in[m] W[..1], odd outputs (rows [Co, 2 Co)) use in[m] W[..2] + in[m + 1] W[..0]; tap blocks are
// ordered (m - 1, m, m + 1) and the two unused blocks are zero.
__global__ void convt_weight_planes_kernel(const float* __restrict__ w, int Ci, int Co, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld) {
  const size_t total = (size_t)2 * Co * 3 * Ci;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ci);
    const int tap = (int)((i / Ci) % 3);
    const int n = (int)(i / ((size_t)3 * Ci));
    const int o = n % Co, odd = n / Co;
    int j = -1;
    if (!odd) j = tap == 0 ? 3 : (tap == 1 ? 1 : -1);
    else j = tap == 1 ? 2 : (tap == 2 ? 0 : -1);
    const float v = j >= 0 ? w[((size_t)c * Co + o) * 4 + j] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[(size_t)n * ld + (size_t)tap * Ci + c] = h;
    lo[(size_t)n * ld + (size_t)tap * Ci + c] = l;
  }
}

inline dim3 grid_1d(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return dim3((unsigned)g);
}

}  // namespace

cudaError_t launch_unet_input(const UnetInputParams& p, cudaStream_t stream) {
  unet_input_kernel<<<grid_1d((size_t)p.B * p.L * (p.ld / 2), 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_unet_emb(const TokenParams& p, cudaStream_t stream) {
  unet_emb_kernel<<<p.num_seqs, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t configure_groupnorm_kernel() {
  return cudaFuncSetAttribute(groupnorm_mish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 256 * 4);
}

cudaError_t launch_groupnorm_mish(const GroupNormParams& p, int num_seqs, cudaStream_t stream) {
  const int Cg = p.C / p.groups;
  if (p.C % p.groups || Cg % 4 || (size_t)p.L * Cg * 4 > 224 * 256 * 4) {
    set_last_error("launch_groupnorm_mish: unsupported shape C=%d groups=%d L=%d", p.C, p.groups, p.L);
    return cudaErrorInvalidValue;
  }
  groupnorm_mish_kernel<<<dim3(p.groups, num_seqs), 512, (size_t)p.L * Cg * 4, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_conv_weight_planes(const float* w, int Co, int Ci, int k, int Cp, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld,
                                      cudaStream_t stream) {
  conv_weight_planes_kernel<<<grid_1d((size_t)Co * k * Cp, 256), 256, 0, stream>>>(w, Co, Ci, k, Cp, hi, lo, ld);
  return cudaGetLastError();
}

cudaError_t launch_convt_weight_planes(const float* w, int Ci, int Co, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld, cudaStream_t stream) {
  convt_weight_planes_kernel<<<grid_1d((size_t)2 * Co * 3 * Ci, 256), 256, 0, stream>>>(w, Ci, Co, hi, lo, ld);
  return cudaGetLastError();
}

}  // namespace cmdi
