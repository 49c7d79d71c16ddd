// DAC ResidualVectorQuantize nearest-code search, all codebooks in one pass over z.
// Reference: edm_tts/models/dac/vector_quantizer.py:146-210 (residual loop), :33-67 + :75-91 (per-level search).
//
// The reference keeps a 1024-d residual per frame and re-reads / re-writes it for each of the 12 levels. Because every
// step is linear, the 8-d projected latent of level i can be written without the residual:
//     e_i = W_in_i (z - sum_{j<i} (W_out_j c_j[idx_j] + b_out_j)) + b_in_i
//         = (W_in_i z + b_in_i) - sum_{j<i} G[i][j][idx_j],       G[i][j][code] = W_in_i (W_out_j c_j[code] + b_out_j)
// so z is read exactly once ([frames,1024] x [1024,96] projection for all levels), and the level loop runs on 8-d
// vectors with the 66 small G tables (2.1 MB, L2 resident). The search itself follows the reference formula
//     dist = |e^|^2 - 2 e^ . c^ + |c^|^2 ,  idx = first argmax(-dist)      (e^, c^ L2-normalised, eps 1e-12)
#pragma once
#include "ptx.cuh"

namespace edm {

constexpr int kRvqLevels = 12;
constexpr int kRvqCodes = 1024;
constexpr int kRvqCbDim = 8;
constexpr int kRvqLatent = 1024;
constexpr int kRvqE = kRvqLevels * kRvqCbDim;  // 96 projected latents per frame
constexpr int kRvqFrames = 64;                 // frames per CTA
constexpr int kRvqKc = 32;                     // channels per staged chunk
constexpr int kRvqWP = 98;                     // smem pitch of the staged w_in chunk (2-way instead of 32-way store conflicts)

struct RvqParams {
  const void* z;        // [B, 1024, T], T contiguous
  int z_is_bf16;
  int B, T;
  int n_levels;         // <= 12
  const float* w_in;    // [96, 1024]  weight-norm folded in_proj, levels stacked
  const float* b_in;    // [96]
  const float* cb_norm; // [12, 1024, 8] L2-normalised codebooks
  const float* cb_n2;   // [12, 1024]    |c^|^2
  const float* g;       // [12, 12, 1024, 8]  G[i][j] (only j < i used)
  long long* codes;     // out [B, n_levels, T] int64
  const long long* forced; // teacher forcing [B, n_levels, T] or nullptr
  float* latents;       // out [B, 96, T] (projected latents e_i before normalisation) or nullptr
};

constexpr uint32_t kRvqSmemBytes = (kRvqKc * kRvqFrames + kRvqKc * kRvqWP + kRvqFrames * (kRvqE + 1) + kRvqCodes * kRvqCbDim + kRvqCodes) * 4;

__global__ void __launch_bounds__(256) rvq_encode_kernel(const RvqParams p) {
  extern __shared__ float smem_f[];
  float* s_z = smem_f;                           // [Kc][64 frames]
  float* s_w = s_z + kRvqKc * kRvqFrames;        // [Kc][96]
  float* s_e = s_w + kRvqKc * kRvqWP;            // [64][97]
  float* s_cb = s_e + kRvqFrames * (kRvqE + 1);  // [1024][8]
  float* s_n2 = s_cb + kRvqCodes * kRvqCbDim;    // [1024]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kRvqFrames;

  // ---- phase A: E[f][n] = sum_c z[b][c][t0+f] * w_in[n][c]; thread = 4 frames x 6 outputs
  const int fg = tid & 15;   // frames 4*fg .. 4*fg+3
  const int ng = tid >> 4;   // outputs 6*ng .. 6*ng+5
  float acc[4][6];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;

  for (int cbase = 0; cbase < kRvqLatent; cbase += kRvqKc) {
    // stage z chunk: Kc x 64 (coalesced along t)
    for (int e = tid; e < kRvqKc * kRvqFrames; e += 256) {
      const int c = e / kRvqFrames, f = e % kRvqFrames;
      const int t = t0 + f;
      float val = 0.f;
      if (t < p.T) {
        const long long off = (static_cast<long long>(b) * kRvqLatent + cbase + c) * p.T + t;
        val = p.z_is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.z)[off]) : static_cast<const float*>(p.z)[off];
      }
      s_z[c * kRvqFrames + f] = val;
    }
    // stage w chunk transposed: s_w[c][n] = w_in[n][cbase + c]
    for (int e = tid; e < kRvqKc * kRvqE; e += 256) {
      const int n = e / kRvqKc, c = e % kRvqKc;
      s_w[c * kRvqWP + n] = __ldg(p.w_in + static_cast<long long>(n) * kRvqLatent + cbase + c);
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < kRvqKc; ++c) {
      const float4 zf = *reinterpret_cast<const float4*>(s_z + c * kRvqFrames + 4 * fg);
      const float2 w0 = *reinterpret_cast<const float2*>(s_w + c * kRvqWP + 6 * ng);
      const float2 w1 = *reinterpret_cast<const float2*>(s_w + c * kRvqWP + 6 * ng + 2);
      const float2 w2 = *reinterpret_cast<const float2*>(s_w + c * kRvqWP + 6 * ng + 4);
      const float zz[4] = {zf.x, zf.y, zf.z, zf.w};
      const float ww[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(zz[i], ww[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) s_e[(4 * fg + i) * (kRvqE + 1) + 6 * ng + j] = acc[i][j] + __ldg(p.b_in + 6 * ng + j);
  __syncthreads();

  // ---- phase B: 12 sequential levels in the 8-d space; 4 threads per frame
  const int f = tid >> 2, sub = tid & 3;
  const int t = t0 + f;
  const bool live = t < p.T;
  int chosen[kRvqLevels];
#pragma unroll
  for (int lvl = 0; lvl < kRvqLevels; ++lvl) {
    if (lvl >= p.n_levels) break;
    // stage this level's normalised codebook
    for (int e = tid; e < kRvqCodes * kRvqCbDim / 4; e += 256)
      reinterpret_cast<float4*>(s_cb)[e] = __ldg(reinterpret_cast<const float4*>(p.cb_norm + static_cast<long long>(lvl) * kRvqCodes * kRvqCbDim) + e);
    for (int e = tid; e < kRvqCodes; e += 256) s_n2[e] = __ldg(p.cb_n2 + lvl * kRvqCodes + e);

    // latent e (components 2*sub, 2*sub+1 owned by this thread)
    float e0 = s_e[f * (kRvqE + 1) + lvl * 8 + 2 * sub];
    float e1 = s_e[f * (kRvqE + 1) + lvl * 8 + 2 * sub + 1];
#pragma unroll
    for (int j = 0; j < lvl; ++j) {
      const float2 gj = __ldg(reinterpret_cast<const float2*>(
          p.g + ((static_cast<long long>(lvl) * kRvqLevels + j) * kRvqCodes + chosen[j]) * kRvqCbDim + 2 * sub));
      e0 -= gj.x;
      e1 -= gj.y;
    }
    if (p.latents != nullptr && live) {
      p.latents[(static_cast<long long>(b) * kRvqE + lvl * 8 + 2 * sub) * p.T + t] = e0;
      p.latents[(static_cast<long long>(b) * kRvqE + lvl * 8 + 2 * sub + 1) * p.T + t] = e1;
    }
    float n2 = e0 * e0 + e1 * e1;
    n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
    n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    e0 *= inv;
    e1 *= inv;
    float en[8];
    const int lane = tid & 31, lbase = lane & ~3;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      en[2 * s] = __shfl_sync(0xffffffffu, e0, lbase + s);
      en[2 * s + 1] = __shfl_sync(0xffffffffu, e1, lbase + s);
    }
    float en2 = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) en2 = fmaf(en[d], en[d], en2);
    __syncthreads();  // codebook staged

    float best = -INFINITY;
    int best_idx = 0;
#pragma unroll 4
    for (int m = 0; m < kRvqCodes / 4; ++m) {
      const int k = 4 * m + sub;
      const float4 ca = *reinterpret_cast<const float4*>(s_cb + k * 8);
      const float4 cb = *reinterpret_cast<const float4*>(s_cb + k * 8 + 4);
      float dot = 0.f;
      dot = fmaf(2.0f * en[0], ca.x, dot);
      dot = fmaf(2.0f * en[1], ca.y, dot);
      dot = fmaf(2.0f * en[2], ca.z, dot);
      dot = fmaf(2.0f * en[3], ca.w, dot);
      dot = fmaf(2.0f * en[4], cb.x, dot);
      dot = fmaf(2.0f * en[5], cb.y, dot);
      dot = fmaf(2.0f * en[6], cb.z, dot);
      dot = fmaf(2.0f * en[7], cb.w, dot);
      const float neg = -((en2 - dot) + s_n2[k]);
      if (neg > best) {
        best = neg;
        best_idx = k;
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      if (ob > best || (ob == best && oi < best_idx)) {
        best = ob;
        best_idx = oi;
      }
    }
    const long long oidx = (static_cast<long long>(b) * p.n_levels + lvl) * p.T + t;
    if (live && sub == 0) p.codes[oidx] = best_idx;
    chosen[lvl] = (p.forced != nullptr && live) ? static_cast<int>(p.forced[oidx]) : best_idx;
    __syncthreads();  // before the next level overwrites the staged codebook
  }
}

// ------------------------------------------------------------------------------------------------ codes -> features
// Reference: vector_quantizer.py:212-252 (from_codes / from_codes_unreduced): z_q = sum_i W_out_i c_i[code_i] + b_out_i,
// output channel-major [B, 1024, T] (or [B, L, 1024, T] unreduced). proj[i][code] = W_out_i c_i[code] + b_out_i is
// precomputed ([12, 1024, 1024] fp32), so this is a gather-sum plus a transpose through shared memory.
struct CodesToFeatParams {
  const long long* codes;  // [B, L, T]
  const float* proj;       // [12, 1024 codes, 1024 ch]
  float* out;              // [B, 1024, T] or [B, L, 1024, T]
  int B, L, T;
  int unreduced;
};

__global__ void __launch_bounds__(256) codes_to_features_kernel(const CodesToFeatParams p) {
  __shared__ float tile[32][kRvqLatent / 4 + 1];  // 32 frames x 256 channels (one quarter of the channels per pass)
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  const int n_out = p.unreduced ? p.L : 1;
  for (int o = 0; o < n_out; ++o) {
    const int l_lo = p.unreduced ? o : 0, l_hi = p.unreduced ? o + 1 : p.L;
    for (int cq = 0; cq < 4; ++cq) {
      // gather: thread (f = tid / 8, 8 threads x 32 channels each)
      {
        const int f = tid >> 3, part = tid & 7;
        const int t = t0 + f;
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        if (t < p.T) {
          for (int l = l_lo; l < l_hi; ++l) {
            const long long code = p.codes[(static_cast<long long>(b) * p.L + l) * p.T + t];
            const float4* src = reinterpret_cast<const float4*>(p.proj + (static_cast<long long>(l) * kRvqCodes + code) * kRvqLatent + cq * 256 + part * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = __ldg(src + i);
              acc[4 * i] += v.x;
              acc[4 * i + 1] += v.y;
              acc[4 * i + 2] += v.z;
              acc[4 * i + 3] += v.w;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[f][part * 32 + i] = acc[i];
      }
      __syncthreads();
      // scatter: channel-major, coalesced along t
      for (int e = tid; e < 256 * 32; e += 256) {
        const int c = e >> 5, f = e & 31;
        const int t = t0 + f;
        if (t < p.T) {
          const long long plane = p.unreduced ? (static_cast<long long>(b) * p.L + o) : b;
          p.out[(plane * kRvqLatent + cq * 256 + c) * p.T + t] = tile[f][c];
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace edm
