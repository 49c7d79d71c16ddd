// DAC ResidualVectorQuantize nearest-code search, all codebooks in one pass over z.
// Reference: edm_tts/models/dac/vector_quantizer.py:146-210 (residual loop), :33-67 + :75-91 (per-level search).
//
// The reference keeps a 1024-d residual per frame and re-reads / re-writes it for each of the 12 levels. Because every
// step is linear, the 8-d projected latent of level i can be written without the residual:
//     e_i = W_in_i (z - sum_{j<i} (W_out_j c_j[idx_j] + b_out_j)) + b_in_i
//         = (W_in_i z + b_in_i) - sum_{j<i} G[i][j][idx_j],       G[i][j][code] = W_in_i (W_out_j c_j[code] + b_out_j)
// so z is read exactly once ([frames,1024] x [1024,96] projection for all levels), and the level loop runs on 8-d
// vectors with the 66 small G tables (2.1 MB, L2 resident). The search itself follows the reference formula
//     dist = |e^|^2 - 2 e^ . c^ + |c^|^2 ,  idx = first argmin(dist)      (e^, c^ L2-normalised, eps 1e-12)
//
// Both contractions (the projection and the [frames,8] x [8,1024] distance matrix of every level) run on the tensor
// cores as 3xTF32 (mma.sync m16n8k8 with hi/lo operand splitting), which keeps fp32-level accuracy: the indices must
// match an fp32 reference except at genuine near-ties. bf16 / single-pass tf32 operands would flip ~1 % of the indices.
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
constexpr int kRvqZP = 72;                     // smem pitch of the staged z chunk  [k][frame]   (conflict-free A fragments)
constexpr int kRvqWP = 104;                    // smem pitch of the staged W chunk  [k][output]  (conflict-free B fragments)
constexpr int kRvqEP = 97;                     // smem pitch of the latent tile     [frame][96]

struct RvqParams {
  const void* z;        // [B, 1024, T], T contiguous
  int z_is_bf16;
  int B, T;
  int n_levels;         // <= 12
  const float* w_in_t;  // [1024, 96]  weight-norm folded in_proj, levels stacked, transposed (channel-major)
  const float* b_in;    // [96]
  const float* cb_norm; // [12, 1024, 8] L2-normalised codebooks
  const float* cb_n2;   // [12, 1024]    |c^|^2
  const float* g;       // [12, 12, 1024, 8]  G[i][j] (only j < i used)
  long long* codes;     // out [B, n_levels, T] int64
  const long long* forced; // teacher forcing [B, n_levels, T] or nullptr
  float* latents;       // out [B, 96, T] (projected latents e_i before normalisation) or nullptr
};

constexpr uint32_t kRvqSmemFloats = kRvqKc * kRvqZP + kRvqKc * kRvqWP + kRvqFrames * kRvqEP + kRvqCodes * kRvqCbDim + kRvqCodes +
                                    kRvqFrames * kRvqCbDim + kRvqFrames + 2 * kRvqFrames * 2;
constexpr uint32_t kRvqSmemBytes = kRvqSmemFloats * 4;

__device__ __forceinline__ void tf32_split(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float rem = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a * b with both operands split into tf32 hi + lo parts (the lo*lo term is below fp32 resolution)
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                           uint32_t bl0, uint32_t bl1) {
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bl0, bl1);
  mma_tf32(c, ah, bh0, bh1);
}

__global__ void __launch_bounds__(256, 2) rvq_encode_kernel(const RvqParams p) {
  extern __shared__ float smem_f[];
  float* s_z = smem_f;                            // [Kc][72]
  float* s_w = s_z + kRvqKc * kRvqZP;             // [Kc][104]
  float* s_e = s_w + kRvqKc * kRvqWP;             // [64][97]   projected latents of all levels
  float* s_cb = s_e + kRvqFrames * kRvqEP;        // [1024][8]  normalised codebook of the current level
  float* s_n2 = s_cb + kRvqCodes * kRvqCbDim;     // [1024]
  float* s_en = s_n2 + kRvqCodes;                 // [64][8]    normalised latent of the current level
  float* s_en2 = s_en + kRvqFrames * kRvqCbDim;   // [64]
  float* s_bd = s_en2 + kRvqFrames;               // [2 halves][64] best distance
  int* s_bi = reinterpret_cast<int*>(s_bd + 2 * kRvqFrames);  // [2 halves][64] best index

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kRvqFrames;
  const int mt = warp & 3;        // 16-frame tile of this warp
  const int half = warp >> 2;     // which half of the outputs / codes
  const int qr = lane >> 2, qc = lane & 3;

  // ---- phase A: E[f][n] = sum_c z[b][c][t0+f] * w_in[n][c]  (M = 64 frames, N = 96, K = 1024), 3xTF32
  float acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int cbase = 0; cbase < kRvqLatent; cbase += kRvqKc) {
    for (int e = tid; e < kRvqKc * kRvqFrames; e += 256) {  // z chunk, coalesced along t
      const int c = e / kRvqFrames, f = e % kRvqFrames;
      const int t = t0 + f;
      float val = 0.f;
      if (t < p.T) {
        const long long off = (static_cast<long long>(b) * kRvqLatent + cbase + c) * p.T + t;
        val = p.z_is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.z)[off]) : static_cast<const float*>(p.z)[off];
      }
      s_z[c * kRvqZP + f] = val;
    }
    for (int e = tid; e < kRvqKc * kRvqE; e += 256) {  // W chunk: rows of the transposed weight are contiguous
      const int c = e / kRvqE, n = e % kRvqE;
      s_w[c * kRvqWP + n] = __ldg(p.w_in_t + static_cast<long long>(cbase + c) * kRvqE + n);
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kRvqKc / 8; ++ks) {
      // A fragment (row = frame, col = k): a0 (qr, qc) a1 (qr+8, qc) a2 (qr, qc+4) a3 (qr+8, qc+4)
      uint32_t ah[4], al[4];
      const float* za = s_z + (ks * 8 + qc) * kRvqZP + mt * 16 + qr;
      tf32_split(za[0], ah[0], al[0]);
      tf32_split(za[8], ah[1], al[1]);
      tf32_split(za[4 * kRvqZP], ah[2], al[2]);
      tf32_split(za[4 * kRvqZP + 8], ah[3], al[3]);
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        // B fragment (k, n): b0 (qc, qr) b1 (qc+4, qr)
        const float* wb = s_w + (ks * 8 + qc) * kRvqWP + (half * 6 + nt) * 8 + qr;
        uint32_t bh0, bl0, bh1, bl1;
        tf32_split(wb[0], bh0, bl0);
        tf32_split(wb[4 * kRvqWP], bh1, bl1);
        mma_3xtf32(acc[nt], ah, al, bh0, bh1, bl0, bl1);
      }
    }
    __syncthreads();
  }
  // C fragment: c0 (qr, 2qc) c1 (qr, 2qc+1) c2 (qr+8, 2qc) c3 (qr+8, 2qc+1)
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    const int n = (half * 6 + nt) * 8 + 2 * qc;
    const float b0 = __ldg(p.b_in + n), b1 = __ldg(p.b_in + n + 1);
    float* e0 = s_e + (mt * 16 + qr) * kRvqEP + n;
    e0[0] = acc[nt][0] + b0;
    e0[1] = acc[nt][1] + b1;
    e0[8 * kRvqEP] = acc[nt][2] + b0;
    e0[8 * kRvqEP + 1] = acc[nt][3] + b1;
  }
  __syncthreads();

  // ---- phase B: 12 sequential levels in the 8-d space
  const int f = tid >> 2, sub = tid & 3;  // 4 threads per frame for the latent bookkeeping
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

    // latent e (components 2*sub, 2*sub+1 owned by this thread), normalised copy to smem
    float e0 = s_e[f * kRvqEP + lvl * 8 + 2 * sub];
    float e1 = s_e[f * kRvqEP + lvl * 8 + 2 * sub + 1];
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
    float en2 = e0 * e0 + e1 * e1;
    en2 += __shfl_xor_sync(0xffffffffu, en2, 1);
    en2 += __shfl_xor_sync(0xffffffffu, en2, 2);
    s_en[f * 8 + 2 * sub] = e0;
    s_en[f * 8 + 2 * sub + 1] = e1;
    if (sub == 0) s_en2[f] = en2;
    __syncthreads();  // codebook + normalised latents staged

    // distance tile of this warp: frames [16 mt, 16 mt + 16) x codes [512 half, 512 half + 512)
    uint32_t ah[4], al[4];
    {
      const float* ea = s_en + (mt * 16 + qr) * 8 + qc;
      tf32_split(ea[0], ah[0], al[0]);
      tf32_split(ea[64], ah[1], al[1]);
      tf32_split(ea[4], ah[2], al[2]);
      tf32_split(ea[68], ah[3], al[3]);
    }
    const float en2_r0 = s_en2[mt * 16 + qr], en2_r1 = s_en2[mt * 16 + qr + 8];
    float best0 = INFINITY, best1 = INFINITY;
    int idx0 = 0, idx1 = 0;
#pragma unroll 4
    for (int nt = 0; nt < 64; ++nt) {
      const int code0 = half * 512 + nt * 8;
      const float* cbp = s_cb + (code0 + qr) * 8 + qc;  // B fragment (k = qc / qc+4, n = qr)
      uint32_t bh0, bl0, bh1, bl1;
      tf32_split(cbp[0], bh0, bl0);
      tf32_split(cbp[4], bh1, bl1);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      mma_3xtf32(d, ah, al, bh0, bh1, bl0, bl1);
      const int ca = code0 + 2 * qc;
      const float na = s_n2[ca], nb = s_n2[ca + 1];
      // dist = (|e|^2 - 2 e.c) + |c|^2, first minimum wins (codes are visited in increasing order)
      const float d00 = fmaf(-2.0f, d[0], en2_r0) + na;
      const float d01 = fmaf(-2.0f, d[1], en2_r0) + nb;
      const float d10 = fmaf(-2.0f, d[2], en2_r1) + na;
      const float d11 = fmaf(-2.0f, d[3], en2_r1) + nb;
      if (d00 < best0) { best0 = d00; idx0 = ca; }
      if (d01 < best0) { best0 = d01; idx0 = ca + 1; }
      if (d10 < best1) { best1 = d10; idx1 = ca; }
      if (d11 < best1) { best1 = d11; idx1 = ca + 1; }
    }
    // the 4 lanes of a quad hold interleaved codes of the same two rows
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ob0 = __shfl_xor_sync(0xffffffffu, best0, o), ob1 = __shfl_xor_sync(0xffffffffu, best1, o);
      const int oi0 = __shfl_xor_sync(0xffffffffu, idx0, o), oi1 = __shfl_xor_sync(0xffffffffu, idx1, o);
      if (ob0 < best0 || (ob0 == best0 && oi0 < idx0)) { best0 = ob0; idx0 = oi0; }
      if (ob1 < best1 || (ob1 == best1 && oi1 < idx1)) { best1 = ob1; idx1 = oi1; }
    }
    if (qc == 0) {
      s_bd[half * kRvqFrames + mt * 16 + qr] = best0;
      s_bi[half * kRvqFrames + mt * 16 + qr] = idx0;
      s_bd[half * kRvqFrames + mt * 16 + qr + 8] = best1;
      s_bi[half * kRvqFrames + mt * 16 + qr + 8] = idx1;
    }
    __syncthreads();
    // combine the two code halves (half 0 holds the smaller indices, so it wins ties)
    const float bd0 = s_bd[f], bd1 = s_bd[kRvqFrames + f];
    const int best_idx = (bd1 < bd0) ? s_bi[kRvqFrames + f] : s_bi[f];
    const long long oidx = (static_cast<long long>(b) * p.n_levels + lvl) * p.T + t;
    if (live && sub == 0) p.codes[oidx] = best_idx;
    chosen[lvl] = (p.forced != nullptr && live) ? static_cast<int>(p.forced[oidx]) : best_idx;
    __syncthreads();  // before the next level overwrites the staged codebook / partial results
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
