// Persistent warp-specialised bf16 GEMM for sm_100a: C[M,N] = A[M,K] * B[N,K]^T (both operands K-major, which is
// how activations [tokens, channels] and nn.Linear weights [out, in] already sit in HBM).
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B swizzle, kStages-deep mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x 256 x 16 per instruction, fp32 accum in TMEM)
//   warps 2..9  : epilogue (tcgen05.ld -> registers -> fused op -> global), overlapped with the next tile's mainloop
//                 through two TMEM accumulator buffers (2 x 256 columns = all 512 columns). Warp w reads TMEM lane
//                 quadrant w % 4 and column half (w - 2) / 4; the next chunk's tcgen05.ld is in flight while the
//                 current one is processed.
//
// Fused epilogues cover every GEMM of the conformer block (reference: edm_tts/models/conformer/conformer.py:149-181,
// 113-146) and the logits heads (injection_conformer_wrapper.py:38-63); rounding points follow bf16 autocast:
// Linear/Conv outputs are rounded to bf16 before the activation / residual add, the residual stream stays fp32.
#pragma once
#include "ptx.cuh"

namespace edm {

enum EpiKind : int {
  EPI_BF16 = 0,        // out_bf16 = bf16(acc + bias)
  EPI_SWISH_BF16 = 1,  // h = bf16(acc + bias); out_bf16 = bf16(h * bf16(sigmoid(h)))          (FeedForward up-proj)
  EPI_QKV_ROPE = 2,    // cols < rope_cols: rotary embedding on bf16(acc) per 64-wide head; others bf16(acc)
  EPI_RESID_F32 = 3,   // x_f32[r, c] += scale * bf16(acc + bias)                                (residual branches)
  EPI_F32 = 4,         // out_f32 = acc + bias                                                   (logits heads)
};

struct GemmParams {
  int M, N, K;
  int a_k_offset;    // starting column inside A's tensor map (grouped per-codebook heads share one activation map)
  int b_row_offset;  // starting row inside B's tensor map (stacked per-codebook weights)
  const float* bias;  // [N] (already offset for the group) or nullptr
  void* out;          // bf16 / fp32, row-major with leading dimension ldo (elements); already offset for the group
  long long ldo;
  float scale;
  const float* rope_cos;  // [max_pos, 32] fp32
  const float* rope_sin;
  int seq_len;    // rotary position = row % seq_len
  int rope_cols;  // columns [0, rope_cols) get rotary (q and k parts of the fused QKV projection)
};

constexpr int kGemmBM = 128;
constexpr int kGemmBN = 256;
constexpr int kGemmBK = 64;
constexpr int kGemmStages = 4;
constexpr int kGemmEpiWarps = 8;  // two per TMEM lane quadrant, each owning half of the tile's columns
constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;
constexpr uint32_t kGemmABytes = kGemmBM * kGemmBK * 2;
constexpr uint32_t kGemmBBytes = kGemmBN * kGemmBK * 2;
constexpr uint32_t kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr uint32_t kGemmSmemBytes = kGemmStages * kGemmStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ float sigmoidf_fast(float v) { return 1.0f / (1.0f + __expf(-v)); }

template <int EPI>
__device__ __forceinline__ void gemm_epilogue_32(const GemmParams& p, int row, int col, const uint32_t (&r)[32]) {
  // r: 32 consecutive fp32 accumulators of `row`, columns [col, col + 32)
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if constexpr (EPI == EPI_F32) {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (EPI == EPI_RESID_F32) {
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 x = o[i];
      x.x += p.scale * bf16_round(v[4 * i + 0]);
      x.y += p.scale * bf16_round(v[4 * i + 1]);
      x.z += p.scale * bf16_round(v[4 * i + 2]);
      x.w += p.scale * bf16_round(v[4 * i + 3]);
      o[i] = x;
    }
  } else {
    if constexpr (EPI == EPI_SWISH_BF16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t h2 = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        const uint32_t s2 = pack_bf16x2(sigmoid_tanh(bf16lo(h2)), sigmoid_tanh(bf16hi(h2)));
        const uint32_t o2 = bf16x2_mul(h2, s2);
        v[2 * i] = bf16lo(o2);
        v[2 * i + 1] = bf16hi(o2);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 w;
      w.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      o[i] = w;
    }
  }
}

// One 64-wide head: lo = columns [col, col+32), hi = [col+32, col+64). Reference: conformer.py:45-51 (rotate_half),
// applied to the bf16 projection output in fp32 and rounded to bf16 when SDPA consumes it.
__device__ __forceinline__ void gemm_epilogue_rope64(const GemmParams& p, int row, int col, const uint32_t (&lo)[32],
                                                     const uint32_t (&hi)[32]) {
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + col;
  uint32_t w[32];
  if (col < p.rope_cols) {
    const int pos = row % p.seq_len;
    const float4* c4 = reinterpret_cast<const float4*>(p.rope_cos + static_cast<long long>(pos) * 32);
    const float4* s4 = reinterpret_cast<const float4*>(p.rope_sin + static_cast<long long>(pos) * 32);
    float o_lo[32], o_hi[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 c = __ldg(c4 + i), s = __ldg(s4 + i);
      const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x1 = bf16_round(__uint_as_float(lo[4 * i + j]));
        float x2 = bf16_round(__uint_as_float(hi[4 * i + j]));
        o_lo[4 * i + j] = __fadd_rn(__fmul_rn(x1, cc[j]), __fmul_rn(-x2, ss[j]));
        o_hi[4 * i + j] = __fadd_rn(__fmul_rn(x2, cc[j]), __fmul_rn(x1, ss[j]));
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      w[i] = pack_bf16x2(o_lo[2 * i], o_lo[2 * i + 1]);
      w[16 + i] = pack_bf16x2(o_hi[2 * i], o_hi[2 * i + 1]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      w[i] = pack_bf16x2(__uint_as_float(lo[2 * i]), __uint_as_float(lo[2 * i + 1]));
      w[16 + i] = pack_bf16x2(__uint_as_float(hi[2 * i]), __uint_as_float(hi[2 * i + 1]));
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);
  uint64_t* empty_bar = full_bar + kGemmStages;
  uint64_t* tmem_full_bar = empty_bar + kGemmStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + kGemmBM - 1) / kGemmBM;
  const int num_n = p.N / kGemmBN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / kGemmBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 32 * kGemmEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kGemmStageBytes;
          uint8_t* sb = sa + kGemmABytes;
          mbar_arrive_expect_tx(&full_bar[stage], kGemmStageBytes);
          tma_load_2d(&tma_a, &full_bar[stage], sa, p.a_k_offset + kb * kGemmBK, m_blk * kGemmBM);
          tma_load_2d(&tma_b, &full_bar[stage], sb, kb * kGemmBK, p.b_row_offset + n_blk * kGemmBN);
          if (++stage == kGemmStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, kGemmBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kGemmBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kGemmStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sa + kGemmABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // +32 B per 16-element K step inside the 128 B swizzle atom (start-address field is in 16 B units)
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == kGemmStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;   // which 128 of the tile's 256 columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / num_n, n_blk = t % num_n;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * kGemmBM + quad * 32 + lane;
      const int col0 = n_blk * kGemmBN + half * 128;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kGemmBN + half * 128;
      if constexpr (EPI == EPI_QKV_ROPE) {
        uint32_t lo[2][32], hi[2][32];
        tmem_ld_32x32(taddr, lo[0]);
        tmem_ld_32x32(taddr + 32, hi[0]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld_wait_dep(lo[c]);
          tmem_ld_wait_dep(hi[c]);
          if (c == 0) {
            tmem_ld_32x32(taddr + 64, lo[1]);
            tmem_ld_32x32(taddr + 96, hi[1]);
          }
          if (row < p.M) gemm_epilogue_rope64(p, row, col0 + c * 64, lo[c], hi[c]);
        }
      } else {
        uint32_t r[2][32];
        tmem_ld_32x32(taddr, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait_dep(r[c & 1]);
          if (c + 1 < 4) tmem_ld_32x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
          if (row < p.M) gemm_epilogue_32<EPI>(p, row, col0 + c * 32, r[c & 1]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace edm
