// Non-causal multi-head attention for the conformer block (reference: edm_tts/models/conformer/attend.py:63-115 via
// conformer.py:128-146): softmax(q k^T / sqrt(64)) v, H heads of 64, no mask, one sequence = one batch element.
//
// One CTA = 128 query rows of one (batch, head). Both contractions run on tcgen05 with fp32 accumulators in TMEM:
//   S = Q K^T : A = Q  [128 x 64]  K-major smem,  B = K tile [128 x 64] K-major smem   -> TMEM cols [0,128)
//   O = P V   : A = P  [128 x 128] K-major smem (written by the softmax warps as bf16),
//               B = V tile [128 kv x 64] which TMA delivers with the head dim contiguous = MN-major B operand
//                                                                                        -> TMEM cols [128,192)
//   warps 0..3 : softmax, one query row per thread (TMEM lane == row, so no shuffles): two passes over S in
//                32-column tcgen05.ld chunks (row max, then exp2 + bf16 P into 128B-swizzled smem); the running
//                output lives in registers and is rescaled there, each P V product is read back from TMEM and added.
//   warp 4     : TMA producer (Q once, then a 2-deep K/V ring)
//   warp 5     : TMEM allocator + single-thread MMA issuer
// 112 KB smem + 256 TMEM columns per CTA -> two CTAs per SM overlap each other's softmax and MMA phases.
#pragma once
#include "ptx.cuh"

namespace edm {

struct AttnParams {
  int B, N, H;            // sequences, tokens per sequence, heads (head dim fixed at 64)
  int q_col0, k_col0, v_col0;  // first column of q / k / v inside the fused projection buffer
  __nv_bfloat16* out;     // [B*N, H*64]
  long long ldo;
  float scale_log2e;      // softmax scale * log2(e)
  // MN-major V descriptor knobs (bytes); defaults 1024 / 1024 / 2048, exposed so a bring-up run can sweep them.
  uint32_t v_lbo, v_sbo, v_kstep;
};

constexpr int kAttnThreads = 192;
constexpr uint32_t kAttnTile = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr uint32_t kAttnSmemBytes = kAttnTile /*Q*/ + 2 * kAttnTile /*K*/ + 2 * kAttnTile /*V*/ + 2 * kAttnTile /*P*/ + 256;

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTile;
  uint8_t* sV = sK + 2 * kAttnTile;
  uint8_t* sP = sV + 2 * kAttnTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttnTile);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.N + 127) / 128;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("edm: attention smem base not 1024-aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAttnTile);
      tma_load_3d(&tma_qkv, q_full, sQ, p.q_col0 + h * 64, qt * 128, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * kAttnTile);
        tma_load_3d(&tma_qkv, &kv_full[s], sK + s * kAttnTile, p.k_col0 + h * 64, j * 128, b);
        tma_load_3d(&tma_qkv, &kv_full[s], sV + s * kAttnTile, p.v_col0 + h * 64, j * 128, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      mbar_wait(q_full, 0);
      // S_0
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        // P_j is in smem, S_j and O_{j-1} have been drained from TMEM
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t pdesc0 = umma_desc_sw128(smem_u32(sP), 16, 1024);
        const uint64_t pdesc1 = umma_desc_sw128(smem_u32(sP + kAttnTile), 16, 1024);
        const uint32_t v_addr = smem_u32(sV + s * kAttnTile);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t adesc = (k < 4 ? pdesc0 : pdesc1) + 2 * (k & 3);
          const uint64_t bdesc = umma_desc_sw128(v_addr + k * p.v_kstep, p.v_lbo, p.v_sbo);
          umma_ss(tmem_O, adesc, bdesc, idesc_o, k != 0);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
        if (j + 1 < n_kv) {
          const int s1 = (j + 1) & 1;
          mbar_wait(&kv_full[s1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + s1 * kAttnTile), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
      }
    }
  } else {
    // ---- softmax: thread <-> query row
    const int row_in_tile = warp * 32 + lane;
    const int q_row = qt * 128 + row_in_tile;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    const float sl2 = p.scale_log2e;
    uint8_t* p_row = sP + row_in_tile * 128;
    const int sw = row_in_tile & 7;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = p.N - j * 128;  // columns >= kv_valid are padding
      float mx = m_run;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_S + lane_off + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[i]);
          if (c * 32 + i < kv_valid) mx = fmaxf(mx, s);
        }
      }
      const float alpha = (m_run == -INFINITY) ? 0.f : exp2f((m_run - mx) * sl2);
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_O + lane_off + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(r[i])) * alpha;
        }
      }
      l_run *= alpha;
      const float mb = mx * sl2;
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_S + lane_off + c * 32, r);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = (c * 32 + 2 * i < kv_valid) ? exp2f(fmaf(__uint_as_float(r[2 * i]), sl2, -mb)) : 0.f;
          float p1 = (c * 32 + 2 * i + 1 < kv_valid) ? exp2f(fmaf(__uint_as_float(r[2 * i + 1]), sl2, -mb)) : 0.f;
          lsum += p0 + p1;
          w[i] = pack_bf16x2(p0, p1);
        }
        // 32 kv columns = 64 B = four 16 B chunks of this row; 128B swizzle: chunk index ^= (row & 7)
        uint8_t* half = p_row + (c >> 1) * kAttnTile;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ sw;
          *reinterpret_cast<uint4*>(half + chunk * 16) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
      }
      l_run += lsum;
      m_run = mx;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }
    // last P V product
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_O + lane_off + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(r[i])) * inv_l;
    }
    if (q_row < p.N) {
      uint4* o = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * p.N + q_row) * p.ldo + h * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = make_uint4(pack_bf16x2(o_acc[8 * i + 0], o_acc[8 * i + 1]), pack_bf16x2(o_acc[8 * i + 2], o_acc[8 * i + 3]),
                          pack_bf16x2(o_acc[8 * i + 4], o_acc[8 * i + 5]), pack_bf16x2(o_acc[8 * i + 6], o_acc[8 * i + 7]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem_base);
}

}  // namespace edm
