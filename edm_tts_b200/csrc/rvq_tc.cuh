// DAC ResidualVectorQuantize nearest-code search on the 5th-generation tensor cores (tcgen05, kind::tf32, TMEM accumulators).
// Reference: edm_tts/models/dac/vector_quantizer.py:146-210 (residual loop), :33-67 + :75-91 (per-level search).
//
// Same algebra as rvq.cuh (z is read once, the level loop runs on 8-d latents with the G cross tables), split in two kernels:
//
//   rvq_project_kernel : E[frame, 96] = z[b, :, t] . W_in^T + b_in for all 12 levels at once (256-frame CTA tiles: two 128-row
//                        accumulators share every weight chunk). z [B, 1024, T] is consumed as an
//                        MN-major (frames contiguous) tf32 A operand straight from TMA boxes (128B swizzle with 32 B atoms), so the
//                        channel-major layout of the reference needs no transpose. fp32-level accuracy comes from the 3xTF32
//                        split: four "split" warps rewrite every staged z tile as hi = tf32(z) and lo = tf32(z - hi) in place
//                        (elementwise, so they never need to know the swizzle), the weights are pre-split at pack time, and the
//                        MMA warp accumulates hi*hi + lo*hi + hi*lo into one TMEM accumulator. bf16 z (the reference's autocast
//                        case) lands as bf16 staging boxes and is expanded on chip; being tf32-exact it needs no lo part.
//   rvq_search_kernel  : 128 frames per CTA, two scan warps per TMEM lane quadrant (each half of the columns), the first of them
//                        also owning its frame's bookkeeping. Per level: subtract the G rows of the codes chosen so far,
//                        L2-normalise, write the 3xTF32-split latent row as a K-major A operand (128B swizzle) to shared memory;
//                        the MMA thread multiplies it with the packed codebook [c_hi | c_hi | c_lo | -|c|^2/2] streamed in
//                        128-code chunks by TMA (K = 32 per code), scores land in TMEM (2 x 128 columns, double-buffered) and
//                        each thread scans its own row with tcgen05.ld for the first maximum (groups of eight: FMNMX3 tree on the
//                        ALU pipe, the remembered group's scores saved by predicated FMUL on the FMA pipe). Two CTAs per SM
//                        overlap one tile's level-boundary latency with the other's scan. Both kernels issue their MMAs from
//                        warp-uniform code with one elected lane (see ptx.cuh: umma_ss_warp).
//
//   score(code) = e^ . c^ - |c^|^2 / 2  = -(dist - |e^|^2) / 2   with dist = |e^|^2 - 2 e^ . c^ + |c^|^2 (the reference formula),
//   so arg-max score (first maximum) = the reference's argmax(-dist).
#pragma once
#include "ptx.cuh"

namespace edm {

constexpr int kRtFrames = 128;  // frames per tile = UMMA M
constexpr int kRtE = 96;        // 12 levels x 8 dims
constexpr int kRtLatent = 1024;
constexpr int kRtCodes = 1024;   // codes per codebook
constexpr int kRvqLevels = 12;   // codebooks

// ------------------------------------------------------------------------------------------------ projection
// Two rings: one for the raw z tiles (the only HBM stream: 3 x 32 KB in flight per SM keeps the memory system busy
// across the ~1 us load latency) and a shallow one for what is needed only between "split" and "MMA": the lo tile and the
// (L2-resident) hi / lo weight chunks.
// A CTA tile is 256 frames (two 128-row accumulators): every weight chunk fetched from L2 feeds two MMAs, which halves the
// L2 -> SM weight traffic (with 128-frame tiles it was 1.5x the z stream and the L2 fabric, not HBM, set the pace).
constexpr int kRpFrames = 256;
constexpr int kRpKc = 32;                                            // latent channels per pipeline step
constexpr int kRpZStages = 3;        // fp32 z: 3 x 32 KB; bf16 z: 6 x 16 KB staging tiles (same 96 KB)
constexpr int kRpLwStages = 2;
constexpr uint32_t kRpZBytes = kRpFrames * kRpKc * 4;                // 32 KB: 8 TMA boxes of (32 frames x 32 channels)
constexpr uint32_t kRpWBytes = kRtE * kRpKc * 4;                     // 12 KB: 96 outputs x 32 channels (K-major)
constexpr uint32_t kRpLwBytes = kRpZBytes + 2 * kRpWBytes;           // z_lo | w_hi | w_lo = 56 KB
constexpr uint32_t kRpSmemBytes = kRpZStages * kRpZBytes + kRpLwStages * kRpLwBytes + 1024 + 256;
constexpr int kRpThreads = 352;  // warp 0 z TMA, warp 1 MMA, warps 2-5 split, warps 6-9 epilogue, warp 10 weight TMA

struct RvqProjParams {
  int B, T;
  const float* b_in;  // [96]
  float* e_out;       // [B*T, 96] projected latents (before any residual correction)
  uint32_t a_lbo, a_sbo;  // MN-major A descriptor strides (bytes): 32-frame atoms (TMA boxes) 4096 apart, 4-channel K-groups 512 apart
  int dbg;                // bring-up timing probes: bit 0 = split warps skip their arithmetic, bit 1 = no MMAs are issued
};

// kBf16: z is bf16 (what the DAC encoder produces under autocast). Its tiles land as bf16 staging boxes (6 x 16 KB ring), the
// split warps expand them into the fp32 hi tile of the (hi | weights) slot in the MN-major layout the MMA expects; bf16 values
// are tf32-exact, so the lo tile and its MMA disappear and z costs half the HBM bytes.
template <bool kBf16>
__global__ void __launch_bounds__(kRpThreads, 1)
rvq_project_kernel(const __grid_constant__ CUtensorMap tma_z, const __grid_constant__ CUtensorMap tma_whi,
                   const __grid_constant__ CUtensorMap tma_wlo, const RvqProjParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int kZSt = kBf16 ? 2 * kRpZStages : kRpZStages;
  constexpr uint32_t kZB = kBf16 ? kRpZBytes / 2 : kRpZBytes;   // bytes per z ring slot
  uint8_t* sZ = smem;                                   // fp32: 3 x 32 KB (becomes the hi tile in place); bf16: 6 x 16 KB staging
  uint8_t* sLw = smem + kRpZStages * kRpZBytes;         // kRpLwStages x (lo tile [bf16: hi tile] | w_hi | w_lo)
  uint64_t* z_full = reinterpret_cast<uint64_t*>(sLw + kRpLwStages * kRpLwBytes);
  uint64_t* z_empty = z_full + 2 * kRpZStages;
  uint64_t* w_full = z_empty + 2 * kRpZStages;
  uint64_t* split_bar = w_full + kRpLwStages;
  uint64_t* lw_empty = split_bar + kRpLwStages;
  uint64_t* tfull_bar = lw_empty + kRpLwStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_b = (p.T + kRpFrames - 1) / kRpFrames;
  const int num_tiles = tiles_per_b * p.B;
  constexpr int kNumKc = kRtLatent / kRpKc;
  const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const uint32_t total_steps = static_cast<uint32_t>(my_tiles) * kNumKc;  // pipeline steps of this CTA, numbered it

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_z);
    tma_prefetch_desc(&tma_whi);
    tma_prefetch_desc(&tma_wlo);
    for (int s = 0; s < kZSt; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], kBf16 ? 128 : 1);  // released by the split warps (bf16 staging) or by the MMA commit (fp32, in place)
    }
    for (int s = 0; s < kRpLwStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&split_bar[s], 128);
      mbar_init(&lw_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (uint32_t it = 0; it < total_steps; ++it) {
        const int tile = blockIdx.x + (it / kNumKc) * gridDim.x, kc = it % kNumKc;
        const int b = tile / tiles_per_b, t0 = (tile % tiles_per_b) * kRpFrames;
        const uint32_t zs = it % kZSt;
        mbar_wait(&z_empty[zs], ((it / kZSt) & 1) ^ 1);
        mbar_arrive_expect_tx(&z_full[zs], kZB);
        if constexpr (kBf16) {
#pragma unroll
          for (int i = 0; i < kRpFrames / 64; ++i)  // bf16 boxes: 64 frames (128 B) x 32 channels
            tma_load_2d(&tma_z, &z_full[zs], sZ + zs * kZB + i * 4096, t0 + 64 * i, b * kRtLatent + kc * kRpKc);
        } else {
#pragma unroll
          for (int i = 0; i < kRpFrames / 32; ++i)  // frames beyond T are zero-filled by TMA
            tma_load_2d(&tma_z, &z_full[zs], sZ + zs * kZB + i * 4096, t0 + 32 * i, b * kRtLatent + kc * kRpKc);
        }
      }
    }
  } else if (warp == 10) {
    if (lane == 0) {
      for (uint32_t it = 0; it < total_steps; ++it) {
        const int kc = it % kNumKc;
        const uint32_t ls = it % kRpLwStages;
        mbar_wait(&lw_empty[ls], ((it / kRpLwStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[ls], 2 * kRpWBytes);
        tma_load_2d(&tma_whi, &w_full[ls], sLw + ls * kRpLwBytes + kRpZBytes, kc * kRpKc, 0);
        tma_load_2d(&tma_wlo, &w_full[ls], sLw + ls * kRpLwBytes + kRpZBytes + kRpWBytes, kc * kRpKc, 0);
      }
    }
  } else if (warp == 1) {
    {  // whole warp, warp-uniform control flow; one elected lane issues (umma_*_warp, see ptx.cuh)
      constexpr uint32_t idesc = umma_idesc_tf32(kRtFrames, kRtE, /*a_mn_major=*/1, 0);
      for (uint32_t it = 0; it < total_steps; ++it) {
        const uint32_t tl = it / kNumKc, kc = it % kNumKc;
        const uint32_t acc = tl & 1;
        const uint32_t zs = it % kZSt, ls = it % kRpLwStages;
        if (kc == 0) {
          mbar_wait(&tempty_bar[acc], ((tl >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + acc * 256;
        mbar_wait(&w_full[ls], (it / kRpLwStages) & 1);     // weights (TMA) landed
        mbar_wait(&split_bar[ls], (it / kRpLwStages) & 1);  // z tile rewritten as hi (in place) / lo by the split warps
        tc_fence_after();
        const uint32_t lw = smem_u32(sLw + ls * kRpLwBytes);
        const uint32_t zh = kBf16 ? lw : smem_u32(sZ + zs * kZB);   // fp32 hi tile: expanded into the slot (bf16) or in place
#pragma unroll
        for (int k = 0; k < ((p.dbg & 2) ? 0 : kRpKc / 8); ++k) {
          // A (MN-major, 128B swizzle with 32 B atoms): 8 channels = two 512 B K-groups inside each 32-frame box (4 KB); the
          // four boxes of one 128-frame accumulator are 4 KB apart. B (K-major): +32 B per 8-element K step inside the 128 B
          // swizzle row.
          const uint64_t b_hi = umma_desc_sw128(lw + kRpZBytes, 16, 1024) + 2 * k;
          const uint64_t b_lo = umma_desc_sw128(lw + kRpZBytes + kRpWBytes, 16, 1024) + 2 * k;
#pragma unroll
          for (int m = 0; m < kRpFrames / 128; ++m) {
            const uint64_t a_hi = umma_desc_sw128_base32(zh + m * 16384 + k * 1024, p.a_lbo, p.a_sbo);
            if constexpr (kBf16) {
              umma_ss_tf32_warp(d_tmem + m * 128, a_hi, b_lo, idesc, (kc | k) != 0 ? 1u : 0u);
            } else {
              const uint64_t a_lo = umma_desc_sw128_base32(lw + m * 16384 + k * 1024, p.a_lbo, p.a_sbo);
              umma_ss_tf32_warp(d_tmem + m * 128, a_lo, b_hi, idesc, (kc | k) != 0 ? 1u : 0u);
              umma_ss_tf32_warp(d_tmem + m * 128, a_hi, b_lo, idesc, 1u);
            }
            umma_ss_tf32_warp(d_tmem + m * 128, a_hi, b_hi, idesc, 1u);
          }
        }
        if constexpr (!kBf16) umma_commit_warp(&z_empty[zs]);
        umma_commit_warp(&lw_empty[ls]);
        if (kc == kNumKc - 1) umma_commit_warp(&tfull_bar[acc]);
      }
    }
  } else if (warp < 6) {
    // split warps. fp32 z: tile (as TMA wrote it) -> hi = tf32(z) in place, lo = tf32(z - hi) at the same offset of the lo tile.
    // bf16 z: staging boxes [32 channels][64 frames] (128B swizzle: 16 B chunk k of row c sits at chunk k ^ (c & 7)) -> fp32 hi
    // tile of the slot, 32-frame boxes with 32 B chunks XOR-ed by (c & 3) (the MN-major tf32 layout).
    const int st_tid = threadIdx.x - 64;
    for (uint32_t it = 0; it < total_steps; ++it) {
      const uint32_t zs = it % kZSt, ls = it % kRpLwStages;
      mbar_wait(&z_full[zs], (it / kZSt) & 1);
      mbar_wait(&lw_empty[ls], ((it / kRpLwStages) & 1) ^ 1);
      if constexpr (kBf16) {
        const uint32_t src = smem_u32(sZ + zs * kZB), dst = smem_u32(sLw + ls * kRpLwBytes);
#pragma unroll
        for (int i = 0; i < ((p.dbg & 1) ? 0 : static_cast<int>(kZB) / (128 * 16)); ++i) {
          const int q = st_tid + 128 * i;          // physical 16 B chunk of the staging tile
          const int box = q >> 8, c = (q & 255) >> 3, k = (q & 7) ^ (c & 7);   // frames 64 box + 8 k .. + 7 of channel row c
          const float4 v = lds128(src + q * 16);   // 8 bf16
          const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
          const uint32_t o = dst + (2 * box + (k >> 2)) * 4096 + c * 128 + (((k & 3) ^ (c & 3)) << 5);
          sts128(o, make_float4(__uint_as_float(w[0] << 16), __uint_as_float(w[0] & 0xffff0000u), __uint_as_float(w[1] << 16),
                                __uint_as_float(w[1] & 0xffff0000u)));
          sts128(o + 16, make_float4(__uint_as_float(w[2] << 16), __uint_as_float(w[2] & 0xffff0000u), __uint_as_float(w[3] << 16),
                                     __uint_as_float(w[3] & 0xffff0000u)));
        }
        fence_proxy_async_smem();
        mbar_arrive(&z_empty[zs]);   // the staging slot can be refilled
        mbar_arrive(&split_bar[ls]);
      } else {
        const uint32_t zh = smem_u32(sZ + zs * kZB) + st_tid * 16;
        const uint32_t zl = smem_u32(sLw + ls * kRpLwBytes) + st_tid * 16;
#pragma unroll
        for (int i = 0; i < ((p.dbg & 1) ? 0 : static_cast<int>(kRpZBytes) / (128 * 16)); ++i) {
          const float4 v = lds128(zh + i * 2048);
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
          sts128(zh + i * 2048, h);
          sts128(zl + i * 2048, l);
        }
        fence_proxy_async_smem();
        mbar_arrive(&split_bar[ls]);
      }
    }
  } else if (warp < 10) {
    // epilogue warps: thread <-> TMEM lane = frame of each 128-row accumulator, 96 columns + bias -> E[frame, 96]
    const int quad = warp & 3;
    const int f = quad * 32 + lane;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int tile = blockIdx.x + tl * gridDim.x;
      const int acc = tl & 1;
      const int b = tile / tiles_per_b;
      mbar_wait(&tfull_bar[acc], (tl >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < kRpFrames / 128; ++m) {
        const int t = (tile % tiles_per_b) * kRpFrames + m * 128 + f;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + m * 128;
        uint32_t r0[32], r1[32], r2[32];
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + 32, r1);
        tmem_ld_32x32(taddr + 64, r2);
        tmem_ld_wait_dep(r0);
        tmem_ld_wait_dep(r1);
        tmem_ld_wait_dep(r2);
        if (m == kRpFrames / 128 - 1) {
          tc_fence_before();
          mbar_arrive(&tempty_bar[acc]);
        }
        if (t < p.T) {
          float4* o = reinterpret_cast<float4*>(p.e_out + (static_cast<long long>(b) * p.T + t) * kRtE);
          const float4* b4 = reinterpret_cast<const float4*>(p.b_in);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i);
            o[i] = make_float4(__uint_as_float(r0[4 * i]) + bb.x, __uint_as_float(r0[4 * i + 1]) + bb.y,
                               __uint_as_float(r0[4 * i + 2]) + bb.z, __uint_as_float(r0[4 * i + 3]) + bb.w);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + 8 + i);
            o[8 + i] = make_float4(__uint_as_float(r1[4 * i]) + bb.x, __uint_as_float(r1[4 * i + 1]) + bb.y,
                                   __uint_as_float(r1[4 * i + 2]) + bb.z, __uint_as_float(r1[4 * i + 3]) + bb.w);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + 16 + i);
            o[16 + i] = make_float4(__uint_as_float(r2[4 * i]) + bb.x, __uint_as_float(r2[4 * i + 1]) + bb.y,
                                    __uint_as_float(r2[4 * i + 2]) + bb.z, __uint_as_float(r2[4 * i + 3]) + bb.w);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ level search
constexpr int kRsChunk = 128;                                  // codes per MMA (UMMA N) and per TMA box
constexpr int kRsChunks = 1024 / kRsChunk;
constexpr int kRsRing = 4;
constexpr uint32_t kRsTileBytes = 128 * 128;                   // 128 rows x 32 fp32 (A tile and every codebook chunk)
constexpr uint32_t kRsSmemBytes = kRsTileBytes * (1 + kRsRing) + (12 + 8 + 2) * 128 * 4 + 1024 + 256;
constexpr int kRsThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 search + level bookkeeping, warps 6-9 search only

struct RvqSearchParams {
  int B, T, n_levels;
  const float* e;           // [B*T, 96] from rvq_project_kernel
  const float* g;           // [12, 12, 1024, 8]  G[i][j][code] (j < i used)
  long long* codes;         // out [B, n_levels, T]
  const long long* forced;  // teacher forcing [B, n_levels, T] or nullptr
  float* latents;           // out [B, 96, T] residual-corrected latents before normalisation, or nullptr
  float one;                // 1.0f and 1, passed at run time so the saves of rvq_scan32 stay FMUL / IMAD (FMA pipe)
  int onei;
  int n_mma;                // MMAs per codebook chunk: 4 (K = 32: the 3xTF32 products and the norm term); fewer = bring-up timing probe only
};

// Running first maximum of one frame's scores. The scan is ALU-pipe bound (FSETP / FMNMX / SEL all issue there at half
// rate), so it works on groups of eight: four FMNMX3 / FMNMX for the group maximum, one FSETP against the running best, one
// FMNMX to update it (0.75 ALU instructions per score), and the "remember this group" side (group id + its eight raw scores)
// as predicated FMUL / IMAD by a run-time 1, which issue on the otherwise idle FMA pipe. The winner inside the
// remembered group is resolved once per level.
struct RvqBest {
  float best;
  int gid;       // group (code / 8) holding the running maximum
  float s[8];    // its eight scores
};
template <int kGroup0>
__device__ __forceinline__ void rvq_scan32(const uint32_t (&r)[32], RvqBest& st, int group_base, float one, int onei) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float v0 = __uint_as_float(r[8 * g]), v1 = __uint_as_float(r[8 * g + 1]), v2 = __uint_as_float(r[8 * g + 2]),
                v3 = __uint_as_float(r[8 * g + 3]), v4 = __uint_as_float(r[8 * g + 4]), v5 = __uint_as_float(r[8 * g + 5]),
                v6 = __uint_as_float(r[8 * g + 6]), v7 = __uint_as_float(r[8 * g + 7]);
    const float m = fmaxf(fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)), fmaxf(fmaxf(v4, v5), fmaxf(v6, v7)));
    asm("{\n\t"
        ".reg .pred q;\n\t"
        "setp.gt.f32 q, %10, %0;\n\t"           // strict: an equal later group never replaces an earlier one
        "max.f32 %0, %0, %10;\n\t"
        "@q mul.f32 %1, %11, %19;\n\t"
        "@q mul.f32 %2, %12, %19;\n\t"
        "@q mul.f32 %3, %13, %19;\n\t"
        "@q mul.f32 %4, %14, %19;\n\t"
        "@q mul.f32 %5, %15, %19;\n\t"
        "@q mul.f32 %6, %16, %19;\n\t"
        "@q mul.f32 %7, %17, %19;\n\t"
        "@q mul.f32 %8, %18, %19;\n\t"
        "@q mad.lo.s32 %9, %20, %21, %22;\n\t"
        "}\n"
        : "+f"(st.best), "+f"(st.s[0]), "+f"(st.s[1]), "+f"(st.s[2]), "+f"(st.s[3]), "+f"(st.s[4]), "+f"(st.s[5]), "+f"(st.s[6]),
          "+f"(st.s[7]), "+r"(st.gid)
        : "f"(m), "f"(v0), "f"(v1), "f"(v2), "f"(v3), "f"(v4), "f"(v5), "f"(v6), "f"(v7), "f"(one), "r"(onei), "r"(kGroup0 + g),
          "r"(group_base));
  }
}
// bring-up probe: consumes the loaded registers with one op per 32 scores (measures the TMEM-read / MMA floor of the kernel)
__device__ __forceinline__ void rvq_scan32_probe(const uint32_t (&r)[32], RvqBest& st) {
  st.best = fmaxf(st.best, __uint_as_float(r[0] ^ r[31]));
}

template <int kScan>  // 0 = full scan, 1 = bring-up probe (loads only), 2 = bring-up probe (no loads either)
__global__ void __launch_bounds__(kRsThreads, 2)
rvq_search_kernel(const __grid_constant__ CUtensorMap tma_cb, const RvqSearchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sRing = smem + kRsTileBytes;
  int* s_chosen = reinterpret_cast<int*>(sRing + kRsRing * kRsTileBytes);  // [12][128]
  float* s_nxt = reinterpret_cast<float*>(s_chosen + 12 * 128);            // [8][128] next level's partial latent (parked during the scan)
  float* s_pbest = s_nxt + 8 * 128;                                        // [128] best score of the upper column half (warps 6-9)
  int* s_pidx = reinterpret_cast<int*>(s_pbest + 128);                     // [128] and its code
  uint64_t* ring_full = reinterpret_cast<uint64_t*>(s_pidx + 128);
  uint64_t* ring_empty = ring_full + kRsRing;
  uint64_t* tfull_bar = ring_empty + kRsRing;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* a_ready = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_b = (p.T + kRtFrames - 1) / kRtFrames;
  const int num_tiles = tiles_per_b * p.B;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_cb);
    for (int s = 0; s < kRsRing; ++s) {
      mbar_init(&ring_full[s], 1);
      mbar_init(&ring_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 256);  // both column halves have pulled their part of the buffer
    }
    mbar_init(a_ready, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int l = 0; l < p.n_levels; ++l)
          for (int ch = 0; ch < kRsChunks; ++ch, ++g) {
            const uint32_t s = g % kRsRing;
            mbar_wait(&ring_empty[s], ((g / kRsRing) & 1) ^ 1);
            mbar_arrive_expect_tx(&ring_full[s], kRsTileBytes);
            tma_load_2d(&tma_cb, &ring_full[s], sRing + s * kRsTileBytes, 0, l * 1024 + ch * kRsChunk);
          }
    }
  } else if (warp == 1) {
    {  // whole warp, warp-uniform control flow; one elected lane issues (umma_*_warp, see ptx.cuh)
      constexpr uint32_t idesc = umma_idesc_tf32(kRtFrames, kRsChunk, 0, 0);
      const uint64_t adesc = umma_desc_sw128(smem_u32(sA), 16, 1024);
      uint32_t g = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int l = 0; l < p.n_levels; ++l, ++it) {
          mbar_wait(a_ready, it & 1);  // the 128 latent rows of this level are in shared memory
          for (int ch = 0; ch < kRsChunks; ++ch, ++g) {
            const uint32_t s = g % kRsRing, buf = g & 1;
            mbar_wait(&ring_full[s], (g / kRsRing) & 1);
            mbar_wait(&tempty_bar[buf], ((g >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sRing + s * kRsTileBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < p.n_mma) umma_ss_tf32_warp(tmem_base + buf * kRsChunk, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0 ? 1u : 0u);
            umma_commit_warp(&ring_empty[s]);
            umma_commit_warp(&tfull_bar[buf]);
          }
        }
    }
  } else {
    // Two warps per TMEM lane quadrant: warps 2-5 scan columns [0, 64) of every 128-code chunk and do the per-level
    // bookkeeping of their frame, warps 6-9 scan columns [64, 128) and hand (best, code) over through shared memory
    // (named barrier 1 + quad, 64 threads). Four warps per SM sub-partition keep TMEM loads and compares overlapped.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int f = quad * 32 + lane;
    const uint32_t a_row = smem_u32(sA) + f * 128;
    const int sw = f & 7;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * 64;
    uint32_t g = 0;
    // scans this warp's 64 columns of the 8 chunks of one level -> first maximum (score, code)
    auto scan_level = [&](float& best_out, int& idx_out) {
      RvqBest st;
      st.best = -INFINITY;
      st.gid = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) st.s[i] = 0.f;
      for (int ch = 0; ch < kRsChunks; ++ch, ++g) {
        const uint32_t buf = g & 1;
        const uint32_t taddr = tmem_row + buf * kRsChunk;
        const int group_base = ch * (kRsChunk / 8) + half * 8;
        mbar_wait(&tfull_bar[buf], (g >> 1) & 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        if constexpr (kScan != 2) {
          tmem_ld_32x32(taddr, r0);
          tmem_ld_32x32(taddr + 32, r1);
          tmem_ld_wait_dep(r0);
          tmem_ld_wait_dep(r1);
        } else {  // bring-up probe: no TMEM reads at all (what the MMA / barrier pipeline alone costs)
#pragma unroll
          for (int i = 0; i < 32; ++i) r0[i] = r1[i] = taddr + i;
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[buf]);  // the accumulator buffer may be overwritten by chunk g + 2
        if constexpr (kScan == 0) {
          rvq_scan32<0>(r0, st, group_base, p.one, p.onei);
          rvq_scan32<4>(r1, st, group_base, p.one, p.onei);
        } else {
          rvq_scan32_probe(r0, st);
          rvq_scan32_probe(r1, st);
        }
      }
      // first position of the maximum inside the remembered group
      int pos = 7;
#pragma unroll
      for (int i = 6; i >= 0; --i) pos = st.s[i] == st.best ? i : pos;
      best_out = st.best;
      idx_out = st.gid * 8 + pos;
    };

    if (half == 1) {
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int l = 0; l < p.n_levels; ++l) {
          float best;
          int bidx;
          scan_level(best, bidx);
          s_pbest[f] = best;
          s_pidx[f] = bidx;
          switch (quad) {  // immediate barrier ids keep the CTA at 5 named barriers
            case 0: asm volatile("barrier.arrive 1, 64;" ::: "memory"); break;
            case 1: asm volatile("barrier.arrive 2, 64;" ::: "memory"); break;
            case 2: asm volatile("barrier.arrive 3, 64;" ::: "memory"); break;
            default: asm volatile("barrier.arrive 4, 64;" ::: "memory"); break;
          }
        }
    } else {
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_b, t = (tile % tiles_per_b) * kRtFrames + f;
        const bool live = t < p.T;
        const float4* erow = reinterpret_cast<const float4*>(p.e + (static_cast<long long>(b) * p.T + (live ? t : 0)) * kRtE);
        float cur[8];
        {
          const float4 a = live ? __ldg(erow) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 c = live ? __ldg(erow + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          cur[0] = a.x; cur[1] = a.y; cur[2] = a.z; cur[3] = a.w; cur[4] = c.x; cur[5] = c.y; cur[6] = c.z; cur[7] = c.w;
        }
        for (int l = 0; l < p.n_levels; ++l) {
          if (p.latents != nullptr && live) {
#pragma unroll
            for (int k = 0; k < 8; ++k) p.latents[(static_cast<long long>(b) * kRtE + l * 8 + k) * p.T + t] = cur[k];
          }
          // F.normalize: x / max(|x|_2, 1e-12)
          float n2 = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) n2 = fmaf(cur[k], cur[k], n2);
          const float den = fmaxf(sqrtf(n2), 1e-12f);
          float hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float en = __fdiv_rn(cur[k], den);
            hi[k] = tf32_rna(en);
            lo[k] = tf32_rna(en - hi[k]);
          }
          // A row (K = 32): [e_hi | e_lo | e_hi | 1 1 0 0 0 0 0 0] against B rows [c_hi | c_hi | c_lo | x_hi x_lo 0 ...]
          sts128(a_row + ((0 ^ sw) << 4), make_float4(hi[0], hi[1], hi[2], hi[3]));
          sts128(a_row + ((1 ^ sw) << 4), make_float4(hi[4], hi[5], hi[6], hi[7]));
          sts128(a_row + ((2 ^ sw) << 4), make_float4(lo[0], lo[1], lo[2], lo[3]));
          sts128(a_row + ((3 ^ sw) << 4), make_float4(lo[4], lo[5], lo[6], lo[7]));
          sts128(a_row + ((4 ^ sw) << 4), make_float4(hi[0], hi[1], hi[2], hi[3]));
          sts128(a_row + ((5 ^ sw) << 4), make_float4(hi[4], hi[5], hi[6], hi[7]));
          sts128(a_row + ((6 ^ sw) << 4), make_float4(1.f, 1.f, 0.f, 0.f));
          sts128(a_row + ((7 ^ sw) << 4), make_float4(0.f, 0.f, 0.f, 0.f));
          fence_proxy_async_smem();
          mbar_arrive(a_ready);

          // next level's latent, minus everything that does not depend on this level's decision (loads overlap the scan);
          // parked in shared memory so the scan keeps its registers for the TMEM loads in flight
          if (l + 1 < p.n_levels) {
            float nxt[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) nxt[k] = 0.f;
            if (live) {
              const float4 a = __ldg(erow + 2 * (l + 1)), c = __ldg(erow + 2 * (l + 1) + 1);
              nxt[0] = a.x; nxt[1] = a.y; nxt[2] = a.z; nxt[3] = a.w; nxt[4] = c.x; nxt[5] = c.y; nxt[6] = c.z; nxt[7] = c.w;
            }
            for (int j = 0; j < l; ++j) {
              const int cj = s_chosen[j * 128 + f];
              const float4* gp = reinterpret_cast<const float4*>(p.g + ((static_cast<long long>(l + 1) * 12 + j) * 1024 + cj) * 8);
              const float4 a = __ldg(gp), c = __ldg(gp + 1);
              nxt[0] -= a.x; nxt[1] -= a.y; nxt[2] -= a.z; nxt[3] -= a.w; nxt[4] -= c.x; nxt[5] -= c.y; nxt[6] -= c.z; nxt[7] -= c.w;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s_nxt[k * 128 + f] = nxt[k];
          }

          float best;
          int bidx;
          scan_level(best, bidx);
          // merge with the upper column half: larger score wins, equal scores go to the smaller code (first maximum)
          switch (quad) {
            case 0: asm volatile("barrier.sync 1, 64;" ::: "memory"); break;
            case 1: asm volatile("barrier.sync 2, 64;" ::: "memory"); break;
            case 2: asm volatile("barrier.sync 3, 64;" ::: "memory"); break;
            default: asm volatile("barrier.sync 4, 64;" ::: "memory"); break;
          }
          {
            const float ob = s_pbest[f];
            const int oi = s_pidx[f];
            if (ob > best || (ob == best && oi < bidx)) bidx = oi;
          }

          const long long oidx = (static_cast<long long>(b) * p.n_levels + l) * p.T + t;
          if (live) p.codes[oidx] = bidx;
          const int chosen = (p.forced != nullptr && live) ? static_cast<int>(p.forced[oidx]) : bidx;
          s_chosen[l * 128 + f] = chosen;
          if (l + 1 < p.n_levels) {
            const float4* gp = reinterpret_cast<const float4*>(p.g + ((static_cast<long long>(l + 1) * 12 + l) * 1024 + chosen) * 8);
            const float4 a = __ldg(gp), c = __ldg(gp + 1);
            cur[0] = s_nxt[0 * 128 + f] - a.x; cur[1] = s_nxt[1 * 128 + f] - a.y; cur[2] = s_nxt[2 * 128 + f] - a.z;
            cur[3] = s_nxt[3 * 128 + f] - a.w; cur[4] = s_nxt[4 * 128 + f] - c.x; cur[5] = s_nxt[5 * 128 + f] - c.y;
            cur[6] = s_nxt[6 * 128 + f] - c.z; cur[7] = s_nxt[7 * 128 + f] - c.w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
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
  __shared__ float tile[32][kRtLatent / 4 + 1];  // 32 frames x 256 channels (one quarter of the channels per pass)
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
            long long code = p.codes[(static_cast<long long>(b) * p.L + l) * p.T + t];
            code = code < 0 ? 0 : (code >= kRtCodes ? kRtCodes - 1 : code);  // memory safety; the host mirror validates the range
            const float4* src = reinterpret_cast<const float4*>(p.proj + (static_cast<long long>(l) * kRtCodes + code) * kRtLatent + cq * 256 + part * 32);
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
          p.out[(plane * kRtLatent + cq * 256 + c) * p.T + t] = tile[f][c];
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace edm
