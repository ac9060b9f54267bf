// tcgen05 / TMEM / TMA tap-GEMM for sm_100a: the generator's transposed convolutions (forward, train.py:312)
// and their data gradients (train.py:314) as implicit GEMMs.
//
//   D[row, n] = sum_{tap} sum_{k} A[plane_tap][b][m + dy_tap][n + dx_tap][k] * W[brow_tap + n][k]
//
// * A rows are 128 grid positions (b, m, n) fetched by ONE 5-D TMA box per tap and K block; the (dy, dx) shift
//   moves the box, and TMA's out-of-bounds zero fill supplies the convolution padding and the ragged batch.
// * Operands are 16-bit hi|lo pairs (value = hi + lo; fp16 for the forward pass, bf16 for the data gradients); each
//   K step issues three tcgen05.mma (lo*hi, hi*lo, hi*hi) into one fp32 TMEM accumulator -- the 3-pass split that
//   keeps z_T within the 1e-4 parity budget.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue
//   (tcgen05.ld -> bias/activation/derivative -> bf16 hi|lo or fp32 stores).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "tapgemm_common.cuh"

namespace lsnf {

constexpr int TC_EPI_WARPS = 8;                       // two warps per TMEM lane quarter, each owning half the columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB per hi or lo half

template <int BN>
struct TcCfg {
  static constexpr int B_TILE_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : (BN == 64 ? 4 : 5));
  static constexpr int SPAN = BN >= 64 ? BN / 2 : BN;   // columns per epilogue warp
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128-byte swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;             // leading byte offset (unused for swizzled K-major), encoded 16 B
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

template <int BN>
__device__ __forceinline__ uint32_t umma_idesc(bool fp16) {
  // c_format F32 (bits 4-5 = 1), a/b format (bits 7-9, 10-12): 0 = F16, 1 = BF16; K-major A and B;
  // N>>3 at bit 17, M>>4 at bit 24
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// Epilogue of one 128-row accumulator tile.  Warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32); warps w and w+4
// split the columns.  Row-invariant addressing is hoisted out of the column loop, and the saved-activation signs a
// data-gradient row needs are fetched BEFORE waiting for the accumulator, i.e. while the main loop still runs.
// Reference semantics: bias + LeakyReLU of model.py:56-151; LeakyReLU' for the data gradient (train.py:314).
template <int BN>
__device__ __forceinline__ void tc_epilogue(const StageDev& st, int mtile, int n0, int phase, int split, int warp,
                                            int lane, uint32_t tmem_acc, uint32_t full_bar, uint32_t parity,
                                            bool have_acc) {
  using Cfg = TcCfg<BN>;
  const int quarter = warp & 3;
  const int half = (warp - 2) >> 2;
  const int r = quarter * 32 + lane;
  const RowCtx rc = tile_row(st, mtile, r);
  constexpr int SPAN = Cfg::SPAN;
  constexpr int CH = SPAN >= 32 ? 32 : 16;
  const int c_begin = half * SPAN;
  const bool active = c_begin < BN;
  const int col0 = n0 + c_begin;
  const int epi = st.epi;
  const bool ok = active && rc.valid && col0 < st.n_pad;
  const int lo_off = st.oC;
  const float leak = st.leak;
  const bool out16 = st.out_fp16 != 0;
  uint16_t* o16 = nullptr;
  float* o32 = nullptr;
  const float* bias = nullptr;
  uint4 mk[SPAN / 8];
  if (ok) {
    if (epi == EPI_ACT_HL) {
      const int pos = col0 / st.oC, cb = col0 % st.oC;
      const int mo = rc.m * st.ms + st.ph[phase].mo, no = rc.n * st.ms + st.ph[phase].no;
      o16 = (uint16_t*)st.out + (size_t)rc.b * st.sB + (size_t)mo * st.sH + (size_t)no * st.sW + (size_t)pos * st.sPos + cb;
      bias = st.bias + cb;
    } else if (epi == EPI_GRAD_HL) {
      size_t off;
      if (st.split)
        off = (size_t)((rc.m & 1) * 2 + (rc.n & 1)) * st.sP + (size_t)rc.b * st.sB + (size_t)(rc.m >> 1) * st.sH +
              (size_t)(rc.n >> 1) * st.sW;
      else
        off = (size_t)rc.b * st.sB + (size_t)rc.m * st.sH + (size_t)rc.n * st.sW;
      o16 = (uint16_t*)st.out + off + col0;
      const uint4* mp = reinterpret_cast<const uint4*>(
          (const uint16_t*)st.mask + (((size_t)rc.b * st.Hg + rc.m) * st.Wg + rc.n) * 2 * st.oC + col0);
#pragma unroll
      for (int j = 0; j < SPAN / 8; ++j) mk[j] = __ldg(mp + j);
    } else {
      const size_t row = ((size_t)rc.b * st.Hg + rc.m) * st.Wg + rc.n;
      o32 = (float*)st.out + ((size_t)split * st.rows_total + row) * st.n_pad + col0;
    }
  }
  const float descale = st.descale ? __ldg(st.descale) : 1.f;
  if (have_acc) mbar_wait(full_bar, parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!active) return;
#pragma unroll
  for (int cc = 0; cc < SPAN; cc += CH) {
    uint32_t v[CH];
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c_begin + cc);
    if constexpr (CH == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (!ok) continue;
#pragma unroll
    for (int j = 0; j < CH; j += 8) {
      float f[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = have_acc ? __uint_as_float(v[j + q]) * descale : 0.f;
      if (epi == EPI_ACT_HL) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cc + j));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cc + j + 4));
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        __align__(16) uint16_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float t = f[q] + bb[q];
          t = t > 0.f ? t : t * leak;
          split16(t, out16, hi[q], lo[q]);
        }
        *reinterpret_cast<uint4*>(o16 + cc + j) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(o16 + lo_off + cc + j) = *reinterpret_cast<uint4*>(lo);
      } else if (epi == EPI_GRAD_HL) {
        const uint16_t* mv = reinterpret_cast<const uint16_t*>(&mk[(cc + j) / 8]);
        __align__(16) uint16_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float t = (mv[q] & 0x8000u) ? f[q] * leak : f[q];   // sign bit of the saved activation
          split16(t, out16, hi[q], lo[q]);
        }
        *reinterpret_cast<uint4*>(o16 + cc + j) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(o16 + lo_off + cc + j) = *reinterpret_cast<uint4*>(lo);
      } else {
        *reinterpret_cast<float4*>(o32 + cc + j) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(o32 + cc + j + 4) = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tapgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ StageDev st) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = tiles + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barrier block: full[STAGES], empty[STAGES], tmem_full, then the TMEM base address slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * Cfg::STAGES + 1);
  uint8_t* gen_base = smem_raw + (tiles - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(gen_base + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtile = blockIdx.x, n0 = blockIdx.y * BN;
  const int phase = blockIdx.z / st.ksplit, split = blockIdx.z % st.ksplit;
  const int kblocks = st.Ka / BLOCK_K;
  const int total = st.ph[phase].ntaps * kblocks;
  const int it0 = split * st.it_per_split, it1 = min(total, it0 + st.it_per_split);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int b0, h0, w0;
      tile_origin(st, mtile, b0, h0, w0);
      for (int it = it0; it < it1; ++it) {
        const int i = it - it0, s = i % Cfg::STAGES;
        const uint32_t par = (uint32_t)((i / Cfg::STAGES) & 1);
        mbar_wait(empty_bar(s), par ^ 1u);
        const int t = it / kblocks, kb = it % kblocks;
        int dy, dx, plane, brow, bcol;
        get_tap(st, phase, t, dy, dx, plane, brow, bcol);
        const uint32_t sa = tiles + s * Cfg::STAGE_BYTES;
        const uint32_t sb = sa + 2 * A_TILE_BYTES;
        mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
        tma_load_5d(sa, &tmA, full_bar(s), kb * BLOCK_K, w0 + dx, h0 + dy, b0, plane);
        tma_load_5d(sa + A_TILE_BYTES, &tmA, full_bar(s), st.Ka + kb * BLOCK_K, w0 + dx, h0 + dy, b0, plane);
        tma_load_2d(sb, &tmB, full_bar(s), bcol + kb * BLOCK_K, brow + n0);
        tma_load_2d(sb + Cfg::B_TILE_BYTES, &tmB, full_bar(s), st.b_k + bcol + kb * BLOCK_K, brow + n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc<BN>(st.fp16 != 0);
      for (int it = it0; it < it1; ++it) {
        const int i = it - it0, s = i % Cfg::STAGES;
        const uint32_t par = (uint32_t)((i / Cfg::STAGES) & 1);
        mbar_wait(full_bar(s), par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = tiles + s * Cfg::STAGE_BYTES;
        const uint32_t sb = sa + 2 * A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t a_hi = umma_desc(sa + k * 32), a_lo = umma_desc(sa + A_TILE_BYTES + k * 32);
          const uint64_t b_hi = umma_desc(sb + k * 32), b_lo = umma_desc(sb + Cfg::B_TILE_BYTES + k * 32);
          umma_bf16(tmem_base, a_lo, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_bf16(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(empty_bar(s));  // frees the smem slot once the MMAs above have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    tc_epilogue<BN>(st, mtile, n0, phase, split, warp, lane, tmem_base, tmem_full_bar, 0u, it1 > it0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the wide stages (N tile 256): two CTAs of one cluster compute a 256 x 256
// output tile.  Each CTA loads its own 128 A rows and HALF of the 256 weight rows, so operand traffic per CTA drops
// from 384 to 256 rows per K block (the 3-pass hi|lo split doubles operand bytes per FLOP and makes the single-CTA
// kernel L2-bound); the leader CTA issues the MMAs for both and commits to the barriers of both.
// ---------------------------------------------------------------------------------------------------
constexpr int P_BN = 256;
constexpr int P_B_TILE_BYTES = (P_BN / 2) * BLOCK_K * 2;                 // this CTA's half of the weight rows
constexpr int P_STAGE_BYTES = 2 * A_TILE_BYTES + 2 * P_B_TILE_BYTES;      // 64 KiB per CTA
constexpr int P_STAGES = 3;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;   // 3 x 64 KiB ring + alignment + barriers
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;                            // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in the LEADER CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}

// Persistent: one CTA pair per SM pair walks the tile list (m-pair fastest, so consecutive tiles share weights in
// L2).  The shared-memory ring keeps streaming across tile boundaries, and the 512 TMEM columns hold TWO 256-column
// accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tapgemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ StageDev st) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = tiles + P_STAGES * P_STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (P_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * P_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * P_STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * P_STAGES + 4);
  uint8_t* gen_base = smem_raw + (tiles - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(gen_base + P_STAGES * P_STAGE_BYTES + 8 * (2 * P_STAGES + 4));

  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = st.Ka / BLOCK_K;
  // tile list: t = (phase * n_tiles + ntile) * m_pairs + mpair
  const int mtiles = st.tiles_b * st.tiles_h * st.tiles_w;
  const int m_pairs = (mtiles + 1) >> 1, n_tiles = st.n_pad / P_BN;
  const int num_tiles = m_pairs * n_tiles * st.nphase;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // leader's producer arrives once per use (+ the bytes of both CTAs)
      mbar_init(empty_bar(s), 1);   // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);                     // multicast commit of the tile's last MMA
      mbar_init(tmem_empty_bar(a), 2 * TC_EPI_WARPS);     // every epilogue warp of both CTAs (leader's copy is used)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything can arrive on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs); completions of both land on the LEADER's full barrier =====
      uint32_t i = 0;   // running K-block counter: the ring does not drain between tiles
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        const int mp = t % m_pairs, nt = (t / m_pairs) % n_tiles, phase = t / (m_pairs * n_tiles);
        const int mtile = 2 * mp + (int)cta_rank;
        int b0, h0, w0;
        tile_origin(st, mtile, b0, h0, w0);
        const int nb = nt * P_BN + (int)cta_rank * (P_BN / 2);
        const int total = st.ph[phase].ntaps * kblocks;
        for (int it = 0; it < total; ++it, ++i) {
          const int s = i % P_STAGES;
          const uint32_t par = (i / P_STAGES) & 1u;
          mbar_wait(empty_bar(s), par ^ 1u);
          const int tp = it / kblocks, kb = it % kblocks;
          int dy, dx, plane, brow, bcol;
          get_tap(st, phase, tp, dy, dx, plane, brow, bcol);
          const uint32_t sa = tiles + s * P_STAGE_BYTES;
          const uint32_t sb = sa + 2 * A_TILE_BYTES;
          const uint32_t fb = full_bar(s) & PEER_BIT_MASK;
          if (leader) mbar_expect_tx(full_bar(s), 2 * P_STAGE_BYTES);
          tma2_load_5d(sa, &tmA, fb, kb * BLOCK_K, w0 + dx, h0 + dy, b0, plane);
          tma2_load_5d(sa + A_TILE_BYTES, &tmA, fb, st.Ka + kb * BLOCK_K, w0 + dx, h0 + dy, b0, plane);
          tma2_load_2d(sb, &tmB, fb, bcol + kb * BLOCK_K, brow + nb);
          tma2_load_2d(sb + P_B_TILE_BYTES, &tmB, fb, st.b_k + bcol + kb * BLOCK_K, brow + nb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader CTA only): M = 256 over both CTAs, N = 256 =====
      const uint32_t fmt = st.fp16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t i = 0, lt = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs, ++lt) {
        const int phase = t / (m_pairs * n_tiles);
        const int total = st.ph[phase].ntaps * kblocks;
        const uint32_t acc = lt & 1u;
        // the epilogues of both CTAs must have drained this accumulator (two tiles ago)
        mbar_wait(tmem_empty_bar(acc), ((lt >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * P_BN;
        for (int it = 0; it < total; ++it, ++i) {
          const int s = i % P_STAGES;
          const uint32_t par = (i / P_STAGES) & 1u;
          mbar_wait(full_bar(s), par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = tiles + s * P_STAGE_BYTES;
          const uint32_t sb = sa + 2 * A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t a_hi = umma_desc(sa + k * 32), a_lo = umma_desc(sa + A_TILE_BYTES + k * 32);
            const uint64_t b_hi = umma_desc(sb + k * 32), b_lo = umma_desc(sb + P_B_TILE_BYTES + k * 32);
            umma2_bf16(d_tmem, a_lo, b_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
            umma2_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
            umma2_bf16(d_tmem, a_hi, b_hi, idesc, 1u);
          }
          umma2_commit_both(empty_bar(s));      // frees the slot in both CTAs
        }
        umma2_commit_both(tmem_full_bar(acc));  // this accumulator is complete in both CTAs
      }
    }
  } else {
    // ===== epilogue warps: tile lt reads accumulator lt & 1, then hands it back to the MMA issuer =====
    uint32_t lt = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs, ++lt) {
      const int mp = t % m_pairs, nt = (t / m_pairs) % n_tiles, phase = t / (m_pairs * n_tiles);
      const int mtile = 2 * mp + (int)cta_rank;
      const uint32_t acc = lt & 1u;
      tc_epilogue<P_BN>(st, mtile, nt * P_BN, phase, 0, warp, lane, tmem_base + acc * P_BN, tmem_full_bar(acc),
                        (lt >> 1) & 1u, true);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty_bar(acc));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_relaxed();   // neither CTA may free TMEM or exit while the pair is still working
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host: tensor-map encoding (driver entry point fetched at run time; the library does not link libcuda)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

static bool pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSNF_NO_PAIR"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}
// the CTA-pair kernel serves the wide stages whose M extent gives both CTAs of every pair a tile
static bool use_pair(const StageHost& sh) {
  const StageDev& d = sh.dev;
  return pair_enabled() && d.block_n == 256 && d.n_pad % 256 == 0 && d.ksplit == 1 &&
         (d.tiles_b * d.tiles_h * d.tiles_w) >= 2;
}

int tc_encode_maps(lsnf_plan* plan, StageHost& sh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return LSNF_ERR_CUDA; }
  const StageDev& d = sh.dev;
  {
    const cuuint64_t row = (cuuint64_t)2 * d.Ka * 2;  // bytes per position (hi|lo)
    cuuint64_t dims[5] = {(cuuint64_t)2 * d.Ka, (cuuint64_t)d.aW, (cuuint64_t)d.aH, (cuuint64_t)d.B, (cuuint64_t)d.aP};
    cuuint64_t strides[4] = {row, row * d.aW, row * d.aW * d.aH, row * d.aW * d.aH * d.B};
    cuuint32_t box[5] = {BLOCK_K, (cuuint32_t)d.bW, (cuuint32_t)d.bH, (cuuint32_t)d.bB, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&sh.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)d.a, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)r) + " for stage layer " +
                std::to_string(sh.layer) + " kind " + std::to_string(sh.kind));
      return LSNF_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)2 * d.b_k, (cuuint64_t)d.b_rows};
    cuuint64_t strides[1] = {(cuuint64_t)2 * d.b_k * 2};
    const bool pair = use_pair(sh);
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(pair ? d.block_n / 2 : d.block_n)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&sh.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d.b, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r) + " for stage layer " +
                std::to_string(sh.layer) + " kind " + std::to_string(sh.kind));
      return LSNF_ERR_CUDA;
    }
  }
  sh.maps_ready = true;
  return LSNF_OK;
}

template <int BN>
static int launch_bn(const StageHost& sh, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    LSNF_CUDA(cudaFuncSetAttribute(tapgemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const StageDev& st = sh.dev;
  dim3 grid(st.tiles_b * st.tiles_h * st.tiles_w, st.n_pad / BN, st.nphase * st.ksplit);
  tapgemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(sh.tmA, sh.tmB, st);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

static int launch_pair(const StageHost& sh, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    LSNF_CUDA(cudaFuncSetAttribute(tapgemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    attr_set = true;
  }
  const StageDev& st = sh.dev;
  const int mtiles = st.tiles_b * st.tiles_h * st.tiles_w;
  const int num_tiles = (mtiles + 1) / 2 * (st.n_pad / P_BN) * st.nphase;   // an odd tail pairs with an empty tile
  static int max_pairs = 0;
  if (!max_pairs) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_pairs = sms / 2;
  }
  dim3 grid(2 * std::min(num_tiles, max_pairs), 1, 1);
  tapgemm_tc2_kernel<<<grid, TC_THREADS, P_SMEM_BYTES, s>>>(sh.tmA, sh.tmB, st);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

int launch_tapgemm_tc(const StageHost& sh, cudaStream_t s) {
  if (!sh.maps_ready) { set_error("tensor maps not encoded"); return LSNF_ERR_STATE; }
  if (use_pair(sh)) return launch_pair(sh, s);
  switch (sh.dev.block_n) {
    case 256: return launch_bn<256>(sh, s);
    case 128: return launch_bn<128>(sh, s);
    case 64: return launch_bn<64>(sh, s);
    case 32: return launch_bn<32>(sh, s);
    case 16: return launch_bn<16>(sh, s);
    default: set_error("unsupported N tile"); return LSNF_ERR_INVALID;
  }
}

}  // namespace lsnf
