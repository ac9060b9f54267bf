// CUDA-core twin of the tcgen05 tap-GEMM: same stage tables, same bf16 hi|lo operand buffers, same epilogues.
// It exists so that tests can tell a wrong stage table / packing from a wrong tensor-core kernel; it is never
// selected unless the plan is created with gemm_impl = LSNF_GEMM_SIMT.
#include <type_traits>

#include "tapgemm_common.cuh"

namespace lsnf {

constexpr int SIMT_BN = 64;
constexpr int SIMT_KS = 32;  // channels staged per shared-memory slab

__global__ void __launch_bounds__(256) tapgemm_simt_kernel(const __grid_constant__ StageDev st) {
  __shared__ __align__(16) float As[SIMT_KS][BLOCK_M + 4];
  __shared__ __align__(16) float Bs[SIMT_KS][SIMT_BN + 4];
  const int tid = threadIdx.x;
  const int mtile = blockIdx.x, n0 = blockIdx.y * SIMT_BN;
  const int phase = blockIdx.z / st.ksplit, split = blockIdx.z % st.ksplit;
  const int kblocks = st.Ka / BLOCK_K;
  const int total = st.ph[phase].ntaps * kblocks;
  const int it0 = split * st.it_per_split, it1 = min(total, it0 + st.it_per_split);
  int b0, h0, w0;
  tile_origin(st, mtile, b0, h0, w0);

  // A loader: row = tid/2, 16 channels; B loader: n = tid/4, 8 channels
  const int ar = tid >> 1, ah = tid & 1;
  const int a_b = b0 + ar / (st.bW * st.bH), a_m = h0 + (ar / st.bW) % st.bH, a_n = w0 + ar % st.bW;
  const int bn = tid >> 2, bq = tid & 3;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int it = it0; it < it1; ++it) {
    const int t = it / kblocks, kb = it % kblocks;
    int dy, dx, plane, brow, bcol;
    get_tap(st, phase, t, dy, dx, plane, brow, bcol);
    const int y = a_m + dy, x = a_n + dx;
    const bool a_ok = a_b < st.B && y >= 0 && y < st.aH && x >= 0 && x < st.aW;
    const uint16_t* arow = (const uint16_t*)st.a +
                           ((((size_t)plane * st.B + a_b) * st.aH + y) * st.aW + x) * 2 * st.Ka + kb * BLOCK_K;
    const int brow_i = brow + n0 + bn;
    const bool b_ok = brow_i < st.b_rows && (n0 + bn) < st.n_pad;
    const uint16_t* brw = (const uint16_t*)st.b + (size_t)brow_i * 2 * st.b_k + bcol + kb * BLOCK_K;
    const bool f16 = st.fp16 != 0;
    const bool one = st.passes == 1;   // single pass: the lo halves are neither written nor read
    for (int slab = 0; slab < BLOCK_K / SIMT_KS; ++slab) {
      {
        __align__(16) uint16_t hi[16], lo[16];
        if (a_ok) {
          const uint16_t* p = arow + slab * SIMT_KS + ah * 16;
          *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(p);
          *reinterpret_cast<uint4*>(hi + 8) = *reinterpret_cast<const uint4*>(p + 8);
          *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(p + st.Ka);
          *reinterpret_cast<uint4*>(lo + 8) = *reinterpret_cast<const uint4*>(p + st.Ka + 8);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          As[ah * 16 + i][ar] = a_ok ? join16(hi[i], one ? (uint16_t)0 : lo[i], f16) : 0.f;
      }
      {
        __align__(16) uint16_t hi[8], lo[8];
        if (b_ok) {
          const uint16_t* p = brw + slab * SIMT_KS + bq * 8;
          *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(p);
          *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(p + st.b_k);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          Bs[bq * 8 + i][bn] = b_ok ? join16(hi[i], one ? (uint16_t)0 : lo[i], f16) : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < SIMT_KS; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const float descale = st.descale ? __ldg(st.descale) : 1.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const RowCtx rc = tile_row(st, mtile, ty * 8 + i);
    uint32_t* word = nullptr;
    uint32_t w = epilogue_store<4>(st, phase, split, rc, n0 + tx * 4, acc[i], descale, &word) << (4 * (tx & 7));
    if (st.epi == EPI_ACT_HL) {   // uniform: 8 lanes (32 channels) assemble one word of LeakyReLU sign bits
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if ((tx & 7) == 0 && word) *word = w;
    }
  }
}

int launch_tapgemm_simt(const StageHost& sh, cudaStream_t s) {
  const StageDev& st = sh.dev;
  dim3 grid(st.tiles_b * st.tiles_h * st.tiles_w, (st.n_pad + SIMT_BN - 1) / SIMT_BN, st.nphase * st.ksplit);
  tapgemm_simt_kernel<<<grid, 256, 0, s>>>(st);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

}  // namespace lsnf
