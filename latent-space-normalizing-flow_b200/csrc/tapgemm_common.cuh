// Tile -> row-grid mapping and the fused epilogues shared by the tcgen05 tap-GEMM and its SIMT twin.
#pragma once
#include "lsnf_internal.cuh"

namespace lsnf {

struct RowCtx {
  int b, m, n;  // sample and position in the stage's row grid
  bool valid;
};

__device__ __forceinline__ void tile_origin(const StageDev& st, int mtile, int& b0, int& h0, int& w0) {
  const int wt = mtile % st.tiles_w, ht = (mtile / st.tiles_w) % st.tiles_h, bt = mtile / (st.tiles_w * st.tiles_h);
  b0 = bt * st.bB; h0 = ht * st.bH; w0 = wt * st.bW;
}

__device__ __forceinline__ RowCtx tile_row(const StageDev& st, int mtile, int r) {
  int b0, h0, w0;
  tile_origin(st, mtile, b0, h0, w0);
  RowCtx rc;
  rc.n = w0 + r % st.bW;
  rc.m = h0 + (r / st.bW) % st.bH;
  rc.b = b0 + r / (st.bW * st.bH);
  rc.valid = rc.b < st.B && rc.n < st.Wg;   // ragged batch; rows of a partial tile (weight-gradient stages)
  return rc;
}

template <int NV>
__device__ __forceinline__ void split_pack(const float* v, bool fp16, uint16_t* hi, uint16_t* lo) {
#pragma unroll
  for (int j = 0; j < NV; ++j) split16(v[j], fp16, hi[j], lo[j]);
}

// Applies the stage's epilogue to NV consecutive accumulator columns [col, col+NV) of one row and stores them.
// NV is 4 or 8; col % NV == 0.  Reference semantics: bias + LeakyReLU of model.py:56-151, and for the data gradient
// the LeakyReLU derivative autograd applies (train.py:314).  `pre_mask` (optional) holds the NV saved-activation
// halves already fetched by the caller.
// word of the sign-bit array that holds output channel `cb` of the activation element at element offset `off`
// (offsets are multiples of 2*oC: one hi|lo row per position)
__device__ __forceinline__ uint32_t* act_bits_word(const StageDev& st, size_t off, int cb) {
  return st.mbits + off / (size_t)(2 * st.oC) * (size_t)(st.oC >> 5) + (cb >> 5);
}
__device__ __forceinline__ const uint32_t* grad_bits_row(const StageDev& st, const RowCtx& rc) {
  return st.mbits + (((size_t)rc.b * st.Hg + rc.m) * st.Wg + rc.n) * (size_t)(st.oC >> 5);
}

// Returns the NV LeakyReLU sign bits of an EPI_ACT_HL store (bit j = output j is negative), and through `bits_word`
// the word they belong to (bit position col % 32); 0 / nullptr otherwise.  The caller assembles whole words.
template <int NV>
__device__ __forceinline__ uint32_t epilogue_store(const StageDev& st, int phase, int split_idx, const RowCtx& rc,
                                                   int col, const float* acc_in, float descale,
                                                   uint32_t** bits_word = nullptr) {
  if (bits_word) *bits_word = nullptr;
  if (!rc.valid || col >= st.n_pad) return 0u;
  uint32_t signs = 0u;
  using Vec = typename std::conditional<NV == 8, uint4, uint2>::type;
  float acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = acc_in[j] * descale;
  if (st.epi == EPI_ACT_HL) {
    const int pos = col / st.oC, cb = col % st.oC;
    const int mo = rc.m * st.ms + st.ph[phase].mo, no = rc.n * st.ms + st.ph[phase].no;
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float t = acc[j] + __ldg(st.bias + (col + j) % st.bias_mod);
      v[j] = t > 0.f ? t : t * st.leak;
    }
    __align__(16) uint16_t hi[NV], lo[NV];
    split_pack<NV>(v, st.out_fp16 != 0, hi, lo);
    uint16_t* o = (uint16_t*)st.out + (size_t)rc.b * st.sB + (size_t)mo * st.sH + (size_t)no * st.sW +
                  (size_t)pos * st.sPos + cb;
    *reinterpret_cast<Vec*>(o) = *reinterpret_cast<Vec*>(hi);
    *reinterpret_cast<Vec*>(o + st.oC) = *reinterpret_cast<Vec*>(lo);
#pragma unroll
    for (int j = 0; j < NV; ++j) signs |= (uint32_t)(hi[j] >> 15) << j;   // sign of the hi half = sign of the value
    if (bits_word) *bits_word = act_bits_word(st, (size_t)(o - cb - (uint16_t*)st.out), cb);
  } else if (st.epi == EPI_GRAD_HL) {
    const uint32_t mw = __ldg(grad_bits_row(st, rc) + (col >> 5)) >> (col & 31);
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ((mw >> j) & 1u) ? acc[j] * st.leak : acc[j];  // sign of the saved activation
    __align__(16) uint16_t hi[NV], lo[NV];
    split_pack<NV>(v, st.out_fp16 != 0, hi, lo);
    size_t off;
    if (st.split)
      off = (size_t)((rc.m & 1) * 2 + (rc.n & 1)) * st.sP + (size_t)rc.b * st.sB + (size_t)(rc.m >> 1) * st.sH +
            (size_t)(rc.n >> 1) * st.sW;
    else
      off = (size_t)rc.b * st.sB + (size_t)rc.m * st.sH + (size_t)rc.n * st.sW;
    uint16_t* o = (uint16_t*)st.out + off + col;
    *reinterpret_cast<Vec*>(o) = *reinterpret_cast<Vec*>(hi);
    if (!st.out_single) *reinterpret_cast<Vec*>(o + st.oC) = *reinterpret_cast<Vec*>(lo);
  } else {  // EPI_PARTIAL: raw fp32 rows
    const size_t row = ((size_t)rc.b * st.Hg + rc.m) * st.Wg + rc.n;
    float* o = (float*)st.out + ((size_t)split_idx * st.rows_total + row) * st.n_pad + col;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (NV == 8) *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  return signs;
}

// tap of iteration `t` of a phase: table entry, or generated for the first layer's data gradient
__device__ __forceinline__ void get_tap(const StageDev& st, int phase, int t, int& dy, int& dx, int& plane,
                                        int& brow, int& bcol) {
  if (st.tap_gen) {
    dy = t / st.tap_gen; dx = t % st.tap_gen; plane = 0; brow = 0; bcol = t * st.Ka;
  } else {
    const TapDev& tp = st.ph[phase].taps[t];
    dy = tp.dy; dx = tp.dx; plane = tp.plane; brow = tp.brow; bcol = tp.bsh;
  }
}

}  // namespace lsnf
