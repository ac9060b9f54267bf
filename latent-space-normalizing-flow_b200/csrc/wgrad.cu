// Generator parameter gradients (train.py:390-394: loss_g = mse_sum(G(z_k), x) / B; loss_g.backward()).
//
// The data-gradient chain of the Langevin loop already leaves, for every layer l, the gradient w.r.t. its
// pre-activation output (gpre_l, LeakyReLU' applied) next to the layer's input activation act_{l-1}.  The weight
// gradient of a ConvTranspose2d is then, per kernel tap (ky, kx),
//     dW[ci][co][ky][kx] = sum_{b, iy, ix} act[b][iy][ix][ci] * gpre[b][iy*s - p + ky][ix*s - p + kx][co]
// i.e. a GEMM whose K axis is (batch x positions).  The tcgen05 tap-GEMM wants K-major operands, so both tensors are
// first TRANSPOSED to [channel][position] with the positions of every sample laid out on a grid with one zero halo
// row above and below (Hp = H + 2, Wp = W rounded up to 8): a vertical tap shift dy then is the constant offset dy*Wp
// along the flattened K axis and reads the zero halo outside the image.  TMA box origins must be 16-byte aligned
// along the innermost axis (an odd element offset raises an illegal-instruction fault), so the horizontal shift
// dx = +-1 cannot be a coordinate offset: the transposed gradient is stored twice per phase plane, as is and
// pre-shifted by one column (`xvar`), and each tap picks its copy.  One 2-D TMA box per K block feeds each operand
// (wgrad stages of tapgemm_tc_kernel).  Values are re-split to bf16 hi|lo pairs (fp32 range, 16 significant bits; 3 MMA passes).
// A finalize kernel sums the split-K partials and writes the gradient in the parameter's own [C_in][C_out][k][k]
// layout; bias gradients are row sums of the transposed gradient.
#include <algorithm>

#include "lsnf_internal.cuh"

namespace lsnf {

// ---------------------------------------------------------------------------------------------------
// [P][B][H][W][2C] 16-bit hi|lo (position-major)  ->  [P][c_rows][2*Kp] bf16 hi|lo (channel-major, zero halo)
// ---------------------------------------------------------------------------------------------------

constexpr int TR_POS = 32, TR_CH = 64;

__global__ void __launch_bounds__(256) transpose_hl_kernel(TransArgs a) {
  __shared__ uint16_t hi[TR_CH][TR_POS + 2], lo[TR_CH][TR_POS + 2];
  const int tid = threadIdx.x;
  const long long npos = (long long)a.B * a.H * a.W;
  const long long pos0 = (long long)blockIdx.x * TR_POS;
  const int c0 = blockIdx.y * TR_CH, p = blockIdx.z;
  {
    const int pr = tid >> 3, cq = tid & 7;          // position within the tile, group of 8 channels
    const long long pos = pos0 + pr;
    const int c = c0 + cq * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (pos < npos && c < a.C) {
      const int x = (int)(pos % a.W), y = (int)((pos / a.W) % a.H), b = (int)(pos / ((long long)a.W * a.H));
      const uint16_t* s = a.src + (long long)p * a.s_plane + (long long)b * a.s_b + (long long)y * a.s_h + (long long)x * a.s_w + c;
      __align__(16) uint16_t h8[8], l8[8];
      *reinterpret_cast<uint4*>(h8) = *reinterpret_cast<const uint4*>(s);
      if (!a.src_single) *reinterpret_cast<uint4*>(l8) = *reinterpret_cast<const uint4*>(s + a.C);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = join16(h8[j], a.src_single ? (uint16_t)0 : l8[j], a.src_fp16 != 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint16_t h, l;
      split16(v[j], false, h, l);
      hi[cq * 8 + j][pr] = h; lo[cq * 8 + j][pr] = l;
    }
  }
  __syncthreads();
  {
    const int cr = tid >> 2, pq = tid & 3;          // channel row, group of 8 positions
    const int c = c0 + cr;
    if (c < a.C) {
      const int nvar = a.xvar ? 2 : 1;
      uint16_t* row = a.dst + ((long long)p * nvar * a.c_rows + c) * 2 * a.Kp;
      uint16_t* row1 = row + (long long)a.c_rows * 2 * a.Kp;   // the column-shifted copy
      const int sh = (p & 1) ? 1 : -1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long pos = pos0 + pq * 8 + j;
        if (pos < npos) {
          const int x = (int)(pos % a.W), y = (int)((pos / a.W) % a.H), b = (int)(pos / ((long long)a.W * a.H));
          const long long k = ((long long)b * a.Hp + y + a.halo) * a.Wp + x;
          row[k] = hi[cr][pq * 8 + j];
          row[a.Kp + k] = lo[cr][pq * 8 + j];
          if (a.xvar && x + sh >= 0 && x + sh < a.W) {
            row1[k + sh] = hi[cr][pq * 8 + j];
            row1[a.Kp + k + sh] = lo[cr][pq * 8 + j];
          }
        }
      }
    }
  }
}

int launch_transpose_hl(const TransArgs& a, cudaStream_t s) {
  const long long npos = (long long)a.B * a.H * a.W;
  dim3 grid((unsigned)((npos + TR_POS - 1) / TR_POS), (unsigned)((a.C + TR_CH - 1) / TR_CH), (unsigned)a.P);
  transpose_hl_kernel<<<grid, 256, 0, s>>>(a);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// split-K partials [tap][split][C_in][n_pad] -> dW in the parameter's layout [C_in][C_out][kk], times `scale`
// generic source index: tap * s_tap + split * s_split + ci * s_ci + co
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) wgrad_finalize_kernel(FinalizeArgs a) {
  __shared__ float tile[64][33];   // [tap][co]
  const int ci = blockIdx.y, co0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int co = co0 + lane;
  for (int tap = grp; tap < a.kk; tap += 8) {
    float v = 0.f;
    if (co < a.C_out) {
      const float* p = a.part + (long long)tap * a.s_tap + (long long)ci * a.s_ci + co;
      for (int s = 0; s < a.ksplit; ++s) v += __ldcg(p + (long long)s * a.s_split);   // fixed order: deterministic
    }
    tile[tap][lane] = v * a.scale;
  }
  __syncthreads();
  const int ncol = min(32, a.C_out - co0);
  float* o = a.out + ((long long)ci * a.C_out + co0) * a.kk;
  for (int i = threadIdx.x; i < ncol * a.kk; i += blockDim.x) o[i] = tile[i % a.kk][i / a.kk];
}

int launch_wgrad_finalize(const FinalizeArgs& a, cudaStream_t s) {
  if (a.kk > 64) { set_error("wgrad finalize: more than 64 taps"); return LSNF_ERR_INVALID; }
  dim3 grid((unsigned)((a.C_out + 31) / 32), (unsigned)a.C_in);
  wgrad_finalize_kernel<<<grid, 256, 0, s>>>(a);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// bias gradient: db[c] = scale * sum over the selected rows (sel[j] + c) of a transposed hi|lo matrix of ALL their
// 2*Kp elements (value = hi + lo, the halo is zero).  One CTA per output channel.
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) bias_rowsum_kernel(RowSumArgs a) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float acc = 0.f;
  for (int j = 0; j < a.nsel; ++j) {
    const uint4* row = reinterpret_cast<const uint4*>(a.src + (long long)(a.sel[j] + c) * 2 * a.Kp);
    const long long n16 = 2 * a.Kp / 8;   // Kp is a multiple of 64
    for (long long i = threadIdx.x; i < n16; i += blockDim.x) {
      const uint4 q = __ldg(row + i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc += __uint_as_float(w[t] << 16);            // bf16 -> fp32: the low element
        acc += __uint_as_float(w[t] & 0xFFFF0000u);    // the high element
      }
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.out[c] = red[0] * a.scale;
}

int launch_bias_rowsum(const RowSumArgs& a, cudaStream_t s) {
  bias_rowsum_kernel<<<a.C, 256, 0, s>>>(a);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// loss_g = scale * sum (x_hat - x)^2 (train.py:393), one CTA, fixed summation order
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mse_sum_kernel(const float* __restrict__ xh, const float* __restrict__ x,
                                                       long long n, float scale, float* __restrict__ out) {
  __shared__ float red[1024];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = xh[i] - x[i];
    acc = fmaf(d, d, acc);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0] * scale;
}

int launch_mse_sum(const float* xh, const float* x, long long n, float scale, float* out, cudaStream_t s) {
  mse_sum_kernel<<<1, 1024, 0, s>>>(xh, x, n, scale, out);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

}  // namespace lsnf
