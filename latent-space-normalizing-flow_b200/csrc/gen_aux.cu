// Small HBM-bound kernels around the generator tap-GEMMs: weight packing, latent splitting, the
// reconstruction-gradient seed (with its im2col), split-K reduction and the fused Langevin update.
#include "lsnf_internal.cuh"

namespace lsnf {

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------------------------
// weight packing: ConvTranspose2d weight [ci][co][k][k] fp32 -> K-major bf16 hi|lo operand of one stage
// (model.py:57-149 weights; layouts documented at pack_index in lsnf_internal.cuh)
// ---------------------------------------------------------------------------------------------------
__global__ void pack_stage_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, PackGeom g,
                                  int rows, int nz_valid) {
  const long long total = (long long)rows * g.ka;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / g.ka), col = (int)(i % g.ka);
    int ci = -1, co = -1, tap = -1;
    if (g.kind == 0) {
      ci = col;
      if (g.first) { tap = row / g.co; co = row % g.co; }
      else { tap = row / g.n_pad; co = row % g.n_pad; }
    } else if (g.first || g.last) {
      ci = row; tap = col / g.co; co = col % g.co;
    } else {
      tap = row / g.n_pad; ci = row % g.n_pad; co = col;
    }
    float v = 0.f;
    if (ci < g.ci && co < g.co && tap < g.k * g.k && ci < nz_valid)
      v = w[((long long)ci * g.co + co) * g.k * g.k + tap];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    out[(long long)row * 2 * g.ka + col] = hi;
    out[(long long)row * 2 * g.ka + g.ka + col] = lo;
  }
}

int launch_pack_stage(const lsnf_plan* plan, const StageHost& st, const float* w, cudaStream_t s) {
  PackGeom g;
  g.kind = st.kind; g.first = st.first; g.last = st.last; g.k = st.k; g.ci = st.ci; g.co = st.co;
  g.n_pad = st.info.n_pad; g.ka = st.info.b_k;
  const int rows = st.info.b_rows;
  const long long total = (long long)rows * g.ka;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148 * 16);
  pack_stage_kernel<<<blocks, threads, 0, s>>>(w, (__nv_bfloat16*)(plan->ws + st.b_off), g, rows, st.ci);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// z [B][nz] fp32 -> bf16 hi|lo [B][2*kp] (operand of the first generator layer, train.py:312)
// ---------------------------------------------------------------------------------------------------
__global__ void split_z_kernel(const float* __restrict__ z, __nv_bfloat16* __restrict__ zhl, int B, int nz, int kp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kp) return;
  const int b = i / kp, j = i % kp;
  const float v = j < nz ? z[(size_t)b * nz + j] : 0.f;
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  zhl[(size_t)b * 2 * kp + j] = hi;
  zhl[(size_t)b * 2 * kp + kp + j] = lo;
}

int launch_split_z(const lsnf_plan* plan, const float* z, cudaStream_t s) {
  const int n = plan->cfg.batch * plan->kp;
  split_z_kernel<<<(n + 255) / 256, 256, 0, s>>>(z, (__nv_bfloat16*)(plan->ws + plan->off_zhl), plan->cfg.batch,
                                                 plan->cfg.nz, plan->kp);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// seed of the reconstruction gradient (train.py:313-314):
//   g[b,c,oy,ox] = (x_hat - x) / sigma^2 * (1 - x_hat^2)          (MSE-sum derivative times tanh')
// written directly as the im2col operand of the last layer's data gradient:
//   A[b][iy][ix][tap*nc + c] = g[b][c][iy*s - p + ky][ix*s - p + kx]  (0 outside the image), 64 columns.
// ---------------------------------------------------------------------------------------------------
__global__ void recon_grad_im2col_kernel(const float* __restrict__ xhat, const float* __restrict__ x,
                                         __nv_bfloat16* __restrict__ a, int B, int nc, int img, int hin, int k,
                                         int s, int p, float inv_sigma2) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * hin * hin * BLOCK_K;
  if (i >= total) return;
  const int col = (int)(i % BLOCK_K);
  const long long row = i / BLOCK_K;
  const int ix = (int)(row % hin), iy = (int)((row / hin) % hin), b = (int)(row / ((long long)hin * hin));
  float v = 0.f;
  if (col < k * k * nc) {
    const int tap = col / nc, c = col % nc;
    const int oy = iy * s - p + tap / k, ox = ix * s - p + tap % k;
    if (oy >= 0 && oy < img && ox >= 0 && ox < img) {
      const size_t o = (((size_t)b * nc + c) * img + oy) * img + ox;
      const float xh = xhat[o];
      v = (xh - x[o]) * inv_sigma2 * (1.f - xh * xh);
    }
  }
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  a[row * 2 * BLOCK_K + col] = hi;
  a[row * 2 * BLOCK_K + BLOCK_K + col] = lo;
}

int launch_recon_grad_im2col(const lsnf_plan* plan, const float* x, float sigma, cudaStream_t s) {
  const auto& y = plan->layers[plan->n_layers - 1];
  const long long total = (long long)plan->cfg.batch * y.hin * y.hin * BLOCK_K;
  recon_grad_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
      (const float*)(plan->ws + plan->off_xhat), x, (__nv_bfloat16*)(plan->ws + plan->off_im2col), plan->cfg.batch,
      plan->cfg.nc, plan->img, y.hin, y.k, y.s, y.p, 1.f / (sigma * sigma));
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// split-K partials [S][B][nzp] -> grad_z [B][nz]  (standalone lsnf_generator_dgrad only; the Langevin loop
// folds this sum into the update kernel)
// ---------------------------------------------------------------------------------------------------
__global__ void reduce_partial_kernel(const float* __restrict__ part, float* __restrict__ g, int S, int B, int nz,
                                      int nzp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nz) return;
  const int b = i / nz, j = i % nz;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += part[((size_t)s * B + b) * nzp + j];
  g[i] = acc;
}

int launch_reduce_partial(const lsnf_plan* plan, float* grad_z, cudaStream_t s) {
  const int n = plan->cfg.batch * plan->cfg.nz;
  reduce_partial_kernel<<<(n + 255) / 256, 256, 0, s>>>((const float*)(plan->ws + plan->off_partial), grad_z,
                                                       plan->ksplit_first, plan->cfg.batch, plan->cfg.nz, plan->nzp);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

// ---------------------------------------------------------------------------------------------------
// fused Langevin update (train.py:324-329): one warp per sample, 4 latent elements per lane per pass
//   z <- z - s^2/2 (grad_g + grad_f) + s * noise ;  per-sample |grad_g|, |grad_f| ; next step's bf16 hi|lo z
// Noise: injected eps, or Philox4x32-10 keyed by (seed, global sample index, step, element quad) drawn in
// registers (oracle/philox.py is the checker).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u0 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-8f;  // 2^-24
  const float u1 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-8f;
  const float r = sqrtf(-2.f * logf(u0));
  float sn, cs;
  sincospif(2.f * u1, &sn, &cs);
  n0 = r * cs; n1 = r * sn;
}

__global__ void __launch_bounds__(256) langevin_update_kernel(
    float* __restrict__ z, const float* __restrict__ gg, const float* __restrict__ partial, int nsplit, int nzp,
    const float* __restrict__ gf, const float* __restrict__ eps, __nv_bfloat16* __restrict__ zhl, int B, int nz,
    int kp, float step, int with_noise, uint64_t seed, uint64_t sample_offset, uint32_t step_idx,
    float* __restrict__ norm_scratch, unsigned int* __restrict__ ticket, float* __restrict__ gnorms) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp < B) {
    const int b = warp;
    float sg = 0.f, sf = 0.f;
    for (int q = lane; q < nz / 4; q += 32) {
      const size_t o = (size_t)b * nz + 4 * q;
      float4 g;
      if (partial) {
        g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < nsplit; ++s) {
          const float4 p = *reinterpret_cast<const float4*>(partial + ((size_t)s * B + b) * nzp + 4 * q);
          g.x += p.x; g.y += p.y; g.z += p.z; g.w += p.w;
        }
      } else {
        g = *reinterpret_cast<const float4*>(gg + o);
      }
      const float4 f = *reinterpret_cast<const float4*>(gf + o);
      float4 v = *reinterpret_cast<const float4*>(z + o);
      const float h = 0.5f * step * step;
      v.x = v.x - h * (g.x + f.x); v.y = v.y - h * (g.y + f.y);
      v.z = v.z - h * (g.z + f.z); v.w = v.w - h * (g.w + f.w);
      if (eps) {
        const float4 e = *reinterpret_cast<const float4*>(eps + o);
        v.x += step * e.x; v.y += step * e.y; v.z += step * e.z; v.w += step * e.w;
      } else if (with_noise) {
        const uint64_t sample = sample_offset + (uint64_t)b;
        uint32_t r[4];
        philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), step_idx, (uint32_t)q, (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        float n0, n1, n2, n3;
        box_muller(r[0], r[1], n0, n1);
        box_muller(r[2], r[3], n2, n3);
        v.x += step * n0; v.y += step * n1; v.z += step * n2; v.w += step * n3;
      }
      *reinterpret_cast<float4*>(z + o) = v;
      sg += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
      sf += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
      if (zhl) {
        __nv_bfloat16 hi[4], lo[4];
        split_bf16(v.x, hi[0], lo[0]); split_bf16(v.y, hi[1], lo[1]);
        split_bf16(v.z, hi[2], lo[2]); split_bf16(v.w, hi[3], lo[3]);
        __nv_bfloat16* row = zhl + (size_t)b * 2 * kp;
        *reinterpret_cast<uint2*>(row + 4 * q) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(row + kp + 4 * q) = *reinterpret_cast<uint2*>(lo);
      }
    }
    if (gnorms) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sf += __shfl_xor_sync(0xffffffffu, sf, o);
      }
      if (lane == 0) { norm_scratch[b] = sqrtf(sg); norm_scratch[B + b] = sqrtf(sf); }
    }
  }
  if (!gnorms) return;
  // the last block to finish averages the per-sample norms in a fixed order (deterministic diagnostics)
  __shared__ bool is_last;
  __shared__ float red[2][256];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float a = 0.f, c = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    a += __ldcg(norm_scratch + b);
    c += __ldcg(norm_scratch + B + b);
  }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    gnorms[0] = red[0][0] / (float)B;
    gnorms[1] = red[1][0] / (float)B;
    *ticket = 0u;
  }
}

int launch_update(const lsnf_plan* plan, float* z, const float* gg, const float* partial, int nsplit,
                  const float* gf, float step, const float* eps, int with_noise, uint64_t seed,
                  uint64_t sample_offset, uint32_t step_idx, const uint32_t*, float* gnorms, int write_zhl,
                  cudaStream_t s) {
  const int B = plan->cfg.batch;
  const int blocks = (B * 32 + 255) / 256;
  // scalars block: [0] = ticket of the last-block-done reduction
  unsigned int* ticket = (unsigned int*)(plan->ws + plan->off_scalars);
  float* scratch = (float*)(plan->ws + plan->off_norms);
  __nv_bfloat16* zhl = (write_zhl && plan->n_layers) ? (__nv_bfloat16*)(plan->ws + plan->off_zhl) : nullptr;
  langevin_update_kernel<<<blocks, 256, 0, s>>>(z, gg, partial, nsplit, plan->nzp, gf, eps, zhl, B, plan->cfg.nz,
                                                plan->kp, step, with_noise, seed, sample_offset, step_idx, scratch,
                                                ticket, gnorms);
  LSNF_CUDA(cudaGetLastError());
  return LSNF_OK;
}

}  // namespace lsnf
