// Fused multi-tensor Adam (train.py:294-295, :398, :415: torch.optim.Adam, betas (0.5, 0.999), optional weight decay
// added to the gradient): one launch updates up to ADAM_MAX parameter tensors in place -- parameters, exp_avg and
// exp_avg_sq are the caller's own fp32 tensors (the torch Parameters and the optimizer state), gradients come from
// the flat buffers the gradient kernels fill.  HBM-bound: 16 B read + 12 B written per element.
#include <algorithm>
#include <cmath>

#include "lsnf_internal.cuh"

namespace lsnf {

constexpr int ADAM_MAX = 48;         // tensors per launch (the table travels as a kernel parameter, < 4 KB)
constexpr int ADAM_CHUNK = 4096;     // elements per CTA

struct AdamTable {
  float* p[ADAM_MAX];
  const float* g[ADAM_MAX];
  float* m[ADAM_MAX];
  float* v[ADAM_MAX];
  long long size[ADAM_MAX];
  int chunk0[ADAM_MAX + 1];   // first CTA of every tensor
  int kk[ADAM_MAX];           // > 1: the gradient is stored tap-major [kk][outer][inner] while the parameter is
  int inner[ADAM_MAX];        // [outer][inner][kk] (ConvTranspose2d weights, generator weight-gradient kernels)
  int n;
};

struct AdamHyper {
  float step_size;      // lr / (1 - beta1^t)
  float bias2_sqrt;     // sqrt(1 - beta2^t)
  float beta1, beta2, eps, weight_decay;
  const float* grad_scale;   // device scalar multiplied into every gradient (gradient-norm clipping), or null
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable t, AdamHyper h) {
  int ti = 0;
  while (ti + 1 < t.n && (int)blockIdx.x >= t.chunk0[ti + 1]) ++ti;
  const long long base = (long long)(blockIdx.x - t.chunk0[ti]) * ADAM_CHUNK;
  const long long n = t.size[ti];
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const int kk = t.kk[ti], inner = t.inner[ti];
  const long long outer = kk > 1 ? n / ((long long)inner * kk) : 0;
  const float gs = h.grad_scale ? __ldg(h.grad_scale) : 1.f;
  for (long long i = base + threadIdx.x; i < min(n, base + ADAM_CHUNK); i += blockDim.x) {
    long long gi = i;
    if (kk > 1) {
      const long long tap = i % kk, oi = i / kk;           // oi = outer_index * inner + inner_index
      gi = tap * outer * inner + oi;
    }
    const float pv = p[i];
    float gr = g[gi] * gs;
    if (h.weight_decay != 0.f) gr = fmaf(h.weight_decay, pv, gr);
    const float mv = m[i] + (1.f - h.beta1) * (gr - m[i]);                       // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = v[i] * h.beta2 + (1.f - h.beta2) * gr * gr;                 // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = sqrtf(vv) / h.bias2_sqrt + h.eps;
    m[i] = mv; v[i] = vv;
    p[i] = pv - h.step_size * (mv / denom);                                      // param.addcdiv_(exp_avg, denom, -step_size)
  }
}

int launch_adam(int n, float* const* params, const float* const* grads, float* const* m, float* const* v,
                const int64_t* sizes, const int32_t* grad_kk, const int32_t* grad_inner, float lr, float beta1,
                float beta2, float eps, float weight_decay, int64_t step, const float* grad_scale, cudaStream_t s) {
  AdamHyper h;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  h.step_size = (float)((double)lr / bc1);
  h.bias2_sqrt = (float)sqrt(bc2);
  h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay; h.grad_scale = grad_scale;
  for (int first = 0; first < n; first += ADAM_MAX) {
    AdamTable t;
    t.n = std::min(ADAM_MAX, n - first);
    int chunks = 0;
    for (int i = 0; i < t.n; ++i) {
      const int j = first + i;
      if (!params[j] || !grads[j] || !m[j] || !v[j] || sizes[j] < 0) { set_error("adam: null tensor"); return LSNF_ERR_INVALID; }
      t.p[i] = params[j]; t.g[i] = grads[j]; t.m[i] = m[j]; t.v[i] = v[j]; t.size[i] = sizes[j];
      t.kk[i] = grad_kk ? grad_kk[j] : 1; t.inner[i] = grad_inner ? grad_inner[j] : 1;
      if (t.kk[i] > 1 && (t.inner[i] <= 0 || sizes[j] % ((long long)t.inner[i] * t.kk[i]))) {
        set_error("adam: tap-major gradient layout does not divide the tensor"); return LSNF_ERR_INVALID;
      }
      t.chunk0[i] = chunks;
      chunks += (int)((sizes[j] + ADAM_CHUNK - 1) / ADAM_CHUNK);
    }
    t.chunk0[t.n] = chunks;
    if (chunks == 0) continue;
    adam_kernel<<<chunks, 256, 0, s>>>(t, h);
    LSNF_CUDA(cudaGetLastError());
  }
  return LSNF_OK;
}

}  // namespace lsnf
