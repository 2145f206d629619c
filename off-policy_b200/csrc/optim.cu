// Gradient partial reduction, loss / PER finalisation, global-norm clip + Adam, Polyak target update.
//
// reference: torch.nn.utils.clip_grad_norm_(params, max_grad_norm) + torch.optim.Adam.step (qmix.py:190-193),
// soft_update (utils/util.py:123-134), PER priorities (qmix.py:176-181), train_info (qmix.py:195-198).
//
// Pipeline:  k_grad_reduce  : grad[i] = sum over per-CTA partials (deterministic order); block `gridDim.x-1`
//                             also finalises the scalar sums, the PER priorities and bumps the Adam step count
//            [NCCL all-reduce of grad[0 .. P+4) when world_size > 1]
//            k_adam         : every CTA recomputes ||grad||^2 over the whole (L2-resident) buffer in the same
//                             order -- no grid-wide barrier needed -- then clips and updates its own slice.
#include <math.h>

#include "mx_internal.h"
#include "mx_kernels.h"

__global__ void __launch_bounds__(256) k_grad_reduce(OptimArgs a) {
  const int tid = threadIdx.x;
  MX_PDL_WAIT();
  if (blockIdx.x == gridDim.x - 1) {
    // ---- scalar sums: sum(1-bad), loss numerator, sum Q_tot(1-bad) ----
    __shared__ float red[3][256];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int i = tid; i < a.spart_n; i += blockDim.x) { s0 += a.spart[i * 8]; s1 += a.spart[i * 8 + 1]; s2 += a.spart[i * 8 + 2]; }
    red[0][tid] = s0; red[1][tid] = s1; red[2][tid] = s2;
    __syncthreads();
    if (tid == 0) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      for (int i = 0; i < (int)blockDim.x; ++i) { t0 += red[0][i]; t1 += red[1][i]; t2 += red[2][i]; }
      a.grad[a.P + 0] = t0;
      a.grad[a.P + 1] = t1;
      a.grad[a.P + 2] = t2;
      a.grad[a.P + 3] = (float)(a.B * a.T);
      // Adam step count (1-based) and the running powers beta^t for the bias corrections (first step: 1 * beta)
      const double t_old = a.adam_t[0];
      a.adam_t[0] = t_old + 1.0;
      a.adam_t[1] = (t_old == 0.0 ? 1.0 : a.adam_t[1]) * (double)a.beta1;
      a.adam_t[2] = (t_old == 0.0 ? 1.0 : a.adam_t[2]) * (double)a.beta2;
    }
    // ---- PER: new priority = (1-nu) * mean_t|e| + nu * max_t|e| + eps  (mean over all T, masked steps are zeros) ----
    if (a.prio) {
      for (int b = tid; b < a.B; b += blockDim.x) {
        float mx = 0.f, sm = 0.f;
        for (int t = 0; t < a.T; ++t) {
          const float e = fabsf(a.err[(size_t)b * a.T + t]);
          sm += e;
          mx = fmaxf(mx, e);
        }
        a.prio[b] = (1.f - a.per_nu) * (sm / (float)a.T) + a.per_nu * mx + a.per_eps;
      }
    }
    return;
  }
  const long long i = (long long)blockIdx.x * blockDim.x + tid;
  if (i >= a.P) return;
  int parts = 0;
  for (int s = 0; s < a.nseg; ++s)
    if (i >= a.seg_begin[s] && i < a.seg_end[s]) parts = a.seg_parts[s];
  // eight independent chains, unrolled twice: 16 partial loads in flight per thread (fixed order -> deterministic result)
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f, g5 = 0.f, g6 = 0.f, g7 = 0.f;
  const float* gp = a.gpart + i;
  int p = 0;
#pragma unroll 2
  for (; p + 8 <= parts; p += 8) {
    g0 += gp[(size_t)p * a.P];
    g1 += gp[(size_t)(p + 1) * a.P];
    g2 += gp[(size_t)(p + 2) * a.P];
    g3 += gp[(size_t)(p + 3) * a.P];
    g4 += gp[(size_t)(p + 4) * a.P];
    g5 += gp[(size_t)(p + 5) * a.P];
    g6 += gp[(size_t)(p + 6) * a.P];
    g7 += gp[(size_t)(p + 7) * a.P];
  }
  for (; p < parts; ++p) g0 += gp[(size_t)p * a.P];
  g0 += g4; g1 += g5; g2 += g6; g3 += g7;
  a.grad[i] = (g0 + g1) + (g2 + g3);
}

__global__ void __launch_bounds__(1024) k_adam(OptimArgs a) {
  __shared__ double red[32];
  __shared__ float s_scale, s_step, s_bc2s;
  const int tid = threadIdx.x;
  MX_PDL_WAIT();
  const float denom = a.grad[a.P + 0];
  const float invd = 1.0f / denom;
  // ||g||^2 over the full vector, identical summation order in every CTA (no grid-wide barrier needed)
  float sf = 0.f;
  const long long P4 = a.P / 4;
#pragma unroll 4
  for (long long i = tid; i < P4; i += blockDim.x) {
    const float4 g = mx_ld4(a.grad + 4 * i);
    const float gx = g.x * invd, gy = g.y * invd, gz = g.z * invd, gw = g.w * invd;
    sf += (gx * gx + gy * gy) + (gz * gz + gw * gw);
  }
  double ds = mx_warp_sum_d((double)sf);
  if ((tid & 31) == 0) red[tid >> 5] = ds;
  __syncthreads();
  if (tid < 32) {
    double t = tid < (int)(blockDim.x >> 5) ? red[tid] : 0.0;
    t = mx_warp_sum_d(t);
    if (tid == 0) {
      const float norm = (float)sqrt(t);
      float coef = a.max_grad_norm / (norm + 1e-6f);       // clip_grad_norm_: always applied, clamped to 1
      if (coef > 1.f) coef = 1.f;
      s_scale = coef * invd;
      s_step = a.lr / (float)(1.0 - a.adam_t[1]);          // beta1^t, beta2^t maintained by k_grad_reduce
      s_bc2s = (float)sqrt(1.0 - a.adam_t[2]);
      if (blockIdx.x == 0) {
        a.info[0] = a.grad[a.P + 1] * invd;                // loss
        a.info[1] = norm;                                  // grad_norm (pre-clip)
        a.info[2] = a.grad[a.P + 2] / a.grad[a.P + 3];     // Q_tot mean over all (t,b)
        a.info[3] = denom;
      }
    }
  }
  __syncthreads();
  const float scale = s_scale, step = s_step, bc2s = s_bc2s;
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.P; i += (long long)gridDim.x * blockDim.x) {
    float g = a.grad[i] * scale;
    if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, a.theta[i], g);     // torch Adam: grad.add(param, alpha=weight_decay)
    float m = a.adam_m[i], v = a.adam_v[i];
    m = m + (g - m) * (1.f - a.beta1);                   // exp_avg.lerp_(grad, 1 - beta1)
    v = v * a.beta2 + (1.f - a.beta2) * g * g;           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float den = sqrtf(v) / bc2s + a.eps;
    const float th_new = a.theta[i] - step * (m / den);
    a.theta[i] = th_new;
    a.adam_m[i] = m;
    a.adam_v[i] = v;
    if (a.fuse_polyak) a.theta_tgt[i] = a.theta_tgt[i] * (1.0f - a.tau) + th_new * a.tau;   // util.py:132-134 fused epilogue
  }
}

__global__ void __launch_bounds__(256) k_polyak(float* __restrict__ tgt, const float* __restrict__ src, long long n4, float tau) {
  MX_PDL_WAIT();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 t = mx_ld4(tgt + 4 * i);
    const float4 s = mx_ld4(src + 4 * i);
    t.x = t.x * (1.0f - tau) + s.x * tau;                // util.py:132-134
    t.y = t.y * (1.0f - tau) + s.y * tau;
    t.z = t.z * (1.0f - tau) + s.z * tau;
    t.w = t.w * (1.0f - tau) + s.w * tau;
    mx_st4(tgt + 4 * i, t);
  }
}

int mx_launch_grad_reduce(const OptimArgs& a, cudaStream_t s) {
  const int grid = (int)((a.P + 255) / 256) + 1;
  MX_LAUNCH_PDL(k_grad_reduce, dim3(grid), dim3(256), 0, s, a);
  MX_COUNT();
  MX_MARK("k_grad_reduce", s);
  return MX_CHECK_LAUNCH("grad_reduce");
}
int mx_launch_adam(const OptimArgs& a, cudaStream_t s) {
  int grid = (int)((a.P + 4095) / 4096);
  const int sms = mx_num_sms();
  if (grid > sms) grid = sms;
  MX_LAUNCH_PDL(k_adam, dim3(grid), dim3(1024), 0, s, a);
  MX_PDL_THETA_WRITTEN();
  MX_COUNT();
  MX_MARK("k_adam", s);
  return MX_CHECK_LAUNCH("adam");
}
int mx_launch_polyak(float* tgt, const float* src, long long n, float tau, cudaStream_t s) {
  const long long n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  const int sms = mx_num_sms();
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  MX_LAUNCH_PDL(k_polyak, dim3(grid), dim3(256), 0, s, tgt, src, n4, tau);
  MX_PDL_THETA_WRITTEN();
  MX_COUNT();
  MX_MARK("k_polyak", s);
  return MX_CHECK_LAUNCH("polyak");
}
