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
  // 64 float4 columns (256 parameters) per block x 4 slices of the partial index: thread (c, sl) sums partials sl, sl+4, ... of
  // its column with 16-byte loads (8 in flight), then the four slices are added in fixed order -> deterministic result.
  __shared__ float4 sl_sum[4][64];
  __shared__ float sq[2];
  const int c = tid & 63, sl = tid >> 6;
  const long long i = ((long long)blockIdx.x * 64 + c) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < a.P) {
    int parts = 0;
    for (int s = 0; s < a.nseg; ++s)
      if (i >= a.seg_begin[s] && i < a.seg_end[s]) parts = a.seg_parts[s];     // segment bounds are multiples of 4 floats
    const float* gp = a.gpart + i;
    float4 b0 = acc, b1 = acc, b2 = acc, b3 = acc;
    int p = sl;
    for (; p + 28 < parts; p += 32) {
      const float4 v0 = mx_ld4(gp + (size_t)p * a.P), v1 = mx_ld4(gp + (size_t)(p + 4) * a.P);
      const float4 v2 = mx_ld4(gp + (size_t)(p + 8) * a.P), v3 = mx_ld4(gp + (size_t)(p + 12) * a.P);
      const float4 v4 = mx_ld4(gp + (size_t)(p + 16) * a.P), v5 = mx_ld4(gp + (size_t)(p + 20) * a.P);
      const float4 v6 = mx_ld4(gp + (size_t)(p + 24) * a.P), v7 = mx_ld4(gp + (size_t)(p + 28) * a.P);
      b0.x += v0.x; b0.y += v0.y; b0.z += v0.z; b0.w += v0.w;
      b1.x += v1.x; b1.y += v1.y; b1.z += v1.z; b1.w += v1.w;
      b2.x += v2.x; b2.y += v2.y; b2.z += v2.z; b2.w += v2.w;
      b3.x += v3.x; b3.y += v3.y; b3.z += v3.z; b3.w += v3.w;
      b0.x += v4.x; b0.y += v4.y; b0.z += v4.z; b0.w += v4.w;
      b1.x += v5.x; b1.y += v5.y; b1.z += v5.z; b1.w += v5.w;
      b2.x += v6.x; b2.y += v6.y; b2.z += v6.z; b2.w += v6.w;
      b3.x += v7.x; b3.y += v7.y; b3.z += v7.z; b3.w += v7.w;
    }
    for (; p < parts; p += 4) {
      const float4 v = mx_ld4(gp + (size_t)p * a.P);
      b0.x += v.x; b0.y += v.y; b0.z += v.z; b0.w += v.w;
    }
    acc.x = (b0.x + b1.x) + (b2.x + b3.x); acc.y = (b0.y + b1.y) + (b2.y + b3.y);
    acc.z = (b0.z + b1.z) + (b2.z + b3.z); acc.w = (b0.w + b1.w) + (b2.w + b3.w);
  }
  sl_sum[sl][c] = acc;
  __syncthreads();
  if (tid < 64) {
    const float4 s0 = sl_sum[0][tid], s1 = sl_sum[1][tid], s2 = sl_sum[2][tid], s3 = sl_sum[3][tid];
    float4 g;
    g.x = (s0.x + s1.x) + (s2.x + s3.x); g.y = (s0.y + s1.y) + (s2.y + s3.y);
    g.z = (s0.z + s1.z) + (s2.z + s3.z); g.w = (s0.w + s1.w) + (s2.w + s3.w);
    const long long j = ((long long)blockIdx.x * 64 + tid) * 4;
    float q = 0.f;
    if (j < a.P) {
      mx_st4(a.grad + j, g);
      q = (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
    }
    q = mx_warp_sum(q);                       // per-block sum of squares: lets k_adam skip re-reading the whole gradient
    if ((tid & 31) == 0) sq[tid >> 5] = q;
  }
  __syncthreads();
  if (tid == 0 && a.normpart) a.normpart[blockIdx.x] = sq[0] + sq[1];
}

__global__ void __launch_bounds__(1024) k_adam(OptimArgs a) {
  __shared__ double red[32];
  __shared__ float s_scale, s_step, s_bc2s;
  const int tid = threadIdx.x;
  MX_PDL_WAIT();
  const float denom = a.grad[a.P + 0];
  const float invd = 1.0f / denom;
  // ||g||^2 over the full vector, identical summation order in every CTA (no grid-wide barrier needed)
  double ds;
  if (a.normpart) {       // single GPU: per-block sums of squares of the numerators left by k_grad_reduce
    double sp = 0.0;
    for (int i = tid; i < a.normpart_n; i += blockDim.x) sp += (double)a.normpart[i];
    ds = mx_warp_sum_d(sp) * ((double)invd * (double)invd);
  } else {                // data parallel: the gradient was all-reduced after k_grad_reduce
    float sf = 0.f;
    const long long P4 = a.P / 4;
#pragma unroll 4
    for (long long i = tid; i < P4; i += blockDim.x) {
      const float4 g = mx_ld4(a.grad + 4 * i);
      const float gx = g.x * invd, gy = g.y * invd, gz = g.z * invd, gw = g.w * invd;
      sf += (gx * gx + gy * gy) + (gz * gz + gw * gw);
    }
    ds = mx_warp_sum_d((double)sf);
  }
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
  const int grid = mx_grad_reduce_blocks(a.P) + 1;
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
