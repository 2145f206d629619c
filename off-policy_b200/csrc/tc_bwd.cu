// Weight gradients of the agent network's time-batched layers on tcgen05 (3xTF32): k_wgrad_tc.
//
//   dW_ih = dgi^T x2     dW_hh = [dgi_r, dgi_z, dgi_n * r]^T h_{t-1}     dW2 = da2^T x1     dW1 = da1^T x0     + every bias gradient
//
// These are reductions over the ROWS of the step (K = rows), so both MMA operands are the row-major activations read TRANSPOSED:
// A[feature][row] and B[feature][row], K-major.  A thread stages (feature, 4 consecutive rows) pairs: four scalar loads from the
// L2-resident activations, TF32 hi / lo split, two 16-byte stores into the core-matrix layout; within a warp the 32 pairs are
// (8 features) x (4 row quads), which is 512 contiguous bytes of the tile -- no bank conflicts, every store a full wavefront.
// Normalised activations (x2, x1, x0) are recomputed from the saved pre-norm rows and statistics while staging; bias gradients fall out
// of a column of ones appended to B.  Three accumulators live in TMEM for the whole CTA (rows are split over CTAs in chunks of 64):
//   D1[128][144] = [dgi_r | dgi_z]^T     . [x2 | h_prev | 1]   -> dW_ih[0:128], dW_hh[0:128], db_ih[0:128] (= db_hh[0:128])
//   D2[128][144] = [dgi_n | dgi_n * r]^T . [x2 | h_prev | 1]   -> dW_ih[128:192] (rows 0-63, cols 0-63), dW_hh[128:192] (rows 64-127, cols 64-127)
//   D3[128][..]  = [da2   | da1]^T       . [x1 | x0 | 1]       -> dW2 (rows 0-63, cols 0-63), dW1 (rows 64-127, cols 64..), db2, db1
// The off-diagonal blocks of D2 / D3 are products nobody needs (half of two of the three MMAs): the tensor core is not the limit here.
// Each CTA writes its sums as ONE gradient partial (its row of gpart), like k_front_bwd does for the LayerNorm parameters it keeps.
#include "mx_internal.h"
#include "mx_kernels.h"
#include "mx_tc.cuh"

#include <string.h>

int g_mx_wgrad_tc = 0;        // off until timed on a B200 (emulator-verified): mx_set_option("wgrad_tc", 1)

#define WG_ROWS 64            // rows per MMA group (the K extent of one staged tile)
#define WG_DSTRIDE 160        // TMEM column stride between the three accumulators

struct WgradSmem { int o_ahi, o_alo, o_bhi, o_blo, total; };
static WgradSmem wgrad_smem() {
  WgradSmem s;
  int o = 0;
  s.o_ahi = o; o += 128 * WG_ROWS * 4;
  s.o_alo = o; o += 128 * WG_ROWS * 4;
  s.o_bhi = o; o += 144 * WG_ROWS * 4;
  s.o_blo = o; o += 144 * WG_ROWS * 4;
  s.total = o;
  return s;
}

// one (feature, row quad) pair -> hi / lo tiles
__device__ __forceinline__ void wg_put(char* hi, char* lo, int feat, int kq, const float (&x)[4]) {
  float4 h, l;
  h.x = tc::to_tf32(x[0]); h.y = tc::to_tf32(x[1]); h.z = tc::to_tf32(x[2]); h.w = tc::to_tf32(x[3]);
  l.x = x[0] - h.x; l.y = x[1] - h.y; l.z = x[2] - h.z; l.w = x[3] - h.w;
  const uint32_t o = tc::core_off_bytes(feat, 4 * kq, WG_ROWS);
  *reinterpret_cast<float4*>(hi + o) = h;
  *reinterpret_cast<float4*>(lo + o) = l;
}
// pair index -> (feature, row quad): lanes of a warp cover 8 consecutive features x 4 consecutive row quads
__device__ __forceinline__ void wg_pair(int p, int* feat, int* kq) {
  const int fr = p & 7, kl = (p >> 3) & 3, rest = p >> 5;
  *feat = (rest >> 2) * 8 + fr;
  *kq = (rest & 3) * 4 + kl;
}

struct WgradArgs {
  FrontBwdArgs f;
  int nchunks, Kp16;
};

#define WG_THREADS 512        // staging is load-latency bound: 16 warps keep enough loads in flight; warps 0-3 own the TMEM lanes in the epilogue
__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc(WgradArgs w, WgradSmem sm, int swap_ls) {
  MX_DYN_SMEM_RAW(smem_raw);
  __shared__ __align__(8) tc::Bar bar_s;
  __shared__ uint32_t tmem_s;
  __shared__ float par_s[6 * 64];            // ln2 g,b | ln1 g,b | fn g,b
  const FrontBwdArgs& a = w.f;
  const MxNetLayout L = a.L;
  const float* __restrict__ th = a.theta;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int I = L.in_dim, N = a.N, T1 = a.T1 > 0 ? a.T1 : a.T + 1;
  const int Kp16 = w.Kp16, N3 = 64 + Kp16 + 16, ones3 = 64 + Kp16;
  char* base = reinterpret_cast<char*>(smem_raw);
  char *a_hi = base + sm.o_ahi, *a_lo = base + sm.o_alo, *b_hi = base + sm.o_bhi, *b_lo = base + sm.o_blo;
  const uint32_t bar = tc::bar_addr(&bar_s);
  if (warp == 0) tc::tmem_alloc<512>(&tmem_s);
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init_fence();
  }
  for (int i = tid; i < 64; i += blockDim.x) {
    par_s[i] = th[L.ln2_g + i]; par_s[64 + i] = th[L.ln2_b + i]; par_s[128 + i] = th[L.ln1_g + i]; par_s[192 + i] = th[L.ln1_b + i];
    par_s[256 + i] = i < I ? th[L.fn_g + i] : 0.f; par_s[320 + i] = i < I ? th[L.fn_b + i] : 0.f;
  }
  MX_PDL_WAIT();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = tmem_s;
  uint32_t phase = 0;
  int iter = 0;
  for (int chunk = blockIdx.x; chunk < w.nchunks; chunk += gridDim.x, ++iter) {
    const int row0 = chunk * WG_ROWS;
    const uint32_t acc0 = iter > 0 ? 1u : 0u;
    // ---- B = [x2 | h_prev | 1 | 0..]  (144 features) ----
#pragma unroll 2
    for (int p = tid; p < 144 * 16; p += blockDim.x) {
      int n, kq;
      wg_pair(p, &n, &kq);
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = row0 + 4 * kq + j;
        float v = 0.f;
        if (m < a.M) {
          if (n < 64) v = (a.u2[(size_t)m * MX_H + n] - a.st2[2 * (size_t)m]) * a.st2[2 * (size_t)m + 1] * par_s[n] + par_s[64 + n];
          else if (n < 128) {
            if (((m / N) % T1) > 0) v = a.hall[(size_t)(m - N) * MX_H + (n - 64)];
            else if (a.h0) v = a.h0[(size_t)m * MX_H + (n - 64)];
          } else if (n == 128) v = 1.f;
        }
        x[j] = v;
      }
      wg_put(b_hi, b_lo, n, kq, x);
    }
    // ---- A = [dgi_r | dgi_z] -> D1 ; then A = [dgi_n | dgi_n * r] -> D2 (same B) ----
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 2
      for (int p = tid; p < 128 * 16; p += blockDim.x) {
        int f, kq;
        wg_pair(p, &f, &kq);
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int m = row0 + 4 * kq + j;
          float v = 0.f;
          if (m < a.M) {
            if (pass == 0) v = a.dgi[(size_t)m * MX_G + f];
            else {
              v = a.dgi[(size_t)m * MX_G + 2 * MX_H + (f & 63)];
              if (f >= 64) v *= a.gates[(size_t)m * MX_G + (f - 64)];
            }
          }
          x[j] = v;
        }
        wg_put(a_hi, a_lo, f, kq, x);
      }
      tc::fence_async_smem();
      tc::fence_before();
      __syncthreads();
      tc::fence_after();
      if (tid == 0) tc::issue_layer_acc(tmem_base + pass * WG_DSTRIDE, a_hi, a_lo, b_hi, b_lo, 144, WG_ROWS, swap_ls, acc0, bar);
      tc::mbar_wait(bar, phase);      // the MMAs have read the tiles: A (and after the second pass B) may be refilled
      phase ^= 1;
      tc::fence_after();
    }
    // ---- A = [da2 | da1], B = [x1 | x0 | 1] -> D3 ----
#pragma unroll 2
    for (int p = tid; p < 128 * 16; p += blockDim.x) {
      int f, kq;
      wg_pair(p, &f, &kq);
      const float* src = f < 64 ? a.da2_out : a.da1_out;
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = row0 + 4 * kq + j;
        x[j] = m < a.M ? src[(size_t)m * MX_H + (f & 63)] : 0.f;
      }
      wg_put(a_hi, a_lo, f, kq, x);
    }
#pragma unroll 2
    for (int p = tid; p < N3 * 16; p += blockDim.x) {
      int n, kq;
      wg_pair(p, &n, &kq);
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = row0 + 4 * kq + j;
        float v = 0.f;
        if (m < a.M) {
          if (n < 64) v = (a.u1[(size_t)m * MX_H + n] - a.st1[2 * (size_t)m]) * a.st1[2 * (size_t)m + 1] * par_s[128 + n] + par_s[192 + n];
          else if (n < 64 + I) {
            const int c = n - 64;
            const float xr = a.X[(size_t)m * a.ldx + c];
            v = a.feature_norm ? (xr - a.st0[2 * (size_t)m]) * a.st0[2 * (size_t)m + 1] * par_s[256 + c] + par_s[320 + c] : xr;
          } else if (n == ones3) v = 1.f;
        }
        x[j] = v;
      }
      wg_put(b_hi, b_lo, n, kq, x);
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (tid == 0) tc::issue_layer_acc(tmem_base + 2 * WG_DSTRIDE, a_hi, a_lo, b_hi, b_lo, N3, WG_ROWS, swap_ls, acc0, bar);
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::fence_after();
  }
  // ---- epilogue: thread r (warps 0-3) = accumulator row r -> this CTA's gradient partial ----
  float* gp = a.gpart + (size_t)blockIdx.x * a.P;
  if (warp < 4) {
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  const bool any = iter > 0;      // a CTA without rows writes zeros (its TMEM was never written)
  const int r = tid;
  float v[64];
  float t[32];
  // D1: gates r, z
  if (any) tc::tmem_ld64(trow, v);
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4)
    *reinterpret_cast<float4*>(gp + L.wih + (size_t)r * MX_H + 4 * c4) = any ? make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (any) tc::tmem_ld64(trow + 64, v);
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4)
    *reinterpret_cast<float4*>(gp + L.whh + (size_t)r * MX_H + 4 * c4) = any ? make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (any) tc::tmem_ld32(trow + 128, t);
  gp[L.bih + r] = any ? t[0] : 0.f;
  gp[L.bhh + r] = any ? t[0] : 0.f;
  // D2: n gate.  rows 0-63: dW_ih[128 + r] = cols 0-63, db_ih ; rows 64-127: dW_hh[128 + r - 64] = cols 64-127, db_hh
  if (any) tc::tmem_ld64(trow + WG_DSTRIDE + (r < 64 ? 0 : 64), v);
  {
    float* dst = gp + (r < 64 ? L.wih : L.whh) + (size_t)(128 + (r & 63)) * MX_H;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4)
      *reinterpret_cast<float4*>(dst + 4 * c4) = any ? make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (any) tc::tmem_ld32(trow + WG_DSTRIDE + 128, t);
  gp[(r < 64 ? L.bih : L.bhh) + 128 + (r & 63)] = any ? t[0] : 0.f;
  // D3: rows 0-63: dW2[r] = cols 0-63, db2 ; rows 64-127: dW1[r - 64][0:I] = cols 64.., db1
  if (any) tc::tmem_ld64(trow + 2 * WG_DSTRIDE + (r < 64 ? 0 : 64), v);
  if (r < 64) {
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4)
      *reinterpret_cast<float4*>(gp + L.w2 + (size_t)r * MX_H + 4 * c4) = any ? make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    float* dst = gp + L.w1 + (size_t)(r - 64) * I;
#pragma unroll
    for (int c = 0; c < 64; ++c)
      if (c < I) dst[c] = any ? v[c] : 0.f;
  }
  if (any) tc::tmem_ld32(trow + 2 * WG_DSTRIDE + ones3, t);      // 32 columns from the ones column on (inside this accumulator's stride)
  gp[(r < 64 ? L.b2 : L.b1) + (r & 63)] = any ? t[0] : 0.f;
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

bool mx_wgrad_tc_usable(const FrontBwdArgs& a) {
  return g_mx_wgrad_tc && a.da2_out && a.da1_out && !a.no_gru && !a.skip_wgrad && a.L.in_dim <= 64 && a.M >= 1;
}

extern int g_mx_tc_swap;
int mx_launch_wgrad_tc(const FrontBwdArgs& a, int nparts, cudaStream_t s) {
  WgradArgs w;
  w.f = a;
  w.nchunks = mx_ceil_div(a.M, WG_ROWS);
  w.Kp16 = mx_round_up(a.L.in_dim, 16);
  WgradSmem sm = wgrad_smem();
#if !MX_EMU
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total) != cudaSuccess) { mx_set_error("wgrad_tc: smem %d too large", sm.total); return 1; }
    configured = true;
  }
#endif
  MX_LAUNCH_PDL(k_wgrad_tc, dim3(nparts), dim3(WG_THREADS), (size_t)sm.total, s, w, sm, g_mx_tc_swap);
  MX_COUNT();
  MX_MARK("k_wgrad_tc", s);
  return MX_CHECK_LAUNCH("wgrad_tc");
}
