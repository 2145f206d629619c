// Backward kernels of the recurrent agent network (BPTT through the GRU over the full episode).
//
//   k_qhead_bwd   d(q_taken) -> Linear head grads, post-GRU LayerNorm backward -> dL/dh_t (row-parallel)
//   k_gru_bwd     serial reverse recurrence: gate derivatives + dh_{t-1} = z*dh + W_hh^T dgh  (W_hh^T slices in registers)
//   k_front_bwd   time-batched: dW_ih/db_ih, dW_hh/db_hh, then back through LN2/fc2/LN1/fc1/LN0 with all
//                 weight gradients reduced per CTA into its gradient partial
//
// reference: autograd of nn.GRU / nn.LayerNorm / nn.Linear invoked by loss.backward() at qmix.py:191 for the
// module stack in algorithms/utils/{mlp,rnn,act}.py.  Gradients are numerators (see mixer.cu).
#include "mx_internal.h"
#include "mx_kernels.h"
#include "mx_tile.cuh"

// =====================================================================================================
// Q head backward (one warp per row-step)
// =====================================================================================================
__global__ void __launch_bounds__(256) k_qhead_bwd(QHeadBwdArgs a) {
  __shared__ float wq_s[32 * MX_H];
  __shared__ float dwq_s[32 * MX_H];
  __shared__ float dbq_s[32];
  __shared__ float dg_s[MX_H], db_s[MX_H], lg_s[MX_H], lb_s[MX_H];
  const int tid = threadIdx.x, lane = tid & 31;
  const int A = a.A;
  for (int i = tid; i < A * MX_H; i += blockDim.x) { wq_s[i] = a.theta[a.wq + i]; dwq_s[i] = 0.f; }
  for (int i = tid; i < 32; i += blockDim.x) dbq_s[i] = 0.f;
  for (int i = tid; i < MX_H; i += blockDim.x) { dg_s[i] = 0.f; db_s[i] = 0.f; lg_s[i] = a.theta[a.lno_g + i]; lb_s[i] = a.theta[a.lno_b + i]; }
  __syncthreads();
  const int wglobal = blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
  const int wtotal = gridDim.x * (blockDim.x >> 5);
  const int T1 = a.T + 1, N = a.N;
  float dg0 = 0.f, dg1 = 0.f, db0 = 0.f, db1 = 0.f;
  for (int m = wglobal; m < a.M; m += wtotal) {
    const int n = m % N;
    const int bt = m / N;
    const int t = bt % T1, b = bt / T1;
    float* out = a.dh_out + (size_t)m * MX_H;
    if (t >= a.T) {                       // Q at the bootstrap step only feeds the (detached) target
      out[lane] = 0.f; out[lane + 32] = 0.f;
      continue;
    }
    const size_t e = (size_t)b * a.T + t;
    const int act = a.act_idx[e * N + n];
    const float dqv = a.dq_taken[e * N + n];
    const float* h = a.hall + (size_t)m * MX_H;
    const float mean = a.sto[2 * (size_t)m], rstd = a.sto[2 * (size_t)m + 1];
    const float xh0 = (h[lane] - mean) * rstd, xh1 = (h[lane + 32] - mean) * rstd;
    const float y0 = xh0 * lg_s[lane] + lb_s[lane], y1 = xh1 * lg_s[lane + 32] + lb_s[lane + 32];
    const float dy0 = dqv * wq_s[act * MX_H + lane], dy1 = dqv * wq_s[act * MX_H + lane + 32];
    atomicAdd(&dwq_s[act * MX_H + lane], dqv * y0);
    atomicAdd(&dwq_s[act * MX_H + lane + 32], dqv * y1);
    if (lane == 0) atomicAdd(&dbq_s[act], dqv);
    dg0 += dy0 * xh0; dg1 += dy1 * xh1; db0 += dy0; db1 += dy1;
    // LayerNorm backward
    const float dx0 = dy0 * lg_s[lane], dx1 = dy1 * lg_s[lane + 32];
    const float c1 = mx_warp_sum(dx0 + dx1) * (1.f / MX_H);
    const float c2 = mx_warp_sum(dx0 * xh0 + dx1 * xh1) * (1.f / MX_H);
    out[lane] = rstd * (dx0 - c1 - xh0 * c2);
    out[lane + 32] = rstd * (dx1 - c1 - xh1 * c2);
  }
  atomicAdd(&dg_s[lane], dg0); atomicAdd(&dg_s[lane + 32], dg1);
  atomicAdd(&db_s[lane], db0); atomicAdd(&db_s[lane + 32], db1);
  __syncthreads();
  float* gp = a.gpart + (size_t)blockIdx.x * a.P;
  for (int i = tid; i < A * MX_H; i += blockDim.x) gp[a.wq + i] = dwq_s[i];
  for (int i = tid; i < A; i += blockDim.x) gp[a.bq + i] = dbq_s[i];
  for (int i = tid; i < MX_H; i += blockDim.x) { gp[a.lno_g + i] = dg_s[i]; gp[a.lno_b + i] = db_s[i]; }
}

// =====================================================================================================
// GRU backward through time
// =====================================================================================================
template <int RPC>
__global__ void __launch_bounds__(MX_G) k_gru_bwd(GruBwdArgs a) {
  __shared__ __align__(16) float dgh_s[RPC][MX_G];
  __shared__ float part_s[RPC][3][MX_H];
  const int j = threadIdx.x;
  const int p = j / MX_H, k = j % MX_H;      // this thread sums rows p*64..p*64+63 of W_hh^T column k
  const int row0 = blockIdx.x * RPC;
  float wT[MX_H];
#pragma unroll
  for (int jj = 0; jj < MX_H; ++jj) wT[jj] = a.theta[a.whh + (p * MX_H + jj) * MX_H + k];
  const int T1 = a.T + 1, N = a.N;
  // the gate phase is done by threads idx < RPC*64 (strided); each owns dh_carry for its (row, i) pairs
  float carry[(RPC * MX_H + MX_G - 1) / MX_G];
#pragma unroll
  for (int c = 0; c < (RPC * MX_H + MX_G - 1) / MX_G; ++c) carry[c] = 0.f;
  // zero the t == T rows of dgi
  for (int idx = j; idx < RPC * MX_G; idx += MX_G) {
    const int r = idx / MX_G, c = idx % MX_G;
    const int row = row0 + r;
    if (row < a.R) {
      const size_t mm = (((size_t)(row / N) * T1) + a.T) * N + (row % N);
      a.dgi[mm * MX_G + c] = 0.f;
    }
  }
  for (int t = a.T - 1; t >= 0; --t) {
    int ci = 0;
    for (int idx = j; idx < RPC * MX_H; idx += MX_G, ++ci) {
      const int r = idx / MX_H, i = idx % MX_H;
      const int row = row0 + r;
      float d_r = 0.f, d_z = 0.f, d_n = 0.f, d_hn = 0.f, dhz = 0.f;
      if (row < a.R) {
        const size_t mm = (((size_t)(row / N) * T1) + t) * N + (row % N);
        const float rg = a.gates[mm * MX_G + i], zg = a.gates[mm * MX_G + MX_H + i], ng = a.gates[mm * MX_G + 2 * MX_H + i];
        const float hn = a.hn[mm * MX_H + i];
        const float hp = t > 0 ? a.hall[(mm - N) * MX_H + i] : 0.f;
        const float dh = a.dh_out[mm * MX_H + i] + carry[ci];
        const float dn = dh * (1.f - zg);
        const float dz = dh * (hp - ng);
        d_n = dn * (1.f - ng * ng);              // d pre-activation of n
        d_z = dz * zg * (1.f - zg);
        d_r = d_n * hn * rg * (1.f - rg);
        d_hn = d_n * rg;                         // gradient reaching W_hn h + b_hn
        dhz = dh * zg;
        a.dgi[mm * MX_G + i] = d_r;
        a.dgi[mm * MX_G + MX_H + i] = d_z;
        a.dgi[mm * MX_G + 2 * MX_H + i] = d_n;
      }
      dgh_s[r][i] = d_r; dgh_s[r][MX_H + i] = d_z; dgh_s[r][2 * MX_H + i] = d_hn;
      carry[ci] = dhz;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPC; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int jj = 0; jj < MX_H; jj += 4) {
        const float4 d4 = mx_ld4(&dgh_s[r][p * MX_H + jj]);
        acc = fmaf(wT[jj], d4.x, acc);
        acc = fmaf(wT[jj + 1], d4.y, acc);
        acc = fmaf(wT[jj + 2], d4.z, acc);
        acc = fmaf(wT[jj + 3], d4.w, acc);
      }
      part_s[r][p][k] = acc;
    }
    __syncthreads();
    ci = 0;
    for (int idx = j; idx < RPC * MX_H; idx += MX_G, ++ci) {
      const int r = idx / MX_H, i = idx % MX_H;
      carry[ci] += part_s[r][0][i] + part_s[r][1][i] + part_s[r][2][i];
    }
    // (part_s / dgh_s are rewritten only after the next barrier pair)
  }
}

// =====================================================================================================
// front backward (time-batched)
// =====================================================================================================
struct FrontBwdSmem {
  int ldi, ld64, ldg;
  int o_dgi, o_dgn, o_x, o_hp, o_u, o_da, o_x0, o_xh0, o_wc, o_col, total;
};
static FrontBwdSmem front_bwd_smem(int in_dim, int TM) {
  FrontBwdSmem s;
  const int I64 = mx_round_up(in_dim, 64);
  s.ldi = mx_ld(I64); s.ld64 = mx_ld(64); s.ldg = mx_ld(MX_G);
  int o = 0;
  s.o_dgi = o; o += TM * s.ldg;     // dgi tile (r,z,n)
  s.o_dgn = o; o += TM * s.ld64;    // dgi_n * r  (gradient reaching W_hn h)
  s.o_x = o; o += TM * s.ld64;      // x2, later x1
  s.o_hp = o; o += TM * s.ld64;     // h_{t-1}
  s.o_u = o; o += TM * s.ld64;      // u2, later u1 (post-ReLU, pre-LN)
  s.o_da = o; o += TM * s.ld64;     // gradient w.r.t. the Linear output (after ReLU mask)
  s.o_x0 = o; o += TM * s.ldi;      // LN0 output (fc1 input)
  s.o_xh0 = o; o += TM * s.ldi;     // normalised input before the affine
  s.o_wc = o; o += 64 * s.ld64;
  s.o_col = o; o += 2 * I64 + 4 * 64;   // column accumulators for LayerNorm gains/biases
  s.total = o;
  return s;
}

// LayerNorm backward on a 64-wide row spread over a half-warp; v = upstream grad w.r.t. LN output (cols 4tx+j),
// u = LN input.  Returns grad w.r.t. LN input, masked by ReLU (u > 0), and accumulates dgamma/dbeta column sums.
template <int RM>
MX_DEVINL void ln64_bwd_relu(float (&v)[RM][4], const float* u_s, int ld, const float* st, int m0, int M, const float* gamma, float* dg_col,
                             float* db_col) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int r = ty * RM + i, m = m0 + r;
    float mean = 0.f, rstd = 0.f;
    if (m < M) { mean = st[2 * (size_t)m]; rstd = st[2 * (size_t)m + 1]; }
    float xh[4], dx[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * tx + j;
      const float uu = u_s[r * ld + c];
      xh[j] = (uu - mean) * rstd;
      atomicAdd(&dg_col[c], v[i][j] * xh[j]);
      atomicAdd(&db_col[c], v[i][j]);
      dx[j] = v[i][j] * gamma[c];
      s1 += dx[j]; s2 += dx[j] * xh[j];
    }
    s1 = mx_row16_sum(s1) * (1.f / 64.f);
    s2 = mx_row16_sum(s2) * (1.f / 64.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float uu = u_s[r * ld + 4 * tx + j];
      const float du = rstd * (dx[j] - s1 - xh[j] * s2);
      v[i][j] = uu > 0.f ? du : 0.f;
    }
  }
}

template <int RM>
__global__ void __launch_bounds__(MX_TILE_THREADS) k_front_bwd(FrontBwdArgs a, FrontBwdSmem sm) {
  constexpr int TM = 16 * RM;
  MX_DYN_SMEM(smem);
  const MxNetLayout L = a.L;
  const float* __restrict__ th = a.theta;
  const int I = L.in_dim, I64 = mx_round_up(I, 64);
  float* dgi_s = smem + sm.o_dgi; float* dgn_s = smem + sm.o_dgn; float* x_s = smem + sm.o_x; float* hp_s = smem + sm.o_hp;
  float* u_s = smem + sm.o_u; float* da_s = smem + sm.o_da; float* x0_s = smem + sm.o_x0; float* xh0_s = smem + sm.o_xh0;
  float* Wc = smem + sm.o_wc;
  float* col0g = smem + sm.o_col; float* col0b = col0g + I64;          // LN0 gain / bias grads   [I64] each
  float* col1g = col0b + I64; float* col1b = col1g + 64;               // LN1
  float* col2g = col1b + 64; float* col2b = col2g + 64;                // LN2
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ntiles = (a.M + TM - 1) / TM;
  const int T1 = a.T + 1, N = a.N;
  float* gp = a.gpart + (size_t)blockIdx.x * a.P;
  for (int i = tid; i < 2 * I64 + 4 * 64; i += MX_TILE_THREADS) col0g[i] = 0.f;
  int iter = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    const int m0 = tile * TM;
    const bool accum = iter > 0;
    __syncthreads();
    // ---- stage the tile: dgi, dgi_n*r, h_{t-1}, u2, x2 = LN2(u2) ----
    for (int idx = tid; idx < TM * MX_G; idx += MX_TILE_THREADS) {
      const int r = idx / MX_G, c = idx % MX_G, m = m0 + r;
      dgi_s[r * sm.ldg + c] = m < a.M ? a.dgi[(size_t)m * MX_G + c] : 0.f;
    }
    for (int idx = tid; idx < TM * MX_H; idx += MX_TILE_THREADS) {
      const int r = idx / MX_H, c = idx % MX_H, m = m0 + r;
      float dn = 0.f, hp = 0.f, u2 = 0.f, x2 = 0.f;
      if (m < a.M) {
        dn = a.dgi[(size_t)m * MX_G + 2 * MX_H + c] * a.gates[(size_t)m * MX_G + c];
        const int t = (m / N) % T1;
        if (t > 0) hp = a.hall[(size_t)(m - N) * MX_H + c];
        u2 = a.u2[(size_t)m * MX_H + c];
        x2 = (u2 - a.st2[2 * (size_t)m]) * a.st2[2 * (size_t)m + 1] * th[L.ln2_g + c] + th[L.ln2_b + c];
      }
      dgn_s[r * sm.ld64 + c] = dn; hp_s[r * sm.ld64 + c] = hp; u_s[r * sm.ld64 + c] = u2; x_s[r * sm.ld64 + c] = x2;
    }
    __syncthreads();
    // ---- GRU weight gradients: dW_ih = dgi^T x2 ; dW_hh = [dgi_r, dgi_z, dgi_n*r]^T h_{t-1} ----
    for (int nb = 0; nb < 3; ++nb) {
      mx_wgrad_block(dgi_s + nb * 64, sm.ldg, x_s, sm.ld64, TM, gp + L.wih, MX_G, MX_H, nb * 64, 0, accum);
      const float* dgh = nb < 2 ? dgi_s + nb * 64 : dgn_s;
      const int ldd = nb < 2 ? sm.ldg : sm.ld64;
      // rows nb*64.. of W_hh: reuse the block routine on the 64-row slice
      mx_wgrad_block(dgh, ldd, hp_s, sm.ld64, TM, gp + L.whh + nb * 64 * MX_H, 64, MX_H, 0, 0, accum);
    }
    mx_colsum(dgi_s, sm.ldg, TM, MX_G, gp + L.bih, accum);
    mx_colsum(dgi_s, sm.ldg, TM, 2 * MX_H, gp + L.bhh, accum);
    mx_colsum(dgn_s, sm.ld64, TM, MX_H, gp + L.bhh + 2 * MX_H, accum);
    // ---- dx2 = dgi . W_ih  (three 64-row chunks) ----
    float v[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) v[i][jx] = 0.f;
    for (int nc = 0; nc < 3; ++nc) {
      __syncthreads();
      mx_stage_weight(Wc, sm.ld64, th + L.wih, MX_G, MX_H, MX_H, nc * 64, 0, 64);
      __syncthreads();
      mx_mm_nn<RM>(dgi_s + nc * 64, sm.ldg, Wc, sm.ld64, v);
    }
    // ---- LN2 backward + ReLU mask -> da2 ----
    ln64_bwd_relu<RM>(v, u_s, sm.ld64, a.st2, m0, a.M, th + L.ln2_g, col2g, col2b);
    __syncthreads();     // x_s (x2), u_s (u2) no longer needed by anyone
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) da_s[(ty * RM + i) * sm.ld64 + 4 * tx + jx] = v[i][jx];
    // stage u1 and x1 = LN1(u1)
    for (int idx = tid; idx < TM * MX_H; idx += MX_TILE_THREADS) {
      const int r = idx / MX_H, c = idx % MX_H, m = m0 + r;
      float u1 = 0.f, x1 = 0.f;
      if (m < a.M) {
        u1 = a.u1[(size_t)m * MX_H + c];
        x1 = (u1 - a.st1[2 * (size_t)m]) * a.st1[2 * (size_t)m + 1] * th[L.ln1_g + c] + th[L.ln1_b + c];
      }
      u_s[r * sm.ld64 + c] = u1; x_s[r * sm.ld64 + c] = x1;
    }
    __syncthreads();
    // ---- fc2: dW2 = da2^T x1, db2 ; dx1 = da2 . W2 ----
    mx_wgrad_block(da_s, sm.ld64, x_s, sm.ld64, TM, gp + L.w2, MX_H, MX_H, 0, 0, accum);
    mx_colsum(da_s, sm.ld64, TM, MX_H, gp + L.b2, accum);
    mx_stage_weight(Wc, sm.ld64, th + L.w2, MX_H, MX_H, MX_H, 0, 0, 64);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) v[i][jx] = 0.f;
    mx_mm_nn<RM>(da_s, sm.ld64, Wc, sm.ld64, v);
    ln64_bwd_relu<RM>(v, u_s, sm.ld64, a.st1, m0, a.M, th + L.ln1_g, col1g, col1b);
    __syncthreads();     // da_s (da2) consumed by everyone
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) da_s[(ty * RM + i) * sm.ld64 + 4 * tx + jx] = v[i][jx];
    // stage x0 = LN0(x) and the normalised input
    for (int idx = tid; idx < TM * I64; idx += MX_TILE_THREADS) {
      const int r = idx / I64, c = idx % I64, m = m0 + r;
      float xh = 0.f, x0 = 0.f;
      if (m < a.M && c < I) {
        const float x = a.X[(size_t)m * a.ldx + c];
        if (a.feature_norm) {
          xh = (x - a.st0[2 * (size_t)m]) * a.st0[2 * (size_t)m + 1];
          x0 = xh * th[L.fn_g + c] + th[L.fn_b + c];
        } else x0 = x;
      }
      x0_s[r * sm.ldi + c] = x0; xh0_s[r * sm.ldi + c] = xh;
    }
    __syncthreads();
    // ---- fc1: dW1 = da1^T x0, db1 ; dx0 = da1 . W1 (only for the LN0 gain/bias) ----
    for (int kb = 0; kb * 64 < I; ++kb) mx_wgrad_block(da_s, sm.ld64, x0_s + kb * 64, sm.ldi, TM, gp + L.w1, MX_H, I, 0, kb * 64, accum);
    mx_colsum(da_s, sm.ld64, TM, MX_H, gp + L.b1, accum);
    if (a.feature_norm) {
      for (int kb = 0; kb * 64 < I; ++kb) {
        __syncthreads();
        mx_stage_weight(Wc, sm.ld64, th + L.w1, MX_H, I, I, 0, kb * 64, 64);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) v[i][jx] = 0.f;
        mx_mm_nn<RM>(da_s, sm.ld64, Wc, sm.ld64, v);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) {
            const int c = kb * 64 + 4 * tx + jx;
            if (c < I) {
              atomicAdd(&col0g[c], v[i][jx] * xh0_s[(ty * RM + i) * sm.ldi + c]);
              atomicAdd(&col0b[c], v[i][jx]);
            }
          }
      }
    }
  }
  __syncthreads();
  // ---- LayerNorm gain / bias gradients accumulated over all tiles of this CTA ----
  for (int c = tid; c < MX_H; c += MX_TILE_THREADS) {
    gp[L.ln2_g + c] = col2g[c]; gp[L.ln2_b + c] = col2b[c];
    gp[L.ln1_g + c] = col1g[c]; gp[L.ln1_b + c] = col1b[c];
  }
  for (int c = tid; c < I; c += MX_TILE_THREADS) { gp[L.fn_g + c] = col0g[c]; gp[L.fn_b + c] = col0b[c]; }
  if (iter == 0) {
    // a CTA without tiles still owns a partial: publish zeros for the slices this kernel is responsible for
    for (int i = tid; i < MX_H * I; i += MX_TILE_THREADS) gp[L.w1 + i] = 0.f;
    for (int i = tid; i < MX_H * MX_H; i += MX_TILE_THREADS) gp[L.w2 + i] = 0.f;
    for (int i = tid; i < MX_G * MX_H; i += MX_TILE_THREADS) { gp[L.wih + i] = 0.f; gp[L.whh + i] = 0.f; }
    for (int i = tid; i < MX_G; i += MX_TILE_THREADS) { gp[L.bih + i] = 0.f; gp[L.bhh + i] = 0.f; }
    for (int i = tid; i < MX_H; i += MX_TILE_THREADS) { gp[L.b1 + i] = 0.f; gp[L.b2 + i] = 0.f; }
  }
}

// =====================================================================================================
// launchers
// =====================================================================================================
int mx_launch_qhead_bwd(const QHeadBwdArgs& a, int* nparts_used, cudaStream_t s) {
  int grid = mx_ceil_div(a.M, 8 * 4);   // ~4 rows per warp
  const int cap = mx_num_sms();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  MX_LAUNCH(k_qhead_bwd, dim3(grid), dim3(256), 0, s, a);
  MX_COUNT();
  MX_MARK("k_qhead_bwd", s);
  *nparts_used = grid;
  return MX_CHECK_LAUNCH("qhead_bwd");
}

int mx_launch_gru_bwd(const GruBwdArgs& a, cudaStream_t s) {
  const int sms = mx_num_sms();
  int rpc = 1;
  while (rpc < 4 && mx_ceil_div(a.R, rpc) > 2 * sms) rpc *= 2;
  dim3 grid(mx_ceil_div(a.R, rpc));
  if (rpc == 1) MX_LAUNCH(k_gru_bwd<1>, grid, dim3(MX_G), 0, s, a);
  else if (rpc == 2) MX_LAUNCH(k_gru_bwd<2>, grid, dim3(MX_G), 0, s, a);
  else MX_LAUNCH(k_gru_bwd<4>, grid, dim3(MX_G), 0, s, a);
  MX_COUNT();
  MX_MARK("k_gru_bwd", s);
  return MX_CHECK_LAUNCH("gru_bwd");
}

int mx_launch_front_bwd(const FrontBwdArgs& a, int* nparts_used, cudaStream_t s) {
  const int RM = 2, TM = 16 * RM;
  FrontBwdSmem sm = front_bwd_smem(a.L.in_dim, TM);
  const size_t smem = (size_t)sm.total * sizeof(float) + 16;
  const int ntiles = mx_ceil_div(a.M, TM);
  int grid = mx_num_sms();
  if (grid > ntiles) grid = ntiles;
  auto kern = k_front_bwd<2>;
#if !MX_EMU
  if (smem > 227 * 1024) { mx_set_error("front_bwd: %zu bytes of shared memory needed (obs_dim too large)", smem); return 1; }
  static size_t configured = 0;
  if (smem > configured) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
#endif
  MX_LAUNCH(kern, dim3(grid), dim3(MX_TILE_THREADS), smem, s, a, sm);
  MX_COUNT();
  MX_MARK("k_front_bwd", s);
  *nparts_used = grid;
  return MX_CHECK_LAUNCH("front_bwd");
}
