// Shared-memory tiled FP32 micro-kernels for the time-batched dense layers (256 threads per CTA).
//
// Thread map: tx = tid & 15, ty = tid >> 4.  A CTA works on a tile of TM = 16*RM rows; thread (ty,tx) owns
// rows ty*RM .. ty*RM+RM-1.  Because ty is constant over 16 consecutive lanes, every ROW of a 64-wide
// output lives in one half-warp, so LayerNorm row statistics are 4 xor-shuffles (no shared memory).
//
// Weights are streamed through a 64-row staging chunk `Wc[64][ldw]` in their natural PyTorch layout
// W[n][k] (row n = output feature).  Leading dimensions are "round_up(K,8)+4" floats: every 128-bit
// shared load below is then bank-conflict free (odd multiple of 4 banks between consecutive rows).
//
//   NT  (forward):    acc[i][j] += sum_k A[r_i][k] * Wc[tx+16j][k]        (both operands k-contiguous)
//   NN  (data grad):  acc[i][j] += sum_n dY[r_i][n] * Wc[n][kb + 4tx + j] (contraction over the chunk rows)
//   TN  (weight grad): dW[4tn+i][4tk+j] = sum_r dY[r][nb+4tn+i] * X[r][kb+4tk+j]
#pragma once
#include "mx_common.cuh"

#define MX_TILE_THREADS 256

static inline int mx_ld(int k) { return mx_round_up(k, 8) + 4; }   // host helper: padded leading dimension
MX_DEVINL int mx_ld_dev(int k) { return ((k + 7) / 8) * 8 + 4; }

// Stage rows [row0, row0+64) x cols [col0, col0+ncols_pad) of a row-major global matrix W[nrows][ldg] into
// Wc[64][ldw] with cp.async (16-byte copies when rows are 16-byte aligned, 4-byte copies otherwise); rows / cols
// outside the matrix are zero-filled.  Returns after this thread's copies have landed; the caller's
// __syncthreads() publishes the chunk.
MX_DEVINL void mx_stage_weight(float* Wc, int ldw, const float* __restrict__ W, int nrows, int ncols_total, int ldg, int row0, int col0,
                               int ncols_pad, bool wait = true) {      // wait = false: one committed cp.async group, the caller waits (pipelined staging)
  const int tid = threadIdx.x;
  const bool vec = ((ldg & 3) == 0) && ((col0 & 3) == 0) && ((ncols_total & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (vec) {
    const int nc4 = ncols_pad >> 2;
    for (int r = tid >> 4; r < 64; r += MX_TILE_THREADS / 16) {
      const int gr = row0 + r;
      for (int c4 = tid & 15; c4 < nc4; c4 += 16) {
        const int gc = col0 + 4 * c4;
        float* dst = Wc + r * ldw + 4 * c4;
        if (gr < nrows && gc < ncols_total) mx_cp16(dst, W + (size_t)gr * ldg + gc);
        else mx_st4(dst, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  } else {
    for (int r = tid >> 5; r < 64; r += MX_TILE_THREADS / 32) {
      const int gr = row0 + r;
      for (int c = tid & 31; c < ncols_pad; c += 32) {
        const int gc = col0 + c;
        float* dst = Wc + r * ldw + c;
        if (gr < nrows && gc < ncols_total) mx_cp4(dst, W + (size_t)gr * ldg + gc);
        else *dst = 0.f;
      }
    }
  }
  mx_cp_commit();
  if (wait) mx_cp_wait<0>();
}

// Copy `nrows_tile` rows of `ncols` floats (ncols % 4 == 0, rows 16-byte aligned, row stride ldg) starting at global
// row m0 into dst[r][ldd]; rows >= M are zero-filled.  Asynchronous: caller commits / waits.
MX_DEVINL void mx_stage_rows(float* dst, int ldd, const float* __restrict__ src, size_t ldg, int m0, int M, int nrows_tile, int ncols) {
  const int tid = threadIdx.x;
  const int nc4 = ncols >> 2;
  for (int r = tid >> 4; r < nrows_tile; r += MX_TILE_THREADS / 16) {
    const int m = m0 + r;
    for (int c4 = tid & 15; c4 < nc4; c4 += 16) {
      float* d = dst + r * ldd + 4 * c4;
      if (m < M) mx_cp16(d, src + (size_t)m * ldg + 4 * c4);
      else mx_st4(d, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

template <int RM>
MX_DEVINL void mx_mm_nt(const float* __restrict__ A_s, int lda, const float* __restrict__ Wc, int ldw, int kpad, float (&acc)[RM][4]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* a0 = A_s + (ty * RM) * lda;
  const float* w0 = Wc + tx * ldw;
  for (int k = 0; k < kpad; k += 4) {
    float4 a[RM], w[4];
#pragma unroll
    for (int i = 0; i < RM; ++i) a[i] = mx_ld4(a0 + i * lda + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = mx_ld4(w0 + (16 * j) * ldw + k);
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
      }
  }
}

// dY_s: [TM][ldy]; contraction over chunk rows n = 0..63 (global n = nbase + n, caller offsets dY_s by nbase);
// Wc holds W[nbase+n][kb .. kb+63]; output columns 4tx .. 4tx+3 of the k-block.
template <int RM>
MX_DEVINL void mx_mm_nn(const float* __restrict__ dY_s, int ldy, const float* __restrict__ Wc, int ldw, float (&acc)[RM][4]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* d0 = dY_s + (ty * RM) * ldy;
  const float* w0 = Wc + tx * 4;
  for (int n = 0; n < 64; n += 4) {
    float4 d[RM], w[4];
#pragma unroll
    for (int i = 0; i < RM; ++i) d[i] = mx_ld4(d0 + i * ldy + n);
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = mx_ld4(w0 + (n + q) * ldw);
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      acc[i][0] = fmaf(d[i].x, w[0].x, acc[i][0]); acc[i][1] = fmaf(d[i].x, w[0].y, acc[i][1]);
      acc[i][2] = fmaf(d[i].x, w[0].z, acc[i][2]); acc[i][3] = fmaf(d[i].x, w[0].w, acc[i][3]);
      acc[i][0] = fmaf(d[i].y, w[1].x, acc[i][0]); acc[i][1] = fmaf(d[i].y, w[1].y, acc[i][1]);
      acc[i][2] = fmaf(d[i].y, w[1].z, acc[i][2]); acc[i][3] = fmaf(d[i].y, w[1].w, acc[i][3]);
      acc[i][0] = fmaf(d[i].z, w[2].x, acc[i][0]); acc[i][1] = fmaf(d[i].z, w[2].y, acc[i][1]);
      acc[i][2] = fmaf(d[i].z, w[2].z, acc[i][2]); acc[i][3] = fmaf(d[i].z, w[2].w, acc[i][3]);
      acc[i][0] = fmaf(d[i].w, w[3].x, acc[i][0]); acc[i][1] = fmaf(d[i].w, w[3].y, acc[i][1]);
      acc[i][2] = fmaf(d[i].w, w[3].z, acc[i][2]); acc[i][3] = fmaf(d[i].w, w[3].w, acc[i][3]);
    }
  }
}

// Weight gradient of one 64(n) x 64(k) output block over the TM rows of the tile.
//   dY_s: [TM][ldy] (caller offsets to column nb), X_s: [TM][ldx] (caller offsets to column kb)
//   dW:   global partial, row-major [N][K]; writes rows nb+4tn+i < N, cols kb+4tk+j < K.
MX_DEVINL void mx_wgrad_block(const float* __restrict__ dY_s, int ldy, const float* __restrict__ X_s, int ldx, int TM, float* __restrict__ dW,
                              int N, int K, int nb, int kb, bool accumulate) {
  const int tk = threadIdx.x & 15, tn = threadIdx.x >> 4;
  const int n0 = 4 * tn, k0 = 4 * tk;
  if (nb + n0 >= N || kb + k0 >= K) return;     // whole 4x4 block outside (no barriers inside: safe)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int r = 0; r < TM; ++r) {
    float4 d = mx_ld4(dY_s + r * ldy + n0);
    float4 x = mx_ld4(X_s + r * ldx + k0);
    acc[0][0] = fmaf(d.x, x.x, acc[0][0]); acc[0][1] = fmaf(d.x, x.y, acc[0][1]); acc[0][2] = fmaf(d.x, x.z, acc[0][2]); acc[0][3] = fmaf(d.x, x.w, acc[0][3]);
    acc[1][0] = fmaf(d.y, x.x, acc[1][0]); acc[1][1] = fmaf(d.y, x.y, acc[1][1]); acc[1][2] = fmaf(d.y, x.z, acc[1][2]); acc[1][3] = fmaf(d.y, x.w, acc[1][3]);
    acc[2][0] = fmaf(d.z, x.x, acc[2][0]); acc[2][1] = fmaf(d.z, x.y, acc[2][1]); acc[2][2] = fmaf(d.z, x.z, acc[2][2]); acc[2][3] = fmaf(d.z, x.w, acc[2][3]);
    acc[3][0] = fmaf(d.w, x.x, acc[3][0]); acc[3][1] = fmaf(d.w, x.y, acc[3][1]); acc[3][2] = fmaf(d.w, x.z, acc[3][2]); acc[3][3] = fmaf(d.w, x.w, acc[3][3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = nb + n0 + i;
    if (n >= N) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = kb + k0 + j;
      if (k < K) {
        float* p = dW + (size_t)n * K + k;
        *p = accumulate ? (*p + acc[i][j]) : acc[i][j];
      }
    }
  }
}

// Column sums of dY_s[TM][ldy] over the tile rows for columns [0, ncols): db[c] (+)= sum_r dY_s[r][c]
MX_DEVINL void mx_colsum(const float* __restrict__ dY_s, int ldy, int TM, int ncols, float* __restrict__ db, bool accumulate) {
  for (int c = threadIdx.x; c < ncols; c += MX_TILE_THREADS) {
    float s = 0.f;
    for (int r = 0; r < TM; ++r) s += dY_s[r * ldy + c];
    db[c] = accumulate ? (db[c] + s) : s;
  }
}

// Sum over the 16 lanes that share `ty` (one tile row): lanes differ in tx = lane & 15.
MX_DEVINL float mx_row16_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// LayerNorm over a full 64-wide row held as acc[i][0..3] by the 16 lanes of a half-warp (cols tx+16j or 4tx+j,
// the statistics do not care).  Returns mean / rstd per owned row.
template <int RM>
MX_DEVINL void mx_row_stats64(const float (&v)[RM][4], float (&mean)[RM], float (&rstd)[RM]) {
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    float s = mx_row16_sum(v[i][0] + v[i][1] + v[i][2] + v[i][3]);
    float m = s * (1.0f / 64.0f);
    float d0 = v[i][0] - m, d1 = v[i][1] - m, d2 = v[i][2] - m, d3 = v[i][3] - m;
    float q = mx_row16_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
    mean[i] = m;
    rstd[i] = rsqrtf(q * (1.0f / 64.0f) + MX_LN_EPS);
  }
}
