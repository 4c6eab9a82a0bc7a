// tile16.cuh -- forward / targets / backward of ONE 16-row tile of the train step in 256 threads, shared by
//   * train_cluster.cu: one agent over a 4-CTA cluster, 16 of the 64 rows of a tile per CTA, and
//   * train_fused.cu:   the tail of a batch of 65..80 rows (the sweep draws batch sizes up to 70,
//                       hyperparameter_optimization.py:121): 64 rows go through the 64-row tile code, the few rows that are
//                       left through this one at ~0.4x the cost of a second 64-row tile.
// All 48 "forward rows" -- (theta, s) | (theta, s') | (theta^-, s') -- run together; the backward produces a full-size
// partial gradient that is ADDED into G.  Same arithmetic as the 64-row code (q_learning_functions.py:42-64, :31-39, :23;
// dddqn.py:24-34); summation orders differ.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "tile_ops.cuh"

namespace dqn {
namespace t16 {

using namespace tile;

constexpr int R = 16;             // rows of the tile
constexpr int FR = 3 * R;         // 48 forward rows: [0,16) (theta,s) | [16,32) (theta,s') | [32,48) (theta^-,s')
constexpr int XS = 2 * R + 4;     // 36: row stride of X [d][col], cols [0,16) = s, [16,32) = s'
constexpr int HS = FR + 4;        // 52: row stride of H1 / H2 [unit][forward row]
constexpr int RS = R + 4;         // 20: row stride of the k-major backward buffers
constexpr int DRS = kH2 + 4;      // 68: row stride of Dh2R [row][unit]
constexpr int HC = kHeadCols;
constexpr int WS2 = kW2Stride;    // 68

// shared-memory footprint of the tile's activations (floats), in the order of `Bufs`
__host__ __device__ inline int floats(int D) {
  return (D + 1) * XS + kH1 * HS + (kH2 + 1) * HS + kH2 * RS + R * DRS + kH1 * RS + HC * RS + 4 * HC * FR + R * 4;
}

struct Bufs {
  const float* W;     // theta   (packed layout, common.cuh)
  const float* Wt;    // theta^-
  float* G;           // gradient accumulator (packed layout)
  float* X;           // [D+1][36]   row D = ones (the caller keeps it set)
  float* H1;          // [32][52]
  float* H2;          // [65][52]    row 64 = ones (the caller keeps it set)
  float* Dh2T;        // [64 j][20]
  float* Dh2R;        // [16 r][68]
  float* Dh1T;        // [32 k][20]
  float* DhdT;        // [8 c][20]
  float* Scr;         // [4 parts][8 c][48 forward rows]
  float* Meta;        // [16][4]  raw action lo, action hi, reward, done
  float* Red;         // [16..31] head-bias partials (scratch)
  int pW2, pWh, D;
  // carve the activation buffers out of `base` (floats(D) floats, 16-byte aligned)
  __device__ __forceinline__ void carve(float* base) {
    float* o = base;
    X = o; o += (D + 1) * XS;
    H1 = o; o += kH1 * HS;
    H2 = o; o += (kH2 + 1) * HS;
    Dh2T = o; o += kH2 * RS;
    Dh2R = o; o += R * DRS;
    Dh1T = o; o += kH1 * RS;
    DhdT = o; o += HC * RS;
    Scr = o; o += 4 * HC * FR;
    Meta = o;
  }
};

// X and Meta hold the tile's rows (rows >= B zero-filled).  Every thread of the 256-thread CTA calls this; it starts and
// ends without a barrier (the caller orders its own writes before, and reads of G after).  Returns, in threads 0..15, the
// sum of the tile's per-sample losses (not yet divided by B).
// NB = 0: the CTA has exactly 256 threads (barrier 0).  NB = 256: the first 256 threads of a larger CTA call this and meet on named
// barrier 1 (train_fused_tc.cu).
template <int NB>
__device__ __forceinline__ void tile_sync() {
  if constexpr (NB == 0) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(NB) : "memory");
}

template <int A, int NB = 0>
__device__ __forceinline__ float step(const Bufs& bf, int row_base, int B, float fB, float gamma, bool l2loss, const TapsDev& taps) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int D = bf.D;
  const float* const W = bf.W;
  const float* const Wt = bf.Wt;
  float* const G = bf.G;
  float* const X = bf.X;
  float* const H1 = bf.H1;
  float* const H2 = bf.H2;
  float* const Dh2T = bf.Dh2T;
  float* const Dh2R = bf.Dh2R;
  float* const Dh1T = bf.Dh1T;
  float* const DhdT = bf.DhdT;
  float* const Scr = bf.Scr;
  float* const Meta = bf.Meta;
  float* const Red = bf.Red;
  float loss_part = 0.f;
  // ---- forward, all 48 forward rows at once (threads 0..191: 12 row tiles x 16 column tiles) ----
  const int ct = t & 15, rt = t >> 4;
  const bool fwd = rt < FR / 4;
  const float* Wsel = rt < 8 ? W : Wt;                 // forward rows 32..47 use theta^-
  const int xcol = rt < 8 ? 4 * rt : 4 * (rt - 4);     // forward rows 32..47 read the s' columns again
  if (fwd) {   // layer 1: 4 rows x 2 units
    u64 acc[4][1];
    op_init<2>(Wsel + D * kH1 + 2 * ct, acc);
    op_tile4<2>(X + xcol, XS, Wsel + 2 * ct, kH1, D, acc);
    op_store_relu<2>(H1 + (2 * ct) * HS + 4 * rt, HS, acc);
  }
  tile_sync<NB>();
  if (fwd) {   // layer 2: 4 rows x 4 units
    u64 acc[4][2];
    op_init<4>(Wsel + bf.pW2 + kH1 * WS2 + 4 * ct, acc);
    op_tile4<4>(H1 + 4 * rt, HS, Wsel + bf.pW2 + 4 * ct, WS2, kH1, acc);
    op_store_relu<4>(H2 + (4 * ct) * HS + 4 * rt, HS, acc);
  }
  tile_sync<NB>();
  if (t < 4 * FR) {   // head: split-K over 4 thread groups, forward row fr = t % 48
    const int fr = t % FR, part = t / FR;
    const float* wh = (fr < 2 * R ? W : Wt) + bf.pWh;
    constexpr int NP = (A + 2) / 2;
    u64 acc2[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) acc2[q] = part == 0 ? ld64(wh + kH2 * HC + 2 * q) : 0ull;
#pragma unroll 8
    for (int k = 16 * part; k < 16 * part + 16; ++k) {
      const float h = H2[k * HS + fr];
      const u64 hh = pack2(h, h);
      const u64x2 w0 = ld2x64(wh + k * HC);
      ffma2(acc2[0], hh, w0.lo);
      if constexpr (NP > 1) ffma2(acc2[1], hh, w0.hi);
      if constexpr (NP > 2) { const u64x2 w1 = ld2x64(wh + k * HC + 4); ffma2(acc2[2], hh, w1.lo); if constexpr (NP > 3) ffma2(acc2[3], hh, w1.hi); }
    }
    float acc[2 * NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) unpack2(acc2[q], acc[2 * q], acc[2 * q + 1]);
#pragma unroll
    for (int c = 0; c <= A; ++c) Scr[(part * HC + c) * FR + fr] = acc[c];
  }
  tile_sync<NB>();
  if (t < R) {   // targets / Huber / d(head) for this CTA's 16 samples (half of warp 0)
    const int i = t;
    float hd[3][1 + A];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int c = 0; c <= A; ++c) {
        float v = Scr[(0 * HC + c) * FR + g * R + i];
#pragma unroll
        for (int p = 1; p < 4; ++p) v += Scr[(p * HC + c) * FR + g * R + i];
        hd[g][c] = v;
      }
    float q[A], nq[A], nqt[A];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float ms = 0.f;
#pragma unroll
      for (int j = 1; j <= A; ++j) ms += hd[g][j];
      ms = ms / (float)A;
#pragma unroll
      for (int j = 0; j < A; ++j) {
        const float v = hd[g][0] + hd[g][1 + j] - ms;
        if (g == 0) q[j] = v; else if (g == 1) nq[j] = v; else nqt[j] = v;
      }
    }
    int astar = 0; float best = nq[0];
#pragma unroll
    for (int j = 1; j < A; ++j) if (nq[j] > best) { best = nq[j]; astar = j; }
    const float4 meta = ld4(Meta + 4 * i);
    int a = __float_as_int(meta.x);
    a = a < 0 ? 0 : (a >= A ? A - 1 : a);
    const float rew = meta.z;
    const float done = __float_as_uint(meta.w) ? 1.f : 0.f;
    float qa = q[0], nt = nqt[0];
#pragma unroll
    for (int j = 1; j < A; ++j) { if (j == a) qa = q[j]; if (j == astar) nt = nqt[j]; }
    const float tv = rew + (1.0f - done) * (gamma * nt - qa);
    const float tgt = qa + tv;
    const float e = qa - tgt;
    const float ae = fabsf(e), quad = fminf(ae, 1.0f);
    const int grow = row_base + i;
    const bool valid = grow < B;
    const float l = valid ? (l2loss ? 0.5f * e * e : 0.5f * quad * quad + (ae - quad)) : 0.f;
    const float gi = valid ? (l2loss ? e : fminf(fmaxf(e, -1.0f), 1.0f)) / fB : 0.f;
    DhdT[0 * RS + i] = gi;
    float dsum[1 + A];
    dsum[0] = gi;
#pragma unroll
    for (int j = 0; j < A; ++j) {
      const float dadv = (j == a ? gi : 0.f) - gi / (float)A;
      DhdT[(1 + j) * RS + i] = dadv;
      dsum[1 + j] = dadv;
    }
    float lsum = l;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {      // reduce over the 16 active lanes
      lsum += __shfl_xor_sync(0x0000ffffu, lsum, o);
#pragma unroll
      for (int c = 0; c <= A; ++c) dsum[c] += __shfl_xor_sync(0x0000ffffu, dsum[c], o);
    }
    loss_part = lsum;
    if (i == 0) {
#pragma unroll
      for (int c = 0; c <= A; ++c) Red[16 + c] = dsum[c];
    }
    if (taps.enabled && valid) {
#pragma unroll
      for (int j = 0; j < A; ++j) {
        if (taps.q) taps.q[grow * A + j] = q[j];
        if (taps.next_q) taps.next_q[grow * A + j] = nq[j];
        if (taps.next_q_tm) taps.next_q_tm[grow * A + j] = nqt[j];
        if (taps.targets) taps.targets[grow * A + j] = (j == a) ? tgt : q[j];
      }
      if (taps.max_actions) taps.max_actions[grow] = astar;
    }
  }
  tile_sync<NB>();

  // ---- backward ----
  {  // dh2 (16 rows x 64 units): thread = row r, 4 units
    const int r = t & 15, jt = t >> 4;
    if (t <= A) G[bf.pWh + kH2 * HC + t] += Red[16 + t];
    float dh[1 + A];
#pragma unroll
    for (int c = 0; c <= A; ++c) dh[c] = DhdT[c * RS + r];
    float o[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * jt + jj;
      const float4 w0 = ld4(W + bf.pWh + j * HC), w1 = ld4(W + bf.pWh + j * HC + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float v = 0.f;
#pragma unroll
      for (int c = 0; c <= A; ++c) v = fmaf(dh[c], wv[c], v);
      o[jj] = H2[j * HS + r] > 0.f ? v : 0.f;
      Dh2T[j * RS + r] = o[jj];
    }
    st4(Dh2R + r * DRS + 4 * jt, o[0], o[1], o[2], o[3]);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
      for (int s = 8; s > 0; s >>= 1) o[jj] += __shfl_xor_sync(0xffffffffu, o[jj], s);
    }
    if (r == 0) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) G[bf.pW2 + kH1 * WS2 + 4 * jt + jj] += o[jj];
    }
  }
  tile_sync<NB>();
  {  // dW2: warp = 16 x 16 block, reduction over this CTA's 16 rows
    const int mi = lane & 7, ni = lane >> 3, k0 = 16 * (warp & 1) + mi, j0 = 16 * (warp >> 1) + ni;
    const float* ap[2] = {H1 + k0 * HS, H1 + (k0 + 8) * HS};
    const float* bp[4] = {Dh2T + j0 * RS, Dh2T + (j0 + 4) * RS, Dh2T + (j0 + 8) * RS, Dh2T + (j0 + 12) * RS};
    float acc[2][4];
    dot_tile<2, 4, 4>(ap, bp, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) G[bf.pW2 + (k0 + 8 * i) * WS2 + j0 + 4 * j] += acc[i][j];
  }
  {  // dWh: 4-way split over the 16 rows inside a warp
    const int j = warp * 8 + (lane & 7), part = lane >> 3;
    const float* ap[1] = {H2 + j * HS + 4 * part};
    const float* bp[1 + A];
#pragma unroll
    for (int c = 0; c <= A; ++c) bp[c] = DhdT + c * RS + 4 * part;
    float acc[1][1 + A];
    dot_tile<1, 1 + A, 1>(ap, bp, acc);
#pragma unroll
    for (int c = 0; c <= A; ++c) {
      float v = acc[0][c];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (part == 0) G[bf.pWh + j * HC + c] += v;
    }
  }
  {  // dh1 (16 rows x 32 units), reduction over 64 units
    const int r = t & 15, kk = t >> 4;
    const float* ap[1] = {Dh2R + r * DRS};
    const float* bp[2] = {W + bf.pW2 + kk * WS2, W + bf.pW2 + (kk + 16) * WS2};
    float acc[1][2];
    dot_tile<1, 2, 16>(ap, bp, acc);
    Dh1T[kk * RS + r] = H1[kk * HS + r] > 0.f ? acc[0][0] : 0.f;
    Dh1T[(kk + 16) * RS + r] = H1[(kk + 16) * HS + r] > 0.f ? acc[0][1] : 0.f;
  }
  tile_sync<NB>();
  {  // [dW1; db1]
    const int hcol = t & 31, mt = t >> 5;
    const int d1 = mt + 8 <= D ? mt + 8 : D, d2 = mt + 16 <= D ? mt + 16 : D;
    const float* bp[1] = {Dh1T + hcol * RS};
    const float* ap[3] = {X + mt * XS, X + d1 * XS, X + d2 * XS};
    float acc[3][1];
    dot_tile<3, 1, 4>(ap, bp, acc);
    if (mt <= D) G[mt * kH1 + hcol] += acc[0][0];
    if (mt + 8 <= D) G[(mt + 8) * kH1 + hcol] += acc[1][0];
    if (mt + 16 <= D) G[(mt + 16) * kH1 + hcol] += acc[2][0];
  }
  return loss_part;
}

}  // namespace t16
}  // namespace dqn
