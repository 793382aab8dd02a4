// H36M evaluation math on sm_100a: MPJPE (reference valid_bilinear.py:53-60, per-action fp64
// accumulation of :61-70) and Procrustes-aligned P-MPJPE (not in the reference; SURVEY.md
// Appendix B gives the specification the oracle follows).
//
// Both are HBM-bound: 2 * J * 3 * 4 bytes are read per pose (384 B at J = 16) and at most 8 bytes
// written.  Layout of the work:
//   - a persistent grid (3 CTAs per SM for the TMA-fed J = 16 kernel, 4 for the generic one) walks chunks of 128 poses;
//   - the 128 threads of a CTA stream the chunk with 128-bit coalesced loads (24 float4 in flight per
//     thread) into shared memory with an odd row pitch (3J+1 words -> conflict-free per-pose reads);
//   - then ONE THREAD PER POSE pulls its 2*3J values into registers and does the whole per-pose
//     computation there (no cross-lane traffic): centroids, 3x3 correlation, optimal rotation via
//     Horn's quaternion form -- largest eigenvalue of the 4x4 key matrix by Newton iteration on its
//     characteristic quartic (the QCP method), eigenvector from the adjugate -- scale, aligned
//     distances.  Degenerate poses (repeated top eigenvalue, zero extent) fall back to a cyclic
//     Jacobi eigen-solver / the unaligned distance, so the result is always finite.
//   - per-action sums are staged in fixed-point (1/256 mm) shared-memory buckets per chunk, collected into fp64 per-CTA totals and
//     flushed with one double atomic per action per CTA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "ptx.cuh"

namespace b200 {

constexpr int kEvalThreads = 128;
constexpr int kMaxActions = 64;

struct EvalParams {
  const float* pred;
  const float* gt;
  const float* mean;    // [3J] or nullptr
  const float* stddev;  // [3J] or nullptr
  const int32_t* action;
  float* dist_out;
  float* pdist_out;
  double* acc_sum;
  double* pacc_sum;
  long long* acc_cnt;
  int64_t n;
  int32_t n_actions;
};

// ---- 4x4 symmetric eigen fallback (cyclic Jacobi), only for degenerate poses ------------------
__device__ __noinline__ void jacobi4_top(const float* Nm, float& lam, float* q) {
  float A[4][4], V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { A[i][j] = Nm[i * 4 + j]; V[i][j] = (i == j) ? 1.f : 0.f; }
  for (int sweep = 0; sweep < 8; ++sweep) {
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 4; ++r) {
        const float apq = A[p][r];
        if (fabsf(apq) < 1e-20f) continue;
        const float theta = (A[r][r] - A[p][p]) / (2.f * apq);
        const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
        const float c = rsqrtf(t * t + 1.f), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const float akp = A[k][p], akq = A[k][r];
          A[k][p] = c * akp - s * akq;
          A[k][r] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          const float apk = A[p][k], aqk = A[r][k];
          A[p][k] = c * apk - s * aqk;
          A[r][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          const float vkp = V[k][p], vkq = V[k][r];
          V[k][p] = c * vkp - s * vkq;
          V[k][r] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (A[i][i] > A[best][best]) best = i;
  lam = A[best][best];
  for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}

// sqrt(x) for x >= 0: one MUFU.SQRT (sqrt.approx.ftz.f32, ~1 ulp, exact 0 for 0; distances are ~1e1..1e3 mm and the
// tolerance is 0.05 mm).  The x * rsqrt(x) form it replaces cost a compare, a select and a multiply on top, 32 times
// per pose.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float det3(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
  return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// Largest eigenvalue / eigenvector of Horn's key matrix for the normalised correlation S (3x3,
// S[a][b] = sum_j p_a g_b with |P0|_F = |G0|_F = 1).  Returns lambda_max = s1 + s2 + sign(det) s3.
__device__ __forceinline__ void horn_top(const float S[3][3], float& lam, float q[4]) {
  const float Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2];
  const float Syx = S[1][0], Syy = S[1][1], Syz = S[1][2];
  const float Szx = S[2][0], Szy = S[2][1], Szz = S[2][2];
  float Nm[16];
  Nm[0] = Sxx + Syy + Szz; Nm[1] = Syz - Szy;        Nm[2] = Szx - Sxz;         Nm[3] = Sxy - Syx;
  Nm[4] = Nm[1];           Nm[5] = Sxx - Syy - Szz;  Nm[6] = Sxy + Syx;         Nm[7] = Szx + Sxz;
  Nm[8] = Nm[2];           Nm[9] = Nm[6];            Nm[10] = -Sxx + Syy - Szz; Nm[11] = Syz + Szy;
  Nm[12] = Nm[3];          Nm[13] = Nm[7];           Nm[14] = Nm[11];           Nm[15] = -Sxx - Syy + Szz;

  // characteristic polynomial l^4 + c2 l^2 + c1 l + c0 (the key matrix is traceless)
  const float c2 = -2.f * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy + Syz * Syz + Szx * Szx +
                           Szy * Szy + Szz * Szz);
  const float c1 = -8.f * det3(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz);
  const float a = Nm[0], b = Nm[1], c = Nm[2], d = Nm[3], e = Nm[5], f = Nm[6], g = Nm[7], h = Nm[10], i = Nm[11],
              j = Nm[15];
  // det of the symmetric 4x4 [[a b c d][b e f g][c f h i][d g i j]] by 2x2 minors of rows (0,1)x(2,3)
  const float s0 = a * e - b * b, s1 = a * f - b * c, s2 = a * g - b * d, s3 = b * f - e * c, s4 = b * g - e * d,
              s5 = c * g - f * d;
  const float t5 = h * j - i * i, t4 = f * j - i * g, t3 = f * i - h * g, t2 = c * j - i * d, t1 = c * i - h * d,
              t0 = c * g - f * d;
  const float c0 = s0 * t5 - s1 * t4 + s2 * t3 + s3 * t2 - s4 * t1 + s5 * t0;

  // Newton from the upper bound (|P0|^2 + |G0|^2)/2 = 1: monotone convergence to the top root
  float l = 1.0f;
#pragma unroll 1
  for (int it = 0; it < 24; ++it) {
    const float l2 = l * l;
    const float pv = (l2 + c2) * l2 + c1 * l + c0;
    const float dv = (4.f * l2 + 2.f * c2) * l + c1;
    if (!(fabsf(dv) > 1e-30f)) break;
    const float step = __fdividef(pv, dv);   // Newton corrects its own rounding: an approximate quotient is enough
    l -= step;
    if (fabsf(step) < 1e-7f * fabsf(l)) break;
  }
  lam = l;

  // eigenvector: a column of adj(N - lam I) (= const * q q^T when the top eigenvalue is simple)
  const float m00 = Nm[0] - l, m11 = Nm[5] - l, m22 = Nm[10] - l, m33 = Nm[15] - l;
  const float m01 = Nm[1], m02 = Nm[2], m03 = Nm[3], m12 = Nm[6], m13 = Nm[7], m23 = Nm[11];
  // First column alone, when it is a good one: adj = c q q^T, so C00 = c q0^2 and |column 0|^2 = c^2 q0^2, i.e.
  // C00^2 / |column 0|^2 = q0^2 -- the column's conditioning is known without the other three diagonals.  A pose that
  // needs less than a ~143 degree turn (q0^2 > 0.1) is done after 4 of the 10 cofactors; everything else (and a vanishing
  // adjugate = repeated top eigenvalue) takes the full pivoted path below.
  const float C00 = det3(m11, m12, m13, m12, m22, m23, m13, m23, m33);
  {
    const float C01 = -det3(m01, m12, m13, m02, m22, m23, m03, m23, m33);
    const float C02 = det3(m01, m11, m13, m02, m12, m23, m03, m13, m33);
    const float C03 = -det3(m01, m11, m12, m02, m12, m22, m03, m13, m23);
    const float nn0 = C00 * C00 + C01 * C01 + C02 * C02 + C03 * C03;
    if (nn0 > 1e-16f && C00 * C00 > 0.1f * nn0) {
      const float r = rsqrtf(nn0);
      q[0] = C00 * r; q[1] = C01 * r; q[2] = C02 * r; q[3] = C03 * r;
      return;
    }
  }
  // diagonal cofactors
  const float C11 = det3(m00, m02, m03, m02, m22, m23, m03, m23, m33);
  const float C22 = det3(m00, m01, m03, m01, m11, m13, m03, m13, m33);
  const float C33 = det3(m00, m01, m02, m01, m11, m12, m02, m12, m22);
  // off-diagonal cofactors C_ij = (-1)^(i+j) det(minor_ij)
  const float C01 = -det3(m01, m12, m13, m02, m22, m23, m03, m23, m33);
  const float C02 = det3(m01, m11, m13, m02, m12, m23, m03, m13, m33);
  const float C03 = -det3(m01, m11, m12, m02, m12, m22, m03, m13, m23);
  const float C12 = -det3(m00, m01, m03, m02, m12, m23, m03, m13, m33);
  const float C13 = det3(m00, m01, m02, m02, m12, m22, m03, m13, m23);
  const float C23 = -det3(m00, m01, m02, m01, m11, m12, m03, m13, m23);
  const float a0 = fabsf(C00), a1 = fabsf(C11), a2 = fabsf(C22), a3 = fabsf(C33);
  float q0, q1, q2, q3;
  if (a0 >= a1 && a0 >= a2 && a0 >= a3) { q0 = C00; q1 = C01; q2 = C02; q3 = C03; }
  else if (a1 >= a2 && a1 >= a3) { q0 = C01; q1 = C11; q2 = C12; q3 = C13; }
  else if (a2 >= a3) { q0 = C02; q1 = C12; q2 = C22; q3 = C23; }
  else { q0 = C03; q1 = C13; q2 = C23; q3 = C33; }
  const float nn = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
  if (nn > 1e-16f) {
    const float r = rsqrtf(nn);
    q[0] = q0 * r; q[1] = q1 * r; q[2] = q2 * r; q[3] = q3 * r;
  } else {
    jacobi4_top(Nm, lam, q);  // repeated top eigenvalue (collinear / reflected-degenerate pose)
    const float r = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= r; q[1] *= r; q[2] *= r; q[3] *= r;
  }
}

// Per-pose metrics from un-normalised coordinates (mm).  P and G are modified (centred) when DO_P.
template <int J, bool DO_M, bool DO_P>
__device__ __forceinline__ void pose_metrics(float* P, float* G, float& dist, float& pdist) {
  constexpr int W = 3 * J;
  dist = 0.f;
  pdist = 0.f;
  if (DO_M) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float dx = P[3 * j] - G[3 * j], dy = P[3 * j + 1] - G[3 * j + 1], dz = P[3 * j + 2] - G[3 * j + 2];
      dist += fast_sqrt(dx * dx + dy * dy + dz * dz);    // valid_bilinear.py:60
    }
  }
  if (DO_P) {
    float mp[3] = {0.f, 0.f, 0.f}, mg[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int a = 0; a < 3; ++a) { mp[a] += P[3 * j + a]; mg[a] += G[3 * j + a]; }
#pragma unroll
    for (int a = 0; a < 3; ++a) { mp[a] *= (1.f / J); mg[a] *= (1.f / J); }
    float nP2 = 0.f, nG2 = 0.f;
    float S[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float pc[3], gc[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        pc[a] = P[3 * j + a] - mp[a];
        gc[a] = G[3 * j + a] - mg[a];
        P[3 * j + a] = pc[a];
        G[3 * j + a] = gc[a];
        nP2 += pc[a] * pc[a];
        nG2 += gc[a] * gc[a];
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a][b] += pc[a] * gc[b];
    }
    // zero-extent convention shared with oracle/pose_metrics.py (DEGENERATE_REL); the raw second moment
    // is sum p^2 = sum (p - m)^2 + J |m|^2
    const float rawP2 = nP2 + float(J) * (mp[0] * mp[0] + mp[1] * mp[1] + mp[2] * mp[2]);
    const float rawG2 = nG2 + float(J) * (mg[0] * mg[0] + mg[1] * mg[1] + mg[2] * mg[2]);
    if (nP2 > 1e-10f * rawP2 && nG2 > 1e-10f * rawG2) {
      const float inv = rsqrtf(nP2) * rsqrtf(nG2);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a][b] *= inv;
      float lam, q[4];
      horn_top(S, lam, q);
      const float scale = lam * sqrtf(nG2) * rsqrtf(nP2);   // tr * nG / nP (Appendix B)
      const float q0 = q[0], qx = q[1], qy = q[2], qz = q[3];
      // scale * R (the similarity transform applied to the centred prediction)
      const float s2 = 2.f * scale;
      const float R00 = scale * (q0 * q0 + qx * qx - qy * qy - qz * qz), R01 = s2 * (qx * qy - q0 * qz),
                  R02 = s2 * (qx * qz + q0 * qy);
      const float R10 = s2 * (qy * qx + q0 * qz), R11 = scale * (q0 * q0 - qx * qx + qy * qy - qz * qz),
                  R12 = s2 * (qy * qz - q0 * qx);
      const float R20 = s2 * (qz * qx - q0 * qy), R21 = s2 * (qz * qy + q0 * qx),
                  R22 = scale * (q0 * q0 - qx * qx - qy * qy + qz * qz);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float px = P[3 * j], py = P[3 * j + 1], pz = P[3 * j + 2];
        const float ax = (R00 * px + R01 * py + R02 * pz) - G[3 * j];
        const float ay = (R10 * px + R11 * py + R12 * pz) - G[3 * j + 1];
        const float az = (R20 * px + R21 * py + R22 * pz) - G[3 * j + 2];
        pdist += fast_sqrt(ax * ax + ay * ay + az * az);
      }
    } else {
      // zero-extent pose: alignment undefined -> unaligned distance of the (uncentred) poses
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float dx = (P[3 * j] + mp[0]) - (G[3 * j] + mg[0]);
        const float dy = (P[3 * j + 1] + mp[1]) - (G[3 * j + 1] + mg[1]);
        const float dz = (P[3 * j + 2] + mp[2]) - (G[3 * j + 2] + mg[2]);
        pdist += fast_sqrt(dx * dx + dy * dy + dz * dz);
      }
    }
  }
}

// ---- J = 16, packed-pair form of pose_metrics (TMA kernel) ------------------------------------------------------
// sm_100 has two-wide fp32 instructions (FADD2 / FMUL2 / FFMA2 on an aligned register pair, with per-operand negation and a
// scalar-broadcast operand form).  A pose row [x0 y0 z0 x1 y1 z1 ...] read with 128-bit shared-memory loads already sits
// in such pairs -- (x, y), (z, x'), (y', z') for every two joints -- so everything that treats the 48 numbers alike
// (un-normalise, centroid, centring, norms, differences) runs on 24 pairs, and the two places that mix coordinates keep
// pair ACCUMULATORS instead: the correlation matrix as column pairs (S0b, S1b) += (px, py) * gb for even joints and
// (S1b, S2b) += (py', pz') * gb' for odd ones (5 instead of 9 instructions per joint), the aligned residual as
// (ax, ay) = (R00, R10) px + (R01, R11) py + (R02, R12) pz - (gx, gy) (3 + 3 instead of 12).  ncu had the scalar form at
// 1 533 instructions per pose with the issue slots 51 % busy and DRAM at 51 %: the P-MPJPE pass was instruction-bound.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <bool DO_M, bool DO_P>
__device__ __forceinline__ void pose_metrics16(float2* P, float2* G, float& dist, float& pdist) {
  constexpr int J = 16;
  dist = 0.f;
  pdist = 0.f;
  if (DO_M) {
    float da = 0.f, db = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 d0 = sub2(P[3 * i], G[3 * i]), d1 = sub2(P[3 * i + 1], G[3 * i + 1]), d2 = sub2(P[3 * i + 2], G[3 * i + 2]);
      const float2 q0 = mul2(d0, d0), q1 = mul2(d1, d1), q2 = mul2(d2, d2);
      da += fast_sqrt(q0.x + q0.y + q1.x);    // valid_bilinear.py:60
      db += fast_sqrt(q1.y + q2.x + q2.y);
    }
    dist = da + db;
  }
  if (DO_P) {
    float2 a0 = f2(0.f, 0.f), a1 = a0, a2 = a0, b0 = a0, b1 = a0, b2 = a0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a0 = add2(a0, P[3 * i]); a1 = add2(a1, P[3 * i + 1]); a2 = add2(a2, P[3 * i + 2]);
      b0 = add2(b0, G[3 * i]); b1 = add2(b1, G[3 * i + 1]); b2 = add2(b2, G[3 * i + 2]);
    }
    const float mpx = (a0.x + a1.y) * (1.f / J), mpy = (a0.y + a2.x) * (1.f / J), mpz = (a1.x + a2.y) * (1.f / J);
    const float mgx = (b0.x + b1.y) * (1.f / J), mgy = (b0.y + b2.x) * (1.f / J), mgz = (b1.x + b2.y) * (1.f / J);
    const float2 MP0 = f2(mpx, mpy), MP1 = f2(mpz, mpx), MP2 = f2(mpy, mpz);
    const float2 MG0 = f2(mgx, mgy), MG1 = f2(mgz, mgx), MG2 = f2(mgy, mgz);
    float2 np = f2(0.f, 0.f), ng = np;
    float2 C0 = np, C1 = np, C2 = np, R2 = np, D0 = np, D1 = np, D2 = np, E = np;
    float s22 = 0.f, s00 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 p0 = sub2(P[3 * i], MP0), p1 = sub2(P[3 * i + 1], MP1), p2 = sub2(P[3 * i + 2], MP2);
      const float2 g0 = sub2(G[3 * i], MG0), g1 = sub2(G[3 * i + 1], MG1), g2 = sub2(G[3 * i + 2], MG2);
      P[3 * i] = p0; P[3 * i + 1] = p1; P[3 * i + 2] = p2;
      G[3 * i] = g0; G[3 * i + 1] = g1; G[3 * i + 2] = g2;
      np = fma2(p0, p0, np); np = fma2(p1, p1, np); np = fma2(p2, p2, np);
      ng = fma2(g0, g0, ng); ng = fma2(g1, g1, ng); ng = fma2(g2, g2, ng);
      // even joint: p = (p0.x, p0.y, p1.x), g = (g0.x, g0.y, g1.x)
      C0 = fma2(p0, bc2(g0.x), C0);     // (S00, S10)
      C1 = fma2(p0, bc2(g0.y), C1);     // (S01, S11)
      C2 = fma2(p0, bc2(g1.x), C2);     // (S02, S12)
      R2 = fma2(bc2(p1.x), g0, R2);     // (S20, S21)
      s22 = fmaf(p1.x, g1.x, s22);
      // odd joint: p = (p1.y, p2.x, p2.y), g = (g1.y, g2.x, g2.y)
      D0 = fma2(p2, bc2(g1.y), D0);     // (S10, S20)
      D1 = fma2(p2, bc2(g2.x), D1);     // (S11, S21)
      D2 = fma2(p2, bc2(g2.y), D2);     // (S12, S22)
      E = fma2(bc2(p1.y), g2, E);       // (S01, S02)
      s00 = fmaf(p1.y, g1.y, s00);
    }
    const float nP2 = np.x + np.y, nG2 = ng.x + ng.y;
    float S[3][3];
    S[0][0] = C0.x + s00;  S[0][1] = C1.x + E.x;  S[0][2] = C2.x + E.y;
    S[1][0] = C0.y + D0.x; S[1][1] = C1.y + D1.x; S[1][2] = C2.y + D2.x;
    S[2][0] = R2.x + D0.y; S[2][1] = R2.y + D1.y; S[2][2] = s22 + D2.y;
    // zero-extent convention shared with oracle/pose_metrics.py (DEGENERATE_REL), as in pose_metrics
    const float rawP2 = nP2 + float(J) * (mpx * mpx + mpy * mpy + mpz * mpz);
    const float rawG2 = nG2 + float(J) * (mgx * mgx + mgy * mgy + mgz * mgz);
    if (nP2 > 1e-10f * rawP2 && nG2 > 1e-10f * rawG2) {
      const float inv = rsqrtf(nP2) * rsqrtf(nG2);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a][b] *= inv;
      float lam, q[4];
      horn_top(S, lam, q);
      const float scale = lam * sqrtf(nG2) * rsqrtf(nP2);   // tr * nG / nP (Appendix B)
      const float q0 = q[0], qx = q[1], qy = q[2], qz = q[3];
      const float s2 = 2.f * scale;
      const float R00 = scale * (q0 * q0 + qx * qx - qy * qy - qz * qz), R01 = s2 * (qx * qy - q0 * qz),
                  R02 = s2 * (qx * qz + q0 * qy);
      const float R10 = s2 * (qy * qx + q0 * qz), R11 = scale * (q0 * q0 - qx * qx + qy * qy - qz * qz),
                  R12 = s2 * (qy * qz - q0 * qx);
      const float R20 = s2 * (qz * qx - q0 * qy), R21 = s2 * (qz * qy + q0 * qx),
                  R22 = scale * (q0 * q0 - qx * qx - qy * qy + qz * qz);
      const float2 U0 = f2(R00, R10), U1 = f2(R01, R11), U2 = f2(R02, R12);   // rows 0, 1 of a column
      const float2 V0 = f2(R10, R20), V1 = f2(R11, R21), V2 = f2(R12, R22);   // rows 1, 2 of a column
      float pa = 0.f, pb = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 p0 = P[3 * i], p1 = P[3 * i + 1], p2 = P[3 * i + 2];
        const float2 g0 = G[3 * i], g1 = G[3 * i + 1], g2 = G[3 * i + 2];
        float2 t = fma2(U0, bc2(p0.x), neg2(g0));
        t = fma2(U1, bc2(p0.y), t);
        t = fma2(U2, bc2(p1.x), t);
        float az = fmaf(R20, p0.x, -g1.x);
        az = fmaf(R21, p0.y, az);
        az = fmaf(R22, p1.x, az);
        pa += fast_sqrt(fmaf(az, az, fmaf(t.y, t.y, t.x * t.x)));
        float2 u = fma2(V0, bc2(p1.y), neg2(g2));
        u = fma2(V1, bc2(p2.x), u);
        u = fma2(V2, bc2(p2.y), u);
        float ax = fmaf(R00, p1.y, -g1.y);
        ax = fmaf(R01, p2.x, ax);
        ax = fmaf(R02, p2.y, ax);
        pb += fast_sqrt(fmaf(ax, ax, fmaf(u.y, u.y, u.x * u.x)));
      }
      pdist = pa + pb;
    } else {
      // zero-extent pose: alignment undefined -> unaligned distance of the (uncentred) poses
      const float2 T0 = sub2(MP0, MG0), T1 = sub2(MP1, MG1), T2 = sub2(MP2, MG2);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 d0 = add2(sub2(P[3 * i], G[3 * i]), T0), d1 = add2(sub2(P[3 * i + 1], G[3 * i + 1]), T1),
                     d2 = add2(sub2(P[3 * i + 2], G[3 * i + 2]), T2);
        pdist += fast_sqrt(d0.x * d0.x + d0.y * d0.y + d1.x * d1.x);
        pdist += fast_sqrt(d1.y * d1.y + d2.x * d2.x + d2.y * d2.y);
      }
    }
  }
}

// Per-action accumulation (valid_bilinear.py:61-70) staged through shared memory.  Every pose adds its two fp32
// results and a count to the CTA's staging buckets with native shared-memory atomics (3 instructions per pose, a
// few-way conflict among the 32 lanes); once per chunk, right behind the barrier the chunk loop has anyway, thread a
// moves bucket a -- the fixed-point sum of at most a couple of hundred values -- into its float64 accumulator with an
// atomic exchange (read-and-clear: an add that arrives later simply stays for the next round).  The float64 totals
// leave the CTA once, at the end.  This replaced (i) a second kernel re-reading the per-pose results (11-18 us of
// kernel plus ~18 us of launch gap on a 62 us pass) and (ii) warp-shuffle reductions per bucket (+24-29 us).
// Buckets are FIXED-POINT (1/256 mm) unsigned 32-bit counters: a 32-bit integer add is a native shared-memory atomic, a float
// add is a compare-and-swap loop (ncu, r4x: 8 trips per warp with 15 actions, 7 % of the P-MPJPE kernel's samples).  A bucket
// holds at most two chunks (256 poses: the collecting thread may be overtaken by early adds of the next chunk) before it
// is cleared, so per-pose values below kFixMax cannot overflow it; anything larger (metres of error per joint) goes to
// the float buckets through the slow path.  Rounding: <= 1/512 mm per pose, unbiased -- three orders of magnitude below the
// 0.05 mm tolerance of the per-action means.
constexpr float kFixScale = 256.0f, kFixMax = 4.0e4f;   // 256 poses x 4e4 mm x 256 = 2.6e9 < 2^32
struct ActionStage {
  unsigned int s[kMaxActions], ps[kMaxActions];
  float sf[kMaxActions], psf[kMaxActions];
  unsigned int c[kMaxActions];
};
struct ActionTotals {
  double s = 0.0, ps = 0.0;
  unsigned long long c = 0ull;
};
__device__ __forceinline__ void stage_add1(unsigned int* fix, float* flt, int act, float v) {
  if (v < kFixMax) atomicAdd(&fix[act], __float2uint_rn(v * kFixScale));   // (negative / NaN never happen: sums of square roots)
  else atomicAdd(&flt[act], v);
}
template <bool DO_M, bool DO_P>
__device__ __forceinline__ void stage_add(ActionStage& st, int act, float dist, float pdist) {
  if (DO_M) stage_add1(st.s, st.sf, act, dist);
  if (DO_P) stage_add1(st.ps, st.psf, act, pdist);
  atomicAdd(&st.c[act], 1u);
}
template <bool DO_M, bool DO_P>
__device__ __forceinline__ void stage_collect(ActionStage& st, ActionTotals& tot, int tid, int n_actions) {
  if (tid < n_actions) {
    if (DO_M) tot.s += double(atomicExch(&st.s[tid], 0u)) * (1.0 / double(kFixScale)) + double(atomicExch(&st.sf[tid], 0.f));
    if (DO_P) tot.ps += double(atomicExch(&st.ps[tid], 0u)) * (1.0 / double(kFixScale)) + double(atomicExch(&st.psf[tid], 0.f));
    tot.c += atomicExch(&st.c[tid], 0u);
  }
}
template <bool DO_M, bool DO_P>
__device__ __forceinline__ void stage_flush(ActionStage& st, ActionTotals& tot, const EvalParams& prm, int tid) {
  __syncthreads();   // the last chunk's adds are in
  stage_collect<DO_M, DO_P>(st, tot, tid, prm.n_actions);
  if (tid < prm.n_actions && tot.c != 0ull) {
    if (DO_M && prm.acc_sum) atomicAdd(prm.acc_sum + tid, tot.s);
    if (DO_P && prm.pacc_sum) atomicAdd(prm.pacc_sum + tid, tot.ps);
    if (prm.acc_cnt) atomicAdd(reinterpret_cast<unsigned long long*>(prm.acc_cnt) + tid, tot.c);
  }
}

template <int J, bool DO_M, bool DO_P>
__global__ void __launch_bounds__(kEvalThreads) eval_kernel(const EvalParams prm) {
  constexpr int W = 3 * J;          // floats per pose
  constexpr int PITCH = W | 1;      // odd pitch: per-pose rows hit distinct banks
  constexpr int CH = kEvalThreads;  // poses per chunk
  extern __shared__ float sm[];
  float* sP = sm;
  float* sG = sm + CH * PITCH;
  float* sMean = sG + CH * PITCH;
  float* sStd = sMean + W;
  __shared__ ActionStage stage;
  ActionTotals tot;
  const int tid = threadIdx.x;
  for (int i = tid; i < kMaxActions; i += CH) { stage.s[i] = 0u; stage.ps[i] = 0u; stage.sf[i] = 0.f; stage.psf[i] = 0.f; stage.c[i] = 0u; }
  const bool unnorm = prm.mean != nullptr;
  for (int i = tid; i < W; i += CH) {
    sMean[i] = unnorm ? prm.mean[i] : 0.f;
    sStd[i] = unnorm ? prm.stddev[i] : 1.f;
  }
  __syncthreads();

  const int64_t nchunks = (prm.n + CH - 1) / CH;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t p0 = ch * CH;
    const int np = int(min(int64_t(CH), prm.n - p0));
    const int nfl = np * W;                       // floats in this chunk (per tensor)
    const float* gp = prm.pred + p0 * W;
    const float* gg = prm.gt + p0 * W;
    // chunk base is 16-byte aligned: p0 * W * 4 = ch * 128 * W * 4
    constexpr int V4 = CH * W / 4;                // float4 per tensor per full chunk
    constexpr int PER = (V4 + CH - 1) / CH;
    float4 rp[PER], rg[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int v = tid + k * CH;
      if (4 * v + 3 < nfl) {
        rp[k] = __ldcs(reinterpret_cast<const float4*>(gp) + v);
        rg[k] = __ldcs(reinterpret_cast<const float4*>(gg) + v);
      } else {
        float tp[4], tg[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = 4 * v + e;
          tp[e] = (idx < nfl) ? gp[idx] : 0.f;
          tg[e] = (idx < nfl) ? gg[idx] : 0.f;
        }
        rp[k] = make_float4(tp[0], tp[1], tp[2], tp[3]);
        rg[k] = make_float4(tg[0], tg[1], tg[2], tg[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int v = tid + k * CH;
      if (v < V4) {
        const float tp[4] = {rp[k].x, rp[k].y, rp[k].z, rp[k].w};
        const float tg[4] = {rg[k].x, rg[k].y, rg[k].z, rg[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = 4 * v + e;
          const int pose = idx / W, w = idx - pose * W;
          sP[pose * PITCH + w] = tp[e];
          sG[pose * PITCH + w] = tg[e];
        }
      }
    }
    __syncthreads();

    {
      const bool valid = tid < np;
      float dist = 0.f, pdist = 0.f;
      int act = 0;
      if (valid) {
        float P[W], G[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          P[w] = sStd[w] * sP[tid * PITCH + w] + sMean[w];   // valid_bilinear.py:53
          G[w] = sStd[w] * sG[tid * PITCH + w] + sMean[w];   // valid_bilinear.py:54
        }
        pose_metrics<J, DO_M, DO_P>(P, G, dist, pdist);
        const int64_t gi = p0 + tid;
        if (DO_M && prm.dist_out) prm.dist_out[gi] = dist;
        if (DO_P && prm.pdist_out) prm.pdist_out[gi] = pdist;
        act = prm.action ? prm.action[gi] : 0;
        act = min(max(act, 0), prm.n_actions - 1);
      }
      if (valid) stage_add<DO_M, DO_P>(stage, act, dist, pdist);
    }
    __syncthreads();
    stage_collect<DO_M, DO_P>(stage, tot, tid, prm.n_actions);
  }
  stage_flush<DO_M, DO_P>(stage, tot, prm, tid);
}

// ------------------------------------------------------------------------------------------------
// J = 16 fast path: TMA-fed.  128 poses of both tensors arrive as six 2-D boxes of {16 floats, 128
// poses} (64-byte rows, SWIZZLE_64B), so that thread r can pull ITS pose out of shared memory with
// conflict-free 128-bit loads (unit u of row r sits at u ^ ((r >> 1) & 3)).  A CTA owns ONE 48 KB
// buffer: as soon as every thread holds its pose in registers the buffer is handed back to the TMA
// unit for the next chunk, which lands while the per-pose math runs -- 3 CTAs (12 warps, 168 registers each) per SM hide
// the serial parts of the Procrustes solve, and no data passes through registers on its way into
// shared memory.  Per-action sums: staged through fixed-point shared-memory buckets (stage_add); when the caller
// keeps no per-pose output, the in-kernel warp walk.
// ------------------------------------------------------------------------------------------------
constexpr int kTmaPoses = kEvalThreads;             // poses per chunk
constexpr int kSubBytes = kTmaPoses * 64;           // one 16-float column group of one tensor
constexpr int kStageBytes = 6 * kSubBytes;          // 3 groups x {pred, gt} = 48 KB

#ifndef B200_EVAL_CTAS
#define B200_EVAL_CTAS 3   // resident CTAs per SM of the TMA-fed kernel: 3 = 168 registers, no spills (4 = 128 registers spilled 164 bytes per thread; measured 1M poses: P-MPJPE 91.2 -> 83.6 us, MPJPE 70.8 -> 67.2 us)
#endif
template <bool DO_M, bool DO_P, bool ACC>
__global__ void __launch_bounds__(kEvalThreads, B200_EVAL_CTAS)
eval_kernel_tma(const __grid_constant__ CUtensorMap tm_pred, const __grid_constant__ CUtensorMap tm_gt, const EvalParams prm) {
  constexpr int W = 48;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t full;
  __shared__ float4 sMean4[W / 4], sStd4[W / 4];
  float* const sMean = reinterpret_cast<float*>(sMean4);
  float* const sStd = reinterpret_cast<float*>(sStd4);
  __shared__ ActionStage stage;
  ActionTotals tot;
  const int tid = threadIdx.x;
  for (int i = tid; i < kMaxActions; i += kEvalThreads) {
    stage.s[i] = 0u; stage.ps[i] = 0u; stage.sf[i] = 0.f; stage.psf[i] = 0.f; stage.c[i] = 0u;
  }
  const bool unnorm = prm.mean != nullptr;
  for (int i = tid; i < W; i += kEvalThreads) {
    sMean[i] = unnorm ? prm.mean[i] : 0.f;
    sStd[i] = unnorm ? prm.stddev[i] : 1.f;
  }
  if (tid == 0) {
    mbar_init(&full, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int64_t nchunks = (prm.n + kTmaPoses - 1) / kTmaPoses;
  auto issue = [&](int64_t ch) {
    mbar_expect_tx(&full, kStageBytes);
    const int row = int(ch * kTmaPoses);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      tma_load_2d(smem + g * kSubBytes, &tm_pred, &full, 16 * g, row);
      tma_load_2d(smem + (3 + g) * kSubBytes, &tm_gt, &full, 16 * g, row);
    }
  };
  if (tid == 0 && int64_t(blockIdx.x) < nchunks) issue(blockIdx.x);
  uint32_t parity = 0;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, parity ^= 1u) {
    // this pose's action id: requested BEFORE the chunk is waited for -- read where it is used (after the per-pose math) the
    // load sat out a full DRAM round trip per chunk: 25 % of the P-MPJPE kernel's stall samples (ncu source view, r4x)
    int act = 0;
    if (ACC && prm.action && ch * kTmaPoses + tid < prm.n) act = __ldg(prm.action + ch * kTmaPoses + tid);
    mbar_wait(&full, parity, 900);
    float4 pr[12], gr[12];
    {
      const uint8_t* st = smem + tid * 64;
      const int sw = (tid >> 1) & 3;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pr[4 * g + u] = *reinterpret_cast<const float4*>(st + g * kSubBytes + ((u ^ sw) << 4));
          gr[4 * g + u] = *reinterpret_cast<const float4*>(st + (3 + g) * kSubBytes + ((u ^ sw) << 4));
        }
      }
    }
    // Un-normalise BEFORE the barrier: the arithmetic consumes every value the 24 shared-memory loads above deliver, so a
    // thread that arrives at the barrier has its reads of the buffer behind it -- and the TMA write of the next chunk, issued
    // right after the barrier, cannot overtake them.  That is what lets the re-issue go without fence.proxy.async (the reads
    // are done, and nothing was WRITTEN to the buffer through the generic proxy): the fence + MEMBAR.ALL.CTA in front of the
    // issue were 20 % of the issuing warp's samples (r4y), i.e. time the buffer sat idle in a pass that waits for HBM.
    // (MPJPE alone has almost no math behind the barrier: there the short path to the re-issue is loads -> barrier -> fence,
    // and the un-normalisation follows -- measured 66 us against 70 us with the order below)
    constexpr bool kEarly = DO_P;
    float2 P[24], G[24];
    auto unnormalise = [&]() {
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float4 sd = sStd4[k], mn = sMean4[k];
        const float2 sd0 = f2(sd.x, sd.y), sd1 = f2(sd.z, sd.w), mn0 = f2(mn.x, mn.y), mn1 = f2(mn.z, mn.w);
        P[2 * k] = fma2(sd0, f2(pr[k].x, pr[k].y), mn0);           // valid_bilinear.py:53
        P[2 * k + 1] = fma2(sd1, f2(pr[k].z, pr[k].w), mn1);
        G[2 * k] = fma2(sd0, f2(gr[k].x, gr[k].y), mn0);           // valid_bilinear.py:54
        G[2 * k + 1] = fma2(sd1, f2(gr[k].z, gr[k].w), mn1);
      }
    };
    if (kEarly) unnormalise();
    __syncthreads();   // every thread holds its pose: the buffer goes back to the TMA unit
    if (tid == 0 && ch + gridDim.x < nchunks) {
      if (!kEarly) fence_proxy_async_all();
      issue(ch + gridDim.x);
    }
    if (ACC) stage_collect<DO_M, DO_P>(stage, tot, tid, prm.n_actions);   // (the previous chunk's adds are all in)
    if (!kEarly) unnormalise();
    const int64_t p0 = ch * kTmaPoses;
    const int np = int(min(int64_t(kTmaPoses), prm.n - p0));
    const bool valid = tid < np;
    float dist = 0.f, pdist = 0.f;
    if (valid) {
      pose_metrics16<DO_M, DO_P>(P, G, dist, pdist);
      const int64_t gi = p0 + tid;
      if (DO_M && prm.dist_out) prm.dist_out[gi] = dist;
      if (DO_P && prm.pdist_out) prm.pdist_out[gi] = pdist;
      if (ACC) {
        act = min(max(act, 0), prm.n_actions - 1);
        stage_add<DO_M, DO_P>(stage, act, dist, pdist);
      }
    }
  }
  if (ACC) stage_flush<DO_M, DO_P>(stage, tot, prm, tid);
}

template <bool DO_M, bool DO_P, bool ACC>
static cudaError_t launch_eval_tma_t(const EvalParams& prm, const CUtensorMap& tp, const CUtensorMap& tg, cudaStream_t stream) {
  const size_t smem = size_t(kStageBytes) + 1024;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(eval_kernel_tma<DO_M, DO_P, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t nchunks = (prm.n + kTmaPoses - 1) / kTmaPoses;
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>(nchunks, int64_t(sms) * B200_EVAL_CTAS)));
  eval_kernel_tma<DO_M, DO_P, ACC><<<grid, kEvalThreads, smem, stream>>>(tp, tg, prm);
  return cudaGetLastError();
}

// tp / tg: tensor maps of pred / gt viewed as [n][48] fp32 with box {16, 128} and SWIZZLE_64B (encoded by
// the caller, api.cu); J = 16 only.
cudaError_t launch_eval_tma(const EvalParams& prm, const CUtensorMap* tp, const CUtensorMap* tg, bool do_m, bool do_p,
                            cudaStream_t stream, int* launches) {
  // one kernel: per-pose results (when the caller keeps them) and the per-action float64 sums in the same pass
  *launches = 1;
  if (do_m && do_p) return launch_eval_tma_t<true, true, true>(prm, *tp, *tg, stream);
  if (do_m) return launch_eval_tma_t<true, false, true>(prm, *tp, *tg, stream);
  return launch_eval_tma_t<false, true, true>(prm, *tp, *tg, stream);
}
int eval_tma_box_rows() { return kTmaPoses; }

template <int J, bool DO_M, bool DO_P>
static cudaError_t launch_eval_t(const EvalParams& prm, cudaStream_t stream) {
  constexpr int W = 3 * J, PITCH = W | 1;
  const size_t smem = size_t(2 * kEvalThreads * PITCH + 2 * W) * sizeof(float);
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(eval_kernel<J, DO_M, DO_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t nchunks = (prm.n + kEvalThreads - 1) / kEvalThreads;
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>(nchunks, int64_t(sms) * 4)));
  eval_kernel<J, DO_M, DO_P><<<grid, kEvalThreads, smem, stream>>>(prm);
  return cudaGetLastError();
}

cudaError_t launch_eval(const EvalParams& prm, int joints, bool do_m, bool do_p, cudaStream_t stream) {
  if (joints == 16) {
    if (do_m && do_p) return launch_eval_t<16, true, true>(prm, stream);
    if (do_m) return launch_eval_t<16, true, false>(prm, stream);
    return launch_eval_t<16, false, true>(prm, stream);
  }
  if (joints == 17) {
    if (do_m && do_p) return launch_eval_t<17, true, true>(prm, stream);
    if (do_m) return launch_eval_t<17, true, false>(prm, stream);
    return launch_eval_t<17, false, true>(prm, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200
