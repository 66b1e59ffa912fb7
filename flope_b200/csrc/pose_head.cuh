// flope_b200: fused pose-head epilogue (sm_100a).
//
//   fc_rot            : Linear(2048, 9) + bias                 sunflower/models/posenet.py:19,33
//   procrustes        : (B,9) -> (B,3,3) row-major, projected onto SO(3)
//                       sunflower/utils/conversion.py:54-58 (roma.special_procrustes)
//   nullify yaw       : Euler 'zyx', zero the z angle, recompose  sunflower/utils/mvg.py:240-251
//
// One CTA per crop: the 2048-long dot products are a coalesced block reduction (shuffle tree + smem),
// then thread 0 does the 3x3 projection in fp64 (9 values per crop; costs nothing and keeps the
// near-degenerate random-init heads of SURVEY.md H3 stable).
#pragma once
#include "common.cuh"

namespace flope {

// Jacobi eigen-decomposition of a symmetric 3x3 (fp64): A = V diag(w) V^T, V orthonormal columns.
__device__ inline void jacobi_eig3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int pi = (pq == 2) ? 1 : 0;
      const int qi = (pq == 0) ? 1 : 2;
      const double apq = A[pi][qi];
      if (fabs(apq) < 1e-300) continue;
      const double theta = (A[qi][qi] - A[pi][pi]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) {           // A <- A J
        const double akp = A[k][pi], akq = A[k][qi];
        A[k][pi] = c * akp - s * akq;
        A[k][qi] = s * akp + c * akq;
      }
      for (int k = 0; k < 3; ++k) {           // A <- J^T A
        const double apk = A[pi][k], aqk = A[qi][k];
        A[pi][k] = c * apk - s * aqk;
        A[qi][k] = s * apk + c * aqk;
      }
      for (int k = 0; k < 3; ++k) {           // V <- V J
        const double vkp = V[k][pi], vkq = V[k][qi];
        V[k][pi] = c * vkp - s * vkq;
        V[k][qi] = s * vkp + c * vkq;
      }
    }
  }
  w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

// R = U diag(1,1,det(U V^T)) V^T for M = U S V^T.  With V from the eigenvectors of M^T M sorted by
// descending eigenvalue:  u1 = M v1/|.|, u2 = Gram-Schmidt(M v2), and the third term is
// det(V) * (u1 x u2) v3^T, which equals d*u3*v3^T without ever needing sigma3 or its sign.
__device__ inline void special_procrustes3(const double M[3][3], double R[3][3]) {
  double A[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = M[0][i] * M[0][j] + M[1][i] * M[1][j] + M[2][i] * M[2][j];
  jacobi_eig3(A, V, w);
  int o0 = 0, o1 = 1, o2 = 2;                                     // sort descending
  if (w[o0] < w[o1]) { int t = o0; o0 = o1; o1 = t; }
  if (w[o0] < w[o2]) { int t = o0; o0 = o2; o2 = t; }
  if (w[o1] < w[o2]) { int t = o1; o1 = o2; o2 = t; }
  double v1[3], v2[3], v3[3];
  for (int k = 0; k < 3; ++k) { v1[k] = V[k][o0]; v2[k] = V[k][o1]; v3[k] = V[k][o2]; }
  const double detV = v1[0] * (v2[1] * v3[2] - v2[2] * v3[1]) - v1[1] * (v2[0] * v3[2] - v2[2] * v3[0]) +
                      v1[2] * (v2[0] * v3[1] - v2[1] * v3[0]);
  double u1[3], u2[3], u3[3];
  for (int i = 0; i < 3; ++i) {
    u1[i] = M[i][0] * v1[0] + M[i][1] * v1[1] + M[i][2] * v1[2];
    u2[i] = M[i][0] * v2[0] + M[i][1] * v2[1] + M[i][2] * v2[2];
  }
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  if (n1 < 1e-150) { u1[0] = 1; u1[1] = 0; u1[2] = 0; n1 = 1; }    // M == 0: any rotation is optimal
  for (int i = 0; i < 3; ++i) u1[i] /= n1;
  const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  if (n2 < 1e-150 * (n1 + 1.0) || n2 < 1e-12 * n1) {               // rank 1: pick any unit vector orthogonal to u1
    const int k = (fabs(u1[0]) <= fabs(u1[1]) && fabs(u1[0]) <= fabs(u1[2])) ? 0 : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    const double de = u1[k];
    for (int i = 0; i < 3; ++i) u2[i] = e[i] - de * u1[i];
    n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  }
  for (int i = 0; i < 3; ++i) u2[i] /= n2;
  u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
  u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
  u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
  const double s3 = detV >= 0 ? 1.0 : -1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = u1[i] * v1[j] + u2[i] * v2[j] + s3 * u3[i] * v3[j];
}

// SciPy as_euler('zyx') -> zero the first (z) angle -> from_euler('zyx'):  R' = Rx(gamma) Ry(beta)
__device__ inline void nullify_yaw3(const double R[3][3], double Y[3][3]) {
  const double beta = atan2(R[0][2], hypot(R[0][0], R[0][1]));
  const double gamma = atan2(-R[1][2], R[2][2]);
  double sb, cb, sg, cg;
  sincos(beta, &sb, &cb);
  sincos(gamma, &sg, &cg);
  Y[0][0] = cb;       Y[0][1] = 0.0; Y[0][2] = sb;
  Y[1][0] = sg * sb;  Y[1][1] = cg;  Y[1][2] = -sg * cb;
  Y[2][0] = -cg * sb; Y[2][1] = sg;  Y[2][2] = cg * cb;
}

// One CTA (128 threads) per crop.  feat: (n, K) fp32 features (post-ReLU) or nullptr when r9_in / R_in is given.
// Outputs (each nullable): r9 (n,9) f32, R (n,9) f32 row-major, R_yaw (n,9) f64 row-major.
constexpr int kHeadThreads = 128;
__global__ void __launch_bounds__(kHeadThreads) pose_head_kernel(const float* __restrict__ feat, int K,
                                                                 const float* __restrict__ w_rot,
                                                                 const float* __restrict__ b_rot,
                                                                 const float* __restrict__ r9_in, int n,
                                                                 float* __restrict__ r9_out, float* __restrict__ R_out,
                                                                 double* __restrict__ Ryaw_out,
                                                                 const float* __restrict__ R_in) {
  __shared__ float s_part[kHeadThreads / 32][9];
  griddep_wait();
  griddep_launch();
  const int crop = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (feat) {
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float4* f4 = reinterpret_cast<const float4*>(feat + (size_t)crop * K);
    for (int k = threadIdx.x; k < (K >> 2); k += kHeadThreads) {
      const float4 f = f4[k];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_rot + (size_t)j * K) + k);
        acc[j] = fmaf(f.x, w.x, fmaf(f.y, w.y, fmaf(f.z, w.z, fmaf(f.w, w.w, acc[j]))));
      }
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      if (lane == 0) s_part[warp][j] = acc[j];
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  float r9[9];
  if (feat) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float t = b_rot[j];
#pragma unroll
      for (int w = 0; w < kHeadThreads / 32; ++w) t += s_part[w][j];
      r9[j] = t;
    }
  } else if (r9_in) {
#pragma unroll
    for (int j = 0; j < 9; ++j) r9[j] = r9_in[(size_t)crop * 9 + j];
  }
  if (!R_in && !R_out && !Ryaw_out) {           // fc_rot only (flope_posenet_forward): no projection wanted
    if (r9_out)
      for (int j = 0; j < 9; ++j) r9_out[(size_t)crop * 9 + j] = r9[j];
    return;
  }
  double R[3][3];
  if (R_in) {                                   // yaw-only mode (mvg.nullify_yaw_batch mirror)
    for (int j = 0; j < 9; ++j) R[j / 3][j % 3] = (double)R_in[(size_t)crop * 9 + j];
  } else {
    double M[3][3];
    for (int j = 0; j < 9; ++j) M[j / 3][j % 3] = (double)r9[j];
    special_procrustes3(M, R);
    if (r9_out)
      for (int j = 0; j < 9; ++j) r9_out[(size_t)crop * 9 + j] = r9[j];
    if (R_out)
      for (int j = 0; j < 9; ++j) R_out[(size_t)crop * 9 + j] = (float)R[j / 3][j % 3];
  }
  if (Ryaw_out) {
    // the reference hands the fp32 rotation to SciPy, so yaw nullification starts from the fp32-rounded R
    double Rf[3][3], Y[3][3];
    for (int j = 0; j < 9; ++j) Rf[j / 3][j % 3] = (double)(float)R[j / 3][j % 3];
    nullify_yaw3(Rf, Y);
    for (int j = 0; j < 9; ++j) Ryaw_out[(size_t)crop * 9 + j] = Y[j / 3][j % 3];
  }
}

}  // namespace flope
