// Rotation / frame helpers (device functions) used by the regressor tail and the env-step kernels.
// torchgeometry==0.1.2 semantics; reference call sites
// motion/models/baseops.py:120-130 (cont2rotmat), :155-162 (rotmat2aa), :587-591 (update_transl_glorot).
#pragma once
#include "common.cuh"

namespace eg {

// F.normalize(v, eps=1e-12)
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
  const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
  x /= n; y /= n; z /= n;
}

// RotConverter.cont2rotmat: x[6] viewed as [3 rows, 2 cols]; R row-major with columns b1,b2,b3
__device__ __forceinline__ void cont6d_to_rotmat(const float* x, float* R) {
  float a0 = x[0], a1 = x[2], a2 = x[4];
  normalize3(a0, a1, a2);
  const float c0 = x[1], c1 = x[3], c2 = x[5];
  const float dot = a0 * c0 + a1 * c1 + a2 * c2;
  float b0 = c0 - dot * a0, b1 = c1 - dot * a1, b2 = c2 - dot * a2;
  normalize3(b0, b1, b2);
  const float d0 = a1 * b2 - a2 * b1, d1 = a2 * b0 - a0 * b2, d2 = a0 * b1 - a1 * b0;
  R[0] = a0; R[1] = b0; R[2] = d0;
  R[3] = a1; R[4] = b1; R[5] = d1;
  R[6] = a2; R[7] = b2; R[8] = d2;
}

// tgm.angle_axis_to_rotation_matrix (3x3 part)
__device__ __forceinline__ void tgm_aa_to_rotmat(const float* aa, float* R) {
  const float rx = aa[0], ry = aa[1], rz = aa[2];
  const float theta2 = rx * rx + ry * ry + rz * rz;
  if (theta2 > 1e-6f) {
    const float theta = sqrtf(theta2);
    const float inv = theta + 1e-6f;
    const float wx = rx / inv, wy = ry / inv, wz = rz / inv;
    const float c = cosf(theta), s = sinf(theta), omc = 1.0f - c;
    R[0] = c + wx * wx * omc;
    R[3] = wz * s + wx * wy * omc;
    R[6] = -wy * s + wx * wz * omc;
    R[1] = wx * wy * omc - wz * s;
    R[4] = c + wy * wy * omc;
    R[7] = wx * s + wy * wz * omc;
    R[2] = wy * s + wx * wz * omc;
    R[5] = -wx * s + wy * wz * omc;
    R[8] = c + wz * wz * omc;
  } else {
    R[0] = 1.f; R[1] = -rz; R[2] = ry;
    R[3] = rz; R[4] = 1.f; R[5] = -rx;
    R[6] = -ry; R[7] = rx; R[8] = 1.f;
  }
}

// tgm.rotation_matrix_to_angle_axis = rotation_matrix_to_quaternion (on the transpose, 4 cases,
// eps 1e-6) followed by quaternion_to_angle_axis
__device__ __forceinline__ void tgm_rotmat_to_aa(const float* R, float* aa) {
  // M = R^T : M[i][j] = R[j*3+i]
  const float m00 = R[0], m11 = R[4], m22 = R[8];
  const float m01 = R[3], m10 = R[1], m02 = R[6], m20 = R[2], m12 = R[7], m21 = R[5];
  float q0, q1, q2, q3, t;
  if (m22 < 1e-6f) {
    if (m00 > m11) {
      t = 1.f + m00 - m11 - m22;
      q0 = m12 - m21; q1 = t; q2 = m01 + m10; q3 = m20 + m02;
    } else {
      t = 1.f - m00 + m11 - m22;
      q0 = m20 - m02; q1 = m01 + m10; q2 = t; q3 = m12 + m21;
    }
  } else {
    if (m00 < -m11) {
      t = 1.f - m00 - m11 + m22;
      q0 = m01 - m10; q1 = m20 + m02; q2 = m12 + m21; q3 = t;
    } else {
      t = 1.f + m00 + m11 + m22;
      q0 = t; q1 = m12 - m21; q2 = m20 - m02; q3 = m01 - m10;
    }
  }
  const float sq = sqrtf(t);
  q0 = q0 / sq * 0.5f; q1 = q1 / sq * 0.5f; q2 = q2 / sq * 0.5f; q3 = q3 / sq * 0.5f;
  const float sin2 = q1 * q1 + q2 * q2 + q3 * q3;
  const float sn = sqrtf(sin2);
  const float two_theta = 2.0f * (q0 < 0.0f ? atan2f(-sn, -q0) : atan2f(sn, q0));
  const float k = sin2 > 0.0f ? two_theta / sn : 2.0f;
  aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
}

// CanonicalCoordinateExtractor.get_new_coordinate_torch (baseops.py:214-225):
// x = (j2 - j1) with z := 0, normalised (no epsilon); z = (0,0,1); y = normalise(z x x); R = [x y z] columns
__device__ __forceinline__ void new_coordinate(const float* j0, const float* j1, const float* j2, float* R, float* T) {
  float x0 = j2[0] - j1[0], x1 = j2[1] - j1[1];
  const float nx = sqrtf(x0 * x0 + x1 * x1 + 0.0f);
  x0 /= nx; x1 /= nx;
  // y = cross((0,0,1), (x0,x1,0)) = (-x1, x0, 0)
  float y0 = -x1, y1 = x0;
  const float ny = sqrtf(y0 * y0 + y1 * y1 + 0.0f);
  y0 /= ny; y1 /= ny;
  R[0] = x0; R[1] = y0; R[2] = 0.f;
  R[3] = x1; R[4] = y1; R[5] = 0.f;
  R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
  T[0] = j0[0]; T[1] = j0[1]; T[2] = j0[2];
}

}  // namespace eg
