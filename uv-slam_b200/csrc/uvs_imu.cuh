// uvs_imu.cuh — IMU preintegration factor.
//
// IMUFactor::Evaluate (factor/imu_factor.h:19-182) + IntegrationBase::evaluate
// (factor/integration_base.h:160-186).  Two phases (see k_imu in uvs_sweep.cu): one THREAD per factor computes
// the frame geometry (a long dependent chain of quaternion algebra - doing it once per lane instead of once
// per warp cuts the issued instructions 3x), then one WARP per factor applies the sqrt_info (15x15
// upper-triangular) left-multiply so that consecutive lanes write consecutive doubles of the record.
#pragma once
#include "uvs_math.cuh"

namespace uvs {

// bottom-right 3x3 of Qleft(q) = w I + [u]x  /  Qright(q) = w I - [u]x   (utility/utility.h:46-64)
__device__ __forceinline__ m33 qleft_br(q4 q) {
  m33 r = skew(qvec(q));
  r.a[0] += q.w; r.a[4] += q.w; r.a[8] += q.w;
  return r;
}

__device__ __forceinline__ void put33(double *J, int ncols, int r0, int c0, const m33 &M, double sgn) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) J[(r0 + i) * ncols + c0 + j] = sgn * M.a[3 * i + j];
}

__device__ __forceinline__ m33 ld33(const double *__restrict__ M15, int r, int c) {
  m33 B;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) B.a[3 * i + j] = __ldg(M15 + (r + i) * 15 + c + j);
  return B;
}

struct ImuIn {
  const double *pose_i, *sb_i, *pose_j, *sb_j;   // state blocks
  const double *dp, *dq, *dv, *lin_ba, *lin_bg;  // preintegration constants
  double sum_dt;
  const double *jac;        // 15x15 row-major
  const double *sqrt_info;  // 15x15 upper
};

// Geometry of one factor by ONE thread: unweighted residual raw[15] and, when kJac, the twelve 3x3 blocks the
// unweighted Jacobian is made of (compact form, 9 doubles each, IMU_NBLK blocks; element e at comp[e * cstride]):
//   0 RiT   1 [a]x   2 -(Qleft(qj^-1 qi) Qright(dq^))_br   3 [b]x   4 RiT*T   5 dp_dba   6 dp_dbg
//   7 -Qleft(qj^-1 qi dq)_br dq_dbg   8 dv_dba   9 dv_dbg   10 Qleft(dq^^-1 qi^-1 qj)_br   11 I
constexpr int IMU_NBLK = 12;
constexpr int IMU_COMP = 9 * IMU_NBLK;

template <bool kJac>
__device__ __forceinline__ void imu_geometry(const ImuIn &in, const double g[3], double raw[15], double *comp, int cstride) {
  d3 Pi, Pj; q4 Qi, Qj;
  load_pose(in.pose_i, Pi, Qi);
  load_pose(in.pose_j, Pj, Qj);
  const d3 Vi = mk3(__ldg(in.sb_i), __ldg(in.sb_i + 1), __ldg(in.sb_i + 2));
  const d3 Bai = mk3(__ldg(in.sb_i + 3), __ldg(in.sb_i + 4), __ldg(in.sb_i + 5));
  const d3 Bgi = mk3(__ldg(in.sb_i + 6), __ldg(in.sb_i + 7), __ldg(in.sb_i + 8));
  const d3 Vj = mk3(__ldg(in.sb_j), __ldg(in.sb_j + 1), __ldg(in.sb_j + 2));
  const d3 Baj = mk3(__ldg(in.sb_j + 3), __ldg(in.sb_j + 4), __ldg(in.sb_j + 5));
  const d3 Bgj = mk3(__ldg(in.sb_j + 6), __ldg(in.sb_j + 7), __ldg(in.sb_j + 8));
  const d3 G = mk3(g[0], g[1], g[2]);
  const double T = in.sum_dt;
  const m33 dp_dba = ld33(in.jac, 0, 9), dp_dbg = ld33(in.jac, 0, 12), dq_dbg = ld33(in.jac, 3, 12);
  const m33 dv_dba = ld33(in.jac, 6, 9), dv_dbg = ld33(in.jac, 6, 12);
  const d3 dba = Bai - mk3(__ldg(in.lin_ba), __ldg(in.lin_ba + 1), __ldg(in.lin_ba + 2));
  const d3 dbg = Bgi - mk3(__ldg(in.lin_bg), __ldg(in.lin_bg + 1), __ldg(in.lin_bg + 2));
  const q4 delta_q = mkq(__ldg(in.dq), __ldg(in.dq + 1), __ldg(in.dq + 2), __ldg(in.dq + 3));
  const d3 th = mvec(dq_dbg, dbg);
  const q4 cdq = qmul(delta_q, mkq(th.x / 2.0, th.y / 2.0, th.z / 2.0, 1.0));   // corrected_delta_q (not unit)
  const d3 cdv = mk3(__ldg(in.dv), __ldg(in.dv + 1), __ldg(in.dv + 2)) + mvec(dv_dba, dba) + mvec(dv_dbg, dbg);
  const d3 cdp = mk3(__ldg(in.dp), __ldg(in.dp + 1), __ldg(in.dp + 2)) + mvec(dp_dba, dba) + mvec(dp_dbg, dbg);
  const q4 Qi_inv = qinv(Qi);
  const d3 a = qrot(Qi_inv, (0.5 * T * T) * G + Pj - Pi - T * Vi);
  const d3 b = qrot(Qi_inv, T * G + Vj - Vi);
  const q4 qij = qmul(Qi_inv, Qj);
  const q4 qlast = qmul(qinv(cdq), qij);
  const d3 rq = 2.0 * qvec(qlast);
  const d3 rp = a - cdp, rv = b - cdv, rba = Baj - Bai, rbg = Bgj - Bgi;
  raw[0] = rp.x; raw[1] = rp.y; raw[2] = rp.z; raw[3] = rq.x; raw[4] = rq.y; raw[5] = rq.z;
  raw[6] = rv.x; raw[7] = rv.y; raw[8] = rv.z; raw[9] = rba.x; raw[10] = rba.y; raw[11] = rba.z;
  raw[12] = rbg.x; raw[13] = rbg.y; raw[14] = rbg.z;
  if (!kJac) return;
  auto store = [&](int blk, const m33 &M, double sgn) {
#pragma unroll
    for (int k = 0; k < 9; k++) comp[(size_t)(9 * blk + k) * cstride] = sgn * M.a[k];
  };
  const m33 RiT = qmat(Qi_inv);
  const q4 qji = qmul(qinv(Qj), Qi);
  store(0, RiT, 1.0);
  store(1, skew(a), 1.0);
  {
    // -(Qleft(Qj^-1 Qi) Qright(corrected_delta_q)).bottomRightCorner<3,3>()        imu_factor.h:100-101
    const d3 u = qvec(qji), v = qvec(cdq);
    const m33 Lbr = qleft_br(qji);
    m33 Rbr = skew(v);
#pragma unroll
    for (int k = 0; k < 9; k++) Rbr.a[k] = -Rbr.a[k];
    Rbr.a[0] += cdq.w; Rbr.a[4] += cdq.w; Rbr.a[8] += cdq.w;
    m33 M = mmul(Lbr, Rbr);
    M.a[0] += u.x * -v.x; M.a[1] += u.x * -v.y; M.a[2] += u.x * -v.z;
    M.a[3] += u.y * -v.x; M.a[4] += u.y * -v.y; M.a[5] += u.y * -v.z;
    M.a[6] += u.z * -v.x; M.a[7] += u.z * -v.y; M.a[8] += u.z * -v.z;
    store(2, M, -1.0);
  }
  store(3, skew(b), 1.0);
  { m33 M = RiT; for (int k = 0; k < 9; k++) M.a[k] *= T; store(4, M, 1.0); }
  store(5, dp_dba, 1.0); store(6, dp_dbg, 1.0);
  // -Qleft(Qj^-1 Qi delta_q).bottomRightCorner<3,3>() dq_dbg   (delta_q, not corrected)   imu_factor.h:128
  store(7, mmul(qleft_br(qmul(qji, delta_q)), dq_dbg), -1.0);
  store(8, dv_dba, 1.0); store(9, dv_dbg, 1.0);
  store(10, qleft_br(qlast), 1.0);
  { m33 I; for (int k = 0; k < 9; k++) I.a[k] = (k % 4 == 0) ? 1.0 : 0.0; store(11, I, 1.0); }
}

// where the compact blocks go in the 15 x 30 unweighted Jacobian: {row0, col0, block, sign}
struct ImuPut { signed char r0, c0, blk, sgn; };
__constant__ ImuPut c_imu_puts[18] = {
    {0, 0, 0, -1}, {0, 3, 1, 1}, {3, 3, 2, 1}, {6, 3, 3, 1},                                   // pose_i
    {0, 6, 4, -1}, {0, 9, 5, -1}, {0, 12, 6, -1}, {3, 12, 7, 1}, {6, 6, 0, -1}, {6, 9, 8, -1},  // speed-bias_i
    {6, 12, 9, -1}, {9, 9, 11, -1}, {12, 12, 11, -1},
    {0, 15, 0, 1}, {3, 18, 10, 1},                                                             // pose_j
    {6, 21, 0, 1}, {9, 24, 11, 1}, {12, 27, 11, 1}};                                           // speed-bias_j

}  // namespace uvs
