// uvs_imu.cuh — warp-cooperative IMU preintegration factor and the prior residual.
//
// IMUFactor::Evaluate (factor/imu_factor.h:19-182) + IntegrationBase::evaluate
// (factor/integration_base.h:160-186).  One warp per factor: every lane computes the (cheap) frame
// geometry redundantly, the 15x30 raw Jacobian is assembled in shared memory and the
// sqrt_info (15x15 upper-triangular) left-multiply is spread over the lanes so that consecutive
// lanes write consecutive doubles of the record.
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

// Computes the weighted residual (returned for row `lane` < 15, 0 otherwise) and, when kJac, fills
// Jraw (15x30 row-major, shared memory, this warp's slice) with the UNWEIGHTED Jacobian.
template <bool kJac>
__device__ __forceinline__ double imu_eval_warp(const ImuIn &in, const double g[3], int lane, double *Jraw) {
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
  const d3 rq = 2.0 * qvec(qmul(qinv(cdq), qij));
  double raw[15];
  {
    const d3 rp = a - cdp, rv = b - cdv, rba = Baj - Bai, rbg = Bgj - Bgi;
    raw[0] = rp.x; raw[1] = rp.y; raw[2] = rp.z;
    raw[3] = rq.x; raw[4] = rq.y; raw[5] = rq.z;
    raw[6] = rv.x; raw[7] = rv.y; raw[8] = rv.z;
    raw[9] = rba.x; raw[10] = rba.y; raw[11] = rba.z;
    raw[12] = rbg.x; raw[13] = rbg.y; raw[14] = rbg.z;
  }
  double res = 0.0;
  if (lane < 15) {
#pragma unroll
    for (int k = 0; k < 15; k++) res += __ldg(in.sqrt_info + lane * 15 + k) * raw[k];
  }
  if (kJac) {
    for (int e = lane; e < 450; e += 32) Jraw[e] = 0.0;
    __syncwarp();
    const m33 RiT = qmat(Qi_inv);
    const q4 qji = qmul(qinv(Qj), Qi);
    switch (lane) {
      // pose_i: columns 0-5
      case 0: put33(Jraw, 30, 0, 0, RiT, -1.0); break;
      case 1: put33(Jraw, 30, 0, 3, skew(a), 1.0); break;
      case 2: {
        // -(Qleft(Qj^-1 Qi) Qright(corrected_delta_q)).bottomRightCorner<3,3>()        imu_factor.h:100-101
        // rows 1..3 of the 4x4 product restricted to columns 1..3 (order w,x,y,z)
        const d3 u = qvec(qji), v = qvec(cdq);
        const m33 Lbr = qleft_br(qji);
        m33 Rbr = skew(v);
#pragma unroll
        for (int k = 0; k < 9; k++) Rbr.a[k] = -Rbr.a[k];
        Rbr.a[0] += cdq.w; Rbr.a[4] += cdq.w; Rbr.a[8] += cdq.w;
        m33 M = mmul(Lbr, Rbr);
        // + column 0 of L rows (= u) times row 0 of R columns 1..3 (= -v)
        M.a[0] += u.x * -v.x; M.a[1] += u.x * -v.y; M.a[2] += u.x * -v.z;
        M.a[3] += u.y * -v.x; M.a[4] += u.y * -v.y; M.a[5] += u.y * -v.z;
        M.a[6] += u.z * -v.x; M.a[7] += u.z * -v.y; M.a[8] += u.z * -v.z;
        put33(Jraw, 30, 3, 3, M, -1.0);
      } break;
      case 3: put33(Jraw, 30, 6, 3, skew(b), 1.0); break;
      // speed-bias_i: columns 6-14
      case 4: { m33 M = RiT; for (int k = 0; k < 9; k++) M.a[k] *= T; put33(Jraw, 30, 0, 6, M, -1.0); } break;
      case 5: put33(Jraw, 30, 0, 9, dp_dba, -1.0); break;
      case 6: put33(Jraw, 30, 0, 12, dp_dbg, -1.0); break;
      // -Qleft(Qj^-1 Qi delta_q).bottomRightCorner<3,3>() dq_dbg   (delta_q, not corrected)   imu_factor.h:128
      case 7: put33(Jraw, 30, 3, 12, mmul(qleft_br(qmul(qji, delta_q)), dq_dbg), -1.0); break;
      case 8: put33(Jraw, 30, 6, 6, RiT, -1.0); break;
      case 9: put33(Jraw, 30, 6, 9, dv_dba, -1.0); break;
      case 10: put33(Jraw, 30, 6, 12, dv_dbg, -1.0); break;
      case 11: Jraw[9 * 30 + 9] = Jraw[10 * 30 + 10] = Jraw[11 * 30 + 11] = -1.0; break;
      case 12: Jraw[12 * 30 + 12] = Jraw[13 * 30 + 13] = Jraw[14 * 30 + 14] = -1.0; break;
      // pose_j: columns 15-20
      case 13: put33(Jraw, 30, 0, 15, RiT, 1.0); break;
      case 14: put33(Jraw, 30, 3, 18, qleft_br(qmul(qinv(cdq), qij)), 1.0); break;
      // speed-bias_j: columns 21-29
      case 15: put33(Jraw, 30, 6, 21, RiT, 1.0); break;
      case 16: Jraw[9 * 30 + 24] = Jraw[10 * 30 + 25] = Jraw[11 * 30 + 26] = 1.0; break;
      case 17: Jraw[12 * 30 + 27] = Jraw[13 * 30 + 28] = Jraw[14 * 30 + 29] = 1.0; break;
      default: break;
    }
    __syncwarp();
  }
  return res;
}

}  // namespace uvs
