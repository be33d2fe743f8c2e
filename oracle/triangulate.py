"""CPU restatement (numpy) of the reference's triangulation — TEST INFRASTRUCTURE ONLY: nothing under uv-slam_b200/
may import this file; tests/ use it as the checker of uvs_triangulate_points / uvs_triangulate_lines.

Follows, line by line:
  FeatureManager::triangulate      /root/reference/vins_estimator/src/feature_manager.cpp:427-481
  FeatureManager::triangulateLine  /root/reference/vins_estimator/src/feature_manager.cpp:504-589
  FeatureManager::calcPluckerLine  /root/reference/vins_estimator/src/feature_manager.cpp:827-902
  FeatureManager::setLineOrtho     /root/reference/vins_estimator/src/feature_manager.cpp:333-423 (validity test)
Eigen pieces restated from their published algorithms (Eigen is not in this image, version unpinned by the reference):
JacobiSVD(...).matrixV().rightCols<1>() = right singular vector of the smallest singular value (numpy.linalg.svd);
Matrix3d::eulerAngles(0, 1, 2) as implemented in Eigen 3.3 (Geometry/EulerAngles.h).  PARITY UNPINNED in the same sense
as the rest of oracle/: the reference ships no test vectors for these functions.
"""
import numpy as np


def triangulate_point(Rs, Ps, ric, tic, start_frame, pts, init_depth=5.0):
    """feature_manager.cpp:435-478 for one feature; pts[n][3] = feature_per_frame[k].point"""
    i = start_frame
    t0 = Ps[i] + Rs[i] @ tic
    R0 = Rs[i] @ ric
    A = np.zeros((2 * len(pts), 4))
    for k, pt in enumerate(pts):
        j = i + k
        t1 = Ps[j] + Rs[j] @ tic
        R1 = Rs[j] @ ric
        t = R0.T @ (t1 - t0)
        R = R0.T @ R1
        P = np.zeros((3, 4))
        P[:, :3] = R.T
        P[:, 3] = -R.T @ t
        f = pt / np.linalg.norm(pt)
        A[2 * k] = f[0] * P[2] - f[2] * P[0]
        A[2 * k + 1] = f[1] * P[2] - f[2] * P[1]
    V = np.linalg.svd(A, full_matrices=False)[2][-1]
    depth = V[2] / V[3]
    return init_depth if depth < 0.1 else depth


def euler_angles_012(m):
    """Eigen 3.3 MatrixBase::eulerAngles(0, 1, 2): m = Rx(a) Ry(b) Rz(c)"""
    i, j, k = 0, 1, 2     # odd = 0
    r0 = np.arctan2(m[j, k], m[k, k])
    c2 = np.hypot(m[i, i], m[i, j])
    if r0 > 0:
        r0 -= np.pi
        r1 = np.arctan2(-m[i, k], -c2)
    else:
        r1 = np.arctan2(-m[i, k], c2)
    s1, c1 = np.sin(r0), np.cos(r0)
    r2 = np.arctan2(s1 * m[k, i] - c1 * m[j, i], c1 * m[j, j] - s1 * m[k, j])
    return -np.array([r0, r1, r2])


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def calc_plucker_line(prev_sp, prev_ep, curr_sp, curr_ep, origin_prev, origin_curr):
    """feature_manager.cpp:827-902 -> (direction, normal)"""
    pn = skew(prev_sp) @ prev_ep
    prev_plane = np.array([pn[0], pn[1], pn[2], -(pn @ origin_prev)])
    cn = skew(curr_sp) @ curr_ep
    curr_plane = np.array([cn[0], cn[1], cn[2], -(cn @ origin_curr)])
    dual = np.outer(prev_plane, curr_plane) - np.outer(curr_plane, prev_plane)
    direction = np.array([dual[2, 1], dual[0, 2], dual[1, 0]])
    normal = np.array([dual[0, 3], dual[1, 3], dual[2, 3]])
    return direction, normal


def triangulate_line(Rs, Ps, ric, tic, frame_first, frame_last, sp_first, ep_first, sp_last, ep_last):
    """feature_manager.cpp:527-586 for one line -> orthonormal_vec[4]"""
    R_left = Rs[frame_first] @ ric
    t_left = Rs[frame_first] @ tic + Ps[frame_first]
    R_right = Rs[frame_last] @ ric
    t_right = Rs[frame_last] @ tic + Ps[frame_last]
    R_rel = R_left.T @ R_right                 # q_left.inverse() * q_right
    t_rel = R_left.T @ (t_right - t_left)
    direction, normal = calc_plucker_line(sp_first, ep_first, R_rel @ sp_last, R_rel @ ep_last, np.zeros(3), t_rel)
    n_w = R_left @ normal + skew(t_left) @ R_left @ direction
    d_w = R_left @ direction
    psi = np.zeros((3, 3))
    psi[:, 0] = n_w / np.linalg.norm(n_w)
    psi[:, 1] = d_w / np.linalg.norm(d_w)
    c = np.cross(n_w, d_w)
    psi[:, 2] = c / np.linalg.norm(c)
    out = np.zeros(4)
    out[:3] = euler_angles_012(psi)
    out[3] = np.arctan2(np.linalg.norm(d_w), np.linalg.norm(n_w))
    return out


def line_solve_flag(Rs, Ps, ric, tic, start_frame, ortho, sp, ep):
    """feature_manager.cpp:346-415 for one line -> (solve_flag, D_s_w, D_e_w); ortho = the feature's stored
    orthonormal_vec, sp / ep = start_point / end_point of its first observation (z = 1)"""
    a, b, c, phi = ortho
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    psi = Rx @ Ry @ Rz                                   # roll * pitch * yaw
    n_w = np.cos(phi) * psi[:, 0]
    d_w = np.sin(phi) * psi[:, 1]
    R_wc = Rs[start_frame] @ ric
    t_wc = Rs[start_frame] @ tic + Ps[start_frame]
    T_cw = np.zeros((6, 6))
    T_cw[:3, :3] = R_wc.T
    T_cw[:3, 3:] = skew(-R_wc.T @ t_wc) @ R_wc.T
    T_cw[3:, 3:] = R_wc.T
    l_c = T_cw @ np.concatenate([n_w, d_w])
    n_c, d_c = l_c[:3], l_c[3:]
    L_c = np.zeros((4, 4))
    L_c[:3, :3] = skew(n_c)
    L_c[:3, 3] = d_c
    L_c[3, :3] = -d_c
    with np.errstate(divide="ignore", invalid="ignore"):
        slope = -1.0 * (ep[0] - sp[0]) / (ep[1] - sp[1])
        sp_p = np.array([sp[0] + 1.0, slope + sp[1], 1.0])
        ep_p = np.array([ep[0] + 1.0, slope + ep[1], 1.0])
        pi_s = np.append(np.cross(sp, sp_p), 0.0)
        pi_e = np.append(np.cross(ep, ep_p), 0.0)
        D_s, D_e = L_c @ pi_s, L_c @ pi_e
        D_s3, D_e3 = D_s[:3] / D_s[3], D_e[:3] / D_e[3]
    flag = 2 if (D_s3[2] < 0 or D_e3[2] < 0) else 1
    return flag, R_wc @ D_s3 + t_wc, R_wc @ D_e3 + t_wc
