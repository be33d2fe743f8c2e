"""Device-resident sliding window (include/uvs.h uvs_window_*) against the host-side restatement of the
reference's per-frame bookkeeping (tools/fm_ref.py): a seeded frame stream is fed frame by frame to
  (a) the resident window: uvs_window_push_frame / _upload / uvs_solve / _marginalize / _slide - observations, IMU records
      and the prior never leave the device again, and
  (b) a second handle that gets every window packed from scratch on the host (uvs_upload_windows),
and the two must hold the same problem (bit for bit) after every one of 22 consecutive slides of both kinds."""
import numpy as np
import pytest

import uvs_b200
from tools import fm_ref
from tools import gen_sequence as gs

pytestmark = pytest.mark.gpu

INT_ARRAYS = ("proj_frame_i", "proj_frame_j", "proj_point", "line_frame", "line_idx", "vp_frame", "vp_line", "imu_frame_i")
F64_ARRAYS = ("proj_pts_i", "proj_pts_j", "line_sp", "line_ep", "vp_dir", "imu_delta_p", "imu_delta_q", "imu_delta_v", "imu_sum_dt",
              "imu_lin_ba", "imu_lin_bg", "imu_jacobian", "imu_covariance", "prior_J", "prior_r")
COUNTS = ("n_frames", "n_points", "n_lines", "n_proj", "n_line_obs", "n_vp_obs", "n_imu", "prior_n")


def test_resident_window_follows_the_reference_bookkeeping():
    W = 10
    seq = gs.Sequence(n_frames=W + 1 + 22, n_points=150, n_lines=50, seed=3)
    opts = uvs_b200.default_options(max_num_iterations=4)
    dev, ref = uvs_b200.Solver(0), uvs_b200.Solver(0)
    dev.window_create(W, 5, 1024, 512)
    host = fm_ref.HostWindow(W, 5)
    nrng = np.random.default_rng(5)
    prior = None
    slides, h2d_frame, kinds_seen = 0, [], set()
    for k, fr in enumerate(seq.frames):
        h2d0 = dev.h2d_bytes()
        host.push(k, fr)
        dev.window_push_frame(fr.point_id, fr.point_xyz, fr.line_id, fr.line_sp, fr.line_ep, fr.line_vp, imu=fr.imu)
        if len(host.frame_ids) < W + 1:
            continue
        if slides == 7:   # FeatureManager::removeFailures / removeOutlier: a few tracks go, in both books
            drop_p = [t.id for t in host.eligible_points()[3:6]]
            drop_l = [t.id for t in host.eligible_lines()[1:2]]
            host.points = [t for t in host.points if t.id not in drop_p]
            host.lines = [t for t in host.lines if t.id not in drop_l]
            dev.window_remove_tracks(drop_p, drop_l)
        # ---- the state the estimator packs (vector2double): truth + noise for this window's frames / eligible landmarks
        pose, sb = seq.noisy_pose_sb(host.frame_ids, nrng)
        ep, el = host.eligible_points(), host.eligible_lines()
        inv = np.array([seq.inv_depth_of(t.id, host.frame_ids[t.start]) for t in ep]) * (1 + nrng.normal(0, 0.05, len(ep)))
        ortho = np.array([seq.ortho_of(t.id) for t in el]).reshape(-1, 4) + nrng.normal(0, 0.01, (len(el), 4))
        w_host = host.pack(pose, sb, seq.ex_pose(), inv, ortho, seq.ric, seq.tic, prior)
        counts = dev.window_counts()
        assert [counts[c] for c in COUNTS] == [getattr(w_host, c) for c in COUNTS], (k, counts)
        w_dev = uvs_b200.Window(pose=pose.copy(), speed_bias=sb.copy(), ex_pose=seq.ex_pose(), inv_depth=inv.copy(), ortho=ortho.copy(),
                                line_ric=seq.ric.copy(), line_tic=seq.tic.copy())
        dev.window_upload(w_dev, opts)
        h2d_frame.append(dev.h2d_bytes() - h2d0)
        # ---- what the device assembled == what the reference's assembly loops produce, bit for bit
        got = dev.download_factors(counts)
        for name in INT_ARRAYS + F64_ARRAYS:
            assert np.array_equal(getattr(got, name), getattr(w_host, name)), (k, slides, name)
        # ---- and it solves like the host-packed window (FP64 reductions are order-dependent: 1e-9)
        ref.upload([w_host], opts)
        sd, sr = dev.solve()[0], ref.solve()[0]
        dev.download(); ref.download()
        assert sd.num_iterations == sr.num_iterations
        assert abs(sd.final_cost - sr.final_cost) <= 1e-9 * abs(sr.final_cost), (k, sd.final_cost, sr.final_cost)
        # two solves of the SAME upload differ at this level too (order of the FP64 reductions); line parameters can be
        # weakly observable, so they get the looser bar
        assert np.abs(w_dev.pose - w_host.pose).max() < 1e-7 and np.abs(w_dev.speed_bias - w_host.speed_bias).max() < 1e-7, k
        assert np.abs(w_dev.inv_depth - w_host.inv_depth).max() < 1e-6 and np.abs(w_dev.ortho - w_host.ortho).max() < 1e-4, k
        # ---- next prior: built on the device from the resident window, kept there; the host path builds its own
        flag = fm_ref.MARGIN_SECOND_NEW if slides % 3 == 2 else fm_ref.MARGIN_OLD
        kinds_seen.add(flag)
        pd, pr = dev.window_marginalize(flag), ref.marginalize(0, flag)
        assert (pd is None) == (pr is None)
        if pd is not None:
            assert pd["n"] == pr["n"] and np.array_equal(pd["kinds"], pr["kinds"]) and np.array_equal(pd["ids"], pr["ids"])
            assert np.allclose(pd["x0"], pr["x0"], atol=1e-9)
            assert np.abs(pd["A"] - pr["A"]).max() <= 1e-6 * np.abs(pr["A"]).max()
            prior = pd
        elif flag == fm_ref.MARGIN_OLD:
            prior = None
        merged = merged_samples = None
        if flag == fm_ref.MARGIN_SECOND_NEW and len(host.imu) >= 2:
            merged_samples = gs.merge_samples(host.imu_samples[-2], host.imu_samples[-1])
            merged = gs.imu_record(merged_samples)
        host.slide(flag, merged, merged_samples)
        dev.window_slide(flag, merged)
        slides += 1
    assert slides >= 20 and kinds_seen == {0, 1}
    # per frame the resident window takes the new frame's observations, one IMU record, the packed state and a plan
    per_frame = int(np.median(h2d_frame))
    full = len(w_host.to_bytes())
    print("resident window: %d B host-to-device per frame (median of %d frames; a from-scratch upload of the same window: %d B)"
          % (per_frame, len(h2d_frame), full))
    assert per_frame < 20 * 1024 and per_frame < full / 6
    dev.close(); ref.close()


def test_resident_window_argument_checks():
    s = uvs_b200.Solver(0)
    with pytest.raises(uvs_b200.UvsError):
        s.window_counts()                                  # no resident window yet
    s.window_create(10, 5, 4, 4)
    z3, z2 = np.zeros((0, 3)), np.zeros((0, 2))
    s.window_push_frame([1, 2], [[0, 0, 1], [0.1, 0, 1]], [], z2, z2, z3)
    with pytest.raises(uvs_b200.UvsError):               # every later frame needs its IMU record
        s.window_push_frame([1, 2], [[0, 0, 1], [0.1, 0, 1]], [], z2, z2, z3)
    with pytest.raises(uvs_b200.UvsError):               # fewer than two frames
        s.window_upload(uvs_b200.Window(pose=np.zeros((1, 7)), speed_bias=np.zeros((1, 9)), ex_pose=np.zeros(7)))
    s.close()
