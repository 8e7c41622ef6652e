// Exercises the C++ host mirror (airdos_b200/host/airdos_host.hpp) end to end.
// usage: test_host <image.raw 640x480 u8> <out.bin>
// Writes: int32 n, n x adb_keypoint, n x 32 desc bytes, then BA: 3 doubles = pose 1 translation after LocalBundleAdjustment.
// Without a GPU it prints NO_DEVICE and exits 0 (the CPU test only checks that it compiles, links and loads).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../airdos_b200/host/airdos_host.hpp"

int main(int argc, char** argv) {
    if (adb_device_count() == 0) {
        try {
            ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 12, 7, 640, 480);
        } catch (const std::exception& e) {
            std::printf("NO_DEVICE (%s)\n", e.what());
            return 0;
        }
        return 1;   // must not succeed without a device
    }
    if (argc < 3) return 2;
    std::vector<uint8_t> img(640 * 480);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    std::fclose(f);
    // the reference's own five-argument constructor (include/ORBextractor.h:51-52, src/Tracking.cc:160-163): provisioned by the first frame
    ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 12, 7);
    if (ex.GetLevels() != 8 || ex.GetScaleFactors()[1] != 1.2f || ex.mnFeaturesPerLevel[0] != 217) return 7;   // getters work before the first frame
    std::vector<adb_keypoint> kps;
    std::vector<uint8_t> desc;
    airdos::ImageView im; im.data = img.data(); im.cols = 640; im.rows = 480; im.step = 640;
    ex(im, airdos::ImageView(), kps, desc);
    {   // the sized constructor gives the same answer
        ORB_SLAM2::ORBextractor sized(1000, 1.2f, 8, 12, 7, 640, 480);
        std::vector<adb_keypoint> ks; std::vector<uint8_t> ds;
        sized(im, airdos::ImageView(), ks, ds);
        if (ks.size() != kps.size() || ds != desc || std::memcmp(ks.data(), kps.data(), ks.size() * sizeof(adb_keypoint)) != 0) return 8;
    }
    // empty image: silent return
    std::vector<adb_keypoint> k2; std::vector<uint8_t> d2;
    ex(airdos::ImageView(), airdos::ImageView(), k2, d2);
    if (!k2.empty()) return 4;
    if (ORB_SLAM2::ORBmatcher::DescriptorDistance(desc.data(), desc.data()) != 0) return 5;
    if (ex.ImagePyramid()[1].size() != (size_t)533 * 400 || ex.GetLevels() != 8) return 6;

    // a 3-pose, 12-point toy window: points on a plane at z = 5, exact observations, perturbed pose 1
    const double fx = 500, cx = 320, cy = 240, bf = 100;
    std::vector<double> q = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1}, t = {0, 0, 0, -0.5, 0, 0, -1.0, 0, 0};
    std::vector<uint8_t> fixed = {1, 0, 0};
    std::vector<double> X, obs, info;
    std::vector<int32_t> ep, el;
    for (int i = 0; i < 12; ++i) {
        const double x = -1.5 + 0.4 * (i % 4) + 0.05 * i, y = -0.8 + 0.6 * (i / 4), z = 5.0 + 0.3 * (i % 3);
        X.insert(X.end(), {x, y, z});
        for (int k = 0; k < 3; ++k) {
            const double xc = x + t[3 * k], u = xc / z * fx + cx, v = y / z * fx + cy;
            ep.push_back(k); el.push_back(i);
            obs.insert(obs.end(), {u, v, u - bf / z});
            info.push_back(1.0);
        }
    }
    t[3] += 0.02; t[4] -= 0.01;   // perturb pose 1
    adb_ba_problem P;
    std::memset(&P, 0, sizeof(P));
    P.fx = fx; P.fy = fx; P.cx = cx; P.cy = cy; P.bf = bf;
    P.n_poses = 3; P.pose_q = q.data(); P.pose_t = t.data(); P.pose_fixed = fixed.data();
    P.n_points = 12; P.points = X.data();
    P.n_edges = (int)ep.size(); P.edge_pose = ep.data(); P.edge_point = el.data(); P.edge_obs = obs.data(); P.edge_info = info.data();
    adb_ba_result R;
    std::memset(&R, 0, sizeof(R));
    bool stop = false;
    if (!ORB_SLAM2::Optimizer::LocalBundleAdjustment(P, &stop, R)) return 7;
    if (std::fabs(t[3] + 0.5) > 1e-3 || std::fabs(t[4]) > 1e-3) { std::printf("BA did not recover pose 1: %g %g\n", t[3], t[4]); return 8; }
    stop = true;
    if (ORB_SLAM2::Optimizer::LocalBundleAdjustment(P, &stop, R)) return 9;   // stop flag set: early return

    // global BA entry: one round, robust off (LoopClosing's call), same toy window
    stop = false;
    t[3] += 0.02;
    if (!ORB_SLAM2::Optimizer::GlobalBundleAdjustemnt(P, 10, &stop, false, R)) return 10;
    if (R.iterations_run[1] != 0 || std::fabs(t[3] + 0.5) > 1e-3) return 11;

    // SearchByProjection(Current, Last): the extracted frame against itself (identity poses, every key-point a map point at
    // 5 m): each valid query must come back holding its own key-point
    {
        const int nk = (int)kps.size();
        std::vector<float> ur(nk, -1.f), xw(3 * (size_t)nk), ang(nk), sf(8);
        std::vector<int32_t> oct(nk), match(nk, -1);
        std::vector<uint8_t> flags(nk, 3);
        sf[0] = 1.f; for (int l = 1; l < 8; ++l) sf[l] = sf[l - 1] * 1.2f;
        for (int i = 0; i < nk; ++i) {
            xw[3 * i] = (kps[i].x - 320.f) * 5.f / 500.f; xw[3 * i + 1] = (kps[i].y - 240.f) * 5.f / 500.f; xw[3 * i + 2] = 5.f;
            oct[i] = kps[i].octave; ang[i] = kps[i].angle;
        }
        const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        adb_proj_search S;
        std::memset(&S, 0, sizeof(S));
        S.n_kp = nk; S.kps = kps.data(); S.u_right = ur.data(); S.desc = desc.data();
        S.min_x = 0; S.min_y = 0; S.max_x = 640; S.max_y = 480; S.grid_inv_w = 64.f / 640.f; S.grid_inv_h = 48.f / 480.f;
        S.n_q = nk; S.q_flags = flags.data(); S.q_desc = desc.data(); S.q_angle = ang.data();
        S.last_xw = xw.data(); S.last_octave = oct.data(); S.tcw_cur = I4; S.tcw_last = I4;
        S.fx = 500; S.fy = 500; S.cx = 320; S.cy = 240; S.mbf = 100; S.mb = 0.2f; S.scale_factors = sf.data(); S.n_levels = 8; S.th = 7; S.mono = 0;
        S.kp_match = match.data();
        ORB_SLAM2::ORBmatcher matcher(0.9f, true);
        const int nm = matcher.SearchByProjection(S);
        int self = 0;
        for (int i = 0; i < nk; ++i) self += match[i] == i;
        if (nm < nk * 9 / 10 || self < nk * 9 / 10) { std::printf("self search: %d matches, %d on themselves of %d\n", nm, self, nk); return 12; }
    }

    f = std::fopen(argv[2], "wb");
    const int32_t n = (int32_t)kps.size();
    std::fwrite(&n, 4, 1, f);
    std::fwrite(kps.data(), sizeof(adb_keypoint), n, f);
    std::fwrite(desc.data(), 32, n, f);
    std::fwrite(&t[3], 8, 3, f);
    std::fclose(f);
    std::printf("OK n=%d ba_trials=%d pose1=(%.6f %.6f %.6f)\n", n, R.trials_run, t[3], t[4], t[5]);
    return 0;
}
