// oracle/pipeline_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.cpp header).
//
// The per-frame front-end of the reference as one CPU job, used for the CPU-baseline timing of
// bench.py: Frame::Frame (src/Frame.cc:61-131) = ExtractORB(left) + ExtractORB(right) +
// ComputeStereoMatches, for a batch of stereo pairs spread over `threads` host threads.  The
// reference runs the two extractors of one pair on two threads (src/Frame.cc:81-84); here every
// worker thread owns whole pairs, which keeps all cores busy with the same total work.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
int orb_oracle_extract(const uint8_t* img, int w, int h, int pitch, const uint8_t* mask, int mpitch, int nfeatures, float sf,
                       int nlevels, int iniTh, int minTh, void* kps, uint8_t* desc, int cap, uint8_t* pyr_out,
                       int32_t* cand_counts);
void orb_oracle_params(int nfeatures, float sf, int nlevels, int w, int h, int32_t* lw, int32_t* lh, int32_t* quota,
                       float* scale, int32_t* umax16);
void match_oracle_stereo(const void* kl, const uint8_t* dl, int nl, const void* kr, const uint8_t* dr, int nr,
                         const uint8_t* pyrL, const uint8_t* pyrR, const int64_t* off, const int32_t* lw, const int32_t* lh,
                         int nlevels, const float* scaleF, const float* invScaleF, float mb, float mbf, int stage,
                         float* uRight, float* depth, int32_t* ham_idx, int32_t* ham_dist);

// pairs: [n_pairs][2][h][w] u8.  counts: [n_pairs][2] key-points (left, right); matched: [n_pairs]
// stereo matches with depth > 0.  Returns the total number of key-points.
long pipeline_oracle_stereo_batch(const uint8_t* pairs, int n_pairs, int w, int h, int nfeatures, float sf, int nlevels,
                                  int iniTh, int minTh, float mb, float mbf, int threads, int32_t* counts, int32_t* matched) {
    std::vector<int32_t> lw(nlevels), lh(nlevels), quota(nlevels), umax(16);
    std::vector<float> scale(nlevels), inv(nlevels);
    orb_oracle_params(nfeatures, sf, nlevels, w, h, lw.data(), lh.data(), quota.data(), scale.data(), umax.data());
    std::vector<int64_t> off(nlevels);
    int64_t tot = 0;
    for (int l = 0; l < nlevels; ++l) { off[l] = tot; tot += (int64_t)lw[l] * lh[l]; inv[l] = 1.0f / scale[l]; }
    const int cap = nfeatures + 8 * nlevels + 64;
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            std::vector<uint8_t> kl((size_t)cap * 24), kr((size_t)cap * 24), dl((size_t)cap * 32), dr((size_t)cap * 32);
            std::vector<uint8_t> pl((size_t)tot), pr((size_t)tot);
            std::vector<float> ur(cap), dp(cap);
            std::vector<int32_t> hi(cap), hd(cap);
            for (int p = t; p < n_pairs; p += threads) {
                const uint8_t* L = pairs + (size_t)p * 2 * w * h;
                const uint8_t* R = L + (size_t)w * h;
                const int nl = orb_oracle_extract(L, w, h, w, nullptr, 0, nfeatures, sf, nlevels, iniTh, minTh, kl.data(), dl.data(), cap, pl.data(), nullptr);
                const int nr = orb_oracle_extract(R, w, h, w, nullptr, 0, nfeatures, sf, nlevels, iniTh, minTh, kr.data(), dr.data(), cap, pr.data(), nullptr);
                counts[2 * p] = nl; counts[2 * p + 1] = nr;
                int m = 0;
                if (nl > 0 && nr > 0) {
                    match_oracle_stereo(kl.data(), dl.data(), nl, kr.data(), dr.data(), nr, pl.data(), pr.data(), off.data(), lw.data(),
                                        lh.data(), nlevels, scale.data(), inv.data(), mb, mbf, 0, ur.data(), dp.data(), hi.data(), hd.data());
                    for (int i = 0; i < nl; ++i) m += dp[i] > 0;
                }
                matched[p] = m;
            }
        });
    for (auto& th : pool) th.join();
    long total = 0;
    for (int p = 0; p < 2 * n_pairs; ++p) total += counts[p] > 0 ? counts[p] : 0;
    return total;
}
}
