/* airdos_b200.h -- C-ABI of libairdos_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for the three numeric classes on AirDOS's hot path.  The reference has no
 * FFI of its own; the seam is the C++ interface of ORB_SLAM2::ORBextractor / ORBmatcher /
 * Optimizer (namespace ORB_SLAM2, libORB_SLAM2.so).  Every entry point below names the
 * reference interface it replaces (paths relative to the reference tree).  INTEGRATION.md
 * shows the shim a maintainer drops into src/ORBextractor.cc, src/Frame.cc and
 * src/Optimizer.cc to forward those methods here.
 *
 * Conventions
 *   - plain C types only; every function returns an adb_status (0 = ok) and never aborts;
 *   - "host" entry points take host pointers and do the H2D / D2H copies themselves;
 *     "_device" entry points take device pointers (inputs already resident in HBM) and leave
 *     their results resident; results are fetched with the *_results / *_download calls;
 *   - handles are not thread-safe but independent: one handle per calling thread, each owns
 *     its CUDA stream and scratch (the reference runs one extractor per image thread,
 *     src/Frame.cc:81-84);
 *   - there is NO CPU fallback: without a CUDA device every create call returns
 *     ADB_ERR_NO_DEVICE.
 */
#ifndef AIRDOS_B200_H
#define AIRDOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADB_VERSION 100

typedef enum adb_status {
    ADB_OK = 0,
    ADB_ERR_INVALID = 1,     /* bad argument / unsupported shape */
    ADB_ERR_NO_DEVICE = 2,   /* no CUDA device, or not sm_100 */
    ADB_ERR_CUDA = 3,        /* a CUDA call failed; see adb_last_error() */
    ADB_ERR_CAPACITY = 4,    /* an output or scratch capacity was exceeded (nothing silently dropped) */
    ADB_ERR_NOT_POSDEF = 5,  /* BA: reduced system not positive definite in every LM trial */
    ADB_ERR_STOPPED = 6      /* BA: stop flag was set before optimisation started (no write-back) */
} adb_status;

const char* adb_last_error(void);   /* thread-local message of the last non-OK status */
int adb_version(void);
int adb_device_count(void);

/* ---------------------------------------------------------------------------------------
 * Key-point record: the cv::KeyPoint fields the SLAM code reads (pt, size, angle, response,
 * octave); 24 bytes, little endian.  Replaces std::vector<cv::KeyPoint>& _keypoints of
 * ORBextractor::operator() (include/ORBextractor.h:59-61). */
typedef struct adb_keypoint {
    float x, y;      /* pt, level-0 pixel coordinates */
    float size;      /* 31 * scale[octave], truncated to int as the reference does */
    float angle;     /* degrees [0, 360) */
    float response;  /* FAST score */
    int32_t octave;
} adb_keypoint;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (include/ORBextractor.h:51-52, src/ORBextractor.cc:411-472) plus the sizes the handle must
 * provision device memory for. */
typedef struct adb_orb_config {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;       /* 1..16 */
    int32_t ini_th_fast;
    int32_t min_th_fast;
    int32_t width, height; /* image size this handle is built for (fixed per handle) */
    int32_t max_batch;     /* frames processed per call (>= 1) */
    int32_t device;        /* CUDA device ordinal */
} adb_orb_config;

typedef struct adb_orb* adb_orb_t;

/* ctor / dtor of ORBextractor. */
adb_status adb_orb_create(const adb_orb_config* cfg, adb_orb_t* out);
adb_status adb_orb_destroy(adb_orb_t h);

/* Getters: GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:63-85) and the per-level quota
 * mnFeaturesPerLevel (src/ORBextractor.cc:433-445). */
int32_t adb_orb_levels(adb_orb_t h);
int32_t adb_orb_capacity(adb_orb_t h); /* max key-points per frame the extractor can emit */
adb_status adb_orb_level_info(adb_orb_t h, int32_t level, int32_t* w, int32_t* h_, int32_t* pitch,
                              float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                              int32_t* quota);

/* ORBextractor::operator()(image, mask, keypoints, descriptors)  (src/ORBextractor.cc:1054-1119)
 * for one frame, host buffers.  mask may be NULL (= all 255); otherwise CV_8UC1 semantics of
 * Frame::ExtractORB (src/Frame.cc:551-571): 0 = rejected.  kps / desc receive at most `cap`
 * records (desc is cap x 32 bytes); *n_out the count.  An empty image (w == 0 || h == 0)
 * returns ADB_OK with *n_out = 0, as the reference silently returns. */
adb_status adb_orb_extract(adb_orb_t h, const uint8_t* image, int32_t w, int32_t h_, int32_t pitch,
                           const uint8_t* mask, int32_t mask_pitch,
                           adb_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* n_out);

/* Same for n_frames frames stored frame_stride bytes apart (host memory; pinned memory makes
 * the copies asynchronous).  masks may be NULL.  kps: [n_frames][cap], desc: [n_frames][cap][32],
 * counts: [n_frames]. */
adb_status adb_orb_extract_batch(adb_orb_t h, int32_t n_frames, const uint8_t* images, size_t frame_stride,
                                 int32_t w, int32_t h_, int32_t pitch,
                                 const uint8_t* masks, size_t mask_frame_stride, int32_t mask_pitch,
                                 adb_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* counts);

/* Device-resident variant: images (and masks) are device pointers; results stay in HBM.
 * Asynchronous on the handle's stream; adb_orb_sync waits for it. */
adb_status adb_orb_extract_batch_device(adb_orb_t h, int32_t n_frames, const uint8_t* d_images, size_t frame_stride,
                                        int32_t w, int32_t h_, int32_t pitch,
                                        const uint8_t* d_masks, size_t mask_frame_stride, int32_t mask_pitch);
adb_status adb_orb_sync(adb_orb_t h);
void* adb_orb_stream(adb_orb_t h); /* cudaStream_t of the handle */

/* Measurement hooks.  adb_orb_profile(h, 1) makes every following extract call record CUDA events
 * on the handle's stream around its five stages; adb_orb_stage_ms then returns the device time
 * of the last call's stages {pyramid, FAST cells, quad-tree, level blur, orientation + descriptors} in ms.
 * adb_orb_launch_count = kernels this handle has launched since it was created. */
adb_status adb_orb_profile(adb_orb_t h, int32_t enable);
adb_status adb_orb_stage_ms(adb_orb_t h, float* ms5);
int64_t adb_orb_launch_count(adb_orb_t h);

/* Multi-GPU fusion of the descriptor all-gather (BASELINE configs[2]) into the extraction: besides its own result
 * buffers the handle scatters every record it produces straight into up to ADB_MAX_GATHER peer buffers -- device
 * pointers mapped over NVLink (CUDA IPC / symmetric memory), each laid out like the handle's own results
 * ([max_batch][capacity] key-points, [max_batch][capacity][32] descriptors, [max_batch] counts) and already offset to
 * this rank's slot.  The stores are issued by the descriptor kernel's epilogue, so the transfer overlaps the compute
 * tile by tile and no separate collective runs.  t == NULL or t->n == 0 switches it off.  The caller synchronises the
 * ranks (barrier) before reading the gathered buffers. */
#define ADB_MAX_GATHER 8
typedef struct adb_gather_targets {
    int32_t n;
    int32_t multicast;   /* 1: the (single) target is an NVLS multicast mapping of the symmetric buffers (NVSwitch replicates every
                            store to all ranks, own copy included): written with multimem.st, one copy leaves the GPU instead of n */
    adb_keypoint* kps[ADB_MAX_GATHER];
    uint8_t* desc[ADB_MAX_GATHER];
    int32_t* counts[ADB_MAX_GATHER];
} adb_gather_targets;
adb_status adb_orb_set_gather(adb_orb_t h, const adb_gather_targets* t);

/* Device pointers to the resident results of the last extract call:
 * kps [max_batch][capacity], desc [max_batch][capacity][32], counts [max_batch]. */
adb_status adb_orb_results_device(adb_orb_t h, const adb_keypoint** d_kps, const uint8_t** d_desc,
                                  const int32_t** d_counts, int32_t* capacity);
/* Copy the resident results of frames [first, first + n) to host buffers laid out as in
 * adb_orb_extract_batch.  Returns ADB_ERR_CAPACITY if a frame holds more than cap key-points. */
adb_status adb_orb_download(adb_orb_t h, int32_t first, int32_t n, adb_keypoint* kps, uint8_t* desc,
                            int32_t cap, int32_t* counts);

/* ORBextractor::mvImagePyramid[level] (public member read by Frame::ComputeStereoMatches,
 * src/Frame.cc:836,926-943; include/ORBextractor.h:86): copy level `level` of frame `frame` of the
 * last extract call to a host buffer (rows dst_pitch apart).  which = 0 image, 1 mask pyramid. */
adb_status adb_orb_get_pyramid(adb_orb_t h, int32_t frame, int32_t level, int32_t which, uint8_t* dst, int32_t dst_pitch);

/* Diagnostics for the parity tests: FAST candidates handed to the quad-tree for (frame, level),
 * in the reference's vToDistributeKeys order, as int32 triples (x, y, score) relative to the
 * (16,16) cell origin.  *n receives the count (<= cap written). */
adb_status adb_orb_debug_candidates(adb_orb_t h, int32_t frame, int32_t level, int32_t* xys, int32_t cap, int32_t* n);

/* ---------------------------------------------------------------------------------------
 * ORBmatcher::DescriptorDistance(a, b)  (src/ORBmatcher.cc:1647-1663): 256-bit Hamming
 * distance of two 32-byte descriptors; host inline helper, identical on every platform. */
int32_t adb_hamming_distance(const uint8_t* a, const uint8_t* b);

typedef struct adb_matcher* adb_matcher_t;
adb_status adb_matcher_create(int32_t device, adb_matcher_t* out);
adb_status adb_matcher_destroy(adb_matcher_t m);

/* The best / second-best scan every ORBmatcher::Search* / Fuse runs over its candidate list
 * (src/ORBmatcher.cc:85-114, 216-225 and the other nine call sites): for query q the candidates
 * are targets cand_idx[cand_off[q] .. cand_off[q+1]) in list order (cand_off == NULL: all nt
 * targets 0..nt-1).  Update rule: strict '<', i.e. the first candidate in list order wins
 * ties.  Outputs start at best = second = 256, idx = -1.  Host buffers.  The host variant checks the lists first (cand_off
 * monotone from 0, every cand_idx inside [0, nt)): ADB_ERR_INVALID; the device variant trusts lists that are already in HBM. */
adb_status adb_match_best2(adb_matcher_t m, const uint8_t* q_desc, int32_t nq, const uint8_t* t_desc, int32_t nt,
                           const int32_t* cand_off, const int32_t* cand_idx,
                           int32_t* best_idx, int32_t* best_d, int32_t* second_d);
/* Device-resident variant (all pointers in HBM; asynchronous on `stream`, a cudaStream_t or NULL). */
adb_status adb_match_best2_device(adb_matcher_t m, const uint8_t* q_desc, int32_t nq, const uint8_t* t_desc, int32_t nt,
                                  const int32_t* cand_off, const int32_t* cand_idx,
                                  int32_t* best_idx, int32_t* best_d, int32_t* second_d, void* stream);

/* MapPoint::ComputeDistinctiveDescriptors()  (src/MapPoint.cc:245-310), batched over map points: `desc` holds the
 * descriptors of all observations, point p owning rows point_ptr[p] .. point_ptr[p+1]).  Per point the observation whose
 * sorted Hamming distances to all observations of the point (itself included) have the least median (element
 * 0.5 * (N - 1)) wins, the first one on ties.  best_idx[p] = index within the point's rows (-1: no observations);
 * best_desc (optional) receives the winning 32 bytes (32 zero bytes for a point without observations: the reference returns
 * before touching mDescriptor, src/MapPoint.cc:259, so the caller keeps its old descriptor when best_idx[p] is -1).
 * At most ADB_MAX_OBSERVATIONS rows per point.  Host buffers. */
#define ADB_MAX_OBSERVATIONS 128
adb_status adb_distinctive_descriptors(adb_matcher_t m, const uint8_t* desc, const int32_t* point_ptr, int32_t n_points,
                                       int32_t* best_idx, uint8_t* best_desc);

/* ---------------------------------------------------------------------------------------
 * Guided window searches on the Frame grid, candidate generation included:
 *   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)          src/ORBmatcher.cc:45-129
 *   ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono)  src/ORBmatcher.cc:1328-1470
 * with Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:534-549, 700-712, 645-698) and
 * ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1601-1642) behind them.  Both searches are one loop over queries
 * (map points with a predicted pixel): the key-points of the 64 x 48 grid cells under the query's window, in the
 * reference's order (cell column, cell row, insertion order), minus the ones already held by an observed map point,
 * minus the ones whose right coordinate is off; best / second-best Hamming; accept; assign.  The loop is sequential in
 * the reference (an earlier query closes a key-point to later ones); the kernel iterates to the same fixed point.
 * One problem = one frame; all pointers are host memory. */
typedef struct adb_proj_search {
    /* the current Frame */
    int32_t n_kp;                  /* Frame::N (<= ADB_SEARCH_MAX) */
    const adb_keypoint* kps;       /* mvKeysUn: pt, octave, angle */
    const float* u_right;          /* mvuRight */
    const uint8_t* desc;           /* mDescriptors, n_kp x 32 */
    const uint8_t* taken;          /* NULL, or [n_kp]: mvpMapPoints[i] != NULL && Observations() > 0 on entry */
    float min_x, min_y, max_x, max_y;   /* mnMinX .. mnMaxY */
    float grid_inv_w, grid_inv_h;       /* mfGridElementWidthInv / mfGridElementHeightInv */
    /* queries, variant 1 (map points, src/ORBmatcher.cc:45-129): filled by the caller from MapPoint::mTrackProj* */
    int32_t n_q;                   /* <= ADB_SEARCH_MAX */
    const float* q_u;              /* mTrackProjX */
    const float* q_v;              /* mTrackProjY */
    const float* q_ur;             /* mTrackProjXR */
    const float* q_radius;         /* r * mvScaleFactors[nPredictedLevel] (r = RadiusByViewingCos(...) [* th]) */
    const int32_t* q_min_level;    /* nPredictedLevel - 1 */
    const int32_t* q_max_level;    /* nPredictedLevel */
    const uint8_t* q_flags;        /* bit 0: takes part (mbTrackInView && !isBad); bit 1: Observations() > 0 */
    const uint8_t* q_desc;         /* pMP->GetDescriptor(), n_q x 32 */
    const float* q_angle;          /* LastFrame.mvKeysUn[i].angle; only read when check_orientation != 0 */
    int32_t use_ratio;             /* 1: second-best + same-level ratio rule (variant 1) */
    float nn_ratio;                /* mfNNratio */
    int32_t check_orientation;     /* mbCheckOrientation (variant 2) */
    /* queries, variant 2 (last frame, src/ORBmatcher.cc:1328-1393): when last_xw != NULL the library projects the last
     * frame's map points itself and q_u .. q_max_level are ignored.  q_flags bit 0 then means "mvpMapPoints[i] != NULL
     * && !mvbOutlier[i]"; q_desc / q_angle / n_q are the last frame's. */
    const float* last_xw;          /* [n_q][3] pMP->GetWorldPos() */
    const int32_t* last_octave;    /* [n_q] LastFrame.mvKeys[i].octave */
    const float* tcw_cur;          /* row-major 4x4 CurrentFrame.mTcw */
    const float* tcw_last;         /* row-major 4x4 LastFrame.mTcw */
    float fx, fy, cx, cy, mbf, mb;
    const float* scale_factors;    /* CurrentFrame.mvScaleFactors */
    int32_t n_levels;
    float th;
    int32_t mono;                  /* bMono */
    /* queries, variant 1 with the visibility test on the device: when mp_xw != NULL the library runs Frame::isInFrustum
     * (src/Frame.cc:587-643) + MapPoint::PredictScale (src/MapPoint.cc:405-420) on every map point whose q_flags bit 0 is
     * set (= the point reached the test in Tracking::SearchLocalPoints, src/Tracking.cc:1319-1331) and builds the queries
     * from the result: q_u .. q_max_level are ignored, radius = RadiusByViewingCos(viewCos) [* th if th != 1] *
     * mvScaleFactors[level] (src/ORBmatcher.cc:55-69).  Uses tcw_cur, fx .. mbf, scale_factors, n_levels, th from above. */
    const float* mp_xw;            /* [n_q][3] GetWorldPos() */
    const float* mp_normal;        /* [n_q][3] GetNormal() */
    const float* mp_min_distance;  /* [n_q] mfMinDistance (the test uses 0.8f * it) */
    const float* mp_max_distance;  /* [n_q] mfMaxDistance (the test uses 1.2f * it, PredictScale the raw value) */
    const float* ow;               /* [3] Frame::mOw */
    float view_cos_limit;          /* 0.5 */
    float log_scale_factor;        /* Frame::mfLogScaleFactor */
    float* q_track;                /* optional out [n_q][4]: mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos (0 when not in view) */
    int32_t* q_level;              /* optional out [n_q]: mnTrackScaleLevel, -1 = mbTrackInView false */
    /* fuse != 0: the candidate search of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:825-975) on the
     * mp_* inputs: the "frame" arrays are the KeyFrame's (mvKeysUn, mvuRight, mDescriptors, grid), q_flags bit 0 =
     * pMP && !isBad() && !IsInKeyFrame(pKF).  Projection and tests as at :845-885 (x = Xc * invz first, IsInImage with a
     * half-open range, PO.dot(Pn) < 0.5 * dist3D), window radius th * mvScaleFactors[level], candidates of levels
     * [level - 1, level] that pass the chi2 gate on the reprojection error (:913-936, needs inv_level_sigma2), best
     * Hamming only, accepted when <= TH_LOW.  No key-point is closed to later queries: the Replace / AddObservation
     * bookkeeping of :948-968 stays on the host, driven by q_best_idx (>= 0 and q_best_dist <= 50 = "fuse this one").
     * kp_match / n_matches: the last accepted query per key-point / the number of accepted queries (= nFused). */
    int32_t fuse;
    const float* inv_level_sigma2; /* [n_levels] KeyFrame::mvInvLevelSigma2 */
    /* results */
    int32_t* kp_match;             /* [n_kp] CurrentFrame.mvpMapPoints after the call: -1 untouched, -2 set to NULL by the rotation
                                      check, >= 0 index of the query (map point) now held */
    int32_t* q_best_idx;           /* [n_q] optional: bestIdx of every query (-1: no candidate) */
    int32_t* q_best_dist;          /* [n_q] optional: bestDist (256: no candidate) */
    int32_t n_matches;             /* return value of the reference function */
} adb_proj_search;
#define ADB_SEARCH_MAX 8192
/* Runs n_problems independent searches (one thread block each).  Sizes, required pointers and every number that indexes a
 * per-level table on the device (last_octave; the key-points' octaves when fuse != 0) are checked on the host first:
 * ADB_ERR_INVALID, nothing launched. */
adb_status adb_search_by_projection(adb_matcher_t m, adb_proj_search* problems, int32_t n_problems);
/* Device time in ms of the kernels of the last adb_search_by_projection call. */
adb_status adb_search_last_ms(adb_matcher_t m, float* ms);

/* ---------------------------------------------------------------------------------------
 * Vocabulary-bucket searches: ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)  (src/ORBmatcher.cc:159-288)
 * and ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  (src/ORBmatcher.cc:657-823, with
 * CheckDistEpipolarLine, :131-157).  DBoW2 stays on the host (out of scope): the shim walks the two FeatureVectors as the
 * reference does (:183-245 / :692-775) and hands over, per common vocabulary node, the feature indices of both sides
 * in the reference's order (CSR).  Everything else -- Hamming scans, the first-come rule of SearchByBoW (a matched
 * key-point is closed to later queries), ratio test, epipole / epipolar-line gates, rotation histogram -- runs on the device.
 * One problem = one (key-frame, frame) or (key-frame, key-frame) pair; all pointers are host memory. */
typedef struct adb_bow_search {
    int32_t mode;                  /* 0 = SearchByBoW, 1 = SearchForTriangulation */
    /* side 1 (queries): pKF / pKF1 */
    int32_t n1;
    const adb_keypoint* kps1;      /* mvKeysUn */
    const float* u_right1;         /* mvuRight (mode 1) */
    const uint8_t* desc1;          /* mDescriptors */
    const uint8_t* flags1;         /* [n1] bit 0: takes part.  mode 0: map point present and not bad (:196-202);
                                      mode 1: no map point yet and (stereo || !bOnlyStereo) (:703-712) */
    /* side 2 (targets): F / pKF2 */
    int32_t n2;
    const adb_keypoint* kps2;      /* mode 0: F.mvKeys (angle, :228); mode 1: pKF2->mvKeysUn */
    const float* u_right2;
    const uint8_t* desc2;
    const uint8_t* flags2;         /* [n2] bit 0: may be taken.  mode 0: all 1; mode 1: no map point and (stereo || !bOnlyStereo) (:726-735) */
    /* common vocabulary nodes, ascending node id */
    int32_t n_buckets;
    const int32_t* b_ptr1;         /* [n_buckets + 1] into b_idx1 */
    const int32_t* b_idx1;         /* feature indices of side 1 per node, in FeatureVector order */
    const int32_t* b_ptr2;
    const int32_t* b_idx2;
    float nn_ratio;                /* mfNNratio (mode 0) */
    int32_t check_orientation;     /* mbCheckOrientation */
    /* mode 1 only */
    const float* f12;              /* row-major 3x3 */
    float ex, ey;                  /* epipole of camera 1 in image 2 (:663-670) */
    const float* scale_factors2;   /* pKF2->mvScaleFactors */
    const float* level_sigma2_2;   /* pKF2->mvLevelSigma2 */
    int32_t n_levels;
    /* results */
    int32_t* match21;              /* mode 0, [n2]: vpMapPointMatches as the index (side 1) of the key-point whose map point it holds, -1 none */
    int32_t* match12;              /* mode 1, [n1]: vMatches12 (index in side 2 or -1) */
    int32_t n_matches;
} adb_bow_search;
/* The bucket pointers (start at 0, monotone), the bucket indices (inside [0, n1) / [0, n2)) and, in mode 1, the octaves of side 2
 * (inside [0, n_levels)) are checked on the host first: ADB_ERR_INVALID, nothing launched. */
adb_status adb_search_by_bow(adb_matcher_t m, adb_bow_search* problems, int32_t n_problems);

/* Frame::ComputeStereoMatches()  (src/Frame.cc:829-1003) for the n_frames frames resident in
 * the two extractor handles (frame i of `left` against frame i of `right`): row-band Hamming
 * match, 11x11 SAD sub-pixel refinement on the pyramids, median-distance cut.
 * mb = baseline (mbf / fx), mbf = baseline * fx.  Outputs, host, [n_frames][cap] (cap >= the
 * left frame's key-point count, else ADB_ERR_CAPACITY): u_right = mvuRight, depth = mvDepth
 * (-1 where unmatched); best_idx / best_dist (optional, may be NULL) = right index and Hamming
 * distance of the row-band stage (-1 / 100 where no candidate beat TH_HIGH). */
adb_status adb_stereo_match(adb_orb_t left, adb_orb_t right, int32_t n_frames, float mb, float mbf,
                            float* u_right, float* depth, int32_t* best_idx, int32_t* best_dist, int32_t cap);
/* Device-resident variant: results stay in HBM ([max_batch][capacity] arrays owned by `left`). */
adb_status adb_stereo_match_device(adb_orb_t left, adb_orb_t right, int32_t n_frames, float mb, float mbf);
adb_status adb_stereo_results_device(adb_orb_t left, const float** d_u_right, const float** d_depth,
                                     const int32_t** d_best_idx, const int32_t** d_best_dist);

/* The hot part of the stereo Frame constructor (src/Frame.cc:80-100: ExtractORB(0, imLeft) and ExtractORB(1, imRight) on two
 * threads, then ComputeStereoMatches()) for n_frames stereo pairs in HOST memory, as one pipeline: per chunk of frames the two
 * uploads, the two extractions, the stereo matcher of the chunk and the downloads of everything the chunk produced overlap with
 * the neighbouring chunks.  Arguments as in adb_orb_extract_batch (left / right images share one geometry; masks for both images or
 * for neither) and adb_stereo_match; the results are identical to adb_orb_extract_batch(left) + adb_orb_extract_batch(right) +
 * adb_stereo_match, which is also what batches below 32 frames run. */
adb_status adb_stereo_frames_batch(adb_orb_t left, adb_orb_t right, int32_t n_frames, const uint8_t* images_left, const uint8_t* images_right,
                                   size_t frame_stride, int32_t w, int32_t h, int32_t pitch, const uint8_t* masks_left, const uint8_t* masks_right,
                                   size_t mask_frame_stride, int32_t mask_pitch, adb_keypoint* kps_left, uint8_t* desc_left, int32_t* counts_left,
                                   adb_keypoint* kps_right, uint8_t* desc_right, int32_t* counts_right, int32_t cap, float mb, float mbf,
                                   float* u_right, float* depth, int32_t* best_idx, int32_t* best_dist);


/* ---------------------------------------------------------------------------------------
 * Bundle adjustment: Optimizer::LocalBundleAdjustment (src/Optimizer.cc:431-731) and
 * Optimizer::LocalBundleAdjustmentHumanTrajactory (src/Optimizer.cc:1496-2222) as one flat
 * problem.  The g2o graph (vertices looked up by id, edges with virtual linearizeOplus) becomes
 * index arrays; what stays identical is the arithmetic: g2o::EdgeStereoSE3ProjectXYZ /
 * EdgeSE3ProjectXYZ residuals and Jacobians (Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:
 * 103-234), Huber IRLS (core/robust_kernel_impl.cpp:78-92, core/base_binary_edge.hpp:55-121),
 * Schur complement on the marginalised points (core/block_solver.hpp:354-486), Levenberg-
 * Marquardt control (core/optimization_algorithm_levenberg.cpp:61-189) and the two-round
 * schedule with chi2 gates of the two Optimizer functions.
 *
 * Poses are g2o::SE3Quat (unit quaternion x,y,z,w + translation, world -> camera), exactly
 * what Converter::toSE3Quat (src/Converter.cc:37-47) hands to g2o; adb_ba_pose_from_tcw /
 * adb_ba_pose_to_tcw do that float <-> double conversion for the shim.
 * All arrays are host memory; in/out arrays are updated in place on success. */
typedef struct adb_ba_problem {
    double fx, fy, cx, cy, bf;       /* KeyFrame::fx.. mbf (src/Optimizer.cc:571-575, 596-600) */
    /* g2o::VertexSE3Expmap per key-frame (local + fixed), src/Optimizer.cc:494-516 */
    int32_t n_poses;
    double* pose_q;                  /* [n_poses][4] in/out */
    double* pose_t;                  /* [n_poses][3] in/out */
    const uint8_t* pose_fixed;       /* [n_poses] */
    /* g2o::VertexSBAPointXYZ, marginalised, per static MapPoint, src/Optimizer.cc:541-548 */
    int32_t n_points;
    double* points;                  /* [n_points][3] in/out */
    /* EdgeStereoSE3ProjectXYZ (obs[2] >= 0) / EdgeSE3ProjectXYZ (obs[2] < 0), src/Optimizer.cc:550-618 */
    int32_t n_edges;
    const int32_t* edge_pose;        /* [n_edges] */
    const int32_t* edge_point;       /* [n_edges] */
    const double* edge_obs;          /* [n_edges][3] u, v, u_right */
    const double* edge_info;         /* [n_edges] invSigma2 (information = invSigma2 * I) */
    /* ---- articulated-human part (all counts may be 0), src/Optimizer.cc:1732-1957 ----
     * Residuals of the rigidity and motion edges are the reference's; their JACOBIANS deviate on purpose:
     * the reference's EdgeRigidBodyDouble::linearizeOplus reads never-assigned members (undefined behaviour) and
     * LandmarkMotionTernaryEdge::linearizeOplus rescales its Jacobian on every call; this library uses the analytic rigidity
     * Jacobian and the first-call motion Jacobian. */
    int32_t n_joints;                /* MapHumanKey vertices: VertexSBAPointXYZ, NOT marginalised */
    double* joints;                  /* [n_joints][3] in/out */
    int32_t n_joint_edges;           /* EdgeStereoSE3ProjectXYZ pose <-> joint, information SigmaHuman * I */
    const int32_t* jedge_pose;
    const int32_t* jedge_joint;
    const double* jedge_obs;         /* [n_joint_edges][3] */
    const double* jedge_info;        /* [n_joint_edges] */
    int32_t n_dists;                 /* VertexDistanceDouble (include/g2o_vertex_distance.h:28-45) */
    double* dists;                   /* [n_dists] in/out */
    int32_t n_rigid_edges;           /* EdgeRigidBodyDouble (include/g2o_edge_rigidbody.h:67-149) */
    const int32_t* redge_i;
    const int32_t* redge_j;
    const int32_t* redge_dist;
    const double* redge_info;        /* [n_rigid_edges] SigmaRigidity */
    int32_t n_motions;               /* VertexSE3 (include/g2o_vertex_se3.h:65-133): Isometry3 as quaternion + translation */
    double* motion_q;                /* [n_motions][4] in/out */
    double* motion_t;                /* [n_motions][3] in/out */
    int32_t n_motion_edges;          /* LandmarkMotionTernaryEdge (include/g2o_dyn_slam3d.h:11-101) */
    const int32_t* medge_p1;
    const int32_t* medge_p2;
    const int32_t* medge_motion;
    const double* medge_dt;          /* [n_motion_edges] delta_t */
    const double* medge_info;        /* [n_motion_edges] SigmaMotion */
} adb_ba_problem;

typedef struct adb_ba_options {
    int32_t iterations[2];           /* optimize(5), optimize(10): src/Optimizer.cc:625,667; round 2 skipped when 0 */
    int32_t max_trials;              /* maxTrialsAfterFailure = 10 */
    double tau;                      /* 1e-5: lambda_0 = tau * max diag(H) */
    double chi2_mono, chi2_stereo;   /* 5.991 / 7.815 gates (src/Optimizer.cc:640,653) */
    double chi2_rigid, chi2_motion;  /* thRanSacRigidity / thRanSacMotion gates (src/Optimizer.cc:2002,2011) */
    double huber_mono, huber_stereo; /* (float)sqrt(5.991), (float)sqrt(7.815) (src/Optimizer.cc:538-539) */
    double huber_rigid, huber_motion;/* thRanSacRigidity (not rooted, D.11), (float)sqrt(thRanSacMotion) */
    int32_t robust[2];               /* Huber kernel on in round 1 / round 2: {1, 0} for the two local-BA functions
                                        (src/Optimizer.cc:605-609, 660-665); Optimizer::BundleAdjustment(..., bRobust)
                                        (src/Optimizer.cc:52-230) is one round with robust[0] = bRobust */
} adb_ba_options;

#define ADB_BA_TRACE_COLS 5          /* lambda, chi2 before, chi2 after, rho, accepted (0/1) per LM trial */
typedef struct adb_ba_result {
    int32_t iterations_run[2];       /* outer LM iterations executed per round */
    int32_t trials_run;              /* LM trials (= reduced-system solves) over both rounds */
    int32_t stopped;                 /* 1 if the stop flag ended optimisation early */
    double chi2_initial;             /* robust chi2 before the first iteration */
    double chi2_round[2];            /* robust chi2 of the accepted state at the end of each round */
    double lambda_final;
    /* optional caller-allocated outputs (NULL to skip) */
    uint8_t* edge_outlier;           /* [n_edges]  1 = erase list entry (src/Optimizer.cc:671-699) */
    uint8_t* jedge_outlier;          /* [n_joint_edges] */
    uint8_t* redge_outlier;          /* [n_rigid_edges] */
    uint8_t* medge_outlier;          /* [n_motion_edges] */
    double* edge_chi2;               /* [n_edges] chi2 as the gates saw it */
    double* trace;                   /* [trace_cap][ADB_BA_TRACE_COLS] */
    int32_t trace_cap, trace_len;
} adb_ba_result;

void adb_ba_default_options(adb_ba_options* opt);   /* the constants of Optimizer::LocalBundleAdjustment */
/* The constants of Optimizer::BundleAdjustment / GlobalBundleAdjustemnt(pMap, nIterations, pbStopFlag, nLoopKF, bRobust)
 * (src/Optimizer.cc:52-230): one round of nIterations, Huber deltas sqrt(5.99) / sqrt(7.815), no chi2 gating between
 * rounds (outlier flags are still reported against 5.991 / 7.815 for the caller to ignore).  The problem is the same flat
 * adb_ba_problem: every key-frame a pose (id 0 fixed), every map point a marginalised point. */
void adb_ba_global_options(adb_ba_options* opt, int32_t n_iterations, int32_t robust);

/* Converter::toSE3Quat (src/Converter.cc:37-47) / Converter::toCvMat(SE3Quat) (src/Converter.cc:74-82):
 * row-major 4x4 float Tcw <-> quaternion (x,y,z,w) + translation in double. */
void adb_ba_pose_from_tcw(const float* tcw16, double* q4, double* t3);
void adb_ba_pose_to_tcw(const double* q4, const double* t3, float* tcw16);

typedef struct adb_ba* adb_ba_t;
adb_status adb_ba_create(int32_t device, adb_ba_t* out);
adb_status adb_ba_destroy(adb_ba_t s);
/* Runs the two-round schedule.  `stop` is Optimizer's bool* pbStopFlag (may be NULL), polled before
 * optimisation and between LM iterations / trials: set before the start -> ADB_ERR_STOPPED and
 * nothing is written (src/Optimizer.cc:620-622); set during round 1 -> round 2 is skipped but
 * gates and write-back still run (src/Optimizer.cc:627-631). */
adb_status adb_ba_solve(adb_ba_t s, adb_ba_problem* prob, const adb_ba_options* opt, volatile const uint8_t* stop,
                        adb_ba_result* res);
/* Device time in ms of the stages of the last adb_ba_solve: {linearise, schur, reduced solve,
 * back-substitution + evaluation, everything else, whole LM loop from the first linearisation to the
 * last trial (device events, includes the host's LM decisions between trials)}; and the kernel
 * launches this handle has made. */
adb_status adb_ba_stage_ms(adb_ba_t s, float* ms6);
int64_t adb_ba_launch_count(adb_ba_t s);

/* Pinning hook for the tests: evaluates the device implementations of the reference's LEAF arithmetic on arrays of n inputs --
 * Edge(Stereo)SE3ProjectXYZ[OnlyPose]::computeError / linearizeOplus (types_six_dof_expmap.cpp:103-364), VertexSE3Expmap::oplusImpl
 * (types_six_dof_expmap.h:73-76), EdgeRigidBodyDouble::computeError (include/g2o_edge_rigidbody.h:139-149),
 * LandmarkMotionTernaryEdge::computeError (include/g2o_dyn_slam3d.h:65-76), VertexSE3::oplusImpl (include/g2o_vertex_se3.h:113-122),
 * RobustKernelHuber::robustify (core/robust_kernel_impl.cpp:78-92) --
 * so that they can be compared with tests/golden/ba_leaf_ref.npz / lm_ref.npz (values produced by those reference sources themselves).
 * All arrays are host memory, doubles; out = n records of ADB_BA_LEAF_RECORD doubles (layout: airdos_b200/csrc/ba.cu, ba_leaf_kernel). */
#define ADB_BA_LEAF_RECORD 122
typedef struct adb_ba_leaf_io {
    int32_t n;
    double fx, fy, cx, cy, bf;
    const double* pose_q; const double* pose_t; const double* x; const double* obs; const double* pose_update;   /* [n][4|3|3|3|6] */
    const double* joint_a; const double* joint_b; const double* bone;                                             /* [n][3|3|1] */
    const double* motion_q; const double* motion_t; const double* motion_dt; const double* motion_update;         /* [n][4|3|1|6] */
    double* out;
    /* optional (both or neither, may be NULL): out[120], out[121] = rho(e2), rho'(e2) of RobustKernelHuber::robustify after setDelta(delta)
     * (Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-92; delta^2 is kept in a float there, core/robust_kernel_impl.h:84) */
    const double* huber_delta; const double* huber_e2;                                                            /* [n] */
} adb_ba_leaf_io;
adb_status adb_ba_leaf_eval(adb_ba_t s, const adb_ba_leaf_io* io);

/* g2o::LinearSolverDense<MatrixType>::solve (Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:64-113: dense Eigen LDLT of the
 * whole non-marginalised block, `!ldlt.isPositive()` -> false) and LinearSolverEigen::solve (linear_solver_eigen.h:92-115), as
 * adb_ba_solve uses them internally: solves A x = b for a symmetric positive definite A (row-major n x n, only the lower
 * triangle is read) by the in-cluster FP64 tensor-core Cholesky (chol.cu).  Host pointers.  *info = 0, or the 1-based index
 * of the 32-column panel start whose pivot was not positive (x is then meaningless, like g2o's rejected step).
 * cluster = CTAs in the thread-block cluster (8 or 16; 0 = chosen by order).  reps >= 1 repeats the device solve and
 * ms_per_solve (may be NULL) returns the mean device time of one solve (CUDA events around the kernel). */
adb_status adb_dense_solve(int32_t device, int32_t n, const double* a, const double* b, double* x, int32_t* info,
                           int32_t cluster, int32_t reps, float* ms_per_solve);

/* Optimizer::PoseOptimization(Frame*)  (src/Optimizer.cc:232-429): pose-only robust LM over the frame's
 * MapPoint correspondences (g2o::Edge(Stereo)SE3ProjectXYZOnlyPose), 4 rounds x 10 iterations with
 * chi2 re-classification between rounds, batched over frames (the reference calls it once per frame:
 * n_frames = 1).  Correspondence arrays are CSR over frames and float like the Frame / MapPoint members
 * they come from.  All pointers are host memory. */
typedef struct adb_pose_problem {
    double fx, fy, cx, cy, bf;       /* Frame::fx .. mbf */
    int32_t n_frames;
    const int32_t* frame_ptr;        /* [n_frames + 1] */
    double* pose_q;                  /* [n_frames][4] Converter::toSE3Quat(pFrame->mTcw), in/out */
    double* pose_t;                  /* [n_frames][3] in/out */
    const float* xw;                 /* [n][3] MapPoint::GetWorldPos() */
    const float* obs;                /* [n][3] mvKeysUn[i].pt.x, .pt.y, mvuRight[i] (< 0: monocular) */
    const float* inv_sigma2;         /* [n] mvInvLevelSigma2[octave] */
    uint8_t* outlier;                /* [n] out: mvbOutlier */
    int32_t* n_inliers;              /* [n_frames] out: return value (nInitialCorrespondences - nBad; 0 if < 3 correspondences) */
} adb_pose_problem;
adb_status adb_pose_optimize(adb_ba_t s, adb_pose_problem* prob);

#ifdef __cplusplus
}
#endif
#endif /* AIRDOS_B200_H */
