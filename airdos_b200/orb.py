"""Host-side mirror of ORB_SLAM2::ORBextractor / ORBmatcher over the C-ABI.

Names and argument meaning follow include/ORBextractor.h:51-86 and include/ORBmatcher.h:41-83
of the reference so that the parity tests read like calls into the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE, check, lib, ptr


class HostResults:
    """Reusable host buffers for extract_batch (pinned: the D2H copies run at full PCIe rate)."""

    def __init__(self, n_frames: int, cap: int, pinned: bool = False):
        self._keep = []
        def mk(shape, dtype):
            if pinned:
                import torch
                nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
                t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
                self._keep.append(t)
                return t.numpy()[:nbytes].view(dtype).reshape(shape)
            return np.zeros(shape, dtype)
        self.kps = mk((n_frames, cap), KP_DTYPE)
        self.desc = mk((n_frames, cap, 32), np.uint8)
        self.counts = mk((n_frames,), np.int32)
        self.cap = cap


class HostStereo:
    def __init__(self, n_frames: int, cap: int, pinned: bool = False):
        self._keep = []
        def mk(dtype):
            if pinned:
                import torch
                t = torch.empty(n_frames * cap * 4, dtype=torch.uint8).pin_memory()
                self._keep.append(t)
                return t.numpy().view(dtype).reshape(n_frames, cap)
            return np.zeros((n_frames, cap), dtype)
        self.u_right, self.depth = mk(np.float32), mk(np.float32)
        self.best_idx, self.best_dist = mk(np.int32), mk(np.int32)
        self.cap = cap


class ORBextractor:
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) -- the reference's own five arguments
    (include/ORBextractor.h:51-52): the device buffers are then provisioned by the first call from the size of its image.
    width / height given up front provision at construction (nothing is allocated inside a call); max_batch frames per call."""

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int,
                 width: int = 0, height: int = 0, max_batch: int = 1, device: int = 0):
        cfg = capi.OrbConfig(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height, max_batch, device)
        self._h = C.c_void_p()
        check(lib().adb_orb_create(C.byref(cfg), C.byref(self._h)))
        self.width, self.height, self.max_batch = width, height, max_batch
        self.nlevels = lib().adb_orb_levels(self._h)
        self.capacity = lib().adb_orb_capacity(self._h)
        self._info = [self._level_info(l) for l in range(self.nlevels)]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().adb_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _level_info(self, l):
        w, h, p, q = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        s, i, s2, i2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        check(lib().adb_orb_level_info(self._h, l, C.byref(w), C.byref(h), C.byref(p), C.byref(s), C.byref(i),
                                       C.byref(s2), C.byref(i2), C.byref(q)))
        return dict(w=w.value, h=h.value, pitch=p.value, scale=s.value, inv_scale=i.value, sigma2=s2.value,
                    inv_sigma2=i2.value, quota=q.value)

    # include/ORBextractor.h:63-85
    def GetLevels(self): return self.nlevels
    def GetScaleFactors(self): return [i["scale"] for i in self._info]
    def GetInverseScaleFactors(self): return [i["inv_scale"] for i in self._info]
    def GetScaleSigmaSquares(self): return [i["sigma2"] for i in self._info]
    def GetInverseScaleSigmaSquares(self): return [i["inv_sigma2"] for i in self._info]
    def level_sizes(self): return [(i["w"], i["h"]) for i in self._info]
    def quotas(self): return [i["quota"] for i in self._info]

    def __call__(self, image: np.ndarray, mask: np.ndarray | None = None):
        """operator()(image, mask, keypoints, descriptors) for one CV_8UC1 image -> (kps, desc)."""
        image = np.ascontiguousarray(image, np.uint8)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        k, d, c = self.extract_batch(image[None], None if mask is None else np.ascontiguousarray(mask, np.uint8)[None])
        return k[0, :c[0]].copy(), d[0, :c[0]].copy()

    def extract_batch(self, images: np.ndarray, masks: np.ndarray | None = None, out: HostResults | None = None):
        """images u8 [F, H, W] (host) -> kps [F, cap], desc [F, cap, 32], counts [F]."""
        images = np.ascontiguousarray(images, np.uint8)
        f, h, w = images.shape
        if out is None:
            out = HostResults(f, self.capacity)
        kps, desc, counts = out.kps[:f], out.desc[:f], out.counts[:f]
        if masks is not None:
            masks = np.ascontiguousarray(masks, np.uint8)
            assert masks.shape == images.shape
        check(lib().adb_orb_extract_batch(self._h, f, ptr(images), h * w, w, h, w, ptr(masks), h * w, w,
                                          ptr(kps), ptr(desc), self.capacity, ptr(counts)))
        self._after_call(w, h)
        return kps, desc, counts

    def _after_call(self, w: int, h: int):
        """A lazily provisioned handle learns its level sizes with the first image (and again when the size changes)."""
        if (self.width, self.height) != (w, h):
            self.width, self.height = w, h
            self._info = [self._level_info(l) for l in range(self.nlevels)]
            self.capacity = lib().adb_orb_capacity(self._h)      # the creation-time upper bound becomes the provisioned row count

    def extract_batch_device(self, d_images: int, n_frames: int, frame_stride: int | None = None, pitch: int | None = None,
                             d_masks: int | None = None, width: int | None = None, height: int | None = None):
        """Device-resident inputs (raw device pointer); asynchronous, results stay in HBM."""
        if width is not None and height is not None:
            w, h = width, height
        else:
            w, h = self.width, self.height
            if not w or not h:
                raise ValueError("a lazily provisioned extractor needs width / height with its first device-resident call")
        pitch = pitch or w
        frame_stride = frame_stride or pitch * h
        check(lib().adb_orb_extract_batch_device(self._h, n_frames, ptr(d_images), frame_stride, w, h,
                                                 pitch, ptr(d_masks), frame_stride, pitch))
        self._after_call(w, h)

    def sync(self):
        check(lib().adb_orb_sync(self._h))

    def stream(self) -> int:
        return int(lib().adb_orb_stream(self._h) or 0)

    def set_gather(self, kps_ptrs=(), desc_ptrs=(), counts_ptrs=(), multicast: bool = False):
        """Peer-mapped device pointers (one triple per rank, already offset to this rank's slot) that the descriptor
        kernel writes every record to as well; empty = off.  multicast: the single triple is an NVLS multicast mapping."""
        t = capi.GatherTargets()
        t.n = len(kps_ptrs)
        t.multicast = int(bool(multicast))
        for g, (a, b, c) in enumerate(zip(kps_ptrs, desc_ptrs, counts_ptrs)):
            t.kps[g], t.desc[g], t.counts[g] = int(a), int(b), int(c)
        check(lib().adb_orb_set_gather(self._h, C.byref(t)))

    def profile(self, enable: bool = True):
        check(lib().adb_orb_profile(self._h, int(enable)))

    def stage_ms(self):
        """Device ms of the last profiled call: (pyramid, fast_cells, quadtree, blur7_level, orient_describe)."""
        ms = (C.c_float * 5)()
        check(lib().adb_orb_stage_ms(self._h, ms))
        return tuple(ms)

    def launch_count(self) -> int:
        return int(lib().adb_orb_launch_count(self._h))

    def results_device(self):
        k, d, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        cap = C.c_int32()
        check(lib().adb_orb_results_device(self._h, C.byref(k), C.byref(d), C.byref(c), C.byref(cap)))
        return k.value, d.value, c.value, cap.value

    def download(self, first: int, n: int):
        kps = np.zeros((n, self.capacity), KP_DTYPE)
        desc = np.zeros((n, self.capacity, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        check(lib().adb_orb_download(self._h, first, n, ptr(kps), ptr(desc), self.capacity, ptr(counts)))
        return kps, desc, counts

    def pyramid(self, frame: int = 0, which: int = 0):
        """mvImagePyramid (which=0) / mvMaskPyramid (which=1) of `frame` of the last call."""
        out = []
        for l, (w, h) in enumerate(self.level_sizes()):
            a = np.zeros((h, w), np.uint8)
            check(lib().adb_orb_get_pyramid(self._h, frame, l, which, ptr(a), w))
            out.append(a)
        return out

    def debug_candidates(self, frame: int, level: int):
        n = C.c_int32()
        check(lib().adb_orb_debug_candidates(self._h, frame, level, None, 0, C.byref(n)))
        a = np.zeros((max(n.value, 1), 3), np.int32)
        check(lib().adb_orb_debug_candidates(self._h, frame, level, ptr(a), n.value, C.byref(n)))
        return a[:n.value]


class ORBmatcher:
    """The Hamming primitives of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:41-83)."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # src/ORBmatcher.cc:37-39

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, device: int = 0):
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self._m = C.c_void_p()
        check(lib().adb_matcher_create(device, C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None) is not None and self._m:
            lib().adb_matcher_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray) -> int:
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        assert a.size == 32 and b.size == 32
        return int(lib().adb_hamming_distance(ptr(a), ptr(b)))

    def best2(self, q_desc: np.ndarray, t_desc: np.ndarray, cand_off: np.ndarray | None = None,
              cand_idx: np.ndarray | None = None):
        """best / second-best Hamming scan over candidate lists -> (best_idx, best_d, second_d)."""
        q = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
        nq, nt = len(q), len(t)
        bi = np.full(nq, -1, np.int32); bd = np.full(nq, 256, np.int32); sd = np.full(nq, 256, np.int32)
        if cand_off is not None:
            cand_off = np.ascontiguousarray(cand_off, np.int32); cand_idx = np.ascontiguousarray(cand_idx, np.int32)
            assert len(cand_off) == nq + 1
        if nq:
            check(lib().adb_match_best2(self._m, ptr(q), nq, ptr(t), nt, ptr(cand_off), ptr(cand_idx), ptr(bi), ptr(bd), ptr(sd)))
        return bi, bd, sd


    # -- guided window searches (candidate generation on the device) ---------------------------------
    def search_by_projection(self, problems: list[dict]):
        """Batch of independent searches.  Each problem is a dict with the current frame (`kps`, `u_right`, `desc`,
        optional `taken`, `bounds` = (mnMinX, mnMinY, mnMaxX, mnMaxY)) and either projected queries (`q_u`, `q_v`, `q_ur`,
        `q_radius`, `q_min_level`, `q_max_level`, `q_flags`, `q_desc`; SearchByProjection(F, vpMapPoints, th),
        src/ORBmatcher.cc:45-129) or the last frame (`last_xw`, `last_octave`, `q_flags`, `q_desc`, `q_angle`, `tcw_cur`,
        `tcw_last`, `cam` = (fx, fy, cx, cy, mbf, mb), `scale_factors`, `th`, `mono`; SearchByProjection(Current, Last, th,
        bMono), src/ORBmatcher.cc:1328-1470) or the local map points before the visibility test (`mp_xw`, `mp_normal`,
        `mp_min_distance`, `mp_max_distance`, `ow`, `tcw_cur`, `cam`, `scale_factors`, `log_scale_factor`, `th`; Frame::isInFrustum
        + SearchByProjection as Tracking::SearchLocalPoints runs them, src/Tracking.cc:1319-1343; two more outputs: q_track
        [n_q, 4] = mTrackProjX / Y / XR / ViewCos and q_level = mnTrackScaleLevel or -1).  With `fuse` = 1 (+ `inv_level_sigma2`)
        the same inputs run the candidate search of ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:825-975).
        Returns a list of (nmatches, kp_match, q_best_idx, q_best_dist[, q_track, q_level])."""
        prep = self.prepare_projection(problems)
        return self.run_prepared_projection(prep)

    def prepare_projection(self, problems: list[dict]):
        """Fills the adb_proj_search array of a batch once (what the C++ shim does with plain pointer assignments); the batch can then
        be run any number of times with run_prepared_projection -- the C-ABI call proper, host buffers in and out."""
        from .capi import ProjSearch
        n = len(problems)
        arr = (ProjSearch * n)()
        keep = []   # numpy arrays referenced by the structs

        def a(x, dt):
            x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data

        outs = []
        for i, pr in enumerate(problems):
            S = arr[i]
            kps = np.ascontiguousarray(pr["kps"], KP_DTYPE); keep.append(kps)
            S.n_kp = len(kps); S.kps = kps.ctypes.data
            S.u_right = a(pr["u_right"], np.float32); S.desc = a(pr["desc"], np.uint8)
            S.taken = a(pr["taken"], np.uint8) if pr.get("taken") is not None else None
            mnx, mny, mxx, mxy = [np.float32(v) for v in pr["bounds"]]
            S.min_x, S.min_y, S.max_x, S.max_y = mnx, mny, mxx, mxy
            S.grid_inv_w = np.float32(64) / np.float32(mxx - mnx)      # src/Frame.cc:107-108
            S.grid_inv_h = np.float32(48) / np.float32(mxy - mny)
            S.n_q = len(pr["q_flags"])
            S.q_flags = a(pr["q_flags"], np.uint8); S.q_desc = a(pr["q_desc"], np.uint8)
            track = level = None
            if "mp_xw" in pr:      # variant 1 with Frame::isInFrustum on the device
                S.mp_xw = a(pr["mp_xw"], np.float32); S.mp_normal = a(pr["mp_normal"], np.float32)
                S.mp_min_distance = a(pr["mp_min_distance"], np.float32); S.mp_max_distance = a(pr["mp_max_distance"], np.float32)
                S.ow = a(pr["ow"], np.float32); S.tcw_cur = a(pr["tcw_cur"], np.float32)
                S.fx, S.fy, S.cx, S.cy, S.mbf, S.mb = [float(v) for v in pr["cam"]]
                sf = np.ascontiguousarray(pr["scale_factors"], np.float32); keep.append(sf)
                S.scale_factors = sf.ctypes.data; S.n_levels = len(sf)
                S.th = float(pr.get("th", 1.0)); S.view_cos_limit = float(pr.get("view_cos_limit", 0.5))
                S.log_scale_factor = float(pr["log_scale_factor"])
                S.use_ratio = int(pr.get("use_ratio", 1)); S.nn_ratio = float(pr.get("nn_ratio", self.mfNNratio)); S.check_orientation = 0
                if pr.get("fuse"):     # ORBmatcher::Fuse candidate search (src/ORBmatcher.cc:825-975)
                    S.fuse = 1; S.inv_level_sigma2 = a(pr["inv_level_sigma2"], np.float32)
                track = np.zeros((S.n_q, 4), np.float32); level = np.full(S.n_q, -1, np.int32); keep += [track, level]
                S.q_track, S.q_level = track.ctypes.data, level.ctypes.data
            elif "last_xw" in pr:
                S.last_xw = a(pr["last_xw"], np.float32); S.last_octave = a(pr["last_octave"], np.int32)
                S.tcw_cur = a(pr["tcw_cur"], np.float32); S.tcw_last = a(pr["tcw_last"], np.float32)
                S.fx, S.fy, S.cx, S.cy, S.mbf, S.mb = [float(v) for v in pr["cam"]]
                sf = np.ascontiguousarray(pr["scale_factors"], np.float32); keep.append(sf)
                S.scale_factors = sf.ctypes.data; S.n_levels = len(sf)
                S.th = float(pr["th"]); S.mono = int(pr.get("mono", 0))
                S.q_angle = a(pr["q_angle"], np.float32)
                S.check_orientation = int(pr.get("check_orientation", self.mbCheckOrientation)); S.use_ratio = 0
            else:
                for k in ("q_u", "q_v", "q_ur", "q_radius"):
                    setattr(S, k, a(pr[k], np.float32))
                S.q_min_level = a(pr["q_min_level"], np.int32); S.q_max_level = a(pr["q_max_level"], np.int32)
                S.use_ratio = int(pr.get("use_ratio", 1)); S.nn_ratio = float(pr.get("nn_ratio", self.mfNNratio))
                S.check_orientation = int(pr.get("check_orientation", 0))
                if S.check_orientation:
                    S.q_angle = a(pr["q_angle"], np.float32)
            km = np.full(S.n_kp, -1, np.int32); bi = np.full(S.n_q, -1, np.int32); bd = np.full(S.n_q, 256, np.int32)
            S.kp_match, S.q_best_idx, S.q_best_dist = km.ctypes.data, bi.ctypes.data, bd.ctypes.data
            outs.append((km, bi, bd) if track is None else (km, bi, bd, track, level))
        return arr, n, outs, keep

    def run_prepared_projection(self, prep):
        arr, n, outs, _keep = prep
        check(lib().adb_search_by_projection(self._m, C.byref(arr), n))
        return [(int(arr[i].n_matches),) + outs[i] for i in range(n)]

    def SearchByProjection(self, problem: dict):
        """One frame: returns (nmatches, kp_match, q_best_idx, q_best_dist)."""
        return self.search_by_projection([problem])[0]

    def search_by_bow(self, problems: list[dict]):
        """ORBmatcher::SearchByBoW(pKF, F, ...) (mode 0, src/ORBmatcher.cc:159-288) / SearchForTriangulation(pKF1, pKF2, F12, ...)
        (mode 1, src/ORBmatcher.cc:657-823) on the per-node feature lists of the two FeatureVectors.  Problem dict: `mode`,
        `kps1`, `desc1`, `flags1`, `kps2`, `desc2`, `b_ptr1`, `b_idx1`, `b_ptr2`, `b_idx2`; mode 0: `nn_ratio`; mode 1:
        `u_right1`, `u_right2`, `flags2`, `f12`, `epipole`, `scale_factors2`, `level_sigma2_2`.
        Returns a list of (nmatches, match): match21 [n2] for mode 0 (index on side 1 or -1), match12 [n1] for mode 1."""
        from .capi import BowSearch
        n = len(problems)
        arr = (BowSearch * n)()
        keep, outs = [], []

        def a(x, dt):
            x = np.ascontiguousarray(x, dt); keep.append(x); return x.ctypes.data

        for i, pr in enumerate(problems):
            S = arr[i]
            S.mode = int(pr["mode"])
            k1 = np.ascontiguousarray(pr["kps1"], KP_DTYPE); k2 = np.ascontiguousarray(pr["kps2"], KP_DTYPE); keep += [k1, k2]
            S.n1, S.n2 = len(k1), len(k2)
            S.kps1, S.kps2 = k1.ctypes.data, k2.ctypes.data
            S.desc1 = a(pr["desc1"], np.uint8); S.desc2 = a(pr["desc2"], np.uint8); S.flags1 = a(pr["flags1"], np.uint8)
            S.n_buckets = len(pr["b_ptr1"]) - 1
            S.b_ptr1 = a(pr["b_ptr1"], np.int32); S.b_idx1 = a(pr["b_idx1"], np.int32)
            S.b_ptr2 = a(pr["b_ptr2"], np.int32); S.b_idx2 = a(pr["b_idx2"], np.int32)
            S.nn_ratio = float(pr.get("nn_ratio", self.mfNNratio))
            S.check_orientation = int(pr.get("check_orientation", self.mbCheckOrientation))
            if S.mode == 1:
                S.u_right1 = a(pr["u_right1"], np.float32); S.u_right2 = a(pr["u_right2"], np.float32); S.flags2 = a(pr["flags2"], np.uint8)
                S.f12 = a(pr["f12"], np.float32); S.ex, S.ey = [float(v) for v in pr["epipole"]]
                sf = np.ascontiguousarray(pr["scale_factors2"], np.float32); keep.append(sf)
                S.scale_factors2 = sf.ctypes.data; S.n_levels = len(sf); S.level_sigma2_2 = a(pr["level_sigma2_2"], np.float32)
                out = np.full(S.n1, -1, np.int32); S.match12 = out.ctypes.data
            else:
                out = np.full(S.n2, -1, np.int32); S.match21 = out.ctypes.data
            outs.append(out)
        check(lib().adb_search_by_bow(self._m, C.byref(arr), n))
        return [(int(arr[i].n_matches), outs[i]) for i in range(n)]

    def search_last_ms(self) -> float:
        ms = C.c_float()
        check(lib().adb_search_last_ms(self._m, C.byref(ms)))
        return float(ms.value)


def compute_distinctive_descriptors(matcher: ORBmatcher, desc: np.ndarray, point_ptr: np.ndarray):
    """MapPoint::ComputeDistinctiveDescriptors for a batch of map points (CSR over observation descriptors) ->
    (best_idx [P], best_desc [P, 32])."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    point_ptr = np.ascontiguousarray(point_ptr, np.int32)
    npts = len(point_ptr) - 1
    bi = np.zeros(npts, np.int32); bd = np.zeros((npts, 32), np.uint8)
    check(lib().adb_distinctive_descriptors(matcher._m, ptr(desc), ptr(point_ptr), npts, ptr(bi), ptr(bd)))
    return bi, bd


def compute_stereo_matches(left: ORBextractor, right: ORBextractor, n_frames: int, mb: float, mbf: float,
                           out: HostStereo | None = None):
    """Frame::ComputeStereoMatches for the frames resident in two extractor handles ->
    (uRight [F, cap], depth [F, cap], best_idx [F, cap], best_dist [F, cap])."""
    cap = left.capacity
    if out is None:
        out = HostStereo(n_frames, cap)
    ur, dp, bi, bd = out.u_right[:n_frames], out.depth[:n_frames], out.best_idx[:n_frames], out.best_dist[:n_frames]
    check(lib().adb_stereo_match(left._h, right._h, n_frames, mb, mbf, ptr(ur), ptr(dp), ptr(bi), ptr(bd), cap))
    return ur, dp, bi, bd


def stereo_frames_batch(left: ORBextractor, right: ORBextractor, images_left: np.ndarray, images_right: np.ndarray, mb: float, mbf: float,
                        masks_left: np.ndarray | None = None, masks_right: np.ndarray | None = None, out_left: HostResults | None = None,
                        out_right: HostResults | None = None, out_stereo: HostStereo | None = None):
    """The stereo Frame constructor's hot part (src/Frame.cc:80-100) for F stereo pairs in host memory, as one chunk pipeline:
    -> ((kps, desc, counts) left, (kps, desc, counts) right, (uRight, depth, best_idx, best_dist))."""
    images_left = np.ascontiguousarray(images_left, np.uint8); images_right = np.ascontiguousarray(images_right, np.uint8)
    assert images_left.shape == images_right.shape
    f, h, w = images_left.shape
    cap = left.capacity
    out_left = out_left or HostResults(f, cap); out_right = out_right or HostResults(f, cap); out_stereo = out_stereo or HostStereo(f, cap)
    if masks_left is not None:
        masks_left = np.ascontiguousarray(masks_left, np.uint8); masks_right = np.ascontiguousarray(masks_right, np.uint8)
        assert masks_left.shape == images_left.shape and masks_right.shape == images_left.shape
    kl, dl, cl = out_left.kps[:f], out_left.desc[:f], out_left.counts[:f]
    kr, dr, cr = out_right.kps[:f], out_right.desc[:f], out_right.counts[:f]
    ur, dp, bi, bd = out_stereo.u_right[:f], out_stereo.depth[:f], out_stereo.best_idx[:f], out_stereo.best_dist[:f]
    check(lib().adb_stereo_frames_batch(left._h, right._h, f, ptr(images_left), ptr(images_right), h * w, w, h, w, ptr(masks_left), ptr(masks_right),
                                        h * w, w, ptr(kl), ptr(dl), ptr(cl), ptr(kr), ptr(dr), ptr(cr), cap, mb, mbf, ptr(ur), ptr(dp), ptr(bi), ptr(bd)))
    left._after_call(w, h); right._after_call(w, h)
    return (kl, dl, cl), (kr, dr, cr), (ur, dp, bi, bd)


def stereo_match_device(left: ORBextractor, right: ORBextractor, n_frames: int, mb: float, mbf: float):
    """Device-resident ComputeStereoMatches (asynchronous on the left handle's stream)."""
    check(lib().adb_stereo_match_device(left._h, right._h, n_frames, mb, mbf))
