"""airdos_b200 -- B200-native hot path of AirDOS (ORB extract + Hamming match + sparse BA).

The package is a thin host-side mirror of the reference's three numeric classes
(ORB_SLAM2::ORBextractor, ORBmatcher, Optimizer) over the C-ABI of libairdos_b200.so.
"""
from .capi import AdbError, KP_DTYPE, LIB_PATH  # noqa: F401
from .orb import ORBextractor, ORBmatcher, compute_distinctive_descriptors, compute_stereo_matches, stereo_frames_batch  # noqa: F401
