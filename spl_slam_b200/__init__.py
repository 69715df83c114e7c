"""B200-native point-line feature front-end (drop-in for SPL-SLAM's ORB / LSD+LBD / Hamming hot path)."""
from .api import (Context, ORBextractor, Lineextractor, FldLineextractor, Linematcher, ORBmatcher, PlfError, load,
                  KEYPOINT_DTYPE, KEYLINE_DTYPE, distribute_octree, GridParams, features_in_area, ORBVocabulary, Camera,
                  undistort_keypoints, undistort_keylines)
