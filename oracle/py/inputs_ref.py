"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's input-path point arithmetic.

  merge_sweeps_np   second/data/nuscenes_dataset.py:166-185 -- the statements are kept in the reference's own order and
                    dtypes (float32 sweep array, float64 calibration, in-place stores) because the rounding points matter
  crop_convex_np    second/core/geometry.py:358-395 ``_points_in_convex_polygon_3d_jit`` (sign >= 0 -> outside), float64

tests/test_inputs.py pins both against the UNMODIFIED reference functions in the build container.
"""
import numpy as np


def merge_sweeps_np(sweeps, rotations, translations, time_lags):
    out = []
    for i, pts in enumerate(sweeps):
        p = np.array(pts, dtype=np.float32, copy=True)
        if i == 0:
            dt = np.zeros(p.shape[0], np.float32)
        else:
            p[:, :3] = p[:, :3] @ np.asarray(rotations[i], np.float64).T       # float64 product stored as float32
            p[:, :3] += np.asarray(translations[i], np.float64)                # float32 += float64
            dt = np.full(p.shape[0], time_lags[i], dtype=np.float32)
        out.append(np.concatenate([p[:, :3], dt[:, None]], axis=1))
    return np.concatenate(out, axis=0).astype(np.float32)


def crop_convex_np(points, planes):
    p = points[:, :3].astype(np.float64)
    keep = np.ones(points.shape[0], bool)
    for a, b, c, d in np.asarray(planes, np.float64):
        s = p[:, 0] * a + p[:, 1] * b + p[:, 2] * c + d
        keep &= ~(s >= 0)
    return points[keep]
