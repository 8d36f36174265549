"""Oracle voxeliser (C) against an independent pure-Python statement of the sequential algorithm
(the loop shape of second/utils/simplevis.py:34-50; semantics SURVEY.md App. A)."""
import numpy as np
import pytest


def python_voxelizer(points, voxel_size, pc_range, T, max_voxels):
    vs = np.asarray(voxel_size, np.float32)
    lo = np.asarray(pc_range[:3], np.float32)
    grid = np.round((np.asarray(pc_range[3:], np.float32) - lo) / vs).astype(np.int64)
    lookup = {}
    voxels, coors, nums = [], [], []
    for i in range(points.shape[0]):
        c = np.floor((points[i, :3] - lo) / vs).astype(np.int64)   # fp32 sub, fp32 div, floor
        if np.any(c < 0) or np.any(c >= grid):
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))
        vid = lookup.get(key)
        if vid is None:
            if len(voxels) >= max_voxels:
                continue
            vid = len(voxels)
            lookup[key] = vid
            voxels.append(np.zeros((T, points.shape[1]), np.float32))
            coors.append(key)
            nums.append(0)
        if nums[vid] < T:
            voxels[vid][nums[vid]] = points[i]
            nums[vid] += 1
    if not voxels:
        return (np.zeros((0, T, points.shape[1]), np.float32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32))
    return np.stack(voxels), np.array(coors, np.int32), np.array(nums, np.int32)


@pytest.mark.parametrize("T,max_voxels,n", [(5, 20000, 3000), (1, 20000, 2000), (3, 150, 3000), (100, 500, 4000)])
def test_oracle_voxelizer_matches_python_loop(oracle, T, max_voxels, n):
    rng = np.random.default_rng(T * 7 + n)
    pc_range = [0, -4, -3, 7.04, 4, 1]
    vs = [0.05, 0.05, 0.1] if T < 100 else [0.16, 0.16, 4]
    pts = rng.uniform([-0.5, -4.5, -3.5, 0], [7.5, 4.5, 1.5, 1], (n, 4)).astype(np.float32)
    pts[::7] = pts[3]            # repeated points -> multi-point voxels
    pts[5::11, :3] = pts[2, :3]
    gen = oracle.utils.VoxelGeneratorV2(vs, pc_range, T, max_voxels)
    res = gen.generate(pts, max_voxels)
    v, c, m = python_voxelizer(pts, vs, pc_range, T, max_voxels)
    assert res["voxel_num"] == v.shape[0]
    np.testing.assert_array_equal(res["coordinates"], c)
    np.testing.assert_array_equal(res["num_points_per_voxel"], m)
    np.testing.assert_array_equal(res["voxels"], v)
    # scratch grid restored: second call gives the same answer
    res2 = gen.generate(pts, max_voxels)
    np.testing.assert_array_equal(res2["coordinates"], c)


def test_oracle_voxelizer_empty_and_boundaries(oracle):
    gen = oracle.utils.VoxelGeneratorV2([0.5, 0.5, 0.5], [0, 0, 0, 2, 2, 2], 2, 100)
    res = gen.generate(np.zeros((0, 4), np.float32), 100)
    assert res["voxel_num"] == 0 and res["voxels"].shape == (0, 2, 4)
    pts = np.array([[0, 0, 0, 1], [2.0, 1, 1, 1], [1.9999, 1.9999, 1.9999, 1], [-1e-7, 0, 0, 1],
                    [np.nan, 0, 0, 1]], np.float32)
    res = gen.generate(pts, 100)
    # upper bound exclusive, tiny negative floors to -1 (rejected), NaN rejected
    np.testing.assert_array_equal(res["coordinates"], [[0, 0, 0], [3, 3, 3]])
    assert gen.grid_size.tolist() == [4, 4, 4] and gen.grid_size.dtype == np.int64


def test_generate_multi_gpu_is_padded(oracle):
    gen = oracle.utils.VoxelGeneratorV2([0.5, 0.5, 0.5], [0, 0, 0, 2, 2, 2], 2, 100)
    pts = np.array([[0.1, 0.1, 0.1, 1], [1.1, 0.1, 0.1, 2]], np.float32)
    res = gen.generate_multi_gpu(pts, 7)
    assert res["voxels"].shape == (7, 2, 4) and res["voxel_num"] == 2 and res["coordinates"].shape == (7, 3)
