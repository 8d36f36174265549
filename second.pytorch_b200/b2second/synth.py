"""Seeded synthetic LiDAR clouds (no datasets offline) -- SURVEY.md §8(d) / App. D spec.

KITTI-style: 64-beam ray cast (elevation -24.8..+2 deg, azimuth +-45 deg step 0.09 deg) onto the
ground plane z=-1.73 plus random boxes, sigma=1 cm noise, cropped to the point-cloud range, uniformly
subsampled (index-sorted) to P points, reflectance U[0,1).  Point layout ``[x, y, z, r]`` as
second/data/kitti_dataset.py:202-204.

NuScenes-style: 360 deg, 32 beams, 10 sweeps, layout ``[x, y, z, dt]`` as
second/data/nuscenes_dataset.py:166-185 (intensity column dropped, time lag appended).
"""
import numpy as np


def _raycast(rng, elev_deg, azim_deg, n_boxes, box_range, ground_z=-1.73, max_range=120.0):
    el = np.deg2rad(elev_deg)[:, None]
    az = np.deg2rad(azim_deg)[None, :]
    dx = (np.cos(el) * np.cos(az)).ravel()
    dy = (np.cos(el) * np.sin(az)).ravel()
    dz = (np.sin(el) * np.ones_like(az)).ravel()
    t = np.full(dx.shape, max_range, dtype=np.float64)
    # ground plane
    down = dz < -1e-6
    tg = np.where(down, ground_z / np.where(down, dz, -1.0), max_range)
    t = np.minimum(t, tg)
    # axis-aligned boxes standing on the ground (slab test)
    cx = rng.uniform(box_range[0], box_range[1], n_boxes)
    cy = rng.uniform(box_range[2], box_range[3], n_boxes)
    sx = rng.uniform(1.5, 4.5, n_boxes)
    sy = rng.uniform(1.5, 4.5, n_boxes)
    sz = rng.uniform(1.2, 2.5, n_boxes)
    for b in range(n_boxes):
        lo = np.array([cx[b] - sx[b] / 2, cy[b] - sy[b] / 2, ground_z])
        hi = np.array([cx[b] + sx[b] / 2, cy[b] + sy[b] / 2, ground_z + sz[b]])
        with np.errstate(divide="ignore", invalid="ignore"):
            t0x, t1x = lo[0] / dx, hi[0] / dx
            t0y, t1y = lo[1] / dy, hi[1] / dy
            t0z, t1z = lo[2] / dz, hi[2] / dz
        tn = np.maximum(np.maximum(np.minimum(t0x, t1x), np.minimum(t0y, t1y)), np.minimum(t0z, t1z))
        tf = np.minimum(np.minimum(np.maximum(t0x, t1x), np.maximum(t0y, t1y)), np.maximum(t0z, t1z))
        hit = (tn <= tf) & (tn > 0.5)
        t = np.where(hit & (tn < t), tn, t)
    ok = t < max_range
    pts = np.stack([dx * t, dy * t, dz * t], axis=1)[ok]
    return pts


def kitti_cloud(seed, num_points=29000, pc_range=(0, -40, -3, 70.4, 40, 1)):
    """-> float32 [P,4] (x,y,z,r).  P=20000 gives ~13.7k car.fhd voxels, P~29000 gives ~17k."""
    rng = np.random.default_rng(seed)
    elev = np.linspace(-24.8, 2.0, 64)
    azim = np.arange(-45.0, 45.0, 0.09)
    pts = _raycast(rng, elev, azim, 40, (5, 65, -35, 35))
    pts = pts + rng.normal(0.0, 0.01, pts.shape)
    r = np.asarray(pc_range, dtype=np.float64)
    m = ((pts[:, 0] >= r[0]) & (pts[:, 0] < r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] < r[4])
         & (pts[:, 2] >= r[2]) & (pts[:, 2] < r[5]))
    pts = pts[m]
    if pts.shape[0] > num_points:
        sel = np.sort(rng.choice(pts.shape[0], num_points, replace=False))
        pts = pts[sel]
    refl = rng.uniform(0.0, 1.0, (pts.shape[0], 1))
    return np.concatenate([pts, refl], axis=1).astype(np.float32)


def nuscenes_cloud(seed, num_points=300000, sweeps=10, pc_range=(-50, -50, -5, 50, 50, 3)):
    """-> float32 [P,4] (x,y,z,dt): 10 merged 360-degree sweeps, ego drifting forward."""
    rng = np.random.default_rng(seed)
    elev = np.linspace(-30.0, 10.0, 32)
    azim = np.arange(-180.0, 180.0, 0.28)
    per = []
    for s in range(sweeps):
        rs = np.random.default_rng(seed)  # same scene, shifted ego
        pts = _raycast(rs, elev, azim, 60, (-45, 45, -45, 45), ground_z=-1.84)
        pts = pts + rng.normal(0.0, 0.02, pts.shape)
        pts[:, 0] -= 0.5 * s  # ego motion 10 m/s at 20 Hz
        dt = np.full((pts.shape[0], 1), 0.05 * s)
        per.append(np.concatenate([pts, dt], axis=1))
    pts = np.concatenate(per, axis=0)
    r = np.asarray(pc_range, dtype=np.float64)
    m = ((pts[:, 0] >= r[0]) & (pts[:, 0] < r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] < r[4])
         & (pts[:, 2] >= r[2]) & (pts[:, 2] < r[5]))
    pts = pts[m]
    if pts.shape[0] > num_points:
        sel = np.sort(rng.choice(pts.shape[0], num_points, replace=False))
        pts = pts[sel]
    return pts.astype(np.float32)
