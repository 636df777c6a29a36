"""Seeded synthetic sensor streams (scenes S1-S4). No dataset, no network.

RGB-D: an axis-aligned room x in [-3,3], y in [-1.5,1.5], z in [-2,2] m with one sphere
(r = 0.4 m at (0.5, 0.3, 1.2)), ray-cast analytically in the camera model of the reference
(camera.cuh:84-91: ray = (ifx*(col-cx-0.5), ify*(row-cy-0.5), 1), so depth = camera z), quantised to
1/6553.5 m like Replica PNGs (configurations/replica.cfg:22). LiDAR: 128 beams x 1024 azimuths in a
60 x 60 x 15 m box, spherical model.
"""
import numpy as np

# configurations/scannet.cfg:20 (640x480)
SCANNET_K = (577.590698, 578.729797, 318.905426, 242.683609)
ROOM_MIN = np.array([-3.0, -1.5, -2.0])
ROOM_MAX = np.array([3.0, 1.5, 2.0])
SPHERE_C = np.array([0.5, 0.3, 1.2])
SPHERE_R = 0.4
DEPTH_SCALE = 6553.5

# configurations/replica.cfg:1-18
REPLICA_PARAMS = dict(
    sdf_truncation=0.07,
    sdf_truncation_scale=0.0,
    integration_weight_sample=1,
    virtual_voxel_size=0.01,
    n_frames_invalidate_voxels=100,
    voxel_extents_scale=1,
    viewer_active=False,
    marching_cubes_threshold=1.5,
    min_weight_threshold=5,
    min_depth=0.01,
    max_depth=30.0,
    sdf_var_threshold=0.0,
    vertices_merging_threshold=0.0,
    projective_sdf=True,
)

# configurations/vbr.cfg:1-22
VBR_PARAMS = dict(
    sdf_truncation=0.40,
    sdf_truncation_scale=0.0,
    integration_weight_sample=1,
    virtual_voxel_size=0.20,
    n_frames_invalidate_voxels=0,
    voxel_extents_scale=1,
    viewer_active=False,
    marching_cubes_threshold=1.5,
    min_weight_threshold=50,
    min_depth=0.2,
    max_depth=100.0,
    sdf_var_threshold=0.0,
    vertices_merging_threshold=0.0,
    projective_sdf=True,
)


def intrinsics(width=640, height=480):
    """scannet.cfg intrinsics, scaled with the resolution (S4 = 2x)."""
    s = width / 640.0
    fx, fy, cx, cy = SCANNET_K
    return fx * s, fy * s, cx * s, cy * s


def orbit_pose(k, n_frames, radius=1.0):
    """Scene S2: camera centre (cos t, 0, sin t)*radius, optical axis radially outward, +y_cam = world +y.
    Returns (translation[3] f32, quaternion xyzw[4] f32, R[3,3] f64)."""
    th = 2.0 * np.pi * k / n_frames
    z = np.array([np.cos(th), 0.0, np.sin(th)])
    y = np.array([0.0, 1.0, 0.0])
    x = np.cross(y, z)
    R = np.stack([x, y, z], axis=1)
    t = radius * z
    return t.astype(np.float32), rot_to_quat_xyzw(R).astype(np.float32), R


def rot_to_quat_xyzw(R):
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w, x, y, z = (R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w, x, y, z = (R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w, x, y, z = (R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s
    return np.array([x, y, z, w])


def quat_to_matrix_f32(t, q):
    """cam_in_world as Eigen::Quaternionf(w,x,y,z).toRotationMatrix() builds it in float32
    (geowrapper.cpp:86-92) — the matrix every implementation receives for a (t, q) pose."""
    f = np.float32
    x, y, z, w = (f(v) for v in q)
    tx, ty, tz = f(2) * x, f(2) * y, f(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    T = np.eye(4, dtype=np.float32)
    T[0, 0], T[0, 1], T[0, 2] = f(1) - (tyy + tzz), txy - twz, txz + twy
    T[1, 0], T[1, 1], T[1, 2] = txy + twz, f(1) - (txx + tzz), tyz - twx
    T[2, 0], T[2, 1], T[2, 2] = txz - twy, tyz + twx, f(1) - (txx + tyy)
    T[:3, 3] = np.asarray(t, np.float32)
    return T


def _hash_rgb(p):
    q = np.floor(p * 16.0).astype(np.int64)
    h = (q[..., 0] * 73856093) ^ (q[..., 1] * 19349669) ^ (q[..., 2] * 83492791)
    h = (h & 0xFFFFFFFF).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0x5BD1E995)) & np.uint32(0xFFFFFFFF)
    h ^= h >> np.uint32(15)
    return np.stack([(h & 0xFF), (h >> 8) & 0xFF, (h >> 16) & 0xFF], axis=-1).astype(np.uint8)


def render_rgbd(R, t, width=640, height=480, noise_sigma=0.0, rng=None, K=None):
    """Ray-cast the room + sphere. Returns (depth f32 [H,W] metres, rgb u8 [H,W,3])."""
    fx, fy, cx, cy = K if K is not None else intrinsics(width, height)
    cols, rows = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    d_cam = np.stack([(cols - cx - 0.5) / fx, (rows - cy - 0.5) / fy, np.ones_like(cols)], axis=-1)
    d = d_cam @ np.asarray(R, np.float64).T
    o = np.asarray(t, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_hi = np.where(d > 0, (ROOM_MAX - o) / d, np.where(d < 0, (ROOM_MIN - o) / d, np.inf))
    t_box = t_hi.min(axis=-1)
    oc = o - SPHERE_C
    a = (d * d).sum(-1)
    b = 2.0 * (d * oc).sum(-1)
    c = (oc * oc).sum() - SPHERE_R**2
    disc = b * b - 4 * a * c
    with np.errstate(invalid="ignore"):
        t_sph = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    t_sph = np.where(t_sph > 1e-6, t_sph, np.inf)
    depth = np.minimum(t_box, t_sph)
    hit = o + depth[..., None] * d
    rgb = _hash_rgb(hit)
    if noise_sigma > 0:
        depth = depth + (rng or np.random.default_rng(1)).normal(0.0, noise_sigma, depth.shape)
    depth = np.round(depth * DEPTH_SCALE) / DEPTH_SCALE
    return depth.astype(np.float32), rgb


def rgbd_frame(k=0, n_frames=1000, width=640, height=480, orbit=True, noise_sigma=0.0, seed=1):
    """One frame of S2/S4 (orbit=True) or S1 (orbit=False: identity pose at the room centre).
    Returns (t[3] f32, q_xyzw[4] f32, depth, rgb)."""
    if orbit:
        t, q, R = orbit_pose(k, n_frames)
    else:
        t, q, R = np.zeros(3, np.float32), np.array([0, 0, 0, 1], np.float32), np.eye(3)
    rng = np.random.default_rng(seed + k) if noise_sigma > 0 else None
    depth, rgb = render_rgbd(R, t, width, height, noise_sigma, rng)
    return t, q, depth, rgb


# ---- LiDAR (scene S3) ----
LIDAR_ROWS, LIDAR_COLS = 128, 1024
LIDAR_BOX_MIN = np.array([-30.0, -30.0, -1.8])
LIDAR_BOX_MAX = np.array([30.0, 30.0, 13.2])


def lidar_frame(k=0, rows=LIDAR_ROWS, cols=LIDAR_COLS, seed=2, noise_sigma=0.0):
    """128 beams x 1024 azimuths, vertical FOV +-22.5 deg, sensor at (k*1 m - 20 m, 0, 0) inside a
    60 x 60 x 15 m box (z-up, ground 1.8 m below the sensor). Returns (cam_in_world[4,4] f32,
    points [N,3] f32 in the sensor frame); rays farther than 100 m or nearer than 0.2 m are dropped."""
    az = (np.arange(cols) + 0.5) / cols * 2 * np.pi - np.pi
    el = np.deg2rad(-22.5 + (np.arange(rows) + 0.5) / rows * 45.0)
    A, E = np.meshgrid(az, el)
    d = np.stack([np.cos(A) * np.cos(E), np.sin(A) * np.cos(E), np.sin(E)], axis=-1).reshape(-1, 3)
    o = np.array([-20.0 + 1.0 * k, 0.0, 0.0])
    with np.errstate(divide="ignore", invalid="ignore"):
        t_hi = np.where(d > 0, (LIDAR_BOX_MAX - o) / d, np.where(d < 0, (LIDAR_BOX_MIN - o) / d, np.inf))
    r = t_hi.min(axis=-1)
    if noise_sigma > 0:
        r = r + np.random.default_rng(seed + k).normal(0.0, noise_sigma, r.shape)
    keep = (r > 0.2) & (r < 100.0)
    pts = (d * r[:, None])[keep].astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = o.astype(np.float32)
    return T, pts


def render_rgbd_torch(R, t, width=640, height=480, device="cuda"):
    """Same ray-cast as render_rgbd, evaluated with torch on `device` (bench.py generates its
    1000+ frame stream this way). Returns (depth f32 [H,W], rgb u8 [H,W,3]) tensors on `device`."""
    import torch

    f64 = torch.float64
    fx, fy, cx, cy = intrinsics(width, height)
    cols = torch.arange(width, dtype=f64, device=device)[None, :].expand(height, width)
    rows = torch.arange(height, dtype=f64, device=device)[:, None].expand(height, width)
    d_cam = torch.stack([(cols - cx - 0.5) / fx, (rows - cy - 0.5) / fy, torch.ones_like(cols)], dim=-1)
    Rt = torch.as_tensor(np.asarray(R, np.float64), device=device)
    d = d_cam @ Rt.T
    o = torch.as_tensor(np.asarray(t, np.float64), device=device)
    lo = torch.as_tensor(ROOM_MIN, device=device)
    hi = torch.as_tensor(ROOM_MAX, device=device)
    inf = torch.full_like(d, float("inf"))
    t_hi = torch.where(d > 0, (hi - o) / d, torch.where(d < 0, (lo - o) / d, inf))
    t_box = t_hi.min(dim=-1).values
    oc = o - torch.as_tensor(SPHERE_C, device=device)
    a = (d * d).sum(-1)
    b = 2.0 * (d * oc).sum(-1)
    c = (oc * oc).sum() - SPHERE_R**2
    disc = b * b - 4 * a * c
    t_sph = torch.where(disc > 0, (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a), torch.full_like(a, float("inf")))
    t_sph = torch.where(t_sph > 1e-6, t_sph, torch.full_like(a, float("inf")))
    depth = torch.minimum(t_box, t_sph)
    hit = o + depth[..., None] * d
    q = torch.floor(hit * 16.0).to(torch.int64)
    h = ((q[..., 0] * 73856093) ^ (q[..., 1] * 19349669) ^ (q[..., 2] * 83492791)) & 0xFFFFFFFF
    h = h ^ (h >> 13)
    h = (h * 0x5BD1E995) & 0xFFFFFFFF
    h = h ^ (h >> 15)
    rgb = torch.stack([h & 0xFF, (h >> 8) & 0xFF, (h >> 16) & 0xFF], dim=-1).to(torch.uint8)
    depth = torch.round(depth * DEPTH_SCALE) / DEPTH_SCALE
    return depth.to(torch.float32).contiguous(), rgb.contiguous()
