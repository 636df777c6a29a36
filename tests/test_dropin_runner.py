"""GPU: the drop-in boundary, executed. The reference's own application apps/rgbd_runner.py runs
UNMODIFIED (staged next to the other reference build outputs in oracle/_ref/apps by oracle/Makefile;
git-ignored, never part of this repository) with `PYTHONPATH=<this repo>`, so that its
`from mrhash.src.pygeowrapper import GeoWrapper` (rgbd_runner.py:9) resolves to mrhash_b200 through the
import shim in mrhash/. Input: a temporary directory in the Replica layout the runner's DepthReader
expects (apps/utils/depth_reader.py:27-45: results/*.png 16-bit depth, results/*.jpg colour, traj.txt)
written from the synthetic stream, and configurations/replica.cfg with only the paths, the end frame
and the camera (resolution / intrinsics of the synthetic stream) replaced. natsort and matplotlib are
not installed in this image: two-line stand-ins are put on the path (natural sort of the file names;
camera.py only imports pyplot). Checks: the runner exits 0, writes the mesh and the two point clouds,
and the voxel cloud equals, record for record, the one a direct GeoWrapper session produces from the
same decoded files."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from test_parity_mesh import _ply_payload_sorted

from mrhash_b200 import GeoWrapper, synth

pytestmark = pytest.mark.gpu

APPS = os.path.join(ROOT, "oracle", "_ref", "apps")
CFG = os.path.join(ROOT, "oracle", "_ref", "configurations", "replica.cfg")
W, H, N = 320, 240, 8


def _write_dataset(root):
    from PIL import Image

    os.makedirs(os.path.join(root, "results"))
    poses = []
    for k in range(N):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=2000, width=W, height=H)
        Image.fromarray(np.round(depth * 6553.5).astype(np.uint16)).save(os.path.join(root, "results", f"depth{k:06d}.png"))
        Image.fromarray(rgb).save(os.path.join(root, "results", f"frame{k:06d}.jpg"), quality=95)
        poses.append(np.asarray(synth.quat_to_matrix_f32(t, q), np.float64).reshape(16))
    np.savetxt(os.path.join(root, "traj.txt"), np.stack(poses), delimiter=" ")


@pytest.mark.skipif(not os.path.exists(os.path.join(APPS, "rgbd_runner.py")), reason="reference runner not staged (oracle/Makefile, target ref)")
def test_unmodified_rgbd_runner_runs_on_mrhash_b200(tmp_path):
    import yaml

    data, results, stubs = str(tmp_path / "replica_room"), str(tmp_path / "out"), str(tmp_path / "stubs")
    _write_dataset(data)
    os.makedirs(os.path.join(stubs, "matplotlib"))
    open(os.path.join(stubs, "natsort.py"), "w").write(
        "import re\n\ndef natsorted(seq):\n    return sorted(seq, key=lambda p: [int(s) if s.isdigit() else s for s in re.split(r'(\\d+)', str(p))])\n"
    )
    open(os.path.join(stubs, "matplotlib", "__init__.py"), "w").write("")
    open(os.path.join(stubs, "matplotlib", "pyplot.py"), "w").write("")
    cfg = yaml.safe_load(open(CFG))
    fx, fy, cx, cy = synth.intrinsics(W, H)
    cfg["data_path"], cfg["results_path"], cfg["end_frame"] = data, results, -1
    cfg["sensor"]["resolution"] = [W, H]
    cfg["sensor"]["intrinsics"] = [float(fx), float(fy), float(cx), float(cy)]
    cfg_path = str(tmp_path / "replica.cfg")
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, stubs]), MRH_NUM_SDF_BLOCKS="60000")
    r = subprocess.run([sys.executable, os.path.join(APPS, "rgbd_runner.py"), cfg_path], cwd=APPS, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = sorted(os.listdir(results))
    mesh = [f for f in out if f.startswith("mesh_") and f.endswith(".ply")]
    hashp = [f for f in out if f.startswith("hash_points_")]
    voxp = [f for f in out if f.startswith("voxel_points_")]
    assert len(mesh) == 1 and len(hashp) == 1 and len(voxp) == 1, out
    assert os.path.getsize(os.path.join(results, mesh[0])) > 100000

    # the same frames, decoded exactly as the runner's DepthReader decodes them, through a direct session
    sys.path.insert(0, APPS)
    sys.path.insert(0, stubs)
    from pathlib import Path

    from utils.depth_reader import DepthReader

    reader = DepthReader(Path(data), min_range=cfg["sensor"]["min_depth"], max_range=cfg["sensor"]["max_depth"], depth_scaling=cfg["sensor"]["depth_scaling"], sensor_hz=30)
    g = GeoWrapper(
        sdf_truncation=cfg["map"]["sdf_truncation"], sdf_truncation_scale=cfg["map"]["sdf_truncation_scale"], integration_weight_sample=cfg["map"]["integration_weight_sample"],
        virtual_voxel_size=cfg["map"]["virtual_voxel_size"], n_frames_invalidate_voxels=cfg["map"]["n_frames_invalidate_voxels"], voxel_extents_scale=cfg["streamer"]["voxel_extents_scale"],
        viewer_active=False, marching_cubes_threshold=cfg["mesh"]["marching_cubes_threshold"], min_weight_threshold=cfg["mesh"]["min_weight_threshold"],
        sdf_var_threshold=cfg["mesh"]["sdf_var_threshold"], vertices_merging_threshold=cfg["mesh"]["vertices_merging_threshold"], projective_sdf=True,
        min_depth=cfg["sensor"]["min_depth"], max_depth=cfg["sensor"]["max_depth"], num_sdf_blocks=60000, hash_num_buckets=30000, max_num_triangles=2_000_000,
    )
    g.setCamera(fx, fy, cx, cy, H, W, cfg["sensor"]["min_depth"], cfg["sensor"]["max_depth"], 0)
    for frame, pose, quat, depth_img, rgb_img in reader:
        g.setCurrPose(pose, quat)
        g.setDepthImage(depth_img)
        g.setRGBImage(rgb_img)
        g.compute()
    g.streamAllOut()
    hp, vp = str(tmp_path / "direct_hash.ply"), str(tmp_path / "direct_voxel.ply")
    g.serializeData(hp, vp)
    pa, da = _ply_payload_sorted(os.path.join(results, voxp[0]))
    pb, db = _ply_payload_sorted(vp)
    assert pa == pb and da.shape == db.shape and np.array_equal(da, db)
    print(f"[drop-in] rgbd_runner.py (unmodified) on mrhash_b200: {N} frames, {len(da)} voxel points identical to a direct session; outputs {out}")
