"""The reference's YAML config surface (configurations/*.cfg, read by apps/rgbd_runner.py:31-75 and
apps/ply_runner.py / kitti_runner.py) mapped onto GeoWrapper's constructor keywords, so that a
reference .cfg file drives this implementation unchanged:

    from mrhash_b200.config import load_config, wrapper_from_config
    geo, sensor = wrapper_from_config("configurations/replica.cfg")      # GeoWrapper + camera set
    kwargs, sensor = load_config("configurations/vbr.cfg")                # or just the keyword dict

Sections and keys are the reference's: map.{sdf_truncation, sdf_truncation_scale,
integration_weight_sample, virtual_voxel_size, n_frames_invalidate_voxels},
streamer.voxel_extents_scale, mesh.{marching_cubes_threshold, min_weight_threshold, sdf_var_threshold,
vertices_merging_threshold}, sensor.{min_depth, max_depth[, intrinsics, resolution, depth_scaling, hz,
rosbag_topic]}, data_path, results_path, end_frame."""
import yaml

_MAP_KEYS = ("sdf_truncation", "sdf_truncation_scale", "integration_weight_sample", "virtual_voxel_size", "n_frames_invalidate_voxels")
_MESH_KEYS = ("marching_cubes_threshold", "min_weight_threshold", "sdf_var_threshold", "vertices_merging_threshold")


def load_config(path_or_dict):
    """Returns (GeoWrapper keyword dict, sensor dict). Missing required keys raise KeyError naming them,
    like the runner's own dictionary look-ups do."""
    if isinstance(path_or_dict, dict):
        cf = path_or_dict
    else:
        with open(path_or_dict, "r") as f:
            cf = yaml.safe_load(f)
    kw = {k: cf["map"][k] for k in _MAP_KEYS}
    kw["voxel_extents_scale"] = cf["streamer"]["voxel_extents_scale"]
    kw.update({k: cf["mesh"][k] for k in _MESH_KEYS})
    sensor = dict(cf["sensor"])
    kw["min_depth"] = sensor["min_depth"]
    kw["max_depth"] = sensor["max_depth"]
    kw["viewer_active"] = False  # every runner passes False (rgbd_runner.py:113)
    kw["projective_sdf"] = True  # rgbd_runner.py:118
    for extra in ("data_path", "results_path", "end_frame"):
        if extra in cf:
            sensor[extra] = cf[extra]
    return kw, sensor


def wrapper_from_config(path_or_dict, camera_model=0, **sizing):
    """GeoWrapper built from a reference config; the camera is set when the config carries pinhole
    intrinsics (sensor.intrinsics = [fx, fy, cx, cy], sensor.resolution = [cols, rows]). `sizing` takes
    the keyword-only additions (num_sdf_blocks=, hash_num_buckets=, max_num_triangles=, device=, ...)."""
    from .geowrapper import GeoWrapper

    kw, sensor = load_config(path_or_dict)
    geo = GeoWrapper(**kw, **sizing)
    if "intrinsics" in sensor and "resolution" in sensor:
        fx, fy, cx, cy = sensor["intrinsics"]
        cols, rows = sensor["resolution"]
        geo.setCamera(fx, fy, cx, cy, rows, cols, kw["min_depth"], kw["max_depth"], camera_model)
    return geo, sensor
