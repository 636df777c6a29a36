"""GPU parity of extractMesh (marching cubes) and serializeData against the unmodified reference
kernels (oracle/_ref) on the same map state. Triangle order is atomic-append order in both
implementations, so soups are compared as canonicalised sets."""
import os

import numpy as np
import pytest

from compare import canonical_triangles, compare_dumps
from oracle_lib import Oracle, RefCuda, ref_available
from test_parity_rgbd import NUM_BLOCKS, NUM_BUCKETS, feed, make_all

from mrhash_b200 import GeoWrapper, _capi, synth

pytestmark = pytest.mark.gpu

MAX_TRIS = 2_000_000


def build_state(n_frames, with_ref=True, width=640, height=480):
    params = dict(synth.REPLICA_PARAMS)
    fx, fy, cx, cy = synth.intrinsics(width, height)
    ours = GeoWrapper(**params, num_sdf_blocks=NUM_BLOCKS, hash_num_buckets=NUM_BUCKETS, max_num_triangles=MAX_TRIS)
    ours.setCamera(fx, fy, cx, cy, height, width, params["min_depth"], params["max_depth"], 0)
    ref = None
    if with_ref and ref_available():
        ref = RefCuda(params, NUM_BLOCKS, NUM_BUCKETS, max_num_triangles=MAX_TRIS)
        ref.set_camera(fx, fy, cx, cy, height, width, params["min_depth"], params["max_depth"], 0)
    for k in range(n_frames):
        # small steps so that voxels collect weight >= min_weight_threshold
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=2000, width=width, height=height)
        feed(ours, [ref], t, q, depth, rgb)
    return ours, ref, params


def test_marching_cubes_matches_reference(tmp_path):
    ours, ref, _ = build_state(8)
    state = ours.dumpState()
    if ref is not None:
        assert compare_dumps(state, ref.dump())["ok"]
    ply = str(tmp_path / "mesh.ply")
    ours.extractMesh(ply)
    # ASCII PLY as geowrapper.cpp:194-229 writes it: default ostream precision (= %g), uchar colours
    lines = open(ply).read().split("\n")
    assert lines[:4] == ["ply", "format ascii 1.0", f"element vertex {len(ours.getVertices())}", "property float x"]
    hdr_end = lines.index("end_header")
    v0 = lines[hdr_end + 1].split()
    assert len(v0) == 6 and v0[0] == "%g" % ours.getVertices()[0, 0]
    f0 = lines[hdr_end + 1 + len(ours.getVertices())].split()
    assert f0[0] == "3" and [int(x) for x in f0[1:]] == ours.getFaces()[0].tolist()
    assert len(lines) == hdr_end + 1 + len(ours.getVertices()) + len(ours.getFaces()) + 1
    # every vertex and face line, byte for byte, against printf's %g / %d (csrc/mrh_fmt.h is a fast exact %g)
    Vp, Fp, Cp = ours.getVertices(), ours.getFaces(), ours.getColors()
    want_v = ["%g %g %g %d %d %d" % (v[0], v[1], v[2], int(c[0]) & 0xFF, int(c[1]) & 0xFF, int(c[2]) & 0xFF) for v, c in zip(Vp.tolist(), Cp.tolist())]
    assert lines[hdr_end + 1 : hdr_end + 1 + len(Vp)] == want_v
    want_f = ["3 %d %d %d" % tuple(f) for f in Fp.tolist()]
    assert lines[hdr_end + 1 + len(Vp) : hdr_end + 1 + len(Vp) + len(Fp)] == want_f
    mine = ours.getTriangles()
    assert len(mine) > 10000
    # every block left the device (streamAllOut inside extractMesh) and sits in the host store
    assert ours.getStats()["live_blocks"] == 0 and ours.storeSize() == len(state[0])
    V, F, C = ours.getVertices(), ours.getFaces(), ours.getColors()
    assert F.min() >= 0 and F.max() < len(V) and len(C) == len(V)
    assert len(np.unique(V, axis=0)) == len(V)  # welded: no duplicate vertex
    assert (F[:, 0] != F[:, 1]).all() and (F[:, 0] != F[:, 2]).all() and (F[:, 1] != F[:, 2]).all()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    theirs, n = ref.extract_triangles(MAX_TRIS)
    assert n == len(theirs)
    print(f"triangles ours={len(mine)} ref={n}")
    assert len(mine) == n
    a, b = canonical_triangles(mine, 6), canonical_triangles(theirs, 6)
    diff = np.abs(a - b).max()
    print("max |vertex diff| after canonical sort:", diff)
    assert diff <= 1e-6
    # bit-exact positions and colours as multisets
    ka = np.sort(np.ascontiguousarray(mine.reshape(len(mine), -1)).view([("", mine.dtype)] * 18).ravel())
    kb = np.sort(np.ascontiguousarray(theirs.reshape(len(theirs), -1)).view([("", theirs.dtype)] * 18).ravel())
    same = int((ka == kb).sum())
    print(f"bit-identical triangles: {same}/{n}")
    assert same == n


def test_halo_path_equals_generic_path():
    """The shared-memory sampler and the per-read hash sampler must give the same soup bit for bit."""
    a, _, _ = build_state(6, with_ref=False)
    b, _, _ = build_state(6, with_ref=False)
    a.extractMesh(None)
    b.extractMesh(None, force_generic=True)
    ta, tb = a.getTriangles(), b.getTriangles()
    assert len(ta) == len(tb) and len(ta) > 0
    ka = np.sort(np.ascontiguousarray(ta.reshape(len(ta), -1)).view([("", ta.dtype)] * 18).ravel())
    kb = np.sort(np.ascontiguousarray(tb.reshape(len(tb), -1)).view([("", tb.dtype)] * 18).ravel())
    assert (ka == kb).all()


def test_in_place_meshing_equals_round_trip(tmp_path, monkeypatch):
    """A fully resident one-region map is meshed where it lies; MRH_MESH_ROUND_TRIP=1 forces the
    reference's streamAllOut -> streamInToGPU -> streamAllOut sequence. Same soup, same host store,
    same device state afterwards."""
    a, _, _ = build_state(6, with_ref=False)
    b, _, _ = build_state(6, with_ref=False)
    blocks = a.getStats()["live_blocks"]
    a.extractMesh(str(tmp_path / "a.ply"))
    monkeypatch.setenv("MRH_MESH_ROUND_TRIP", "1")
    b.extractMesh(str(tmp_path / "b.ply"))
    monkeypatch.delenv("MRH_MESH_ROUND_TRIP")
    ta, tb = a.getTriangles(), b.getTriangles()
    assert len(ta) == len(tb) and len(ta) > 0
    ka = np.sort(np.ascontiguousarray(ta.reshape(len(ta), -1)).view([("", ta.dtype)] * 18).ravel())
    kb = np.sort(np.ascontiguousarray(tb.reshape(len(tb), -1)).view([("", tb.dtype)] * 18).ravel())
    assert (ka == kb).all()
    for g in (a, b):
        st = g.getStats()
        assert st["live_blocks"] == 0 and g.storeSize() == blocks
        assert st["heap_free"] == NUM_BLOCKS
    # and the map comes back from the store the same either way
    a.serializeData(str(tmp_path / "ha.ply"), str(tmp_path / "va.ply"))
    b.serializeData(str(tmp_path / "hb.ply"), str(tmp_path / "vb.ply"))
    for name in ("v", "h"):
        pa, ra = _ply_payload_sorted(str(tmp_path / f"{name}a.ply"))
        pb, rb = _ply_payload_sorted(str(tmp_path / f"{name}b.ply"))
        assert pa == pb and ra.shape == rb.shape and (ra == rb).all()


def _read_ply_points(path):
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            header += f.readline()
        data = f.read()
    lines = header.decode().split("\n")
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    props = [l.split()[1:] for l in lines if l.startswith("property")]
    dt = np.dtype([(name, {"float": "<f4", "uchar": "u1"}[t]) for t, name in props])
    return np.frombuffer(data, dt, n), [name for _, name in props]


def test_serialize_data_round_trip(tmp_path):
    """serializeData (streamer.cpp:104-160): one point per voxel with weight > 0, one per block."""
    ours, _, params = build_state(3, with_ref=False)
    entries, voxels = ours.dumpState()
    ours.streamAllOut()
    hp, vp = str(tmp_path / "hash.ply"), str(tmp_path / "voxel.ply")
    ours.serializeData(hp, vp)
    hpts, hprops = _read_ply_points(hp)
    vpts, vprops = _read_ply_points(vp)
    assert vprops == ["x", "y", "z", "sdf", "weight", "red", "green", "blue", "alpha"]
    assert hprops == ["x", "y", "z", "weight", "red", "green", "blue", "alpha"]
    w = voxels["weight"]
    assert len(vpts) == int((w > 0).sum())
    assert len(hpts) == int(((w > 0).sum(axis=1) > 0).sum())
    assert (vpts["red"] == 255).all() and (vpts["green"] == 0).all()
    # voxel positions / sdf / weight against the dump
    s = np.float32(params["virtual_voxel_size"])
    bi, vi = np.nonzero(w > 0)
    lx, ly, lz = vi % 8, (vi // 8) % 8, vi // 64
    exp = np.stack(
        [
            (entries[bi, 0] * 8).astype(np.float32) * s + lx.astype(np.float32) * s,
            (entries[bi, 1] * 8).astype(np.float32) * s + ly.astype(np.float32) * s,
            (entries[bi, 2] * 8).astype(np.float32) * s + lz.astype(np.float32) * s,
            voxels["sdf"][bi, vi],
            w[bi, vi].astype(np.float32),
        ],
        axis=1,
    )
    got = np.stack([vpts["x"], vpts["y"], vpts["z"], vpts["sdf"], vpts["weight"]], axis=1)
    assert np.array_equal(exp[np.lexsort(exp.T[::-1])], got[np.lexsort(got.T[::-1])])
    # a second streamAllOut + clearBuffers leaves nothing behind
    ours.clearBuffers()
    assert ours.storeSize() == 0


def _first_seen_weld(tris, eps):
    """numpy restatement of MeshExtractor::processTriangles (mesh_extractor.cpp:9-76,156-259):
    vertices numbered in first-seen order of the soup, degenerate and repeated faces dropped."""
    P = np.ascontiguousarray(tris[:, :, :3].reshape(-1, 3))
    Cc = tris[:, :, 3:].reshape(-1, 3)
    if eps == 0.0:
        keys = P.view(np.uint32).astype(np.int64)
    else:
        keys = np.floor(P.astype(np.float64) * (1.0 / eps)).astype(np.int64)
    _, first_idx, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")
    rank = np.empty(len(order), np.int64)
    rank[order] = np.arange(len(order))
    ids = rank[np.asarray(inverse).ravel()].reshape(-1, 3)
    V = P[first_idx[order]].astype(np.float64)
    C = Cc[first_idx[order]].astype(np.float64)
    F = ids[(ids[:, 0] != ids[:, 1]) & (ids[:, 0] != ids[:, 2]) & (ids[:, 1] != ids[:, 2])]
    _, ff = np.unique(F, axis=0, return_index=True)
    return V, F[np.sort(ff)].astype(np.int32), C


@pytest.mark.parametrize("eps", [0.0, 0.004])
def test_device_weld_matches_first_seen_merge(eps):
    params = dict(synth.REPLICA_PARAMS)
    params["vertices_merging_threshold"] = eps
    fx, fy, cx, cy = synth.intrinsics(640, 480)
    g = GeoWrapper(**params, num_sdf_blocks=NUM_BLOCKS, hash_num_buckets=NUM_BUCKETS, max_num_triangles=MAX_TRIS)
    g.setCamera(fx, fy, cx, cy, 480, 640, params["min_depth"], params["max_depth"], 0)
    for k in range(7):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=2000)
        feed(g, [], t, q, depth, rgb)
    g.extractMesh(None)
    tris = g.getTriangles()
    assert len(tris) > 10000
    V, F, C = _first_seen_weld(tris, eps)
    gv, gf, gc = g.getVertices(), g.getFaces(), g.getColors()
    print(f"[weld eps={eps}] triangles {len(tris)} -> vertices {len(gv)} faces {len(gf)} (restatement {len(V)} / {len(F)})")
    assert gv.shape == V.shape and gf.shape == F.shape
    assert np.array_equal(gv, V) and np.array_equal(gc, C) and np.array_equal(gf, F)
    # welding twice gives the same mesh (the device weld is deterministic)
    g2v = g.getVertices()
    assert np.array_equal(g2v, gv)


# ---------------------------------------------------------------------------------------------
# a14 pinned against the reference's OWN host code: mesh_extractor.cpp is compiled, unmodified,
# into oracle/_ref (oracle/Makefile) and driven through ref_process_soup.
# ---------------------------------------------------------------------------------------------
def _weld_ours(g, soup, eps):
    """mrh_weld_device_soup over a soup uploaded to the device."""
    import torch

    g._set("VerticesMergingThreshold", eps)
    d = torch.from_numpy(np.ascontiguousarray(soup, np.float32)).cuda()
    _capi.check(_capi.lib().mrh_weld_device_soup(g._h, d.data_ptr(), len(soup), None))
    return g.getVertices(), g.getFaces(), g.getColors()


def _hand_soups():
    """Soups that hit every branch of processTriangles, among them the hand meshes of the reference's
    tests/test_marching_cubes.cpp:126-215 (REMOVE_DUPL_VERTICES basic_zero / basic_nonzero) as triangles."""
    rng = np.random.default_rng(5)
    out = {}
    # tests/test_marching_cubes.cpp:126-139: vertices 0 and 3 coincide
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 0]], np.float32)
    f = np.array([[0, 1, 2], [3, 1, 2]])
    out["ref_basic_zero"] = np.concatenate([v[f], np.zeros((2, 3, 3), np.float32)], axis=2)
    # :171-184: vertex 3 within epsilon of vertex 0
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.0005, 0.0005, 0.0005]], np.float32)
    out["ref_basic_nonzero"] = np.concatenate([v[f], np.ones((2, 3, 3), np.float32)], axis=2)
    # :91-101 duplicate faces, after welding
    v = rng.random((6, 3)).astype(np.float32)
    f = np.array([[0, 1, 2], [2, 3, 4], [0, 1, 2], [4, 5, 0], [2, 3, 4]])
    out["duplicate_faces"] = np.concatenate([v[f], rng.random((5, 3, 3)).astype(np.float32)], axis=2)
    # degenerate triangles (two corners weld together) and fully collapsed ones
    v = rng.random((5, 3)).astype(np.float32)
    f = np.array([[0, 0, 1], [1, 2, 3], [4, 4, 4], [3, 2, 1]])
    out["degenerate"] = np.concatenate([v[f], rng.random((4, 3, 3)).astype(np.float32)], axis=2)
    # signed zeros: vertices on the coordinate planes, reached as -0.0 and +0.0
    z = np.array([[0.0, 0.5, 1.0], [-0.0, 0.5, 1.0], [1.0, -0.0, 0.0], [1.0, 0.0, -0.0], [2.0, 2.0, 2.0]], np.float32)
    f = np.array([[0, 2, 4], [1, 3, 4], [0, 3, 4], [1, 2, 4]])
    out["signed_zero"] = np.concatenate([z[f], rng.random((4, 3, 3)).astype(np.float32)], axis=2)
    # a few thousand triangles on a coarse lattice: heavy welding, many repeated and degenerate faces
    lattice = (rng.integers(-6, 7, (4000, 3, 3)) * np.float32(0.25)).astype(np.float32)
    out["lattice"] = np.concatenate([lattice, rng.random((4000, 3, 3)).astype(np.float32)], axis=2)
    return out


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("eps", [0.0, 0.004])
def test_weld_matches_reference_process_triangles(eps):
    params = dict(synth.REPLICA_PARAMS)
    g = GeoWrapper(**params, num_sdf_blocks=2000, hash_num_buckets=1000, max_num_triangles=100000)
    ref = RefCuda(params, 2000, 1000, max_num_triangles=100000)
    for name, soup in _hand_soups().items():
        if name == "signed_zero" and eps == 0.0:
            continue  # own test below
        gv, gf, gc = _weld_ours(g, soup, eps)
        rv, rf, rc = ref.process_soup(soup, eps)
        assert gv.shape == rv.shape and gf.shape == rf.shape, (name, gv.shape, rv.shape, gf.shape, rf.shape)
        assert np.array_equal(gv, rv) and np.array_equal(gf, rf) and np.array_equal(gc, rc), name


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_weld_of_a_real_soup_matches_reference_process_triangles():
    """The 7-frame room: ~75 k triangles through the reference's processTriangles and through the device weld."""
    for eps in (0.0, 0.004):
        params = dict(synth.REPLICA_PARAMS)
        params["vertices_merging_threshold"] = eps
        fx, fy, cx, cy = synth.intrinsics(640, 480)
        g = GeoWrapper(**params, num_sdf_blocks=NUM_BLOCKS, hash_num_buckets=NUM_BUCKETS, max_num_triangles=MAX_TRIS)
        g.setCamera(fx, fy, cx, cy, 480, 640, params["min_depth"], params["max_depth"], 0)
        for k in range(7):
            t, q, depth, rgb = synth.rgbd_frame(k, n_frames=2000)
            feed(g, [], t, q, depth, rgb)
        g.extractMesh(None)
        tris = g.getTriangles()
        ref = RefCuda(params, 2000, 1000, max_num_triangles=len(tris) + 16)
        rv, rf, rc = ref.process_soup(tris, eps)
        gv, gf, gc = g.getVertices(), g.getFaces(), g.getColors()
        print(f"[weld vs reference host code, eps={eps}] {len(tris)} triangles -> {len(gv)} vertices / {len(gf)} faces")
        assert np.array_equal(gv, rv) and np.array_equal(gf, rf) and np.array_equal(gc, rc)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_weld_signed_zero_and_nan_follow_the_reference_map():
    """The reference keys its exact map on the BYTES of a vertex (Vector3dHash, mesh_extractor.cuh:25-30)
    and compares with a == b (:32-36): -0.0 and +0.0 meet only if their hashes share a bucket, a NaN
    vertex never equals anything. Ours: bit-pattern keys, NaN vertices each their own class. The vertex
    COUNT is compared (the reference's can be lower by the signed-zero pairs that collide in a bucket)."""
    params = dict(synth.REPLICA_PARAMS)
    g = GeoWrapper(**params, num_sdf_blocks=2000, hash_num_buckets=1000, max_num_triangles=1000)
    ref = RefCuda(params, 2000, 1000, max_num_triangles=1000)
    soup = _hand_soups()["signed_zero"]
    gv, gf, _ = _weld_ours(g, soup, 0.0)
    rv, rf, _ = ref.process_soup(soup, 0.0)
    assert len(gv) == 5  # bit patterns: (+0, .5, 1), (-0, .5, 1), (1, -0, 0), (1, 0, -0), (2, 2, 2)
    assert len(rv) in (3, 4, 5), len(rv)  # 5 unless a +-0 pair lands in one bucket of libstdc++'s table
    if len(rv) == 5:
        assert np.array_equal(gv, rv) and np.array_equal(gf, rf)
    nan = soup.copy()
    nan[0, 0, 0] = np.nan
    nan[1, 0, 0] = np.nan  # two vertices with identical NaN payloads
    gv, gf, _ = _weld_ours(g, nan, 0.0)
    rv, rf, _ = ref.process_soup(nan, 0.0)
    assert len(gv) == len(rv) and np.array_equal(gf, rf)
    assert np.array_equal(np.isnan(gv), np.isnan(rv))


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_reference_mesh_utilities_known_answers():
    """The reference's own gtest cases (tests/test_marching_cubes.cpp:91-215), run against the reference's
    own code in oracle/_ref: the harness and the Eigen stand-in it is compiled against behave."""
    params = dict(synth.REPLICA_PARAMS)
    ref = RefCuda(params, 2000, 1000, max_num_triangles=16)
    faces = np.array([[0, 1, 2], [2, 3, 4], [0, 1, 2], [4, 5, 6], [2, 3, 4]], np.int32)
    assert len(ref.remove_duplicate_faces(faces)) == len(faces) - 2  # :91-101
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 0]], np.float64)
    f = np.array([[0, 1, 2], [3, 1, 2]], np.int32)
    uv, uf, mp = ref.remove_duplicate_vertices(v, f, 0.0)  # :126-169
    assert len(uv) == 3 and np.array_equal(uv, v[:3]) and list(mp) == [0, 1, 2, 0] and np.array_equal(uf, [[0, 1, 2], [0, 1, 2]])
    v[3] = [0.0005, 0.0005, 0.0005]
    uv, uf, mp = ref.remove_duplicate_vertices(v, f, 0.001)  # :171-214
    assert len(uv) == 3 and list(mp) == [0, 1, 2, 0]


# ---------------------------------------------------------------------------------------------
# a15 pinned against the reference's OWN streamer: streamer.cpp / streamer.cu are compiled,
# unmodified, into oracle/_ref and driven the way GeoWrapper drives them.
# ---------------------------------------------------------------------------------------------
def _ply_payload_sorted(path):
    pts, props = _read_ply_points(path)
    raw = np.frombuffer(pts.tobytes(), np.uint8).reshape(len(pts), -1)
    order = np.lexsort(raw.T[::-1])
    return props, raw[order]


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_serialize_data_matches_reference_streamer(tmp_path):
    """streamAllOut + serializeData (streamer.cpp:104-160, 216-281): the voxel cloud and the per-block
    cloud (mean colour, mean weight) written by the reference's Streamer and by mrhash_b200 hold the same
    records, byte for byte, once both files are sorted (the reference walks an unordered_map of chunks)."""
    ours, ref, params = build_state(6)
    ref.streamer_create()
    assert compare_dumps(ours.dumpState(), ref.dump())["ok"]
    ours.streamAllOut()
    ref.stream_all_out()
    assert ref.grid_blocks() == ours.storeSize()
    oh, ov = str(tmp_path / "o_hash.ply"), str(tmp_path / "o_voxel.ply")
    rh, rv = str(tmp_path / "r_hash.ply"), str(tmp_path / "r_voxel.ply")
    ours.serializeData(oh, ov)
    ref.serialize_data(rh, rv)
    for a, b, what in ((ov, rv, "voxel"), (oh, rh, "hash")):
        pa, da = _ply_payload_sorted(a)
        pb, db = _ply_payload_sorted(b)
        assert pa == pb, (what, pa, pb)
        assert da.shape == db.shape, (what, da.shape, db.shape)
        assert np.array_equal(da, db), (what, int((da != db).any(axis=1).sum()))
    print(f"[serializeData vs reference] {len(_read_ply_points(ov)[0])} voxel points, {len(_read_ply_points(oh)[0])} block points identical")
