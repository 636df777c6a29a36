"""Comparator of canonical state dumps (entries [n,5] = x,y,z,resolution,ptr sorted by key,
voxels [n,512] of the reference's 12-byte Voxel). The parity contract: identical key
sets and resolutions, bit-exact weight / rgb, sdf and sum_squared within a tolerance. Slot / pool
indices (ptr) are layout- and race-dependent and not compared."""
import numpy as np


def _keys(entries):
    e = entries.astype(np.int64)
    return (e[:, 0] + (1 << 20)) << 42 | (e[:, 1] + (1 << 20)) << 21 | (e[:, 2] + (1 << 20))


def compare_dumps(a, b, sdf_rtol=1e-5, sdf_atol=1e-7):
    """Returns a dict report; report['ok'] is the strict verdict."""
    ea, va = a
    eb, vb = b
    ka, kb = _keys(ea), _keys(eb)
    common, ia, ib = np.intersect1d(ka, kb, return_indices=True)
    rep = {
        "n_a": len(ka),
        "n_b": len(kb),
        "only_a": int(len(ka) - len(common)),
        "only_b": int(len(kb) - len(common)),
    }
    xa, xb = va[ia], vb[ib]
    res_a, res_b = ea[ia, 3], eb[ib, 3]
    rep["resolution_mismatch"] = int((res_a != res_b).sum())
    # only the first 64 voxels of a resolution-1 record are meaningful
    lane = np.arange(512)[None, :]
    valid = (res_a[:, None] == 0) | (lane < 64)
    rep["weight_mismatch"] = int(((xa["weight"] != xb["weight"]) & valid).sum())
    rgb_bad = ((xa["r"] != xb["r"]) | (xa["g"] != xb["g"]) | (xa["b"] != xb["b"])) & valid
    rep["rgb_mismatch"] = int(rgb_bad.sum())
    for f in ("sdf", "sum_squared"):
        # NaN (a NaN depth pixel poisons the voxels under it, in the reference too) must meet NaN; the
        # payload bits of a NaN are not part of the contract (GPU: canonical 0x7FFFFFFF, x86: propagated)
        na, nb = np.isnan(xa[f]), np.isnan(xb[f])
        both = na & nb
        with np.errstate(invalid="ignore"):
            d = np.abs(xa[f].astype(np.float64) - xb[f].astype(np.float64))
            tol = sdf_atol + sdf_rtol * np.maximum(np.abs(xa[f]), np.abs(xb[f]))
            bad = ((d > tol) | (na != nb)) & valid
        d = np.where(both | (na != nb), 0.0, d)
        rep[f + "_mismatch"] = int(bad.sum())
        rep[f + "_max_abs_diff"] = float(d[valid].max()) if valid.any() else 0.0
        rep[f + "_bitexact"] = bool(((xa[f].view(np.uint32) == xb[f].view(np.uint32)) | both)[valid].all() and not bad.any()) if valid.any() else True
        rep[f + "_nan"] = int((both & valid).sum())
    rep["voxels_compared"] = int(valid.sum())
    rep["ok"] = all(rep[k] == 0 for k in ("only_a", "only_b", "resolution_mismatch", "weight_mismatch", "rgb_mismatch", "sdf_mismatch", "sum_squared_mismatch"))
    return rep


def canonical_triangles(tris, decimals=5):
    """Triangle soup [T,3,6] -> sorted array of rounded vertex positions, rotation-invariant per triangle."""
    if len(tris) == 0:
        return np.zeros((0, 9))
    p = np.round(tris[:, :, :3].astype(np.float64), decimals)
    # rotate each triangle so that its lexicographically smallest vertex comes first (orientation kept)
    keys = p[:, :, 0] * 1e6 + p[:, :, 1] * 1e3 + p[:, :, 2]
    first = np.argmin(keys, axis=1)
    idx = (first[:, None] + np.arange(3)[None, :]) % 3
    p = np.take_along_axis(p, idx[:, :, None], axis=1).reshape(-1, 9)
    order = np.lexsort(p.T[::-1])
    return p[order]
