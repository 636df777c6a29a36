"""GPU: BASELINE configs[4] at its full size - extractMesh over >= 97 657 blocks (50 M voxels) - against
the reference's own marching-cubes kernels on the same map: same triangle count, and the same
multiset of bit-identical triangles (positions and colours), not only "within 1e-4"."""
import os
import sys

import pytest

from oracle_lib import ref_available

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_50m_voxel_mesh_is_the_reference_kernels_soup():
    import bench_mesh

    info = bench_mesh.run(97657, 0.003, with_ref=True, compare=True)
    print({k: info[k] for k in ("blocks", "voxels_M", "triangles", "vertices", "faces", "extract_mesh_total_s", "reference")})
    assert info["blocks"] >= 97657 and info["voxels_M"] >= 50.0
    ref = info["reference"]
    assert info["triangles"] == ref["ref_triangles"] > 4_000_000
    assert ref["bit_identical_triangles"] == info["triangles"]
    # the weld neither loses nor invents geometry: every face indexes a vertex, no degenerate face survives
    assert 0 < info["faces"] <= info["triangles"] and 0 < info["vertices"] <= 3 * info["triangles"]
