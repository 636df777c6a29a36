"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, exports every symbol that
include/mrhash_b200.h declares, and fails LOUDLY (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT, has_gpu

from mrhash_b200 import _capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mrhash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrh_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = C.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes binding covers the whole header too
    assert sorted(_capi.SIGNATURES) == names


def test_library_is_built_for_sm_100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    assert all("sm_100a" in l for l in out.splitlines() if l.strip().startswith("ELF file"))


def test_abi_version_and_defaults(built_lib):
    lib = _capi.lib()
    assert lib.mrh_abi_version() == 1
    p = _capi.Params()
    assert lib.mrh_params_default(C.byref(p)) == 0
    # configurations/replica.cfg:1-18
    assert abs(p.sdf_truncation - 0.07) < 1e-7 and p.integration_weight_sample == 1 and abs(p.virtual_voxel_size - 0.01) < 1e-7
    assert p.n_frames_invalidate_voxels == 100 and p.min_weight_threshold == 5 and p.projective_sdf == 1


def test_product_does_not_touch_the_oracle():
    """The product (mrhash_b200/, include/) must not link, load or mention anything under oracle/."""
    bad = []
    for base in ("mrhash_b200", "include", "mrhash"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"oracle_lib|libmrh_oracle|libref_harness|oracle/", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu(built_lib):
    from mrhash_b200 import GeoWrapper, synth

    with pytest.raises(RuntimeError, match="no CUDA device|CPU path"):
        GeoWrapper(**synth.REPLICA_PARAMS, num_sdf_blocks=100, hash_num_buckets=100)
