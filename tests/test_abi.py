"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol the header
declares (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "df3d_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(df3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(lib_built)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"libdf3d_b200.so does not export: {missing}"


def test_python_binding_covers_header(lib_built):
    from deepfly3d_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib.df3d_abi_version() == 2


def test_argument_validation_without_gpu(lib_built):
    """Entry points reject bad arguments before touching CUDA and report through df3d_last_error."""
    from deepfly3d_b200 import _lib

    lib = _lib.lib
    rc = lib.df3d_triangulate_dlt(None, None, 7, 1, 38, None, None)
    assert rc == -1 and b"null pointer" in lib.df3d_last_error()
    order = (ctypes.c_int * 7)(0, 1, 2, 3, 4, 5, 5)
    rc = lib.df3d_pack_points2d(ctypes.c_void_p(8), 7, 1, 19, 64, 128, order, 960, 480, ctypes.c_void_p(8), None, None)
    assert rc == -1 and b"permutation" in lib.df3d_last_error()
    assert lib.df3d_pack_points2d(ctypes.c_void_p(8), 6, 1, 19, 64, 128, order, 960, 480, ctypes.c_void_p(8), None, None) == -4
    d = _lib.HGDesc(8, 19, 256, 256, 4)
    assert lib.df3d_hg_param_count(ctypes.byref(d)) == 25566232
    d_bad = _lib.HGDesc(8, 19, 250, 256, 4)
    assert lib.df3d_hg_param_count(ctypes.byref(d_bad)) == 0
    assert lib.df3d_bundle_adjust_workspace_bytes(7, 15, 38) > 0
    assert lib.df3d_bundle_adjust_launches(None) == 2 + 8 * 20          # default solver: LSMR (per evaluation: 3 passes + their finishes, one cooperative LSMR kernel, apply)


def test_sharded_bundle_adjust_plan_is_a_host_function(lib_built):
    """df3d_ba_sharded_plan needs no GPU: block count, where the per-block partial sums live in the workspace, doubles per
    block of the four passes; block counts that do not split evenly over the ranks are refused."""
    from deepfly3d_b200 import _lib, ops

    nb, off, pd = ops.ba_sharded_plan(7, 8000, 38, 8)
    assert nb == 4 * 148 and nb % 8 == 0                      # 304 000 points / 128 per block, capped
    assert off % 256 == 0 and 0 < off < _lib.lib.df3d_bundle_adjust_workspace_bytes(7, 8000, 38)
    assert pd == [36 * 7 + 12 * 7 + 6, 36 * 7 + 6 * 7 + 36 * 49 + 6 * 7 + 2, 4, 3]
    assert off + nb * max(pd) * 8 <= _lib.lib.df3d_bundle_adjust_workspace_bytes(7, 8000, 38)
    assert ops.ba_sharded_plan(7, 26, 38, 2)[0] == 8          # 988 points -> 8 blocks
    assert ops.ba_sharded_plan(7, 26, 38, 3) is None          # 8 blocks over 3 ranks
    assert ops.ba_sharded_plan(7, 3, 38, 4) is None           # 1 block over 4 ranks


def test_flatten_matches_param_count(lib_built):
    from deepfly3d_b200 import _lib, hourglass
    from oracle import hourglass as ohg

    for stacks in (2, 8):
        blob, S, K = hourglass.flatten_state_dict(ohg.make_model(stacks).state_dict())
        d = _lib.HGDesc(S, K, 256, 256, 1)
        assert (S, K) == (stacks, 19)
        assert blob.size == _lib.lib.df3d_hg_param_count(ctypes.byref(d))
