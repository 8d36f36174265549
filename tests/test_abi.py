"""The C-ABI library loads and exports every symbol include/b2second.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import REPO


def _header_symbols():
    text = open(os.path.join(REPO, "include", "b2second.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from spconv import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libb2second.so does not export %s" % s


def test_bindings_cover_header():
    from spconv import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_header_arity_matches_bindings():
    from spconv import _lib
    text = open(os.path.join(REPO, "include", "b2second.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else params.count(",") + 1
        assert n == len(args), "%s: header has %d parameters, binding %d" % (name, n, len(args))


def test_version_and_error_string():
    from spconv import _lib
    lib = _lib.load()
    assert lib.b2s_version() >= 100
    assert isinstance(lib.b2s_last_error(), bytes)


def test_workspace_queries_are_host_only():
    from spconv import _lib
    lib = _lib.load()
    assert lib.b2s_voxelize_hash_capacity(20000) == 65536
    assert lib.b2s_voxelize_workspace_bytes(20000, 1, 40000, 5) > 0
    assert lib.b2s_rulebook_conv_workspace_bytes(1, (ctypes.c_int * 3)(21, 800, 704)) > 0
    assert lib.b2s_nms_workspace_bytes(1, 4096, 1000) > 0


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    import spconv
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = spconv.SparseConvTensor(torch.zeros(3, 4), torch.zeros(3, 4, dtype=torch.int32), [8, 8, 8], 1)
    with pytest.raises(RuntimeError):
        spconv.SubMConv3d(4, 16, 3, bias=False)(x)
    with pytest.raises(RuntimeError):
        x.dense()
    with pytest.raises(RuntimeError):
        spconv.utils.VoxelGeneratorV2([0.1] * 3, [0, 0, 0, 1, 1, 1], 5).generate(torch.zeros(4, 4).numpy(), 100)


def test_argument_checks_run_before_any_cuda_call():
    """host-side contract of the sparse-conv entry points: bad arguments are rejected with -2 and a message before any
    CUDA call (so this runs without a GPU); the supported (Cin, Cout) table is a host query."""
    from spconv import _lib
    lib = _lib.load()
    i3 = (ctypes.c_int * 3)
    dummy = ctypes.c_void_p(256)                      # never dereferenced: every call below fails its argument check
    # tile plan: K must be 1..27, ksize must multiply to K, tile_mask (and perm when sorting) are required
    assert lib.b2s_sparse_tile_plan(dummy, None, 28, None, dummy, 1000, 0, None, dummy, None) == -2
    assert b"K must be 1..27" in lib.b2s_last_error()
    assert lib.b2s_sparse_tile_plan(dummy, None, 27, i3(3, 3, 2), dummy, 1000, 0, None, dummy, None) == -2
    assert b"ksize" in lib.b2s_last_error()
    assert lib.b2s_sparse_tile_plan(dummy, None, 27, i3(3, 3, 3), dummy, 1000, 1, None, dummy, None) == -2
    assert lib.b2s_sparse_tile_plan(dummy, None, 27, i3(3, 3, 3), dummy, 1000, 0, None, None, None) == -2
    assert lib.b2s_sparse_tile_plan(None, None, 27, i3(3, 3, 3), dummy, 0, 0, None, dummy, None) == 0    # no rows: nothing to do
    # tensor-pipe conv: channel table, K range, 16-byte row alignment, 32-bit row offsets
    assert [lib.b2s_sparse_conv_tc_supported(c, 64) for c in (8, 16, 32, 64, 4, 128)] == [1, 1, 1, 1, 0, 0]
    assert [lib.b2s_sparse_conv_tc_supported(64, c) for c in (16, 32, 64, 8, 128)] == [1, 1, 1, 0, 0]

    def conv(cin=64, cout=64, K=27, in_stride=128, rows_in=1000, feat=256, out_stride=128):
        return lib.b2s_sparse_conv_tc_plan(ctypes.c_void_p(feat), ctypes.c_void_p(feat + 128), in_stride, rows_in, cin, dummy,
                                           dummy, dummy, K, dummy, 1000, None, None, None, None, 1, dummy,
                                           ctypes.c_void_p(512), out_stride, cout, None, None)
    assert conv(cin=48) == -2 and b"built for Cin" in lib.b2s_last_error()
    assert conv(K=0) == -2 and conv(K=28) == -2
    assert conv(in_stride=60) == -2 and b"16-byte aligned" in lib.b2s_last_error()
    assert conv(feat=260) == -2
    assert conv(out_stride=60) == -2
    assert conv(rows_in=1 << 25, in_stride=128) == -2 and b"4 GiB" in lib.b2s_last_error()
