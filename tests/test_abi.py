"""The C-ABI library loads and exports exactly the symbols include/svs.h declares (no compute calls)."""
import os
import re

import svolsdf_b200._lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'svs.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(svs_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    syms = header_symbols()
    assert 'svs_sdf_outputs_backward' in syms and 'svs_composite_forward' in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    lib = L.load()
    for name in header_symbols():
        assert hasattr(lib, name), 'missing export: ' + name


def test_ctypes_table_matches_header():
    assert sorted(L.SIGNATURES.keys()) == header_symbols()


def test_abi_version_and_engine_query():
    lib = L.load()
    assert lib.svs_abi_version() == 6
    assert lib.svs_has_engine(L.ENGINE_FP32) == 1


def test_descriptor_validation_needs_no_gpu():
    lib = L.load()
    d = L.make_desc(L.NET_SDF, [39, 256, 256, 256, 256, 256, 256, 256, 256],
                    [256, 256, 256, 217, 256, 256, 256, 256, 257], d_in=3, n_freqs=6, skip_layer=4)
    n = lib.svs_mlp_wbuf_floats(d, L.ENGINE_FP32)
    assert n == sum(o * ((i + 3) // 4 * 4) + (o + 3) // 4 * 4 for i, o in
                    zip([39, 256, 256, 256, 256, 256, 256, 256, 256], [256, 256, 256, 217, 256, 256, 256, 256, 257]))
    assert lib.svs_sdf_ldy(d) == 260
    # the tcgen05 engine appends the bf16 weight images (forward + transposed) to the fp32 block
    nb = lib.svs_mlp_wbuf_floats(d, L.ENGINE_TC)
    assert nb > n and lib.svs_has_engine(L.ENGINE_TC) == 1
    assert lib.svs_sdf_saved_floats(d, 1000, L.ENGINE_TC) == 8 * (1 + 8 * 4 + 8 * 4) * 16384 // 4
    bad = L.make_desc(L.NET_SDF, [39, 256], [256, 257], d_in=3, n_freqs=5)
    assert lib.svs_mlp_wbuf_floats(bad, L.ENGINE_FP32) == -1
    assert b'PE width' in lib.svs_last_error()


def test_no_cpu_fallback():
    import pytest
    import torch
    with pytest.raises(L.SvsError):
        L.ptr(torch.zeros(4))


def test_argument_validation_and_empty_inputs_need_no_gpu():
    """error behaviour of the C ABI: bad arguments come back as SVS_ERR_INVALID (-1) with a message, empty inputs are a
    no-op that returns 0 — both before anything touches the device"""
    import ctypes as C
    lib = L.load()
    p = C.c_void_p(4096)            # never dereferenced: validation and the empty-input return come first
    # compositor: empty ray set ok, S out of range / missing beta rejected
    assert lib.svs_composite_forward(p, p, None, None, p, 1e-4, None, None, 0, 98, 0, p, None, None, None, None, None) == 0
    assert lib.svs_composite_forward(p, p, None, None, p, 1e-4, None, None, 8, 300, 0, p, None, None, None, None, None) < 0
    assert b'S <= 256' in lib.svs_last_error()
    assert lib.svs_composite_forward(p, p, None, None, None, 1e-4, None, None, 8, 98, 0, p, None, None, None, None, None) < 0
    assert lib.svs_composite_backward(p, p, None, p, 1e-4, None, None, 0, 98, 0, None, None, None, None, p, None, None, None) == 0
    # MVS cost lookup: view count, null volumes, image size
    v = (L.MvsView * 9)()
    for i in range(9):
        v[i].cost, v[i].z_near, v[i].z_far, v[i].Dz, v[i].H, v[i].W = 4096, 4096, 4096, 4, 6, 8
    assert lib.svs_cost_mapping(p, 0, 20, v, 3, 72, 96, 1, None, p, p, p, None) == 0          # no samples
    assert lib.svs_cost_mapping(p, 4, 20, v, 9, 72, 96, 1, None, p, p, p, None) < 0           # > 8 views
    assert b'n_views' in lib.svs_last_error()
    assert lib.svs_cost_mapping(p, 4, 20, v, 0, 72, 96, 1, None, p, p, p, None) < 0
    assert lib.svs_cost_mapping(p, 4, 20, v, 3, 1, 96, 1, None, p, p, p, None) < 0            # degenerate image
    # ... fused with the loss's MVS term
    f = C.c_float
    assert lib.svs_mvs_loss(p, 0, 20, v, 3, 72, 96, 1, None, p, f(0.5), f(0.0), p, p, p, None) == 0     # no rays
    assert lib.svs_mvs_loss(p, 4, 300, v, 3, 72, 96, 1, None, p, f(0.5), f(0.0), p, p, p, None) < 0    # D > 256
    assert lib.svs_mvs_loss(p, 4, 20, v, 3, 72, 96, 1, None, None, f(0.5), f(0.0), p, p, p, None) < 0  # no weights
    assert lib.svs_mvs_loss(p, 4, 20, v, 3, 72, 96, 1, None, p, f(-1.0), f(0.0), p, p, p, None) < 0    # negative exponent
    v[1].cost = None
    assert lib.svs_cost_mapping(p, 4, 20, v, 3, 72, 96, 1, None, p, p, p, None) < 0
    assert b'view 1' in lib.svs_last_error()
    assert lib.svs_cost_mapping(None, 4, 20, v, 3, 72, 96, 1, None, p, p, p, None) < 0


def test_dy_staging_index_arithmetic_is_exact():
    """The backward chain converts a tile's fp32 dy rows from a flat staging buffer: row of flat element e = umulhi(e,
    ceil(2^32 / ldy)) (csrc/mlp_tc_chains.cuh: dy_magic; csrc/mlp_tc.cuh: PRO_DY).  Exact for every element of a 128-row
    tile at every row length the path accepts (128 * ldy < 65536), and the 16-byte bulk / direct-load split of a partial
    last tile covers every valid float exactly once."""
    import numpy as np
    for ldy in list(range(1, 64)) + [257, 258, 300, 320, 511]:
        if 128 * ldy >= 65536:
            continue
        magic = (1 << 32) // ldy + 1
        e = np.arange(128 * ldy, dtype=np.uint64)
        row = (e * np.uint64(magic)) >> np.uint64(32)
        assert np.array_equal(row, e // np.uint64(ldy)), ldy
    ldy = 257
    for rows in (1, 2, 3, 57, 127, 128):
        total = (rows * ldy * 4) & ~15           # bytes moved by bulk copies
        n_el = rows * ldy
        from_smem = sum(1 for e in range(n_el) if e * 4 + 4 <= total)
        direct = sum(1 for e in range(n_el) if not (e * 4 + 4 <= total))
        assert from_smem == total // 4 and from_smem + direct == n_el and direct <= 3
