"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/*.h declares, geometry derivation matches the oracle (and therefore tron.cu:905-961),
RA files and binary16 conversions match the oracle / known answers.  No compute calls."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    import tron_b200 as t
    for sym in t.EXPORTED_SYMBOLS:
        assert hasattr(lib, sym), sym
    # every function name declared in include/*.h must be in the list and in the .so
    declared = set()
    for h in ("tron.h", "ra.h", "float16.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        if h == "float16.h":                       # C++-linkage block has mangled names; check extern "C" part
            text = text[text.index('extern "C"'):]
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
            name = m.group(1)
            if name in ("defined", "sizeof", "if") or name.isupper():
                continue
            declared.add(name)
    missing = sorted(n for n in declared if not hasattr(lib, n))
    assert not missing, missing
    assert lib.tron_version() == 100


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    import tron_b200 as t
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(t.TronError, match="no CUDA device"):
        t.Plan(t.make_config([1, 1, 64, 64, 1], adjoint=True))


def test_product_does_not_link_or_import_the_oracle():
    import tron_b200 as t
    out = subprocess.run(["nm", "-D", t.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tron_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "build.py", os.path.join(dirpath, f)


GEOM_CASES = [
    ([6, 1, 512, 20271, 1], dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=21)),     # BASELINE cfg2
    ([32, 1, 512, 205824, 1], dict(adjoint=True, golden=True, undersamp=1.5703125, prof_slide=804)),  # cfg3
    ([16, 1, 256, 42107, 1], dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=21)),    # cfg4
    ([16, 1, 256, 42000, 1], dict(adjoint=True, golden=True, undersamp=0.08203125, prof_slide=21)),
    ([64, 1, 1024, 1024, 1], dict(adjoint=False, gridos=2.0, kernwidth=6.0)),                   # cfg5 forward
    ([64, 1, 2048, 2048, 1], dict(adjoint=True, gridos=2.0, kernwidth=6.0)),                    # cfg5 adjoint
    ([1, 1, 256, 256, 1], dict(adjoint=False)),                                                 # cfg1 forward
    ([1, 1, 512, 512, 1], dict(adjoint=True)),                                                  # cfg1 adjoint
    ([6, 1, 512, 512, 1], dict(adjoint=True, golden=True, gridos=1.5)),
    ([2, 1, 128, 300, 1], dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=7, skip_angles=3)),
    ([2, 1, 100, 77, 1], dict(adjoint=True, gridos=1.28)),
    ([4, 1, 96, 96, 1], dict(adjoint=False, undersamp=0.37, gridos=1.25)),
]


@pytest.mark.parametrize("dims,flags", GEOM_CASES)
def test_geometry_matches_oracle(lib, oracle, dims, flags):
    import tron_b200 as t
    g = t.geometry(t.make_config(dims, **flags))
    o = oracle.config(dims, flags.get("adjoint", False), golden=flags.get("golden", False),
                      gridos=flags.get("gridos", 2.0), kernwidth=flags.get("kernwidth", 2.0),
                      undersamp=flags.get("undersamp", 1.0), prof_slide=flags.get("prof_slide", 0),
                      skip_angles=flags.get("skip_angles", 0))
    for k in ("nc", "nt", "nro", "npe1", "npe2", "npe1work", "nx", "ny", "nz", "nxos", "nyos"):
        assert getattr(g, k) == getattr(o, k), k
    assert [int(x) for x in g.out_dims] == [int(x) for x in o.out_dims]
    assert int(g.out_elems) == int(o.out_elems)


def test_geometry_known_answers(lib):
    import tron_b200 as t
    g = t.geometry(t.make_config(*GEOM_CASES[0][:1], **GEOM_CASES[0][1]))
    assert (g.npe1work, g.nz, g.nx, g.nxos) == (204, 956, 256, 512)
    g = t.geometry(t.make_config(*GEOM_CASES[1][:1], **GEOM_CASES[1][1]))
    assert (g.npe1work, g.nz) == (804, 256)
    g = t.geometry(t.make_config(*GEOM_CASES[2][:1], **GEOM_CASES[2][1]))
    assert (g.npe1work, g.nz, g.nx) == (128, 2000, 128)
    g = t.geometry(t.make_config(*GEOM_CASES[3][:1], **GEOM_CASES[3][1]))
    assert (g.npe1work, g.nz) == (21, 2000)


def test_geometry_rejections(lib):
    import tron_b200 as t
    bad = [
        (dict(dims=[3, 1, 64, 64, 1], adjoint=True), "even"),                 # tron.cu:963
        (dict(dims=[2, 2, 64, 64, 1], adjoint=True), "nt"),
        (dict(dims=[1, 1, 64, 64, 2], adjoint=False), "dims\\[4\\]"),
        (dict(dims=[1, 1, 64, 64, 1], adjoint=False, niter=3), "CGNR"),       # tron.cu:753-755: adjoint only
        (dict(dims=[2, 1, 64, 64, 1], adjoint=True, niter=3, gridos=1.5), "CGNR"),   # needs nro == nxos
        (dict(dims=[4, 1, 64, 64, 1], adjoint=True, niter=2, coils=(0, 2)), "coil shards"),
        (dict(dims=[4, 1, 64, 64, 1], adjoint=True, coils=(0, 2)), "sos_partial"),             # RSS of a partial sum
        (dict(dims=[4, 1, 64, 64, 1], adjoint=True, coils=(0, 2), per_coil_out=True, sos_partial=True), "sos_partial"),
        (dict(dims=[4, 1, 64, 64, 1], adjoint=False, coils=(2, 4)), "dims\\[0\\]"),                 # forward: shard the input
        (dict(dims=[4, 1, 64, 64, 1], adjoint=True, coil_combine=1, per_coil_out=True), "Walsh"),
        (dict(dims=[4, 1, 64, 64, 1], adjoint=True, coil_combine=2), "coil_combine"),
        (dict(dims=[1, 1, 64, 64, 1], adjoint=True, koosh=True), "koosh"),
    ]
    for kw, pat in bad:
        with pytest.raises(t.TronError, match=pat):
            t.geometry(t.make_config(**kw))
    with pytest.raises(t.TronError):
        t.geometry(t.make_config([1, 1, 50, 64, 1], adjoint=True, gridos=1.0))   # nxos = 25, odd
    with pytest.raises(t.TronError, match="slice shard"):
        t.geometry(t.make_config([2, 1, 64, 640, 1], adjoint=True, prof_slide=64, slices=(3, 99)))
    with pytest.raises(t.TronError, match="coil"):
        t.geometry(t.make_config([6, 1, 64, 64, 1], adjoint=True, coils=(1, 3)))


def test_slice_shards_cover_the_range(lib):
    import tron_b200 as t
    dims, flags = GEOM_CASES[0]
    full = t.geometry(t.make_config(dims, **flags))
    spoke = full.nc * full.nro
    for world in (1, 2, 3, 4, 8):
        prev_hi, out_sum = 0, 0
        for rank in range(world):
            lo, hi = t.shard_slices(full.nz, rank, world)
            assert lo == prev_hi
            prev_hi = hi
            g = t.geometry(t.make_config(dims, slices=(lo, hi), **flags))
            assert int(g.shard_in_offset) == spoke * lo * 21
            assert int(g.shard_in_elems) == spoke * ((hi - lo - 1) * 21 + 204)       # window + halo
            assert int(g.shard_in_offset + g.shard_in_elems) <= int(full.in_elems)
            assert int(g.shard_out_offset) == lo * 256 * 256
            out_sum += int(g.shard_out_elems)
        assert prev_hi == full.nz and out_sum == int(full.out_elems)


# ------------------------------------------------------------------------- RA container
def test_ra_header_layout_and_roundtrip(lib, tmp_path):
    import tron_b200 as t
    a = (np.arange(2 * 1 * 4 * 3 * 1) * (1 + 0.5j)).astype(np.complex64)
    f = str(tmp_path / "x.ra")
    t.ra_write(f, a, dims=[2, 1, 4, 3, 1])
    raw = open(f, "rb").read()
    assert len(raw) == 88 + a.nbytes                               # 48 + 8*5 header bytes for 5-D
    head = struct.unpack("<11Q", raw[:88])
    assert head[0] == 0x7961727261776172 == 8746397786917265778    # ra.h:51, rawrite.m:47
    assert raw[:8] == b"rawarray"
    assert head[1:6] == (0, 4, 8, a.nbytes, 5)                     # flags, eltype, elbyte, size, ndims
    assert head[6:] == (2, 1, 4, 3, 1)
    b, dims, et, eb = t.ra_read(f)
    assert dims == [2, 1, 4, 3, 1] and (et, eb) == (4, 8) and np.array_equal(b, a)
    # sizes of the two data files of the reference (git-LFS pointers): SURVEY F1
    assert 88 + 256 * 256 * 8 == 524376
    assert 88 + 6 * 512 * 20271 * 8 == 498180184


def test_ra_payload_beyond_2_gib(lib, tmp_path):
    """A single read(2)/write(2) moves at most 0x7ffff000 bytes on Linux; the reference's loop loses the tail of larger
    payloads (ra.cu:153-158).  2 GiB + 4 KiB through ra_write / ra_read: every chunk boundary and the tail."""
    import shutil
    import tron_b200 as t
    if shutil.disk_usage(str(tmp_path)).free < 6 * (1 << 30):
        pytest.skip("needs 6 GiB of scratch space")
    n = (1 << 29) + 1024                                   # float32 elements: 2 GiB + 4 KiB
    a = np.empty(n, dtype=np.float32)
    a.view(np.uint32)[:] = np.arange(n, dtype=np.uint32)    # position-coded bit patterns
    f = str(tmp_path / "big.ra")
    t.ra_write(f, a, dims=[n], eltype=3)                   # ra.h:63-72: 3 = float
    assert os.path.getsize(f) == 48 + 8 + 4 * n
    b, dims, et, eb = t.ra_read(f)
    assert dims == [n] and (et, eb) == (3, 4) and b.shape == a.shape
    bu = b.view(np.uint32)
    assert np.array_equal(bu[-4096:], np.arange(n - 4096, n, dtype=np.uint32))           # the tail
    for edge in (0x7ffff000 // 4, (1 << 30) // 4, (1 << 31) // 4):                         # syscall cap, I/O chunks
        assert np.array_equal(bu[edge - 8:edge + 8], np.arange(edge - 8, edge + 8, dtype=np.uint32))
    assert np.array_equal(bu[::4099], np.arange(0, n, 4099, dtype=np.uint32))
    os.remove(f)


def test_ra_matches_oracle_writer(lib, oracle, tmp_path):
    """Bytes written by the product are bytes the oracle's reader understands, and vice versa."""
    import tron_b200 as t
    from oracle.oracle import _ptr
    a = np.arange(30, dtype=np.float32)
    f1 = str(tmp_path / "a.ra")
    t.ra_write(f1, a, dims=[5, 6], eltype=3)

    class ORa(C.Structure):
        _fields_ = [("flags", C.c_uint64), ("eltype", C.c_uint64), ("elbyte", C.c_uint64), ("size", C.c_uint64),
                    ("ndims", C.c_uint64), ("dims", C.POINTER(C.c_uint64)), ("data", C.POINTER(C.c_uint8))]
    r = ORa()
    assert oracle.lib.oracle_ra_read(C.byref(r), f1.encode()) == 0
    assert (r.eltype, r.elbyte, r.size, r.ndims) == (3, 4, 120, 2) and [r.dims[0], r.dims[1]] == [5, 6]
    f2 = str(tmp_path / "b.ra")
    assert oracle.lib.oracle_ra_write(C.byref(r), f2.encode()) == 0
    assert open(f1, "rb").read() == open(f2, "rb").read()
    oracle.lib.oracle_ra_free(C.byref(r))


def test_ra_errors_and_extras(lib, tmp_path):
    import tron_b200 as t
    bad = str(tmp_path / "bad.ra")
    open(bad, "wb").write(b"notarawarrayfile" * 8)
    with pytest.raises(t.TronError):
        t.ra_read(bad)
    with pytest.raises(t.TronError):
        t.ra_read(str(tmp_path / "missing.ra"))
    trunc = str(tmp_path / "trunc.ra")
    t.ra_write(trunc, np.zeros(16, dtype=np.complex64), dims=[16])
    data = open(trunc, "rb").read()
    open(trunc, "wb").write(data[:-8])
    with pytest.raises(t.TronError):
        t.ra_read(trunc)
    # reshape / squash / diff / convert
    f = str(tmp_path / "c.ra")
    a = (np.arange(12) - 3.25j).astype(np.complex64)
    t.ra_write(f, a, dims=[1, 3, 1, 4])
    r1, r2 = t.api.RaStruct(), t.api.RaStruct()
    assert lib.ra_read(C.byref(r1), f.encode()) == 0 and lib.ra_read(C.byref(r2), f.encode()) == 0
    assert lib.ra_diff(C.byref(r1), C.byref(r2)) == 0
    assert lib.ra_squash(C.byref(r1)) == 0 and r1.ndims == 2 and [r1.dims[0], r1.dims[1]] == [3, 4]
    assert lib.ra_diff(C.byref(r1), C.byref(r2)) == 1
    nd = (C.c_uint64 * 3)(2, 3, 2)
    assert lib.ra_reshape(C.byref(r1), nd, 3) == 0 and r1.ndims == 3
    assert lib.ra_reshape(C.byref(r1), nd, 2) != 0
    lib.ra_convert(C.byref(r2), 4, 4)                         # complex64 -> complex-half
    assert (r2.elbyte, r2.size) == (4, 48)
    half = np.ctypeslib.as_array(r2.data, shape=(48,)).view(np.float16)
    assert np.array_equal(half, a.view(np.float32).astype(np.float16))
    lib.ra_convert(C.byref(r2), 4, 8)
    back = np.ctypeslib.as_array(r2.data, shape=(96,)).view(np.complex64)
    assert np.array_equal(back, a)                            # these values are exact in half
    lib.ra_free(C.byref(r1)); lib.ra_free(C.byref(r2))


# ------------------------------------------------------------------------- binary16
def test_half_to_float_exhaustive(lib, oracle):
    allh = np.arange(65536, dtype=np.uint16)
    mine = np.array([lib.tron_halfbits_to_floatbits(int(h)) for h in allh], dtype=np.uint32)
    assert np.array_equal(mine, oracle.half_to_float_bits(allh))
    nan = np.isnan(allh.view(np.float16))
    assert np.array_equal(mine[~nan], allh.view(np.float16).astype(np.float32).view(np.uint32)[~nan])
    mine64 = np.array([lib.tron_halfbits_to_doublebits(int(h)) for h in allh[~nan]], dtype=np.uint64)
    assert np.array_equal(mine64, allh[~nan].view(np.float16).astype(np.float64).view(np.uint64))


def test_float_to_half_matches_oracle_and_ieee(lib, oracle):
    rng = np.random.default_rng(5)
    bits = np.concatenate([
        rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32),
        np.array([0, 0x80000000, 0x7f800000, 0xff800000, 0x7fc00000, 0x7f800001, 0x477fe000, 0x477ff000,
                  0x33000000, 0x33000001, 0x32ffffff, 0x387fc000, 0x387fe000, 0x38800000, 0x3f800000,
                  0x3f801000, 0x3f803000, 0x3f801001], dtype=np.uint32),
        (rng.integers(0x33000000, 0x38800000, 50000, dtype=np.uint64)).astype(np.uint32)])      # subnormal halves
    src = bits.view(np.float32)
    dst = np.zeros(bits.size, dtype=np.uint16)
    lib.tron_float_to_half_array(dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), bits.size)
    want = oracle.float_to_half_bits(bits)
    assert np.array_equal(dst, want)
    # against IEEE round-to-nearest-even (numpy): identical except the reference's
    # subnormal sticky-bit quirk and NaN payloads
    with np.errstate(over="ignore"):
        ieee = src.astype(np.float16).view(np.uint16)
    diff = dst != ieee
    f = np.abs(src[diff])
    finite = ~np.isnan(f)
    assert np.all((f[finite] > 2.0 ** -25) & (f[finite] < 2.0 ** -14))
    assert np.all(np.abs(dst[diff][finite].astype(np.int32) - ieee[diff][finite].astype(np.int32)) <= 1)
    assert lib.tron_floatbits_to_halfbits(0x33000001) == 0 and int(ieee[list(bits).index(0x33000001)]) == 1
    # doubles
    for v in (0.0, 1.0, -2.5, 65504.0, 65520.0, 1e-8, 6.1e-5, float("inf")):
        b = struct.unpack("<Q", struct.pack("<d", v))[0]
        with np.errstate(over="ignore"):
            assert lib.tron_doublebits_to_halfbits(b) == int(np.float64(v).astype(np.float16).view(np.uint16))


def test_ra_reader_rejects_crafted_headers(lib, tmp_path):
    """A file cannot claim pinned storage (the in-memory RA_FLAG_PINNED_DATA bit would send ra_free to
    cudaFreeHost with a malloc'ed pointer), and its size field must equal prod(dims) * elbyte."""
    import tron_b200 as t
    f = str(tmp_path / "ok.ra")
    a = np.arange(24, dtype=np.float32)
    t.ra_write(f, a, dims=[4, 6], eltype=3)
    raw = bytearray(open(f, "rb").read())
    pinned = bytearray(raw)
    struct.pack_into("<Q", pinned, 8, 1 << 62)                    # flags word: the in-memory pinned bit
    fp = str(tmp_path / "pinned.ra")
    open(fp, "wb").write(pinned)
    r = t.api.RaStruct()
    assert lib.ra_read(C.byref(r), fp.encode()) == 0
    assert r.flags & (1 << 62) == 0
    lib.ra_free(C.byref(r))
    wrong = bytearray(raw)
    struct.pack_into("<Q", wrong, 32, len(a) * 4 + 4)             # size field off by one element
    fw = str(tmp_path / "wrong.ra")
    open(fw, "wb").write(wrong + b"\0\0\0\0")
    with pytest.raises(t.TronError):
        t.ra_read(fw)
    huge = bytearray(raw)
    struct.pack_into("<2Q", huge, 48, 1 << 40, 1 << 40)           # dims whose product wraps 64 bits with elbyte
    fh = str(tmp_path / "huge.ra")
    open(fh, "wb").write(huge)
    with pytest.raises(t.TronError):
        t.ra_read(fh)
    with pytest.raises(t.TronError, match="2\\^60"):
        t.geometry(t.make_config([1 << 30, 1, 1 << 30, 1 << 30, 1], adjoint=True))


def test_build_records_the_hash_of_its_sources(lib):
    """load/build use the library as is only when it was built from the sources in the tree."""
    from tron_b200 import build
    assert os.path.isfile(build.HASHFILE)
    assert open(build.HASHFILE).read().strip() == build.source_hash()
