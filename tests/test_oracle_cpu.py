"""CPU tests of the oracle (oracle/tron_oracle.c): pinned against the golden vectors that the
UNMODIFIED reference produced on a B200 (tests/golden/*.npz, made by tests/golden/make_golden.py),
plus self-consistency of its building blocks."""
import os

import numpy as np
import pytest

from util import PARITY_CASES, case_input, rel_l2, synth_complex

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg(oracle, dims, flags):
    return oracle.config(dims, flags.get("adjoint", False), golden=flags.get("golden", False),
                         gridos=flags.get("gridos", 2.0), kernwidth=flags.get("kernwidth", 2.0),
                         undersamp=flags.get("undersamp", 1.0), prof_slide=flags.get("prof_slide", 0),
                         skip_angles=flags.get("skip_angles", 0))


@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_oracle_matches_reference_golden(oracle, name):
    """Whole pipeline, every case of the parity matrix, against the reference's own output."""
    dims, flags = PARITY_CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = _cfg(oracle, dims, flags)
    assert [int(x) for x in cfg.out_dims] == [int(x) for x in gold["out_dims"]]
    geom = dict(zip([str(k) for k in gold["geom_keys"]], [int(v) for v in gold["geom"]]))
    for k in ("nc", "nro", "npe1", "npe1work", "nx", "nz", "nxos"):
        assert getattr(cfg, k) == geom[k], k
    oracle.set_trig_table(gold["ct"], gold["st"])
    try:
        got = oracle.recon(cfg, case_input(name))
    finally:
        oracle.set_trig_table(None)
    assert rel_l2(got, gold["out"]) <= 1e-5, rel_l2(got, gold["out"])


@pytest.mark.parametrize("name", ["P1_gold", "P3_6ch", "P4_fwdG"])
def test_oracle_with_libm_trig_stays_close(oracle, name):
    """Without the SFU table the oracle uses libm sin/cos: a few edge taps flip, nothing more."""
    dims, flags = PARITY_CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    got = oracle.recon(_cfg(oracle, dims, flags), case_input(name))
    assert rel_l2(got, gold["out"]) <= 2e-4


def test_oracle_stage_vectors(oracle):
    g = np.load(os.path.join(GOLD, "stages.npz"))
    s = synth_complex((24, 64, 2), stream=101)
    u = synth_complex((64, 64, 2), stream=102)
    oracle.set_trig_table(g["grid_golden_ct"], g["grid_golden_st"])
    got = oracle.grid(s, 64, 2, 64, 24, W=2.0, skip=4, golden=True)
    assert rel_l2(got, g["grid_golden"]) <= 1e-5
    oracle.set_trig_table(g["grid_linear_ct"], g["grid_linear_st"])
    got = oracle.grid(s, 64, 2, 64, 24, W=2.0, skip=0, golden=False)
    assert rel_l2(got, g["grid_linear"]) <= 1e-5
    oracle.set_trig_table(g["degrid_golden_ct"], g["degrid_golden_st"])
    got = oracle.degrid(u, 64, 2, 64, 20, W=2.0, skip=2, golden=True)
    assert rel_l2(got, g["degrid_golden"]) <= 1e-5
    oracle.set_trig_table(g["degrid_linear_ct"], g["degrid_linear_st"])
    got = oracle.degrid(u, 64, 2, 64, 20, W=2.0, skip=0, golden=False)
    assert rel_l2(got, g["degrid_linear"]) <= 1e-5
    oracle.set_trig_table(None)
    ones = np.ones((32, 32, 1), dtype=np.complex64)
    assert rel_l2(oracle.deapod(ones, 32, 1, 2.0, 2.0), g["deapod_adj"]) <= 1e-5
    assert rel_l2(oracle.deapod(np.ones((64, 64, 1), dtype=np.complex64), 64, 1, 2.0, 1.0), g["deapod_fwd"]) <= 1e-5


@pytest.mark.parametrize("n", [8, 16, 24, 15, 96, 128])
def test_oracle_fft2_is_cufft_convention(oracle, n):
    """Unnormalised; sign -1 = FORWARD (e^{-i}), +1 = INVERSE (e^{+i}); channel-interleaved batch."""
    a = synth_complex((n, n, 3), stream=n)
    f = oracle.fft2(a, n, 3, -1)
    assert rel_l2(f, np.fft.fft2(a.astype(np.complex128), axes=(0, 1))) < 1e-6
    b = oracle.fft2(a, n, 3, +1)
    assert rel_l2(b, np.fft.ifft2(a.astype(np.complex128), axes=(0, 1)) * n * n) < 1e-6


def test_oracle_fftshift_pad_crop_quirks(oracle):
    n, nc = 6, 2
    a = synth_complex((n, n, nc), stream=3)
    fwd = oracle.fftshift(a, n, nc, inverse=False)
    assert np.array_equal(fwd, np.roll(a, (n // 2, n // 2), axis=(0, 1)))
    inv = oracle.fftshift(a, n, nc, inverse=True)
    assert np.array_equal(inv, np.roll(a, (n - n // 2, n - n // 2), axis=(0, 1)))
    odd = synth_complex((5, 5, 1), stream=4)
    assert np.array_equal(oracle.fftshift(oracle.fftshift(odd, 5, 1, False), 5, 1, True), odd)
    # pad drops source row 0 and column 0 (tron.cu:449-450)
    src = synth_complex((4, 4, nc), stream=5)
    p = oracle.pad(src, 8, 4, nc)
    assert np.all(p[2, :, :] == 0) and np.all(p[:, 2, :] == 0)
    assert np.array_equal(p[3:6, 3:6, :], src[1:4, 1:4, :])
    assert np.count_nonzero(p) == 9 * nc
    # crop takes the centred window
    big = synth_complex((8, 8, nc), stream=6)
    assert np.array_equal(oracle.crop(big, 4, 8, nc), big[2:6, 2:6, :])


def test_oracle_kernel_scalars(oracle):
    L = oracle.lib
    # I0 rational approximation against the series
    for x in (0.0, 0.5, 1.0, 3.0, 9.36, 14.0):
        k = np.arange(0, 60)
        fact = np.cumprod(np.concatenate(([1.0], np.arange(1, 60, dtype=np.float64))))
        series = float(np.sum((x * x / 4.0) ** k / fact ** 2))
        assert abs(L.oracle_besseli0(x) - series) / series < 1e-6
    # support and symmetry of the window, value at the edge = I0(0)/(2W)
    for W in (2.0, 2.5, 3.0, 6.0):
        assert L.oracle_gridkernel(W, W) == 0.0 and L.oracle_gridkernel(-W, W) == 0.0
        assert L.oracle_gridkernel(0.3, W) == L.oracle_gridkernel(-0.3, W)
        assert abs(L.oracle_gridkernel(np.nextafter(np.float32(W), np.float32(0)), W) - 0.5 / W) < 1e-3
    # linear-angle conventions differ between the directions
    assert abs(L.oracle_spoke_angle_grid(0, 64, 0, 0) - np.pi / 2) < 1e-6
    assert L.oracle_spoke_angle_degrid(0, 64, 0, 0) == 0.0
    assert abs(L.oracle_spoke_angle_grid(16, 64, 0, 0) - np.pi) < 1e-6
    assert abs(L.oracle_spoke_angle_degrid(16, 64, 0, 0) - np.pi / 4) < 1e-6
    # golden angle: f32 product of the absolute index, reduced to [0, 2pi)
    for pe in (0, 1, 7, 20270, 205823):
        want = np.fmod(np.float32(1.9416089796736116) * np.float32(pe), np.float32(2 * np.pi))
        assert L.oracle_spoke_angle_grid(pe, 204, 0, 1) == np.float32(want)
        assert L.oracle_spoke_angle_grid(pe - 3, 204, 3, 1) == np.float32(want) if pe >= 3 else True


def test_oracle_grid_band_and_dc_quirks(oracle):
    """F4: support is square n annulus; F5: r = 0 is visited twice near the centre."""
    nxos, nro, npe = 32, 32, 9
    hits = oracle.grid_hits(nxos, nro, npe, W=2.0, skip=0, golden=True)
    cell, pe, r, ridx = hits.T
    X, Y = cell % nxos - nxos // 2, cell // nxos - nxos // 2
    R = np.hypot(X, Y).astype(np.float32)
    assert np.all(np.abs(r) <= np.minimum(np.floor(R + 2.0), nxos // 2 - 1))
    assert np.all(np.abs(r) >= np.maximum(np.ceil(R - 2.0), 0))
    assert np.array_equal(ridx, r)                      # nro == nxos
    centre = (X == 0) & (Y == 0) & (r == 0)
    assert centre.sum() == 2 * npe                      # aligned + anti-aligned loop
    # a textbook scatter gridder would apply more taps than the band allows
    t = np.array([oracle.lib.oracle_spoke_angle_grid(int(p), npe, 0, 1) for p in range(npe)], dtype=np.float32)
    square = 0
    for p in range(npe):
        for rr in range(-(nxos // 2 - 1), nxos // 2):
            kx, ky = rr * np.cos(t[p]), rr * np.sin(t[p])
            square += (np.floor(kx + 2) - np.ceil(kx - 2) + 1) * (np.floor(ky + 2) - np.ceil(ky - 2) + 1)
    assert len(hits) - npe < square                     # (the doubled r=0 taps excluded)
    # with gridos 1.5 the radius index is truncated toward zero
    hits = oracle.grid_hits(24, 32, 5, W=2.0, skip=0, golden=True)
    assert np.array_equal(hits[:, 3], np.trunc(hits[:, 2] * 32 / 24).astype(np.int32))


def test_oracle_adjoint_linearity_and_window(oracle):
    dims, flags = PARITY_CASES["P2_slide"]
    cfg = _cfg(oracle, dims, flags)
    x = case_input("P2_slide")
    out = oracle.recon(cfg, x).reshape(cfg.nz, cfg.nx, cfg.nx)
    # slice z only depends on spokes [z*slide, z*slide + npe1work)
    x2 = x.copy().reshape(cfg.npe1, cfg.nro, cfg.nc)
    x2[: 2 * cfg.prof_slide] = 0
    out2 = oracle.recon(cfg, x2.ravel()).reshape(cfg.nz, cfg.nx, cfg.nx)
    assert not np.array_equal(out[0], out2[0]) and not np.array_equal(out[1], out2[1])
    assert np.array_equal(out[2:], out2[2:])


# ---------------------------------------------------------------- Walsh combine, CGNR
def test_oracle_walsh_matches_reference_golden(oracle):
    """coilcombinewalsh (tron.cu:270-302) run by oracle/_ref on a B200 -> tests/golden/walsh.npz."""
    from util import WALSH_CASES, walsh_input
    gold = np.load(os.path.join(GOLD, "walsh.npz"))
    for nimg, nc, npatch in WALSH_CASES:
        got = oracle.walsh(walsh_input(nimg, nc), nimg, nc, npatch)
        want = gold["walsh_%d_%d_%d" % (nimg, nc, npatch)]
        assert rel_l2(got, want) <= 1e-5, (nimg, nc, npatch, rel_l2(got, want))


def test_oracle_walsh_properties(oracle):
    one = synth_complex((12, 12, 1), stream=301)
    assert np.array_equal(oracle.walsh(one, 12, 1, 1), one[:, :, 0])            # tron.cu:277-278
    # rank-one coil images z_c = s_c m: the dominant eigenvector is s/|s| up to the phase of s^H 1,
    # so out = (1^T s)/|1^T s| |s| m
    nimg, nc = 16, 4
    m = synth_complex((nimg, nimg), stream=302)
    s = (np.array([1.0, 0.5j, -0.7, 0.2 + 0.3j])).astype(np.complex64)
    z = (m[:, :, None] * s[None, None, :]).astype(np.complex64)
    out = oracle.walsh(z, nimg, nc, 1)
    ph = s.sum() / abs(s.sum())
    assert rel_l2(out, ph * np.linalg.norm(s) * m) <= 1e-5
    # npatch = 0 is the pixel's own outer product: |out| is the root sum of squares
    z = synth_complex((nimg, nimg, 6), stream=303)
    assert rel_l2(np.abs(oracle.walsh(z, nimg, 6, 0)), np.abs(oracle.sos(z, nimg, 6))) <= 1e-5


def test_oracle_walsh_in_the_slice_loop(oracle):
    dims, flags = PARITY_CASES["P2_slide"]
    x = case_input("P2_slide")
    cfg = oracle.config(dims, True, golden=True, undersamp=0.25, prof_slide=7, skip_angles=3,
                        coil_combine=1, walsh_npatch=1)
    out = oracle.recon(cfg, x).reshape(cfg.nz, cfg.nx, cfg.nx)
    coils = oracle.adj_coils(cfg, x[cfg.nc * cfg.nro * 7 * 2:], peoffset=14)      # slice 2
    assert np.array_equal(out[2], oracle.walsh(coils, cfg.nx, cfg.nc, 1))


def _cg_phantom(oracle, nc, nx, npe):
    import ctypes as C
    from oracle.oracle import _Cfg, _ptr
    y, x = (np.mgrid[0:nx, 0:nx] - nx / 2).astype(np.float32)
    obj = ((x ** 2 + y ** 2) < (0.36 * nx) ** 2).astype(np.float32) + 0.5 * (((x - 4) ** 2 + (y + 3) ** 2) < 16)
    truth = np.zeros((nx, nx, nc), dtype=np.complex64)
    for c in range(nc):
        truth[:, :, c] = obj * np.exp(1j * 0.3 * c * x / nx) * (1 + 0.2 * c)
    fcfg = oracle.config([nc, 1, nx, nx, 1], False, golden=True)
    full = np.zeros((fcfg.npe1work, fcfg.nro, nc), dtype=np.complex64)
    oracle.lib.oracle_nufft_fwd_slice.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p]
    oracle.lib.oracle_nufft_fwd_slice(C.byref(fcfg), _ptr(full), _ptr(truth))
    return truth, np.ascontiguousarray(full[:npe])


def test_oracle_cgnr_converges(oracle):
    """The repaired CGNR (see oracle_cgnr_coils) reduces the error against the phantom monotonically
    over 16 iterations; one iteration is the adjoint image times the optimal step."""
    nc, nx, npe = 2, 32, 64
    truth, samples = _cg_phantom(oracle, nc, nx, npe)
    cfg = oracle.config([nc, 1, 2 * nx, npe, 1], True, golden=True)
    errs = [rel_l2(oracle.cgnr_coils(cfg, samples, 0, it), truth) for it in (1, 2, 4, 8, 16)]
    assert all(b < a for a, b in zip(errs, errs[1:])), errs
    assert errs[-1] < 0.08, errs
    x1 = oracle.cgnr_coils(cfg, samples, 0, 1)
    adj = oracle.adj_coils(cfg, samples)
    inner = (slice(1, None), slice(1, None))          # row 0 / column 0 are cleared (pad drops them)
    ratio = x1[inner] / adj[inner]
    # B uses the forward model's deapodisation table, the plain adjoint its own (F9 quirk): the
    # two differ by a smooth real factor close to a constant
    assert np.std(np.abs(ratio)) / np.mean(np.abs(ratio)) < 0.05
    assert np.all(x1[0] == 0) and np.all(x1[:, 0] == 0)


def test_oracle_cgnr_in_the_slice_loop(oracle):
    dims = [2, 1, 32, 40, 1]
    x = synth_complex((int(np.prod(dims)),), stream=304)
    cfg = oracle.config(dims, True, golden=True, undersamp=0.5, prof_slide=8, skip_angles=1, niter=2)
    out = oracle.recon(cfg, x).reshape(cfg.nz, cfg.nx, cfg.nx)
    z = 2
    coils = oracle.cgnr_coils(cfg, x[cfg.nc * cfg.nro * 8 * z:], peoffset=8 * z, niter=2)
    assert np.array_equal(out[z], oracle.sos(coils, cfg.nx, cfg.nc))
