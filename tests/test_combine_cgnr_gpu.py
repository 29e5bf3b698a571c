"""GPU parity tests of the SURVEY section 8(f) rows N3 / N4: the adaptive (Walsh) coil combine
(tron.cu:270-302) and the CGNR iteration (tron.cu:665-720), through the C ABI, against
  (a) the reference's own coilcombinewalsh kernel (oracle/_ref, same GPU, nc <= 6),
  (b) the committed golden vectors tests/golden/walsh.npz (made by (a)),
  (c) the CPU oracle (oracle/tron_oracle.c) for both.
The reference's CGNR is self-declared broken (tron.cu:670); parity for it is against the
oracle's restatement of the repaired algorithm, plus convergence properties.
"""
import os

import numpy as np
import pytest

from util import WALSH_CASES, rel_l2, synth_complex, walsh_input

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))


def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def walsh_gpu(coil, npatch):
    """(nslices, nimg, nimg, nc) complex64 -> (nslices, nimg, nimg) through tron_coilcombine_walsh_device."""
    import tron_b200 as t
    torch = torch_cuda()
    coil = np.ascontiguousarray(coil, dtype=np.complex64)
    ns, nimg, _, nc = coil.shape
    d_in = torch.from_numpy(coil.view(np.float32)).cuda()
    d_out = torch.zeros((ns, nimg, nimg, 2), dtype=torch.float32, device="cuda")
    t.coilcombine_walsh_device(d_out.data_ptr(), d_in.data_ptr(), nimg, nc, npatch, ns)
    torch.cuda.synchronize()
    return d_out.cpu().numpy().view(np.complex64)[..., 0]


def install_trig(oracle, reflib, cfg, golden, skip):
    """Give the oracle the SFU sin/cos of the spokes (as tests/golden/make_golden.py stores them for
    the fixed cases): libm's differ in the last bits, which moves edge taps (tron_oracle.c)."""
    ntab = skip + cfg.npe1 if golden else cfg.npe1work
    ct, st = reflib.spoke_cs(ntab, cfg.npe1work, 0, golden, False)
    oracle.set_trig_table(ct, st)


# ------------------------------------------------------------------ Walsh combine
@pytest.mark.parametrize("nimg,nc,npatch", WALSH_CASES)
def test_walsh_vs_reference_kernel_and_golden(lib, reflib, oracle, nimg, nc, npatch):
    coil = walsh_input(nimg, nc)
    got = walsh_gpu(coil[None], npatch)[0]
    want = reflib.walsh(coil, nimg, nc, npatch)
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)
    assert rel_l2(got, oracle.walsh(coil, nimg, nc, npatch)) <= TOL_F32
    gold = np.load(os.path.join(HERE, "golden", "walsh.npz"))["walsh_%d_%d_%d" % (nimg, nc, npatch)]
    assert rel_l2(got, gold) <= TOL_F32


@pytest.mark.parametrize("nc,npatch", [(8, 1), (10, 1), (16, 1), (32, 1), (64, 1), (64, 2), (96, 0), (3, 1)])
def test_walsh_many_coils_vs_oracle(lib, oracle, nc, npatch):
    """nc > 6 is beyond the reference kernel (MAXCHAN / NCHAN = 6, tron.h:50-51): oracle only.
    nc > 8 runs the matrix-free warp-per-pixel kernel."""
    nimg = 24
    coil = walsh_input(nimg, nc)
    got = walsh_gpu(coil[None], npatch)[0]
    assert rel_l2(got, oracle.walsh(coil, nimg, nc, npatch)) <= TOL_F32


def test_walsh_batched_slices_and_single_coil(lib, oracle):
    coil = np.stack([walsh_input(20, 6) * (1 + s) for s in range(5)])
    got = walsh_gpu(coil, 1)
    for s in range(5):
        assert rel_l2(got[s], oracle.walsh(coil[s], 20, 6, 1)) <= TOL_F32
    one = walsh_input(16, 1)
    assert np.array_equal(walsh_gpu(one[None], 1)[0], one[:, :, 0])          # tron.cu:277-278
    # a patch of zeros gives 0 (the reference kernel: 0 * (1/0) = NaN), small and wide kernels alike
    for nc in (6, 16):
        z = walsh_input(12, nc)
        z[:4] = 0
        out = walsh_gpu(z[None], 1)[0]
        assert np.all(out[:3] == 0) and np.all(np.isfinite(out.view(np.float32)))
        assert rel_l2(out, oracle.walsh(z, 12, nc, 1)) <= TOL_F32


def test_walsh_phase_property(lib):
    """A common phase on every coil passes through, and so does a real scale."""
    coil = walsh_input(24, 6)
    base = walsh_gpu(coil[None], 1)[0]
    rot = walsh_gpu((coil * np.exp(0.7j).astype(np.complex64))[None], 1)[0]
    assert rel_l2(rot, base * np.exp(0.7j)) <= 1e-5
    assert rel_l2(walsh_gpu((coil * np.float32(3.0))[None], 1)[0], 3.0 * base) <= 1e-5


@pytest.mark.parametrize("dims,flags", [
    ([6, 1, 64, 100, 1], dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=7, skip_angles=3)),
    ([4, 1, 96, 150, 1], dict(adjoint=True, prof_slide=50, undersamp=0.5)),
    ([16, 1, 64, 40, 1], dict(adjoint=True, golden=True)),
])
def test_pipeline_with_walsh_vs_oracle(lib, oracle, reflib, dims, flags):
    """tron -a -w 1: adjoint NUFFT then coilcombinewalsh, i.e. recon_radial2d with tron.cu:766
    enabled instead of tron.cu:764."""
    import tron_b200 as t
    torch_cuda()
    h_in = synth_complex((int(np.prod(dims)),), stream=61)
    cfg = oracle.config(dims, True, golden=flags.get("golden", False), undersamp=flags.get("undersamp", 1.0),
                        prof_slide=flags.get("prof_slide", 0), skip_angles=flags.get("skip_angles", 0),
                        coil_combine=1, walsh_npatch=1)
    install_trig(oracle, reflib, cfg, flags.get("golden", False), flags.get("skip_angles", 0))
    want = oracle.recon(cfg, h_in)
    oracle.set_trig_table(None)
    with t.Plan(t.make_config(dims, coil_combine=1, walsh_npatch=1, **flags)) as p:
        got = p.recon_host(h_in)
        assert p.last_launches() > 0
    assert rel_l2(got, want) <= TOL_F32, rel_l2(got, want)


def test_pipeline_with_walsh_equals_percoil_plus_kernel(lib):
    """The fused pipeline is exactly per_coil_out followed by the stand-alone combine."""
    import tron_b200 as t
    torch_cuda()
    dims = [6, 1, 128, 180, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.3, prof_slide=11)
    h_in = synth_complex((int(np.prod(dims)),), stream=62)
    with t.Plan(t.make_config(dims, per_coil_out=True, **flags)) as p:
        coils = p.recon_host(h_in).reshape(p.geom.nz, 64, 64, 6)
    with t.Plan(t.make_config(dims, coil_combine=1, walsh_npatch=2, batch_slices=8, **flags)) as p:
        got = p.recon_host(h_in).reshape(p.geom.nz, 64, 64)
    assert np.array_equal(got, walsh_gpu(coils, 2))


# ------------------------------------------------------------------ CGNR
def _phantom_samples(oracle, nc, nx, npe, skip=0):
    """Consistent golden-angle data: the oracle's forward model applied to a smooth multi-coil phantom."""
    import ctypes as C
    from oracle.oracle import _Cfg, _ptr
    y, x = (np.mgrid[0:nx, 0:nx] - nx / 2).astype(np.float32)
    obj = ((x ** 2 + y ** 2) < (0.36 * nx) ** 2).astype(np.float32) + 0.5 * (((x - 4) ** 2 + (y + 3) ** 2) < 16)
    truth = np.zeros((nx, nx, nc), dtype=np.complex64)
    for c in range(nc):
        truth[:, :, c] = obj * np.exp(1j * 0.3 * c * x / nx) * (1 + 0.2 * c)
    fcfg = oracle.config([nc, 1, nx, nx, 1], False, golden=True, skip_angles=skip)
    full = np.zeros((fcfg.npe1work, fcfg.nro, nc), dtype=np.complex64)
    oracle.lib.oracle_nufft_fwd_slice.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p]
    oracle.lib.oracle_nufft_fwd_slice(C.byref(fcfg), _ptr(full), _ptr(truth))
    assert npe <= fcfg.npe1work
    return truth, np.ascontiguousarray(full[:npe])


@pytest.mark.parametrize("niter", [1, 2, 4])
def test_cgnr_vs_oracle(lib, oracle, reflib, niter):
    import tron_b200 as t
    torch_cuda()
    nc, nx, npe = 2, 32, 48
    truth, samples = _phantom_samples(oracle, nc, nx, npe)
    dims = [nc, 1, 2 * nx, npe, 1]
    cfg = oracle.config(dims, True, golden=True)
    install_trig(oracle, reflib, cfg, True, 0)
    want = oracle.cgnr_coils(cfg, samples, 0, niter)
    oracle.set_trig_table(None)
    with t.Plan(t.make_config(dims, adjoint=True, golden=True, niter=niter, per_coil_out=True)) as p:
        got = p.recon_host(samples).reshape(nx, nx, nc)
    # every iteration passes through both operators and divides two inner products
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)


def test_cgnr_sliding_windows_vs_oracle_and_rss(lib, oracle, reflib):
    """-i 2 on overlapping golden-angle windows with skip: every slice uses its own angles
    (skip + z*slide) in BOTH operators, and the coil combine follows."""
    import tron_b200 as t
    torch_cuda()
    dims = [4, 1, 64, 60, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=9, skip_angles=2)
    h_in = synth_complex((int(np.prod(dims)),), stream=63)
    cfg = oracle.config(dims, True, golden=True, undersamp=0.5, prof_slide=9, skip_angles=2, niter=2)
    install_trig(oracle, reflib, cfg, True, 2)
    want = oracle.recon(cfg, h_in)
    with t.Plan(t.make_config(dims, niter=2, batch_slices=3, **flags)) as p:
        assert p.geom.nz == 4
        got = p.recon_host(h_in)
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)
    cfgw = oracle.config(dims, True, golden=True, undersamp=0.5, prof_slide=9, skip_angles=2, niter=2,
                         coil_combine=1, walsh_npatch=1)
    with t.Plan(t.make_config(dims, niter=2, coil_combine=1, walsh_npatch=1, **flags)) as p:
        gotw = p.recon_host(h_in)
    wantw = oracle.recon(cfgw, h_in)
    oracle.set_trig_table(None)
    assert rel_l2(gotw, wantw) <= 1e-4


def test_cgnr_converges_on_consistent_data(lib, oracle):
    """Error against the phantom falls with the iteration count and ends well below the plain
    adjoint's (which is only proportional to the image); the same holds for linear angles."""
    import tron_b200 as t
    torch_cuda()
    nc, nx, npe = 2, 64, 128
    truth, samples = _phantom_samples(oracle, nc, nx, npe)
    dims = [nc, 1, 2 * nx, npe, 1]
    errs = []
    for niter in (1, 2, 4, 8, 16):
        with t.Plan(t.make_config(dims, adjoint=True, golden=True, niter=niter, per_coil_out=True)) as p:
            got = p.recon_host(samples).reshape(nx, nx, nc)
        errs.append(rel_l2(got, truth))
    assert all(b < a for a, b in zip(errs, errs[1:])), errs
    assert errs[-1] < 0.1 and errs[-1] < 0.5 * errs[0], errs


def test_cgnr_fp16_input_and_device_api(lib):
    import tron_b200 as t
    torch = torch_cuda()
    dims = [2, 1, 64, 50, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=9, niter=3)
    h_in = synth_complex((int(np.prod(dims)),), stream=64)
    with t.Plan(t.make_config(dims, **flags)) as p:
        want = p.recon_host(h_in)
        d_in = torch.from_numpy(h_in.view(np.float32)).cuda()
        d_out = torch.zeros(want.size * 2, dtype=torch.float32, device="cuda")
        p.recon_device(d_out.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.complex64), want)       # deterministic reductions
    h16 = h_in.view(np.float32).astype(np.float16)
    with t.Plan(t.make_config(dims, half_in=True, **flags)) as p:
        got16 = p.recon_host(h16)
    with t.Plan(t.make_config(dims, **flags)) as p:
        ref16 = p.recon_host(h16.astype(np.float32).view(np.complex64))
    assert rel_l2(got16, ref16) <= 1e-5


def test_legacy_cgnr_symbol(lib, oracle):
    """tron_cgnr_radial2d(d_out, d_in, j, niter) after tron_set_config (tron.cu:665)."""
    import ctypes as C
    import tron_b200 as t
    torch = torch_cuda()
    nc, nx, npe = 2, 32, 40
    dims = [nc, 1, 2 * nx, npe, 1]
    h_in = synth_complex((int(np.prod(dims)),), stream=65)
    cfg = t.make_config(dims, adjoint=True, golden=True)
    assert lib.tron_set_config(C.byref(cfg)) == 0
    d_in = torch.from_numpy(h_in.view(np.float32)).cuda()
    d_out = torch.zeros(nx * nx * nc * 2, dtype=torch.float32, device="cuda")
    lib.tron_cgnr_radial2d(C.c_void_p(d_out.data_ptr()), C.c_void_p(d_in.data_ptr()), 0, 2)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.complex64)
    with t.Plan(t.make_config(dims, adjoint=True, golden=True, niter=2, per_coil_out=True)) as p:
        want = p.recon_host(h_in)
    assert np.array_equal(got, want)
    # Caxpy (tron.cu:658): z = y + alpha x
    x = torch.randn(1000, 2, device="cuda"); y = torch.randn(1000, 2, device="cuda"); z = torch.empty_like(x)
    assert lib.tron_launch_Caxpy(C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(x.data_ptr()),
                                 C.c_float(0.25), 1000, 7, 64, None) == 0
    torch.cuda.synchronize()
    assert torch.allclose(z, y + 0.25 * x, atol=1e-6)
    lib.tron_shutdown()


@pytest.mark.parametrize("nro", [512, 256])
def test_percoil_output_two_stage_fft_matches_radix8_and_rss(lib, monkeypatch, nro):
    """Per-coil images at the benchmark line lengths come from the two-stage pass B
    (p2w_adj_pass_b_coil); they must agree with the radix-8 pass and reproduce the RSS image."""
    import tron_b200 as t
    torch_cuda()
    dims = [6, 1, nro, 90, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.1, prof_slide=13)
    h_in = synth_complex((int(np.prod(dims)),), stream=66)
    nx = nro // 2
    with t.Plan(t.make_config(dims, per_coil_out=True, **flags)) as p:
        coils = p.recon_host(h_in).reshape(p.geom.nz, nx, nx, 6)
    with t.Plan(t.make_config(dims, **flags)) as p:
        rss = p.recon_host(h_in).reshape(p.geom.nz, nx, nx)
    monkeypatch.setenv("TRON_FFT_R8", "1")
    with t.Plan(t.make_config(dims, per_coil_out=True, **flags)) as p:
        coils8 = p.recon_host(h_in).reshape(coils.shape)
    assert rel_l2(coils, coils8) <= 2e-6
    assert rel_l2(np.sqrt((np.abs(coils.astype(np.complex128)) ** 2).sum(-1)), rss.real) <= 2e-6
    assert np.all(rss.imag == 0)


def test_cli_walsh_and_cgnr_flags(lib, tmp_path):
    """`tron -a -G ... -w 1` and `-i 2` write what the plan API computes (same 88-byte header as
    the adjoint output of the reference CLI); `-i` without `-a` is refused with a message."""
    import subprocess
    import tron_b200 as t
    torch_cuda()
    dims = [4, 1, 64, 60, 1]
    h_in = synth_complex((int(np.prod(dims)),), stream=67)
    fin, fout = str(tmp_path / "in.ra"), str(tmp_path / "out.ra")
    t.ra_write(fin, h_in, dims=dims)
    base = ["-a", "-G", "-u", "0.5", "-d", "9"]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=9)
    for extra, kw in ((["-w", "1"], dict(coil_combine=1, walsh_npatch=1)), (["-i", "2"], dict(niter=2)),
                      (["-i", "1", "-w", "2"], dict(niter=1, coil_combine=1, walsh_npatch=2))):
        subprocess.run([t.CLI_PATH] + base + extra + [fin, fout], check=True, timeout=300)
        got, d, eltype, elbyte = t.ra_read(fout)
        with t.Plan(t.make_config(dims, **flags, **kw)) as p:
            want = p.recon_host(h_in)
            assert d == [int(x) for x in p.geom.out_dims] and (eltype, elbyte) == (4, 8)
        assert np.array_equal(got, want)
    r = subprocess.run([t.CLI_PATH, "-i", "2", fin, fout], capture_output=True, text=True)
    assert r.returncode == 1 and "CGNR" in r.stderr


def _fuzz_percoil_cases(n, seed=77):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        nc = int(rng.choice([2, 4, 6, 8]))
        nro = int(rng.choice([32, 48, 64]))
        npe1work = int(rng.integers(8, 40))
        nsl = int(rng.integers(1, 4))
        slide = int(rng.integers(1, npe1work + 1)) if nsl > 1 else 0
        dims = [nc, 1, nro, npe1work + (nsl - 1) * slide, 1]
        flags = dict(adjoint=True, golden=bool(rng.integers(0, 2)), undersamp=(npe1work + 0.5) / nro,
                     prof_slide=slide, skip_angles=int(rng.integers(0, 9)))
        extra = dict(niter=int(rng.integers(0, 4)), coil_combine=int(rng.integers(0, 2)),
                     walsh_npatch=int(rng.integers(0, 3)))
        if extra["niter"] == 0 and extra["coil_combine"] == 0:
            extra["coil_combine"] = 1
        out.append((i, dims, flags, extra))
    return out


@pytest.mark.parametrize("i,dims,flags,extra", _fuzz_percoil_cases(12), ids=lambda v: str(v) if isinstance(v, int) else None)
def test_fuzz_walsh_cgnr_vs_oracle(lib, oracle, reflib, i, dims, flags, extra):
    """Seeded random geometries through -i / -w (any mix) against the oracle's restatement."""
    import tron_b200 as t
    torch_cuda()
    h_in = synth_complex((int(np.prod(dims)),), stream=600 + i)
    cfg = oracle.config(dims, True, golden=flags["golden"], undersamp=flags["undersamp"], prof_slide=flags["prof_slide"],
                        skip_angles=flags["skip_angles"], **extra)
    install_trig(oracle, reflib, cfg, flags["golden"], flags["skip_angles"])
    want = oracle.recon(cfg, h_in)
    oracle.set_trig_table(None)
    with t.Plan(t.make_config(dims, **flags, **extra)) as p:
        got = p.recon_host(h_in)
    assert np.all(np.isfinite(got.view(np.float32)))
    if extra["niter"] and extra["coil_combine"]:
        # the combine normalises by |A x| and fixes the phase by x^H z: where the coil images nearly cancel it
        # amplifies the 5e-5 the iterations accumulate between libm and SFU arithmetic
        assert rel_l2(np.abs(got), np.abs(want)) <= 2e-4 and rel_l2(got, want) <= 1e-3, (dims, flags, extra)
    else:
        assert rel_l2(got, want) <= (1e-4 if extra["niter"] else TOL_F32), (dims, flags, extra, rel_l2(got, want))


def test_walsh_and_cgnr_with_fp16_output(lib):
    """-H (complex-half output) after the Walsh combine and after CGNR + RSS: the float result rounded once."""
    import tron_b200 as t
    torch_cuda()
    dims = [6, 1, 64, 70, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.4, prof_slide=11)
    h_in = synth_complex((int(np.prod(dims)),), stream=68)
    for extra in (dict(coil_combine=1, walsh_npatch=1), dict(niter=2), dict(niter=1, coil_combine=1, walsh_npatch=2)):
        with t.Plan(t.make_config(dims, **flags, **extra)) as p:
            full = p.recon_host(h_in)
        with t.Plan(t.make_config(dims, half_out=True, **flags, **extra)) as p:
            half = p.recon_host(h_in)
        assert half.dtype == np.float16 and half.shape == (full.size, 2)
        want = full.view(np.float32).reshape(-1, 2).astype(np.float16)
        assert np.array_equal(half, want), extra


def test_cgnr_single_coil_complex_output(lib, oracle, reflib):
    """nc = 1: the iterate passes through as a complex image (tron.cu:265-266), with or without -w."""
    import tron_b200 as t
    torch_cuda()
    dims = [1, 1, 64, 40, 1]
    flags = dict(adjoint=True, golden=True, undersamp=0.5, prof_slide=4, skip_angles=2)
    h_in = synth_complex((int(np.prod(dims)),), stream=69)
    cfg = oracle.config(dims, True, golden=True, undersamp=0.5, prof_slide=4, skip_angles=2, niter=3)
    install_trig(oracle, reflib, cfg, True, 2)
    want = oracle.recon(cfg, h_in)
    oracle.set_trig_table(None)
    with t.Plan(t.make_config(dims, niter=3, **flags)) as p:
        got = p.recon_host(h_in)
    with t.Plan(t.make_config(dims, niter=3, coil_combine=1, walsh_npatch=1, **flags)) as p:
        got_w = p.recon_host(h_in)
    assert np.abs(got.imag).max() > 0 and rel_l2(got, want) <= 1e-4
    assert np.array_equal(got, got_w)


# ------------------------------------------------------------------ CGNR pinned to the reference's operators
def _dense_operators_from_reference_kernels(reflib, nx, nro, npe, W=2.0):
    """Dense matrices of the two operators the CGNR iterates with, one coil, built by pushing indicator vectors
    through the REFERENCE's own kernels (oracle/_ref: gridradial2d tron.cu:465-536, degridradial2d 540-577,
    deapodkernel 390-402); the stages between them are the exact linear maps the reference's host code applies
    (pad 435-457, fftshift 161-178 = circular shift by n/2 for even n, unnormalised cuFFT, crop 418-431).
      A (nsamp x npix): pad -> deapod(nxos, sigma 1) -> shift -> FFT forward -> shift -> degridradial2d   (639-649)
      B (npix x nsamp): precompensate ramp -> gridradial2d -> shift -> FFT inverse -> shift -> deapod(nxos, sigma 1)
                        -> crop, row 0 / column 0 cleared, sample ro = 0 dropped    (623-637, repaired: DESIGN 3.6)"""
    n = 2 * nx
    w = (n - nx) // 2
    npix, nsamp = nx * nx, npe * nro
    ones = np.ones((n, n, 1), dtype=np.complex64)
    dw = reflib.deapod(ones, n, 1, W, 1.0)[:, :, 0].astype(np.complex128)        # what deapodkernel multiplies by
    shift = lambda a: np.roll(a, (n // 2, n // 2), axis=(0, 1))
    A = np.zeros((nsamp, npix), dtype=np.complex128)
    for j0 in range(0, npix, 6):
        cols = list(range(j0, min(j0 + 6, npix)))
        u = np.zeros((n, n, 6), dtype=np.complex128)
        for c, j in enumerate(cols):
            row, col = divmod(j, nx)
            if row > 0 and col > 0:                                            # pad drops source row 0 / column 0 (449-450)
                u[row + w, col + w, c] = 1.0
        u *= dw[:, :, None]
        u = shift(np.fft.fft2(shift(u), axes=(0, 1)))
        s = reflib.degrid(u.astype(np.complex64), n, 6, nro, npe, W=W, gridos=2.0, skip=0, golden=True)
        for c, j in enumerate(cols):
            A[:, j] = s[:, :, c].reshape(-1)
    a = (2.0 - 2.0 / npe) / nro
    b = 1.0 / npe
    B = np.zeros((npix, nsamp), dtype=np.complex128)
    for i0 in range(0, nsamp, 6):
        rows = list(range(i0, min(i0 + 6, nsamp)))
        s = np.zeros((npe, nro, 6), dtype=np.complex64)
        for c, i in enumerate(rows):
            pe, ro = divmod(i, nro)
            if ro != 0:                                                        # M: ro = 0 is never gridded (Rhi <= nxos/2 - 1, 499)
                s[pe, ro, c] = np.float32(a) * abs(np.float32(ro) - np.float32(nro // 2)) + np.float32(b)   # 405-416
        g = reflib.grid(s, n, 6, nro, npe, W=W, gridos=2.0, skip=0, golden=True).astype(np.complex128)
        g = shift(np.fft.ifft2(shift(g), axes=(0, 1)) * (n * n))
        g = g * dw[:, :, None]
        img = g[w:w + nx, w:w + nx, :].copy()
        img[0, :, :] = 0
        img[:, 0, :] = 0
        for c, i in enumerate(rows):
            B[:, i] = img[:, :, c].reshape(-1)
    return A, B


@pytest.mark.parametrize("niter", [1, 2, 3, 5])
def test_cgnr_pinned_to_reference_operators_and_textbook_solver(lib, reflib, niter):
    """The reference's own tron_cgnr_radial2d is self-declared broken (tron.cu:670), so the pin has two legs that do
    not involve this repository's arithmetic: the OPERATORS are dense matrices assembled from the reference's
    kernels, the ITERATION is textbook CGNR (Knopp et al. 2007, Alg. 1, the algorithm tron.cu:665-720 names) in
    NumPy float64 on those matrices.  `tron -i niter` must agree to 1e-4."""
    import tron_b200 as t
    torch_cuda()
    nc, nx, npe = 2, 12, 20
    nro = 2 * nx
    A, B = _dense_operators_from_reference_kernels(reflib, nx, nro, npe)
    y = synth_complex((npe, nro, nc), stream=900).astype(np.complex128)
    s = 1.0 / (2 * nx) / npe                                                   # the adjoint's output scale (532), inside B
    wa, wb = (2.0 - 2.0 / npe) / nro, 1.0 / npe
    ro = np.arange(nro)
    Wp = wa * np.abs(ro - nro // 2) + wb                                       # the weights gridding really applies
    Wp[nro // 2] *= 2.0                                                        # r = 0 visited twice (512, 521)
    Wp[0] = 0.0
    Wfull = np.tile(Wp, npe)
    keep = np.tile((ro != 0).astype(np.float64), npe)                          # M
    r = [y[:, :, c].reshape(-1) * keep for c in range(nc)]
    z = [B @ rc for rc in r]
    p = [zc.copy() for zc in z]
    x = [np.zeros(nx * nx, dtype=np.complex128) for _ in range(nc)]
    zz = sum(np.vdot(zc, zc).real for zc in z)
    for it in range(niter):
        v = [A @ pc for pc in p]
        vwv = sum(np.sum(Wfull * np.abs(vc) ** 2) for vc in v)
        alpha = zz / (s * vwv)
        x = [xc + alpha * pc for xc, pc in zip(x, p)]
        if it == niter - 1:
            break
        r = [rc - alpha * keep * vc for rc, vc in zip(r, v)]
        z = [B @ rc for rc in r]
        zz2 = sum(np.vdot(zc, zc).real for zc in z)
        beta = zz2 / zz
        p = [zc + beta * pc for zc, pc in zip(z, p)]
        zz = zz2
    want = np.stack([xc.reshape(nx, nx) for xc in x], axis=-1)
    dims = [nc, 1, nro, npe, 1]
    with t.Plan(t.make_config(dims, adjoint=True, golden=True, niter=niter, per_coil_out=True)) as pl:
        got = pl.recon_host(y.astype(np.complex64)).reshape(nx, nx, nc)
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)
    # the operator pair itself: B is s A^H W' up to the annulus clipping of gridding (< 1e-2 here)
    asym = np.linalg.norm(B - s * (A.conj().T * Wfull[None, :])) / np.linalg.norm(B)
    assert asym < 5e-2, asym
