"""Shared helpers for the tests, bench.py and the golden-vector generator."""
import numpy as np

SEED = 20261017


def synth_complex(shape, stream=0, seed=SEED):
    """i.i.d. N(0,1) real/imag complex64 from a counter-based generator (Philox)."""
    rng = np.random.Generator(np.random.Philox(key=seed + stream))
    a = rng.standard_normal(tuple(shape) + (2,), dtype=np.float32)
    return np.ascontiguousarray(a).view(np.complex64).reshape(shape)


def shepp_logan(n):
    """Analytic Shepp-Logan phantom (modified contrast), complex64 with zero imaginary part."""
    ell = [(1.0, .69, .92, 0, 0, 0), (-.8, .6624, .8740, 0, -.0184, 0), (-.2, .1100, .3100, .22, 0, -18),
           (-.2, .1600, .4100, -.22, 0, 18), (.1, .2100, .2500, 0, .35, 0), (.1, .0460, .0460, 0, .1, 0),
           (.1, .0460, .0460, 0, -.1, 0), (.1, .0460, .0230, -.08, -.605, 0), (.1, .0230, .0230, 0, -.606, 0),
           (.1, .0230, .0460, .06, -.605, 0)]
    y, x = np.mgrid[-1:1:n * 1j, -1:1:n * 1j]
    img = np.zeros((n, n), dtype=np.float32)
    for A, a, b, x0, y0, phi in ell:
        p = np.deg2rad(phi)
        xr = (x - x0) * np.cos(p) + (y - y0) * np.sin(p)
        yr = -(x - x0) * np.sin(p) + (y - y0) * np.cos(p)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1] += A
    return img[::-1].astype(np.complex64)


def rel_l2(a, b):
    a = np.asarray(a).ravel().astype(np.complex128)
    b = np.asarray(b).ravel().astype(np.complex128)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# The small parity matrix, sized so the oracle finishes in seconds.
# name -> (dims, flags)
PARITY_CASES = {
    "P1_lin":   ([1, 1, 64, 64, 1],   dict(adjoint=True)),
    "P1_gold":  ([1, 1, 64, 64, 1],   dict(adjoint=True, golden=True)),
    "P2_slide": ([2, 1, 64, 100, 1],  dict(adjoint=True, golden=True, undersamp=0.25, prof_slide=7, skip_angles=3)),
    "P3_6ch":   ([6, 1, 128, 96, 1],  dict(adjoint=True, golden=True)),
    "P3_k3":    ([6, 1, 128, 96, 1],  dict(adjoint=True, golden=True, kernwidth=3.0)),
    "P3_o15":   ([6, 1, 128, 96, 1],  dict(adjoint=True, golden=True, gridos=1.5)),
    "P3_lin4":  ([4, 1, 96, 150, 1],  dict(adjoint=True, prof_slide=50, undersamp=0.5)),
    "P4_fwd":   ([1, 1, 64, 64, 1],   dict(adjoint=False)),
    "P4_fwdG":  ([1, 1, 64, 64, 1],   dict(adjoint=False, golden=True)),
    "P4_fwdk3": ([1, 1, 64, 64, 1],   dict(adjoint=False, kernwidth=3.0)),
    "P5_fwd4":  ([4, 1, 32, 32, 1],   dict(adjoint=False, golden=True)),
    "P5_fwdu":  ([2, 1, 48, 48, 1],   dict(adjoint=False, undersamp=0.5, skip_angles=5, golden=True)),
}


def case_input(name):
    dims, flags = PARITY_CASES[name]
    stream = sorted(PARITY_CASES).index(name) + 1
    n = int(np.prod(dims))
    return synth_complex((n,), stream=stream)


# coilcombinewalsh vectors (tests/golden/walsh.npz): (nimg, nc, npatch); nc <= 6 because the
# reference zeroes NCHAN^2 = 36 matrix entries whatever nchan is (tron.cu:282)
WALSH_CASES = [(32, 2, 1), (32, 4, 1), (32, 6, 1), (24, 6, 0), (24, 6, 2), (20, 4, 3)]


def walsh_input(nimg, nc):
    """Coil images with a smooth per-coil sensitivity on top of noise, (nimg, nimg, nc) complex64."""
    z = synth_complex((nimg, nimg, nc), stream=200 + nc)
    y, x = np.mgrid[0:nimg, 0:nimg].astype(np.float32) / nimg
    obj = (1.0 + np.cos(3.0 * x) * np.sin(2.0 * y)).astype(np.float32)
    out = np.empty_like(z)
    for c in range(nc):
        sens = np.exp(1j * (0.7 * c + 2.0 * x * (c + 1) / nc)) * (0.5 + (c + 1) / nc * y)
        out[:, :, c] = (obj * sens + 0.2 * z[:, :, c]) * 1e-3
    return np.ascontiguousarray(out.astype(np.complex64))
