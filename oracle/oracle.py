"""ctypes front-ends for the test oracle.  TEST INFRASTRUCTURE ONLY.

``Oracle``  -- the CPU restatement (oracle/tron_oracle.c), usable anywhere.
``RefLib``  -- the unmodified reference compiled in place (oracle/_ref), needs
               a GPU at run time; only the prebuilt .so is used off the build box.

Arrays follow the reference layout: complex64, channel fastest.  From numpy
that is a C-contiguous array with the channel axis LAST, e.g. samples of shape
(npe, nro, nchan) and grids/images of shape (n, n, nchan).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c64 = np.complex64


class _Cfg(C.Structure):
    _fields_ = [("dims", C.c_uint64 * 5), ("adjoint", C.c_int), ("golden_angle", C.c_int),
                ("gridos", C.c_float), ("kernwidth", C.c_float), ("data_undersamp", C.c_float),
                ("prof_slide", C.c_int), ("skip_angles", C.c_int),
                ("nc", C.c_int), ("nt", C.c_int), ("nro", C.c_int), ("npe1", C.c_int),
                ("npe2", C.c_int), ("npe1work", C.c_int),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nxos", C.c_int), ("nyos", C.c_int),
                ("out_dims", C.c_uint64 * 5), ("out_elems", C.c_uint64),
                ("niter", C.c_int), ("coil_combine", C.c_int), ("walsh_npatch", C.c_int)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a):
    a = np.ascontiguousarray(a, dtype=c64)
    return a


class Oracle:
    def __init__(self, path=None):
        if path is None:
            from . import build as _b
            path = _b.build_oracle()
        self.lib = L = C.CDLL(path)
        L.oracle_besseli0.restype = C.c_float
        L.oracle_besseli0.argtypes = [C.c_float]
        L.oracle_gridkernel.restype = C.c_float
        L.oracle_gridkernel.argtypes = [C.c_float, C.c_float]
        L.oracle_gridkernelhat.restype = C.c_float
        L.oracle_gridkernelhat.argtypes = [C.c_float, C.c_float]
        L.oracle_spoke_angle_grid.restype = C.c_float
        L.oracle_spoke_angle_grid.argtypes = [C.c_int] * 4
        L.oracle_spoke_angle_degrid.restype = C.c_float
        L.oracle_spoke_angle_degrid.argtypes = [C.c_int] * 4
        L.oracle_grid_hits.restype = C.c_long
        L.oracle_grid_hits.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.oracle_gridradial2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_float, C.c_int, C.c_int]
        L.oracle_degridradial2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_float, C.c_int, C.c_int]
        L.oracle_precompensate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_fft2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_fftshift.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_crop.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.oracle_pad.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.oracle_deapod.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.oracle_coilcombinesos.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_coilcombinewalsh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_nufft_adj_coils.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_cgnr_coils.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_geometry.argtypes = [C.POINTER(_Cfg)]
        L.oracle_recon_radial2d.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p]
        L.oracle_floatbits_to_halfbits.restype = C.c_uint16
        L.oracle_floatbits_to_halfbits.argtypes = [C.c_uint32]
        L.oracle_halfbits_to_floatbits.restype = C.c_uint32
        L.oracle_halfbits_to_floatbits.argtypes = [C.c_uint16]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_trig_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._trig = None

    # -- configuration ----------------------------------------------------
    def config(self, dims, adjoint, golden=False, gridos=2.0, kernwidth=2.0, undersamp=1.0,
               prof_slide=0, skip_angles=0, niter=0, coil_combine=0, walsh_npatch=1):
        cfg = _Cfg()
        self.lib.oracle_cfg_defaults(C.byref(cfg))
        for i, d in enumerate(dims):
            cfg.dims[i] = int(d)
        cfg.adjoint = int(bool(adjoint)); cfg.golden_angle = int(bool(golden))
        cfg.gridos = gridos; cfg.kernwidth = kernwidth; cfg.data_undersamp = undersamp
        cfg.prof_slide = prof_slide; cfg.skip_angles = skip_angles
        cfg.niter = int(niter); cfg.coil_combine = int(coil_combine); cfg.walsh_npatch = int(walsh_npatch)
        rc = self.lib.oracle_geometry(C.byref(cfg))
        if rc:
            raise ValueError("oracle_geometry rejected the configuration (%d)" % rc)
        return cfg

    def recon(self, cfg, h_in):
        h_in = _c(h_in)
        out = np.zeros(int(cfg.out_elems), dtype=c64)
        rc = self.lib.oracle_recon_radial2d(C.byref(cfg), _ptr(out), _ptr(h_in))
        if rc:
            raise ValueError("oracle_recon_radial2d failed (%d)" % rc)
        return out

    def set_trig_table(self, ct=None, st=None):
        """Install SFU sin/cos values of the spokes (index pe+skip for golden, pe for linear); None clears."""
        if ct is None:
            self._trig = None
            self.lib.oracle_set_trig_table(None, None, 0)
            return
        ct = np.ascontiguousarray(ct, dtype=np.float32); st = np.ascontiguousarray(st, dtype=np.float32)
        self._trig = (ct, st)                       # keep alive
        self.lib.oracle_set_trig_table(_ptr(ct), _ptr(st), int(ct.size))

    # -- kernels -----------------------------------------------------------
    def grid(self, samples, nxos, nchan, nro, npe, W=2.0, skip=0, golden=False):
        samples = _c(samples)
        out = np.zeros((nxos, nxos, nchan), dtype=c64)
        self.lib.oracle_gridradial2d(_ptr(out), _ptr(samples), nxos, nchan, nro, npe, W, skip, int(golden))
        return out

    def degrid(self, grid, n, nrep, nro, npe, W=2.0, skip=0, golden=False):
        grid = _c(grid)
        out = np.zeros((npe, nro, nrep), dtype=c64)
        self.lib.oracle_degridradial2d(_ptr(out), _ptr(grid), n, nrep, nro, npe, W, skip, int(golden))
        return out

    def grid_hits(self, nxos, nro, npe, W=2.0, skip=0, golden=False, maxhits=1 << 24):
        buf = np.zeros((maxhits, 4), dtype=np.int32)
        n = self.lib.oracle_grid_hits(_ptr(buf), maxhits, nxos, nro, npe, W, skip, int(golden))
        if n > maxhits:
            raise ValueError("hit buffer too small: %d" % n)
        return buf[:n]

    def precompensate(self, samples, nchan, nro, npe):
        s = _c(samples).copy()
        self.lib.oracle_precompensate(_ptr(s), nchan, nro, npe)
        return s

    def fft2(self, a, n, nchan, sign):
        a = _c(a).copy()
        self.lib.oracle_fft2(_ptr(a), n, nchan, sign)
        return a

    def deapod(self, a, n, nrep, m, sigma):
        a = _c(a).copy()
        self.lib.oracle_deapod(_ptr(a), n, nrep, m, sigma)
        return a

    def pad(self, src, ndst, nsrc, nchan):
        out = np.zeros((ndst, ndst, nchan), dtype=c64)
        self.lib.oracle_pad(_ptr(out), ndst, _ptr(_c(src)), nsrc, nchan)
        return out

    def crop(self, src, ndst, nsrc, nchan):
        out = np.zeros((ndst, ndst, nchan), dtype=c64)
        self.lib.oracle_crop(_ptr(out), ndst, _ptr(_c(src)), nsrc, nchan)
        return out

    def fftshift(self, src, n, nchan, inverse):
        out = np.zeros((n, n, nchan), dtype=c64)
        self.lib.oracle_fftshift(_ptr(out), _ptr(_c(src)), n, nchan, int(inverse))
        return out

    def sos(self, coilimg, nimg, nchan):
        out = np.zeros((nimg, nimg), dtype=c64)
        self.lib.oracle_coilcombinesos(_ptr(out), _ptr(_c(coilimg)), nimg, nchan)
        return out

    def walsh(self, coilimg, nimg, nchan, npatch=1):
        out = np.zeros((nimg, nimg), dtype=c64)
        self.lib.oracle_coilcombinewalsh(_ptr(out), _ptr(_c(coilimg)), nimg, nchan, npatch)
        return out

    def adj_coils(self, cfg, samples, peoffset=0):
        """Per-coil images (nx, nx, nc) of one adjoint slice (tron.cu:623-637)."""
        out = np.zeros((cfg.nx, cfg.nx, cfg.nc), dtype=c64)
        self.lib.oracle_nufft_adj_coils(C.byref(cfg), _ptr(out), _ptr(_c(samples)), peoffset)
        return out

    def cgnr_coils(self, cfg, samples, peoffset=0, niter=1):
        """Per-coil images (nx, nx, nc) after `niter` CGNR iterations on one window."""
        out = np.zeros((cfg.nx, cfg.nx, cfg.nc), dtype=c64)
        self.lib.oracle_cgnr_coils(C.byref(cfg), _ptr(out), _ptr(_c(samples)), peoffset, niter)
        return out

    def float_to_half_bits(self, bits):
        f = self.lib.oracle_floatbits_to_halfbits
        return np.array([f(int(b)) for b in np.asarray(bits, dtype=np.uint32).ravel()], dtype=np.uint16)

    def half_to_float_bits(self, bits):
        f = self.lib.oracle_halfbits_to_floatbits
        return np.array([f(int(b)) for b in np.asarray(bits, dtype=np.uint16).ravel()], dtype=np.uint32)

    def num_threads(self):
        return int(self.lib.oracle_num_threads())


class RefLib:
    """The unmodified reference (oracle/_ref/libtronref*.so).  GPU required."""

    def __init__(self, widened=False):
        name = "libtronref_mc64.so" if widened else "libtronref.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.isfile(path):
            raise FileNotFoundError(path + " (run oracle/build.py in the build container)")
        self.lib = L = C.CDLL(path)
        L.tronref_configure.restype = C.c_longlong
        L.tronref_configure.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
        L.tronref_recon.restype = C.c_double
        L.tronref_recon.argtypes = [C.c_void_p, C.c_void_p]
        L.tronref_grid.restype = C.c_float
        L.tronref_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.tronref_degrid.restype = C.c_float
        L.tronref_degrid.argtypes = L.tronref_grid.argtypes
        L.tronref_deapod.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        if hasattr(L, "tronref_walsh"):
            L.tronref_walsh.restype = C.c_float
            L.tronref_walsh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.tronref_adj_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        L.tronref_geometry.argtypes = [C.POINTER(C.c_int)]
        L.tronref_maxchan.restype = C.c_int
        L.tronref_host_alloc.restype = C.c_void_p
        L.tronref_host_alloc.argtypes = [C.c_size_t]
        L.tronref_host_free.argtypes = [C.c_void_p]
        L.tronref_spoke_cs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.tronref_floatbits_to_halfbits.restype = C.c_ushort
        L.tronref_floatbits_to_halfbits.argtypes = [C.c_uint]
        L.tronref_halfbits_to_floatbits.restype = C.c_uint
        L.tronref_halfbits_to_floatbits.argtypes = [C.c_ushort]
        self.maxchan = int(L.tronref_maxchan())

    def configure(self, dims, adjoint, golden=False, gridos=2.0, kernwidth=2.0, undersamp=1.0,
                  prof_slide=0, skip_angles=0, verbose=False):
        d = (C.c_ulonglong * 5)(*[int(x) for x in dims])
        od = (C.c_ulonglong * 5)()
        n = self.lib.tronref_configure(d, int(adjoint), int(golden), gridos, kernwidth, undersamp,
                                       prof_slide, skip_angles, int(verbose), od)
        self.out_elems = int(n)
        self.out_dims = [int(x) for x in od]
        g = (C.c_int * 12)()
        self.lib.tronref_geometry(g)
        keys = ("nc", "nt", "nro", "npe1", "npe2", "npe1work", "nx", "ny", "nz", "nxos", "nyos", "prof_slide")
        self.geom = dict(zip(keys, [int(x) for x in g]))
        if self.geom["nc"] * self.geom["nt"] > self.maxchan and adjoint:
            raise ValueError("reference build supports at most %d channels (tron.h:51)" % self.maxchan)
        return self.geom

    def recon(self, h_in, return_seconds=False):
        h_in = _c(h_in)
        out = np.zeros(self.out_elems, dtype=c64)
        sec = self.lib.tronref_recon(_ptr(out), _ptr(h_in))
        return (out, sec) if return_seconds else out

    def grid(self, samples, nxos, nchan, nro, npe, W=2.0, gridos=2.0, skip=0, golden=False, reps=0):
        if nchan > self.maxchan:
            raise ValueError("nchan > MAXCHAN")
        samples = _c(samples)
        out = np.zeros((nxos, nxos, nchan), dtype=c64)
        ms = self.lib.tronref_grid(_ptr(out), _ptr(samples), nxos, nchan, nro, npe, W, gridos, skip,
                                   int(golden), reps)
        return (out, ms) if reps else out

    def degrid(self, grid, n, nrep, nro, npe, W=2.0, gridos=2.0, skip=0, golden=False, reps=0):
        grid = _c(grid)
        out = np.zeros((npe, nro, nrep), dtype=c64)
        ms = self.lib.tronref_degrid(_ptr(out), _ptr(grid), n, nrep, nro, npe, W, gridos, skip,
                                     int(golden), reps)
        return (out, ms) if reps else out

    def deapod(self, a, n, nrep, m, sigma):
        a = _c(a).copy()
        self.lib.tronref_deapod(_ptr(a), n, nrep, m, sigma)
        return a

    def walsh(self, coilimg, nimg, nchan, npatch=1, reps=0):
        """The reference's coilcombinewalsh kernel (tron.cu:270-302) on (nimg, nimg, nchan) coil images.
        It zeroes NCHAN^2 = 36 matrix entries whatever nchan is: defined for nchan <= 6 only."""
        if nchan > 6:
            raise ValueError("coilcombinewalsh is undefined for nchan > NCHAN = 6 (tron.cu:282)")
        out = np.zeros((nimg, nimg), dtype=c64)
        ms = self.lib.tronref_walsh(_ptr(out), _ptr(_c(coilimg)), nimg, nchan, npatch, reps)
        return (out, ms) if reps else out

    def spoke_cs(self, n, npe, skip, golden, degrid):
        """(ct, st) of spokes 0..n-1 as the reference kernels compute them (SFU)."""
        ct = np.zeros(n, dtype=np.float32); st = np.zeros(n, dtype=np.float32)
        rc = self.lib.tronref_spoke_cs(_ptr(ct), _ptr(st), n, npe, skip, int(golden), int(degrid))
        if rc:
            raise RuntimeError("tronref_spoke_cs failed")
        return ct, st

    def adj_stage_ms(self, samples, reps=3):
        ms = (C.c_float * 8)()
        self.lib.tronref_adj_stage_ms(_ptr(_c(samples)), ms, reps)
        return dict(zip(("precompensate", "gridradial2d", "fftshift_inv", "cufft", "fftshift_fwd",
                         "crop", "deapod", "coilcombinesos"), [float(x) for x in ms]))
