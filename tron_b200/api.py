"""ctypes binding of libtron_b200.so -- the host-side mirror of TRON's interface.

Everything here goes through the C ABI declared in include/tron.h; torch is
only used by callers for device memory, streams and torch.distributed.  There
is no CPU fallback: if the shared library is missing or no CUDA device is
present the calls raise.

Reference interface mirrored (file:line in /root/reference/src):
  tron.cu:822-874   command-line flags            -> Config fields
  tron.cu:905-961   geometry from dims + flags    -> geometry()
  tron.cu:726-786   recon_radial2d(h_out, h_in)   -> Plan.recon_host()
  ra.cu:87-174      ra_read / ra_write            -> ra_read() / ra_write()
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtron_b200.so")
CLI_PATH = os.path.join(HERE, "bin", "tron")

c64 = np.complex64


class TronError(RuntimeError):
    pass


class Config(C.Structure):
    """tron_config (include/tron.h)."""
    _fields_ = [("dims", C.c_uint64 * 5), ("adjoint", C.c_int), ("golden_angle", C.c_int),
                ("gridos", C.c_float), ("kernwidth", C.c_float), ("data_undersamp", C.c_float),
                ("prof_slide", C.c_int), ("skip_angles", C.c_int), ("niter", C.c_int), ("koosh", C.c_int),
                ("verbose", C.c_int), ("device", C.c_int),
                ("half_in", C.c_int), ("half_out", C.c_int), ("slice_begin", C.c_int), ("slice_end", C.c_int),
                ("coil_begin", C.c_int), ("coil_end", C.c_int), ("sos_partial", C.c_int),
                ("batch_slices", C.c_int), ("per_coil_out", C.c_int), ("coil_combine", C.c_int),
                ("walsh_npatch", C.c_int)]


class Geometry(C.Structure):
    """tron_geometry (include/tron.h)."""
    _fields_ = [("nc", C.c_int), ("nt", C.c_int), ("nro", C.c_int), ("npe1", C.c_int), ("npe2", C.c_int),
                ("npe1work", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nxos", C.c_int),
                ("nyos", C.c_int), ("prof_slide", C.c_int), ("slice_begin", C.c_int), ("slice_end", C.c_int),
                ("coil_begin", C.c_int), ("coil_end", C.c_int), ("out_dims", C.c_uint64 * 5),
                ("in_elems", C.c_uint64), ("out_elems", C.c_uint64), ("shard_in_offset", C.c_uint64),
                ("shard_in_elems", C.c_uint64), ("shard_out_offset", C.c_uint64), ("shard_out_elems", C.c_uint64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = [int(x) for x in v] if name == "out_dims" else int(v)
        return d


class RaStruct(C.Structure):
    """ra_t (include/ra.h; reference ra.h:38-48)."""
    _fields_ = [("flags", C.c_uint64), ("eltype", C.c_uint64), ("elbyte", C.c_uint64), ("size", C.c_uint64),
                ("ndims", C.c_uint64), ("dims", C.POINTER(C.c_uint64)), ("data", C.POINTER(C.c_uint8))]


# every symbol include/*.h declares (used by the load test)
EXPORTED_SYMBOLS = [
    # tron.h: plan API
    "tron_config_defaults", "tron_geometry_compute", "tron_plan_create", "tron_plan_destroy",
    "tron_plan_geometry", "tron_recon_host", "tron_recon_device", "tron_grid_device",
    "tron_grid_to_interleaved", "tron_degrid_device", "tron_plan_last_stage_ms", "tron_plan_last_launches", "tron_plan_batch_slices",
    "tron_plan_grid_debug", "tron_coilcombine_sos_device", "tron_coilcombine_walsh_device",
    "tron_last_error", "tron_version",
    # tron.h: coil-sharded root sum of squares (NCCL)
    "tron_comm_unique_id", "tron_comm_create", "tron_comm_create_all", "tron_comm_destroy", "tron_comm_rank",
    "tron_comm_size", "tron_coil_reduce", "tron_coil_reduce_all",
    # tron.h: legacy surface
    "tron_set_config", "tron_init", "tron_shutdown", "tron_nufft_adj_radial2d", "tron_nufft_radial2d",
    "recon_radial2d", "recon_radial_2d", "gridradial2d", "degridradial2d",
    "tron_launch_gridradial2d", "tron_launch_degridradial2d", "tron_cgnr_radial2d", "copy", "Caxpy",
    "tron_launch_Caxpy",
    # ra.h
    "ra_read", "ra_write", "ra_free", "ra_query", "ra_reshape", "ra_convert", "ra_squash", "ra_diff",
    "ra_read_header", "ra_read_pinned", "ra_header_bytes",
    # float16.h
    "tron_floatbits_to_halfbits", "tron_doublebits_to_halfbits", "tron_halfbits_to_floatbits",
    "tron_halfbits_to_doublebits", "tron_float_to_half_array", "tron_half_to_float_array",
]

_lib = None


def load_library(path=None):
    """dlopen libtron_b200.so (building it first if nvcc is available and it is stale)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise TronError("%s not found: run `python -m tron_b200.build` (there is no CPU fallback)" % p)
    L = C.CDLL(p)
    L.tron_last_error.restype = C.c_char_p
    L.tron_geometry_compute.argtypes = [C.POINTER(Config), C.POINTER(Geometry)]
    L.tron_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
    L.tron_plan_destroy.argtypes = [C.c_void_p]
    L.tron_plan_geometry.argtypes = [C.c_void_p, C.POINTER(Geometry)]
    L.tron_recon_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.tron_recon_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tron_grid_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tron_grid_to_interleaved.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.tron_degrid_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tron_coilcombine_sos_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tron_coilcombine_walsh_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tron_launch_Caxpy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_size_t, C.c_int, C.c_int,
                                    C.c_void_p]
    L.tron_cgnr_radial2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.tron_cgnr_radial2d.restype = None
    L.tron_plan_last_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.tron_plan_last_launches.argtypes = [C.c_void_p]
    L.tron_plan_batch_slices.argtypes = [C.c_void_p]
    L.tron_plan_batch_slices.restype = C.c_int
    L.tron_set_config.argtypes = [C.POINTER(Config)]
    L.tron_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    L.tron_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
    L.tron_comm_create_all.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]
    L.tron_comm_destroy.argtypes = [C.c_void_p]
    L.tron_comm_rank.argtypes = [C.c_void_p]
    L.tron_comm_size.argtypes = [C.c_void_p]
    L.tron_coil_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.tron_coil_reduce_all.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t,
                                       C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.recon_radial2d.argtypes = [C.c_void_p, C.c_void_p]
    L.recon_radial2d.restype = None
    L.tron_launch_gridradial2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tron_launch_degridradial2d.argtypes = L.tron_launch_gridradial2d.argtypes
    L.ra_read.argtypes = [C.POINTER(RaStruct), C.c_char_p]
    L.ra_read_header.argtypes = [C.POINTER(RaStruct), C.c_char_p]
    L.ra_write.argtypes = [C.POINTER(RaStruct), C.c_char_p]
    L.ra_free.argtypes = [C.POINTER(RaStruct)]
    L.ra_free.restype = None
    L.ra_convert.argtypes = [C.POINTER(RaStruct), C.c_uint64, C.c_uint64]
    L.ra_convert.restype = None
    L.ra_reshape.argtypes = [C.POINTER(RaStruct), C.POINTER(C.c_uint64), C.c_uint64]
    L.ra_squash.argtypes = [C.POINTER(RaStruct)]
    L.ra_diff.argtypes = [C.POINTER(RaStruct), C.POINTER(RaStruct)]
    L.ra_header_bytes.argtypes = [C.POINTER(RaStruct)]
    L.ra_header_bytes.restype = C.c_uint64
    L.tron_floatbits_to_halfbits.argtypes = [C.c_uint32]
    L.tron_floatbits_to_halfbits.restype = C.c_uint16
    L.tron_halfbits_to_floatbits.argtypes = [C.c_uint16]
    L.tron_halfbits_to_floatbits.restype = C.c_uint32
    L.tron_doublebits_to_halfbits.argtypes = [C.c_uint64]
    L.tron_doublebits_to_halfbits.restype = C.c_uint16
    L.tron_halfbits_to_doublebits.argtypes = [C.c_uint16]
    L.tron_halfbits_to_doublebits.restype = C.c_uint64
    L.tron_float_to_half_array.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.tron_float_to_half_array.restype = None
    L.tron_half_to_float_array.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.tron_half_to_float_array.restype = None
    if path is None:
        _lib = L
    return L


def _check(rc, lib):
    if rc != 0:
        raise TronError("libtron_b200 error %d: %s" % (rc, lib.tron_last_error().decode()))


def make_config(dims, adjoint, golden=False, gridos=2.0, kernwidth=2.0, undersamp=1.0, prof_slide=0,
                skip_angles=0, device=-1, half_in=False, half_out=False, slices=None, coils=None,
                sos_partial=False, batch_slices=0, per_coil_out=False, niter=0, koosh=False,
                coil_combine=0, walsh_npatch=1):
    """The flags of the `tron` command line (tron.cu:822-874) as a tron_config."""
    lib = load_library()
    cfg = Config()
    lib.tron_config_defaults(C.byref(cfg))
    for i, d in enumerate(dims):
        cfg.dims[i] = int(d)
    cfg.adjoint = int(bool(adjoint)); cfg.golden_angle = int(bool(golden))
    cfg.gridos = gridos; cfg.kernwidth = kernwidth; cfg.data_undersamp = undersamp
    cfg.prof_slide = prof_slide; cfg.skip_angles = skip_angles; cfg.device = device
    cfg.half_in = int(bool(half_in)); cfg.half_out = int(bool(half_out))
    if slices is not None:
        cfg.slice_begin, cfg.slice_end = int(slices[0]), int(slices[1])
    if coils is not None:
        cfg.coil_begin, cfg.coil_end = int(coils[0]), int(coils[1])
    cfg.sos_partial = int(bool(sos_partial)); cfg.batch_slices = int(batch_slices)
    cfg.per_coil_out = int(bool(per_coil_out)); cfg.niter = int(niter); cfg.koosh = int(bool(koosh))
    cfg.coil_combine = int(coil_combine); cfg.walsh_npatch = int(walsh_npatch)
    return cfg


def geometry(cfg):
    """Host-only geometry derivation (no GPU needed)."""
    lib = load_library()
    g = Geometry()
    _check(lib.tron_geometry_compute(C.byref(cfg), C.byref(g)), lib)
    return g


def shard_slices(nz, rank, world):
    """Contiguous slice range of `rank` when nz slices are split over `world` ranks."""
    base, rem = divmod(nz, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Plan:
    """A tron_plan: geometry, tables, work buffers and streams on one GPU."""

    def __init__(self, cfg):
        self.lib = load_library()
        self.cfg = cfg
        h = C.c_void_p()
        _check(self.lib.tron_plan_create(C.byref(h), C.byref(cfg)), self.lib)
        self.handle = h
        self.geom = Geometry()
        _check(self.lib.tron_plan_geometry(self.handle, C.byref(self.geom)), self.lib)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.tron_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # element sizes of the boundary buffers
    @property
    def in_itemsize(self):
        return 4 if self.cfg.half_in else 8

    @property
    def out_itemsize(self):
        if self.cfg.adjoint and self.cfg.sos_partial and self.geom.nc > 1:
            return 4
        return 4 if self.cfg.half_out else 8

    def out_array(self):
        """A host array for this plan's shard of the output."""
        n = int(self.geom.shard_out_elems)
        if self.out_itemsize == 4 and self.cfg.sos_partial and self.cfg.adjoint:
            return np.zeros(n, dtype=np.float32)
        if self.cfg.half_out:
            return np.zeros((n, 2), dtype=np.float16)
        return np.zeros(n, dtype=c64)

    def recon_host(self, h_in, h_out=None):
        """Whole job on host arrays (numpy); h_in is this plan's shard of the input."""
        h_in = np.ascontiguousarray(h_in)
        if h_in.nbytes < int(self.geom.shard_in_elems) * self.in_itemsize:
            raise TronError("input holds %d bytes, the plan's shard needs %d"
                            % (h_in.nbytes, int(self.geom.shard_in_elems) * self.in_itemsize))
        if h_out is None:
            h_out = self.out_array()
        _check(self.lib.tron_recon_host(self.handle, h_out.ctypes.data_as(C.c_void_p),
                                        h_in.ctypes.data_as(C.c_void_p)), self.lib)
        return h_out

    def recon_host_ptr(self, out_ptr, in_ptr):
        """Whole job on raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        _check(self.lib.tron_recon_host(self.handle, C.c_void_p(out_ptr), C.c_void_p(in_ptr)), self.lib)

    def recon_device(self, d_out_ptr, d_in_ptr, stream=0):
        """Whole job on device pointers, asynchronous on `stream` (0 = the CUDA default stream)."""
        _check(self.lib.tron_recon_device(self.handle, C.c_void_p(d_out_ptr), C.c_void_p(d_in_ptr),
                                          C.c_void_p(stream)), self.lib)

    def grid_device(self, d_grid_ptr, d_samples_ptr, z0, nslices, stream=0):
        _check(self.lib.tron_grid_device(self.handle, C.c_void_p(d_grid_ptr), C.c_void_p(d_samples_ptr),
                                         z0, nslices, C.c_void_p(stream)), self.lib)

    def grid_to_interleaved(self, d_dst_ptr, d_grid_ptr, nslices, stream=0):
        _check(self.lib.tron_grid_to_interleaved(self.handle, C.c_void_p(d_dst_ptr), C.c_void_p(d_grid_ptr),
                                                 nslices, C.c_void_p(stream)), self.lib)

    def degrid_device(self, d_samples_ptr, d_grid_ptr, stream=0):
        _check(self.lib.tron_degrid_device(self.handle, C.c_void_p(d_samples_ptr), C.c_void_p(d_grid_ptr),
                                           C.c_void_p(stream)), self.lib)

    def last_launches(self):
        return int(self.lib.tron_plan_last_launches(self.handle))

    def batch_slices(self):
        """Slices per launch of the device-resident pipeline."""
        return int(self.lib.tron_plan_batch_slices(self.handle))

    def last_stage_ms(self):
        """(gridding/degridding, FFT passes, other) device ms of the last recon_device call;
        only filled when the plan was created with TRON_STAGE_TIMING set (diagnostic mode)."""
        ms = (C.c_float * 3)()
        _check(self.lib.tron_plan_last_stage_ms(self.handle, ms), self.lib)
        return [float(x) for x in ms]


COMM_ID_BYTES = 128


def comm_unique_id():
    """128 opaque bytes rank 0 hands to every other rank (ncclGetUniqueId)."""
    L = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(L.tron_comm_unique_id(buf, COMM_ID_BYTES), L)
    return buf.raw


class Comm:
    """One rank's NCCL communicator for the coil-sharded sum of squares (include/tron.h: tron_comm_*)."""

    def __init__(self, unique_id, rank, nranks, device):
        self.lib = load_library()
        self.handle = C.c_void_p()
        _check(self.lib.tron_comm_create(C.byref(self.handle), unique_id, len(unique_id), rank, nranks, device), self.lib)
        self.rank, self.nranks = rank, nranks

    def coil_reduce(self, d_img_ptr, d_sos_ptr, npix, root=0, half_out=False, stream=0):
        """ncclReduce(sum) of float32[npix] partial sums of squares to `root`, then (sqrt, 0) pixels there."""
        _check(self.lib.tron_coil_reduce(self.handle, C.c_void_p(d_img_ptr), C.c_void_p(d_sos_ptr), npix, root,
                                         int(half_out), C.c_void_p(stream)), self.lib)

    def close(self):
        if self.handle:
            self.lib.tron_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def coilcombine_walsh_device(d_img_ptr, d_coil_ptr, nimg, nchan, npatch=1, nslices=1, stream=0):
    """coilcombinewalsh (tron.cu:270-302) on device pointers: [nslices][nimg][nimg][nchan] -> [nslices][nimg][nimg]."""
    lib = load_library()
    _check(lib.tron_coilcombine_walsh_device(C.c_void_p(d_img_ptr), C.c_void_p(d_coil_ptr), nimg, nchan, npatch,
                                             nslices, C.c_void_p(stream)), lib)


def coilcombine_sos_device(d_img_ptr, d_coil_ptr, nimg, nchan, nslices=1, stream=0):
    """coilcombinesos (tron.cu:255-268) on device pointers."""
    lib = load_library()
    _check(lib.tron_coilcombine_sos_device(C.c_void_p(d_img_ptr), C.c_void_p(d_coil_ptr), nimg, nchan, nslices,
                                           C.c_void_p(stream)), lib)


def recon_radial2d(h_in, dims, **flags):
    """One-shot equivalent of `tron [flags] in.ra out.ra` on arrays: returns (output, out_dims)."""
    cfg = make_config(dims, **flags)
    with Plan(cfg) as p:
        out = p.recon_host(h_in)
        return out, [int(x) for x in p.geom.out_dims]


# ---------------------------------------------------------------- RA files
def ra_write(path, array, dims=None, eltype=4):
    """Write a complex64 (eltype 4) or other numpy array as an RA file through the C library."""
    lib = load_library()
    a = np.ascontiguousarray(array)
    r = RaStruct()
    if dims is None:
        dims = list(a.shape[::-1])              # RA is column-major: first dim fastest
    d = (C.c_uint64 * len(dims))(*[int(x) for x in dims])
    r.flags = 0; r.eltype = eltype
    nel = int(np.prod(dims))
    r.elbyte = a.nbytes // nel
    r.size = a.nbytes; r.ndims = len(dims)
    r.dims = C.cast(d, C.POINTER(C.c_uint64))
    r.data = a.ctypes.data_as(C.POINTER(C.c_uint8))
    rc = lib.ra_write(C.byref(r), path.encode())
    if rc:
        raise TronError("ra_write(%s) failed: %d" % (path, rc))


def ra_read(path):
    """Read an RA file through the C library: returns (flat numpy array, dims, eltype, elbyte)."""
    lib = load_library()
    r = RaStruct()
    rc = lib.ra_read(C.byref(r), path.encode())
    if rc:
        raise TronError("ra_read(%s) failed: %d" % (path, rc))
    try:
        dims = [int(r.dims[i]) for i in range(r.ndims)]
        raw = np.ctypeslib.as_array(r.data, shape=(int(r.size),)).copy()
        eltype, elbyte = int(r.eltype), int(r.elbyte)
    finally:
        lib.ra_free(C.byref(r))
    table = {(4, 8): np.complex64, (4, 16): np.complex128, (3, 4): np.float32, (3, 8): np.float64,
             (3, 2): np.float16, (1, 1): np.int8, (1, 2): np.int16, (1, 4): np.int32, (1, 8): np.int64,
             (2, 1): np.uint8, (2, 2): np.uint16, (2, 4): np.uint32, (2, 8): np.uint64}
    if (eltype, elbyte) == (4, 4):
        arr = raw.view(np.float16).reshape(-1, 2)
    elif (eltype, elbyte) in table:
        arr = raw.view(table[(eltype, elbyte)])
    else:
        arr = raw
    return arr, dims, eltype, elbyte
