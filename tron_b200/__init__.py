"""tron_b200 -- B200-native radial NUFFT engine, drop-in for TRON's hot path.

The product is the C-ABI shared library tron_b200/lib/libtron_b200.so (CUDA,
sm_100a) and the `tron` command line built from tron_b200/csrc; this package
is the thin host-side binding used by the tests and bench.py.
"""
from .api import (Config, Geometry, Plan, TronError, geometry, load_library, make_config, ra_read,
                  ra_write, recon_radial2d, shard_slices, coilcombine_walsh_device, coilcombine_sos_device,
                  EXPORTED_SYMBOLS, LIB_PATH, CLI_PATH, Comm, comm_unique_id, COMM_ID_BYTES)

__all__ = ["Config", "Geometry", "Plan", "TronError", "geometry", "load_library", "make_config", "ra_read",
           "ra_write", "recon_radial2d", "shard_slices", "coilcombine_walsh_device", "coilcombine_sos_device",
           "EXPORTED_SYMBOLS", "LIB_PATH", "CLI_PATH", "Comm", "comm_unique_id", "COMM_ID_BYTES"]
