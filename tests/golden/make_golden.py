"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref) on a GPU.

Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy the .npz files into tests/golden/ and commit them.  Inputs are not
stored: they are regenerated from the seeded Philox stream in tests/util.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.oracle import RefLib  # noqa: E402
from util import PARITY_CASES, WALSH_CASES, case_input, synth_complex, walsh_input  # noqa: E402


def walsh_vectors(ref, outdir):
    """coilcombinewalsh (tron.cu:270-302) alone: its call site (tron.cu:766) is commented out in the
    reference, so the kernel is launched directly by the harness with the reference launch shape."""
    out = {}
    for nimg, nc, npatch in WALSH_CASES:
        out["walsh_%d_%d_%d" % (nimg, nc, npatch)] = ref.walsh(walsh_input(nimg, nc), nimg, nc, npatch)
    np.savez_compressed(os.path.join(outdir, "walsh.npz"), **out)
    print("walsh ok", sorted(out))


def main(outdir, only=None):
    os.makedirs(outdir, exist_ok=True)
    ref = RefLib()
    if only == "walsh":
        return walsh_vectors(ref, outdir)
    for name, (dims, flags) in sorted(PARITY_CASES.items()):
        ref.configure(dims, flags.get("adjoint", False), golden=flags.get("golden", False),
                      gridos=flags.get("gridos", 2.0), kernwidth=flags.get("kernwidth", 2.0),
                      undersamp=flags.get("undersamp", 1.0), prof_slide=flags.get("prof_slide", 0),
                      skip_angles=flags.get("skip_angles", 0))
        out = ref.recon(case_input(name))
        adj, gold = bool(flags.get("adjoint", False)), bool(flags.get("golden", False))
        # SFU sin/cos of every spoke the run touches (index pe+skip for golden, pe for linear)
        ntab = (flags.get("skip_angles", 0) + ref.geom["npe1"]) if gold else ref.geom["npe1work"]
        ct, st = ref.spoke_cs(ntab, ref.geom["npe1work"], 0, gold, not adj)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), out=out, ct=ct, st=st,
                            out_dims=np.array(ref.out_dims, dtype=np.int64),
                            geom=np.array([ref.geom[k] for k in sorted(ref.geom)], dtype=np.int64),
                            geom_keys=np.array(sorted(ref.geom)))
        print(name, ref.geom, float(np.abs(out).max()))
    # stage-level vectors: gridradial2d and degridradial2d alone, deapodkernel alone
    s = synth_complex((24, 64, 2), stream=101)                 # npe=24, nro=64, nchan=2
    g = ref.grid(s, 64, 2, 64, 24, W=2.0, gridos=2.0, skip=4, golden=True)
    gl = ref.grid(s, 64, 2, 64, 24, W=2.0, gridos=2.0, skip=0, golden=False)
    u = synth_complex((64, 64, 2), stream=102)
    d = ref.degrid(u, 64, 2, 64, 20, W=2.0, gridos=2.0, skip=2, golden=True)
    dl = ref.degrid(u, 64, 2, 64, 20, W=2.0, gridos=2.0, skip=0, golden=False)
    ones = np.ones((32, 32, 1), dtype=np.complex64)
    da = ref.deapod(ones, 32, 1, 2.0, 2.0)
    df = ref.deapod(np.ones((64, 64, 1), dtype=np.complex64), 64, 1, 2.0, 1.0)
    cs = {}
    for key, (n, npe, gold, deg) in dict(grid_golden=(28, 24, 1, 0), grid_linear=(24, 24, 0, 0),
                                         degrid_golden=(22, 20, 1, 1), degrid_linear=(20, 20, 0, 1)).items():
        cs[key + "_ct"], cs[key + "_st"] = ref.spoke_cs(n, npe, 0, gold, deg)
    np.savez_compressed(os.path.join(outdir, "stages.npz"), grid_golden=g, grid_linear=gl,
                        degrid_golden=d, degrid_linear=dl, deapod_adj=da, deapod_fwd=df, **cs)
    print("stages ok")
    walsh_vectors(ref, outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden", sys.argv[2] if len(sys.argv) > 2 else None)
