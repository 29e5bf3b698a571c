/*
 * tron_main.cu -- the `tron` command line, drop-in for the reference binary.
 *
 * Same getopt string, defaults, exit codes and output header as
 * /root/reference/src/tron.cu:790-995 (print_usage, main):
 *   tron [-3aGhv] [-B blocks] [-d prof_slide] [-g gpu] [-i niter] [-k width]
 *        [-o gridos] [-r nro] [-s skip_angles] [-T threads] [-u data_undersamp]
 *        <infile.ra> [outfile.ra]
 * no arguments, -h or an unknown flag print the usage to stderr and exit 1;
 * the default output file is img_tron.ra; -B / -T / -r are accepted and ignored
 * (-r is overwritten from the header in the reference too, tron.cu:858-860 vs
 * 909,945; -B/-T tune the reference's launch shape, which this engine picks
 * itself).
 *
 * Additions, none of which change an existing letter:
 *   -H   write the output with half-precision elements (fp16 storage);
 *        a half-precision input (eltype 4 / elbyte 4) is detected from the header.
 *   -w npatch  adaptive (Walsh) coil combine with a (2 npatch + 1)^2 patch instead of the root
 *        sum of squares: the call the reference keeps commented out at tron.cu:766 (npatch 1).
 *   -i niter   runs the CGNR iteration of cgnr.cu (the reference's own is self-declared broken).
 *   -v   additionally prints one machine-readable JSON timing line.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda_runtime.h>

#include "../../include/tron.h"

static void print_usage(void)
{
    fprintf(stderr, "Trajectory-optimized Non-uniform Fast Fourier Transform (B200 engine)\n");
    fprintf(stderr, "Usage: tron [-3aGhHv] [-B blocks] [-d prof_slide] [-g gpu] [-i niter] [-k width] [-o gridos] "
                    "[-r nro] [-s skip_angles] [-T threads] [-u data_undersamp] [-w npatch] <infile.ra> [outfile.ra]\n");
    fprintf(stderr, "\t-3\t\t\t3D koosh ball trajectory (not implemented)\n");
    fprintf(stderr, "\t-a\t\t\tadjoint operation\n");
    fprintf(stderr, "\t-B blocks\t\tnumber of GPU blocks (ignored)\n");
    fprintf(stderr, "\t-d prof_slide\t\tnumber of phase encodes to slide between slices for helical scans\n");
    fprintf(stderr, "\t-g n\t\t\tGPU device to use (default: 0)\n");
    fprintf(stderr, "\t-G\t\t\tgolden angle radial\n");
    fprintf(stderr, "\t-h\t\t\tshow this help\n");
    fprintf(stderr, "\t-H\t\t\twrite half-precision output\n");
    fprintf(stderr, "\t-i\t\t\tnumber of iterations (default: 0)\n");
    fprintf(stderr, "\t-k width\t\twidth of gridding kernel\n");
    fprintf(stderr, "\t-o gridos\t\tgrid oversampling factor\n");
    fprintf(stderr, "\t-r nro\t\t\tnumber of readout points (ignored)\n");
    fprintf(stderr, "\t-s skip_angles\t\tnumber of initial phase encodes to skip\n");
    fprintf(stderr, "\t-T threads\t\tnumber of GPU threads (ignored)\n");
    fprintf(stderr, "\t-u data_undersamp\tinput data undersampling factor\n");
    fprintf(stderr, "\t-v\t\t\tverbose output\n");
    fprintf(stderr, "\t-w npatch\t\tadaptive (Walsh) coil combine, patch half-width npatch\n");
}

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char *argv[])
{
    tron_config cfg;
    tron_config_defaults(&cfg);
    int c;
    opterr = 0;
    while ((c = getopt(argc, argv, "3aB:d:g:Ghi:k:o:r:s:T:u:vHw:")) != -1) {
        switch (c) {
        case '3': cfg.koosh = 1; break;
        case 'a': cfg.adjoint = 1; break;
        case 'B': break;
        case 'd': cfg.prof_slide = atoi(optarg); break;
        case 'g': cfg.device = atoi(optarg); break;
        case 'G': cfg.golden_angle = 1; break;
        case 'h': print_usage(); return 1;
        case 'H': cfg.half_out = 1; break;
        case 'i': cfg.niter = atoi(optarg); break;
        case 'k': cfg.kernwidth = (float)atof(optarg); break;
        case 'o': cfg.gridos = (float)atof(optarg); break;
        case 'u': cfg.data_undersamp = (float)atof(optarg); break;
        case 'r': break;
        case 's': cfg.skip_angles = atoi(optarg); break;
        case 'T': break;
        case 'v': cfg.verbose = 1; break;
        case 'w': cfg.coil_combine = 1; cfg.walsh_npatch = atoi(optarg); break;
        default: print_usage(); return 1;
        }
    }
    if (argc == optind) { print_usage(); return 1; }
    const char *infile = argv[optind];
    const char *outfile = optind + 1 < argc ? argv[optind + 1] : "img_tron.ra";
#define VPRINT if (cfg.verbose) printf

    if (cfg.device >= 0 && cudaSetDevice(cfg.device) != cudaSuccess) {
        fprintf(stderr, "tron: cannot select GPU %d\n", cfg.device);
        return 1;
    }
    VPRINT("Reading %s\n", infile);
    ra_t ra_in;
    if (ra_read_pinned(&ra_in, infile)) return 74;                      /* EX_IOERR, as ra.cu:56-84 */
    if (ra_in.flags & (RA_FLAG_BIG_ENDIAN | RA_FLAG_COMPRESSED)) { fprintf(stderr, "tron: big-endian or compressed RA payloads are not supported (flags 0x%llx)\n", (unsigned long long)ra_in.flags); return 65; }
    if (ra_in.ndims != 5) { fprintf(stderr, "tron: input must be 5-D [nc, nt, d2, d3, d4], got %llu dims\n", (unsigned long long)ra_in.ndims); return 65; }
    if (ra_in.eltype == RA_TYPE_COMPLEX && ra_in.elbyte == 4) cfg.half_in = 1;
    else if (!(ra_in.eltype == RA_TYPE_COMPLEX && ra_in.elbyte == 8)) {
        fprintf(stderr, "tron: input elements must be complex64 or complex-half (eltype %llu, elbyte %llu)\n",
                (unsigned long long)ra_in.eltype, (unsigned long long)ra_in.elbyte);
        return 65;
    }
    for (int i = 0; i < 5; ++i) cfg.dims[i] = ra_in.dims[i];
    VPRINT("indims = {%llu, %llu, %llu, %llu, %llu}\n", (unsigned long long)ra_in.dims[0], (unsigned long long)ra_in.dims[1],
           (unsigned long long)ra_in.dims[2], (unsigned long long)ra_in.dims[3], (unsigned long long)ra_in.dims[4]);
    VPRINT("WARNING: Assuming square Cartesian dimensions for now.\n");

    tron_plan *plan = NULL;
    double t_plan0 = now_s();
    if (tron_plan_create(&plan, &cfg)) { fprintf(stderr, "tron: %s\n", tron_last_error()); return 1; }
    tron_geometry g;
    tron_plan_geometry(plan, &g);
    const size_t out_el = cfg.half_out ? 4 : 8;
    if (ra_in.size < g.in_elems * (cfg.half_in ? 4 : 8)) { fprintf(stderr, "tron: payload smaller than the header's dims\n"); return 65; }

    void *h_out = NULL;
    if (cudaMallocHost(&h_out, g.out_elems * out_el) != cudaSuccess) { fprintf(stderr, "tron: cannot allocate pinned output\n"); return 1; }
    double t_plan1 = now_s();

    VPRINT("Running reconstruction ...\n ");
    VPRINT("nc=%d nro=%d npe1=%d npe1work=%d nx=%d nxos=%d nz=%d prof_slide=%d\n", g.nc, g.nro, g.npe1, g.npe1work,
           g.nx, g.nxos, g.nz, g.prof_slide);
    double t0 = now_s();
    if (tron_recon_host(plan, h_out, ra_in.data)) { fprintf(stderr, "tron: %s\n", tron_last_error()); return 1; }
    double t1 = now_s();
    VPRINT("Elapsed time: %.2f s\n", (t1 - t0) + (t_plan1 - t_plan0));
    if (cfg.verbose)
        printf("{\"tron_b200\": {\"plan_s\": %.6f, \"recon_s\": %.6f, \"slices\": %d, \"launches\": %d}}\n",
               t_plan1 - t_plan0, t1 - t0, g.nz, tron_plan_last_launches(plan));

    VPRINT("Saving result to %s\n", outfile);
    ra_t ra_out;
    memset(&ra_out, 0, sizeof ra_out);
    ra_out.ndims = 5;
    ra_out.dims = (uint64_t *)malloc(5 * sizeof(uint64_t));
    for (int i = 0; i < 5; ++i) ra_out.dims[i] = g.out_dims[i];
    /* the reference always writes dims[0] = 1, also in forward mode where the
     * payload holds nc channels (tron.cu:899); TRON_FIX_HEADER=1 writes nc */
    if (!cfg.adjoint && getenv("TRON_FIX_HEADER")) ra_out.dims[0] = (uint64_t)g.nc;
    ra_out.flags = 0; ra_out.eltype = RA_TYPE_COMPLEX; ra_out.elbyte = out_el;
    ra_out.size = g.out_elems * out_el;
    ra_out.data = (uint8_t *)h_out;
    int rc = ra_write(&ra_out, outfile);

    VPRINT("Cleaning up.\n");
    free(ra_out.dims);
    cudaFreeHost(h_out);
    ra_free(&ra_in);
    tron_plan_destroy(plan);
    return rc ? 74 : 0;
}
