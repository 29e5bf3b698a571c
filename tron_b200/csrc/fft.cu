/*
 * fft.cu -- batched 2-D FFT with the surrounding passes fused in.
 *
 * Replaces, for the adjoint (tron.cu:631-635 and 764):
 *     fftshift(INV) -> cufftExecC2C(INVERSE) -> fftshift(FWD) -> crop -> deapodkernel
 *     -> coilcombinesos
 * and for the forward direction (tron.cu:642-646):
 *     pad -> deapodkernel(n = nxos, sigma = 1) -> fftshift(FWD) -> cufftExecC2C(FORWARD)
 *     -> fftshift(INV)
 * (/root/reference/src/tron.cu:161-178, 205-220, 255-268, 390-457).
 *
 * Both directions are two passes of 1-D transforms over lines that are
 * contiguous in memory; each pass writes its result transposed so that the
 * next pass (or the consumer) again reads contiguous lines:
 *
 *   adjoint  A: grid[ch][y][:]  --FFT_x-->  keep the nx centre frequencies  --> tmp[ch][b][y]
 *            B: tmp[ch][b][:]   --FFT_y-->  keep nx, deapodise, |.|^2 over coils --> image[a][b]
 *   forward  A: image[a][:][ch] (pad + deapodise on load) --FFT-->  tmp[ch][c][a]
 *            B: tmp[ch][c][:]   (pad on load)             --FFT-->  grid[ch][r][c]
 *
 * The two fftshifts never move data: a circular shift by n/2 of the INPUT is a
 * (-1)^k modulation of the output, a shift of the OUTPUT is an index offset in
 * the store.  Crop = only the nx kept outputs are stored (the intermediate is
 * nx/nxos the size of the grid); pad = zeros are written to shared memory, not
 * read from HBM.  cuFFT conventions are kept: unnormalised, FORWARD e^{-i},
 * INVERSE e^{+i}.
 *
 * The line transform is a Stockham autosort FFT in shared memory, radices
 * 8/4/2 plus 3 and 5 (nxos = 384 occurs with -o 1.5), twiddles from a
 * plan-time table computed in double precision.
 */
#include "tron_internal.h"
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace tronb {

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
/* complex add/sub as one packed FP32x2 instruction (FADD2 on sm_100a) */
__device__ __forceinline__ float2 cadd(float2 a, float2 b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
/* multiply by s*i, s = +-1 */
__device__ __forceinline__ float2 cmuli(float2 a, float s) { return make_float2(-s * a.y, s * a.x); }

/* shared-memory index padding: one extra element every 8 keeps the strided
 * Stockham accesses (stride = radix) off a single bank */
__device__ __forceinline__ int phys(int i) { return i + (i >> 4); }
static inline int phys_host(int i) { return i + (i >> 4); }

template <int R> struct Dft;
template <> struct Dft<2> {
    __device__ static void run(float2 *v, float) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};
template <> struct Dft<4> {
    __device__ static void run(float2 *v, float s) {
        float2 p = cadd(v[0], v[2]), q = csub(v[0], v[2]);
        float2 r = cadd(v[1], v[3]), t = cmuli(csub(v[1], v[3]), s);
        v[0] = cadd(p, r); v[2] = csub(p, r);
        v[1] = cadd(q, t); v[3] = csub(q, t);
    }
};
template <> struct Dft<8> {
    __device__ static void run(float2 *v, float s) {
        float2 e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
        Dft<4>::run(e, s); Dft<4>::run(o, s);
        const float h = 0.70710678118654752440f;
        float2 w1 = make_float2(h, s * h), w3 = make_float2(-h, s * h);
        o[1] = cmul(o[1], w1); o[2] = cmuli(o[2], s); o[3] = cmul(o[3], w3);
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = cadd(e[k], o[k]); v[k + 4] = csub(e[k], o[k]); }
    }
};
template <> struct Dft<3> {
    __device__ static void run(float2 *v, float s) {
        const float c = -0.5f, sn = 0.86602540378443864676f;
        float2 t = cadd(v[1], v[2]), d = cmuli(csub(v[1], v[2]), s * sn);
        float2 m = make_float2(fmaf(c, t.x, v[0].x), fmaf(c, t.y, v[0].y));
        v[0] = cadd(v[0], t); v[1] = cadd(m, d); v[2] = csub(m, d);
    }
};
template <> struct Dft<5> {
    __device__ static void run(float2 *v, float s) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        float2 a = cadd(v[1], v[4]), b = cadd(v[2], v[3]);
        float2 c = csub(v[1], v[4]), d = csub(v[2], v[3]);
        float2 m1 = make_float2(v[0].x + c1 * a.x + c2 * b.x, v[0].y + c1 * a.y + c2 * b.y);
        float2 m2 = make_float2(v[0].x + c2 * a.x + c1 * b.x, v[0].y + c2 * a.y + c1 * b.y);
        float2 n1 = cmuli(make_float2(s1 * c.x + s2 * d.x, s1 * c.y + s2 * d.y), s);
        float2 n2 = cmuli(make_float2(s2 * c.x - s1 * d.x, s2 * c.y - s1 * d.y), s);
        v[0] = cadd(v[0], cadd(a, b));
        v[1] = cadd(m1, n1); v[4] = csub(m1, n1);
        v[2] = cadd(m2, n2); v[3] = csub(m2, n2);
    }
};

/* one Stockham stage of radix R over L lines held in shared memory */
template <int R>
__device__ __forceinline__ void stockham_stage(const float2 *src, float2 *dst, const float2 *tw,
                                               int n, int Ns, int L, int pitch, float sgn)
{
    const int nb = n / R;
    const int tstep = n / (Ns * R);
    for (int idx = threadIdx.x; idx < nb * L; idx += blockDim.x) {
        const int l = idx / nb, j = idx - l * nb;
        const int k = j % Ns;
        const float2 *s = src + l * pitch;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = s[phys(j + t * nb)];
        if (Ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) {
                float2 w = tw[t * k * tstep];
                w.y *= sgn;
                v[t] = cmul(v[t], w);
            }
        }
        Dft<R>::run(v, sgn);
        float2 *d = dst + l * pitch;
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) d[phys(j0 + t * Ns)] = v[t];
    }
}

struct Factors { int nfac; int fac[16]; };

/* transforms L lines in place of bufA/bufB ping-pong; returns the buffer holding the result */
__device__ float2 *fft_lines(float2 *bufA, float2 *bufB, const float2 *tw, int n, int L, int pitch,
                             const Factors &f, float sgn)
{
    int Ns = 1;
    float2 *src = bufA, *dst = bufB;
    for (int s = 0; s < f.nfac; ++s) {
        const int R = f.fac[s];
        switch (R) {
        case 8: stockham_stage<8>(src, dst, tw, n, Ns, L, pitch, sgn); break;
        case 4: stockham_stage<4>(src, dst, tw, n, Ns, L, pitch, sgn); break;
        case 2: stockham_stage<2>(src, dst, tw, n, Ns, L, pitch, sgn); break;
        case 3: stockham_stage<3>(src, dst, tw, n, Ns, L, pitch, sgn); break;
        default: stockham_stage<5>(src, dst, tw, n, Ns, L, pitch, sgn); break;
        }
        Ns *= R;
        __syncthreads();
        float2 *t = src; src = dst; dst = t;
    }
    return src;
}

struct PassGeom {
    int n, nkeep, L, pitch;
    Factors f;
};

__device__ __forceinline__ void load_twiddles(float2 *stw, const float2 *tw, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) stw[i] = tw[i];
}

/* ---------------- adjoint pass A: FFT along x, crop, transpose ---------------- */
__global__ void __launch_bounds__(256)
adj_pass_a_kernel(const float2 *__restrict__ grid, float2 *__restrict__ tmp, const float2 *__restrict__ tw,
                  const PassGeom p)
{
    extern __shared__ float2 smem[];
    const int n = p.n, L = p.L, pitch = p.pitch, nkeep = p.nkeep;
    float2 *bufA = smem, *bufB = smem + L * pitch, *stw = smem + 2 * L * pitch;
    const int y0 = blockIdx.x * L;
    const size_t plane = blockIdx.y;
    load_twiddles(stw, tw, n);
    const float2 *g = grid + plane * (size_t)n * n + (size_t)y0 * n;
    for (int idx = threadIdx.x; idx < L * n; idx += blockDim.x) {
        int l = idx / n, j = idx - l * n;
        bufA[l * pitch + phys(j)] = (y0 + l < n) ? g[idx] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2 *res = fft_lines(bufA, bufB, stw, n, L, pitch, p.f, +1.f);
    const int w = (n - nkeep) / 2, h = n / 2;
    float2 *out = tmp + plane * (size_t)nkeep * n + y0;
    for (int idx = threadIdx.x; idx < nkeep * L; idx += blockDim.x) {
        int b = idx / L, l = idx - b * L;
        int k = b + w - h; if (k < 0) k += n;
        float2 v = res[l * pitch + phys(k)];
        if (k & 1) { v.x = -v.x; v.y = -v.y; }
        if (y0 + l < n) out[(size_t)b * n + l] = v;
    }
}

/* ------- adjoint pass B: FFT along y, crop, deapodise, coil combine ------- */
#define PASSB_MAXO 16
__global__ void __launch_bounds__(256)
adj_pass_b_kernel(const float2 *__restrict__ tmp, void *__restrict__ outv, const float *__restrict__ deapod,
                  const float2 *__restrict__ tw, const PassGeom p, int nch, int nc_total, int ch0, int mode,
                  int half_out)
{
    extern __shared__ float2 smem[];
    const int n = p.n, L = p.L, pitch = p.pitch, nkeep = p.nkeep;
    float2 *bufA = smem, *bufB = smem + L * pitch, *stw = smem + 2 * L * pitch;
    const int b0 = blockIdx.x * L;
    const int slice = blockIdx.y;
    const int w = (n - nkeep) / 2, h = n / 2;
    load_twiddles(stw, tw, n);
    float acc[PASSB_MAXO];
#pragma unroll
    for (int o = 0; o < PASSB_MAXO; ++o) acc[o] = 0.f;
    const size_t img = (size_t)nkeep * nkeep;

    for (int ch = 0; ch < nch; ++ch) {
        const float2 *src = tmp + ((size_t)slice * nch + ch) * (size_t)nkeep * n + (size_t)b0 * n;
        for (int idx = threadIdx.x; idx < L * n; idx += blockDim.x) {
            int l = idx / n, j = idx - l * n;
            bufA[l * pitch + phys(j)] = (b0 + l < nkeep) ? src[idx] : make_float2(0.f, 0.f);
        }
        __syncthreads();
        float2 *res = fft_lines(bufA, bufB, stw, n, L, pitch, p.f, +1.f);
#pragma unroll
        for (int o = 0; o < PASSB_MAXO; ++o) {
            int idx = threadIdx.x + o * blockDim.x;
            if (idx < nkeep * L) {
                int a = idx / L, l = idx - a * L;
                if (b0 + l < nkeep) {
                    int k = a + w - h; if (k < 0) k += n;
                    float2 v = res[l * pitch + phys(k)];
                    float s = __ldg(deapod + (size_t)a * nkeep + b0 + l);
                    if (k & 1) s = -s;
                    v.x *= s; v.y *= s;
                    size_t pix = (size_t)a * nkeep + b0 + l;
                    if (mode == 0 || mode == 3) acc[o] += v.x * v.x + v.y * v.y;
                    else if (mode == 1) {
                        if (half_out) ((__half2 *)outv)[(size_t)slice * img + pix] = __float22half2_rn(v);
                        else ((float2 *)outv)[(size_t)slice * img + pix] = v;
                    } else {
                        size_t o2 = ((size_t)slice * img + pix) * nc_total + ch0 + ch;
                        if (half_out) ((__half2 *)outv)[o2] = __float22half2_rn(v);
                        else ((float2 *)outv)[o2] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (mode == 0 || mode == 3) {
#pragma unroll
        for (int o = 0; o < PASSB_MAXO; ++o) {
            int idx = threadIdx.x + o * blockDim.x;
            if (idx < nkeep * L) {
                int a = idx / L, l = idx - a * L;
                if (b0 + l < nkeep) {
                    size_t pix = (size_t)slice * img + (size_t)a * nkeep + b0 + l;
                    if (mode == 3) ((float *)outv)[pix] = acc[o];
                    else {
                        float2 v = make_float2(sqrtf(acc[o]), 0.f);      /* tron.cu:263-264 */
                        if (half_out) ((__half2 *)outv)[pix] = __float22half2_rn(v);
                        else ((float2 *)outv)[pix] = v;
                    }
                }
            }
        }
    }
}

/* ------- forward pass A: pad + deapodise on load, FFT along columns ------- */
__global__ void __launch_bounds__(256)
fwd_pass_a_kernel(const void *__restrict__ imgv, float2 *__restrict__ tmp, const float *__restrict__ deapod,
                  const float2 *__restrict__ tw, const PassGeom p, int nch, int nc_total, int ch0, int half_in)
{
    extern __shared__ float2 smem[];
    const int n = p.n, L = p.L, pitch = p.pitch, nx = p.nkeep;
    float2 *bufA = smem, *bufB = smem + L * pitch, *stw = smem + 2 * L * pitch;
    const int a0 = blockIdx.x * L;
    const int ch = blockIdx.y % nch;                     /* blockIdx.y = image * nch + channel */
    const size_t img0 = (size_t)(blockIdx.y / nch) * nx * nx * nc_total;
    const int w = (n - nx) / 2, h = n / 2;
    load_twiddles(stw, tw, n);
    for (int idx = threadIdx.x; idx < L * pitch; idx += blockDim.x) bufA[idx] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int idx = threadIdx.x; idx < L * nx; idx += blockDim.x) {
        int l = idx / nx, b = idx - l * nx, a = a0 + l;
        /* pad drops source row 0 and column 0 (tron.cu:449-450) */
        if (a >= 1 && a < nx && b >= 1) {
            size_t e = img0 + ((size_t)a * nx + b) * nc_total + ch0 + ch;
            float2 v = half_in ? __half22float2(((const __half2 *)imgv)[e]) : ((const float2 *)imgv)[e];
            float s = __ldg(deapod + (size_t)a * nx + b);
            bufA[l * pitch + phys(b + w)] = make_float2(v.x * s, v.y * s);
        }
    }
    __syncthreads();
    float2 *res = fft_lines(bufA, bufB, stw, n, L, pitch, p.f, -1.f);
    float2 *out = tmp + (size_t)blockIdx.y * n * nx + a0;
    for (int idx = threadIdx.x; idx < n * L; idx += blockDim.x) {
        int c = idx / L, l = idx - c * L;
        int k = c + h; if (k >= n) k -= n;
        float2 v = res[l * pitch + phys(k)];
        if (k & 1) { v.x = -v.x; v.y = -v.y; }
        if (a0 + l < nx) out[(size_t)c * nx + l] = v;
    }
}

/* ------- forward pass B: pad on load, FFT along rows, planar grid out ------- */
__global__ void __launch_bounds__(256)
fwd_pass_b_kernel(const float2 *__restrict__ tmp, float2 *__restrict__ grid, const float2 *__restrict__ tw,
                  const PassGeom p)
{
    extern __shared__ float2 smem[];
    const int n = p.n, L = p.L, pitch = p.pitch, nx = p.nkeep;
    float2 *bufA = smem, *bufB = smem + L * pitch, *stw = smem + 2 * L * pitch;
    const int c0 = blockIdx.x * L;
    const int ch = blockIdx.y;
    const int w = (n - nx) / 2, h = n / 2;
    load_twiddles(stw, tw, n);
    for (int idx = threadIdx.x; idx < L * pitch; idx += blockDim.x) bufA[idx] = make_float2(0.f, 0.f);
    __syncthreads();
    const float2 *src = tmp + ((size_t)ch * n + c0) * nx;
    for (int idx = threadIdx.x; idx < L * nx; idx += blockDim.x) {
        int l = idx / nx, a = idx - l * nx;
        if (c0 + l < n) bufA[l * pitch + phys(a + w)] = src[idx];
    }
    __syncthreads();
    float2 *res = fft_lines(bufA, bufB, stw, n, L, pitch, p.f, -1.f);
    float2 *out = grid + (size_t)ch * n * n + c0;
    for (int idx = threadIdx.x; idx < n * L; idx += blockDim.x) {
        int r = idx / L, l = idx - r * L;
        int k = r + h; if (k >= n) k -= n;
        float2 v = res[l * pitch + phys(k)];
        if (k & 1) { v.x = -v.x; v.y = -v.y; }
        if (c0 + l < n) out[(size_t)r * n + l] = v;
    }
}

/* ====================================================================== */
/* power-of-two fast path: radix-8 butterflies in registers               */
/* ====================================================================== */
/*
 * One line of N points is owned by N/8 threads; thread j holds elements
 * j + q*N/8 (q = 0..7) in registers.  Every Stockham stage is: butterflies in
 * registers -> scatter to shared memory (autosort index) -> barrier -> gather
 * j + q*N/8 again.  Two shared buffers alternate, so one barrier per stage.
 * All index arithmetic is compile-time shifts and masks.
 */
template <int N, int L> struct P2 {
    static constexpr int T = N / 8;                 /* threads per line */
    /* exchange buffers per line.  Two alternate, one barrier per stage -- except for the longest lines, where two
     * buffers of 4 lines fill the SM's shared memory with ONE block (32 of 64 warps; ncu on the 2048-point passes:
     * every unit below 65 %, barrier and scoreboard stalls).  There a single buffer and a second barrier per stage
     * let two blocks share an SM. */
    static constexpr int NBUF = N >= 2048 ? 1 : 2;
    /* line pitch: >= phys(N-1)+1 and == 16/L (mod 16), so that the transposed read of the
     * store phase (L lines x 16/L consecutive outputs per half-warp) hits 16 distinct bank pairs */
    static constexpr int BASE = N + N / 16;
    static constexpr int WANT = L >= 16 ? 1 : 16 / L;
    static constexpr int PITCH = BASE + ((WANT - BASE % 16) + 16) % 16 + (((WANT - BASE % 16) + 16) % 16 == 0 ? 16 : 0);
};

/* phys(base + off) - phys(base) for the offsets the stages use: compile-time constants because the
 * low four bits of the base are known (see the access patterns in p2_stage / p2_stages) */
template <int Ns> __device__ __forceinline__ constexpr int p2_woff(int t)
{
    return Ns % 16 == 0 ? t * (Ns + Ns / 16) : (Ns == 8 ? 8 * t + (t >> 1) : t);
}
template <int T> __device__ __forceinline__ constexpr int p2_roff(int q)
{
    return T % 16 == 0 ? q * (T + T / 16) : q * T + ((q * T) >> 4);       /* T = 8: j < 8, so no carry */
}

/* the threads of one line synchronise among themselves (named barrier), not with the whole CTA */
template <int N>
__device__ __forceinline__ void p2_line_sync(int l)
{
    constexpr int T = N / 8;
    if (T >= 32) {
        /* barrier ids as immediates (a line index is < 8 whenever T >= 32): with a register operand ptxas reserves
         * all 16 named barriers for the block, which caps the blocks per SM */
        switch (l) {
        case 0: asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"n"(T) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"n"(T) : "memory"); break;
        case 3: asm volatile("bar.sync 4, %0;" ::"n"(T) : "memory"); break;
        case 4: asm volatile("bar.sync 5, %0;" ::"n"(T) : "memory"); break;
        case 5: asm volatile("bar.sync 6, %0;" ::"n"(T) : "memory"); break;
        case 6: asm volatile("bar.sync 7, %0;" ::"n"(T) : "memory"); break;
        default: asm volatile("bar.sync 8, %0;" ::"n"(T) : "memory"); break;
        }
    } else __syncthreads();
}

template <int N, int Ns, int R, int SGN>
__device__ __forceinline__ void p2_stage(float2 (&v)[8], float2 *dst, const float2 *stw, int j)
{
    constexpr int T = N / 8, M = 8 / R, TSTEP = N / (Ns * R);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int b = j + m * T;
        const int k = b & (Ns - 1);
        float2 x[R];
#pragma unroll
        for (int t = 0; t < R; ++t) x[t] = v[m + t * M];
        if (Ns > 1) {
            /* w^t from one table read and a depth-3 product chain: the strided reads
             * stw[t*k*TSTEP] serialise on shared-memory banks */
            float2 w1 = stw[phys(k * TSTEP)];
            if (SGN < 0) w1.y = -w1.y;
            x[1] = cmul(x[1], w1);
            if constexpr (R > 2) {
                const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                x[2] = cmul(x[2], w2); x[3] = cmul(x[3], w3);
                if constexpr (R > 4) {
                    const float2 w4 = cmul(w2, w2);
                    x[4] = cmul(x[4], w4); x[5] = cmul(x[5], cmul(w4, w1));
                    x[6] = cmul(x[6], cmul(w3, w3)); x[7] = cmul(x[7], cmul(w4, w3));
                }
            }
        }
        Dft<R>::run(x, (float)SGN);
        float2 *d = dst + phys((b - k) * R + k);
#pragma unroll
        for (int t = 0; t < R; ++t) d[p2_woff<Ns>(t)] = x[t];
    }
}

template <int N, int Ns, int SGN, bool SINGLE = (N >= 2048)>
__device__ __forceinline__ void p2_stages(float2 (&v)[8], float2 *bufA, float2 *bufB, const float2 *stw, int j, int l)
{
    constexpr int REM = N / Ns;
    constexpr int R = REM >= 8 ? 8 : REM;
    p2_stage<N, Ns, R, SGN>(v, bufA, stw, j);
    if constexpr (Ns * R < N) {
        constexpr int T = N / 8;
        p2_line_sync<N>(l);
        const float2 *s = bufA + phys(j);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = s[p2_roff<T>(q)];
        if (SINGLE) p2_line_sync<N>(l);                   /* single buffer: everyone has read before the next stage writes */
        p2_stages<N, Ns * R, SGN, SINGLE>(v, bufB, bufA, stw, j, l);
    } else {
        __syncthreads();                     /* the store phase reads every line of the CTA */
    }
}

/* number of stages decides which buffer holds the natural-order result */
template <int N> __host__ __device__ constexpr int p2_nstages() { int s = 0, m = N; while (m > 1) { m = m >= 8 ? m / 8 : 1; ++s; } return s; }

/* Transform: registers in (element j + q*T of the line), result in shared memory, natural order.
 * Returns the line base inside the result buffer.  lineA/lineB are this line's two buffers. */
template <int N, int SGN, bool SINGLE = (N >= 2048)>
__device__ __forceinline__ float2 *p2_fft(float2 (&v)[8], float2 *lineA, float2 *lineB, const float2 *stw, int j, int l)
{
    p2_stages<N, 1, SGN, SINGLE>(v, lineA, lineB, stw, j, l);
    return (p2_nstages<N>() & 1) ? lineA : lineB;
}

template <int N, int L>
__global__ void __launch_bounds__(L *(N / 8), (N >= 2048 && L * (N / 8) <= 1024 ? 2 : 1))
p2_adj_pass_a(const float2 *__restrict__ grid, float2 *__restrict__ tmp, const float2 *__restrict__ tw, int nkeep,
              int zero_r2)
{
    extern __shared__ float2 smem[];
    constexpr int T = P2<N, L>::T, PITCH = P2<N, L>::PITCH;
    float2 *bufA = smem, *bufB = smem + (P2<N, L>::NBUF - 1) * L * PITCH, *stw = smem + P2<N, L>::NBUF * L * PITCH;
    const int l = threadIdx.x / T, j = threadIdx.x % T;
    const int y0 = blockIdx.x * L;
    const size_t plane = blockIdx.y;
    for (int i = threadIdx.x; i < N; i += L * T) stw[phys(i)] = tw[i];
    const float2 *g = grid + plane * (size_t)N * N + (size_t)(y0 + l) * N;
    float2 v[8];
    const int Yc = y0 + l - N / 2, lim = zero_r2 - Yc * Yc;       /* see p2w_adj_pass_a */
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int X = j + q * T - N / 2;
        v[q] = X * X <= lim ? g[j + q * T] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2 *res = p2_fft<N, +1>(v, bufA + l * PITCH, bufB + l * PITCH, stw, j, l) - l * PITCH;
    const int w = (N - nkeep) / 2, h = N / 2;
    float2 *out = tmp + plane * (size_t)nkeep * N + y0;
    for (int idx = threadIdx.x; idx < nkeep * L; idx += L * T) {
        const int b = idx / L, ll = idx % L;
        const int k = (b + w - h) & (N - 1);
        float2 val = res[ll * PITCH + phys(k)];
        if (k & 1) { val.x = -val.x; val.y = -val.y; }
        out[(size_t)b * N + ll] = val;
    }
}

template <int N, int L, int OUTS>
__global__ void __launch_bounds__(L *(N / 8), (L * (N / 8) <= 256 ? 3 : 1))
p2_adj_pass_b(const float2 *__restrict__ tmp, void *__restrict__ outv, const float *__restrict__ deapod,
              const float2 *__restrict__ tw, int nkeep, int nch, int nc_total, int ch0, int mode, int half_out)
{
    extern __shared__ float2 smem[];
    constexpr int T = P2<N, L>::T, PITCH = P2<N, L>::PITCH;
    float2 *bufA = smem, *bufB = smem + L * PITCH, *stw = smem + 2 * L * PITCH;   /* two buffers: 64 registers hold one block per SM anyway */
    const int l = threadIdx.x / T, j = threadIdx.x % T;
    const int b0 = blockIdx.x * L;
    const int slice = blockIdx.y;
    const int w = (N - nkeep) / 2, h = N / 2;
    for (int i = threadIdx.x; i < N; i += L * T) stw[phys(i)] = tw[i];
    float acc[OUTS];
#pragma unroll
    for (int o = 0; o < OUTS; ++o) acc[o] = 0.f;
    const size_t img = (size_t)nkeep * nkeep;
    const bool line_ok = b0 + l < nkeep;
    const size_t chan_stride = (size_t)nkeep * N;
    const float2 *src = tmp + ((size_t)slice * nch * nkeep + b0 + l) * (size_t)N + j;

    /* software pipeline over the coils: the next coil's line is in flight while this one is transformed */
    float2 v[8], nv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = line_ok ? src[q * T] : make_float2(0.f, 0.f);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
#pragma unroll
            for (int q = 0; q < 8; ++q) nv[q] = line_ok ? src[(size_t)(ch + 1) * chan_stride + q * T] : make_float2(0.f, 0.f);
        }
        __syncthreads();                     /* previous coil's epilogue reads are done */
        float2 *res = p2_fft<N, +1, false>(v, bufA + l * PITCH, bufB + l * PITCH, stw, j, l) - l * PITCH;
#pragma unroll
        for (int o = 0; o < OUTS; ++o) {
            const int idx = threadIdx.x + o * (L * T);
            const int a = idx / L, ll = idx % L;
            if (a < nkeep && b0 + ll < nkeep) {
                const int k = (a + w - h) & (N - 1);
                float2 val = res[ll * PITCH + phys(k)];
                float s = __ldg(deapod + (size_t)a * nkeep + b0 + ll);
                if (k & 1) s = -s;
                val.x *= s; val.y *= s;
                const size_t pix = (size_t)a * nkeep + b0 + ll;
                if (mode == 0 || mode == 3) acc[o] += val.x * val.x + val.y * val.y;
                else if (mode == 1) {
                    if (half_out) ((__half2 *)outv)[(size_t)slice * img + pix] = __float22half2_rn(val);
                    else ((float2 *)outv)[(size_t)slice * img + pix] = val;
                } else {
                    const size_t o2 = ((size_t)slice * img + pix) * nc_total + ch0 + ch;
                    if (half_out) ((__half2 *)outv)[o2] = __float22half2_rn(val);
                    else ((float2 *)outv)[o2] = val;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = nv[q];
    }
    if (mode == 0 || mode == 3) {
#pragma unroll
        for (int o = 0; o < OUTS; ++o) {
            const int idx = threadIdx.x + o * (L * T);
            const int a = idx / L, ll = idx % L;
            if (a < nkeep && b0 + ll < nkeep) {
                const size_t pix = (size_t)slice * img + (size_t)a * nkeep + b0 + ll;
                if (mode == 3) ((float *)outv)[pix] = acc[o];
                else {
                    float2 val = make_float2(sqrtf(acc[o]), 0.f);      /* tron.cu:263-264 */
                    if (half_out) ((__half2 *)outv)[pix] = __float22half2_rn(val);
                    else ((float2 *)outv)[pix] = val;
                }
            }
        }
    }
}

template <int N, int L>
__global__ void __launch_bounds__(L *(N / 8), (N >= 2048 && L * (N / 8) <= 1024 ? 2 : 1))
p2_fwd_pass_a(const void *__restrict__ imgv, float2 *__restrict__ tmp, const float *__restrict__ deapod,
              const float2 *__restrict__ tw, int nx, int nch, int nc_total, int ch0, int half_in, int chan_fastest)
{
    extern __shared__ float2 smem[];
    constexpr int T = P2<N, L>::T, PITCH = P2<N, L>::PITCH;
    float2 *bufA = smem, *bufB = smem + (P2<N, L>::NBUF - 1) * L * PITCH, *stw = smem + P2<N, L>::NBUF * L * PITCH;
    const int l = threadIdx.x / T, j = threadIdx.x % T;
    /* many channels: the channel index runs fastest over the blocks (gridDim.x = planes), so that the blocks which
     * pick their channel's 4 or 8 bytes out of the same 32-byte sectors of the channel-interleaved image run
     * together and share them in L2 (ncu on cfg5: 8.7 GB of DRAM reads for a 0.27 GB image with rows fastest) */
    const int brow = chan_fastest ? blockIdx.y : blockIdx.x, bplane = chan_fastest ? blockIdx.x : blockIdx.y;
    const int a = brow * L + l;
    const int ch = bplane % nch;                     /* bplane = image * nch + channel */
    const size_t img0 = (size_t)(bplane / nch) * nx * nx * nc_total;
    const int w = (N - nx) / 2, h = N / 2;
    for (int i = threadIdx.x; i < N; i += L * T) stw[phys(i)] = tw[i];
    float2 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int b = j + q * T - w;                 /* source column of padded column j + q*T */
        v[q] = make_float2(0.f, 0.f);
        if (a >= 1 && a < nx && b >= 1 && b < nx) {  /* pad drops row 0 and column 0, tron.cu:449-450 */
            const size_t e = img0 + ((size_t)a * nx + b) * nc_total + ch0 + ch;
            float2 x = half_in ? __half22float2(((const __half2 *)imgv)[e]) : ((const float2 *)imgv)[e];
            const float s = __ldg(deapod + (size_t)a * nx + b);
            v[q] = make_float2(x.x * s, x.y * s);
        }
    }
    __syncthreads();
    float2 *res = p2_fft<N, -1>(v, bufA + l * PITCH, bufB + l * PITCH, stw, j, l) - l * PITCH;
    const int a0 = brow * L;
    float2 *out = tmp + (size_t)bplane * N * nx + a0;
    for (int idx = threadIdx.x; idx < N * L; idx += L * T) {
        const int c = idx / L, ll = idx % L;
        const int k = (c + h) & (N - 1);
        float2 val = res[ll * PITCH + phys(k)];
        if (k & 1) { val.x = -val.x; val.y = -val.y; }
        if (a0 + ll < nx) out[(size_t)c * nx + ll] = val;
    }
}

template <int N, int L>
__global__ void __launch_bounds__(L *(N / 8), (N >= 2048 && L * (N / 8) <= 1024 ? 2 : 1))
p2_fwd_pass_b(const float2 *__restrict__ tmp, float2 *__restrict__ grid, const float2 *__restrict__ tw, int nx)
{
    extern __shared__ float2 smem[];
    constexpr int T = P2<N, L>::T, PITCH = P2<N, L>::PITCH;
    float2 *bufA = smem, *bufB = smem + (P2<N, L>::NBUF - 1) * L * PITCH, *stw = smem + P2<N, L>::NBUF * L * PITCH;
    const int l = threadIdx.x / T, j = threadIdx.x % T;
    const int c0 = blockIdx.x * L;
    const int ch = blockIdx.y;
    const int w = (N - nx) / 2, h = N / 2;
    for (int i = threadIdx.x; i < N; i += L * T) stw[phys(i)] = tw[i];
    const float2 *src = tmp + ((size_t)ch * N + c0 + l) * nx;
    float2 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int a = j + q * T - w;
        v[q] = (a >= 0 && a < nx) ? src[a] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2 *res = p2_fft<N, -1>(v, bufA + l * PITCH, bufB + l * PITCH, stw, j, l) - l * PITCH;
    float2 *out = grid + (size_t)ch * N * N + c0;
    for (int idx = threadIdx.x; idx < N * L; idx += L * T) {
        const int r = idx / L, ll = idx % L;
        const int k = (r + h) & (N - 1);
        float2 val = res[ll * PITCH + phys(k)];
        if (k & 1) { val.x = -val.x; val.y = -val.y; }
        out[(size_t)r * N + ll] = val;
    }
}

/* ====================================================================== */
/* two-stage path: N = R1 * R2, R1 elements per thread, ONE exchange       */
/* ====================================================================== */
/*
 * The radix-8 path above moves every element through shared memory three times for N = 512
 * and is bound by that traffic.  Here a line is owned by T = N/R1 = R2 threads (half a warp
 * for 512 = 32 x 16); thread j holds x[j + T q], q < R1:
 *   stage 1  Y_j[k1] = DFT_R1 over q, in registers; written to shared memory (16-byte stores);
 *   stage 2  for k1 = b: X[b + R1 k2] = DFT_R2 over j of w_N^(j b) Y_j[b], again in registers
 *            (R1/R2 butterflies per thread).
 * The threads of a line sit in one warp: the exchange needs __syncwarp() only.
 * Exchange layout: Y_j[t] at j*(R1+2) + t -- the stores of a quarter-warp and the loads of a
 * half-warp each touch every bank once.
 */
__device__ __forceinline__ float2 w32pow(int k, float s)      /* exp(s 2 pi i k / 32), k < 16 */
{
    const float c[16] = { 1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                          0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
                          0.f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
                          -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f };
    const float n[16] = { 0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                          0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                          1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
                          0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f };
    return make_float2(c[k], s * n[k]);
}
/* radix-2 decimation in time on top of the half-size transform */
template <int R> struct DftDit {
    __device__ static void run(float2 *v, float s) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
        Dft<R / 2>::run(e, s); Dft<R / 2>::run(o, s);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const float2 t = k == 0 ? o[0] : (k == R / 4 ? cmuli(o[k], s) : cmul(o[k], w32pow(k * (32 / R), s)));
            v[k] = cadd(e[k], t); v[k + R / 2] = csub(e[k], t);
        }
    }
};
template <> struct Dft<16> { __device__ static void run(float2 *v, float s) { DftDit<16>::run(v, s); } };
template <> struct Dft<32> { __device__ static void run(float2 *v, float s) { DftDit<32>::run(v, s); } };

template <int N, int R1> struct P2W {
    static constexpr int R2 = N / R1;
    static constexpr int T = R2;                    /* threads per line */
    static constexpr int M2 = R1 / R2;              /* stage-2 butterflies per thread */
    static constexpr int XP = R1 + 2;               /* exchange pitch of one stage-1 thread */
    static constexpr int LPX = T * XP;              /* exchange pitch of a line */
    static constexpr int THREADS = 128;             /* small CTAs: five fit the register file at 96 registers */
    static constexpr int L = THREADS / T;           /* lines per CTA */
    /* pitch of the staged kept outputs F[line][a]: the transposed read of a half-warp covers L lines x
     * 16/L consecutive outputs and must touch 16 distinct 8-byte banks: pitch == 16/L (mod 16) */
    __host__ __device__ static constexpr int fpitch(int nkeep)
    {
        constexpr int want = L >= 16 ? 1 : 16 / L;
        return nkeep + ((want - nkeep % 16) + 16) % 16;
    }
    static_assert(R1 >= R2 && R1 % R2 == 0 && T <= 32 && R1 <= 32, "unsupported split");
};

template <int N, int R1, int SGN>
__device__ __forceinline__ void p2w_stage1(float2 (&v)[R1], float2 *xline, int j)
{
    Dft<R1>::run(v, (float)SGN);
    float4 *d = reinterpret_cast<float4 *>(xline + j * P2W<N, R1>::XP);
#pragma unroll
    for (int m = 0; m < R1 / 2; ++m) d[m] = make_float4(v[2 * m].x, v[2 * m].y, v[2 * m + 1].x, v[2 * m + 1].y);
}

/* butterfly b of stage 2: a[k2] = X[b + R1 k2].  tw = exp(+2 pi i k / N) (global table) */
template <int N, int R1>
__device__ __forceinline__ void p2w_stage2_load(float2 (&a)[N / R1], const float2 *xline, int b)
{
    constexpr int R2 = N / R1, XP = P2W<N, R1>::XP;
#pragma unroll
    for (int t = 0; t < R2; ++t) a[t] = xline[t * XP + b];
}
template <int N, int R1, int SGN>
__device__ __forceinline__ void p2w_stage2(float2 (&a)[N / R1], int b, const float2 *__restrict__ tw)
{
    constexpr int R2 = N / R1;
    /* w^t, t < R2: the powers of two come from the table, the rest are products of depth <= 3 */
    float2 w[R2];
#pragma unroll
    for (int t = 1; t < R2; t <<= 1) {
        w[t] = __ldg(tw + b * t);
        if (SGN < 0) w[t].y = -w[t].y;
    }
#pragma unroll
    for (int t = 3; t < R2; ++t) {
        if ((t & (t - 1)) == 0) continue;
        int hi = 1; while (hi * 2 <= t) hi *= 2;      /* t = hi + lo, lo < hi */
        w[t] = cmul(w[hi], w[t - hi]);
    }
#pragma unroll
    for (int t = 1; t < R2; ++t) a[t] = cmul(a[t], w[t]);
    Dft<R2>::run(a, (float)SGN);
}

/* adjoint pass A: grid[plane][y][:] --FFT--> keep nkeep centre outputs --> tmp[plane][a][y].
 * The kept outputs are staged as F[line][a] (odd pitch) over the exchange buffer for the
 * transposed, coalesced store. */
/* LD = 0: plain loads; 1: streaming loads (read once, evict first) */
template <int N, int R1, int LD>
__device__ __forceinline__ void p2w_pass_a_body(const float2 *__restrict__ grid_plane, float2 *__restrict__ tmp_plane,
                                                const float2 *__restrict__ tw, int nkeep, int zero_r2, int y0,
                                                float2 *smem)
{
    using G = P2W<N, R1>;
    const int l = threadIdx.x / G::T, j = threadIdx.x % G::T;
    float2 *xline = smem + l * G::LPX;
    float2 *F = smem;
    const int PF = G::fpitch(nkeep);
    const float2 *g = grid_plane + (size_t)(y0 + l) * N + j;
    {
        /* cells with X^2 + Y^2 > zero_r2 hold no sample (annulus of tron.cu:498-502 beyond nxos/2-1+W):
         * the gridding kernel does not store them and they are not fetched -- a fifth of the grid */
        const int Y = y0 + l - N / 2, X0 = j - N / 2;
        const int lim = zero_r2 - Y * Y;
        float2 v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            const int X = X0 + q * G::T;
            v[q] = X * X <= lim ? (LD ? __ldcs(g + q * G::T) : g[q * G::T]) : make_float2(0.f, 0.f);
        }
        p2w_stage1<N, R1, +1>(v, xline, j);
    }
    __syncwarp();
    float2 a[G::M2][G::R2];
#pragma unroll
    for (int m = 0; m < G::M2; ++m) p2w_stage2_load<N, R1>(a[m], xline, j + m * G::T);
    __syncthreads();                                  /* exchange buffer consumed: F may overwrite it */
    const int shift = N / 2 - (N - nkeep) / 2;       /* output a of frequency k: a = (k + shift) mod N */
#pragma unroll
    for (int m = 0; m < G::M2; ++m) {
        const int b = j + m * G::T;
        p2w_stage2<N, R1, +1>(a[m], b, tw);
        const float sg = (b & 1) ? -1.f : 1.f;       /* input shift by N/2 = (-1)^k on the output */
#pragma unroll
        for (int t = 0; t < G::R2; ++t) {
            const int aa = (b + t * R1 + shift) & (N - 1);
            if (aa < nkeep) F[l * PF + aa] = make_float2(a[m][t].x * sg, a[m][t].y * sg);
        }
    }
    __syncthreads();
    float2 *out = tmp_plane + y0;
    for (int idx = threadIdx.x; idx < nkeep * G::L; idx += G::THREADS) {
        const int aa = idx / G::L, ll = idx % G::L;
        out[(size_t)aa * N + ll] = F[ll * PF + aa];
    }
}

template <int N, int R1>
__global__ void __launch_bounds__(P2W<N, R1>::THREADS, 5)
p2w_adj_pass_a(const float2 *__restrict__ grid, float2 *__restrict__ tmp, const float2 *__restrict__ tw, int nkeep,
               int zero_r2)
{
    extern __shared__ float2 smem[];
    const size_t plane = blockIdx.y;
    p2w_pass_a_body<N, R1, 0>(grid + plane * (size_t)N * N, tmp + plane * (size_t)nkeep * N, tw, nkeep, zero_r2,
                              blockIdx.x * P2W<N, R1>::L, smem);
}

/* adjoint pass B for the usual 2x oversampling (nkeep = N/2), coils combined by sum of squares:
 * tmp[slice][ch][b][:] --FFT--> keep N/2 centre outputs, |.|^2 summed over the coils in registers
 * (a thread owns the same outputs for every coil), deapodised and stored once at the end.
 * mode 0: image = (sqrt(sum), 0) (tron.cu:255-268); mode 3: the partial sum itself (coil shards). */
/* LD = 0: plain loads; 1: ld.global.cg (L2 only): the fused kernel re-uses the intermediate's addresses for
 * later slices, so a line of an earlier slice may still sit in this SM's (incoherent) L1 */
template <int N, int R1, int LD>
__device__ __forceinline__ void p2w_pass_b_sos_body(const float2 *__restrict__ tmp_slice, void *__restrict__ outv,
                                                    size_t out_slice, const float *__restrict__ deapod,
                                                    const float2 *__restrict__ tw, int nch, int mode, int half_out,
                                                    int b0, float2 *smem)
{
    using G = P2W<N, R1>;
    constexpr int nkeep = N / 2, KT = G::R2 / 4;      /* kept k2: [0, KT) and [R2 - KT, R2) */
    const int l = threadIdx.x / G::T, j = threadIdx.x % G::T;
    float2 *xline = smem + l * G::LPX;
    float acc[G::M2][2 * KT];
#pragma unroll
    for (int m = 0; m < G::M2; ++m)
#pragma unroll
        for (int u = 0; u < 2 * KT; ++u) acc[m][u] = 0.f;
    const float2 *src = tmp_slice + (size_t)(b0 + l) * (size_t)N + j;
    for (int ch = 0; ch < nch; ++ch) {
        {
            float2 v[R1];
#pragma unroll
            for (int q = 0; q < R1; ++q) v[q] = LD ? __ldcg(src + q * G::T) : src[q * G::T];
            p2w_stage1<N, R1, +1>(v, xline, j);
        }
        src += (size_t)nkeep * N;
        __syncwarp();
#pragma unroll
        for (int m = 0; m < G::M2; ++m) {
            const int b = j + m * G::T;
            float2 a[G::R2];
            p2w_stage2_load<N, R1>(a, xline, b);
            p2w_stage2<N, R1, +1>(a, b, tw);
#pragma unroll
            for (int u = 0; u < 2 * KT; ++u) {
                const int t = u < KT ? u : G::R2 - 2 * KT + u;
                acc[m][u] = fmaf(a[t].x, a[t].x, fmaf(a[t].y, a[t].y, acc[m][u]));
            }
        }
        __syncwarp();                                 /* the line's exchange buffer is free again */
    }
    /* stage the sums as S[line][a] for the transposed store; a = (k + N/4) mod N for the kept k */
    float *S = reinterpret_cast<float *>(smem);
    constexpr int PS = nkeep + 1;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < G::M2; ++m)
#pragma unroll
        for (int u = 0; u < 2 * KT; ++u) {
            const int t = u < KT ? u : G::R2 - 2 * KT + u;
            const int aa = (j + m * G::T + t * R1 + N / 4) & (N - 1);
            S[l * PS + aa] = acc[m][u];
        }
    __syncthreads();
    const size_t img = (size_t)nkeep * nkeep;
    for (int idx = threadIdx.x; idx < nkeep * G::L; idx += G::THREADS) {
        const int aa = idx / G::L, ll = idx % G::L;
        const size_t pix = (size_t)aa * nkeep + b0 + ll;
        const float d = __ldg(deapod + pix);
        const float sum = S[ll * PS + aa] * d * d;
        if (mode == 3) ((float *)outv)[out_slice * img + pix] = sum;
        else {
            const float2 val = make_float2(sqrtf(sum), 0.f);              /* tron.cu:263-264 */
            if (half_out) ((__half2 *)outv)[out_slice * img + pix] = __float22half2_rn(val);
            else ((float2 *)outv)[out_slice * img + pix] = val;
        }
    }
}

template <int N, int R1>
__global__ void __launch_bounds__(P2W<N, R1>::THREADS, 4)
p2w_adj_pass_b_sos(const float2 *__restrict__ tmp, void *__restrict__ outv, const float *__restrict__ deapod,
                   const float2 *__restrict__ tw, int nch, int mode, int half_out)
{
    extern __shared__ float2 smem[];
    const int slice = blockIdx.y;
    p2w_pass_b_sos_body<N, R1, 0>(tmp + (size_t)slice * nch * (N / 2) * (size_t)N, outv, (size_t)slice, deapod, tw, nch,
                                  mode, half_out, blockIdx.x * P2W<N, R1>::L, smem);
}

/* Both passes in ONE launch, so that the intermediate never has to come back from HBM.
 * Blocks are numbered slice by slice: the A blocks of slice s (nch planes x N/L row blocks), then the
 * B blocks of slice s-1 (nkeep/L line blocks); the B blocks of the last slice close the grid.  A B block
 * waits until the `ready` counter of its slice shows that every A block has published its rows
 * (__threadfence + atomicAdd); all the blocks it waits for have lower indices, i.e. they were
 * dispatched before it, so the wait cannot deadlock.  The intermediate is a ring of `ring` slices:
 * slice s re-uses the slot of slice s - ring once that slice's B blocks have counted themselves `done`,
 * so its 6.3 MB (cfg2) are overwritten while still in the 126 MB L2 and pass B reads them from L2
 * (ld.global.cg: the SM's own L1 could hold the slot's previous contents).  Grid rows are fetched with
 * streaming loads so that they do not push the ring out. */
template <int N, int R1>
__global__ void __launch_bounds__(P2W<N, R1>::THREADS, 4)
p2w_adj_fused(const float2 *__restrict__ grid, float2 *__restrict__ tmp, void *__restrict__ outv,
              const float *__restrict__ deapod, const float2 *__restrict__ tw, int nch, int mode, int half_out,
              int zero_r2, int nslices, int ring, int *__restrict__ ready, int *__restrict__ done)
{
    extern __shared__ float2 smem[];
    using G = P2W<N, R1>;
    constexpr int nkeep = N / 2, NA = N / G::L, NB = nkeep / G::L;
    const int na = nch * NA, per = na + NB;
    const int id = blockIdx.x;
    int s = id / per, r = id - s * per;
    const size_t slot_elems = (size_t)nch * nkeep * N;
    if (s < nslices && r < na) {                          /* ---- pass A of slice s ---- */
        if (s >= ring) {                                  /* the slot's previous reader must be finished */
            if (threadIdx.x == 0) {
                while (atomicAdd(done + s - ring, 0) < NB) __nanosleep(64);
                __threadfence();
            }
            __syncthreads();
        }
        const int ch = r / NA, yb = r - ch * NA;
        p2w_pass_a_body<N, R1, 1>(grid + ((size_t)s * nch + ch) * (size_t)N * N,
                                  tmp + (size_t)(s % ring) * slot_elems + (size_t)ch * nkeep * N, tw, nkeep, zero_r2,
                                  yb * G::L, smem);
        __threadfence();                                  /* this thread's rows are visible device-wide ... */
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(ready + s, 1);    /* ... before the block counts itself */
        return;
    }
    /* ---- pass B of slice s - 1 (of the last slice for the trailing blocks) ---- */
    int bb;
    if (s < nslices) { if (s == 0) return; bb = r - na; s -= 1; }
    else { bb = id - nslices * per; s = nslices - 1; }
    if (threadIdx.x == 0) {
        while (atomicAdd(ready + s, 0) < na) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
    p2w_pass_b_sos_body<N, R1, 1>(tmp + (size_t)(s % ring) * slot_elems, outv, (size_t)s, deapod, tw, nch, mode,
                                  half_out, bb * G::L, smem);
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(done + s, 1); }
}

/* adjoint pass B, nkeep = N/2, coils kept apart: mode 1 (single channel, complex image) and mode 2
 * (per-coil images, channel interleaved -- the input of the Walsh combine and the vectors of CGNR).
 * Same line transform as p2w_adj_pass_b_sos; each coil's kept outputs are staged as F[line][a] in a
 * buffer of their own (the exchange buffer stays live for the next coil) and stored transposed. */
template <int N, int R1>
__global__ void __launch_bounds__(P2W<N, R1>::THREADS, 4)
p2w_adj_pass_b_coil(const float2 *__restrict__ tmp, void *__restrict__ outv, const float *__restrict__ deapod,
                    const float2 *__restrict__ tw, int nch, int nc_total, int ch0, int mode, int half_out)
{
    extern __shared__ float2 smem[];
    using G = P2W<N, R1>;
    constexpr int nkeep = N / 2, KT = G::R2 / 4;
    constexpr int PF = G::fpitch(nkeep);
    const int l = threadIdx.x / G::T, j = threadIdx.x % G::T;
    float2 *xline = smem + l * G::LPX;
    float2 *F = smem + G::L * G::LPX;
    const int b0 = blockIdx.x * G::L;
    const int slice = blockIdx.y;
    const size_t img = (size_t)nkeep * nkeep;
    const float2 *src = tmp + ((size_t)slice * nch * nkeep + b0 + l) * (size_t)N + j;
    for (int ch = 0; ch < nch; ++ch) {
        {
            float2 v[R1];
#pragma unroll
            for (int q = 0; q < R1; ++q) v[q] = src[q * G::T];
            p2w_stage1<N, R1, +1>(v, xline, j);
        }
        src += (size_t)nkeep * N;
        __syncwarp();
#pragma unroll
        for (int m = 0; m < G::M2; ++m) {
            const int b = j + m * G::T;
            float2 a[G::R2];
            p2w_stage2_load<N, R1>(a, xline, b);
            p2w_stage2<N, R1, +1>(a, b, tw);
            const float sg = (b & 1) ? -1.f : 1.f;    /* input shift by N/2 = (-1)^k on the output */
#pragma unroll
            for (int u = 0; u < 2 * KT; ++u) {
                const int t = u < KT ? u : G::R2 - 2 * KT + u;
                const int aa = (b + t * R1 + N / 4) & (N - 1);
                F[l * PF + aa] = make_float2(a[t].x * sg, a[t].y * sg);
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nkeep * G::L; idx += G::THREADS) {
            const int aa = idx / G::L, ll = idx % G::L;
            const size_t pix = (size_t)aa * nkeep + b0 + ll;
            const float d = __ldg(deapod + pix);
            float2 val = F[ll * PF + aa];
            val.x *= d; val.y *= d;
            const size_t o = mode == 1 ? (size_t)slice * img + pix : ((size_t)slice * img + pix) * nc_total + ch0 + ch;
            if (half_out) ((__half2 *)outv)[o] = __float22half2_rn(val);
            else ((float2 *)outv)[o] = val;
        }
        __syncthreads();                              /* F and the exchange buffers are free again */
    }
}

/* forward passes on the two-stage line transform (the radix-8 versions above move every element through shared
 * memory three times and sit at l1tex 74-83 %).  Both transform L lines per CTA and store all N outputs
 * transposed through the staged buffer F[line][output], like p2w_adj_pass_a.
 *   A: image rows (pad + deapodise on load, pad drops row 0 / column 0, tron.cu:449-450) -> tmp[plane][c][a]
 *   B: tmp[plane][c][:] (zero padded to N) -> grid[plane][r][c] */
template <int N, int R1>
__device__ __forceinline__ void p2w_fwd_finish(float2 (&a)[P2W<N, R1>::M2][P2W<N, R1>::R2], float2 *smem, int l, int j,
                                               const float2 *__restrict__ tw, float2 *__restrict__ out, int row_len,
                                               int valid_lines)
{
    using G = P2W<N, R1>;
    float2 *F = smem;
    constexpr int PF = G::fpitch(N);
    __syncthreads();                                  /* exchange buffer consumed: F may overwrite it */
#pragma unroll
    for (int m = 0; m < G::M2; ++m) {
        const int b = j + m * G::T;
        p2w_stage2<N, R1, -1>(a[m], b, tw);
        const float sg = (b & 1) ? -1.f : 1.f;       /* input shift by N/2 = (-1)^k on the output */
#pragma unroll
        for (int t = 0; t < G::R2; ++t) {
            const int c = (b + t * R1 + N / 2) & (N - 1);        /* output shift by N/2 */
            F[l * PF + c] = make_float2(a[m][t].x * sg, a[m][t].y * sg);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * G::L; idx += G::THREADS) {
        const int c = idx / G::L, ll = idx % G::L;
        if (ll < valid_lines) out[(size_t)c * row_len + ll] = F[ll * PF + c];
    }
}

template <int N, int R1>
__global__ void __launch_bounds__(P2W<N, R1>::THREADS, 5)
p2w_fwd_pass_a(const void *__restrict__ imgv, float2 *__restrict__ tmp, const float *__restrict__ deapod,
               const float2 *__restrict__ tw, int nx, int nch, int nc_total, int ch0, int half_in)
{
    extern __shared__ float2 smem[];
    using G = P2W<N, R1>;
    const int l = threadIdx.x / G::T, j = threadIdx.x % G::T;
    float2 *xline = smem + l * G::LPX;
    const int a0 = blockIdx.x * G::L, arow = a0 + l;
    const int ch = blockIdx.y % nch;                  /* blockIdx.y = image * nch + channel */
    const size_t img0 = (size_t)(blockIdx.y / nch) * nx * nx * nc_total;
    const int w = (N - nx) / 2;
    {
        float2 v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            const int b = j + q * G::T - w;           /* source column of padded column j + q*T */
            v[q] = make_float2(0.f, 0.f);
            if (arow >= 1 && arow < nx && b >= 1 && b < nx) {
                const size_t e = img0 + ((size_t)arow * nx + b) * nc_total + ch0 + ch;
                const float2 x = half_in ? __half22float2(((const __half2 *)imgv)[e]) : ((const float2 *)imgv)[e];
                const float sc = __ldg(deapod + (size_t)arow * nx + b);
                v[q] = make_float2(x.x * sc, x.y * sc);
            }
        }
        p2w_stage1<N, R1, -1>(v, xline, j);
    }
    __syncwarp();
    float2 a[G::M2][G::R2];
#pragma unroll
    for (int m = 0; m < G::M2; ++m) p2w_stage2_load<N, R1>(a[m], xline, j + m * G::T);
    p2w_fwd_finish<N, R1>(a, smem, l, j, tw, tmp + (size_t)blockIdx.y * N * nx + a0, nx, nx - a0);
}

template <int N, int R1>
__global__ void __launch_bounds__(P2W<N, R1>::THREADS, 5)
p2w_fwd_pass_b(const float2 *__restrict__ tmp, float2 *__restrict__ grid, const float2 *__restrict__ tw, int nx)
{
    extern __shared__ float2 smem[];
    using G = P2W<N, R1>;
    const int l = threadIdx.x / G::T, j = threadIdx.x % G::T;
    float2 *xline = smem + l * G::LPX;
    const int c0 = blockIdx.x * G::L;
    const int w = (N - nx) / 2;
    const float2 *src = tmp + ((size_t)blockIdx.y * N + c0 + l) * nx;
    {
        float2 v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            const int ai = j + q * G::T - w;
            v[q] = (ai >= 0 && ai < nx) ? src[ai] : make_float2(0.f, 0.f);
        }
        p2w_stage1<N, R1, -1>(v, xline, j);
    }
    __syncwarp();
    float2 a[G::M2][G::R2];
#pragma unroll
    for (int m = 0; m < G::M2; ++m) p2w_stage2_load<N, R1>(a[m], xline, j + m * G::T);
    p2w_fwd_finish<N, R1>(a, smem, l, j, tw, grid + (size_t)blockIdx.y * N * N + c0, N, G::L);
}

template <int N, int R1> struct P2WLaunch {
    using G = P2W<N, R1>;
    static size_t smem_a(int nkeep) { return (size_t)std::max(G::L * G::LPX, G::L * G::fpitch(nkeep)) * sizeof(float2); }
    static int prepare()
    {
        TRON_CUDA(cudaFuncSetAttribute(p2w_adj_pass_a<N, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a(N)));
        TRON_CUDA(cudaFuncSetAttribute(p2w_adj_pass_b_sos<N, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a(N / 2)));
        TRON_CUDA(cudaFuncSetAttribute(p2w_adj_pass_b_coil<N, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_coil()));
        TRON_CUDA(cudaFuncSetAttribute(p2w_adj_fused<N, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a(N)));
        TRON_CUDA(cudaFuncSetAttribute(p2w_fwd_pass_a<N, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a(N)));
        TRON_CUDA(cudaFuncSetAttribute(p2w_fwd_pass_b<N, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a(N)));
        return 0;
    }
    static int fwd(const FftPlan &f, const FwdFftLaunch &a, cudaStream_t s)
    {
        const int nimg = a.nimg > 0 ? a.nimg : 1;
        dim3 ga((f.nkeep + G::L - 1) / G::L, a.nch * nimg);
        p2w_fwd_pass_a<N, R1><<<ga, G::THREADS, smem_a(N), s>>>(a.img, a.tmp, a.deapod, f.tw, f.nkeep, a.nch, a.nc_total,
                                                                a.ch0, a.half_in);
        TRON_CUDA(cudaGetLastError());
        dim3 gb(N / G::L, a.nch * nimg);
        p2w_fwd_pass_b<N, R1><<<gb, G::THREADS, smem_a(N), s>>>(a.tmp, a.grid, f.tw, f.nkeep);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
    /* both passes in one launch (modes 0 / 3, nkeep = N/2); a.sync holds 2 * nslices zeroed counters */
    static int adj_fused(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s)
    {
        const int per = a.nch * (N / G::L) + (N / 2) / G::L;
        TRON_CUDA(cudaMemsetAsync(a.sync, 0, 2 * (size_t)a.nslices * sizeof(int), s));
        p2w_adj_fused<N, R1><<<a.nslices * per + (N / 2) / G::L, G::THREADS, smem_a(N / 2), s>>>(
            a.grid, a.tmp, a.out, a.deapod, f.tw, a.nch, a.mode, a.half_out, a.zero_r2, a.nslices, a.ring,
            a.sync, a.sync + a.nslices);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
    static size_t smem_coil() { return (size_t)(G::L * G::LPX + G::L * G::fpitch(N / 2)) * sizeof(float2); }
    /* per-coil pass B; only for nkeep = N/2 and modes 1 / 2 */
    static int adj_b_coil(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s)
    {
        dim3 gb(f.nkeep / G::L, a.nslices);
        p2w_adj_pass_b_coil<N, R1><<<gb, G::THREADS, smem_coil(), s>>>(a.tmp, a.out, a.deapod, f.tw, a.nch, a.nc_total,
                                                                       a.ch0, a.mode, a.half_out);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
    /* sum-of-squares pass B; only for nkeep = N/2 and modes 0 / 3 */
    static int adj_b_sos(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s)
    {
        dim3 gb(f.nkeep / G::L, a.nslices);
        p2w_adj_pass_b_sos<N, R1><<<gb, G::THREADS, smem_a(N / 2), s>>>(a.tmp, a.out, a.deapod, f.tw, a.nch, a.mode, a.half_out);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
    static int adj_a(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s)
    {
        dim3 ga(N / G::L, a.nslices * a.nch);
        p2w_adj_pass_a<N, R1><<<ga, G::THREADS, smem_a(f.nkeep), s>>>(a.grid, a.tmp, f.tw, f.nkeep, a.zero_r2);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
};

/* line lengths served by the two-stage path: elements per thread (0: radix-8 path only) */
template <int N> struct P2WSplit { static constexpr int R1 = N == 512 ? 32 : (N == 256 ? 16 : 0); };

template <int N, int L> struct P2Launch {
    static constexpr size_t SMEM = (size_t)(P2<N, L>::NBUF * L * P2<N, L>::PITCH + N + N / 8 + 1) * sizeof(float2);
    static constexpr size_t SMEM_B = (size_t)(2 * L * P2<N, L>::PITCH + N + N / 8 + 1) * sizeof(float2);   /* adjoint pass B */
    static constexpr int THREADS = L * (N / 8);
    static int prepare()
    {
        TRON_CUDA(cudaFuncSetAttribute(p2_adj_pass_a<N, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        TRON_CUDA(cudaFuncSetAttribute(p2_adj_pass_b<N, L, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_B));
        TRON_CUDA(cudaFuncSetAttribute(p2_adj_pass_b<N, L, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_B));
        TRON_CUDA(cudaFuncSetAttribute(p2_fwd_pass_a<N, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        TRON_CUDA(cudaFuncSetAttribute(p2_fwd_pass_b<N, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        if constexpr (P2WSplit<N>::R1 != 0) { int rc = P2WLaunch<N, P2WSplit<N>::R1>::prepare(); if (rc) return rc; }
        return 0;
    }
    static int adj(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s)
    {
        bool wide_a = false;
        /* (a single plane gives the two-stage passes 64 / 32 blocks for 148 SMs: the radix-8 passes with their smaller
         * blocks take 38 instead of 42 us for the 512^2 adjoint of BASELINE config 1) */
        if constexpr (P2WSplit<N>::R1 != 0) wide_a = getenv("TRON_FFT_R8") == nullptr && (a.nslices * a.nch >= 4 || getenv("TRON_FFT_P2W") != nullptr);
        if constexpr (P2WSplit<N>::R1 != 0) {
            if (wide_a && a.sync && a.ring > 0 && 2 * f.nkeep == N && (a.mode == 0 || a.mode == 3))
                return P2WLaunch<N, P2WSplit<N>::R1>::adj_fused(f, a, s);
        }
        if (wide_a) {
            if constexpr (P2WSplit<N>::R1 != 0) { int rc = P2WLaunch<N, P2WSplit<N>::R1>::adj_a(f, a, s); if (rc) return rc; }
        } else {
            dim3 ga(N / L, a.nslices * a.nch);
            p2_adj_pass_a<N, L><<<ga, THREADS, SMEM, s>>>(a.grid, a.tmp, f.tw, f.nkeep, a.zero_r2);
            TRON_CUDA(cudaGetLastError());
        }
        if constexpr (P2WSplit<N>::R1 != 0) {
            if (wide_a && 2 * f.nkeep == N && (a.mode == 0 || a.mode == 3)) return P2WLaunch<N, P2WSplit<N>::R1>::adj_b_sos(f, a, s);
            if (wide_a && 2 * f.nkeep == N) return P2WLaunch<N, P2WSplit<N>::R1>::adj_b_coil(f, a, s);
        }
        dim3 gb((f.nkeep + L - 1) / L, a.nslices);
        if (2 * f.nkeep <= N)
            p2_adj_pass_b<N, L, 4><<<gb, THREADS, SMEM_B, s>>>(a.tmp, a.out, a.deapod, f.tw, f.nkeep, a.nch, a.nc_total,
                                                             a.ch0, a.mode, a.half_out);
        else
            p2_adj_pass_b<N, L, 8><<<gb, THREADS, SMEM_B, s>>>(a.tmp, a.out, a.deapod, f.tw, f.nkeep, a.nch, a.nc_total,
                                                             a.ch0, a.mode, a.half_out);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
    static int fwd(const FftPlan &f, const FwdFftLaunch &a, cudaStream_t s)
    {
        if constexpr (P2WSplit<N>::R1 != 0) {
            /* a single plane gives the 128-thread CTAs of the two-stage passes too few blocks to fill the GPU
             * (256^2 forward transform: 40 vs 33 us); from a few planes on they win (CGNR: 149 -> 144 us per slice) */
            const bool force = getenv("TRON_FFT_P2W") != nullptr;
            if (getenv("TRON_FFT_R8") == nullptr && (force || a.nch * (a.nimg > 0 ? a.nimg : 1) >= 4))
                return P2WLaunch<N, P2WSplit<N>::R1>::fwd(f, a, s);
        }
        const int nimg = a.nimg > 0 ? a.nimg : 1;
        const int rows = (f.nkeep + L - 1) / L, planes = a.nch * nimg;
        const int chan_fastest = a.nch >= 8 && rows <= 65535;
        dim3 ga(chan_fastest ? planes : rows, chan_fastest ? rows : planes);
        p2_fwd_pass_a<N, L><<<ga, THREADS, SMEM, s>>>(a.img, a.tmp, a.deapod, f.tw, f.nkeep, a.nch, a.nc_total, a.ch0, a.half_in,
                                                      chan_fastest);
        TRON_CUDA(cudaGetLastError());
        dim3 gb(N / L, a.nch * nimg);
        p2_fwd_pass_b<N, L><<<gb, THREADS, SMEM, s>>>(a.tmp, a.grid, f.tw, f.nkeep);
        TRON_CUDA(cudaGetLastError());
        return 0;
    }
};

/* lines per CTA chosen so that a CTA has 256..512 threads */
#define P2_DISPATCH(n, CALL)                                   \
    switch (n) {                                               \
    case 64:   return P2Launch<64, 16>::CALL;                  \
    case 128:  return P2Launch<128, 16>::CALL;                 \
    case 256:  return P2Launch<256, 8>::CALL;                  \
    case 512:  return P2Launch<512, 4>::CALL;                  \
    case 1024: return P2Launch<1024, 4>::CALL;                 \
    case 2048: return P2Launch<2048, 4>::CALL;                 \
    case 4096: return P2Launch<4096, 1>::CALL;                 \
    default: break;                                            \
    }

static bool is_p2(int n) { return n >= 64 && n <= 4096 && (n & (n - 1)) == 0; }
static int p2_prepare(int n) { P2_DISPATCH(n, prepare()) return 0; }
static int p2_adj(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s) { P2_DISPATCH(f.n, adj(f, a, s)) return TRON_EUNSUPPORTED; }
static int p2_fwd(const FftPlan &f, const FwdFftLaunch &a, cudaStream_t s) { P2_DISPATCH(f.n, fwd(f, a, s)) return TRON_EUNSUPPORTED; }

/* ------- deapodisation tables (reciprocal weights), tron.cu:351-370, 390-402 ------- */
__device__ __forceinline__ float kb_hat(float u, float kernwidth)
{
    float J = 2.0f * kernwidth;
    float beta = 2.34f * 2.0f * kernwidth;
    float r = (float)(3.14159265358979323846 * (double)J * (double)u);
    float q = r * r - beta * beta;
    float y;
    if (q > 0.f) { float z = sqrtf(q); y = __sinf(z) / z; }
    else if (q < 0.f) { float z = sqrtf(-q); y = sinhf(z) / z; }
    else y = 1.f;
    return y;
}

/* table[a][b] = 1 / w(id), id = (a + off)*n + (b + off), evaluated exactly as
 * deapodkernel does for an n x n array with the given sigma */
__global__ void deapod_table_kernel(float *tab, int nt, int n, int off, float m, float sigma)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt * nt; i += gridDim.x * blockDim.x) {
        int a = i / nt, b = i - a * nt;
        size_t id = (size_t)(a + off) * n + (b + off);
        float x = (float)id / (float)n - (float)((n + 1) / 2);      /* tron.cu:395 (float division) */
        float y = (float)(id % n) - (float)((n + 1) / 2);
        float scale = 1.f / (float)n / sigma;
        float wgt = kb_hat(x * scale, m) * kb_hat(y * scale, m);
        tab[i] = 1.0f / (wgt > 0.f ? wgt : 1.f);
    }
}

int launch_deapod_tables(float *adj_tab, float *fwd_tab, int nx, int nxos, float W, float gridos, cudaStream_t s)
{
    int blocks = (nx * nx + 255) / 256; if (blocks > 4096) blocks = 4096;
    if (adj_tab) deapod_table_kernel<<<blocks, 256, 0, s>>>(adj_tab, nx, nx, 0, W, gridos);
    if (fwd_tab) deapod_table_kernel<<<blocks, 256, 0, s>>>(fwd_tab, nx, nxos, (nxos - nx) / 2, W, 1.f);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

/* ---------------- host side ---------------- */
static bool factorize(int n, Factors &f)
{
    f.nfac = 0;
    int m = n;
    while (m % 8 == 0) { f.fac[f.nfac++] = 8; m /= 8; }
    while (m % 4 == 0) { f.fac[f.nfac++] = 4; m /= 4; }
    while (m % 2 == 0) { f.fac[f.nfac++] = 2; m /= 2; }
    while (m % 3 == 0) { f.fac[f.nfac++] = 3; m /= 3; }
    while (m % 5 == 0) { f.fac[f.nfac++] = 5; m /= 5; }
    return m == 1 && f.nfac <= 16;
}

static PassGeom make_geom(const FftPlan &f)
{
    PassGeom p; p.n = f.n; p.nkeep = f.nkeep; p.L = f.lines;
    p.pitch = phys_host(f.n) + 1;
    p.f.nfac = f.nfac;
    for (int i = 0; i < f.nfac; ++i) p.f.fac[i] = f.fac[i];
    return p;
}

int fft_plan_init(FftPlan &f, int n, int nkeep)
{
    Factors fa;
    if (n < 2 || n > 8192 || !factorize(n, fa)) {
        set_error("oversampled grid size %d is not of the form 2^a 3^b 5^c <= 8192", n);
        return TRON_EUNSUPPORTED;
    }
    if (nkeep > n || nkeep > 4096 || nkeep < 1) { set_error("bad image size %d for grid %d", nkeep, n); return TRON_EINVAL; }
    f.n = n; f.nkeep = nkeep; f.nfac = fa.nfac;
    for (int i = 0; i < fa.nfac; ++i) f.fac[i] = fa.fac[i];
    int pitch = phys_host(n) + 1;
    /* lines per CTA: bounded by shared memory (two buffers + twiddles) and by the
     * PASSB_MAXO outputs a thread owns in pass B */
    int L = 16;
    while (L > 1 && ((size_t)(2 * L * pitch + n) * sizeof(float2) > 96 * 1024 || nkeep * L > PASSB_MAXO * 256)) L >>= 1;
    f.lines = L;
    f.smem = (size_t)(2 * L * pitch + n) * sizeof(float2);
    if (f.smem > 227 * 1024) { set_error("grid size %d needs %zu B of shared memory", n, f.smem); return TRON_EUNSUPPORTED; }
    std::vector<float2> tw(n);
    for (int k = 0; k < n; ++k) {
        double a = 2.0 * 3.14159265358979323846 * (double)k / (double)n;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    TRON_CUDA(cudaMalloc(&f.tw, n * sizeof(float2)));
    TRON_CUDA(cudaMemcpy(f.tw, tw.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    if (is_p2(n) && !getenv("TRON_GENERIC_FFT")) { int rc = p2_prepare(n); if (rc) return rc; f.pow2 = 1; }
    TRON_CUDA(cudaFuncSetAttribute(adj_pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem));
    TRON_CUDA(cudaFuncSetAttribute(adj_pass_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem));
    TRON_CUDA(cudaFuncSetAttribute(fwd_pass_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem));
    TRON_CUDA(cudaFuncSetAttribute(fwd_pass_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem));
    return 0;
}

void fft_plan_free(FftPlan &f)
{
    if (f.tw) cudaFree(f.tw);
    f.tw = nullptr;
}

int launch_adj_fft(const FftPlan &f, const AdjFftLaunch &a, cudaStream_t s)
{
    if (f.pow2) return p2_adj(f, a, s);
    PassGeom p = make_geom(f);
    dim3 ga((f.n + p.L - 1) / p.L, a.nslices * a.nch);
    adj_pass_a_kernel<<<ga, 256, f.smem, s>>>(a.grid, a.tmp, f.tw, p);
    TRON_CUDA(cudaGetLastError());
    dim3 gb((f.nkeep + p.L - 1) / p.L, a.nslices);
    adj_pass_b_kernel<<<gb, 256, f.smem, s>>>(a.tmp, a.out, a.deapod, f.tw, p, a.nch, a.nc_total, a.ch0,
                                              a.mode, a.half_out);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

/* does launch_adj_fft run this launch as one kernel (p2w_adj_fused)? */
bool adj_fft_single_launch(const FftPlan &f, const AdjFftLaunch &a)
{
    return f.pow2 && (f.n == 512 || f.n == 256) && getenv("TRON_FFT_R8") == nullptr && a.sync && a.ring > 0 &&
           2 * f.nkeep == f.n && (a.mode == 0 || a.mode == 3);
}

int launch_fwd_fft(const FftPlan &f, const FwdFftLaunch &a, cudaStream_t s)
{
    if (f.pow2) return p2_fwd(f, a, s);
    PassGeom p = make_geom(f);
    const int nimg = a.nimg > 0 ? a.nimg : 1;
    dim3 ga((f.nkeep + p.L - 1) / p.L, a.nch * nimg);
    fwd_pass_a_kernel<<<ga, 256, f.smem, s>>>(a.img, a.tmp, a.deapod, f.tw, p, a.nch, a.nc_total, a.ch0, a.half_in);
    TRON_CUDA(cudaGetLastError());
    dim3 gb((f.n + p.L - 1) / p.L, a.nch * nimg);
    fwd_pass_b_kernel<<<gb, 256, f.smem, s>>>(a.tmp, a.grid, f.tw, p);
    TRON_CUDA(cudaGetLastError());
    return 0;
}

} // namespace tronb
