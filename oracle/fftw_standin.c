/* Stand-in for the five FFTW3 entry points the reference's mex/nddwt.c uses.
 * TEST INFRASTRUCTURE (see oracle/standin/fftw3.h).  Forward DFT (sign -1), split real/imag
 * arrays, arbitrary rank with explicit strides, one batch ("howmany") dimension -- exactly the
 * shape init_fftw_plan builds (nddwt.c:26-47).  1-D kernels: iterative radix-2 for powers of
 * two, Bluestein's chirp-z otherwise.  OpenMP over the independent lines of each dimension. */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_RANK 8
static int g_threads = 1;

typedef struct {
    int n, m;                 /* length, padded power of two for Bluestein (0 if n is 2^k) */
    double *twr, *twi;        /* radix-2 twiddles for size (m ? m : n): exp(-2 pi i k / size) */
    double *chr, *chi;        /* Bluestein chirp exp(-i pi k^2 / n), k < n */
    double *bfr, *bfi;        /* FFT of the conjugate chirp, length m */
} fft1d;

struct fftw_plan_s {
    int rank;
    fftw_iodim dims[MAX_RANK];
    int howmany_rank;
    fftw_iodim howmany[MAX_RANK];
    fft1d k[MAX_RANK];
};

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static void radix2(double *re, double *im, int n, const double *twr, const double *twi, int inverse)
{
    for (int i = 1, j = 0; i < n; i++) {           /* bit reversal */
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len) {
            for (int q = 0; q < half; q++) {
                double wr = twr[q * step], wi = inverse ? -twi[q * step] : twi[q * step];
                double xr = re[s + q + half], xi = im[s + q + half];
                double tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
                re[s + q + half] = re[s + q] - tr; im[s + q + half] = im[s + q] - ti;
                re[s + q] += tr; im[s + q] += ti;
            }
        }
    }
}

static void fft1d_init(fft1d *f, int n)
{
    memset(f, 0, sizeof(*f));
    f->n = n;
    int size = n;
    if (!is_pow2(n)) { int m = 1; while (m < 2 * n - 1) m <<= 1; f->m = m; size = m; }
    f->twr = (double *)malloc(sizeof(double) * size);
    f->twi = (double *)malloc(sizeof(double) * size);
    for (int q = 0; q < size; q++) {
        double a = -2.0 * M_PI * (double)q / (double)size;
        f->twr[q] = cos(a); f->twi[q] = sin(a);
    }
    if (f->m) {
        int m = f->m;
        f->chr = (double *)malloc(sizeof(double) * n); f->chi = (double *)malloc(sizeof(double) * n);
        f->bfr = (double *)calloc(m, sizeof(double)); f->bfi = (double *)calloc(m, sizeof(double));
        for (int q = 0; q < n; q++) {
            long long q2 = ((long long)q * q) % (2LL * n);   /* keep the angle small */
            double a = -M_PI * (double)q2 / (double)n;
            f->chr[q] = cos(a); f->chi[q] = sin(a);
        }
        f->bfr[0] = f->chr[0]; f->bfi[0] = -f->chi[0];
        for (int q = 1; q < n; q++) {
            f->bfr[q] = f->bfr[m - q] = f->chr[q];
            f->bfi[q] = f->bfi[m - q] = -f->chi[q];
        }
        radix2(f->bfr, f->bfi, m, f->twr, f->twi, 0);
    }
}

static void fft1d_free(fft1d *f)
{
    free(f->twr); free(f->twi); free(f->chr); free(f->chi); free(f->bfr); free(f->bfi);
}

/* forward DFT of a contiguous line; work holds 2*m doubles when Bluestein is needed */
static void fft1d_run(const fft1d *f, double *re, double *im, double *work)
{
    int n = f->n;
    if (!f->m) { radix2(re, im, n, f->twr, f->twi, 0); return; }
    int m = f->m;
    double *ar = work, *ai = work + m;
    for (int q = 0; q < n; q++) {
        ar[q] = re[q] * f->chr[q] - im[q] * f->chi[q];
        ai[q] = re[q] * f->chi[q] + im[q] * f->chr[q];
    }
    for (int q = n; q < m; q++) { ar[q] = 0.0; ai[q] = 0.0; }
    radix2(ar, ai, m, f->twr, f->twi, 0);
    for (int q = 0; q < m; q++) {
        double tr = ar[q] * f->bfr[q] - ai[q] * f->bfi[q];
        double ti = ar[q] * f->bfi[q] + ai[q] * f->bfr[q];
        ar[q] = tr; ai[q] = ti;
    }
    radix2(ar, ai, m, f->twr, f->twi, 1);
    double inv = 1.0 / (double)m;
    for (int q = 0; q < n; q++) {
        double tr = ar[q] * inv, ti = ai[q] * inv;
        re[q] = tr * f->chr[q] - ti * f->chi[q];
        im[q] = tr * f->chi[q] + ti * f->chr[q];
    }
}

int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int nthreads) { g_threads = nthreads > 0 ? nthreads : 1; }

fftw_plan fftw_plan_guru_split_dft(int rank, const fftw_iodim *dims, int howmany_rank,
                                   const fftw_iodim *howmany_dims, double *ri, double *ii,
                                   double *ro, double *io, unsigned flags)
{
    (void)ri; (void)ii; (void)ro; (void)io; (void)flags;
    if (rank < 1 || rank > MAX_RANK || howmany_rank < 0 || howmany_rank > MAX_RANK) return NULL;
    fftw_plan p = (fftw_plan)calloc(1, sizeof(*p));
    p->rank = rank; p->howmany_rank = howmany_rank;
    for (int d = 0; d < rank; d++) { p->dims[d] = dims[d]; fft1d_init(&p->k[d], dims[d].n); }
    for (int d = 0; d < howmany_rank; d++) p->howmany[d] = howmany_dims[d];
    return p;
}

void fftw_destroy_plan(fftw_plan p)
{
    if (!p) return;
    for (int d = 0; d < p->rank; d++) fft1d_free(&p->k[d]);
    free(p);
}

/* transform every line along dimension `dim` of one rank-`rank` array, in place in (re, im) */
static void transform_dim(const fftw_plan p, int dim, double *re, double *im)
{
    int rank = p->rank, n = p->dims[dim].n;
    long stride = p->dims[dim].os;
    long nlines = 1;
    for (int d = 0; d < rank; d++) if (d != dim) nlines *= p->dims[d].n;
    int m = p->k[dim].m;
#pragma omp parallel num_threads(g_threads)
    {
        double *lr = (double *)malloc(sizeof(double) * (size_t)(2 * n + 2 * (m ? m : 1)));
        double *li = lr + n, *work = li + n;
#pragma omp for schedule(static)
        for (long line = 0; line < nlines; line++) {
            long rem = line, base = 0;
            for (int d = 0; d < rank; d++) {
                if (d == dim) continue;
                long c = rem % p->dims[d].n; rem /= p->dims[d].n;
                base += c * (long)p->dims[d].os;
            }
            for (int q = 0; q < n; q++) { lr[q] = re[base + q * stride]; li[q] = im[base + q * stride]; }
            fft1d_run(&p->k[dim], lr, li, work);
            for (int q = 0; q < n; q++) { re[base + q * stride] = lr[q]; im[base + q * stride] = li[q]; }
        }
        free(lr);
    }
}

void fftw_execute_split_dft(const fftw_plan p, double *ri, double *ii, double *ro, double *io)
{
    long batches = 1, bis = 0, bos = 0;
    if (p->howmany_rank >= 1) { batches = p->howmany[0].n; bis = p->howmany[0].is; bos = p->howmany[0].os; }
    long numel = 1;
    for (int d = 0; d < p->rank; d++) numel *= p->dims[d].n;
    for (long b = 0; b < batches; b++) {
        double *sr = ri + b * bis, *si = ii + b * bis, *dr = ro + b * bos, *di = io + b * bos;
        if (dr != sr) {   /* out of place: the reference always uses is == os, dense column-major */
            memcpy(dr, sr, sizeof(double) * (size_t)numel);
            memcpy(di, si, sizeof(double) * (size_t)numel);
        }
        for (int d = 0; d < p->rank; d++) transform_dim(p, d, dr, di);
    }
}
