/* Declarations-only stand-in for the subset of FFTW3's API that the reference's
 * mex/nddwt.c calls (nddwt.c:12,26-27,55,101-103,128,137-138).  TEST INFRASTRUCTURE:
 * FFTW3 is not installed in this image, so oracle/_ref links the reference's own nddwt.c
 * against oracle/fftw_standin.c, which implements these five entry points with a plain
 * split-complex mixed radix-2 / Bluestein DFT.  Signatures follow the public FFTW 3.3 manual
 * ("Guru Split-Array DFTs"); no FFTW source was consulted or copied. */
#ifndef NDDWT_ORACLE_FFTW3_STANDIN_H
#define NDDWT_ORACLE_FFTW3_STANDIN_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fftw_plan_s *fftw_plan;
typedef struct { int n; int is; int os; } fftw_iodim;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
fftw_plan fftw_plan_guru_split_dft(int rank, const fftw_iodim *dims,
                                   int howmany_rank, const fftw_iodim *howmany_dims,
                                   double *ri, double *ii, double *ro, double *io,
                                   unsigned flags);
void fftw_execute_split_dft(const fftw_plan p, double *ri, double *ii, double *ro, double *io);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
