/* FFTW3-API shim for building the BOUT++ reference sources as a CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  FFTW3 (any 3.x; version unpinned by the reference's
 * cmake/FindFFTW.cmake) is an external dependency absent from /root/reference and from this
 * image.  This header declares the subset called by src/invert/fft_fftw.cxx
 * (reference call sites :156,167,220,233,293,310,368,388,428,445,484,507) with FFTW's published
 * semantics: r2c = unnormalised forward DFT X_k = sum_j x_j exp(-2 pi i jk/n), k = 0..n/2;
 * c2r = unnormalised backward DFT of the Hermitian extension (imaginary parts of the DC and
 * Nyquist bins ignored).
 */
#ifndef BOUT_B200_ORACLE_FFTW_SHIM_H
#define BOUT_B200_ORACLE_FFTW_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct bout_b200_fftw_plan_s* fftw_plan;

#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_ESTIMATE (1U << 6)

void* fftw_malloc(size_t n);
void fftw_free(void* p);
fftw_plan fftw_plan_dft_r2c_1d(int n, double* in, fftw_complex* out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex* in, double* out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
