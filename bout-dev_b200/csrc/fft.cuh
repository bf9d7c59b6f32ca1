// Shared-memory real FFT along z for one row per thread group (fp64).
//
// Semantics are the reference's bout::fft::rfft / irfft (src/invert/fft_fftw.cxx:116-240): forward
// transform scaled by 1/nz, backward unscaled, FFTW r2c/c2r conventions (the imaginary parts of
// the DC and Nyquist bins are ignored on the way back).  FFTW itself is an external dependency of
// the reference; this is an independent implementation: a length-nz real transform is done as a
// length-N = nz/2 complex Stockham autosort FFT (radix-4 stages, one radix-2 stage if log2 N is
// odd) on the row viewed as packed complex numbers, plus the usual even/odd split.  Twiddles come
// from a table of exp(-2 pi i k / nz) rounded from long double on the host, so the result agrees
// with a direct DFT to ~1e-16 * log2(nz).  Lengths that are not a power of two (or < 8) take a
// direct O(nz^2) path with exactly reduced twiddle indices -- slow, used by small test meshes only.
//
// Calling convention: a CTA processes `rows` rows at once; `tpr` consecutive threads share a row.
// All threads of the CTA must call the functions together (they contain __syncthreads()).
#pragma once

#include "common.cuh"

namespace b200 {

struct cd { // complex double, 16-byte aligned so it moves as one 128-bit transaction
  double re, im;
};
static_assert(sizeof(cd) == 16, "cd must be 16 bytes");

B200_HD cd cmul(cd a, cd b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
B200_HD cd cadd(cd a, cd b) { return {a.re + b.re, a.im + b.im}; }
B200_HD cd csub(cd a, cd b) { return {a.re - b.re, a.im - b.im}; }
B200_HD cd cconj(cd a) { return {a.re, -a.im}; }
B200_HD cd cscale(cd a, double s) { return {a.re * s, a.im * s}; }
B200_HD cd ldtw(const cd* __restrict__ tw, int i) {
  const double2 t = B200_LDG(reinterpret_cast<const double2*>(tw) + i);
  return {t.x, t.y};
}

struct RowFFT {
  int nz;        // real length
  int N;         // nz / 2 (pow2 path)
  bool pow2;     // fast path available
  int tpr;       // threads per row
  int lane;      // this thread's index within its row group
  cd* A;         // smem buffer, N + 1 complex (holds the packed real row on entry)
  cd* B;         // smem buffer, N + 1 complex
  const cd* tw;  // global table exp(-2 pi i k / nz), k = 0..nz-1
  bool wsync = false; // the row is owned by exactly one warp: synchronise with __syncwarp only
};

// Row-level barrier: rows owned by a single warp progress independently of the CTA's other rows.
__device__ __forceinline__ void rowSync(const RowFFT& r) {
  if (r.wsync) {
    __syncwarp();
  } else {
    __syncthreads();
  }
}

__host__ __device__ inline int fftBufElems(int nz) { return nz / 2 + 2; } // cd elements per buffer

// Complex FFT of N points, in -> out ping-pong; returns the buffer holding the result.
template <bool INVERSE>
__device__ __forceinline__ cd* cfftStockham(const RowFFT& r, cd* in, cd* out) {
  const int N = r.N;
  int Ns = 1;
  // radix-4 stages
  while (Ns * 4 <= N) {
    const int tstep = r.nz / (Ns * 4); // table stride for exp(-2 pi i k / (4 Ns))
    for (int j = r.lane; j < N / 4; j += r.tpr) {
      const int k = j & (Ns - 1);
      cd v0 = in[j], v1 = in[j + N / 4], v2 = in[j + N / 2], v3 = in[j + 3 * (N / 4)];
      if (Ns > 1) {
        cd w1 = ldtw(r.tw, k * tstep), w2 = ldtw(r.tw, 2 * k * tstep), w3 = ldtw(r.tw, 3 * k * tstep);
        if (INVERSE) {
          w1 = cconj(w1);
          w2 = cconj(w2);
          w3 = cconj(w3);
        }
        v1 = cmul(v1, w1);
        v2 = cmul(v2, w2);
        v3 = cmul(v3, w3);
      }
      const cd t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3);
      const cd d = csub(v1, v3);
      // forward: multiply by -i; inverse: by +i
      const cd t3 = INVERSE ? cd{-d.im, d.re} : cd{d.im, -d.re};
      const int j0 = ((j - k) << 2) + k; // (j / Ns) * Ns * 4 + k
      out[j0] = cadd(t0, t2);
      out[j0 + Ns] = cadd(t1, t3);
      out[j0 + 2 * Ns] = csub(t0, t2);
      out[j0 + 3 * Ns] = csub(t1, t3);
    }
    rowSync(r);
    cd* t = in;
    in = out;
    out = t;
    Ns *= 4;
  }
  if (Ns < N) { // one radix-2 stage
    const int tstep = r.nz / (Ns * 2);
    for (int j = r.lane; j < N / 2; j += r.tpr) {
      const int k = j & (Ns - 1);
      cd v0 = in[j], v1 = in[j + N / 2];
      cd w = ldtw(r.tw, k * tstep);
      if (INVERSE) {
        w = cconj(w);
      }
      v1 = cmul(v1, w);
      const int j0 = ((j - k) << 1) + k;
      out[j0] = cadd(v0, v1);
      out[j0 + Ns] = csub(v0, v1);
    }
    rowSync(r);
    cd* t = in;
    in = out;
    out = t;
  }
  return in;
}

// Forward real FFT.  On entry r.A holds the real row (nz doubles = N packed complex) and the CTA
// has synchronised after filling it.  Returns the buffer holding X[0..nz/2] (scaled by 1/nz);
// the CTA is synchronised on return.
__device__ __forceinline__ cd* rfftForward(const RowFFT& r) {
  const int nz = r.nz;
  if (r.pow2) {
    const int N = r.N;
    cd* Z = cfftStockham<false>(r, r.A, r.B);
    cd* X = (Z == r.A) ? r.B : r.A;
    const double inv = 1.0 / nz;
    for (int k = r.lane; k <= N; k += r.tpr) {
      const cd zk = Z[k & (N - 1)];
      const cd zc = cconj(Z[(N - k) & (N - 1)]);
      const cd e = cscale(cadd(zk, zc), 0.5);
      const cd dmi = csub(zk, zc);
      const cd o = {0.5 * dmi.im, -0.5 * dmi.re}; // -i/2 * (zk - zc)
      const cd wk = (k == N) ? cd{-1.0, 0.0} : ldtw(r.tw, k);
      X[k] = cscale(cadd(e, cmul(wk, o)), inv);
    }
    rowSync(r);
    return X;
  }
  // direct DFT
  const double* x = reinterpret_cast<const double*>(r.A);
  const double inv = 1.0 / nz;
  for (int k = r.lane; k <= nz / 2; k += r.tpr) {
    double re = 0.0, im = 0.0;
    for (int j = 0; j < nz; ++j) {
      const cd w = ldtw(r.tw, static_cast<int>((static_cast<long long>(j) * k) % nz));
      re += x[j] * w.re;
      im += x[j] * w.im;
    }
    r.B[k] = {re * inv, im * inv};
  }
  rowSync(r);
  return r.B;
}

// Backward real FFT.  X (r.A or r.B, nz/2+1 entries) holds the spectrum and the CTA has
// synchronised.  Returns the buffer holding the real row (nz doubles); CTA synchronised on return.
__device__ __forceinline__ double* rfftBackward(const RowFFT& r, cd* X) {
  const int nz = r.nz;
  cd* other = (X == r.A) ? r.B : r.A;
  if (r.pow2) {
    const int N = r.N;
    // build Z' in `other`
    for (int k = r.lane; k < N; k += r.tpr) {
      cd xk = X[k];
      cd xc = cconj(X[N - k]);
      if (k == 0) { // imaginary parts of DC and Nyquist are ignored (FFTW c2r)
        xk.im = 0.0;
        xc.im = 0.0;
      }
      const cd wk = cconj(ldtw(r.tw, k));
      const cd d = cmul(wk, csub(xk, xc));
      other[k] = cadd(cadd(xk, xc), cd{-d.im, d.re}); // + i * wk * (xk - xc)
    }
    rowSync(r);
    cd* z = cfftStockham<true>(r, other, X);
    return reinterpret_cast<double*>(z);
  }
  double* out = reinterpret_cast<double*>(other);
  for (int j = r.lane; j < nz; j += r.tpr) {
    double acc = X[0].re;
    for (int k = 1; 2 * k < nz; ++k) {
      const cd w = ldtw(r.tw, static_cast<int>((static_cast<long long>(j) * k) % nz));
      acc += 2.0 * (X[k].re * w.re + X[k].im * w.im);
    }
    if (nz % 2 == 0) {
      acc += (j % 2 == 0) ? X[nz / 2].re : -X[nz / 2].re;
    }
    out[j] = acc;
  }
  rowSync(r);
  return out;
}

// Host side: device twiddle table for a given nz (cached per device for the process lifetime).
const cd* twiddleTable(int nz, cudaStream_t stream, int* status);

// Row-group geometry for a CTA of `threads` threads.
struct RowGeom {
  int tpr, rows;
  bool pow2;
  size_t smemBytes;
};
inline RowGeom rowGeometry(int nz, int threads, bool warp_rows = false) {
  RowGeom g;
  g.pow2 = nz >= 8 && (nz & (nz - 1)) == 0;
  int want = g.pow2 ? nz / 8 : 32; // N/4 butterflies per radix-4 stage
  int tpr = 32;
  while (!warp_rows && tpr < want && tpr < threads) {
    tpr *= 2;
  }
  g.tpr = tpr;
  g.rows = threads / tpr;
  g.smemBytes = static_cast<size_t>(g.rows) * 2 * fftBufElems(nz) * sizeof(cd);
  return g;
}

} // namespace b200
