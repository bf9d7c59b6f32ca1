// Register-resident z-FFT core (nz = 128, 256, 512; N = nz/2 complex points per row).
//
// Same transform and normalisation as fft.cuh / fft_fast.cuh (the reference's bout::fft::rfft / irfft,
// src/invert/fft_fftw.cxx:116-240: forward scaled by 1/nz, backward unscaled, FFTW r2c / c2r conventions).
// What changes is where the data lives.  fft_fast.cuh runs four radix-4 Stockham stages through shared
// memory (ncu: L1/TEX pipe 72 % busy, the limiter of every z-FFT kernel of round 1).  Here a row is
// owned by T threads holding B = N/T points each IN REGISTERS:
//
//   N = 256: T = 16, B = 16      N = 128: T = 8, B = 16      N = 64: T = 8, B = 8
//
//   stage 1   thread n1 holds z[n1 + T n2] (n2 = 0..B-1, coalesced 16-byte loads), does a B-point DFT
//             over n2 in registers and multiplies by the twiddles W_N^(n1 k2);
//   exchange  ONE transposition through shared memory (XOR-swizzled, bank-conflict free both ways);
//   stage 2   thread t does B/T T-point DFTs over n1 in registers and ends up holding Z[t + T q].
//
// i.e. one shared-memory round trip per complex transform instead of three or four.  The real-FFT
// split / merge pairs bin k with N - k, which lives in thread T - t: one more pass through the row's
// buffer.  A row's buffer is N + 1 complex numbers (the spectrum X[0 .. N]) and doubles as the exchange
// area, so a CTA needs no shared memory beyond the spectra it keeps.
//
// Every function takes the thread's position (t) as an argument and contains no barrier: the callers
// (kernels) put __syncwarp() between the phases, and tests/emul/ runs the same phases on the host, one
// simulated thread after another, to check the index algebra without a GPU.
#pragma once

#include "fft.cuh"

namespace b200 {

template <int N>
struct RegFft {
  static_assert(N == 64 || N == 128 || N == 256, "register FFT: nz/2 in {64, 128, 256}");
  static constexpr int T = (N == 256) ? 16 : 8; // threads per row
  static constexpr int B = N / T;               // points per thread (16, 16, 8)
  static constexpr int U = B / T;               // stage-2 transforms per thread (1, 2, 1)
  static constexpr int SROW = N + 1;            // complex elements of a row buffer
};

// cos(pi q / 16), sin(pi q / 16), q = 0 .. 16 (rounded from long double)
__host__ __device__ constexpr double regCos16(int q) {
  constexpr double c[17] = {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476,
                            0.5555702330196022, 0.3826834323650898, 0.19509032201612828, 0.0,
                            -0.19509032201612828, -0.3826834323650898, -0.5555702330196022, -0.7071067811865476,
                            -0.8314696123025452, -0.9238795325112867, -0.9807852804032304, -1.0};
  return c[q];
}
__host__ __device__ constexpr double regSin16(int q) {
  constexpr double s[17] = {0.0, 0.19509032201612828, 0.3826834323650898, 0.5555702330196022, 0.7071067811865476,
                            0.8314696123025452, 0.9238795325112867, 0.9807852804032304, 1.0,
                            0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476,
                            0.5555702330196022, 0.3826834323650898, 0.19509032201612828, 0.0};
  return s[q];
}

// ---- small DFTs in registers (forward: exp(-2 pi i nk/n); INV: conjugate twiddles, unscaled) -----------------
template <bool INV>
B200_HD cd mulMinusI(cd a) { // forward: a * (-i); inverse: a * (+i)
  return INV ? cd{-a.im, a.re} : cd{a.im, -a.re};
}
template <bool INV>
B200_HD cd mulW8_1(cd a) { // a * exp(-+ i pi/4)
  constexpr double h = 0.7071067811865476;
  return INV ? cd{h * (a.re - a.im), h * (a.re + a.im)} : cd{h * (a.re + a.im), h * (a.im - a.re)};
}
template <bool INV>
B200_HD cd mulW8_3(cd a) { // a * exp(-+ 3 i pi/4)
  constexpr double h = 0.7071067811865476;
  return INV ? cd{-h * (a.re + a.im), h * (a.re - a.im)} : cd{h * (a.im - a.re), -h * (a.re + a.im)};
}
template <bool INV>
B200_HD cd mulW(cd a, double wr, double wi) { // a * (wr + i wi) forward, a * conj(..) inverse
  return INV ? cd{a.re * wr + a.im * wi, a.im * wr - a.re * wi} : cd{a.re * wr - a.im * wi, a.im * wr + a.re * wi};
}

template <bool INV>
B200_HD void dft4(cd& x0, cd& x1, cd& x2, cd& x3) {
  const cd t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3);
  const cd t3 = mulMinusI<INV>(csub(x1, x3));
  x0 = cadd(t0, t2);
  x1 = cadd(t1, t3);
  x2 = csub(t0, t2);
  x3 = csub(t1, t3);
}

// 16-point DFT, natural order in and out: n = na + 4 nb, k = 4 ka + kb.
template <bool INV>
B200_HD void dft16(cd (&x)[16]) {
  constexpr double c = 0.9238795325112867, s = 0.3826834323650898; // cos, sin of pi/8
#pragma unroll
  for (int na = 0; na < 4; ++na) {
    dft4<INV>(x[na], x[na + 4], x[na + 8], x[na + 12]); // x[na + 4 kb] = sum_nb x[na + 4 nb] W4^(nb kb)
  }
  // W16^(na kb)
  x[1 + 4] = mulW<INV>(x[1 + 4], c, -s);     // W16^1
  x[1 + 8] = mulW8_1<INV>(x[1 + 8]);         // W16^2
  x[1 + 12] = mulW<INV>(x[1 + 12], s, -c);   // W16^3
  x[2 + 4] = mulW8_1<INV>(x[2 + 4]);         // W16^2
  x[2 + 8] = mulMinusI<INV>(x[2 + 8]);       // W16^4
  x[2 + 12] = mulW8_3<INV>(x[2 + 12]);       // W16^6
  x[3 + 4] = mulW<INV>(x[3 + 4], s, -c);     // W16^3
  x[3 + 8] = mulW8_3<INV>(x[3 + 8]);         // W16^6
  x[3 + 12] = mulW<INV>(x[3 + 12], -c, s);   // W16^9
  cd y[16];
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    dft4<INV>(x[4 * kb], x[1 + 4 * kb], x[2 + 4 * kb], x[3 + 4 * kb]); // x[ka + 4 kb] = X[4 ka + kb]
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
      y[4 * ka + kb] = x[ka + 4 * kb];
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    x[k] = y[k];
  }
}

// 8-point DFT, natural order in and out: n = na + 4 nb, k = 2 ka + kb.
template <bool INV>
B200_HD void dft8(cd (&x)[8]) {
#pragma unroll
  for (int na = 0; na < 4; ++na) {
    const cd a = x[na], b = x[na + 4];
    x[na] = cadd(a, b);
    x[na + 4] = csub(a, b);
  }
  x[1 + 4] = mulW8_1<INV>(x[1 + 4]);
  x[2 + 4] = mulMinusI<INV>(x[2 + 4]);
  x[3 + 4] = mulW8_3<INV>(x[3 + 4]);
  cd y[8];
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    dft4<INV>(x[4 * kb], x[1 + 4 * kb], x[2 + 4 * kb], x[3 + 4 * kb]); // x[ka + 4 kb] = X[2 ka + kb]
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
      y[2 * ka + kb] = x[ka + 4 * kb];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = y[k];
  }
}

template <int B, bool INV>
B200_HD void dftB(cd (&x)[B]) {
  if constexpr (B == 16) {
    dft16<INV>(x);
  } else {
    dft8<INV>(x);
  }
}

// ---- the complex transform of one row: stage 1, exchange, stage 2 ---------------------------------------------
// Twiddle bases of thread t: W_N^(t e) for e = 1, 2, 3 and 4, 8, 12; every W_N^(t k2) is then at most one
// product of two correctly rounded table values.  tw = exp(-2 pi i k / nz), k = 0 .. nz-1 (fft.cuh).
template <int N, bool INV>
B200_HD void regfftStage1(cd (&v)[RegFft<N>::B], const cd* __restrict__ tw, int t) {
  constexpr int B = RegFft<N>::B, nz = 2 * N;
  dftB<B, INV>(v);
  cd lo[3], hi[3];
#pragma unroll
  for (int e = 1; e <= 3; ++e) {
    lo[e - 1] = ldtw(tw, (2 * t * e) & (nz - 1));
    if (4 * e < B) {
      hi[e - 1] = ldtw(tw, (2 * t * 4 * e) & (nz - 1));
    }
  }
#pragma unroll
  for (int k2 = 1; k2 < B; ++k2) {
    const int a = k2 >> 2, b = k2 & 3;
    cd w = b == 0 ? hi[a - 1] : (a == 0 ? lo[b - 1] : cmul(hi[a - 1], lo[b - 1]));
    if (INV) {
      w = cconj(w);
    }
    v[k2] = cmul(v[k2], w);
  }
}

// exchange area index of (k2, n1): conflict-free for "fixed k2, lanes n1" and for "fixed n1, lanes k2"
template <int N>
B200_HD int regfftXidx(int k2, int n1) {
  constexpr int T = RegFft<N>::T;
  return k2 * T + (n1 ^ (k2 & (T - 1)));
}

template <int N>
B200_HD void regfftXchgWrite(const cd (&v)[RegFft<N>::B], cd* E, int t) {
#pragma unroll
  for (int k2 = 0; k2 < RegFft<N>::B; ++k2) {
    E[regfftXidx<N>(k2, t)] = v[k2];
  }
}

// after the barrier: gather, stage 2; v[q] = Z[t + T q]
template <int N, bool INV>
B200_HD void regfftStage2(cd (&v)[RegFft<N>::B], const cd* E, int t) {
  constexpr int T = RegFft<N>::T, U = RegFft<N>::U;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    cd w[T];
#pragma unroll
    for (int n1 = 0; n1 < T; ++n1) {
      w[n1] = E[regfftXidx<N>(t + T * u, n1)];
    }
    dftB<T, INV>(w);
#pragma unroll
    for (int k1 = 0; k1 < T; ++k1) {
      v[u + U * k1] = w[k1]; // bin B k1 + t + T u = t + T (u + U k1)
    }
  }
}

// ---- real split (forward) -------------------------------------------------------------------------------------
// exp(-2 pi i (t + T q) / nz) = ws * exp(-i pi q / B), ws = tw[t]
template <int N>
B200_HD cd regfftSplitTw(cd ws, int q) {
  constexpr int step = 16 / RegFft<N>::B;
  return cmul(ws, cd{regCos16(q * step), -regSin16(q * step)});
}

template <int N>
B200_HD void regfftSplitWriteZ(const cd (&v)[RegFft<N>::B], cd* Z, int t) {
#pragma unroll
  for (int q = 0; q < RegFft<N>::B; ++q) {
    Z[t + RegFft<N>::T * q] = v[q];
  }
}

// after the barrier: the partners conj(Z[N - k]) of this thread's bins k = t + T q < N/2 (and of k = N/2 for t = 0)
template <int N>
B200_HD void regfftSplitReadPartners(cd (&pc)[RegFft<N>::B / 2 + 1], const cd* Z, int t) {
  constexpr int T = RegFft<N>::T, B = RegFft<N>::B;
#pragma unroll
  for (int q = 0; q <= B / 2; ++q) {
    const int k = t + T * q;
    pc[q] = (q < B / 2 || t == 0) ? cconj(Z[(N - k) & (N - 1)]) : cd{0.0, 0.0};
  }
}

// after the barrier: X[k] and X[N - k] of each pair into the row buffer (N + 1 entries), scaled by 1/nz
template <int N>
B200_HD void regfftSplitWriteX(const cd (&v)[RegFft<N>::B], const cd (&pc)[RegFft<N>::B / 2 + 1], cd* X,
                               const cd* __restrict__ tw, int t) {
  constexpr int T = RegFft<N>::T, B = RegFft<N>::B;
  constexpr double sc = 0.5 / (2 * N);
  const cd ws = ldtw(tw, t);
#pragma unroll
  for (int q = 0; q <= B / 2; ++q) {
    if (q < B / 2 || t == 0) {
      const int k = t + T * q;
      const cd zk = v[q], zc = pc[q];
      const cd e = cadd(zk, zc);
      const cd dmi = csub(zk, zc);
      const cd o = {dmi.im, -dmi.re}; // -i (zk - zc)
      const cd wo = cmul(regfftSplitTw<N>(ws, q), o);
      X[k] = cscale(cadd(e, wo), sc);
      if (q < B / 2) {
        X[N - k] = cscale(cconj(csub(e, wo)), sc);
      }
    }
  }
}

// The same with a per-bin function applied to every spectrum value as it is written, mod(k, X[k]) -> stored value:
// a kernel whose per-mode operation is a multiplier (z derivative, filters, phase shift) needs no separate pass over
// the row buffer between the two transforms.
template <int N, class Mod>
B200_HD void regfftSplitWriteXMod(const cd (&v)[RegFft<N>::B], const cd (&pc)[RegFft<N>::B / 2 + 1], cd* X,
                                  const cd* __restrict__ tw, int t, Mod mod) {
  constexpr int T = RegFft<N>::T, B = RegFft<N>::B;
  constexpr double sc = 0.5 / (2 * N);
  const cd ws = ldtw(tw, t);
#pragma unroll
  for (int q = 0; q <= B / 2; ++q) {
    if (q < B / 2 || t == 0) {
      const int k = t + T * q;
      const cd zk = v[q], zc = pc[q];
      const cd e = cadd(zk, zc);
      const cd dmi = csub(zk, zc);
      const cd o = {dmi.im, -dmi.re}; // -i (zk - zc)
      const cd wo = cmul(regfftSplitTw<N>(ws, q), o);
      X[k] = mod(k, q, false, cscale(cadd(e, wo), sc));
      if (q < B / 2) {
        X[N - k] = mod(N - k, q, true, cscale(cconj(csub(e, wo)), sc));
      }
    }
  }
}

// ---- real merge (backward): v[n2] = packed complex input element t + T n2 of the inverse transform -------------
// X: the row buffer (spectrum); bins >= nmode, and the DC bin when zero_dc, read as zero.
template <int N>
B200_HD void regfftMergeRead(cd (&v)[RegFft<N>::B], const cd* X, int nmode, bool zero_dc, const cd* __restrict__ tw,
                             int t) {
  constexpr int T = RegFft<N>::T, B = RegFft<N>::B;
  const cd ws = ldtw(tw, t);
#pragma unroll
  for (int n2 = 0; n2 < B; ++n2) {
    const int k = t + T * n2, kc = N - k;
    cd xk = (k < nmode && !(k == 0 && zero_dc)) ? X[k] : cd{0.0, 0.0};
    cd xc = (kc < nmode) ? cconj(X[kc]) : cd{0.0, 0.0};
    if (k == 0) { // imaginary parts of DC and Nyquist are ignored (FFTW c2r)
      xk.im = 0.0;
      xc.im = 0.0;
    }
    const cd wk = cconj(regfftSplitTw<N>(ws, n2));
    const cd d = cmul(wk, csub(xk, xc));
    v[n2] = cadd(cadd(xk, xc), cd{-d.im, d.re});
  }
}

} // namespace b200
