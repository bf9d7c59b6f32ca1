// Fast path of the shared-memory z-FFT for the common power-of-two row lengths (nz = 128 .. 1024).
//
// Same transform and normalisation as fft.cuh (the reference's bout::fft::rfft / irfft,
// src/invert/fft_fftw.cxx:116-240), restructured around what ncu showed for the generic version
// (L1/TEX pipe 90 % busy, 8.8 M shared-store bank conflicts per launch at 516x132x512):
//   * every thread owns exactly one radix-4 butterfly per stage (N/4 threads per row), so the
//     twiddle factors of a thread are the same for every row it ever processes: they are loaded
//     once into registers (FastTw) instead of 3 global/L1 loads per butterfly per stage;
//   * the two ping-pong buffers use the XOR swizzle i ^ ((i >> 2) & 7) on 16-byte elements, which
//     makes every load and store of every Stockham stage bank-conflict free (the natural layout
//     has 4-way conflicts on the first stage's stores and 2-way on the second's);
//   * the real-FFT split / merge step works straight from / to global memory, saving one shared
//     memory round trip per row;
//   * CTAs are persistent over rows (grid sized to the SM count), so the one-off twiddle set-up
//     is amortised.
#pragma once

#include "fft.cuh"

namespace b200 {

template <int N>
struct FftShape {
  static_assert(N == 64 || N == 128 || N == 256 || N == 512, "fast path: nz/2 in {64,128,256,512}");
  static constexpr int NST = (N == 64 || N == 128) ? 3 : 4; // radix-4 stages
  static constexpr bool R2 = (N == 128 || N == 512);        // one trailing radix-2 stage
  static constexpr int TPR = N / 4;                         // threads per row
};

__device__ __forceinline__ int fsw(int i) { return i ^ ((i >> 2) & 7); }

template <int N>
struct FastTw {
  cd w[FftShape<N>::NST][3]; // radix-4 stage s (Ns = 4^s): exp(-2 pi i k m / (4 Ns)), m = 1..3, k = lane & (Ns-1)
  cd w2[2];                  // radix-2 stage, butterflies j = lane and lane + N/4
  cd wp[4];                  // real split/merge: exp(-2 pi i k / (2N)) for k = lane + m N/4
};

template <int N>
__device__ __forceinline__ void fastTwInit(FastTw<N>& t, const cd* __restrict__ tw, int lane) {
  constexpr int nz = 2 * N;
  int Ns = 1;
#pragma unroll
  for (int s = 0; s < FftShape<N>::NST; ++s) {
    const int k = lane & (Ns - 1);
    const int tstep = nz / (Ns * 4);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      t.w[s][m] = ldtw(tw, ((m + 1) * k * tstep) & (nz - 1));
    }
    Ns *= 4;
  }
  if (FftShape<N>::R2) {
    const int tstep = nz / (Ns * 2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + h * (N / 4);
      t.w2[h] = ldtw(tw, (j & (Ns - 1)) * tstep);
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    t.wp[m] = ldtw(tw, lane + m * (N / 4));
  }
}

// Barrier policies of the stage loop: the whole CTA, or only the first COUNT threads of it (named
// barrier ID) when the CTA also holds threads that do not take part in the transform.
struct SyncCta {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int ID, int COUNT>
struct SyncNamed {
  static_assert(COUNT % 32 == 0 && ID > 0 && ID < 16, "named barrier: whole warps, id 1..15");
  __device__ __forceinline__ void operator()() const {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
  }
};

// The threads of ONE row only (TPR consecutive threads, rows aligned to TPR): a warp-level barrier when a
// row lives inside one warp, else a named barrier per row group (id 1 + row-in-CTA, at most 15 rows).
// Rows of a CTA then run decoupled instead of stepping through every stage in lock-step.
struct SyncGroup {
  int id, count;
  __device__ __forceinline__ void operator()() const {
    if (count <= 32) {
      __syncwarp();
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
    }
  }
};

// One radix-4 Stockham stage (stage index s, Ns = 4^s) of this thread's butterfly, in -> out.
template <int N, bool INVERSE>
__device__ __forceinline__ void radix4Stage(const cd* in, cd* out, const FastTw<N>& t, int lane, int s, int Ns) {
  const int j = lane;
  const int k = j & (Ns - 1);
  cd v0 = in[fsw(j)], v1 = in[fsw(j + N / 4)], v2 = in[fsw(j + N / 2)], v3 = in[fsw(j + 3 * (N / 4))];
  if (s > 0) {
    cd w1 = t.w[s][0], w2 = t.w[s][1], w3 = t.w[s][2];
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
  const cd t3 = INVERSE ? cd{-d.im, d.re} : cd{d.im, -d.re};
  const int j0 = ((j - k) << 2) + k;
  out[fsw(j0)] = cadd(t0, t2);
  out[fsw(j0 + Ns)] = cadd(t1, t3);
  out[fsw(j0 + 2 * Ns)] = csub(t0, t2);
  out[fsw(j0 + 3 * Ns)] = csub(t1, t3);
}

// The trailing radix-2 stage (N = 128, 512): this thread's two butterflies, in -> out.
template <int N, bool INVERSE>
__device__ __forceinline__ void radix2Stage(const cd* in, cd* out, const FastTw<N>& t, int lane, int Ns) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + h * (N / 4);
    const int k = j & (Ns - 1);
    const cd v0 = in[fsw(j)];
    cd w = t.w2[h];
    if (INVERSE) {
      w = cconj(w);
    }
    const cd v1 = cmul(in[fsw(j + N / 2)], w);
    const int j0 = ((j - k) << 1) + k;
    out[fsw(j0)] = cadd(v0, v1);
    out[fsw(j0 + Ns)] = csub(v0, v1);
  }
}

// Complex FFT of N points between the two swizzled buffers; returns the buffer holding the result.
// All rows of the CTA step through the stages together (barrier = Sync).
template <int N, bool INVERSE, class Sync = SyncCta>
__device__ __forceinline__ cd* cfftFast(cd* in, cd* out, const FastTw<N>& t, int lane, Sync sync = Sync()) {
  int Ns = 1;
#pragma unroll
  for (int s = 0; s < FftShape<N>::NST; ++s) {
    radix4Stage<N, INVERSE>(in, out, t, lane, s, Ns);
    sync();
    cd* tmp = in;
    in = out;
    out = tmp;
    Ns *= 4;
  }
  if (FftShape<N>::R2) {
    radix2Stage<N, INVERSE>(in, out, t, lane, Ns);
    sync();
    cd* tmp = in;
    in = out;
    out = tmp;
  }
  return in;
}

// Two rows per thread group (independent butterflies interleave in the pipeline) with a hook that
// runs once per stage before the stage's barrier: the caller slots unrelated work (Thomas steps of
// another batch) into the shadow of the shared-memory latency.  `live` = false skips the butterflies
// but keeps the hooks and barriers.  Buffers swap every stage; fftSwaps<N>() tells where the result
// is.  (Issuing all loads of a stage before the hook and the stores after it measured 7 % slower.)
template <int N>
__host__ __device__ constexpr int fftSwaps() {
  return FftShape<N>::NST + (FftShape<N>::R2 ? 1 : 0);
}
template <int N, bool INVERSE, class Hook>
__device__ __forceinline__ void cfftFast2(cd* a0, cd* b0, cd* a1, cd* b1, const FastTw<N>& t, int lane, bool live,
                                          Hook hook) {
  int Ns = 1;
#pragma unroll
  for (int s = 0; s < FftShape<N>::NST; ++s) {
    if (live) {
      radix4Stage<N, INVERSE>(a0, b0, t, lane, s, Ns);
      radix4Stage<N, INVERSE>(a1, b1, t, lane, s, Ns);
    }
    hook(s);
    __syncthreads();
    cd* tmp = a0;
    a0 = b0;
    b0 = tmp;
    tmp = a1;
    a1 = b1;
    b1 = tmp;
    Ns *= 4;
  }
  if (FftShape<N>::R2) {
    if (live) {
      radix2Stage<N, INVERSE>(a0, b0, t, lane, Ns);
      radix2Stage<N, INVERSE>(a1, b1, t, lane, Ns);
    }
    hook(FftShape<N>::NST);
    __syncthreads();
  }
}

// ---- register shortcuts ----------------------------------------------------------------------------------
// The first Stockham stage of thread `lane` reads elements lane + m N/4 (m = 0..3): exactly the four elements
// that thread loads from global memory (or produces in the real-FFT merge step), so it can take them from
// registers; and the outputs of the LAST stage of thread `lane` are again elements lane + m N/4, the ones
// it stores to global memory.  Used together they save two shared-memory round trips of the row per transform.
template <int N, bool INVERSE>
__device__ __forceinline__ void radix4Stage0Regs(const cd (&v)[4], cd* out, int lane) {
  const cd t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]);
  const cd d = csub(v[1], v[3]);
  const cd t3 = INVERSE ? cd{-d.im, d.re} : cd{d.im, -d.re};
  const int j0 = lane << 2;
  out[fsw(j0)] = cadd(t0, t2);
  out[fsw(j0 + 1)] = cadd(t1, t3);
  out[fsw(j0 + 2)] = csub(t0, t2);
  out[fsw(j0 + 3)] = csub(t1, t3);
}

// Stages 1 .. of the transform whose stage 0 was radix4Stage0Regs(v, B): B -> A -> B ...; returns the buffer
// holding the result (every stage through shared memory).
template <int N, bool INVERSE, class Sync>
__device__ __forceinline__ cd* cfftFastFromRegs(const cd (&v)[4], cd* A, cd* B, const FastTw<N>& t, int lane, Sync sync) {
  radix4Stage0Regs<N, INVERSE>(v, B, lane);
  sync();
  cd* in = B;
  cd* out = A;
  int Ns = 4;
#pragma unroll
  for (int s = 1; s < FftShape<N>::NST; ++s) {
    radix4Stage<N, INVERSE>(in, out, t, lane, s, Ns);
    sync();
    cd* tmp = in;
    in = out;
    out = tmp;
    Ns *= 4;
  }
  if (FftShape<N>::R2) {
    radix2Stage<N, INVERSE>(in, out, t, lane, Ns);
    sync();
    cd* tmp = in;
    in = out;
    out = tmp;
  }
  return in;
}

// The same with the last stage's outputs kept in registers: r[m] = element lane + m N/4 of the result.
// Ends with a barrier so the buffers can be reused at once.
template <int N, bool INVERSE, class Sync>
__device__ __forceinline__ void cfftFastRegsToRegs(const cd (&v)[4], cd (&r)[4], cd* A, cd* B, const FastTw<N>& t, int lane,
                                                   Sync sync) {
  radix4Stage0Regs<N, INVERSE>(v, B, lane);
  sync();
  cd* in = B;
  cd* out = A;
  int Ns = 4;
  constexpr int NLAST = FftShape<N>::R2 ? FftShape<N>::NST : FftShape<N>::NST - 1; // radix-4 stages through memory
#pragma unroll
  for (int s = 1; s < NLAST; ++s) {
    radix4Stage<N, INVERSE>(in, out, t, lane, s, Ns);
    sync();
    cd* tmp = in;
    in = out;
    out = tmp;
    Ns *= 4;
  }
  if (FftShape<N>::R2) { // last stage: radix 2 with Ns = N/2; butterfly j = lane + h N/4 yields elements j, j + N/2
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + h * (N / 4);
      const cd v0 = in[fsw(j)];
      cd w = t.w2[h];
      if (INVERSE) {
        w = cconj(w);
      }
      const cd v1 = cmul(in[fsw(j + N / 2)], w);
      r[h] = cadd(v0, v1);      // element lane + h N/4
      r[h + 2] = csub(v0, v1);  // element lane + h N/4 + N/2
    }
  } else { // last stage: radix 4 with Ns = N/4 (k = lane): yields elements lane + m N/4
    constexpr int s = FftShape<N>::NST - 1;
    const int j = lane;
    cd v0 = in[fsw(j)], v1 = in[fsw(j + N / 4)], v2 = in[fsw(j + N / 2)], v3 = in[fsw(j + 3 * (N / 4))];
    cd w1 = t.w[s][0], w2 = t.w[s][1], w3 = t.w[s][2];
    if (INVERSE) {
      w1 = cconj(w1);
      w2 = cconj(w2);
      w3 = cconj(w3);
    }
    v1 = cmul(v1, w1);
    v2 = cmul(v2, w2);
    v3 = cmul(v3, w3);
    const cd t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3);
    const cd d = csub(v1, v3);
    const cd t3 = INVERSE ? cd{-d.im, d.re} : cd{d.im, -d.re};
    r[0] = cadd(t0, t2);
    r[1] = cadd(t1, t3);
    r[2] = csub(t0, t2);
    r[3] = csub(t1, t3);
  }
  sync();
}

// Merge step of the backward real transform into registers: v[m] = packed complex input element lane + m N/4.
template <int N, class Load>
__device__ __forceinline__ void rfftMergeRegs(Load load, cd (&v)[4], const FastTw<N>& t, int lane) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = lane + m * (N / 4);
    cd xk = load(k);
    cd xc = cconj(load(N - k));
    if (k == 0) { // imaginary parts of DC and Nyquist are ignored (FFTW c2r)
      xk.im = 0.0;
      xc.im = 0.0;
    }
    const cd wk = cconj(t.wp[m]);
    const cd d = cmul(wk, csub(xk, xc));
    v[m] = cadd(cadd(xk, xc), cd{-d.im, d.re});
  }
}

// Forward: real row (global, nz doubles, 16-byte aligned) -> X[0 .. nmode-1] (global), scaled by 1/nz.
// A, B: this row's swizzled buffers (N + 8 elements each).  All threads of the CTA call together.
template <int N, class Sync = SyncCta>
__device__ __forceinline__ void rfftForwardFast(const double* __restrict__ src, cd* __restrict__ dst, int nmode,
                                                bool active, cd* A, cd* B, const FastTw<N>& t, int lane,
                                                Sync sync = Sync()) {
  const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
  cd v[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const double2 q = __ldg(s2 + lane + m * (N / 4));
    v[m] = {q.x, q.y};
  }
  const cd* Z = cfftFastFromRegs<N, false, Sync>(v, A, B, t, lane, sync);
  constexpr double inv = 1.0 / (2 * N);
  if (active) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = lane + m * (N / 4);
      if (k < nmode) {
        const cd zk = Z[fsw(k)];
        const cd zc = cconj(Z[fsw((N - k) & (N - 1))]);
        const cd e = cscale(cadd(zk, zc), 0.5);
        const cd dmi = csub(zk, zc);
        const cd o = {0.5 * dmi.im, -0.5 * dmi.re};
        dst[k] = cscale(cadd(e, cmul(t.wp[m], o)), inv);
      }
    }
    if (lane == 0 && N < nmode) { // Nyquist bin k = N (w = -1)
      const cd z0 = Z[fsw(0)];
      dst[N] = {(z0.re - z0.im) * inv, 0.0};
    }
  }
  sync(); // the buffers are reused by the next row
}

// Backward: X[0 .. nmode-1] (global; zero beyond nmode; DC dropped when zero_dc) -> real row (global).
template <int N, class Sync = SyncCta>
__device__ __forceinline__ void rfftBackwardFast(const cd* __restrict__ src, double* __restrict__ dst, int nmode,
                                                 bool zero_dc, bool active, cd* A, cd* B, const FastTw<N>& t,
                                                 int lane, Sync sync = Sync()) {
  auto load = [&](int k) -> cd {
    if (k >= nmode || (k == 0 && zero_dc)) {
      return cd{0.0, 0.0};
    }
    const double2 v = __ldg(reinterpret_cast<const double2*>(src) + k);
    return cd{v.x, v.y};
  };
  cd v[4], r[4];
  rfftMergeRegs<N>(load, v, t, lane);
  cfftFastRegsToRegs<N, true, Sync>(v, r, A, B, t, lane, sync);
  if (active) {
    double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      d2[lane + m * (N / 4)] = make_double2(r[m].re, r[m].im);
    }
  }
}

// ---- pieces for kernels that keep the spectrum on chip (fused LaplaceCyclic sweeps) ------------------
// Split step of the forward transform: Z = complex FFT of the packed real row (swizzled buffer) ->
// X[0 .. nmode-1] scaled by 1/nz, written to `dst` (shared or global memory).
template <int N>
__device__ __forceinline__ void rfftSplitTo(const cd* Z, cd* dst, int nmode, const FastTw<N>& t, int lane) {
  constexpr double inv = 1.0 / (2 * N);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = lane + m * (N / 4);
    if (k < nmode) {
      const cd zk = Z[fsw(k)];
      const cd zc = cconj(Z[fsw((N - k) & (N - 1))]);
      const cd e = cscale(cadd(zk, zc), 0.5);
      const cd dmi = csub(zk, zc);
      const cd o = {0.5 * dmi.im, -0.5 * dmi.re};
      dst[k] = cscale(cadd(e, cmul(t.wp[m], o)), inv);
    }
  }
  if (lane == 0 && N < nmode) { // Nyquist bin k = N (w = -1)
    const cd z0 = Z[fsw(0)];
    dst[N] = {(z0.re - z0.im) * inv, 0.0};
  }
}

// Merge step of the backward transform: X[0 .. nmode-1] from `src` (shared memory; zero beyond
// nmode; DC dropped when zero_dc) -> packed complex input of the inverse FFT in swizzled buffer A.
template <int N>
__device__ __forceinline__ void rfftMergeFrom(const cd* src, cd* A, int nmode, bool zero_dc, const FastTw<N>& t,
                                              int lane) {
  auto load = [&](int k) -> cd {
    if (k >= nmode || (k == 0 && zero_dc)) {
      return cd{0.0, 0.0};
    }
    return src[k];
  };
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = lane + m * (N / 4);
    cd xk = load(k);
    cd xc = cconj(load(N - k));
    if (k == 0) { // imaginary parts of DC and Nyquist are ignored (FFTW c2r)
      xk.im = 0.0;
      xc.im = 0.0;
    }
    const cd wk = cconj(t.wp[m]);
    const cd d = cmul(wk, csub(xk, xc));
    A[fsw(k)] = cadd(cadd(xk, xc), cd{-d.im, d.re});
  }
}

} // namespace b200
