// FFTW3-API shim implementation (test infrastructure; see fftw3.h).
// Power-of-two lengths: radix-2 complex FFT of n/2 points + real split; other lengths: direct
// O(n^2) DFT with exactly reduced twiddle indices.  Twiddles are rounded from long double.
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

using cplx = std::complex<double>;

struct bout_b200_fftw_plan_s {
  int n;
  bool forward; // r2c
  double* rbuf;
  fftw_complex* cbuf;
  bool pow2;
  std::vector<cplx> w;       // w[k] = exp(-2 pi i k / n), k = 0..n-1
  std::vector<int> bitrev;   // for the n/2-point complex FFT
  std::vector<cplx> scratch; // n/2 points
};

namespace {
cplx twiddle(int k, int n) {
  const long double a = -2.0L * 3.141592653589793238462643383279502884L * k / n;
  return {static_cast<double>(cosl(a)), static_cast<double>(sinl(a))};
}

// In-place radix-2 decimation-in-time FFT on m points (m power of two).
// stride-2 table lookup: w_m^k = w_n^{2k}.  inverse => conjugated twiddles.
void cfft(bout_b200_fftw_plan_s* p, cplx* z, int m, bool inverse) {
  for (int i = 0; i < m; ++i) {
    const int j = p->bitrev[i];
    if (j > i) {
      std::swap(z[i], z[j]);
    }
  }
  for (int len = 2; len <= m; len <<= 1) {
    const int half = len >> 1;
    const int step = (p->n / len); // index stride into w (w has n entries)
    for (int base = 0; base < m; base += len) {
      for (int k = 0; k < half; ++k) {
        cplx t = p->w[static_cast<size_t>(k) * step];
        if (inverse) {
          t = std::conj(t);
        }
        const cplx u = z[base + k];
        const cplx v = z[base + k + half] * t;
        z[base + k] = u + v;
        z[base + k + half] = u - v;
      }
    }
  }
}

fftw_plan makePlan(int n, double* r, fftw_complex* c, bool forward) {
  auto* p = new bout_b200_fftw_plan_s;
  p->n = n;
  p->forward = forward;
  p->rbuf = r;
  p->cbuf = c;
  p->pow2 = (n >= 4) && ((n & (n - 1)) == 0);
  p->w.resize(n);
  for (int k = 0; k < n; ++k) {
    p->w[k] = twiddle(k, n);
  }
  if (p->pow2) {
    const int m = n / 2;
    int bits = 0;
    while ((1 << bits) < m) {
      ++bits;
    }
    p->bitrev.resize(m);
    for (int i = 0; i < m; ++i) {
      int r2 = 0;
      for (int b = 0; b < bits; ++b) {
        if ((i >> b) & 1) {
          r2 |= 1 << (bits - 1 - b);
        }
      }
      p->bitrev[i] = r2;
    }
    p->scratch.resize(m);
  }
  return p;
}
} // namespace

extern "C" {
void* fftw_malloc(size_t n) { return std::malloc(n); }
void fftw_free(void* p) { std::free(p); }

fftw_plan fftw_plan_dft_r2c_1d(int n, double* in, fftw_complex* out, unsigned) {
  return makePlan(n, in, out, true);
}
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex* in, double* out, unsigned) {
  return makePlan(n, out, in, false);
}
void fftw_destroy_plan(fftw_plan p) { delete p; }

void fftw_execute(const fftw_plan p) {
  const int n = p->n;
  const int nm = n / 2 + 1;
  if (p->forward) {
    const double* x = p->rbuf;
    if (p->pow2) {
      const int m = n / 2;
      cplx* z = p->scratch.data();
      for (int j = 0; j < m; ++j) {
        z[j] = cplx(x[2 * j], x[2 * j + 1]);
      }
      cfft(p, z, m, false);
      for (int k = 0; k <= m; ++k) {
        const cplx zk = z[k % m];
        const cplx zc = std::conj(z[(m - k) % m]);
        const cplx e = 0.5 * (zk + zc);
        const cplx o = cplx(0.0, -0.5) * (zk - zc);
        const cplx wk = (k == m) ? cplx(-1.0, 0.0) : p->w[k];
        const cplx X = e + wk * o;
        p->cbuf[k][0] = X.real();
        p->cbuf[k][1] = X.imag();
      }
    } else {
      for (int k = 0; k < nm; ++k) {
        cplx acc(0.0, 0.0);
        for (int j = 0; j < n; ++j) {
          acc += x[j] * p->w[(static_cast<long>(j) * k) % n];
        }
        p->cbuf[k][0] = acc.real();
        p->cbuf[k][1] = acc.imag();
      }
    }
  } else {
    double* x = p->rbuf;
    if (p->pow2) {
      const int m = n / 2;
      cplx* z = p->scratch.data();
      auto X = [&](int k) {
        if (k == 0 || k == m) {
          return cplx(p->cbuf[k][0], 0.0);
        }
        return cplx(p->cbuf[k][0], p->cbuf[k][1]);
      };
      for (int k = 0; k < m; ++k) {
        const cplx xk = X(k);
        const cplx xc = std::conj(X(m - k));
        const cplx wk = std::conj(p->w[k]);
        z[k] = (xk + xc) + cplx(0.0, 1.0) * wk * (xk - xc);
      }
      cfft(p, z, m, true);
      for (int j = 0; j < m; ++j) {
        x[2 * j] = z[j].real();
        x[2 * j + 1] = z[j].imag();
      }
    } else {
      for (int j = 0; j < n; ++j) {
        double acc = p->cbuf[0][0];
        for (int k = 1; 2 * k < n; ++k) {
          const cplx t = std::conj(p->w[(static_cast<long>(j) * k) % n]);
          acc += 2.0 * (p->cbuf[k][0] * t.real() - p->cbuf[k][1] * t.imag());
        }
        if (n % 2 == 0) {
          acc += (j % 2 == 0) ? p->cbuf[n / 2][0] : -p->cbuf[n / 2][0];
        }
        x[j] = acc;
      }
    }
  }
}
} // extern "C"
