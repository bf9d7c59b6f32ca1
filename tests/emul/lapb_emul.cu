// TEST INFRASTRUCTURE (never shipped, never loaded by the product): steps SIMULATED threads through the
// per-thread phases of the block-partitioned LaplaceCyclic path (bout-dev_b200/csrc/laplace_block.cuh,
// fft_reg.cuh) on the host, phase by phase, with barriers exactly where the kernels of laplace_cyclic.cu
// have them.  It checks the index algebra of the register FFT, the separator / reduced-system algebra and
// the change detection on a machine without a GPU; the kernels themselves are checked by `-m gpu`.
// Built by tests/test_lapb_emulation.py with `nvcc -x cu` (host code only is executed).
#include "../../bout-dev_b200/csrc/laplace_block.cuh"

#include <cmath>
#include <cstdlib>
#include <vector>

using namespace b200;

namespace {
std::vector<unsigned char> g_ws;
std::vector<cd> g_tw;
unsigned g_gen = 0;

template <int N, int NT, int MODE>
void emulRows(const LapParams& p) {
  using Sh = LapbShape<N, NT>;
  const int n = p.bd.xe - p.bd.xs + 1, nys = p.ye - p.ys + 1;
  const int nblocks = nys * ((n + Sh::ROWS - 1) / Sh::ROWS);
  std::vector<cd> S(static_cast<size_t>(Sh::ROWS) * Sh::SROW);
  std::vector<double> rowc(Sh::ROWS * Sh::NRC);
  static cd v[NT][Sh::B];
  static cd pc[NT][Sh::B / 2 + 1];
  for (int bid = 0; bid < nblocks; ++bid) {
    LapbCta c = lapbCta<N, NT>(p, bid);
    lapbSetYsel(p, c);
    auto row = [&](int tid) { return S.data() + (tid / Sh::T) * Sh::SROW; };
    for (int tid = 0; tid < NT; ++tid) {
      const int t = tid % Sh::T;
      lapbLoadRow<N, NT, MODE>(p, c, tid, v[tid]);
      if (MODE == LAPB_SOLVE) {
        lapbStageRows<N, NT>(p, c, tid, rowc.data());
      }
      regfftStage1<N, false>(v[tid], p.tw, t);
      regfftXchgWrite<N>(v[tid], row(tid), t);
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftStage2<N, false>(v[tid], row(tid), tid % Sh::T);
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftSplitWriteZ<N>(v[tid], row(tid), tid % Sh::T);
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftSplitReadPartners<N>(pc[tid], row(tid), tid % Sh::T);
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftSplitWriteX<N>(v[tid], pc[tid], row(tid), p.tw, tid % Sh::T);
    }
    for (int tid = 0; tid < NT; ++tid) {
      if (p.nmode > N && tid % Sh::T == 0 && c.row0 + tid / Sh::T <= p.bd.xe - p.bd.xs) {
        double* q = p.bnq + static_cast<long long>(c.iyl) * (p.bd.xe - p.bd.xs + 1) + c.row0 + tid / Sh::T;
        if (MODE == LAPB_DOTS) {
          *q = row(tid)[N].re;
        } else {
          row(tid)[N] = {*q, 0.0};
        }
      }
    }
    // per-mode phase.  The fetches of pass 0 sit before the CTA barrier in the kernel; they read global memory
    // only, so a simulated thread may do them right before its run
    for (int tid = 0; tid < NT; ++tid) {
      int kzs[Sh::KPT];
      bool on[Sh::KPT];
      const int blk = lapbModes<N, NT>(tid, kzs);
      for (int ch = 0; ch < Sh::KPT; ++ch) {
        on[ch] = kzs[ch] < p.nmode;
      }
      if (MODE == LAPB_DOTS) {
        LapbDots<Sh::KPT> st;
        lapbDotsFetch<N, NT, Sh::KPT>(p, c, blk, kzs, 0, st);
        lapbDotsFetchScale<N, NT, Sh::KPT>(p, c, blk, kzs, st);
        lapbDotsRun<N, NT, Sh::KPT>(p, c, blk, kzs, on, st, S.data());
      } else {
        LapbChain<Sh::MB, Sh::KPT> st;
        lapbChainLoad<N, NT, Sh::KPT>(p, c, blk, kzs, st);
        lapbChainRun<N, NT, Sh::KPT>(p, c, blk, kzs, on, st, S.data(), rowc.data());
      }
    }
    if (MODE == LAPB_DOTS) {
      continue;
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftMergeRead<N>(v[tid], row(tid), p.nmode, p.zero_dc != 0, p.tw, tid % Sh::T);
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftStage1<N, true>(v[tid], p.tw, tid % Sh::T);
      regfftXchgWrite<N>(v[tid], row(tid), tid % Sh::T);
    }
    for (int tid = 0; tid < NT; ++tid) {
      regfftStage2<N, true>(v[tid], row(tid), tid % Sh::T);
      lapbStoreRow<N, NT>(p, c, tid, v[tid]);
    }
  }
}

template <int MODE>
void emulRowsN(const LapParams& p) {
  switch (p.nz) {
  case 128: emulRows<64, 256, MODE>(p); break;
  case 256: emulRows<128, 256, MODE>(p); break;
  default:
    if (p.bNT == 128) {
      emulRows<256, 128, MODE>(p);
    } else {
      emulRows<256, 256, MODE>(p);
    }
    break;
  }
}
} // namespace

// All pointers are HOST pointers.  Returns bit 0 = zero pivot seen, bit 1 = the factorisation was rebuilt,
// bit 2 = the tables are per-y; -1 when the configuration is not served by the block path.
// `fresh` != 0 drops the emulated workspace first (a new solver object).
// `nt512`: CTA size of the rows kernels at nz = 512 (256 or 128).
extern "C" int lapb_emul_solve(const bout_b200_grid* g, const bout_b200_laplace* L, const double* rhs,
                               const double* x0, double* out, int fresh, int force, int nt512) {
  if (lapCoupled(L) || !lapbShapeOk(g->nz)) {
    return -1;
  }
  if (lapbThreads512() != nt512) {
    lapbThreads512() = nt512;
    fresh = 1;
  }
  const size_t need = lapLayout(g, L).total;
  if (fresh || g_ws.size() != need) {
    g_ws.assign(need, 0);
    force = 1;
  }
  if (static_cast<int>(g_tw.size()) != g->nz) {
    g_tw.resize(g->nz);
    for (int k = 0; k < g->nz; ++k) {
      const long double a = -2.0L * 3.141592653589793238462643383279502884L * k / g->nz;
      g_tw[k] = {static_cast<double>(cosl(a)), static_cast<double>(sinl(a))};
    }
  }
  LapParams p;
  lapFillParams(g, L, g_ws.data(), &p);
  p.tw = g_tw.data();
  p.rhs = rhs;
  p.x0 = x0;
  p.out = out;
  p.gen = ++g_gen;
  p.force = force;
  for (int i = 0; i < p.nx * p.ny; ++i) {
    lapbCompareThread(p, i);
  }
  int flags = 0;
  if (lapbDirty(p)) {
    flags |= 2;
    for (int i = 0; i < p.nx * p.ny; ++i) {
      lapTableEntry(p, i);
    }
    bool bad = false;
    const long long t1 = static_cast<long long>(p.ye - p.ys + 1) * p.bP * p.nmode;
    for (long long idx = 0; idx < t1; ++idx) {
      bad = lapbFactor1Thread(p, idx) || bad;
    }
    if (p.bP > 1) {
      for (int idx = 0; idx < (p.ye - p.ys + 1) * p.nmode; ++idx) {
        bad = lapbFactor2Thread(p, idx) || bad;
      }
    }
    if (bad) {
      p.status[LAPW_STATUS] = 1;
    }
  }
  if (lapbYdep(p)) {
    flags |= 4;
  }
  if (p.bP > 1 || p.nmode > p.nz / 2) {
    emulRowsN<LAPB_DOTS>(p);
    std::vector<cd> D(p.bP);
    const int nys = p.ye - p.ys + 1, nk = p.nmode < p.nz / 2 ? p.nmode : p.nz / 2;
    for (int idx = 0; idx < nys * nk; ++idx) {
      lapbReducedThread(p, (idx / nk) * p.nmode + idx % nk, D.data(), 1);
    }
    if (p.nmode > p.nz / 2) {
      const int n = p.bd.xe - p.bd.xs + 1;
      std::vector<unsigned char> mem(lapbNyqBytes(n, p.bP) + 64);
      const LapbNyq w = lapbNyqCarve(mem.data(), n, p.bP);
      for (int iyl = 0; iyl < nys; ++iyl) {
        const int ysel = lapbYdep(p) ? iyl : 0;
        for (int tid = 0; tid < 64; ++tid) {
          lapbNyqStage(p, iyl, ysel, tid, 64, w);
        }
        for (int jb = 0; jb < p.bP; ++jb) {
          lapbNyqDots(p, ysel, jb, w);
        }
        lapbNyqSeparators(p, w);
        for (int q = 0; q < 2 * p.bP; ++q) {
          lapbNyqSweepHalf(p, ysel, q >> 1, q & 1, w, w.ends);
        }
        for (int q = 0; q < 2 * p.bP; ++q) {
          lapbNyqSubstHalf(p, ysel, q >> 1, q & 1, w, w.ends);
        }
        for (int i = 0; i < n; ++i) {
          p.bnq[static_cast<long long>(iyl) * n + i] = w.r[i];
        }
      }
    }
  }
  emulRowsN<LAPB_SOLVE>(p);
  return flags | (p.status[LAPW_STATUS] != 0 ? 1 : 0);
}

// Forward real transform of one row followed by the backward one, through the register-FFT phases
// (T simulated threads): spec = rfft(row) (nz/2 + 1 complex, scaled 1/nz), back = irfft(spec).
template <int N>
static void emulRoundTrip(const double* rowin, cd* spec, double* back, const cd* tw) {
  using F = RegFft<N>;
  cd buf[F::SROW];
  cd v[F::T][F::B], pc[F::T][F::B / 2 + 1];
  for (int t = 0; t < F::T; ++t) {
    for (int n2 = 0; n2 < F::B; ++n2) {
      const int j = t + F::T * n2;
      v[t][n2] = {rowin[2 * j], rowin[2 * j + 1]};
    }
    regfftStage1<N, false>(v[t], tw, t);
    regfftXchgWrite<N>(v[t], buf, t);
  }
  for (int t = 0; t < F::T; ++t) regfftStage2<N, false>(v[t], buf, t);
  for (int t = 0; t < F::T; ++t) regfftSplitWriteZ<N>(v[t], buf, t);
  for (int t = 0; t < F::T; ++t) regfftSplitReadPartners<N>(pc[t], buf, t);
  for (int t = 0; t < F::T; ++t) regfftSplitWriteX<N>(v[t], pc[t], buf, tw, t);
  for (int k = 0; k <= N; ++k) spec[k] = buf[k];
  for (int t = 0; t < F::T; ++t) regfftMergeRead<N>(v[t], buf, N + 1, false, tw, t);
  for (int t = 0; t < F::T; ++t) {
    regfftStage1<N, true>(v[t], tw, t);
    regfftXchgWrite<N>(v[t], buf, t);
  }
  for (int t = 0; t < F::T; ++t) {
    regfftStage2<N, true>(v[t], buf, t);
    for (int q = 0; q < F::B; ++q) {
      const int j = t + F::T * q;
      back[2 * j] = v[t][q].re;
      back[2 * j + 1] = v[t][q].im;
    }
  }
}

extern "C" int regfft_emul_roundtrip(int nz, const double* rowin, double* spec, double* back) {
  std::vector<cd> tw(nz);
  for (int k = 0; k < nz; ++k) {
    const long double a = -2.0L * 3.141592653589793238462643383279502884L * k / nz;
    tw[k] = {static_cast<double>(cosl(a)), static_cast<double>(sinl(a))};
  }
  cd* sp = reinterpret_cast<cd*>(spec);
  switch (nz) {
  case 128: emulRoundTrip<64>(rowin, sp, back, tw.data()); return 0;
  case 256: emulRoundTrip<128>(rowin, sp, back, tw.data()); return 0;
  case 512: emulRoundTrip<256>(rowin, sp, back, tw.data()); return 0;
  default: return -1;
  }
}
