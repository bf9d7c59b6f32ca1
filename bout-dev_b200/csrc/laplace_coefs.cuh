// Device-side restatement of Laplacian::tridagCoefs and the boundary rows of
// Laplacian::tridagMatrix (reference src/invert/laplace/invert_laplace.cxx:280-375, :416-768).
// Coefficients are regenerated on chip from the Field2D metrics for every (x, y, kz) instead of
// being stored: the metrics are tiny and L1/L2 resident, the spectral data is not.
#pragma once

#include "fft.cuh"

namespace b200 {

// Boundary / global flags (reference include/bout/invert_laplace.hxx:72-104)
enum : int {
  INV_DC_GRAD = 1, INV_AC_GRAD = 2, INV_AC_LAP = 4, INV_SYM = 8, INV_SET = 16, INV_RHS = 32,
  INV_DC_LAP = 64, INV_BNDRY_ONE = 128, INV_DC_GRADPAR = 256, INV_DC_GRADPARINV = 512,
  INV_IN_CYLINDER = 1024
};
enum : int { INV_ZERO_DC = 1, INV_START_NEW = 2, INV_BOTH_BNDRY_ONE = 4, INV_4TH_ORDER = 8, INV_KX_ZERO = 16 };

struct LapCoefs {
  // Field2D metric arrays [nx][ny]
  const double* __restrict__ dx;
  const double* __restrict__ g11;
  const double* __restrict__ g33;
  const double* __restrict__ g13;
  const double* __restrict__ G1;
  const double* __restrict__ G3;
  const double* __restrict__ g_11; // boundary rows only (may be NULL if unused)
  const double* __restrict__ g_22; // boundary rows only (may be NULL if unused)
  // Laplacian coefficient fields (Field2D) -- any may be NULL
  const double* __restrict__ A;
  const double* __restrict__ C1;
  const double* __restrict__ C2;
  const double* __restrict__ D;
  double zlength;
  int nx, ny;
  int all_terms, nonuniform;
};

// invert_laplace.cxx:280-375 (IncIntShear not supported: the option is off in every config).
// The kz-independent part: k[0] = coef1/dx^2, k[1] = coef2, k[2] = coef3/(2 dx), k[3] = coef4/(2 dx),
// k[4] = coef5, from which a = (k0 - k3, -kwave k2), b = (-2 k0 - kwave^2 k1, kwave k4),
// c = (k0 + k3, kwave k2).
B200_HD void tridagCoef5(const LapCoefs& m, int jx, int jy, bool use_c, bool use_d,
                                            double (&k)[5]) {
  const long long i = static_cast<long long>(jx) * m.ny + jy;
  const double dxi = B200_LDG(m.dx + i);
  const double g11 = B200_LDG(m.g11 + i);
  double coef1 = g11, coef2 = B200_LDG(m.g33 + i);
  const double g13 = B200_LDG(m.g13 + i);
  double coef3 = 2. * g13, coef4 = 0.0, coef5 = 0.0;
  if (m.all_terms) {
    coef4 = B200_LDG(m.G1 + i);
    coef5 = B200_LDG(m.G3 + i);
  }
  if (use_d && m.D != nullptr) {
    const double d = B200_LDG(m.D + i);
    coef1 *= d;
    coef2 *= d;
    coef3 *= d;
    coef4 *= d;
    coef5 *= d;
  }
  if (m.nonuniform) {
    if ((jx != 0) && (jx != (m.nx - 1))) {
      coef4 -= 0.5 * ((B200_LDG(m.dx + i + m.ny) - B200_LDG(m.dx + i - m.ny)) / (dxi * dxi)) * coef1;
    }
  }
  if (use_c && m.C1 != nullptr) {
    if ((jx > 0) && (jx < (m.nx - 1))) {
      const double dc2dx_over_c1 =
          (B200_LDG(m.C2 + i + m.ny) - B200_LDG(m.C2 + i - m.ny)) / (2. * dxi * B200_LDG(m.C1 + i));
      coef4 += g11 * dc2dx_over_c1;
      coef5 += g13 * dc2dx_over_c1;
    }
  }
  coef1 /= dxi * dxi;
  coef3 /= 2. * dxi;
  coef4 /= 2. * dxi;
  k[0] = coef1;
  k[1] = coef2;
  k[2] = coef3;
  k[3] = coef4;
  k[4] = coef5;
}

B200_HD void coefsFromK(const double (&k)[5], double kwave, cd& a, cd& b, cd& c) {
  a = {k[0] - k[3], -kwave * k[2]};
  b = {-2.0 * k[0] - kwave * kwave * k[1], kwave * k[4]};
  c = {k[0] + k[3], kwave * k[2]};
}

B200_HD void tridagCoefs(const LapCoefs& m, int jx, int jy, double kwave,
                                            bool use_c, bool use_d, cd& a, cd& b, cd& c) {
  double k[5];
  tridagCoef5(m, jx, jy, use_c, use_d, k);
  coefsFromK(k, kwave, a, b, c);
}

struct LapBndry {
  int inner_flags, outer_flags, global_flags;
  int inbndry, outbndry; // boundary widths (invert_laplace.cxx:457-469)
  int periodic_x, first_x, last_x;
  int xs, xe;            // x range of the systems on this rank (cyclic_laplace.cxx:88-101)
  int nx_local;          // LocalNx
};

// One matrix row of the system for (x, jy, kz) including the boundary rows; `rhs` is modified
// where tridagMatrix zeroes it.  Row index within the system is x - xs.
B200_HD void tridagRow(const LapCoefs& m, const LapBndry& bd, int x, int jy,
                                          int kz, double kwave, cd& a, cd& b, cd& c, cd& rhs) {
  tridagCoefs(m, x, jy, kwave, true, true, a, b, c);
  if (m.A != nullptr) {
    b.re += B200_LDG(m.A + static_cast<long long>(x) * m.ny + jy);
  }
  const int ix = x - bd.xs;
  const int ncx = bd.xe - bd.xs;
  if (!bd.periodic_x) {
    if (bd.first_x && ix < bd.inbndry) {
      const int fl = bd.inner_flags;
      if (!(fl & (INV_RHS | INV_SET))) {
        rhs = {0.0, 0.0};
      }
      const bool setrhs = (fl & INV_SET) || (fl & INV_RHS);
      const long long i = static_cast<long long>(ix) * m.ny + jy; // NB reference indexes by ix
      double A = 0.0, B = 0.5, C = 0.5;
      if (kz == 0) {
        if ((fl & INV_DC_GRAD) && setrhs) {
          B = -1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
          C = 1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
        } else if (fl & INV_DC_GRAD) {
          B = -1.;
          C = 1.;
        } else if (fl & INV_DC_GRADPAR) {
          B = 1.0 / sqrt(B200_LDG(m.g_22 + i));
          C = -1.0 / sqrt(B200_LDG(m.g_22 + i + m.ny));
        } else if (fl & INV_DC_GRADPARINV) {
          B = sqrt(B200_LDG(m.g_22 + i));
          C = -sqrt(B200_LDG(m.g_22 + i + m.ny));
        } else if (fl & INV_DC_LAP) {
          double k = 0.0;
          if (m.A != nullptr) {
            k = sqrt(-B200_LDG(m.A + static_cast<long long>(bd.inbndry) * m.ny + jy));
          }
          B = 1.;
          C = -exp(-k * B200_LDG(m.dx + i) / sqrt(B200_LDG(m.g11 + i)));
        } else if (fl & INV_IN_CYLINDER) {
          B = 1.;
          C = -1.;
        }
      } else {
        if ((fl & INV_AC_GRAD) && setrhs) {
          B = -1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
          C = 1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
        } else if (fl & INV_AC_GRAD) {
          B = -1.;
          C = 1.;
        } else if (fl & INV_AC_LAP) {
          B = 1.0;
          C = -exp(-1.0 * sqrt(B200_LDG(m.g33 + i) / B200_LDG(m.g11 + i)) * kwave * B200_LDG(m.dx + i));
        } else if (fl & INV_IN_CYLINDER) {
          B = 1.;
          C = ((kz % 2) == 0) ? -1. : 1.;
        }
      }
      a = {A, 0.0};
      b = {B, 0.0};
      c = {C, 0.0};
    }
    if (bd.last_x && (ncx - ix) < bd.outbndry) {
      const int fl = bd.outer_flags;
      const int k_from_edge = ncx - ix; // the reference's loop index "ix" on the outer side
      if (!(fl & (INV_RHS | INV_SET))) {
        rhs = {0.0, 0.0};
      }
      const bool setrhs = (fl & INV_SET) || (fl & INV_RHS);
      const long long i = static_cast<long long>(bd.xe - k_from_edge) * m.ny + jy;
      double A = 0.5, B = 0.5, C = 0.0;
      bool write = true;
      if (kz == 0) {
        if ((fl & INV_DC_GRAD) && setrhs) {
          A = -1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
          B = 1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
        } else if (fl & INV_DC_GRAD) {
          A = 1.;
          B = -1.;
        } else if (fl & (INV_DC_GRADPAR | INV_DC_GRADPARINV | INV_DC_LAP)) {
          // the reference loops these three outer cases over inbndry (invert_laplace.cxx:673-703)
          write = k_from_edge < bd.inbndry;
          if (fl & INV_DC_GRADPAR) {
            A = 1.0 / sqrt(B200_LDG(m.g_22 + i - m.ny));
            B = -1.0 / sqrt(B200_LDG(m.g_22 + i));
          } else if (fl & INV_DC_GRADPARINV) {
            A = sqrt(B200_LDG(m.g_22 + i - m.ny));
            B = -sqrt(B200_LDG(m.g_22 + i));
          } else {
            double k = 0.0;
            if (m.A != nullptr) {
              k = sqrt(-B200_LDG(m.A + static_cast<long long>(bd.inbndry) * m.ny + jy));
            }
            B = 1.;
            A = -exp(-k * B200_LDG(m.dx + i) / sqrt(B200_LDG(m.g11 + i)));
          }
        }
      } else {
        if ((fl & INV_AC_GRAD) && setrhs) {
          A = -1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
          B = 1. / sqrt(B200_LDG(m.g_11 + i)) / B200_LDG(m.dx + i);
        } else if (fl & INV_AC_GRAD) {
          A = 1.;
          B = -1.;
        } else if (fl & INV_AC_LAP) {
          A = -exp(-1.0 * sqrt(B200_LDG(m.g33 + i) / B200_LDG(m.g11 + i)) * kwave * B200_LDG(m.dx + i));
          B = 1.0;
        }
      }
      if (write) {
        a = {A, 0.0};
        b = {B, 0.0};
        c = {C, 0.0};
      }
    }
  } else if (kz == 0 && bd.first_x && ix == 0) {
    // doubly periodic: pin the kz = 0 mode at one point (invert_laplace.cxx:755-767)
    a = {0.0, 0.0};
    b = {1.0, 0.0};
    c = {0.0, 0.0};
    rhs = {0.0, 0.0};
  }
}

} // namespace b200
