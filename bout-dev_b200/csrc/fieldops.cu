// Field algebra (subsystem a12): element-wise expressions in one pass.
//
// The reference evaluates `2.*a + b*c` as three loops and two temporaries
// (src/field/generated_fieldops.cxx: one BOUT_FOR per operator).  Here an expression is a short sum
// of products  out = sum_i coef_i * f1_i * f2_i [/ d_i]  evaluated in a single kernel; every
// operator of generated_fieldops.cxx (Field3D (+,-,*,/) {Field3D, Field2D, BoutReal}) is one or two
// such terms.  Field2D operands broadcast over z; division by a Field2D or a scalar multiplies by
// the reciprocal exactly as the reference does (generated_fieldops.cxx:253-274, :543-558).
#include "common.cuh"

namespace b200 {

constexpr int kMaxTerms = 8;

struct EwTerm {
  double coef;
  const double* f1;
  const double* f2;
  const double* div;
  int f1_2d, f2_2d, div_2d;
};

struct EwParams {
  EwTerm t[kMaxTerms];
  int nterms;
  double* out;
  int ny, nz, xs, xe, ys, ye;
};

template <int VEC>
__global__ void __launch_bounds__(256) k_elementwise(const EwParams p) {
  const int nzv = p.nz / VEC;
  const int nyb = p.ye - p.ys + 1;
  const long long total = static_cast<long long>(p.xe - p.xs + 1) * nyb * nzv;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    int zv, yb_, xb_;
    splitIndex(t, nzv, nyb, zv, yb_, xb_);
    const int y = p.ys + yb_, x = p.xs + xb_;
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    const long long i3 = i2 * p.nz + static_cast<long long>(zv) * VEC;
    Vz<VEC> acc;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      acc.v[k] = 0.0;
    }
    for (int it = 0; it < p.nterms; ++it) {
      const EwTerm& tm = p.t[it];
      Vz<VEC> v;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        v.v[k] = tm.coef;
      }
      if (tm.f1 != nullptr) {
        if (tm.f1_2d) {
          const double s = __ldg(tm.f1 + i2);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            v.v[k] = (tm.coef == 1.0) ? s : v.v[k] * s;
          }
        } else {
          const Vz<VEC> a = Vz<VEC>::load(tm.f1 + i3);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            v.v[k] = (tm.coef == 1.0) ? a.v[k] : v.v[k] * a.v[k];
          }
        }
      }
      if (tm.f2 != nullptr) {
        if (tm.f2_2d) {
          const double s = __ldg(tm.f2 + i2);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            v.v[k] *= s;
          }
        } else {
          const Vz<VEC> a = Vz<VEC>::load(tm.f2 + i3);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            v.v[k] *= a.v[k];
          }
        }
      }
      if (tm.div != nullptr) {
        if (tm.div_2d) {
          const double s = 1.0 / __ldg(tm.div + i2);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            v.v[k] *= s;
          }
        } else {
          const Vz<VEC> a = Vz<VEC>::load(tm.div + i3);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            v.v[k] /= a.v[k];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        acc.v[k] = (it == 0) ? v.v[k] : acc.v[k] + v.v[k];
      }
    }
    acc.store(p.out + i3);
  }
}

enum UnaryOp { U_NEG = 0, U_SQRT, U_ABS, U_EXP, U_LOG, U_SIN, U_COS, U_TANH, U_SQ, U_POW, U_COPY };

__global__ void __launch_bounds__(256) k_unary(int op, const double* __restrict__ f, double arg,
                                               double* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double v = f[i];
    double r;
    switch (op) {
    case U_NEG: r = -1.0 * v; break; // field3d.cxx:581 (-1.0 * f)
    case U_SQRT: r = sqrt(v); break;
    case U_ABS: r = fabs(v); break;
    case U_EXP: r = exp(v); break;
    case U_LOG: r = log(v); break;
    case U_SIN: r = sin(v); break;
    case U_COS: r = cos(v); break;
    case U_TANH: r = tanh(v); break;
    case U_SQ: r = v * v; break;
    case U_POW: r = pow(v, arg); break;
    default: r = v; break;
    }
    out[i] = r;
  }
}

} // namespace b200

using namespace b200;

extern "C" int bout_b200_elementwise(const bout_b200_grid* g, int region, int nterms,
                                     const bout_b200_ewterm* terms, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(terms != nullptr && out != nullptr && nterms >= 1 && nterms <= kMaxTerms,
               BOUT_B200_ERR_ARG, "elementwise: 1..%d terms", kMaxTerms);
  Box b;
  if (regionBox(g, region, &b) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  EwParams p{};
  bool vec2 = (g->nz % 2 == 0) && aligned16(out);
  for (int i = 0; i < nterms; ++i) {
    p.t[i] = {terms[i].coef, terms[i].f1, terms[i].f2, terms[i].div,
              terms[i].f1_is2d, terms[i].f2_is2d, terms[i].div_is2d};
    vec2 = vec2 && (terms[i].f1_is2d || aligned16(terms[i].f1)) && (terms[i].f2_is2d || aligned16(terms[i].f2))
           && (terms[i].div_is2d || aligned16(terms[i].div));
  }
  p.nterms = nterms;
  p.out = out;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = b.xs;
  p.xe = b.xe;
  p.ys = b.ys;
  p.ye = b.ye;
  const long long total = static_cast<long long>(b.xe - b.xs + 1) * (b.ye - b.ys + 1) * (vec2 ? g->nz / 2 : g->nz);
  const long long want = (total + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 64 ? (want < 1 ? 1 : want) : 148LL * 64);
  if (vec2) {
    k_elementwise<2><<<blocks, 256, 0, asStream(stream)>>>(p);
  } else {
    k_elementwise<1><<<blocks, 256, 0, asStream(stream)>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_unary(const bout_b200_grid* g, int op, const double* f, double arg,
                               double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && out && op >= 0 && op <= U_COPY, BOUT_B200_ERR_ARG, "unary: bad argument");
  const long long n = static_cast<long long>(g->nx) * g->ny * g->nz;
  const long long want = (n + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 64 ? want : 148LL * 64);
  k_unary<<<blocks, 256, 0, asStream(stream)>>>(op, f, arg, out, n);
  B200_LAUNCH_OK();
  return 0;
}

namespace b200 {
__global__ void __launch_bounds__(256) k_where(const double* __restrict__ test, int test_is2d, const double* __restrict__ a,
                                               double av, const double* __restrict__ b, double bv, double* __restrict__ out,
                                               long long n, int nz) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double t = __ldg(test + (test_is2d ? i / nz : i));
    out[i] = t > 0.0 ? (a != nullptr ? __ldg(a + i) : av) : (b != nullptr ? __ldg(b + i) : bv);
  }
}
} // namespace b200

extern "C" int bout_b200_where(const bout_b200_grid* g, const double* test, int test_is2d, const double* gt0,
                               double gt0_value, const double* le0, double le0_value, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(test && out && out != test, BOUT_B200_ERR_ARG, "where: bad argument");
  const long long n = static_cast<long long>(g->nx) * g->ny * g->nz;
  const long long want = (n + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 64 ? want : 148LL * 64);
  b200::k_where<<<blocks, 256, 0, asStream(stream)>>>(test, test_is2d, gt0, gt0_value, le0, le0_value, out, n, g->nz);
  B200_LAUNCH_OK();
  return 0;
}

// Linear combination of fields, the stage / output-state loops of the explicit integrators beyond rk4 and rk3ssp
// (euler.cxx:178-181, rkscheme.cxx:57-84 and :173-190, split-rk.cxx:226-300): one pass through k_elementwise.
extern "C" int bout_b200_lincomb(const bout_b200_grid* g, int region, int n, const double* coef,
                                 const double* const* x, double* out, void* stream) {
  B200_REQUIRE(coef != nullptr && x != nullptr && n >= 1 && n <= kMaxTerms, BOUT_B200_ERR_ARG,
               "lincomb: 1..%d terms", kMaxTerms);
  bout_b200_ewterm terms[kMaxTerms];
  for (int i = 0; i < n; ++i) {
    B200_REQUIRE(x[i] != nullptr && x[i] != out, BOUT_B200_ERR_ARG, "lincomb: operand %d is null or aliases out", i);
    terms[i] = {coef[i], x[i], nullptr, nullptr, 0, 0, 0};
  }
  return bout_b200_elementwise(g, region, n, terms, out, stream);
}
