// The reference's examples/elm-pb/elm_pb.cxx compiled IN PLACE (by path, not copied; TEST
// INFRASTRUCTURE ONLY) with BOUTMAIN turned into a factory (shim/model_factory.hxx), plus read access
// to the model's post-init equilibrium fields and per-RHS auxiliaries so that tests/golden/ can
// record them (J0, P0, phi0, B0, b0xcv, eta, dnorm, ... and phi, Jpar, Jpar2 after an RHS call).
#include <bout/physicsmodel.hxx>
#include <cstring>
#include <string>
#define private public
#define protected public
#include <elm_pb.cxx> // -I$(REF)/examples/elm-pb
#undef private
#undef protected

namespace {
ELMpb* asElm(PhysicsModel* m) { return dynamic_cast<ELMpb*>(m); }
} // namespace

// 0 = ok, 1 = not an ELMpb model, 2 = unknown name.  kind: 2 = Field2D [nx][ny], 3 = Field3D, 0 = scalar
int boutref_elmpb_get_impl(PhysicsModel* pm, const char* cname, double* out, int* kind) {
  ELMpb* m = asElm(pm);
  if (m == nullptr) {
    return 1;
  }
  const std::string n = cname;
  Mesh* msh = bout::globals::mesh;
  const size_t n2 = static_cast<size_t>(msh->LocalNx) * msh->LocalNy;
  const size_t n3 = n2 * msh->LocalNz;
  auto f2 = [&](const Field2D& f) {
    *kind = 2;
    if (out != nullptr) {
      std::memcpy(out, &f(0, 0), sizeof(double) * n2);
    }
    return 0;
  };
  auto f3 = [&](const Field3D& f) {
    *kind = 3;
    if (out != nullptr) {
      std::memcpy(out, &f(0, 0, 0), sizeof(double) * n3);
    }
    return 0;
  };
  auto sc = [&](double v) {
    *kind = 0;
    if (out != nullptr) {
      out[0] = v;
    }
    return 0;
  };
#define F2(x)    \
  if (n == #x) { \
    return f2(m->x); \
  }
#define F3(x)    \
  if (n == #x) { \
    return f3(m->x); \
  }
#define SC(x)    \
  if (n == #x) { \
    return sc(static_cast<double>(m->x)); \
  }
  F2(J0) F2(P0) F2(phi0) F2(B0) F2(Rxy) F2(Bpxy) F2(Btxy) F2(hthe) F2(I)
  if (n == "b0xcv_x") { return f2(m->b0xcv.x); }
  if (n == "b0xcv_y") { return f2(m->b0xcv.y); }
  if (n == "b0xcv_z") { return f2(m->b0xcv.z); }
  F3(eta) F3(vac_mask) F3(phi) F3(Jpar) F3(Jpar2) F3(Psi_loc) F3(U) F3(P) F3(Psi)
  SC(dnorm) SC(hyperresist) SC(ehyperviscos) SC(Bbar) SC(Lbar) SC(Tbar) SC(Va) SC(density)
  SC(vac_resist) SC(core_resist) SC(vacuum_pressure) SC(vacuum_trans) SC(filter_z_mode)
  SC(low_pass_z) SC(filter_z) SC(zonal_flow) SC(zonal_field) SC(zonal_bkgd) SC(smooth_j_x)
  SC(diamag) SC(diamag_phi0) SC(nonlinear) SC(evolve_jpar) SC(evolve_pressure) SC(nogradparj)
  SC(constn0) SC(laplace_perp) SC(jpar_bndry_width) SC(diffusion_p4) SC(diffusion_u4)
  SC(diffusion_a4) SC(viscos_par) SC(viscos_perp) SC(hyperviscos) SC(sink_Ul) SC(sink_Ur)
  SC(bm_exb_flag) SC(bm_mag_flag) SC(include_rmp) SC(compress) SC(eHall) SC(withflow)
  SC(diamag_grad_t) SC(relax_j_vac) SC(gyroviscous) SC(diffusion_par) SC(heating_P) SC(sink_P)
  SC(damp_width) SC(sheath_boundaries) SC(phi_constraint) SC(phi_boundary_relax) SC(split_n0)
#undef F2
#undef F3
#undef SC
  return 2;
}
