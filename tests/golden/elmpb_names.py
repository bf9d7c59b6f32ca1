"""Names of the elm-pb equilibrium fields / scalars stored in elmpb.npz (shared by the generator and the tests)."""
ELMPB_FIELDS = ["J0", "P0", "phi0", "B0", "b0xcv_x", "b0xcv_y", "b0xcv_z", "eta", "Rxy", "Bpxy", "Btxy",
                "hthe", "I"]
ELMPB_SCALARS = ["dnorm", "hyperresist", "ehyperviscos", "Bbar", "Lbar", "Tbar", "Va", "density",
                 "vac_resist", "core_resist", "filter_z", "filter_z_mode", "low_pass_z", "zonal_flow",
                 "zonal_field", "zonal_bkgd", "smooth_j_x", "diamag", "diamag_phi0", "nonlinear",
                 "evolve_jpar", "evolve_pressure", "nogradparj", "constn0", "laplace_perp", "jpar_bndry_width",
                 "diffusion_p4", "diffusion_u4", "diffusion_a4", "viscos_par", "viscos_perp", "hyperviscos",
                 "sink_Ul", "sink_Ur", "bm_exb_flag", "bm_mag_flag", "include_rmp", "compress", "eHall",
                 "withflow", "diamag_grad_t", "relax_j_vac", "gyroviscous", "diffusion_par", "heating_P",
                 "sink_P", "damp_width", "sheath_boundaries", "phi_constraint", "phi_boundary_relax",
                 "split_n0"]
