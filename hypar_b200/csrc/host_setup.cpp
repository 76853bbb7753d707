// host_setup.cpp -- host-side set-up of one rank: the C restatement of what HyPar's start-up does
// for the data the hot path consumes. Pure host code (no CUDA calls), usable without a GPU.
//
//   partition / rank maps   reference src/MPIFunctions/MPIPartition1D.c, MPIRanknD.c, MPIRank1D.c
//   ghost coordinates       reference src/IOFunctions/ReadArray.c:60-100
//   dxinv                   reference src/Simulation/InitialSolution.c:74-119
//   neighbours, bcperiodic  reference src/MPIFunctions/MPIExchangeBoundariesnD.c:65-76,
//                           src/Simulation/InitializeBoundaries.c:190-201
//   zone extents            reference src/Simulation/InitializeBoundaries.c:380-440, MathFunctions/FindInterval.c
//   gravity field           reference src/PhysicalModels/NavierStokes3D/NavierStokes3DGravityField.c:33-152
//   RK tableaux             reference src/TimeIntegration/TimeExplicitRKInitialize.c:58-79
#include <math.h>
#include <string.h>
#include "hpb_internal.h"
#include "glmgee_tables.h"

extern "C" int hpb_partition1d(int nglobal, int nproc, int rank)
{
  int n = nglobal / nproc;
  if (rank == nproc - 1) n = nglobal - n * (nproc - 1);   // remainder to the last rank
  return n;
}

extern "C" int hpb_rank1d(int ndims, const int* iproc, const int* ip)
{
  int r = 0, f = 1;
  for (int d = 0; d < ndims; d++) { r += f * ip[d]; f *= iproc[d]; }
  return r;
}

extern "C" void hpb_ranknd(int ndims, int rank, const int* iproc, int* ip)
{
  for (int d = 0; d < ndims; d++) { ip[d] = rank % iproc[d]; rank /= iproc[d]; }
}

static void find_interval(double a, double b, const double* x, int N, int* imin, int* imax)
{
  *imax = -1; *imin = N;
  double min_dx = x[1] - x[0];
  for (int i = 2; i < N; i++) { double dx = x[i] - x[i-1]; if (dx < min_dx) min_dx = dx; }
  double tol = 1e-10 * min_dx;
  for (int i = 0; i < N; i++) if (x[i] <= (b + tol)) *imax = i + 1;
  for (int i = N - 1; i > -1; i--) if (x[i] >= (a - tol)) *imin = i;
}

static inline double raiseto(double x, double a) { return exp(a * log(x)); }   // math_ops.h:37

int hpb_setup_host(hpb_solver* h)
{
  const hpb_config& c = h->cfg;
  const int nd = c.ndims, g = c.ghosts;

  // ---- validation: what the device path implements (anything else fails loudly)
  if (nd < 1 || nd > 3) return hpb_fail(HPB_ERR_INVALID, "ndims = %d not supported (1..3)", nd);
  if (g != HPB_G) return hpb_fail(HPB_ERR_INVALID, "ghost = %d: the WENO5 device path needs exactly %d ghost layers", g, HPB_G);
  static const int model_nv[5] = { -1, 3, 4, 5, -1 }, model_nd[5] = { -1, 1, 2, 3, -1 };
  if (c.model < 0 || c.model > HPB_MODEL_BURGERS) return hpb_fail(HPB_ERR_INVALID, "unknown model id %d", c.model);
  if (c.model == HPB_MODEL_BURGERS && c.nvars != 1) return hpb_fail(HPB_ERR_INVALID, "burgers: nvars must be 1");
  if (c.model != HPB_MODEL_LINEAR_ADR && c.model != HPB_MODEL_BURGERS && (c.nvars != model_nv[c.model] || nd != model_nd[c.model]))
    return hpb_fail(HPB_ERR_INVALID, "model %d needs ndims=%d nvars=%d (got %d, %d)", c.model, model_nd[c.model],
                    model_nv[c.model], nd, c.nvars);
  if (c.nvars < 1 || c.nvars > HPB_MAX_NVARS) return hpb_fail(HPB_ERR_INVALID, "nvars = %d not supported", c.nvars);
  if (c.weno_type < 0 || c.weno_type > 3) return hpb_fail(HPB_ERR_INVALID, "unknown WENO weight type %d", c.weno_type);
  const bool glm = (c.rk_type >= HPB_GLMGEE_23 && c.rk_type <= HPB_GLMGEE_RK285EX);
  if ((c.rk_type < HPB_RK_44 || c.rk_type > HPB_RK_33) && !glm)
    return hpb_fail(HPB_ERR_INVALID, "time_scheme_type %d not supported (rk 1fe, 22, 33, 44, ssprk3; glm-gee 23, 24, 25i, 35, "
                                     "exrk2a, rk32g1, rk285ex)", c.rk_type);
  if (glm && c.glm_ee_mode != HPB_GLM_YEPS && c.glm_ee_mode != HPB_GLM_YYT)
    return hpb_fail(HPB_ERR_INVALID, "glm-gee: ee_mode %d (yeps, yyt)", c.glm_ee_mode);
  if ((c.model == HPB_MODEL_NS3D || c.model == HPB_MODEL_NS2D || c.model == HPB_MODEL_EULER1D) &&
      (c.upwind < HPB_UPWIND_ROE || c.upwind > HPB_UPWIND_LLF))
    return hpb_fail(HPB_ERR_INVALID, "upwinding %d not implemented (roe, rusanov, rf-char, llf-char)", c.upwind);
  const bool has_grav = (c.gravity[0] != 0.0 || c.gravity[1] != 0.0 || c.gravity[2] != 0.0);
  if (has_grav && (c.model == HPB_MODEL_LINEAR_ADR || c.model == HPB_MODEL_BURGERS))
    return hpb_fail(HPB_ERR_INVALID, "gravity needs an Euler / Navier-Stokes model");
  if (has_grav && c.model == HPB_MODEL_EULER1D && c.upwind != HPB_UPWIND_LLF && c.upwind != HPB_UPWIND_ROE)   // Euler1DInitialize.c:144-149
    return hpb_fail(HPB_ERR_INVALID, "llf-char or roe upwinding is needed for flows with gravitational forces");
  if (c.model == HPB_MODEL_EULER1D && (c.gravity[1] != 0.0 || c.gravity[2] != 0.0 || c.gravity_type < 0 || c.gravity_type > 1))
    return hpb_fail(HPB_ERR_INVALID, "euler1d: one gravity component, gravity_type 0 or 1");
  if (has_grav && c.model == HPB_MODEL_NS3D && c.upwind != HPB_UPWIND_RUSANOV)      // NavierStokes3DInitialize.c:371-378
    return hpb_fail(HPB_ERR_INVALID, "rusanov upwinding is needed for flows with gravitational forces");
  if (has_grav && c.model == HPB_MODEL_NS2D && c.upwind != HPB_UPWIND_RUSANOV && c.upwind != HPB_UPWIND_ROE &&
      c.upwind != HPB_UPWIND_LLF)                                                   // NavierStokes2DInitialize.c:207-216
    return hpb_fail(HPB_ERR_INVALID, "llf-char, roe or rusanov upwinding is needed for flows with gravitational forces");
  if (c.model == HPB_MODEL_NS2D && c.gravity[2] != 0.0)
    return hpb_fail(HPB_ERR_INVALID, "navierstokes2d: gravity has two components");
  if (c.model == HPB_MODEL_LINEAR_ADR && c.par_scheme != 2 && c.par_scheme != 4)
    return hpb_fail(HPB_ERR_INVALID, "par_space_scheme %d not supported (2, 4)", c.par_scheme);
  if ((c.model == HPB_MODEL_NS2D || c.model == HPB_MODEL_NS3D) && c.Re > 0 && c.par_scheme != 4)
    // InitializeSolvers.c:128-141: par_space_scheme 2 makes FirstDerivativePar the FIRST-order difference; the device's viscous
    // kernels are FirstDerivativeFourthOrderCentral
    return hpb_fail(HPB_ERR_INVALID, "viscous terms are on the B200 path with par_space_scheme 4 only (got %d)", c.par_scheme);
  if (c.par_space_type < HPB_PAR_NC_1STAGE || c.par_space_type > HPB_PAR_CONS_1STAGE)
    return hpb_fail(HPB_ERR_INVALID, "unknown par_space_type %d", c.par_space_type);
  if (c.model == HPB_MODEL_LINEAR_ADR && c.par_space_type != HPB_PAR_NC_1STAGE)
    for (int i = 0; i < nd * c.nvars; i++)
      if (c.diffusion[i] != 0.0)       // LinearADR installs GFunction AND HFunction: every form is a different discretisation
        return hpb_fail(HPB_ERR_INVALID, "LinearADR diffusion is on the B200 path as par_space_type nonconservative-1stage only");
  if (c.hyp_scheme < HPB_SCHEME_WENO5 || c.hyp_scheme > HPB_SCHEME_HCWENO5)
    return hpb_fail(HPB_ERR_INVALID, "hyp_space_scheme %d not supported (weno5, crweno5, hcweno5, cupw5, upw5, 1, 2, 4, muscl2, muscl3)",
                    c.hyp_scheme);
  if (c.muscl_limiter < HPB_LIMITER_GMM || c.muscl_limiter > HPB_LIMITER_SUPERBEE)
    return hpb_fail(HPB_ERR_INVALID, "muscl limiter %d not supported (gmm, minmod, vanleer, superbee)", c.muscl_limiter);
  if (hpb_scheme_is_compact(c.hyp_scheme)) {
    bool split = false;
    for (int d = 0; d < nd; d++) split = split || c.iproc[d] > 1;
    if (split && c.lu_gather_and_solve)
      return hpb_fail(HPB_ERR_INVALID, "lusolver.inp: reducedsolvetype gather-and-solve is not on the B200 path (jacobi)");
    if (c.lu_maxiter < 0) return hpb_fail(HPB_ERR_INVALID, "lusolver.inp: maxiter %d", c.lu_maxiter);
  }
  if (hpb_scheme_is_compact(c.hyp_scheme))
    for (int d = 0; d < nd; d++)
      if (c.iproc[d] > 64) return hpb_fail(HPB_ERR_INVALID, "compact schemes: at most 64 ranks along one dimension");
  if (c.nzones > HPB_MAX_ZONES) return hpb_fail(HPB_ERR_INVALID, "too many boundary zones");
  int nranks = 1;
  for (int d = 0; d < nd; d++) {
    if (c.iproc[d] < 1) return hpb_fail(HPB_ERR_INVALID, "iproc[%d] < 1", d);
    nranks *= c.iproc[d];
  }
  if (c.rank < 0 || c.rank >= nranks) return hpb_fail(HPB_ERR_INVALID, "rank %d outside iproc grid (%d ranks)", c.rank, nranks);
  if (!c.x_global) return hpb_fail(HPB_ERR_INVALID, "x_global is NULL");

  // ---- partition
  hpb_ranknd(nd, c.rank, c.iproc, h->ip);
  Geom& G = h->geo;
  memset(&G, 0, sizeof(G));
  G.ndims = nd; G.nvars = c.nvars; G.g = g;
  for (int d = 0; d < 3; d++) { G.N[d] = 1; G.P[d] = 1; h->is_global[d] = 0; }
  int xo = 0;
  for (int d = 0; d < nd; d++) {
    G.N[d] = hpb_partition1d(c.dim_global[d], c.iproc[d], h->ip[d]);
    if (G.N[d] < g) return hpb_fail(HPB_ERR_INVALID, "local size %d along dim %d is smaller than the number of ghost layers", G.N[d], d);   // quasi-1-D runs (size 3, ghost 3) are in the reference's Examples
    h->is_global[d] = (c.dim_global[d] / c.iproc[d]) * h->ip[d];
    G.P[d] = G.N[d] + 2 * g;
    G.xoff[d] = xo;
    xo += G.P[d];
  }
  for (int d = 0; d < 3; d++) { G.lo_phys[d] = 1; G.hi_phys[d] = 1; }
  for (int d = 0; d < nd; d++) { G.lo_phys[d] = (h->ip[d] == 0); G.hi_phys[d] = (h->ip[d] == c.iproc[d] - 1); }
  G.st[0] = 1; G.st[1] = G.P[0]; G.st[2] = (long long)G.P[0] * G.P[1];
  G.npg = (long long)G.P[0] * G.P[1] * G.P[2];

  // ---- periodic flags and neighbours
  for (int d = 0; d < 3; d++) h->bcperiodic[d] = 0;
  for (int n = 0; n < c.nzones; n++) {
    const hpb_boundary_zone& z = c.zones[n];
    if (z.dim < 0 || z.dim >= nd) return hpb_fail(HPB_ERR_INVALID, "boundary zone %d: dim %d is invalid (ndims = %d)", n, z.dim, nd);
    if (z.face != 1 && z.face != -1) return hpb_fail(HPB_ERR_INVALID, "boundary zone %d: face must be +1/-1", n);
    if (z.type < 0 || z.type > HPB_BC_SPONGE)
      return hpb_fail(HPB_ERR_INVALID, "boundary zone %d: boundary type %d not implemented", n, z.type);
    if (z.type == HPB_BC_SLIP_WALL && (c.model == HPB_MODEL_LINEAR_ADR || c.model == HPB_MODEL_BURGERS))
      return hpb_fail(HPB_ERR_INVALID, "slip-wall needs an Euler/Navier-Stokes model");
    if ((z.type == HPB_BC_NOSLIP_WALL || (z.type >= HPB_BC_SUBSONIC_INFLOW && z.type <= HPB_BC_SUPERSONIC_OUTFLOW)) &&
        c.model != HPB_MODEL_NS2D && c.model != HPB_MODEL_NS3D)      // the reference has 2-D and 3-D branches only
      return hpb_fail(HPB_ERR_INVALID, "boundary zone %d: boundary type %d needs a 2-D or 3-D Navier-Stokes model", n, z.type);
    if (z.type == HPB_BC_PERIODIC && c.iproc[z.dim] > 1) h->bcperiodic[z.dim] = 1;
  }
  for (int d = 0; d < 3; d++) h->neighbor[2*d] = h->neighbor[2*d+1] = -1;
  for (int d = 0; d < nd; d++) {
    int nip[3] = { h->ip[0], h->ip[1], h->ip[2] };
    if (h->ip[d] == 0) nip[d] = c.iproc[d] - 1; else nip[d]--;
    if (!(h->ip[d] == 0 && !h->bcperiodic[d])) h->neighbor[2*d] = hpb_rank1d(nd, c.iproc, nip);
    nip[d] = h->ip[d];
    if (h->ip[d] == c.iproc[d] - 1) nip[d] = 0; else nip[d]++;
    if (!(h->ip[d] == c.iproc[d] - 1 && !h->bcperiodic[d])) h->neighbor[2*d+1] = hpb_rank1d(nd, c.iproc, nip);
    if (c.iproc[d] == 1) h->neighbor[2*d] = h->neighbor[2*d+1] = -1;   // bcperiodic is 0 then
  }

  // ---- coordinates with ghosts and dxinv
  h->x_h.assign(xo, 0.0);
  h->dxinv_h.assign(xo, 0.0);
  int goff = 0;
  for (int d = 0; d < nd; d++) {
    const int n = G.N[d], ng = c.dim_global[d], i0 = h->is_global[d];
    const double* xg = c.x_global + goff;
    double* X = h->x_h.data() + G.xoff[d];
    for (int i = 0; i < n; i++) X[g + i] = xg[i0 + i];
    if (h->ip[d] == 0) {
      for (int i = 0; i < g; i++) { int delta = g - i; X[i] = X[g] + ((double)delta) * (X[g] - X[g+1]); }
    } else {
      for (int i = 0; i < g; i++) X[i] = xg[i0 - g + i];
    }
    if (h->ip[d] == c.iproc[d] - 1) {
      for (int i = n + g; i < n + 2*g; i++) { int delta = i - (n + g - 1); X[i] = X[n+g-1] + ((double)delta) * (X[n+g-1] - X[n+g-2]); }
    } else {
      for (int i = 0; i < g; i++) X[n + g + i] = xg[i0 + n + i];
    }
    // dxinv over the globally extended grid: every owner computes 2/(x[i+1]-x[i-1]) from its local x,
    // whose internal-face ghosts are the neighbour's true coordinates
    std::vector<double> xe(ng + 2*g), de(ng + 2*g, 0.0);
    for (int i = 0; i < ng; i++) xe[g + i] = xg[i];
    for (int i = 0; i < g; i++) {
      xe[i] = xg[0] + ((double)(g - i)) * (xg[0] - xg[1]);
      xe[ng + g + i] = xg[ng-1] + ((double)(i + 1)) * (xg[ng-1] - xg[ng-2]);
    }
    for (int i = 0; i < ng; i++) de[g + i] = 2.0 / (xe[g + i + 1] - xe[g + i - 1]);
    double* DX = h->dxinv_h.data() + G.xoff[d];
    for (int i = 0; i < n + 2*g; i++) DX[i] = de[i0 + i];
    if (h->ip[d] == 0) for (int i = 0; i < g; i++) DX[i] = DX[g];
    if (h->ip[d] == c.iproc[d] - 1) for (int i = n + g; i < n + 2*g; i++) DX[i] = DX[n + g - 1];
    goff += ng;
  }

  // ---- boundary zone extents
  h->zones.clear();
  for (int n = 0; n < c.nzones; n++) {
    const hpb_boundary_zone& z = c.zones[n];
    ZoneDev zd; memset(&zd, 0, sizeof(zd));
    zd.type = z.type; zd.dim = z.dim; zd.face = z.face; zd.on = 0;
    for (int d = 0; d < 3; d++) { zd.is[d] = 0; zd.ie[d] = 1; zd.wall[d] = z.wall_velocity[d]; }
    zd.rho = z.flow_density; zd.pressure = z.flow_pressure;
    for (int v = 0; v < HPB_MAX_NVARS; v++) zd.val[v] = z.dirichlet[v];
    zd.xs = z.xmin[z.dim]; zd.xe = z.xmax[z.dim];
    const bool edge = (z.face == 1) ? (h->ip[z.dim] == 0) : (h->ip[z.dim] == c.iproc[z.dim] - 1);
    if (z.type == HPB_BC_SPONGE) {            // InitializeBoundaries.c:381-395: an interior box on every rank it overlaps
      zd.on = 1;
      for (int d = 0; d < nd; d++) {
        int is, ie;
        find_interval(z.xmin[d], z.xmax[d], h->x_h.data() + G.xoff[d] + g, G.N[d], &is, &ie);
        zd.is[d] = is; zd.ie[d] = ie;
        if ((ie - is) <= 0) zd.on = 0;
      }
    } else if (edge) {
      zd.on = 1;
      for (int d = 0; d < nd; d++) {
        if (d == z.dim) {
          if (z.face == 1) { zd.is[d] = -g; zd.ie[d] = 0; }
          else             { zd.is[d] = G.N[d]; zd.ie[d] = G.N[d] + g; }
        } else {
          int is, ie;
          find_interval(z.xmin[d], z.xmax[d], h->x_h.data() + G.xoff[d] + g, G.N[d], &is, &ie);
          zd.is[d] = is; zd.ie[d] = ie;
          if ((ie - is) <= 0) zd.on = 0;
        }
      }
    }
    h->zones.push_back(zd);
  }

  // ---- physics parameters
  Phys& P = h->phys;
  memset(&P, 0, sizeof(P));
  P.model = c.model; P.weno = c.weno_type; P.no_limiting = c.no_limiting;
  P.interp_char = (c.interp_char && c.nvars > 1) ? 1 : 0;        // WENOInitialize.c:156
  P.upwind = c.upwind; P.par_scheme = c.par_scheme; P.has_grav = has_grav ? 1 : 0;
  P.scheme = c.hyp_scheme; P.muscl_limiter = c.muscl_limiter; P.muscl_eps = c.muscl_eps;
  P.hc_rc = c.weno_rc; P.hc_xi = c.weno_xi;
  P.eps = c.weno_eps; P.gamma = c.gamma;
  P.Re = c.Re / c.Minf;                                          // NavierStokes3DInitialize.c:368
  P.Pr = c.Pr;
  P.RT = c.p_ref / c.rho_ref;
  for (int d = 0; d < 3; d++) P.grav[d] = c.gravity[d];
  for (int i = 0; i < 15; i++) { P.adv[i] = c.advection[i]; P.diff[i] = c.diffusion[i]; }

  // ---- gravity field (NavierStokes3DGravityField.c:33-152; NavierStokes2DGravityField.c:33-150 is the same with
  // (gx x + gy y) and y as the vertical of HB 3); 1.0 everywhere for the other models
  h->gravf_h.assign((size_t)G.npg, 1.0);
  h->gravg_h.assign((size_t)G.npg, 1.0);
  if (c.model == HPB_MODEL_NS3D || c.model == HPB_MODEL_NS2D) {
    const bool is3 = (c.model == HPB_MODEL_NS3D);
    double p0 = c.p_ref, rho0 = c.rho_ref, RT = p0 / rho0, gamma = c.gamma, R = c.R;
    double Cp = gamma * R / (gamma - 1.0), T0 = p0 / (rho0 * R);
    double gx = c.gravity[0], gy = c.gravity[1], gz = c.gravity[2], Nbv = c.N_bv;
    if (c.HB == 3 && is3 && (gx != 0 || gy != 0))
      return hpb_fail(HPB_ERR_INVALID, "HB = 3 is implemented only for gravity force along the z-coordinate");
    if (c.HB == 3 && !is3 && gx != 0)
      return hpb_fail(HPB_ERR_INVALID, "HB = 3 is implemented only for gravity force along the y-coordinate");
    const double gv = is3 ? gz : gy;          // vertical component
    const double* X = h->x_h.data();
    for (int k = 0; k < G.P[2]; k++) for (int j = 0; j < G.P[1]; j++) for (int i = 0; i < G.P[0]; i++) {
      size_t p = i + (size_t)G.P[0] * (j + (size_t)G.P[1] * k);
      double xc = X[G.xoff[0] + i], yc = X[G.xoff[1] + j], zc = is3 ? X[G.xoff[2] + k] : 0.0;
      const double phi = is3 ? (gx*xc+gy*yc+gz*zc) : (gx*xc+gy*yc);
      const double vc = is3 ? zc : yc;
      double f, gg;
      if (c.HB == 1) {
        f  = exp( phi/RT);
        gg = exp(-phi/RT);
      } else if (c.HB == 2) {
        f  = raiseto((1.0-phi/(Cp*T0)), (-1.0 /(gamma-1.0)));
        gg = raiseto((1.0-phi/(Cp*T0)), (gamma/(gamma-1.0)));
      } else if (c.HB == 3) {
        double Pexner = 1 + ((gv*gv)/(Cp*T0*Nbv*Nbv)) * (exp(-(Nbv*Nbv/gv)*vc)-1.0);
        f  = raiseto(Pexner, (-1.0 /(gamma-1.0))) * exp(Nbv*Nbv*vc/gv);
        gg = raiseto(Pexner, (gamma/(gamma-1.0)));
      } else { f = gg = 1.0; }
      h->gravf_h[p] = f; h->gravg_h[p] = gg;
    }
    for (int d = 0; d < G.ndims; d++) for (int side = 0; side < 2; side++) {
      if (side == 0 && h->ip[d] != 0) continue;
      if (side == 1 && h->ip[d] != c.iproc[d] - 1) continue;
      int b[3] = { G.N[0], G.N[1], G.N[2] };
      b[d] = g;
      for (int k = 0; k < b[2]; k++) for (int j = 0; j < b[1]; j++) for (int i = 0; i < b[0]; i++) {
        int ib[3] = { i, j, k }, i1[3] = { i, j, k }, i2[3] = { i, j, k };
        if (side == 0) { i1[d] = ib[d] - g;        i2[d] = g - 1 - ib[d]; }
        else           { i1[d] = ib[d] + G.N[d];   i2[d] = G.N[d] - 1 - ib[d]; }
        const int g2 = (G.ndims > 2) ? g : 0;     // no ghost layers along an unused third dimension
        size_t p1 = (i1[0]+g) + (size_t)G.P[0] * ((i1[1]+g) + (size_t)G.P[1] * (i1[2]+g2));
        size_t p2 = (i2[0]+g) + (size_t)G.P[0] * ((i2[1]+g) + (size_t)G.P[1] * (i2[2]+g2));
        h->gravf_h[p1] = h->gravf_h[p2]; h->gravg_h[p1] = h->gravg_h[p2];
      }
    }
  }

  // ---- LinearADR spatially varying advection (LinearADRAdvectionField.c:25-193): this rank's block with ghosts. Interior
  // from the global field; internal faces from the neighbour's interior (= the global field); physical faces: periodic copy
  // (one rank along the dimension) or mirror extrapolation; edges and corners are never filled (zero).
  h->advf_h.clear();
  if (c.advection_field) {
    if (c.model != HPB_MODEL_LINEAR_ADR || c.nvars != 1)
      return hpb_fail(HPB_ERR_INVALID, "advection_field: linear-advection-diffusion-reaction with nvars = 1 only");
    const int NG[3] = { c.dim_global[0], nd > 1 ? c.dim_global[1] : 1, nd > 2 ? c.dim_global[2] : 1 };
    h->advf_h.assign((size_t)nd * G.npg, 0.0);
    bool dper[3] = { false, false, false };
    for (int n = 0; n < c.nzones; n++) if (c.zones[n].type == HPB_BC_PERIODIC) dper[c.zones[n].dim] = true;
    auto gidx = [&](int i0, int i1, int i2) { return (size_t)i0 + (size_t)NG[0] * ((size_t)i1 + (size_t)NG[1] * (size_t)i2); };
    auto lidx = [&](int i0, int i1, int i2) {        // local indices, may be ghosts
      size_t p = (size_t)(i0 + g);
      if (nd > 1) p += (size_t)G.P[0] * (size_t)(i1 + g);
      if (nd > 2) p += (size_t)G.P[0] * (size_t)G.P[1] * (size_t)(i2 + g);
      return p; };
    // interior + ghost layers along ONE dimension at a time (the transverse indices stay interior)
    for (int dd = -1; dd < nd; dd++) {
      int lo[3] = { 0, 0, 0 }, hi[3] = { G.N[0], G.N[1], G.N[2] };
      for (int pass = 0; pass < (dd < 0 ? 1 : 2); pass++) {
        if (dd >= 0) { lo[dd] = pass ? G.N[dd] : -g; hi[dd] = pass ? G.N[dd] + g : 0; }
        for (int k = lo[2]; k < hi[2]; k++) for (int j = lo[1]; j < hi[1]; j++) for (int i = lo[0]; i < hi[0]; i++) {
          int li[3] = { i, j, k }, gi[3];
          bool ok = true;
          for (int d = 0; d < 3; d++) gi[d] = (d < nd ? h->is_global[d] + li[d] : 0);
          if (dd >= 0) {
            const int d = dd, n = G.N[d];
            const bool low = (li[d] < 0);
            const bool internal = low ? (h->ip[d] > 0) : (h->ip[d] < c.iproc[d] - 1);
            // (sic) a periodic dimension split among ranks: the exchange wraps the end ranks' outer ghosts, and the
            // else-branch of LinearADRAdvectionField.c:142-189 then overwrites them with the mirror image
            if (internal) { /* neighbour's interior: the global index as it is */ }
            else if (dper[d] && c.iproc[d] == 1) gi[d] = h->is_global[d] + (low ? li[d] + n : li[d] - n);   // periodic copy
            else gi[d] = h->is_global[d] + (low ? (-1 - li[d]) : (2 * n - 1 - li[d]));                      // mirror
            ok = (gi[d] >= 0 && gi[d] < NG[d]);
          }
          if (!ok) continue;
          const size_t p = lidx(li[0], li[1], li[2]), q = gidx(gi[0], gi[1], gi[2]);
          for (int d = 0; d < nd; d++) h->advf_h[(size_t)d * G.npg + p] = c.advection_field[q * nd + d];
        }
        if (dd >= 0) { lo[dd] = 0; hi[dd] = G.N[dd]; }
      }
    }
  }

  if (c.model == HPB_MODEL_EULER1D) {
    // Euler1DGravityField.c:25-95: ONE field S, used as both grav_f and grav_g here; type 0 exp(-g x) with mirror copies into
    // the ghosts of physical faces, type 1 exp(sin(2 pi x) / (2 pi)) without
    const double* X = h->x_h.data();
    const int n = G.N[0];
    for (int i = 0; i < G.P[0]; i++) {
      double S = 1.0;
      if (c.gravity_type == 0) S = exp(-c.gravity[0] * X[i]);
      else { const double pi = 4.0 * atan(1.0); const double phi = -sin(2 * pi * X[i]) / (2 * pi); S = exp(-phi); }
      h->gravf_h[i] = S;
    }
    if (c.gravity_type != 1) {
      if (h->ip[0] == 0)              for (int k = 0; k < g; k++) h->gravf_h[k] = h->gravf_h[g + (g - 1 - k)];
      if (h->ip[0] == c.iproc[0] - 1) for (int k = 0; k < g; k++) h->gravf_h[g + n + k] = h->gravf_h[g + (n - 1 - k)];
    }
    for (int i = 0; i < G.P[0]; i++) h->gravg_h[i] = h->gravf_h[i];
  }

  // ---- RK tableau
  RKTableau& T = h->rk;
  memset(&T, 0, sizeof(T));
  if (c.rk_type >= HPB_GLMGEE_23) {             // TimeGLMGEEInitialize.c:41-495 (tables: glmgee_tables.h)
    const glmgee_table& M = GLMGEE_TABLES[c.rk_type - HPB_GLMGEE_23];
    const int m = c.glm_ee_mode, s = M.s;
    T.ns = s; T.glm = 1; T.mode = m; T.gamma = M.gamma;
    for (int k = 0; k < s * s; k++) T.A[k] = M.A[m][k];
    for (int k = 0; k < s; k++) { T.b[k] = M.B[m][k]; T.b1[k] = M.B[m][s + k]; T.c[k] = M.c[m][k]; }
    for (int k = 0; k < 2 * s; k++) T.C[k] = M.C[m][k];
    for (int k = 0; k < 4; k++) T.D[k] = M.D[m][k];
  } else if (c.rk_type == HPB_RK_44) {
    T.ns = 4;
    T.A[4] = 0.5; T.A[9] = 0.5; T.A[14] = 1.0;
    T.c[0] = 0.0; T.c[1] = T.c[2] = 0.5; T.c[3] = 1.0;
    T.b[0] = T.b[3] = 1.0/6.0; T.b[1] = T.b[2] = 1.0/3.0;
  } else if (c.rk_type == HPB_RK_SSPRK3) {
    T.ns = 3;
    T.A[3] = 1.0; T.A[6] = 0.25; T.A[7] = 0.25;
    T.c[1] = 1.0; T.c[2] = 0.5; T.c[0] = 0.0;
    T.b[0] = T.b[1] = 1.0/6.0; T.b[2] = 2.0/3.0;
  } else if (c.rk_type == HPB_RK_1FE) {         // TimeExplicitRKInitialize.c:27-36
    T.ns = 1;
    T.b[0] = 1.0;
  } else if (c.rk_type == HPB_RK_22) {          // :37-48
    T.ns = 2;
    T.A[2] = 1.0; T.c[1] = 1.0;
    T.b[0] = T.b[1] = 0.5;
  } else {                                      // HPB_RK_33, :49-60
    T.ns = 3;
    T.A[3] = 2.0/3.0; T.A[6] = 2.0/3.0-1.0/4.0; T.A[7] = 1.0/4.0;
    T.c[1] = 2.0/3.0; T.c[2] = 2.0/3.0;
    T.b[0] = 1.0/4.0; T.b[1] = -1.0/4.0; T.b[2] = 1.0;
  }
  return HPB_OK;
}
