// TEST-ONLY host harness (never loaded by the product): compiles the per-point mathematics of
// turtlefsi_b200/csrc/tfsi_math.cuh with g++ and drives it cell by cell exactly as the CUDA
// kernels do (phase 0 base items -> tasks -> block updates), returning DENSE element matrices /
// vectors so that the formulas can be checked against oracle/forms.py without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>
#define TFSI_NO_CUDA 1
#include "../../turtlefsi_b200/csrc/tfsi_math.cuh"

using namespace tfsi;

namespace {

template <int G>
void jac_cell(const QuadTab<MAXQ_J>& tab, const RegionConst& rc, const double* Xp, const double* u,
              const double* u1, int ndof, double* Ae)
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    double X[G + 1][G];
    for (int v = 0; v <= G; ++v) for (int c = 0; c < G; ++c) X[v][c] = Xp[v * G + c];
    double K[G][G];
    double adet = cell_geometry<G>(X, K);
    const int nq = tab.nq;
    auto cf = [&](int k) { return u[k]; };
    auto cf1 = [&](int k) { return u1[k]; };
    auto addblock = [&](int r0, int c0, int nr, int ncl, const double* Z, int ldz) {
        for (int i = 0; i < nr; ++i) for (int j = 0; j < ncl; ++j) Ae[(r0 + i) * ndof + c0 + j] += Z[i * ldz + j];
    };
    if (rc.kind == 0) {
        typedef FluidBase<G> FB;
        std::vector<double> base((size_t)FB::COUNT * nq);
        for (int q = 0; q < nq; ++q) {
            auto out = [&](int item, double v) { base[(size_t)item * nq + q] = v; };
            fluid_base<G>(tab, q, K, adet, rc, cf, cf1, out);
        }
        constexpr int NJ = HalfCols<G, 0>::NJ;
        for (int fld = 0; fld < 2; ++fld)
            for (int a = 0; a < G; ++a)
                for (int b = 0; b < G; ++b)
                    for (int half = 0; half < 2; ++half) {
                        double Z[N2][NJ];
                        memset(Z, 0, sizeof Z);
                        for (int q = 0; q < nq; ++q) {
                            auto bv = [&](int item) { return base[(size_t)item * nq + q]; };
                            double C[G + 1][G + 1];
                            if (fld == 0) fluid_C_psi_dd<G>(bv, rc, a, b, C); else fluid_C_psi_dv<G>(bv, rc, a, b, C);
                            if (half == 0) block_p2p2<G, 0, 7>(tab, q, C, Z); else block_p2p2<G, 1, 7>(tab, q, C, Z);
                        }
                        for (int i = 0; i < N2; ++i) for (int jj = 0; jj < NJ; ++jj) {
                            int j = half == 0 ? HalfCols<G, 0>::col(jj) : HalfCols<G, 1>::col(jj);
                            Ae[(G * N2 + a * N2 + i) * ndof + fld * G * N2 + b * N2 + j] += Z[i][jj];
                        }
                    }
        for (int a = 0; a < G; ++a) {
            double Z[N2][N1];
            memset(Z, 0, sizeof Z);
            for (int q = 0; q < nq; ++q) {
                auto bv = [&](int item) { return base[(size_t)item * nq + q]; };
                double c[G];
                fluid_c_psi_dp<G>(bv, a, c);
                block_gradp2_p1<G>(tab, q, c, Z);
            }
            addblock(G * N2 + a * N2, 2 * G * N2, N2, N1, &Z[0][0], N1);
        }
        for (int fld = 0; fld < 2; ++fld)
            for (int b = 0; b < G; ++b) {
                double Z[N1][N2];
                memset(Z, 0, sizeof Z);
                for (int q = 0; q < nq; ++q) {
                    auto bv = [&](int item) { return base[(size_t)item * nq + q]; };
                    double c[G];
                    if (fld == 0) fluid_c_gamma_dd<G>(bv, b, c); else fluid_c_gamma_dv<G>(bv, b, c);
                    block_p1_gradp2<G>(tab, q, c, Z);
                }
                addblock(2 * G * N2, fld * G * N2 + b * N2, N1, N2, &Z[0][0], N2);
            }
    } else {
        const bool energy = rc.mat != MAT_SVK;
        const int count = energy ? SolidBaseGen<G>::COUNT : SolidBase<G>::COUNT;
        std::vector<double> base((size_t)count * nq);
        for (int q = 0; q < nq; ++q) {
            auto out = [&](int item, double v) { base[(size_t)item * nq + q] = v; };
            if (energy) solid_base_energy<G>(tab, q, K, adet, rc, cf, out);
            else solid_base<G>(tab, q, K, adet, rc, cf, out);
        }
        constexpr int NJ = HalfCols<G, 0>::NJ;
        for (int a = 0; a < G; ++a)
            for (int b = 0; b < G; ++b)
                for (int half = 0; half < 2; ++half) {
                    double Z[N2][NJ];
                    memset(Z, 0, sizeof Z);
                    for (int q = 0; q < nq; ++q) {
                        auto bv = [&](int item) { return base[(size_t)item * nq + q]; };
                        double C[G + 1][G + 1];
                        if (energy) solid_C_psi_dd_energy<G>(bv, a, b, C);
                        else solid_C_psi_dd<G>(bv, rc, a, b, C);
                        if (half == 0) block_p2p2<G, 0, 4>(tab, q, C, Z); else block_p2p2<G, 1, 4>(tab, q, C, Z);
                    }
                    for (int i = 0; i < N2; ++i) for (int jj = 0; jj < NJ; ++jj) {
                        int j = half == 0 ? HalfCols<G, 0>::col(jj) : HalfCols<G, 1>::col(jj);
                        Ae[(G * N2 + a * N2 + i) * ndof + b * N2 + j] += Z[i][jj];
                    }
                }
    }
}

template <int G>
void res_cell(const QuadTab<MAXQ_R>& tab, const RegionConst& rc, const double* Xp, const double* u,
              const double* u1, int ndof, double* Fe)
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    double X[G + 1][G];
    for (int v = 0; v <= G; ++v) for (int c = 0; c < G; ++c) X[v][c] = Xp[v * G + c];
    double K[G][G];
    double adet = cell_geometry<G>(X, K);
    double alfa_cell = rc.alfa;
    if ((rc.ext == 1 && rc.ext_sub == 2) || rc.ext == 3) alfa_cell = 1.0 / (adet / (G == 2 ? 2.0 : 6.0));
    auto cf = [&](int k) { return u[k]; };
    auto cf1 = [&](int k) { return u1[k]; };
    const bool bih = ndof > 2 * G * N2 + N1;
    for (int q = 0; q < tab.nq; ++q) {
        PointFlux<G> f;
        if (rc.ext == 1 && rc.ext_sub == 3) residual_point<G, LIFT_VOLUME_CHANGE>(tab, q, K, adet, rc, alfa_cell, cf, cf1, f);
        else if (rc.ext == 3) residual_point<G, LIFT_ELASTIC>(tab, q, K, adet, rc, alfa_cell, cf, cf1, f);
        else if (rc.kind == 1 && (rc.mat != MAT_SVK || rc.visco)) residual_point<G, LIFT_SOLID_GENERAL>(tab, q, K, adet, rc, alfa_cell, cf, cf1, f);
        else residual_point<G>(tab, q, K, adet, rc, alfa_cell, cf, cf1, f);
        for (int c = 0; c < G; ++c)
            for (int i = 0; i < N2; ++i) {
                double zp = tab.phi2[q][i] * f.psi_val[c], zd = tab.phi2[q][i] * f.phi_val[c], zb = tab.phi2[q][i] * f.beta_val[c];
                for (int s = 0; s < G; ++s) {
                    zp += tab.dphi2[q][i][s] * f.psi_grad[c][s];
                    zd += tab.dphi2[q][i][s] * f.phi_grad[c][s];
                    zb += tab.dphi2[q][i][s] * f.beta_grad[c][s];
                }
                Fe[c * N2 + i] += zd;
                Fe[G * N2 + c * N2 + i] += zp;
                if (bih) Fe[2 * G * N2 + N1 + c * N2 + i] += zb;
            }
        for (int m = 0; m < N1; ++m) Fe[2 * G * N2 + m] += tab.phi1[q][m] * f.gamma;
    }
}

// phi x delta d block of the frozen non-linear lifting Jacobian (lin_jacobian_kernel, gen_kind 1 | 2) of one cell
template <int G>
void lift_cell(const QuadTab<MAXQ_J>& tab, int gen_kind, const double* Xp, const double* u, int ndof, double* Ae)
{
    constexpr int N2 = Dim<G>::N2;
    constexpr int NJ = HalfCols<G, 0>::NJ;
    double X[G + 1][G];
    for (int v = 0; v <= G; ++v) for (int c = 0; c < G; ++c) X[v][c] = Xp[v * G + c];
    double K[G][G];
    const double adet = cell_geometry<G>(X, K);
    const double Ey = 1.0 / (adet / (G == 2 ? 2.0 : 6.0)), nu = 0.25;
    const double a_lam = nu * Ey / ((1.0 + nu) * (1.0 - 2.0 * nu)), a_mu = Ey / (2.0 * (1.0 + nu));
    for (int ca = 0; ca < G; ++ca)
        for (int cb = 0; cb < G; ++cb)
            for (int half = 0; half < 2; ++half) {
                double Z[N2][NJ];
                memset(Z, 0, sizeof Z);
                for (int q = 0; q < tab.nq; ++q) {
                    const double jw = tab.w[q] * adet;
                    double Gp[G][G];
                    for (int aa = 0; aa < G; ++aa) {
                        double Gd[G] = {};
                        for (int j = 0; j < N2; ++j) for (int t = 0; t < G; ++t) Gd[t] += u[aa * N2 + j] * tab.dphi2[q][j][t];
                        for (int b = 0; b < G; ++b) { double sacc = 0; for (int t = 0; t < G; ++t) sacc += Gd[t] * K[t][b]; Gp[aa][b] = sacc; }
                    }
                    double D[G][G][G][G];
                    lift_flux_derivative<G>(gen_kind, Gp, a_mu, a_lam, D);
                    double C[G + 1][G + 1];
                    C[0][0] = 0.0;
                    for (int s = 0; s < G; ++s) {
                        C[0][1 + s] = 0; C[1 + s][0] = 0;
                        for (int t = 0; t < G; ++t) {
                            double z = 0;
                            for (int s2 = 0; s2 < G; ++s2) for (int t2 = 0; t2 < G; ++t2) z += K[s][s2] * D[ca][s2][cb][t2] * K[t][t2];
                            C[1 + s][1 + t] = jw * z;
                        }
                    }
                    if (half == 0) block_p2p2<G, 0, 5>(tab, q, C, Z); else block_p2p2<G, 1, 5>(tab, q, C, Z);
                }
                for (int i = 0; i < N2; ++i) for (int jj = 0; jj < NJ; ++jj) {
                    int j = half == 0 ? HalfCols<G, 0>::col(jj) : HalfCols<G, 1>::col(jj);
                    Ae[(ca * N2 + i) * ndof + cb * N2 + j] += Z[i][jj];
                }
            }
}

// one viscoelastic block (gen_kind 3..6 of lin_jacobian_kernel) of one cell: rows psi, columns delta d | delta v
template <int G>
void visco_cell(const QuadTab<MAXQ_J>& tab, const RegionConst& rc, int gen_kind, const double* Xp, const double* u,
                const double* u1, int ndof, double* Ae)
{
    constexpr int N2 = Dim<G>::N2;
    constexpr int NJ = HalfCols<G, 0>::NJ;
    double X[G + 1][G];
    for (int v = 0; v <= G; ++v) for (int c = 0; c < G; ++c) X[v][c] = Xp[v * G + c];
    double K[G][G];
    const double adet = cell_geometry<G>(X, K);
    const int trial_off = (gen_kind == GEN_VISCO_F1_DV) ? G * N2 : 0;
    for (int ca = 0; ca < G; ++ca)
        for (int cb = 0; cb < G; ++cb)
            for (int half = 0; half < 2; ++half) {
                double Z[N2][NJ];
                memset(Z, 0, sizeof Z);
                for (int q = 0; q < tab.nq; ++q) {
                    const double jw = tab.w[q] * adet;
                    auto phys_grad = [&](const double* vec, int off, double (&out)[G][G]) {
                        for (int aa = 0; aa < G; ++aa) {
                            double Gr[G] = {};
                            for (int j = 0; j < N2; ++j) for (int t = 0; t < G; ++t) Gr[t] += vec[off + aa * N2 + j] * tab.dphi2[q][j][t];
                            for (int b = 0; b < G; ++b) { double sacc = 0; for (int t = 0; t < G; ++t) sacc += Gr[t] * K[t][b]; out[aa][b] = sacc; }
                        }
                    };
                    double Fm[G][G], Gw[G][G], D[G][G][G][G];
                    phys_grad(gen_kind == GEN_VISCO_F2_LIN ? u1 : u, 0, Fm);
                    if (gen_kind == GEN_VISCO_F2_DD) {
                        double G1[G][G];
                        phys_grad(u1, 0, G1);
                        for (int aa = 0; aa < G; ++aa) for (int b = 0; b < G; ++b) Gw[aa][b] = Fm[aa][b] - G1[aa][b];
                    } else if (gen_kind == GEN_VISCO_F2_LIN) {
                        for (int aa = 0; aa < G; ++aa) for (int b = 0; b < G; ++b) Gw[aa][b] = 0.0;
                    } else phys_grad(u, G * N2, Gw);
                    for (int aa = 0; aa < G; ++aa) Fm[aa][aa] += 1.0;
                    visco_flux_derivative<G>(gen_kind, rc, Fm, Gw, D);
                    double C[G + 1][G + 1];
                    C[0][0] = 0.0;
                    for (int s = 0; s < G; ++s) {
                        C[0][1 + s] = 0; C[1 + s][0] = 0;
                        for (int t = 0; t < G; ++t) {
                            double z = 0;
                            for (int s2 = 0; s2 < G; ++s2) for (int t2 = 0; t2 < G; ++t2) z += K[s][s2] * D[ca][s2][cb][t2] * K[t][t2];
                            C[1 + s][1 + t] = jw * z;
                        }
                    }
                    if (half == 0) block_p2p2<G, 0, 5>(tab, q, C, Z); else block_p2p2<G, 1, 5>(tab, q, C, Z);
                }
                for (int i = 0; i < N2; ++i) for (int jj = 0; jj < NJ; ++jj) {
                    int j = half == 0 ? HalfCols<G, 0>::col(jj) : HalfCols<G, 1>::col(jj);
                    Ae[(G * N2 + ca * N2 + i) * ndof + trial_off + cb * N2 + j] += Z[i][jj];
                }
            }
}

}  // namespace

extern "C" {

// solid region with any material model / viscoelasticity. part 0: element matrices of J_nonlinear (jac_solid_kernel +
// the viscoelastic blocks), 1: the state-dependent block of J_linear ("Form2", frozen in A_pre), 2: element vectors of F
int emul_solid_general(int g, int nqJ, const double* ptsJ, const double* wtsJ, int nqR, const double* ptsR, const double* wtsR,
                       double rho, double mu, double lambda, double dt, double theta, int material, double C01, double C10,
                       double C11, double Jm, int visco, double mu_visc, double lam_visc, int part, int64_t nc,
                       const double* X, const double* u, const double* u1, int ndof, double* out)
{
    RegionConst rc = make_region_const(1, rho, mu, lambda, dt, theta, 1, 1, 1, 0, 0.0, 0.0);
    set_solid_model(rc, material, C01, C10, C11, Jm, visco, mu_visc, lam_visc);
    if (part == 2) {
        QuadTab<MAXQ_R> tab;
        if (tabulate<MAXQ_R>(g, nqR, ptsR, wtsR, tab)) return -1;
        for (int64_t c = 0; c < nc; ++c) {
            if (g == 3) res_cell<3>(tab, rc, X + c * 12, u + c * ndof, u1 + c * ndof, ndof, out + c * ndof);
            else res_cell<2>(tab, rc, X + c * 6, u + c * ndof, u1 + c * ndof, ndof, out + c * ndof);
        }
        return 0;
    }
    QuadTab<MAXQ_J> tab;
    if (tabulate<MAXQ_J>(g, nqJ, ptsJ, wtsJ, tab)) return -1;
    for (int64_t c = 0; c < nc; ++c) {
        const double* Xc = X + c * (g + 1) * g;
        const double *uc = u + c * ndof, *u1c = u1 + c * ndof;
        double* Ae = out + c * ndof * ndof;
        auto visco_blk = [&](int kind) {
            if (g == 3) visco_cell<3>(tab, rc, kind, Xc, uc, u1c, ndof, Ae); else visco_cell<2>(tab, rc, kind, Xc, uc, u1c, ndof, Ae);
        };
        if (part == 0) {
            if (g == 3) jac_cell<3>(tab, rc, Xc, uc, u1c, ndof, Ae); else jac_cell<2>(tab, rc, Xc, uc, u1c, ndof, Ae);
            if (visco == 1) { visco_blk(GEN_VISCO_F1_DD); visco_blk(GEN_VISCO_F1_DV); }
            else if (visco == 2) visco_blk(GEN_VISCO_F2_DD);
        } else if (visco == 2) visco_blk(GEN_VISCO_F2_LIN);
    }
    return 0;
}

// Ae[nc][ndof][ndof] += phi x delta d block of the non-linear lifting (gen_kind 1 volume_change, 2 elastic) at state u
int emul_lifting_jacobian(int g, int nq, const double* pts, const double* wts, int gen_kind, int64_t nc, const double* X,
                          const double* u, int ndof, double* Ae)
{
    QuadTab<MAXQ_J> tab;
    if (tabulate<MAXQ_J>(g, nq, pts, wts, tab)) return -1;
    for (int64_t c = 0; c < nc; ++c) {
        if (g == 3) lift_cell<3>(tab, gen_kind, X + c * 12, u + c * ndof, ndof, Ae + c * ndof * ndof);
        else lift_cell<2>(tab, gen_kind, X + c * 6, u + c * ndof, ndof, Ae + c * ndof * ndof);
    }
    return 0;
}

// Ae[nc][ndof][ndof] += element matrices of J_nonlinear ; X[nc][(g+1)*g], u/u1[nc][ndof]
int emul_jacobian(int g, int nq, const double* pts, const double* wts, int kind, double rho, double mu,
                  double lambda, double dt, double theta, int64_t nc, const double* X, const double* u,
                  const double* u1, int ndof, double* Ae)
{
    QuadTab<MAXQ_J> tab;
    if (tabulate<MAXQ_J>(g, nq, pts, wts, tab)) return -1;
    RegionConst rc = make_region_const(kind, rho, mu, lambda, dt, theta, 1, 1, 1, 0, 0.0, 0.0);
    for (int64_t c = 0; c < nc; ++c) {
        if (g == 3) jac_cell<3>(tab, rc, X + c * 12, u + c * ndof, u1 + c * ndof, ndof, Ae + c * ndof * ndof);
        else jac_cell<2>(tab, rc, X + c * 6, u + c * ndof, u1 + c * ndof, ndof, Ae + c * ndof * ndof);
    }
    return 0;
}

// Fe[nc][ndof] += element vectors of F = F_lin + F_nonlin
int emul_residual(int g, int nq, const double* pts, const double* wts, int kind, double rho, double mu,
                  double lambda, double dt, double theta, int fluid_on, int solid_on, int ext, int ext_sub,
                  double hmin, double gravity, int64_t nc, const double* X, const double* u, const double* u1,
                  int ndof, double* Fe)
{
    QuadTab<MAXQ_R> tab;
    if (tabulate<MAXQ_R>(g, nq, pts, wts, tab)) return -1;
    RegionConst rc = make_region_const(kind, rho, mu, lambda, dt, theta, fluid_on, solid_on, ext, ext_sub, hmin, gravity);
    for (int64_t c = 0; c < nc; ++c) {
        if (g == 3) res_cell<3>(tab, rc, X + c * 12, u + c * ndof, u1 + c * ndof, ndof, Fe + c * ndof);
        else res_cell<2>(tab, rc, X + c * 6, u + c * ndof, u1 + c * ndof, ndof, Fe + c * ndof);
    }
    return 0;
}

}  // extern "C"
