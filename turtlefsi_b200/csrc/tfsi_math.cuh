// Per-quadrature-point mathematics of the monolithic FSI forms, shared by the sm_100a kernels
// (tfsi_kernels.cuh). Everything here is a pure inline function on small fixed-size arrays.
//
// What is computed (reference files):
//   fluid residual / Jacobian   turtleFSI/modules/fluid.py:32-56  (N1..N6, L1, L2, L4)
//   solid residual / Jacobian   turtleFSI/modules/solid.py:29-68  (S1..S5, SVK via common.py:84-109)
//   mesh lifting                turtleFSI/modules/laplace.py:33-46, biharmonic.py:31-36
//   split into F_lin / F_nonlin and their Gateaux derivatives  modules/newtonsolver.py:18-24
//
// Formulation (B200-first, not FFC's): every element-matrix block (test component a, trial
// component b) is   Z_ij = sum_q  T_is(q) * C'_st(q) * U_jt(q)
// with T/U the REFERENCE-element tabulation (value slot 0, reference-gradient slots 1..G),
// identical for all cells and therefore read from constant memory as warp-uniform operands,
// and C' a small per-point tensor in which the geometry and the deformation are folded:
// B = K F^-1 maps reference gradients to gradients on the deformed cell, g_i = B^T ghat_i.
// Sum factorisation: W_sj = C'_st U_jt, then Z_ij += T_is W_sj.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define TFSI_HD __host__ __device__ __forceinline__
#define TFSI_HD_NOINLINE __host__ __device__ __noinline__
#else
#define TFSI_HD inline
#define TFSI_HD_NOINLINE inline
#endif

namespace tfsi {

constexpr int MAXQ_J = 32;   // Jacobian rule: <= 32 points (degree <= 6 in 3-D: 24; rules beyond the shared-memory table
                             // of one launch are integrated in chunks)
constexpr int MAXQ_R = 64;   // residual rule: 24 (3-D degree 6), 36 (2-D degree 10), 64 (3-D degree 7)

template <int MAXQ>
struct QuadTab {
    int nq;
    int pad_;
    double w[MAXQ];
    double phi2[MAXQ][10];
    double dphi2[MAXQ][10][3];
    double phi1[MAXQ][4];
    double d4[MAXQ][4];    // 4 lambda_v - 1   (P2 vertex-function gradient factor)
    double l4[MAXQ][4];    // 4 lambda_v       (P2 edge-function gradient factors)
};

template <int G> struct Dim {
    static constexpr int N1 = G + 1;
    static constexpr int N2 = (G == 2) ? 6 : 10;
    static constexpr int NS = G * (G + 1) / 2;     // packed symmetric
};

TFSI_HD int symi(int a, int b, int G) {            // packed index of a symmetric GxG matrix
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return lo * G - lo * (lo - 1) / 2 + (hi - lo);
}

// ---------------------------------------------------------------------------------------------
// small dense helpers
template <int G> TFSI_HD double det(const double (&F)[G][G]);
template <> TFSI_HD double det<2>(const double (&F)[2][2]) { return F[0][0] * F[1][1] - F[0][1] * F[1][0]; }
template <> TFSI_HD double det<3>(const double (&F)[3][3]) {
    return F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1])
         - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0])
         + F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
}
template <int G> TFSI_HD double inv(const double (&F)[G][G], double (&Fi)[G][G]);
template <> TFSI_HD double inv<2>(const double (&F)[2][2], double (&Fi)[2][2]) {
    double d = det<2>(F), r = 1.0 / d;
    Fi[0][0] = F[1][1] * r; Fi[0][1] = -F[0][1] * r;
    Fi[1][0] = -F[1][0] * r; Fi[1][1] = F[0][0] * r;
    return d;
}
template <> TFSI_HD double inv<3>(const double (&F)[3][3], double (&Fi)[3][3]) {
    double c00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    double c01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    double c02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    double d = F[0][0] * c00 + F[0][1] * c01 + F[0][2] * c02, r = 1.0 / d;
    Fi[0][0] = c00 * r;
    Fi[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * r;
    Fi[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * r;
    Fi[1][0] = c01 * r;
    Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * r;
    Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * r;
    Fi[2][0] = c02 * r;
    Fi[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * r;
    Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * r;
    return d;
}

// cell geometry: X[G+1][G] vertex coordinates -> K = J_geo^-1 (K[t][b] = d xi_t / d x_b), |det|
template <int G> TFSI_HD double cell_geometry(const double (&X)[G + 1][G], double (&K)[G][G]) {
    double Jg[G][G];
#pragma unroll
    for (int b = 0; b < G; ++b)
#pragma unroll
        for (int t = 0; t < G; ++t) Jg[b][t] = X[t + 1][b] - X[0][b];
    double d = inv<G>(Jg, K);
    return d < 0 ? -d : d;
}

// ---------------------------------------------------------------------------------------------
// physical constants of one region, pre-combined on the host
struct RegionConst {
    double c1;      // rho_f theta0 / k      (fluid.py:32)
    double c1lin;   // rho_f theta1 / k      (fluid.py:33)
    double rk;      // rho_f / k             (fluid.py:55)
    double th0rho;  // theta0 rho_f          (fluid.py:36)
    double th1rho;  // theta1 rho_f          (fluid.py:38)
    double mu;      // mu_f | mu_s
    double lam;     // lambda_s
    double th0, th1;
    double rho_s_k;        // rho_s / k                   (solid.py:41)
    double delta_rho_k;    // delta rho_s / k             (solid.py:42)
    double delta_rho;      // delta rho_s                 (solid.py:43)
    double alfa;           // laplace: constant / small_constant value; biharmonic: alpha_u
    double grav_rho;       // gravity * rho_s (solid.py:65-68), 0 if none
    int kind;              // 0 fluid, 1 solid
    int fluid_on, solid_on;
    int ext;               // 0 none, 1 laplace, 2 biharmonic, 3 elastic
    int ext_sub;           // laplace: 0 constant, 1 small_constant, 2 volume, 3 volume_change
    int mat;               // solid material model: 0 StVenantKirchoff (closed form), else MAT_* (energy based, common.py:110-125)
    double mC01, mC10, mC11, mJm;   // MooneyRivlin C01, C10, C11; Gent Jm
    int visco;             // viscoelasticity (solid.py:46-59): 0 none, 1 "Form1" (rate from v), 2 "Form2" (rate from (d - d1)/k)
    int pad2_;
    double mu_visc, lam_visc;   // Svisc_D = mu_visc tr(D) I + 2 lambda_visc D (common.py:129-139)
    double inv_k;               // 1 / k
};

// material models of turtleFSI/modules/common.py:100-127 (the `material_model` strings, :104-121)
enum { MAT_SVK = 0, MAT_SVK_ENERGY = 1, MAT_NEO_HOOKEAN = 2, MAT_MOONEY_RIVLIN = 3, MAT_GENT = 4 };

// ---------------------------------------------------------------------------------------------
// FLUID, non-linear Jacobian: per-point base items (stored lane-contiguous in shared memory).
// Items are PRE-SCALED in phase 0 so that every C' entry costs <= 4 FMAs and few loads:
//   k4 = theta0 mu J w ;  sigma = mu (H + H^T)
template <int G> struct FluidBase {
    static constexpr int JW = 0;                    // J w |detJgeo|
    static constexpr int B = 1;                     // B[t][c] = (K F^-1)[t][c]          G*G
    static constexpr int H = B + G * G;             // H = grad v F^-1                    G*G
    static constexpr int MJ = H + G * G;            // k4 * B B^T                         packed sym
    static constexpr int BUJ = MJ + Dim<G>::NS;     // Jw * B u,  u = th0 rho v - rho/k (d - d1)
    static constexpr int XJ = BUJ + G;              // Jw * (c1 (v - v1) + H u)
    static constexpr int PJ = XJ + G;               // Jw * (theta0 B sigma - p B)        G*G
    static constexpr int RJ = PJ + G * G;           // k4 * B H                           G*G
    static constexpr int MASSJ = RJ + G * G;        // pair mode: the FROZEN mass factor rho_f theta1 / k J(d1_setup) w |detJgeo|
                                                    // of fluid.py:33 (part of A_pre; 0 when A_pre is already in the matrix)
    static constexpr int COUNT = MASSJ + 1;         // 50 (3-D), 25 (2-D): same padded stride as 49 / 24
};

// coefficient access: cf(k) returns local dof k of the cell at time n, cf1(k) at n-1
// (local layout [d comps | v comps | p | w comps]); out(item, value) stores a base item.
template <int G, class Tab, class CF, class CF1, class OUT>
TFSI_HD void fluid_base(const Tab& tab, int q, const double (&K)[G][G], double adet,
                        const RegionConst& rc, CF cf, CF1 cf1, OUT out, double massj = 0.0)
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    typedef FluidBase<G> I;
    double Gd[G][G], Gv[G][G], v[G], dd[G], dv[G];
#pragma unroll
    for (int a = 0; a < G; ++a) {
        v[a] = 0; dd[a] = 0; dv[a] = 0;
#pragma unroll
        for (int t = 0; t < G; ++t) { Gd[a][t] = 0; Gv[a][t] = 0; }
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            double dj = cf(a * N2 + j), vj = cf(G * N2 + a * N2 + j);
            double d1j = cf1(a * N2 + j), v1j = cf1(G * N2 + a * N2 + j);
            double ph = tab.phi2[q][j];
            v[a] += vj * ph;
            dd[a] += (dj - d1j) * ph;
            dv[a] += (vj - v1j) * ph;
#pragma unroll
            for (int t = 0; t < G; ++t) {
                Gd[a][t] += dj * tab.dphi2[q][j][t];
                Gv[a][t] += vj * tab.dphi2[q][j][t];
            }
        }
    }
    double F[G][G], Fi[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int b = 0; b < G; ++b) {
            double s = (a == b) ? 1.0 : 0.0;
#pragma unroll
            for (int t = 0; t < G; ++t) s += Gd[a][t] * K[t][b];
            F[a][b] = s;
        }
    double J = inv<G>(F, Fi);
    double Bm[G][G], Hm[G][G];
#pragma unroll
    for (int t = 0; t < G; ++t)
#pragma unroll
        for (int c = 0; c < G; ++c) {
            double s = 0;
#pragma unroll
            for (int b = 0; b < G; ++b) s += K[t][b] * Fi[b][c];
            Bm[t][c] = s;
        }
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int c = 0; c < G; ++c) {
            double s = 0;
#pragma unroll
            for (int t = 0; t < G; ++t) s += Gv[a][t] * Bm[t][c];
            Hm[a][c] = s;
        }
    double p = 0;
#pragma unroll
    for (int m = 0; m < N1; ++m) p += cf(2 * G * N2 + m) * tab.phi1[q][m];
    double u[G];
#pragma unroll
    for (int a = 0; a < G; ++a) u[a] = rc.th0rho * v[a] - rc.rk * dd[a];

    const double Jw = J * tab.w[q] * adet;
    const double k4 = rc.th0 * rc.mu * Jw;
    out(I::JW, Jw);
    out(I::MASSJ, massj);
#pragma unroll
    for (int t = 0; t < G; ++t) {
        double bu = 0, hu = 0;
#pragma unroll
        for (int c = 0; c < G; ++c) { bu += Bm[t][c] * u[c]; hu += Hm[t][c] * u[c]; }
        out(I::BUJ + t, Jw * bu);
        out(I::XJ + t, Jw * (rc.c1 * dv[t] + hu));
#pragma unroll
        for (int c = 0; c < G; ++c) {
            out(I::B + t * G + c, Bm[t][c]);
            out(I::H + t * G + c, Hm[t][c]);
            double r = 0, s = 0;
#pragma unroll
            for (int e = 0; e < G; ++e) {
                r += Bm[t][e] * Hm[e][c];
                s += Bm[t][e] * (Hm[e][c] + Hm[c][e]);
            }
            out(I::RJ + t * G + c, k4 * r);
            out(I::PJ + t * G + c, k4 * s - Jw * p * Bm[t][c]);
            if (c >= t) {
                double m = 0;
#pragma unroll
                for (int e = 0; e < G; ++e) m += Bm[t][e] * Bm[c][e];
                out(I::MJ + symi(t, c, G), k4 * m);
            }
        }
    }
}

// C' of the block  psi_a x delta v_b   (rows: velocity test functions, cols: velocity trial)
//   N1  c1 J phi_j phi_i d_ab ; N2+N6  d_ab J (g_j.u) phi_i + th0 rho J H_ab phi_j phi_i ;
//   N4  th0 mu J (d_ab g_i.g_j + g_ja g_ib)
template <int G, class BV>
TFSI_HD void fluid_C_psi_dv(BV bv, const RegionConst& rc, int a, int b, double (&C)[G + 1][G + 1])
{
    typedef FluidBase<G> I;
    const double Jw = bv(I::JW);
    const double dab = (a == b) ? 1.0 : 0.0;
    C[0][0] = Jw * (rc.c1 * dab + rc.th0rho * bv(I::H + a * G + b)) + dab * bv(I::MASSJ);
    const double k4 = Jw * rc.th0 * rc.mu;
    double kBa[G];
#pragma unroll
    for (int t = 0; t < G; ++t) {
        C[0][1 + t] = dab * bv(I::BUJ + t);
        C[1 + t][0] = 0.0;
        kBa[t] = k4 * bv(I::B + t * G + a);
    }
#pragma unroll
    for (int s = 0; s < G; ++s) {
        const double Bsb = bv(I::B + s * G + b);
#pragma unroll
        for (int t = 0; t < G; ++t)
            C[1 + s][1 + t] = dab * bv(I::MJ + symi(s, t, G)) + kBa[t] * Bsb;
    }
}

// C' of the block  psi_a x delta d_b  (shape derivatives of N1..N4, N6):
//   C'_st = PJ_sa B_tb - PJ_ta B_sb - H_ab MJ_st - B_ta RJ_sb
template <int G, class BV>
TFSI_HD void fluid_C_psi_dd(BV bv, const RegionConst& rc, int a, int b, double (&C)[G + 1][G + 1])
{
    typedef FluidBase<G> I;
    const double Hab = bv(I::H + a * G + b);
    const double XJa = bv(I::XJ + a);
    C[0][0] = -bv(I::JW) * rc.rk * Hab;
    double Ba[G], Bb[G], Pa[G], Rb[G];
#pragma unroll
    for (int t = 0; t < G; ++t) {
        Ba[t] = bv(I::B + t * G + a);
        Bb[t] = bv(I::B + t * G + b);
        Pa[t] = bv(I::PJ + t * G + a);
        Rb[t] = bv(I::RJ + t * G + b);
        C[0][1 + t] = XJa * Bb[t] - Hab * bv(I::BUJ + t);
        C[1 + t][0] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < G; ++s)
#pragma unroll
        for (int t = 0; t < G; ++t)
            C[1 + s][1 + t] = Pa[s] * Bb[t] - Pa[t] * Bb[s] - Hab * bv(I::MJ + symi(s, t, G)) - Ba[t] * Rb[s];
}

// vectors of the small blocks: psi_a x delta p (test gradient slots), gamma x delta v_b and
// gamma x delta d_b (trial gradient slots)
template <int G, class BV>
TFSI_HD void fluid_c_psi_dp(BV bv, int a, double (&c)[G]) {
    typedef FluidBase<G> I;
    const double Jw = bv(I::JW);
#pragma unroll
    for (int s = 0; s < G; ++s) c[s] = -Jw * bv(I::B + s * G + a);
}
template <int G, class BV>
TFSI_HD void fluid_c_gamma_dv(BV bv, int b, double (&c)[G]) {
    typedef FluidBase<G> I;
    const double Jw = bv(I::JW);
#pragma unroll
    for (int t = 0; t < G; ++t) c[t] = Jw * bv(I::B + t * G + b);
}
template <int G, class BV>
TFSI_HD void fluid_c_gamma_dd(BV bv, int b, double (&c)[G]) {
    typedef FluidBase<G> I;
    const double Jw = bv(I::JW);
    double trH = 0, Hb[G];
#pragma unroll
    for (int e = 0; e < G; ++e) { trH += bv(I::H + e * G + e); Hb[e] = bv(I::H + e * G + b); }
#pragma unroll
    for (int t = 0; t < G; ++t) {
        double r = 0;                                     // (B H)_tb
#pragma unroll
        for (int e = 0; e < G; ++e) r += bv(I::B + t * G + e) * Hb[e];
        c[t] = Jw * (bv(I::B + t * G + b) * trH - r);
    }
}

// ---------------------------------------------------------------------------------------------
// SOLID, non-linear Jacobian (theta0 dP : grad psi, SVK)
template <int G> struct SolidBase {
    static constexpr int JW = 0;                        // theta0 w |detJgeo|
    static constexpr int KSK = 1;                       // K S K^T  packed sym
    static constexpr int KF = KSK + Dim<G>::NS;         // KF[a][s] = sum_c K[s][c] F[a][c]
    static constexpr int FFT = KF + G * G;              // F F^T packed sym
    static constexpr int KK = FFT + Dim<G>::NS;         // K K^T packed sym
    static constexpr int COUNT = KK + Dim<G>::NS;       // 28 (3-D), 14 (2-D)
};

template <int G, class Tab, class CF, class OUT>
TFSI_HD void solid_base(const Tab& tab, int q, const double (&K)[G][G], double adet,
                        const RegionConst& rc, CF cf, OUT out)
{
    constexpr int N2 = Dim<G>::N2;
    typedef SolidBase<G> I;
    double Gd[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a) {
#pragma unroll
        for (int t = 0; t < G; ++t) Gd[a][t] = 0;
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            double dj = cf(a * N2 + j);
#pragma unroll
            for (int t = 0; t < G; ++t) Gd[a][t] += dj * tab.dphi2[q][j][t];
        }
    }
    double F[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int b = 0; b < G; ++b) {
            double s = (a == b) ? 1.0 : 0.0;
#pragma unroll
            for (int t = 0; t < G; ++t) s += Gd[a][t] * K[t][b];
            F[a][b] = s;
        }
    // E = (F^T F - I)/2 ; S = 2 mu E + lam tr(E) I          (common.py:84-88,104-109)
    double S[G][G], trE = 0;
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int b = 0; b < G; ++b) {
            double s = 0;
#pragma unroll
            for (int c = 0; c < G; ++c) s += F[c][a] * F[c][b];
            double E = 0.5 * (s - ((a == b) ? 1.0 : 0.0));
            S[a][b] = 2.0 * rc.mu * E;
            if (a == b) trE += E;
        }
#pragma unroll
    for (int a = 0; a < G; ++a) S[a][a] += rc.lam * trE;
    out(I::JW, rc.th0 * tab.w[q] * adet);
    double KS[G][G];
#pragma unroll
    for (int s = 0; s < G; ++s)
#pragma unroll
        for (int c = 0; c < G; ++c) {
            double x = 0;
#pragma unroll
            for (int e = 0; e < G; ++e) x += K[s][e] * S[e][c];
            KS[s][c] = x;
        }
#pragma unroll
    for (int s = 0; s < G; ++s)
#pragma unroll
        for (int t = s; t < G; ++t) {
            double x = 0, y = 0, z = 0;
#pragma unroll
            for (int c = 0; c < G; ++c) {
                x += KS[s][c] * K[t][c];
                y += F[s][c] * F[t][c];
                z += K[s][c] * K[t][c];
            }
            out(I::KSK + symi(s, t, G), x);
            out(I::FFT + symi(s, t, G), y);
            out(I::KK + symi(s, t, G), z);
        }
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int s = 0; s < G; ++s) {
            double x = 0;
#pragma unroll
            for (int c = 0; c < G; ++c) x += K[s][c] * F[a][c];
            out(I::KF + a * G + s, x);
        }
}

// ---------------------------------------------------------------------------------------------
// ENERGY-BASED solid materials (common.py:110-125: P = diff(W, F) with F a ufl variable; W_St_Venant_Kirchoff :214-225,
// W_Neo_Hookean :228-240, W_Gent :242-256, W_Mooney_Rivlin :258-273). Every W of the reference is a function of the
// invariants I1 = tr(F^T F), I2 = (I1^2 - tr((F^T F)^2)) / 2 and J = det F, so
//     P = W_1 dI1/dF + W_2 dI2/dF + W_J dJ/dF ,   dI1/dF = 2 F,  dI2/dF = 2 (I1 F - F C),  dJ/dF = cof F = J F^-T
//     dP/dF = sum_ab W_ab (d a/dF) (x) (d b/dF) + W_1 d2I1 + W_2 d2I2 + W_J d2J
// and only the SCALAR function W(I1, I2, J) differs between the models. Its gradient and Hessian come from evaluating
// it on hyper-dual numbers (f + a e1 + b e2 + ab e1 e2, e1^2 = e2^2 = 0): exact derivatives, one definition per model.
struct HDual { double v, a, b, ab; };
TFSI_HD HDual hd(double v) { return HDual{v, 0.0, 0.0, 0.0}; }
TFSI_HD HDual operator+(HDual x, HDual y) { return HDual{x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; }
TFSI_HD HDual operator-(HDual x, HDual y) { return HDual{x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; }
TFSI_HD HDual operator*(HDual x, HDual y) {
    return HDual{x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b, x.ab * y.v + x.a * y.b + x.b * y.a + x.v * y.ab};
}
TFSI_HD HDual operator*(double c, HDual x) { return HDual{c * x.v, c * x.a, c * x.b, c * x.ab}; }
TFSI_HD HDual operator+(HDual x, double c) { return HDual{x.v + c, x.a, x.b, x.ab}; }
TFSI_HD HDual operator-(HDual x, double c) { return HDual{x.v - c, x.a, x.b, x.ab}; }
// g(x) for a scalar function with derivatives g1, g2 at x.v
TFSI_HD HDual hd_chain(HDual x, double g0, double g1, double g2) {
    return HDual{g0, g1 * x.a, g1 * x.b, g1 * x.ab + g2 * x.a * x.b};
}
TFSI_HD HDual hd_log(HDual x) { return hd_chain(x, log(x.v), 1.0 / x.v, -1.0 / (x.v * x.v)); }
TFSI_HD HDual hd_pow(HDual x, double p) {
    const double f = pow(x.v, p);
    return hd_chain(x, f, p * f / x.v, p * (p - 1.0) * f / (x.v * x.v));
}

// W(I1, I2, J) of the model rc.mat in G dimensions (the reference's expressions, incl. their hard-coded 3s)
TFSI_HD HDual solid_energy(const RegionConst& rc, int G, HDual I1, HDual I2, HDual J) {
    switch (rc.mat) {
    case MAT_SVK_ENERGY: {        // lambda/2 tr(E)^2 + mu tr(E E),  tr E = (I1 - G)/2, tr(E E) = (I1^2 - 2 I2 - 2 I1 + G)/4
        const HDual trE = 0.5 * (I1 - (double)G);
        return (0.5 * rc.lam) * (trE * trE) + (0.25 * rc.mu) * (I1 * I1 - 2.0 * I2 - 2.0 * I1 + (double)G);
    }
    case MAT_NEO_HOOKEAN: {       // C1 (I1 - G - 2 ln J) + D1 (J - 1)^2,  C1 = mu/2, D1 = lambda/2
        const HDual Jm1 = J - 1.0;
        return (0.5 * rc.mu) * (I1 - (double)G - 2.0 * hd_log(J)) + (0.5 * rc.lam) * (Jm1 * Jm1);
    }
    case MAT_GENT: {              // -(mu/2) (Jm ln(1 - (I1 - 3)/Jm) + 2 ln J)
        const HDual arg = hd(1.0) - (1.0 / rc.mJm) * (I1 - 3.0);
        return (-0.5 * rc.mu) * (rc.mJm * hd_log(arg) + 2.0 * hd_log(J));
    }
    case MAT_MOONEY_RIVLIN: {     // C01 (Ib2 - 3) + C10 (Ib1 - 3) + C11 (Ib2 - 3)(Ib1 - 3) + (1/D1)(J - 1)^2, D1 = 2/(lambda + 2 mu/3)
        const HDual Ib1 = hd_pow(J, -2.0 / 3.0) * I1, Ib2 = hd_pow(J, -4.0 / 3.0) * I2;
        const HDual Jm1 = J - 1.0;
        return rc.mC01 * (Ib2 - 3.0) + rc.mC10 * (Ib1 - 3.0) + rc.mC11 * ((Ib2 - 3.0) * (Ib1 - 3.0)) +
               (0.5 * (rc.lam + 2.0 * rc.mu / 3.0)) * (Jm1 * Jm1);
    }
    default: return hd(0.0);
    }
}

// gradient w[3] = (W_1, W_2, W_J) and Hessian h[6] = (W_11, W_12, W_1J, W_22, W_2J, W_JJ) at (I1, I2, J)
TFSI_HD void solid_energy_derivatives(const RegionConst& rc, int G, double I1, double I2, double J, double (&w)[3], double (&h)[6]) {
    const double x[3] = {I1, I2, J};
    int k = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            HDual v[3] = {hd(x[0]), hd(x[1]), hd(x[2])};
            v[i].a = 1.0;
            v[j].b = 1.0;
            const HDual W = solid_energy(rc, G, v[0], v[1], v[2]);
            if (j == i) w[i] = W.a;
            h[k++] = W.ab;
        }
}

// kinematic pieces shared by stress and tangent
template <int G> struct MatKin {
    double C[G][G], B[G][G], cof[G][G], G2[G][G];     // F^T F, F F^T, cof F = J F^-T, dI2/dF = 2 (I1 F - F C)
    double I1, I2, J;
};
template <int G>
TFSI_HD void material_kinematics(const double (&F)[G][G], MatKin<G>& k) {
    double Fi[G][G];
    k.J = inv<G>(F, Fi);
    double trC2 = 0;
    k.I1 = 0;
    for (int i = 0; i < G; ++i)
        for (int j = 0; j < G; ++j) {
            double c = 0, b = 0;
            for (int m = 0; m < G; ++m) { c += F[m][i] * F[m][j]; b += F[i][m] * F[j][m]; }
            k.C[i][j] = c; k.B[i][j] = b;
            k.cof[i][j] = k.J * Fi[j][i];
            trC2 += c * c;
            if (i == j) k.I1 += c;
        }
    k.I2 = 0.5 * (k.I1 * k.I1 - trC2);
    for (int a = 0; a < G; ++a)
        for (int j = 0; j < G; ++j) {
            double fc = 0;
            for (int m = 0; m < G; ++m) fc += F[a][m] * k.C[m][j];
            k.G2[a][j] = 2.0 * (k.I1 * F[a][j] - fc);
        }
}

// first Piola-Kirchhoff stress of an energy-based model (kept out of line: the St.Venant-Kirchhoff hot path of the
// residual kernel must not pay registers for it)
template <int G>
TFSI_HD_NOINLINE void solid_piola_energy(const RegionConst& rc, const double (&F)[G][G], double (&P)[G][G]) {
    MatKin<G> k;
    material_kinematics<G>(F, k);
    double w[3], h[6];
    solid_energy_derivatives(rc, G, k.I1, k.I2, k.J, w, h);
    for (int a = 0; a < G; ++a)
        for (int j = 0; j < G; ++j) P[a][j] = 2.0 * w[0] * F[a][j] + w[1] * k.G2[a][j] + w[2] * k.cof[a][j];
}

// per-point base of the non-linear solid Jacobian for energy-based models: the tangent dP/dF contracted with the
// geometry on both gradient slots and pre-scaled,  T[(a,s)][(b,t)] = theta0 w |detJ| sum_JL K_sJ (dP_aJ/dF_bL) K_tL,
// symmetric in (a,s) <-> (b,t) (hyperelasticity): n (n + 1) / 2 entries, n = G*G.
template <int G> struct SolidBaseGen {
    static constexpr int JW = 0;
    static constexpr int TAN = 1;
    static constexpr int NT = (G * G) * (G * G + 1) / 2;   // 45 (3-D), 10 (2-D)
    static constexpr int COUNT = TAN + NT;
};

template <int G, class Tab, class CF, class OUT>
TFSI_HD void solid_base_energy(const Tab& tab, int q, const double (&K)[G][G], double adet,
                               const RegionConst& rc, CF cf, OUT out)
{
    constexpr int N2 = Dim<G>::N2;
    typedef SolidBaseGen<G> I;
    double F[G][G];
    for (int a = 0; a < G; ++a) {
        double Gd[G];
        for (int t = 0; t < G; ++t) Gd[t] = 0;
        for (int j = 0; j < N2; ++j) {
            const double dj = cf(a * N2 + j);
            for (int t = 0; t < G; ++t) Gd[t] += dj * tab.dphi2[q][j][t];
        }
        for (int b = 0; b < G; ++b) {
            double s = (a == b) ? 1.0 : 0.0;
            for (int t = 0; t < G; ++t) s += Gd[t] * K[t][b];
            F[a][b] = s;
        }
    }
    MatKin<G> k;
    material_kinematics<G>(F, k);
    double w[3], h[6];
    solid_energy_derivatives(rc, G, k.I1, k.I2, k.J, w, h);
    const double jw = rc.th0 * tab.w[q] * adet;
    out(I::JW, jw);
    // geometry-contracted pieces  X~[a][s] = sum_J K[s][J] X[a][J]
    double Ft[G][G], G2t[G][G], Ct[G][G], KK[G][G], KCK[G][G];
    for (int a = 0; a < G; ++a)
        for (int s = 0; s < G; ++s) {
            double x = 0, y = 0, z = 0;
            for (int j = 0; j < G; ++j) { x += K[s][j] * F[a][j]; y += K[s][j] * k.G2[a][j]; z += K[s][j] * k.cof[a][j]; }
            Ft[a][s] = x; G2t[a][s] = y; Ct[a][s] = z;
        }
    for (int s = 0; s < G; ++s)
        for (int t = 0; t < G; ++t) {
            double x = 0, y = 0;
            for (int j = 0; j < G; ++j) {
                x += K[s][j] * K[t][j];
                double kc = 0;
                for (int l = 0; l < G; ++l) kc += k.C[l][j] * K[t][l];
                y += K[s][j] * kc;
            }
            KK[s][t] = x; KCK[s][t] = y;
        }
    // T_alpha[b][t] = sum_beta W_alpha,beta (d beta/dF)~[b][t]   with (d I1/dF)~ = 2 F~
    const double H[3][3] = {{h[0], h[1], h[2]}, {h[1], h[3], h[4]}, {h[2], h[4], h[5]}};
    double T[3][G][G];
    for (int al = 0; al < 3; ++al)
        for (int b = 0; b < G; ++b)
            for (int t = 0; t < G; ++t)
                T[al][b][t] = 2.0 * H[al][0] * Ft[b][t] + H[al][1] * G2t[b][t] + H[al][2] * Ct[b][t];
    const double wJ_J = w[2] / k.J;
    for (int a = 0; a < G; ++a)
        for (int s = 0; s < G; ++s)
            for (int b = 0; b < G; ++b)
                for (int t = 0; t < G; ++t) {
                    const int p = a * G + s, r = b * G + t;
                    if (r < p) continue;
                    const double dab = (a == b) ? 1.0 : 0.0;
                    double v = 2.0 * Ft[a][s] * T[0][b][t] + G2t[a][s] * T[1][b][t] + Ct[a][s] * T[2][b][t];
                    v += dab * ((2.0 * w[0] + 2.0 * w[1] * k.I1) * KK[s][t] - 2.0 * w[1] * KCK[s][t]);
                    v += 2.0 * w[1] * (2.0 * Ft[b][t] * Ft[a][s] - Ft[a][t] * Ft[b][s] - k.B[a][b] * KK[s][t]);
                    v += wJ_J * (Ct[b][t] * Ct[a][s] - Ct[a][t] * Ct[b][s]);
                    out(I::TAN + symi(p, r, G * G), jw * v);
                }
}

// C' (gradient slots only) of psi_a x delta d_b for energy-based models: read from the tangent table
template <int G, class BV>
TFSI_HD void solid_C_psi_dd_energy(BV bv, int a, int b, double (&C)[G + 1][G + 1])
{
    typedef SolidBaseGen<G> I;
    C[0][0] = 0.0;
#pragma unroll
    for (int t = 0; t < G; ++t) { C[0][1 + t] = 0.0; C[1 + t][0] = 0.0; }
#pragma unroll
    for (int s = 0; s < G; ++s)
#pragma unroll
        for (int t = 0; t < G; ++t) C[1 + s][1 + t] = bv(I::TAN + symi(a * G + s, b * G + t, G * G));
}

// C' (gradient slots only) of  psi_a x delta d_b :
//   d_ab G_j.S G_i + mu (f_a.G_j)(f_b.G_i) + mu (F F^T)_ab G_i.G_j + lam (f_a.G_i)(f_b.G_j)
template <int G, class BV>
TFSI_HD void solid_C_psi_dd(BV bv, const RegionConst& rc, int a, int b, double (&C)[G + 1][G + 1])
{
    typedef SolidBase<G> I;
    const double jw = bv(I::JW);
    const double dab = (a == b) ? 1.0 : 0.0;
    const double muFF = rc.mu * bv(I::FFT + symi(a, b, G));
    C[0][0] = 0.0;
#pragma unroll
    for (int t = 0; t < G; ++t) { C[0][1 + t] = 0.0; C[1 + t][0] = 0.0; }
#pragma unroll
    for (int s = 0; s < G; ++s)
#pragma unroll
        for (int t = 0; t < G; ++t)
            C[1 + s][1 + t] = jw * (dab * bv(I::KSK + symi(s, t, G))
                                    + rc.mu * bv(I::KF + b * G + s) * bv(I::KF + a * G + t)
                                    + muFF * bv(I::KK + symi(s, t, G))
                                    + rc.lam * bv(I::KF + a * G + s) * bv(I::KF + b * G + t));
}

// ---------------------------------------------------------------------------------------------
// Sum-factorised block updates, exploiting the SPARSITY of the P2 reference gradients:
//   vertex function v :  ghat = (4 lam_v - 1) dlam_v            dlam_0 = -(1,..,1), dlam_k = e_k
//   edge function (a,b): ghat = 4 lam_a dlam_b + 4 lam_b dlam_a
// so a vertex column/row costs 2 FMAs per slot pair instead of G+1, an edge 3.  The columns of
// one block are split into two halves of equal cost (a thread owns N2 x N2/2 accumulators).
// MODE bits: 1 = C[0][0] present, 2 = C[0][t] present, 4 = C[s][t] present (C[s][0] is
// structurally zero in every form of this path).
template <int G> struct P2Edges;
template <> struct P2Edges<2> {
    TFSI_HD static constexpr int a(int e) { return e == 0 ? 1 : 0; }
    TFSI_HD static constexpr int b(int e) { return e == 2 ? 1 : 2; }
};
template <> struct P2Edges<3> {
    TFSI_HD static constexpr int a(int e) { return e == 0 ? 2 : (e <= 2 ? 1 : 0); }
    TFSI_HD static constexpr int b(int e) { return e == 0 ? 3 : (e == 1 ? 3 : (e == 2 ? 2 : (e == 3 ? 3 : (e == 4 ? 2 : 1)))); }
};
// column lists of the two halves (2 vertices + 3 edges each in 3-D)
template <int G, int HALF> struct HalfCols;
template <int HALF> struct HalfCols<3, HALF> {
    static constexpr int NJ = 5;
    TFSI_HD static constexpr int col(int jj) { return HALF == 0 ? (jj < 2 ? jj : jj + 2) : (jj < 2 ? jj + 2 : jj + 5); }
};
template <int HALF> struct HalfCols<2, HALF> {
    static constexpr int NJ = 3;
    TFSI_HD static constexpr int col(int jj) { return HALF == 0 ? (jj == 0 ? 0 : jj + 2) : (jj < 2 ? jj + 1 : 5); }
};

template <int G, int HALF, int MODE, class Tab>
TFSI_HD void block_p2p2(const Tab& tab, int q, const double (&C)[G + 1][G + 1],
                        double (&Z)[Dim<G>::N2][HalfCols<G, HALF>::NJ])
{
    constexpr int N2 = Dim<G>::N2;
    constexpr int NJ = HalfCols<G, HALF>::NJ;
    typedef P2Edges<G> E;
    // row sums over the gradient slots (used by every function touching vertex 0)
    double Cs[G + 1];
#pragma unroll
    for (int s = 0; s <= G; ++s) {
        double x = 0.0;
        if ((s == 0 && (MODE & 2)) || (s > 0 && (MODE & 4))) {
            x = C[s][1];                                  // not 0.0 + C: `0.0 + x` is a real DADD (signed zero)
#pragma unroll
            for (int t = 2; t <= G; ++t) x += C[s][t];
        }
        Cs[s] = x;
    }
    // stage 1 of column jj + 1 is written BEFORE stage 2 of column jj: with only 3 warps per scheduler the FP64
    // pipe needs several independent FMAs per warp, and the stage-2 chains of a column all wait for its W
    auto stage1 = [&](int jj, double (&W)[G + 1]) {
        const int j = HalfCols<G, HALF>::col(jj);
        if (j == 0) {
            const double D = tab.d4[q][0];
            W[0] = 0.0;
            if (MODE & 1) W[0] = C[0][0] * tab.phi2[q][0];
            if (MODE & 2) W[0] -= D * Cs[0];
#pragma unroll
            for (int s = 1; s <= G; ++s) W[s] = (MODE & 4) ? -D * Cs[s] : 0.0;
        } else if (j <= G) {
            const double D = tab.d4[q][j];
            W[0] = 0.0;
            if (MODE & 1) W[0] = C[0][0] * tab.phi2[q][j];
            if (MODE & 2) W[0] += D * C[0][j];
#pragma unroll
            for (int s = 1; s <= G; ++s) W[s] = (MODE & 4) ? D * C[s][j] : 0.0;
        } else {
            const int ea = E::a(j - G - 1), eb = E::b(j - G - 1);
            const double La = tab.l4[q][ea], Lb = tab.l4[q][eb];
            W[0] = 0.0;
            if (MODE & 1) W[0] = C[0][0] * tab.phi2[q][j];
            if (ea == 0) {
                if (MODE & 2) { W[0] = fma(La, C[0][eb], W[0]); W[0] = fma(-Lb, Cs[0], W[0]); }
#pragma unroll
                for (int s = 1; s <= G; ++s) W[s] = (MODE & 4) ? La * C[s][eb] - Lb * Cs[s] : 0.0;
            } else {
                if (MODE & 2) { W[0] = fma(La, C[0][eb], W[0]); W[0] = fma(Lb, C[0][ea], W[0]); }
#pragma unroll
                for (int s = 1; s <= G; ++s) W[s] = (MODE & 4) ? La * C[s][eb] + Lb * C[s][ea] : 0.0;
            }
        }
    };
    auto stage2 = [&](int jj, const double (&W)[G + 1]) {
        double Wsum = 0.0;
        if (MODE & 4) {
            Wsum = W[1];
#pragma unroll
            for (int s = 2; s <= G; ++s) Wsum += W[s];
        }
        // one FMA per term (`z += a*b + c*d` would cost DMUL + DFMA + DADD), written breadth-first over the rows
        // so that the FMAs of one dependency chain are a whole pass apart
        if (MODE & 3) {
#pragma unroll
            for (int i = 0; i < N2; ++i) Z[i][jj] = fma(tab.phi2[q][i], W[0], Z[i][jj]);
        }
        if (MODE & 4) {
#pragma unroll
            for (int i = 0; i < N2; ++i) {
                if (i == 0) Z[i][jj] = fma(-tab.d4[q][0], Wsum, Z[i][jj]);
                else if (i <= G) Z[i][jj] = fma(tab.d4[q][i], W[i], Z[i][jj]);
                else Z[i][jj] = fma(tab.l4[q][E::a(i - G - 1)], W[E::b(i - G - 1)], Z[i][jj]);
            }
#pragma unroll
            for (int i = G + 1; i < N2; ++i) {
                const int ea = E::a(i - G - 1), eb = E::b(i - G - 1);
                if (ea == 0) Z[i][jj] = fma(-tab.l4[q][eb], Wsum, Z[i][jj]);
                else Z[i][jj] = fma(tab.l4[q][eb], W[ea], Z[i][jj]);
            }
        }
    };
    double Wa[G + 1], Wb[G + 1];
    stage1(0, Wa);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        if (jj % 2 == 0) {
            if (jj + 1 < NJ) stage1(jj + 1, Wb);
            stage2(jj, Wa);
        } else {
            if (jj + 1 < NJ) stage1(jj + 1, Wa);
            stage2(jj, Wb);
        }
    }
}

// z_i += phi_i W[0] + ghat_i . W[1..G]  for all P2 functions i (sparse reference gradients, as stage 2)
template <int G, class Tab>
TFSI_HD void test_p2_sparse(const Tab& tab, int q, const double (&W)[G + 1], double (&Z)[Dim<G>::N2])
{
    constexpr int N2 = Dim<G>::N2;
    typedef P2Edges<G> E;
    double Wsum = W[1];
#pragma unroll
    for (int s = 2; s <= G; ++s) Wsum += W[s];
#pragma unroll
    for (int i = 0; i < N2; ++i) {
        double z = fma(tab.phi2[q][i], W[0], Z[i]);
        if (i == 0) z = fma(-tab.d4[q][0], Wsum, z);
        else if (i <= G) z = fma(tab.d4[q][i], W[i], z);
        else {
            const int ea = E::a(i - G - 1), eb = E::b(i - G - 1);
            if (ea == 0) { z = fma(tab.l4[q][0], W[eb], z); z = fma(-tab.l4[q][eb], Wsum, z); }
            else { z = fma(tab.l4[q][ea], W[eb], z); z = fma(tab.l4[q][eb], W[ea], z); }
        }
        Z[i] = z;
    }
}

// P2 gradient-test x P1 value-trial :  Z[i][m] += (ghat_i . c) chi_m
template <int G, class Tab>
TFSI_HD void block_gradp2_p1(const Tab& tab, int q, const double (&c)[G],
                             double (&Z)[Dim<G>::N2][Dim<G>::N1])
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    double W[N1][G];
#pragma unroll
    for (int m = 0; m < N1; ++m)
#pragma unroll
        for (int s = 0; s < G; ++s) W[m][s] = c[s] * tab.phi1[q][m];
#pragma unroll
    for (int i = 0; i < N2; ++i)
#pragma unroll
        for (int m = 0; m < N1; ++m) {
            double z = Z[i][m];
#pragma unroll
            for (int s = 0; s < G; ++s) z += tab.dphi2[q][i][s] * W[m][s];
            Z[i][m] = z;
        }
}

// P1 value-test x P2 gradient-trial :  Z[m][j] += chi_m (c . ghat_j)
template <int G, class Tab>
TFSI_HD void block_p1_gradp2(const Tab& tab, int q, const double (&c)[G],
                             double (&Z)[Dim<G>::N1][Dim<G>::N2])
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
#pragma unroll
    for (int j = 0; j < N2; ++j) {
        double w = 0;
#pragma unroll
        for (int t = 0; t < G; ++t) w += c[t] * tab.dphi2[q][j][t];
#pragma unroll
        for (int m = 0; m < N1; ++m) Z[m][j] += tab.phi1[q][m] * w;
    }
}

// ---------------------------------------------------------------------------------------------
// RESIDUAL: per-point fluxes of F = F_lin + F_nonlin for one cell (b = -F is scattered).
// fval[a], fgrad[a][s]: coefficient of phi_i / ghat_i,s in row (field, a, i).
template <int G> struct PointFlux {
    double psi_val[G], psi_grad[G][G];   // velocity-test rows
    double phi_val[G], phi_grad[G][G];   // displacement-test rows
    double beta_val[G], beta_grad[G][G]; // w-test rows (biharmonic)
    double gamma;                        // pressure-test rows
};

// LIFT: 1 = laplace "volume_change" (alfa = 1/J(d^n), laplace.py:33-34), 2 = elastic lifting (elastic.py:18-27)
// compiled in; a template flag so that the default instantiation keeps its register allocation
// 3 = general solid (energy-based material models common.py:110-125 and / or viscoelasticity solid.py:46-59) compiled in
constexpr int LIFT_DEFAULT = 0, LIFT_VOLUME_CHANGE = 1, LIFT_ELASTIC = 2, LIFT_SOLID_GENERAL = 3;
template <int G, int LIFT = 0, class Tab, class CF, class CF1>
TFSI_HD void residual_point(const Tab& tab, int q, const double (&K)[G][G], double adet,
                            const RegionConst& rc, double alfa_cell, CF cf, CF1 cf1,
                            PointFlux<G>& f)
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    const double jw = tab.w[q] * adet;
    // field values and reference gradients at n and n-1. Gradients through the barycentric derivatives of
    // the P2 basis, which are sparse (vertex i: (4 lambda_i - 1) e_i; edge (a,b): 4 lambda_b e_a + 4 lambda_a e_b):
    // 16 FMAs + 3 differences per component instead of 30 FMAs; d/dxi_t = d/dlambda_{t+1} - d/dlambda_0.
    typedef P2Edges<G> E;
    double d[G], v[G], d1[G], v1[G], Gd[G][G], Gv[G][G], Gd1[G][G], Gv1[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a) {
        double Ld[G + 1], Lv[G + 1], Ld1[G + 1], Lv1[G + 1];
        d[a] = v[a] = d1[a] = v1[a] = 0;
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            const double dj = cf(a * N2 + j), vj = cf(G * N2 + a * N2 + j);
            const double d1j = cf1(a * N2 + j), v1j = cf1(G * N2 + a * N2 + j);
            const double ph = tab.phi2[q][j];
            d[a] = fma(dj, ph, d[a]); v[a] = fma(vj, ph, v[a]); d1[a] = fma(d1j, ph, d1[a]); v1[a] = fma(v1j, ph, v1[a]);
            if (j <= G) {
                const double g = tab.d4[q][j];
                Ld[j] = dj * g; Lv[j] = vj * g; Ld1[j] = d1j * g; Lv1[j] = v1j * g;
            } else {
                const int ea = E::a(j - G - 1), eb = E::b(j - G - 1);
                const double la = tab.l4[q][ea], lb = tab.l4[q][eb];
                Ld[ea] = fma(dj, lb, Ld[ea]); Ld[eb] = fma(dj, la, Ld[eb]);
                Lv[ea] = fma(vj, lb, Lv[ea]); Lv[eb] = fma(vj, la, Lv[eb]);
                Ld1[ea] = fma(d1j, lb, Ld1[ea]); Ld1[eb] = fma(d1j, la, Ld1[eb]);
                Lv1[ea] = fma(v1j, lb, Lv1[ea]); Lv1[eb] = fma(v1j, la, Lv1[eb]);
            }
        }
#pragma unroll
        for (int t = 0; t < G; ++t) {
            Gd[a][t] = Ld[t + 1] - Ld[0]; Gv[a][t] = Lv[t + 1] - Lv[0];
            Gd1[a][t] = Ld1[t + 1] - Ld1[0]; Gv1[a][t] = Lv1[t + 1] - Lv1[0];
        }
    }
#pragma unroll
    for (int a = 0; a < G; ++a) {
        f.psi_val[a] = f.phi_val[a] = f.beta_val[a] = 0;
#pragma unroll
        for (int s = 0; s < G; ++s) f.psi_grad[a][s] = f.phi_grad[a][s] = f.beta_grad[a][s] = 0;
    }
    f.gamma = 0;
    // deformation gradients
    double F[G][G], F1[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int b = 0; b < G; ++b) {
            double s = (a == b) ? 1.0 : 0.0, s1 = s;
#pragma unroll
            for (int t = 0; t < G; ++t) { s += Gd[a][t] * K[t][b]; s1 += Gd1[a][t] * K[t][b]; }
            F[a][b] = s; F1[a][b] = s1;
        }
    if (rc.kind == 0) {
        if (rc.fluid_on) {
            double Fi[G][G], Fi1[G][G];
            double J = inv<G>(F, Fi), J1 = inv<G>(F1, Fi1);
            double B[G][G], B1[G][G], H[G][G], H1[G][G];
#pragma unroll
            for (int t = 0; t < G; ++t)
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    double s = 0, s1 = 0;
#pragma unroll
                    for (int b = 0; b < G; ++b) { s += K[t][b] * Fi[b][c]; s1 += K[t][b] * Fi1[b][c]; }
                    B[t][c] = s; B1[t][c] = s1;
                }
#pragma unroll
            for (int a = 0; a < G; ++a)
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    double s = 0, s1 = 0;
#pragma unroll
                    for (int t = 0; t < G; ++t) { s += Gv[a][t] * B[t][c]; s1 += Gv1[a][t] * B1[t][c]; }
                    H[a][c] = s; H1[a][c] = s1;
                }
            double p = 0;
#pragma unroll
            for (int m = 0; m < N1; ++m) p += cf(2 * G * N2 + m) * tab.phi1[q][m];
            double trH = 0;
#pragma unroll
            for (int a = 0; a < G; ++a) {
                trH += H[a][a];
                double Hu = 0, H1v1 = 0;
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    Hu += H[a][c] * (rc.th0rho * v[c] - rc.rk * (d[c] - d1[c]));   // N2 + N6
                    H1v1 += H1[a][c] * v1[c];                                        // L2
                }
                f.psi_val[a] = jw * ((rc.c1 * J + rc.c1lin * J1) * (v[a] - v1[a])     // N1 + L1
                                     + J * Hu + rc.th1rho * J1 * H1v1);
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    // [B (-p I + th0 sigma)]_{s,a} J + th1 J1 [B1 sigma1]_{s,a}       N3, N4, L4
                    double bs = 0, bs1 = 0;
#pragma unroll
                    for (int c = 0; c < G; ++c) {
                        bs += B[s][c] * (H[c][a] + H[a][c]);
                        bs1 += B1[s][c] * (H1[c][a] + H1[a][c]);
                    }
                    f.psi_grad[a][s] = jw * (J * (-p * B[s][a] + rc.th0 * rc.mu * bs)
                                             + rc.th1 * J1 * rc.mu * bs1);
                }
            }
            f.gamma = jw * J * trH;                                                  // N5
        }
        if (rc.ext == 1) {                       // laplace.py:45  alfa grad d : grad phi
            double al = alfa_cell;
            if (LIFT == LIFT_VOLUME_CHANGE) {    // laplace.py:33-34 "volume_change": alfa = 1 / J(d^n), pointwise
                double Fi_[G][G];
                al = 1.0 / inv<G>(F, Fi_);
            }
#pragma unroll
            for (int a = 0; a < G; ++a)
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    double x = 0;                // (K K^T Ghat_d^T)_{s,a}
#pragma unroll
                    for (int b = 0; b < G; ++b) {
                        double gdab = 0;         // grad d [a][b]
#pragma unroll
                        for (int t = 0; t < G; ++t) gdab += Gd[a][t] * K[t][b];
                        x += K[s][b] * gdab;
                    }
                    f.phi_grad[a][s] = jw * al * x;
                }
        } else if (LIFT == LIFT_ELASTIC && rc.ext == 3) {
            // elastic.py:18-27: J S_linear(d) F^-T : grad(phi);  eps(d) = sym(grad d F^-1) (common.py:56-60),
            // E = 1 / CellVolume, nu = 0.25; alfa_cell carries E
            const double nu = 0.25, a_lam = nu * alfa_cell / ((1.0 + nu) * (1.0 - 2.0 * nu)), a_mu = alfa_cell / (2.0 * (1.0 + nu));
            double Fi[G][G];
            const double J = inv<G>(F, Fi);
            double tre = 0, S[G][G];             // grad d F^-1 = I - F^-1
#pragma unroll
            for (int a = 0; a < G; ++a) tre += 1.0 - Fi[a][a];
#pragma unroll
            for (int a = 0; a < G; ++a)
#pragma unroll
                for (int b = 0; b < G; ++b)
                    S[a][b] = a_mu * (((a == b) ? 2.0 : 0.0) - Fi[a][b] - Fi[b][a]) + ((a == b) ? a_lam * tre : 0.0);
#pragma unroll
            for (int a = 0; a < G; ++a) {
                double P[G];                     // (J S F^-T)[a][b]
#pragma unroll
                for (int b = 0; b < G; ++b) {
                    double x = 0;
#pragma unroll
                    for (int c = 0; c < G; ++c) x += S[a][c] * Fi[b][c];
                    P[b] = J * x;
                }
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    double x = 0;
#pragma unroll
                    for (int b = 0; b < G; ++b) x += K[s][b] * P[b];
                    f.phi_grad[a][s] = jw * x;
                }
            }
        } else if (rc.ext == 2) {                // biharmonic.py:35-36
            double w[G], Gw[G][G];
#pragma unroll
            for (int a = 0; a < G; ++a) {
                w[a] = 0;
#pragma unroll
                for (int t = 0; t < G; ++t) Gw[a][t] = 0;
#pragma unroll
                for (int j = 0; j < N2; ++j) {
                    double wj = cf(2 * G * N2 + N1 + a * N2 + j);
                    w[a] += wj * tab.phi2[q][j];
#pragma unroll
                    for (int t = 0; t < G; ++t) Gw[a][t] += wj * tab.dphi2[q][j][t];
                }
            }
#pragma unroll
            for (int a = 0; a < G; ++a) {
                f.beta_val[a] = jw * rc.alfa * w[a];
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    double xd = 0, xw = 0;
#pragma unroll
                    for (int b = 0; b < G; ++b) {
                        double gd = 0, gw = 0;
#pragma unroll
                        for (int t = 0; t < G; ++t) { gd += Gd[a][t] * K[t][b]; gw += Gw[a][t] * K[t][b]; }
                        xd += K[s][b] * gd; xw += K[s][b] * gw;
                    }
                    f.beta_grad[a][s] = -jw * rc.alfa * xd;
                    f.phi_grad[a][s] = jw * rc.alfa * xw;
                }
            }
        }
    } else if (rc.solid_on) {
        // solid.py:41-43, 62-63, 65-68
        auto piolaK = [&](const double (&Fm)[G][G], double scale, double (&out)[G][G]) {
            double S[G][G], trE = 0;
#pragma unroll
            for (int a = 0; a < G; ++a)
#pragma unroll
                for (int b = 0; b < G; ++b) {
                    double s = 0;
#pragma unroll
                    for (int c = 0; c < G; ++c) s += Fm[c][a] * Fm[c][b];
                    double E = 0.5 * (s - ((a == b) ? 1.0 : 0.0));
                    S[a][b] = 2.0 * rc.mu * E;
                    if (a == b) trE += E;
                }
#pragma unroll
            for (int a = 0; a < G; ++a) S[a][a] += rc.lam * trE;
            // P = F S ; out[a][s] += scale * sum_b K[s][b] P[a][b]
#pragma unroll
            for (int a = 0; a < G; ++a) {
                double P[G];
#pragma unroll
                for (int b = 0; b < G; ++b) {
                    double x = 0;
#pragma unroll
                    for (int c = 0; c < G; ++c) x += Fm[a][c] * S[c][b];
                    P[b] = x;
                }
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    double x = 0;
#pragma unroll
                    for (int b = 0; b < G; ++b) x += K[s][b] * P[b];
                    out[a][s] += scale * x;
                }
            }
        };
        auto piolaE = [&](const double (&Fm)[G][G], double scale, double (&out)[G][G]) {   // energy-based models
            double P[G][G];
            solid_piola_energy<G>(rc, Fm, P);
#pragma unroll
            for (int a = 0; a < G; ++a)
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    double x = 0;
#pragma unroll
                    for (int b = 0; b < G; ++b) x += K[s][b] * P[a][b];
                    out[a][s] += scale * x;
                }
        };
        if (LIFT != LIFT_SOLID_GENERAL || rc.mat == MAT_SVK) {
            piolaK(F, jw * rc.th0, f.psi_grad);
            piolaK(F1, jw * rc.th1, f.psi_grad);
        } else {
            piolaE(F, jw * rc.th0, f.psi_grad);
            piolaE(F1, jw * rc.th1, f.psi_grad);
        }
        if (LIFT == LIFT_SOLID_GENERAL && rc.visco) {
            // solid.py:46-59: + theta0 F(d) Svisc_D(w) : grad psi + theta1 F(d1) Svisc_D(w1) : grad psi with the rate
            // w = v, w1 = v1 ("Form1") or w = w1 = (d - d1) / k ("Form2"); Svisc_D of common.py:129-139
            auto viscK = [&](const double (&Fm)[G][G], const double (&Gw)[G][G], double scale, double (&out)[G][G]) {
                double S[G][G], tr = 0;
#pragma unroll
                for (int a = 0; a < G; ++a) tr += Gw[a][a];
#pragma unroll
                for (int a = 0; a < G; ++a)
#pragma unroll
                    for (int b = 0; b < G; ++b) S[a][b] = rc.lam_visc * (Gw[a][b] + Gw[b][a]) + ((a == b) ? rc.mu_visc * tr : 0.0);
#pragma unroll
                for (int a = 0; a < G; ++a) {
                    double P[G];
#pragma unroll
                    for (int b = 0; b < G; ++b) {
                        double x = 0;
#pragma unroll
                        for (int c = 0; c < G; ++c) x += Fm[a][c] * S[c][b];
                        P[b] = x;
                    }
#pragma unroll
                    for (int s = 0; s < G; ++s) {
                        double x = 0;
#pragma unroll
                        for (int b = 0; b < G; ++b) x += K[s][b] * P[b];
                        out[a][s] += scale * x;
                    }
                }
            };
            double W0[G][G], W1[G][G];
#pragma unroll
            for (int a = 0; a < G; ++a)
#pragma unroll
                for (int b = 0; b < G; ++b) {
                    if (rc.visco == 1) {
                        double x = 0, y = 0;
#pragma unroll
                        for (int t = 0; t < G; ++t) { x += Gv[a][t] * K[t][b]; y += Gv1[a][t] * K[t][b]; }
                        W0[a][b] = x; W1[a][b] = y;
                    } else {
                        W0[a][b] = W1[a][b] = rc.inv_k * (F[a][b] - F1[a][b]);
                    }
                }
            viscK(F, W0, jw * rc.th0, f.psi_grad);
            viscK(F1, W1, jw * rc.th1, f.psi_grad);
        }
#pragma unroll
        for (int a = 0; a < G; ++a) {
            f.psi_val[a] = jw * rc.rho_s_k * (v[a] - v1[a]);
            if (a == 1) f.psi_val[a] += jw * rc.grav_rho;     // -(0,-g rho) . psi
            f.phi_val[a] = jw * (rc.delta_rho_k * (d[a] - d1[a]) - rc.delta_rho * (rc.th0 * v[a] + rc.th1 * v1[a]));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Derivative of the non-linear mesh-lifting fluxes with respect to grad d, for the FROZEN A_pre (newtonsolver.py:26-27):
//   D[a][s][b][t] = d P_as / d (grad d)_bt   at the state grad d = Gp (physical gradient)
//   kind 1  laplace "volume_change":  P = (1/J) grad d                         (laplace.py:33-34,45)
//   kind 2  elastic:                  P = J S_linear(d) F^-T, S = lam tr(eps) I + 2 mu eps, eps = sym(grad d F^-1)
//                                                                               (elastic.py:24-25, common.py:56-60,141-145)
template <int G>
TFSI_HD void lift_flux_derivative(int kind, const double (&Gp)[G][G], double a_mu, double a_lam, double (&D)[G][G][G][G])
{
    double F[G][G], Fi[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int b = 0; b < G; ++b) F[a][b] = Gp[a][b] + ((a == b) ? 1.0 : 0.0);
    const double J = inv<G>(F, Fi);
    if (kind == 1) {
        for (int a = 0; a < G; ++a) for (int s = 0; s < G; ++s) for (int b = 0; b < G; ++b) for (int t = 0; t < G; ++t)
            D[a][s][b][t] = ((((a == b) && (s == t)) ? 1.0 : 0.0) - Fi[t][b] * Gp[a][s]) / J;
        return;
    }
    double S[G][G], SFt[G][G], tre = 0;
    for (int a = 0; a < G; ++a) tre += 1.0 - Fi[a][a];
    for (int a = 0; a < G; ++a) for (int b = 0; b < G; ++b)
        S[a][b] = a_mu * (((a == b) ? 2.0 : 0.0) - Fi[a][b] - Fi[b][a]) + ((a == b) ? a_lam * tre : 0.0);
    for (int a = 0; a < G; ++a) for (int s = 0; s < G; ++s) {
        double x = 0;
        for (int c = 0; c < G; ++c) x += S[a][c] * Fi[s][c];
        SFt[a][s] = x;
    }
    for (int b = 0; b < G; ++b) for (int t = 0; t < G; ++t) {
        // delta(grad d) = e_b e_t^T :  delta(grad d F^-1) = F^-1 e_b e_t^T F^-1
        double trdM = 0;
        for (int i = 0; i < G; ++i) trdM += Fi[i][b] * Fi[t][i];
        const double dJ = J * Fi[t][b];
        for (int a = 0; a < G; ++a) for (int s = 0; s < G; ++s) {
            double x = 0;                                   // (dS F^-T)[a][s]
            for (int c = 0; c < G; ++c) {
                const double dS = a_mu * (Fi[a][b] * Fi[t][c] + Fi[c][b] * Fi[t][a]) + ((a == c) ? a_lam * trdM : 0.0);
                x += dS * Fi[s][c];
            }
            D[a][s][b][t] = dJ * SFt[a][s] + J * x - J * SFt[a][t] * Fi[s][b];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Derivative of the viscoelastic fluxes of the solid (solid.py:46-59, Svisc_D of common.py:129-139) with respect to the
// gradient of the trial field, D[a][s][b][t] = d Phi_as / d (grad u)_bt, Phi the tensor contracted with grad psi:
//   kind 3  "Form1", non-linear, u = d :  Phi = theta0 F(d) Sv(grad v)          -> theta0 d_ab Sv_ts
//   kind 4  "Form1", non-linear, u = v :  Phi = theta0 F(d) Sv(grad v)          -> theta0 F dSv
//   kind 5  "Form2", non-linear, u = d :  Phi = theta0/k F(d) Sv(grad(d - d1))  -> theta0/k (d_ab Sv_ts + F dSv)
//   kind 6  "Form2", "linear",   u = d :  Phi = theta1/k F(d1) Sv(grad(d - d1)) -> theta1/k F(d1) dSv   (frozen in A_pre)
// with (F dSv)[a][s][b][t] = mu_visc d_bt F_as + lambda_visc (F_ab d_st + F_at d_sb).
// Fm = F(d) (kinds 3-5) or F(d1) (kind 6); Gw = grad v (3) or grad(d - d1) (5), unused otherwise.
constexpr int GEN_VISCO_F1_DD = 3, GEN_VISCO_F1_DV = 4, GEN_VISCO_F2_DD = 5, GEN_VISCO_F2_LIN = 6;
template <int G>
TFSI_HD void visco_flux_derivative(int kind, const RegionConst& rc, const double (&Fm)[G][G], const double (&Gw)[G][G],
                                   double (&D)[G][G][G][G])
{
    const double scale = (kind == GEN_VISCO_F1_DD || kind == GEN_VISCO_F1_DV) ? rc.th0
                       : (kind == GEN_VISCO_F2_DD ? rc.th0 * rc.inv_k : rc.th1 * rc.inv_k);
    const double rate = (kind == GEN_VISCO_F2_DD) ? rc.inv_k : 1.0;     // Sv of (d - d1)/k
    double Sv[G][G], tr = 0;
    for (int a = 0; a < G; ++a) tr += Gw[a][a];
    for (int a = 0; a < G; ++a) for (int b = 0; b < G; ++b)
        Sv[a][b] = rate * (rc.lam_visc * (Gw[a][b] + Gw[b][a]) + ((a == b) ? rc.mu_visc * tr : 0.0));
    const double th = (kind == GEN_VISCO_F2_DD) ? rc.th0 : scale;
    for (int a = 0; a < G; ++a) for (int s = 0; s < G; ++s) for (int b = 0; b < G; ++b) for (int t = 0; t < G; ++t) {
        double x = 0.0;
        if (kind == GEN_VISCO_F1_DD || kind == GEN_VISCO_F2_DD) x += (a == b) ? th * Sv[t][s] : 0.0;
        if (kind != GEN_VISCO_F1_DD)
            x += scale * (((b == t) ? rc.mu_visc * Fm[a][s] : 0.0)
                          + rc.lam_visc * (((s == t) ? Fm[a][b] : 0.0) + ((s == b) ? Fm[a][t] : 0.0)));
        D[a][s][b][t] = x;
    }
}

// ---------------------------------------------------------------------------------------------
// host helpers shared by the ABI and the test harness
// P2 / P1 tabulation on the UFC reference simplex (vertices then edge midpoints, UFC edge order)
template <int MAXQ>
inline int tabulate(int g, int nq, const double* pts, const double* wts, QuadTab<MAXQ>& tab) {
    if (nq < 1 || nq > MAXQ) return -1;
    static const int E2[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    static const int E3[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
    memset(&tab, 0, sizeof tab);
    tab.nq = nq;
    const int ne = (g == 2) ? 3 : 6;
    for (int q = 0; q < nq; ++q) {
        double lam[4], dl[4][3];
        double s = 0;
        for (int c = 0; c < g; ++c) s += pts[q * g + c];
        lam[0] = 1.0 - s;
        for (int c = 0; c < g; ++c) {
            lam[c + 1] = pts[q * g + c];
            dl[0][c] = -1.0;
            for (int i = 1; i <= g; ++i) dl[i][c] = (i == c + 1) ? 1.0 : 0.0;
        }
        tab.w[q] = wts[q];
        for (int i = 0; i <= g; ++i) {
            tab.phi1[q][i] = lam[i];
            tab.d4[q][i] = 4.0 * lam[i] - 1.0;
            tab.l4[q][i] = 4.0 * lam[i];
            tab.phi2[q][i] = lam[i] * (2.0 * lam[i] - 1.0);
            for (int c = 0; c < g; ++c) tab.dphi2[q][i][c] = (4.0 * lam[i] - 1.0) * dl[i][c];
        }
        for (int e = 0; e < ne; ++e) {
            const int i = (g == 2) ? E2[e][0] : E3[e][0], j = (g == 2) ? E2[e][1] : E3[e][1];
            tab.phi2[q][g + 1 + e] = 4.0 * lam[i] * lam[j];
            for (int c = 0; c < g; ++c)
                tab.dphi2[q][g + 1 + e][c] = 4.0 * (lam[i] * dl[j][c] + lam[j] * dl[i][c]);
        }
    }
    return 0;
}


inline RegionConst make_region_const(int kind, double rho, double mu, double lambda, double dt, double theta,
                                     int fluid_on, int solid_on, int ext, int ext_sub, double hmin, double gravity)
{
    RegionConst rc;
    memset(&rc, 0, sizeof rc);
    const double k = dt, th0 = theta, th1 = 1.0 - theta, delta = 1.0e7;   // solid.py:29
    rc.kind = kind; rc.th0 = th0; rc.th1 = th1; rc.mu = mu; rc.lam = lambda;
    rc.fluid_on = fluid_on; rc.solid_on = solid_on;
    if (kind == 0) {
        rc.c1 = rho * th0 / k; rc.c1lin = rho * th1 / k; rc.rk = rho / k;
        rc.th0rho = th0 * rho; rc.th1rho = th1 * rho;
        rc.ext = ext; rc.ext_sub = ext_sub;
        if (ext == 1) rc.alfa = (ext_sub == 1) ? 0.01 * hmin * hmin : 1.0;    // laplace.py:37-40
        else if (ext == 2) rc.alfa = 0.01;                                      // biharmonic.py:31
        else if (ext == 3) rc.alfa = 1.0;                                       // elastic.py:18 (E = alfa / CellVolume)
    } else {
        rc.rho_s_k = rho / k; rc.delta_rho_k = delta * rho / k; rc.delta_rho = delta * rho;
        rc.grav_rho = gravity * rho;
    }
    rc.inv_k = 1.0 / k;
    return rc;
}

// material model (MAT_*) and viscoelasticity of a solid region (common.py:104-139, solid.py:46-59)
inline void set_solid_model(RegionConst& rc, int material, double C01, double C10, double C11, double Jm,
                            int visco, double mu_visc, double lam_visc)
{
    rc.mat = material; rc.mC01 = C01; rc.mC10 = C10; rc.mC11 = C11; rc.mJm = Jm;
    rc.visco = visco; rc.mu_visc = mu_visc; rc.lam_visc = lam_visc;
}

}  // namespace tfsi
