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
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define TFSI_HD __host__ __device__ __forceinline__
#else
#define TFSI_HD inline
#endif

namespace tfsi {

constexpr int MAXQ_J = 16;   // Jacobian rule: <= 15 points (degree <= 5 in 3-D, <= 6 in 2-D)
constexpr int MAXQ_R = 64;   // residual rule: 24 (3-D degree 6), 36 (2-D degree 10), 64 (3-D degree 7)

template <int MAXQ>
struct QuadTab {
    int nq;
    int pad_;
    double w[MAXQ];
    double phi2[MAXQ][10];
    double dphi2[MAXQ][10][3];
    double phi1[MAXQ][4];
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
    int ext;               // 0 none, 1 laplace, 2 biharmonic
    int ext_sub;           // laplace: 0 constant, 1 small_constant, 2 volume
    int pad_;
};

// ---------------------------------------------------------------------------------------------
// FLUID, non-linear Jacobian: per-point base items (stored lane-contiguous in shared memory)
template <int G> struct FluidBase {
    static constexpr int JW = 0;                    // J w |detJgeo|
    static constexpr int B = 1;                     // B[t][c]       G*G
    static constexpr int H = B + G * G;             // H = grad v F^-1
    static constexpr int M = H + G * G;             // M = B B^T     packed sym
    static constexpr int BU = M + Dim<G>::NS;       // B u,  u = th0 rho v - rho/k (d - d1)
    static constexpr int XA = BU + G;               // c1 (v - v1) + H u
    static constexpr int P = XA + G;                // pressure
    static constexpr int S = P + 1;                 // S = B sigma,  sigma = mu (H + H^T)
    static constexpr int R = S + G * G;             // R = B H
    static constexpr int COUNT = R + G * G;         // 50 (3-D), 25 (2-D)
};

// coefficient access: cf(k) returns local dof k of the cell at time n, cf1(k) at n-1
// (local layout [d comps | v comps | p | w comps]); out(item, value) stores a base item.
template <int G, class Tab, class CF, class CF1, class OUT>
TFSI_HD void fluid_base(const Tab& tab, int q, const double (&K)[G][G], double adet,
                        const RegionConst& rc, CF cf, CF1 cf1, OUT out)
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

    out(I::JW, J * tab.w[q] * adet);
    out(I::P, p);
#pragma unroll
    for (int t = 0; t < G; ++t) {
        double bu = 0, hu = 0;
#pragma unroll
        for (int c = 0; c < G; ++c) { bu += Bm[t][c] * u[c]; hu += Hm[t][c] * u[c]; }
        out(I::BU + t, bu);
        out(I::XA + t, rc.c1 * dv[t] + hu);
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
            out(I::R + t * G + c, r);
            out(I::S + t * G + c, rc.mu * s);
            if (c >= t) {
                double m = 0;
#pragma unroll
                for (int e = 0; e < G; ++e) m += Bm[t][e] * Bm[c][e];
                out(I::M + symi(t, c, G), m);
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
    C[0][0] = Jw * (rc.c1 * dab + rc.th0rho * bv(I::H + a * G + b));
    const double k4 = Jw * rc.th0 * rc.mu;
#pragma unroll
    for (int t = 0; t < G; ++t) {
        C[0][1 + t] = dab * Jw * bv(I::BU + t);
        C[1 + t][0] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < G; ++s)
#pragma unroll
        for (int t = 0; t < G; ++t)
            C[1 + s][1 + t] = k4 * (dab * bv(I::M + symi(s, t, G)) + bv(I::B + t * G + a) * bv(I::B + s * G + b));
}

// C' of the block  psi_a x delta d_b  (shape derivatives of N1..N4, N6)
template <int G, class BV>
TFSI_HD void fluid_C_psi_dd(BV bv, const RegionConst& rc, int a, int b, double (&C)[G + 1][G + 1])
{
    typedef FluidBase<G> I;
    const double Jw = bv(I::JW);
    const double Hab = bv(I::H + a * G + b);
    const double Xa = bv(I::XA + a);
    const double p = bv(I::P);
    C[0][0] = -Jw * rc.rk * Hab;
#pragma unroll
    for (int t = 0; t < G; ++t) {
        C[0][1 + t] = Jw * (Xa * bv(I::B + t * G + b) - Hab * bv(I::BU + t));
        C[1 + t][0] = 0.0;
    }
    const double muH = rc.mu * Hab;
#pragma unroll
    for (int s = 0; s < G; ++s) {
        const double Bsa = bv(I::B + s * G + a), Bsb = bv(I::B + s * G + b);
        const double Ssa = bv(I::S + s * G + a), Rsb = bv(I::R + s * G + b);
#pragma unroll
        for (int t = 0; t < G; ++t) {
            const double Bta = bv(I::B + t * G + a), Btb = bv(I::B + t * G + b);
            const double Sta = bv(I::S + t * G + a);
            double n3 = -p * (Bsa * Btb - Bsb * Bta);
            double n4 = Ssa * Btb - Sta * Bsb - muH * bv(I::M + symi(s, t, G)) - rc.mu * Bta * Rsb;
            C[1 + s][1 + t] = Jw * (n3 + rc.th0 * n4);
        }
    }
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
    double trH = 0;
#pragma unroll
    for (int e = 0; e < G; ++e) trH += bv(I::H + e * G + e);
#pragma unroll
    for (int t = 0; t < G; ++t) c[t] = Jw * (bv(I::B + t * G + b) * trH - bv(I::R + t * G + b));
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
// Sum-factorised block updates.  MODE bits: 1 = C[0][0] present, 2 = C[0][t] present,
// 4 = C[s][t] present (C[s][0] is structurally zero in every form of this path).
template <int G, int J0, int NJ, int MODE, class Tab>
TFSI_HD void block_p2p2(const Tab& tab, int q, const double (&C)[G + 1][G + 1],
                        double (&Z)[Dim<G>::N2][NJ])
{
    constexpr int N2 = Dim<G>::N2;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const int j = J0 + jj;
        double W[G + 1];
        W[0] = 0.0;
        if (MODE & 1) W[0] = C[0][0] * tab.phi2[q][j];
        if (MODE & 2) {
#pragma unroll
            for (int t = 0; t < G; ++t) W[0] += C[0][1 + t] * tab.dphi2[q][j][t];
        }
#pragma unroll
        for (int s = 0; s < G; ++s) {
            double x = 0.0;
            if (MODE & 4) {
#pragma unroll
                for (int t = 0; t < G; ++t) x += C[1 + s][1 + t] * tab.dphi2[q][j][t];
            }
            W[1 + s] = x;
        }
#pragma unroll
        for (int i = 0; i < N2; ++i) {
            double z = Z[i][jj];
            if (MODE & 3) z += tab.phi2[q][i] * W[0];
            if (MODE & 4) {
#pragma unroll
                for (int s = 0; s < G; ++s) z += tab.dphi2[q][i][s] * W[1 + s];
            }
            Z[i][jj] = z;
        }
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

template <int G, class Tab, class CF, class CF1>
TFSI_HD void residual_point(const Tab& tab, int q, const double (&K)[G][G], double adet,
                            const RegionConst& rc, double alfa_cell, CF cf, CF1 cf1,
                            PointFlux<G>& f)
{
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    const double jw = tab.w[q] * adet;
    // field values and reference gradients at n and n-1
    double d[G], v[G], d1[G], v1[G], Gd[G][G], Gv[G][G], Gd1[G][G], Gv1[G][G];
#pragma unroll
    for (int a = 0; a < G; ++a) {
        d[a] = v[a] = d1[a] = v1[a] = 0;
#pragma unroll
        for (int t = 0; t < G; ++t) Gd[a][t] = Gv[a][t] = Gd1[a][t] = Gv1[a][t] = 0;
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            double dj = cf(a * N2 + j), vj = cf(G * N2 + a * N2 + j);
            double d1j = cf1(a * N2 + j), v1j = cf1(G * N2 + a * N2 + j);
            double ph = tab.phi2[q][j];
            d[a] += dj * ph; v[a] += vj * ph; d1[a] += d1j * ph; v1[a] += v1j * ph;
#pragma unroll
            for (int t = 0; t < G; ++t) {
                double g = tab.dphi2[q][j][t];
                Gd[a][t] += dj * g; Gv[a][t] += vj * g; Gd1[a][t] += d1j * g; Gv1[a][t] += v1j * g;
            }
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
                    f.phi_grad[a][s] = jw * alfa_cell * x;
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
        piolaK(F, jw * rc.th0, f.psi_grad);
        piolaK(F1, jw * rc.th1, f.psi_grad);
#pragma unroll
        for (int a = 0; a < G; ++a) {
            f.psi_val[a] = jw * rc.rho_s_k * (v[a] - v1[a]);
            if (a == 1) f.psi_val[a] += jw * rc.grav_rho;     // -(0,-g rho) . psi
            f.phi_val[a] = jw * (rc.delta_rho_k * (d[a] - d1[a]) - rc.delta_rho * (rc.th0 * v[a] + rc.th1 * v1[a]));
        }
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
    } else {
        rc.rho_s_k = rho / k; rc.delta_rho_k = delta * rho / k; rc.delta_rho = delta * rho;
        rc.grav_rho = gravity * rho;
    }
    return rc;
}

}  // namespace tfsi
