/* ORACLE / CPU BASELINE (test infrastructure only — never linked or loaded by the product).
 *
 * Plain-C restatement ("port") of what the reference executes on the CPU for the hot path:
 * an FFC-style dense element tensor per cell (tabulate_tensor: loop over quadrature points,
 * then over test/trial basis functions) followed by DOLFIN/PETSc-style insertion
 * (MatSetValuesLocal(ADD_VALUES): binary search of each column in the sorted CSR row).
 *   forms      turtleFSI/modules/fluid.py:32-56, solid.py:41-63, laplace.py:45
 *   derivative turtleFSI/modules/newtonsolver.py:22-24 (hand-derived Gateaux derivative,
 *              written with gradients on the DEFORMED cell  g_i = F^-T grad(phi_i) — a different
 *              formulation from the CUDA kernels' reference-gradient/C' tensors)
 *   assembly   turtleFSI/modules/newtonsolver.py:75-79 (J_nonlinear), :86-88 (-F)
 * Third-party origin: FFC/DOLFIN/PETSc 2019.1 (un-vendored); validated against the numpy
 * complex-step oracle (oracle/forms.py) in tests/test_oracle_c.py.  PARITY vs DOLFIN: UNPINNED
 * at entry level (see oracle/fem_tables.py header).
 * Build: oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp). Lifting: none | laplace/constant | biharmonic cell terms.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXQ 64

typedef struct {
    int32_t marker, kind; /* 0 fluid, 1 solid */
    double rho, mu, lambda;
} oracle_region;

typedef struct {
    int g, n1, n2, nq;
    double w[MAXQ], phi2[MAXQ][10], dphi2[MAXQ][10][3], phi1[MAXQ][4];
} tab_t;

static void tabulate(int g, int nq, const double* pts, const double* wts, tab_t* t)
{
    static const int E2[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    static const int E3[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
    memset(t, 0, sizeof *t);
    t->g = g; t->n1 = g + 1; t->n2 = g == 2 ? 6 : 10; t->nq = nq;
    int ne = g == 2 ? 3 : 6;
    for (int q = 0; q < nq; ++q) {
        double lam[4], dl[4][3] = {{0}};
        double s = 0;
        for (int c = 0; c < g; ++c) s += pts[q * g + c];
        lam[0] = 1 - s;
        for (int c = 0; c < g; ++c) { lam[c + 1] = pts[q * g + c]; dl[0][c] = -1; dl[c + 1][c] = 1; }
        t->w[q] = wts[q];
        for (int i = 0; i <= g; ++i) {
            t->phi1[q][i] = lam[i];
            t->phi2[q][i] = lam[i] * (2 * lam[i] - 1);
            for (int c = 0; c < g; ++c) t->dphi2[q][i][c] = (4 * lam[i] - 1) * dl[i][c];
        }
        for (int e = 0; e < ne; ++e) {
            int i = g == 2 ? E2[e][0] : E3[e][0], j = g == 2 ? E2[e][1] : E3[e][1];
            t->phi2[q][g + 1 + e] = 4 * lam[i] * lam[j];
            for (int c = 0; c < g; ++c) t->dphi2[q][g + 1 + e][c] = 4 * (lam[i] * dl[j][c] + lam[j] * dl[i][c]);
        }
    }
}

static double inv_g(int g, double F[3][3], double Fi[3][3])
{
    if (g == 2) {
        double d = F[0][0] * F[1][1] - F[0][1] * F[1][0];
        Fi[0][0] = F[1][1] / d; Fi[0][1] = -F[0][1] / d; Fi[1][0] = -F[1][0] / d; Fi[1][1] = F[0][0] / d;
        return d;
    }
    double c00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    double c01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    double c02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    double d = F[0][0] * c00 + F[0][1] * c01 + F[0][2] * c02;
    Fi[0][0] = c00 / d; Fi[1][0] = c01 / d; Fi[2][0] = c02 / d;
    Fi[0][1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) / d;
    Fi[0][2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) / d;
    Fi[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) / d;
    Fi[1][2] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) / d;
    Fi[2][1] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) / d;
    Fi[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) / d;
    return d;
}

/* geometry: K = J_geo^-1, returns |det J_geo| */
static double geometry(int g, const double* coords, const int32_t* cv, double K[3][3])
{
    double Jg[3][3];
    for (int b = 0; b < g; ++b)
        for (int t = 0; t < g; ++t) Jg[b][t] = coords[(int64_t)cv[t + 1] * g + b] - coords[(int64_t)cv[0] * g + b];
    return fabs(inv_g(g, Jg, K));
}

/* value and physical gradient (undeformed cell) of a P2 vector field with local coefficients c[a*n2+j] */
static void eval_vec(const tab_t* t, int q, const double* c, double G[10][3], double val[3], double grad[3][3])
{
    int g = t->g, n2 = t->n2;
    for (int a = 0; a < g; ++a) {
        val[a] = 0;
        for (int b = 0; b < g; ++b) grad[a][b] = 0;
        for (int j = 0; j < n2; ++j) {
            val[a] += c[a * n2 + j] * t->phi2[q][j];
            for (int b = 0; b < g; ++b) grad[a][b] += c[a * n2 + j] * G[j][b];
        }
    }
}

/* element tensor of J_nonlinear for one cell: Ae[ndof*ndof] (row-major), zeroed by the caller */
/* part: 0 = J_nonlinear, 1 = J_linear (A_pre: fluid.py:33, laplace.py:45 alfa = 1, solid.py:41-43), 2 = both */
static void jac_cell(const tab_t* t, const oracle_region* rg, double dt, double theta, const double* coords,
                     const int32_t* cv, const double* u, const double* u1, int ndof, double* Ae, int part, int ext)
{
    const int g = t->g, n1 = t->n1, n2 = t->n2;
    const int ov = g * n2, op = 2 * g * n2, ow = 2 * g * n2 + n1;
    const double au = 0.01; /* biharmonic.py:31 */
    double K[3][3];
    const double adet = geometry(g, coords, cv, K);
    const double th0 = theta, k = dt, rho = rg->rho, mu = rg->mu, lam = rg->lambda;
    for (int q = 0; q < t->nq; ++q) {
        double G[10][3]; /* physical gradients on the undeformed cell */
        for (int j = 0; j < n2; ++j)
            for (int b = 0; b < g; ++b) {
                double s = 0;
                for (int c = 0; c < g; ++c) s += K[c][b] * t->dphi2[q][j][c];
                G[j][b] = s;
            }
        double d[3], v[3], d1[3], v1[3], Gd[3][3], Gv[3][3], Gd1[3][3], Gv1[3][3];
        eval_vec(t, q, u, G, d, Gd);
        eval_vec(t, q, u + ov, G, v, Gv);
        eval_vec(t, q, u1, G, d1, Gd1);
        eval_vec(t, q, u1 + ov, G, v1, Gv1);
        double F[3][3], Fi[3][3];
        for (int a = 0; a < g; ++a) for (int b = 0; b < g; ++b) F[a][b] = (a == b) + Gd[a][b];
        const double jw = t->w[q] * adet;
        if (part >= 1) {
            const double th1 = 1 - theta, delta = 1.0e7;
            double F1[3][3], F1i[3][3];
            for (int a = 0; a < g; ++a) for (int b = 0; b < g; ++b) F1[a][b] = (a == b) + Gd1[a][b];
            const double J1 = rg->kind == 0 ? inv_g(g, F1, F1i) : 1.0;
            for (int a = 0; a < g; ++a)
                for (int i = 0; i < n2; ++i)
                    for (int j = 0; j < n2; ++j) {
                        const double mm = jw * t->phi2[q][i] * t->phi2[q][j];
                        double kk = 0;
                        for (int c = 0; c < g; ++c) kk += G[i][c] * G[j][c];
                        if (rg->kind == 0) {
                            Ae[(size_t)(ov + a * n2 + i) * ndof + ov + a * n2 + j] += rho / k * th1 * J1 * mm;
                            if (ext == 1) Ae[(size_t)(a * n2 + i) * ndof + a * n2 + j] += jw * kk;
                            if (ext == 2) { /* biharmonic.py:35-36 */
                                Ae[(size_t)(ow + a * n2 + i) * ndof + ow + a * n2 + j] += au * mm;
                                Ae[(size_t)(ow + a * n2 + i) * ndof + a * n2 + j] -= au * jw * kk;
                                Ae[(size_t)(a * n2 + i) * ndof + ow + a * n2 + j] += au * jw * kk;
                            }
                        } else {
                            Ae[(size_t)(ov + a * n2 + i) * ndof + ov + a * n2 + j] += rho / k * mm;
                            Ae[(size_t)(a * n2 + i) * ndof + a * n2 + j] += delta * rho / k * mm;
                            Ae[(size_t)(a * n2 + i) * ndof + ov + a * n2 + j] -= delta * rho * th0 * mm;
                        }
                    }
        }
        if (part == 1) continue;
        if (rg->kind == 0) {
            const double J = inv_g(g, F, Fi), Jw = J * jw;
            double p = 0;
            for (int m = 0; m < n1; ++m) p += u[op + m] * t->phi1[q][m];
            double H[3][3], Sig[3][3], uu[3], Hu[3], trH = 0;
            for (int a = 0; a < g; ++a)
                for (int b = 0; b < g; ++b) {
                    double s = 0;
                    for (int c = 0; c < g; ++c) s += Gv[a][c] * Fi[c][b];
                    H[a][b] = s;
                }
            for (int a = 0; a < g; ++a) {
                trH += H[a][a];
                uu[a] = th0 * rho * v[a] - rho / k * (d[a] - d1[a]);
                for (int b = 0; b < g; ++b) Sig[a][b] = mu * (H[a][b] + H[b][a]);
            }
            for (int a = 0; a < g; ++a) { Hu[a] = 0; for (int c = 0; c < g; ++c) Hu[a] += H[a][c] * uu[c]; }
            const double c1 = rho * th0 / k;
            double gd[10][3], sg[10][3], hg[10][3], gu[10];
            for (int j = 0; j < n2; ++j) {
                for (int b = 0; b < g; ++b) {
                    double s = 0;
                    for (int c = 0; c < g; ++c) s += Fi[c][b] * G[j][c];   /* F^-T G_j */
                    gd[j][b] = s;
                }
                gu[j] = 0;
                for (int b = 0; b < g; ++b) gu[j] += gd[j][b] * uu[b];
                for (int a = 0; a < g; ++a) {
                    double s = 0, h = 0;
                    for (int c = 0; c < g; ++c) { s += Sig[a][c] * gd[j][c]; h += H[c][a] * gd[j][c]; }
                    sg[j][a] = s; hg[j][a] = h;
                }
            }
            for (int a = 0; a < g; ++a)
                for (int i = 0; i < n2; ++i) {
                    double* row = Ae + (size_t)(ov + a * n2 + i) * ndof;
                    const double phi_i = t->phi2[q][i];
                    for (int b = 0; b < g; ++b)
                        for (int j = 0; j < n2; ++j) {
                            const double phi_j = t->phi2[q][j];
                            double gg = 0;
                            for (int c = 0; c < g; ++c) gg += gd[i][c] * gd[j][c];
                            /* d/dv_b,j */
                            double e = th0 * rho * H[a][b] * phi_j * phi_i + th0 * mu * gd[j][a] * gd[i][b];
                            if (a == b) e += c1 * phi_j * phi_i + gu[j] * phi_i + th0 * mu * gg;
                            row[ov + b * n2 + j] += Jw * e;
                            /* d/dd_b,j */
                            double f = phi_i * (c1 * (v[a] - v1[a]) * gd[j][b] + gd[j][b] * Hu[a] - H[a][b] * gu[j] - rho / k * phi_j * H[a][b]);
                            f -= p * (gd[j][b] * gd[i][a] - gd[j][a] * gd[i][b]);
                            f += th0 * (gd[j][b] * sg[i][a] - mu * H[a][b] * gg - mu * gd[j][a] * hg[i][b] - sg[j][a] * gd[i][b]);
                            row[b * n2 + j] += Jw * f;
                        }
                    for (int m = 0; m < n1; ++m) row[op + m] -= Jw * t->phi1[q][m] * gd[i][a];
                }
            for (int m = 0; m < n1; ++m) {
                double* row = Ae + (size_t)(op + m) * ndof;
                for (int b = 0; b < g; ++b)
                    for (int j = 0; j < n2; ++j) {
                        row[ov + b * n2 + j] += Jw * t->phi1[q][m] * gd[j][b];
                        row[b * n2 + j] += Jw * t->phi1[q][m] * (gd[j][b] * trH - hg[j][b]);
                    }
            }
        } else {
            /* theta0 dP : grad psi, P = F (2 mu E + lam tr E I) */
            double S[3][3], trE = 0, FFt[3][3];
            for (int a = 0; a < g; ++a)
                for (int b = 0; b < g; ++b) {
                    double s = 0, f = 0;
                    for (int c = 0; c < g; ++c) { s += F[c][a] * F[c][b]; f += F[a][c] * F[b][c]; }
                    double E = 0.5 * (s - (a == b));
                    S[a][b] = 2 * mu * E;
                    FFt[a][b] = f;
                    if (a == b) trE += E;
                }
            for (int a = 0; a < g; ++a) S[a][a] += lam * trE;
            double fG[10][3], SG[10][3];
            for (int j = 0; j < n2; ++j)
                for (int a = 0; a < g; ++a) {
                    double s = 0, x = 0;
                    for (int c = 0; c < g; ++c) { s += F[a][c] * G[j][c]; x += S[a][c] * G[j][c]; }
                    fG[j][a] = s; SG[j][a] = x;
                }
            for (int a = 0; a < g; ++a)
                for (int i = 0; i < n2; ++i) {
                    double* row = Ae + (size_t)(ov + a * n2 + i) * ndof;
                    for (int b = 0; b < g; ++b)
                        for (int j = 0; j < n2; ++j) {
                            double GG = 0, GSG = 0;
                            for (int c = 0; c < g; ++c) { GG += G[i][c] * G[j][c]; GSG += G[j][c] * SG[i][c]; }
                            double e = mu * fG[j][a] * fG[i][b] + mu * FFt[a][b] * GG + lam * fG[i][a] * fG[j][b];
                            if (a == b) e += GSG;
                            row[b * n2 + j] += th0 * jw * e;
                        }
                }
        }
    }
}

/* element vector of F = F_lin + F_nonlin (laplace/constant lifting) */
static void res_cell(const tab_t* t, const oracle_region* rg, double dt, double theta, const double* coords,
                     const int32_t* cv, const double* u, const double* u1, double* Fe, int ext, double grav)
{
    const int g = t->g, n1 = t->n1, n2 = t->n2;
    const int ov = g * n2, op = 2 * g * n2, ow = 2 * g * n2 + n1;
    const double au = 0.01;
    double K[3][3];
    const double adet = geometry(g, coords, cv, K);
    const double th0 = theta, th1 = 1 - theta, k = dt, rho = rg->rho, mu = rg->mu, lam = rg->lambda;
    for (int q = 0; q < t->nq; ++q) {
        double G[10][3];
        for (int j = 0; j < n2; ++j)
            for (int b = 0; b < g; ++b) {
                double s = 0;
                for (int c = 0; c < g; ++c) s += K[c][b] * t->dphi2[q][j][c];
                G[j][b] = s;
            }
        double d[3], v[3], d1[3], v1[3], Gd[3][3], Gv[3][3], Gd1[3][3], Gv1[3][3];
        eval_vec(t, q, u, G, d, Gd);
        eval_vec(t, q, u + ov, G, v, Gv);
        eval_vec(t, q, u1, G, d1, Gd1);
        eval_vec(t, q, u1 + ov, G, v1, Gv1);
        double F[3][3], F1[3][3];
        for (int a = 0; a < g; ++a) for (int b = 0; b < g; ++b) { F[a][b] = (a == b) + Gd[a][b]; F1[a][b] = (a == b) + Gd1[a][b]; }
        const double jw = t->w[q] * adet;
        double fval[3] = {0, 0, 0}, flux[3][3] = {{0}}, dval[3] = {0, 0, 0}, dflux[3][3] = {{0}}, gam = 0;
        double wval[3] = {0, 0, 0}, wflux[3][3] = {{0}};
        if (rg->kind == 0) {
            double Fi[3][3], Fi1[3][3];
            const double J = inv_g(g, F, Fi), J1 = inv_g(g, F1, Fi1);
            double p = 0;
            for (int m = 0; m < n1; ++m) p += u[op + m] * t->phi1[q][m];
            double H[3][3], H1[3][3];
            for (int a = 0; a < g; ++a)
                for (int b = 0; b < g; ++b) {
                    double s = 0, s1 = 0;
                    for (int c = 0; c < g; ++c) { s += Gv[a][c] * Fi[c][b]; s1 += Gv1[a][c] * Fi1[c][b]; }
                    H[a][b] = s; H1[a][b] = s1;
                }
            for (int a = 0; a < g; ++a) {
                double conv = 0, conv1 = 0, ale = 0;
                for (int c = 0; c < g; ++c) { conv += H[a][c] * v[c]; conv1 += H1[a][c] * v1[c]; ale += H[a][c] * (d[c] - d1[c]); }
                fval[a] = rho / k * (th0 * J + th1 * J1) * (v[a] - v1[a]) + th0 * rho * J * conv + th1 * rho * J1 * conv1 - rho / k * J * ale;
                for (int b = 0; b < g; ++b) {
                    /* [J (-p I + th0 sigma) F^-T + th1 J1 sigma1 F1^-T]_ab */
                    double s = 0;
                    for (int c = 0; c < g; ++c)
                        s += J * (th0 * mu * (H[a][c] + H[c][a]) - (a == c ? p : 0.0)) * Fi[b][c] + th1 * J1 * mu * (H1[a][c] + H1[c][a]) * Fi1[b][c];
                    flux[a][b] = s;
                }
                gam += J * H[a][a];
                if (ext == 1) for (int b = 0; b < g; ++b) dflux[a][b] = Gd[a][b];   /* laplace.py:45, alfa = 1 */
            }
            if (ext == 2) { /* biharmonic.py:35-36 */
                double w[3], Gw[3][3];
                eval_vec(t, q, u + ow, G, w, Gw);
                for (int a = 0; a < g; ++a) {
                    wval[a] = au * w[a];
                    for (int b = 0; b < g; ++b) { wflux[a][b] = -au * Gd[a][b]; dflux[a][b] = au * Gw[a][b]; }
                }
            }
        } else {
            for (int pass = 0; pass < 2; ++pass) {
                double (*Fm)[3] = pass ? F1 : F;
                const double sc = pass ? th1 : th0;
                double S[3][3], trE = 0;
                for (int a = 0; a < g; ++a)
                    for (int b = 0; b < g; ++b) {
                        double s = 0;
                        for (int c = 0; c < g; ++c) s += Fm[c][a] * Fm[c][b];
                        double E = 0.5 * (s - (a == b));
                        S[a][b] = 2 * mu * E;
                        if (a == b) trE += E;
                    }
                for (int a = 0; a < g; ++a) S[a][a] += lam * trE;
                for (int a = 0; a < g; ++a)
                    for (int b = 0; b < g; ++b) {
                        double s = 0;
                        for (int c = 0; c < g; ++c) s += Fm[a][c] * S[c][b];
                        flux[a][b] += sc * s;
                    }
            }
            const double delta = 1.0e7;
            for (int a = 0; a < g; ++a) {
                fval[a] = rho / k * (v[a] - v1[a]) + (a == 1 ? grav * rho : 0.0);   /* solid.py:65-68 */
                dval[a] = delta * rho / k * (d[a] - d1[a]) - delta * rho * (th0 * v[a] + th1 * v1[a]);
            }
        }
        for (int a = 0; a < g; ++a)
            for (int i = 0; i < n2; ++i) {
                double s = fval[a] * t->phi2[q][i], sd = dval[a] * t->phi2[q][i];
                for (int b = 0; b < g; ++b) { s += flux[a][b] * G[i][b]; sd += dflux[a][b] * G[i][b]; }
                Fe[ov + a * n2 + i] += jw * s;
                Fe[a * n2 + i] += jw * sd;
                if (ext == 2) {
                    double sw = wval[a] * t->phi2[q][i];
                    for (int b = 0; b < g; ++b) sw += wflux[a][b] * G[i][b];
                    Fe[ow + a * n2 + i] += jw * sw;
                }
            }
        for (int m = 0; m < n1; ++m) Fe[op + m] += jw * gam * t->phi1[q][m];
    }
}

static const oracle_region* find_region(int nreg, const oracle_region* regs, int marker)
{
    for (int i = 0; i < nreg; ++i) if (regs[i].marker == marker) return regs + i;
    return NULL;
}

/* A_vals[nnz] += J_nonlinear (cells [c0, c1)); insertion by binary search in the sorted row */
int tfsi_oracle_assemble_jacobian(int g, int nq, const double* pts, const double* wts, int64_t c0, int64_t c1,
                                  const double* coords, const int32_t* cells, const int32_t* cell_dofs,
                                  const int32_t* marker, int ndof, int nreg, const oracle_region* regs, double dt,
                                  double theta, const double* x, const double* x1, const int64_t* rowptr,
                                  const int32_t* colidx, double* A_vals, int nthreads, int part, int ext)
{
    tab_t tab;
    if (nq > MAXQ) return -1;
    tabulate(g, nq, pts, wts, &tab);
    const int nloc = ndof;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double* Ae = (double*)malloc(sizeof(double) * ndof * ndof);
        double u[128], u1[128];
#pragma omp for schedule(dynamic, 64)
        for (int64_t c = c0; c < c1; ++c) {
            const oracle_region* rg = find_region(nreg, regs, marker[c]);
            if (!rg) continue;
            const int32_t* dofs = cell_dofs + c * ndof;
            for (int i = 0; i < nloc; ++i) { u[i] = x[dofs[i]]; u1[i] = x1[dofs[i]]; }
            memset(Ae, 0, sizeof(double) * ndof * ndof);
            jac_cell(&tab, rg, dt, theta, coords, cells + c * (g + 1), u, u1, ndof, Ae, part, ext);
            for (int i = 0; i < ndof; ++i) {
                const int64_t lo = rowptr[dofs[i]], hi = rowptr[dofs[i] + 1];
                for (int j = 0; j < ndof; ++j) {
                    const double val = Ae[(size_t)i * ndof + j];
                    if (val == 0.0) continue;
                    int64_t a = lo, b = hi;
                    while (a < b) { int64_t m = (a + b) >> 1; if (colidx[m] < dofs[j]) a = m + 1; else b = m; }
                    if (a >= hi || colidx[a] != dofs[j]) { bad = 1; continue; }
#pragma omp atomic
                    A_vals[a] += val;
                }
            }
        }
        free(Ae);
    }
    return bad ? -2 : 0;
}

/* b[N] -= F (cells [c0, c1)) */
int tfsi_oracle_assemble_residual(int g, int nq, const double* pts, const double* wts, int64_t c0, int64_t c1,
                                  const double* coords, const int32_t* cells, const int32_t* cell_dofs,
                                  const int32_t* marker, int ndof, int nreg, const oracle_region* regs, double dt,
                                  double theta, const double* x, const double* x1, double* b, int nthreads, int ext, double grav)
{
    tab_t tab;
    if (nq > MAXQ) return -1;
    tabulate(g, nq, pts, wts, &tab);
    const int nloc = ndof;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double u[128], u1[128], Fe[128];
#pragma omp for schedule(dynamic, 64)
        for (int64_t c = c0; c < c1; ++c) {
            const oracle_region* rg = find_region(nreg, regs, marker[c]);
            if (!rg) continue;
            const int32_t* dofs = cell_dofs + c * ndof;
            for (int i = 0; i < nloc; ++i) { u[i] = x[dofs[i]]; u1[i] = x1[dofs[i]]; }
            memset(Fe, 0, sizeof Fe);
            res_cell(&tab, rg, dt, theta, coords, cells + c * (g + 1), u, u1, Fe, ext, grav);
            for (int i = 0; i < nloc; ++i) {
#pragma omp atomic
                b[dofs[i]] -= Fe[i];
            }
        }
    }
    return 0;
}

int tfsi_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
