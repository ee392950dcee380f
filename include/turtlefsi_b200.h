/* turtlefsi_b200 — C ABI of the B200-native monolithic FSI assembly hot path.
 *
 * Every entry point replaces one DOLFIN/PETSc call made by turtleFSI's Newton solver
 * (reference: turtleFSI/modules/newtonsolver.py) — the line it replaces is cited at each
 * declaration.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Memory: `tfsi_create` copies its inputs (host OR device pointers, resolved through UVA) into
 * handle-owned device memory.  All vectors/values passed to the assemble/apply calls are DEVICE
 * pointers owned by the caller (a torch.Tensor.data_ptr(), a cudaMalloc, ...); the `_host`
 * variants take HOST pointers and include the host<->device copies.  `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).  Calls are asynchronous on `stream` unless
 * they return a host scalar.
 *
 * Return value: 0 on success, negative error code otherwise; `tfsi_last_error` gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry fails with TFSI_ERR_CUDA.
 */
#ifndef TURTLEFSI_B200_H
#define TURTLEFSI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFSI_OK 0
#define TFSI_ERR_ARG (-1)
#define TFSI_ERR_CUDA (-2)
#define TFSI_ERR_PATTERN (-3)
#define TFSI_ERR_UNSUPPORTED (-4)

#define TFSI_KIND_FLUID 0
#define TFSI_KIND_SOLID 1

#define TFSI_EXT_NONE 0        /* modules/no_extrapolation.py */
#define TFSI_EXT_LAPLACE 1     /* modules/laplace.py          */
#define TFSI_EXT_BIHARMONIC 2  /* modules/biharmonic.py (cell terms) */
#define TFSI_EXT_ELASTIC 3     /* modules/elastic.py:18-27: J S_linear(d) F^-T : grad(phi), E = 1/CellVolume, nu = 0.25;
                                  non-linear in d but part of F_fluid_linear: the residual uses the current state,
                                  its Jacobian is FROZEN in A_pre at the set-up state (newtonsolver.py:26-27) */

#define TFSI_LAPLACE_CONSTANT 0        /* laplace.py:39-40 */
#define TFSI_LAPLACE_SMALL_CONSTANT 1  /* laplace.py:37-38 */
#define TFSI_LAPLACE_VOLUME 2          /* laplace.py:35-36 */
#define TFSI_LAPLACE_VOLUME_CHANGE 3   /* laplace.py:33-34: alfa = 1/J(d^n) in the residual; its Jacobian is part of the
                                          FROZEN A_pre (newtonsolver.py:26-27), evaluated at the state handed to
                                          tfsi_assemble_A_pre (d = 0 in every shipped problem, any state supported) */

/* material_model strings of common.py:104-121. StVenantKirchoff is the closed form P = F (2 mu E + lambda tr(E) I);
 * the others are P = dW/dF of the strain-energy densities common.py:214-273 (their tangent is part of J_nonlinear). */
#define TFSI_MAT_SVK 0             /* "StVenantKirchoff"        */
#define TFSI_MAT_SVK_ENERGY 1      /* "StVenantKirchoffEnergy"  */
#define TFSI_MAT_NEO_HOOKEAN 2     /* "NeoHookean"              */
#define TFSI_MAT_MOONEY_RIVLIN 3   /* "MooneyRivlin"            */
#define TFSI_MAT_GENT 4            /* "Gent"                    */

typedef struct tfsi_handle tfsi_handle;

/* one entry per cell-marker id that carries integrals: dx_f[i] / dx_s[i]
 * (turtleFSI/modules/domain.py:54-103) */
typedef struct {
    int32_t marker;   /* value of the `domains` MeshFunction selecting this region     */
    int32_t kind;     /* TFSI_KIND_FLUID | TFSI_KIND_SOLID                              */
    double rho;       /* rho_f | rho_s                                                  */
    double mu;        /* mu_f  | mu_s                                                   */
    double lambda;    /* lambda_s (solid, St.Venant-Kirchhoff: common.py:104-109)       */
    /* solid regions only (zero = the reference's defaults: StVenantKirchoff, no viscoelasticity) */
    int32_t material;         /* TFSI_MAT_*: solid_properties["material_model"] (common.py:104-121)            */
    int32_t viscoelasticity;  /* 0 none, 1 "Form1", 2 "Form2": solid_properties["viscoelasticity"] (solid.py:46-59) */
    double C01, C10, C11;     /* MooneyRivlin constants (common.py:258-273)                                   */
    double Jm;                /* Gent constant (common.py:242-256)                                            */
    double mu_visc;           /* solid_properties["mu_visc_s"]     (Svisc_D, common.py:129-139)               */
    double lambda_visc;       /* solid_properties["lambda_visc_s"]                                            */
} tfsi_region;

typedef struct {
    int32_t gdim;             /* 2 | 3                                                  */
    int32_t biharmonic;       /* 1: DVP has the 4th sub-space w (monolithic.py:86-89)   */
    int64_t num_vertices;
    int64_t num_cells;
    int64_t num_dofs;         /* N                                                      */
    const double* coords;         /* [num_vertices * gdim]                              */
    const int32_t* cell_vertices; /* [num_cells * (gdim+1)], ascending per cell         */
    const int32_t* cell_dofs;     /* [num_cells * ndof], local layout d|v|p(|w); taken
                                     as INPUT so DOLFIN's dofmap is honoured bit-exactly */
    const int32_t* cell_marker;   /* [num_cells]                                        */
    const int64_t* rowptr;        /* [N+1]  CSR pattern, taken as INPUT                 */
    const int32_t* colidx;        /* [nnz]  ascending within a row                      */
    int32_t num_regions;
    const tfsi_region* regions;
    double dt;                /* k      (monolithic.py:82)                              */
    double theta;             /* theta  (fluid.py:18-19, solid.py:32-33)                */
    int32_t fluid_on;         /* 0 = --fluid no_fluid                                   */
    int32_t solid_on;         /* 0 = --solid no_solid                                   */
    int32_t extrapolation;    /* TFSI_EXT_*                                             */
    int32_t extrapolation_sub_type; /* TFSI_LAPLACE_* (laplace only)                    */
    double hmin;              /* mesh.hmin() (laplace small_constant)                   */
    double gravity;           /* solid.py:65-68; used iff has_gravity                   */
    int32_t has_gravity;
    /* quadrature rules on the UFC reference simplex (weights sum to the reference volume).
     * Jacobian: compiler_parameters["quadrature_degree"] (problems/__init__.py:16-17);
     * residual: the global form-compiler degree. */
    int32_t nq_jacobian;
    const double* qpts_jacobian;  /* [nq * gdim] */
    const double* qwts_jacobian;  /* [nq]        */
    int32_t nq_residual;
    const double* qpts_residual;
    const double* qwts_residual;
    int32_t verify_pattern;   /* 1: check that every cell block fits the CSR pattern    */
    int32_t borrow_pattern;   /* 1: rowptr / colidx are DEVICE arrays that stay valid (and unchanged) for the life of
                                 the handle (or until tfsi_release_pattern): no second copy of the 4 B/entry column
                                 array is made                                                                       */
    int32_t deterministic;    /* 1: colour-ordered scatter (north_star design point 3): the cells of every region are
                                 split into colour classes (no two cells of a class share a mesh vertex), one launch
                                 per class in a fixed order => every CSR entry / residual row receives its additions
                                 in the same order in every run: bitwise reproducible results                        */
} tfsi_problem;

/* newtonsolver.py:13-31 (solver_setup): uploads mesh, dofmap, pattern; builds scatter maps. */
int tfsi_create(const tfsi_problem* prob, tfsi_handle** out);
void tfsi_destroy(tfsi_handle* h);
/* message of the last failing call on `h` (or of the last failing tfsi_create if h == NULL) */
const char* tfsi_last_error(const tfsi_handle* h);
int64_t tfsi_nnz(const tfsi_handle* h);
int32_t tfsi_ndof_cell(const tfsi_handle* h);

/* A_pre = assemble(J_linear, keep_diagonal=True)            newtonsolver.py:26-27
 * x, x1: state at n and n-1 [N]; Apre_vals [nnz] is overwritten. */
int tfsi_assemble_A_pre(tfsi_handle* h, const double* x, const double* x1, double* Apre_vals,
                        void* stream);

/* assemble(J_nonlinear, tensor=A, keep_diagonal=True); A.axpy(1.0, A_pre, True)
 *                                                           newtonsolver.py:75-79
 * A_vals [nnz] is overwritten with J_nonlinear(x, x1) + A_pre (A_pre may be NULL = 0). */
int tfsi_assemble_jacobian(tfsi_handle* h, const double* x, const double* x1,
                           const double* Apre_vals, double* A_vals, void* stream);

/* The same in two halves, for callers that overlap the halo exchange of the iterate with the x-independent half:
 * phase 1 = the initialisation A := A_pre only (x, x1 unused, may be NULL), phase 2 = the cell kernels only
 * (A += J_nonlinear(x, x1)). phase 1 followed by phase 2 on one stream == tfsi_assemble_jacobian. */
int tfsi_assemble_jacobian_phase(tfsi_handle* h, const double* x, const double* x1,
                                 const double* Apre_vals, double* A_vals, void* stream, int32_t phase);

/* Fixed LINEAR facet terms of the forms as a sparse operator A_f in COO form (HOST or DEVICE arrays,
 * duplicates are summed): the exterior-facet terms of the biharmonic lifting with
 * "constrained_disp_vel" (turtleFSI/modules/biharmonic.py:38-41) — and any other term a problem file
 * adds to F_*_linear that is linear in dvp_["n"]. Afterwards tfsi_assemble_A_pre adds A_f and
 * tfsi_assemble_residual subtracts A_f x. Entries must lie inside the CSR pattern. n = 0 removes it. */
int tfsi_set_facet_operator(tfsi_handle* h, const int32_t* rows, const int32_t* cols, const double* vals, int64_t n);

/* The same for linear facet terms that act on the PREVIOUS time level dvp_["n-1"], e.g. the explicit
 * half of the Robin condition  theta1 (k_s d^{n-1} + c_s v^{n-1}).psi ds_s   (turtleFSI/modules/solid.py:84;
 * its implicit half :83 goes to tfsi_set_facet_operator). No Jacobian contribution:
 * tfsi_assemble_residual subtracts A_p x1. HOST arrays, indices range-checked. n = 0 removes it. */
int tfsi_set_facet_operator_prev(tfsi_handle* h, const int32_t* rows, const int32_t* cols, const double* vals, int64_t n);

/* State-independent facet load added to F_solid_linear by a problem file, e.g. the interior-facet
 * pressure  p(t) n('+').psi('+') dS(fsi_id)  (problems/Test_Cylinder/problem_aneu.py:212-214,
 * problems/Stenosis_2D.py:172-174): a sparse vector l (HOST arrays dofs/vals, duplicates summed) and a
 * scalar that the caller updates per time step (pre_solve: p_out_bc_val.t = t);
 * tfsi_assemble_residual subtracts scale * l. Default scale 1. n = 0 removes the load. */
int tfsi_set_facet_load(tfsi_handle* h, const int32_t* dofs, const double* vals, int64_t n);
int tfsi_set_facet_load_scale(tfsi_handle* h, double scale);

/* Optional: register a COMPRESSED copy of A_pre (only its non-zero entries; the linear forms fill
 * ~15-20 % of the pattern). While `Apre_vals` passed to tfsi_assemble_jacobian is the pointer
 * registered here, "A := A_pre" streams A once from a position bitmap + the packed non-zeros instead of
 * reading the full A_pre. The
 * caller must call it again (or tfsi_assemble_A_pre, which invalidates it) if it changes A_pre. */
int tfsi_compress_A_pre(tfsi_handle* h, const double* Apre_vals, void* stream);
int64_t tfsi_compressed_nnz(const tfsi_handle* h);
/* Pass this as `Apre_vals` to use the registered compressed copy whatever became of the dense array (the caller
 * may free it: at 6.3 M tets per GPU the dense A_pre would cost 77 GB). A_pre is IMMUTABLE once compressed: after
 * editing the dense values call tfsi_compress_A_pre again or tfsi_invalidate_A_pre (then the dense pointer is read). */
#define TFSI_APRE_COMPRESSED ((const double*)(uintptr_t)1)
int tfsi_invalidate_A_pre(tfsi_handle* h);

/* A = Matrix(A_pre)                                          newtonsolver.py:28
 * Declares A_vals as the solver's PERSISTENT matrix: it is set to A_pre now (needs tfsi_compress_A_pre), and from
 * then on tfsi_assemble_jacobian(..., A_vals) re-initialises only the rows that the non-linear forms scatter into
 * (velocity and pressure rows of active regions: ~half of the entries); all other rows keep what the library left
 * there — A_pre, or the unit rows written by tfsi_ident_zeros / tfsi_apply_bc_matrix, which are idempotent. The caller
 * must not write to A_vals itself while it is bound. tfsi_set_bc_dofs, tfsi_compress_A_pre, tfsi_invalidate_A_pre
 * make the next assembly a full one. */
int tfsi_bind_matrix(tfsi_handle* h, double* A_vals, void* stream);
/* The same with TWO buffers that the caller alternates between re-assemblies (the previous Jacobian is dead once the
 * next one is assembled). While tfsi_assemble_jacobian(..., A_k) runs, the rows of the OTHER buffer that the cell
 * kernels scatter into are re-initialised BY THE FLUID KERNEL ITSELF, between its tasks: zero-fill with TMA bulk stores
 * (shared -> global from a zero block), and the A_pre non-zeros of those runs are added with REDs by the fluid kernel
 * of the next assembly into that buffer. The separate HBM-bound pass A := A_pre of the single-buffer mode (18 % of the
 * step at 2 M tets) disappears into the FP64-bound kernel. `Apre_vals`: the array registered by tfsi_compress_A_pre or
 * TFSI_APRE_COMPRESSED. The caller must be done with the other buffer (stream-ordered on `stream`) when it calls
 * tfsi_assemble_jacobian. Costs a second matrix in HBM plus 16 B per A_pre non-zero in the rewritten rows; any other
 * order of use, a problem without fluid cells or the deterministic mode stay correct (foreground initialisation). */
int tfsi_bind_matrix_pair(tfsi_handle* h, double* A0_vals, double* A1_vals, const double* Apre_vals, void* stream);
int tfsi_unbind_matrix(tfsi_handle* h);
/* number of fluid launches that carried re-initialisation work for a matrix pair so far (reporting, tests) */
int64_t tfsi_fused_launches(const tfsi_handle* h);
/* entries rewritten by an assembly into the bound matrix (reporting) */
int64_t tfsi_bound_init_entries(const tfsi_handle* h);

/* Frees / forgets the column array (12 GB at 2 M tets) once set-up is complete: afterwards tfsi_assemble_A_pre,
 * tfsi_set_facet_operator and tfsi_csr_matvec fail with TFSI_ERR_UNSUPPORTED; the hot path does not need it. */
int tfsi_release_pattern(tfsi_handle* h);
/* colour classes per region when created with deterministic = 1 (0 otherwise) */
int32_t tfsi_num_colours(const tfsi_handle* h);

/* b = assemble(-F)                                          newtonsolver.py:86-88 */
int tfsi_assemble_residual(tfsi_handle* h, const double* x, const double* x1, double* b,
                           void* stream);

/* A.ident_zeros()  (rows with sum |a_ij| < DOLFIN_EPS get a unit diagonal)  newtonsolver.py:81 */
int tfsi_ident_zeros(tfsi_handle* h, double* A_vals, void* stream);

/* Dirichlet dof index set (fixed for the run; values change)  — DirichletBC objects of the
 * problem file's create_bcs, flattened in list order, later entries win.  bc_dofs: HOST or
 * DEVICE int32 [n]. */
int tfsi_set_bc_dofs(tfsi_handle* h, const int32_t* bc_dofs, int64_t n);
/* [bc.apply(A) for bc in bcs]            newtonsolver.py:82  zero rows, unit diagonal */
int tfsi_apply_bc_matrix(tfsi_handle* h, double* A_vals, void* stream);
/* [bc.apply(b, x) for bc in bcs]         newtonsolver.py:91  b_i = g_i - x_i ; g [n] values */
int tfsi_apply_bc_rhs(tfsi_handle* h, double* b, const double* x, const double* g, int64_t n, void* stream);
/* [bc.apply(x) for bc in bcs]            newtonsolver.py:98  x_i = g_i */
int tfsi_apply_bc_vec(tfsi_handle* h, double* x, const double* g, int64_t n, void* stream);
/* (n = length of g: must equal the size of the set given to tfsi_set_bc_dofs, else TFSI_ERR_ARG) */

/* b.norm('l2') / norm(dvp_res, 'l2')     newtonsolver.py:105-106 ; synchronises, writes the
 * SUM OF SQUARES to *out_sumsq (caller takes sqrt, after an all-reduce when distributed). */
int tfsi_sumsq(tfsi_handle* h, const double* v, int64_t n, double* out_sumsq, void* stream);

/* Host-buffer variants (reference-facing end-to-end call): x, x1 HOST [N]; b_host HOST [N];
 * A_vals stays on the device (consumer is the device sparse LU).  Includes H2D of x, x1, the
 * Jacobian (+A_pre, ident_zeros, BC rows) and residual (+BC) assembly, and D2H of b.
 * g_host: BC values [nbc] or NULL. x1_host may be NULL: the previous time level uploaded by the last call is
 * re-used (it changes once per time step, not per Newton iteration). Synchronous: ordered after the work queued on
 * `stream` (the stream on which the caller produced A_pre / A), runs on two internal streams (copies under
 * kernels: the iterate goes up in chunks of consecutive dofs and the residual batches start as soon as the chunks
 * they read have arrived; b goes back under the Jacobian kernels), returns when b_host and A_vals are final.
 * Pinned host buffers are needed for the copies to overlap. */
int tfsi_newton_assemble_host(tfsi_handle* h, const double* x_host, const double* x1_host,
                              const double* g_host, const double* Apre_vals, double* A_vals,
                              double* b_host, double* out_sumsq_b, int32_t recompute_jacobian, void* stream);

/* y = A x on the handle's CSR pattern (device arrays): residual check / iterative refinement around the
 * linear solve  up_sol.solve(dvp_res.vector(), b)  newtonsolver.py:93 — the solve itself stays outside. */
int tfsi_csr_matvec(tfsi_handle* h, const double* A_vals, const double* x, double* y, void* stream);

/* Halo update of the Newton iterate before an assembly (multi-GPU): packs x[send_idx] into
 * send_buf and unpacks recv_buf into x[recv_idx]; the NCCL send/recv between the two is
 * issued by the host language (torch.distributed) on the same stream. */
int tfsi_halo_pack(tfsi_handle* h, const double* x, const int32_t* send_idx, int64_t n,
                   double* send_buf, void* stream);
int tfsi_halo_unpack(tfsi_handle* h, double* x, const int32_t* recv_idx, int64_t n,
                     const double* recv_buf, void* stream);

/* Set-up utility (DOLFIN's SparsityPatternBuilder stand-in, used when the pattern is not given by
 * the host): expands the CSR adjacency of the P2 mesh nodes into the dof-level column array of a
 * node-interleaved numbering (dofs of node I = node_base[I] .. + node_ncomp[I]).  All pointers
 * DEVICE.  node_cum = exclusive prefix sum of node_ncomp[node_col] (length nnz_node + 1);
 * rowptr [N+1] must already hold the dof-level row starts; colidx [nnz] is written. */
int tfsi_expand_node_pattern(int64_t num_nodes, int64_t nnz_node, const int64_t* node_rowptr,
                             const int32_t* node_col, const int64_t* node_cum, const int64_t* node_base,
                             const int32_t* node_ncomp, const int64_t* rowptr, int32_t* colidx, void* stream);

/* Per-kernel device timing for bench.py's roofline: with profiling on, every launch of the
 * Jacobian (kind 0 fluid cells, 1 solid cells) and residual (kind 2) kernels is bracketed by CUDA
 * events on its own stream; tfsi_kernel_time returns their summed duration and the launch count
 * since profiling was switched on. */
int tfsi_set_profiling(tfsi_handle* h, int32_t on);
int tfsi_kernel_time(tfsi_handle* h, int32_t kind, double* total_ms, int64_t* count);

/* kernels launched since creation (for bench.py's gpu_launches) */
int64_t tfsi_launch_count(const tfsi_handle* h);

/* FP64 FMA peak probe: runs a register-resident DFMA loop on every SM, returns TFLOP/s. */
int tfsi_probe_fp64_peak(double* out_tflops, double* out_ms);

#ifdef __cplusplus
}
#endif
#endif
