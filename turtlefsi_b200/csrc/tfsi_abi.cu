// C ABI of libturtlefsi_b200.so (declared in include/turtlefsi_b200.h). Host orchestration only:
// uploads, scatter-map construction, kernel launches. No CPU compute path exists.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/turtlefsi_b200.h"
#include "tfsi_kernels.cuh"

using namespace tfsi;

namespace {

thread_local std::string g_create_error;

struct Region {
    RegionConst rc;
    int32_t marker;
    int32_t colour;       // colour class (deterministic mode), else 0
    int64_t ncells;
    int32_t* cell_list;   // device
    int32_t* res_cell_list; // device: the same cells ordered by their largest dof (residual batches; null = cell_list)
    uint16_t* rank_blk;   // device, per CTA batch (Jacobian kernels)
    int32_t* dofs_blk;    // device, per CTA batch
    int32_t* res_dofs_blk; // device, per residual CTA batch: [ndof][res_nc]
    size_t rank_blk_elems, dofs_blk_elems;   // elements per batch
    std::vector<int32_t> res_pmax;   // host entry: prefix maximum of the largest dof of every residual batch
    double* geo_blk;      // device, per CTA batch of the Jacobian kernels: [g*g+1][NC] inverse Jacobian and |det|
    double* res_geo_blk;  // the same per residual batch
    double* massJ;        // pair mode, fluid: frozen mass factors [batch][q][NC] (tfsi_assemble_A_pre)
    uint16_t* rank_blk_v; // pair mode, solid: ranks of the velocity columns per CTA batch (tfsi_bind_matrix_pair)
    size_t rank_v_elems;
};

}  // namespace

struct tfsi_handle {
    int dev;                  // CUDA device the handle lives on (every entry switches to it)
    int gdim, biharmonic, ndof, n2, n1;
    int64_t nv, nc, N, nnz;
    double *coords;
    int32_t *cell_vertices, *cell_dofs, *bc_dofs;
    const int32_t* colidx;    // owned copy, or the caller's array (borrow_pattern)
    const int64_t* rowptr;
    int own_pattern;
    int64_t* diagpos;
    double* scratch;          // 1 double (sumsq) + error counters
    int res_nc;               // cells per residual CTA (32 | 16 | 0 = unsupported rule)
    double *hx, *hx1, *hb, *hg;   // device staging for the _host entry (allocated at its first use)
    int hx1_valid;
    int64_t nbc;
    std::vector<Region> regions;
    int ncolours;
    cudaStream_t copy_stream, main_stream;   // host entry: copy_stream = H2D / D2H, main_stream = kernels
    cudaEvent_t ev_entry, ev_in, ev_b;
    QuadTab<MAXQ_J> tabJ;
    QuadTab<MAXQ_R> tabR;
    int fluid_on, solid_on, ext, ext_sub;
    int64_t launches;
    int sm_count;
    int profiling;
    // fixed linear facet operator (tfsi_set_facet_operator)
    int32_t *fo_rows, *fo_cols; int64_t* fo_pos; double* fo_vals; int64_t fo_n;
    int32_t *fp_rows, *fp_cols; double* fp_vals; int64_t fp_n;      // facet operator on the previous time level (residual only)
    int32_t* fl_dofs; double* fl_vals; int64_t fl_n; double fl_scale;  // state-independent facet load
    // compressed copy of A_pre (tfsi_compress_A_pre)
    const double* cpre_src;
    uint32_t* cpre_bitmap; int64_t* cpre_base; double* cpre_val; int64_t cpre_nnz, cpre_words;
    // persistent matrices (tfsi_bind_matrix / tfsi_bind_matrix_pair): word mask of the rows J_nl scatters into, and the
    // state of each bound buffer: CLEAN = holds A_pre in every row an assembly will scatter into (no initialisation
    // needed), USED = assembled into (masked re-initialisation needed), STALE = full initialisation needed
    struct Bound { double* A; int state; int pending; cudaEvent_t ready; } bound[2];
    int nbound;
    uint32_t* tmask; int64_t bound_entries;
    // pair mode (tfsi_bind_matrix_pair): the re-initialisation of the idle matrix is FUSED into the cell kernels — the
    // rows J_nl scatters into are zero-filled by the fluid kernel of the current assembly (TMA bulk stores from a zero
    // block), and the A_pre part of those rows (mass blocks of fluid.py:33 / solid.py:41, facet operators) is added by
    // the kernels of the next assembly into that matrix
    FillPiece* fill_pieces; int64_t n_fill;
    int64_t* fill_scalar; int64_t n_fill_scalar;
    uint8_t* row_flag;                        // rows J_nl scatters into
    const double* apre_last; double apre_sumsq;   // array written by the last tfsi_assemble_A_pre and its sum of squares
    int apre_own;                             // the compressed A_pre is that array (pair mode relies on its structure)
    int64_t fused_launches;                   // fluid launches that carried fill / fix-up work (tests, bench)
    int suppress_bg;                          // host entry: A := A_pre already runs under the H2D copy
    const std::vector<std::pair<int64_t, int64_t>>* res_ranges;   // host entry: residual batches [b0, b1) per region of the next launch
    std::vector<cudaEvent_t> ev_chunk;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[3];   // 0 jac_fluid, 1 jac_solid, 2 residual
    std::string err;
    int id;
};

namespace {

int g_next_id = 1;
int g_tab_owner[64] = {0};     // per device: id of the handle whose tables sit in constant memory

int fail(tfsi_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(h, call)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(h, TFSI_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                  \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                       \
    } while (0)

// every ABI entry runs on the handle's device, whatever the caller's current device is
struct DeviceGuard {
    int prev; bool switched;
    explicit DeviceGuard(const tfsi_handle* h) : prev(-1), switched(false) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->dev) { cudaSetDevice(h->dev); switched = true; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

template <class T>
int upload(tfsi_handle* h, T** dst, const T* src, size_t n) {
    CK(h, cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CK(h, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyDefault));
    return 0;
}

int bind_tables(tfsi_handle* h, cudaStream_t st) {
    int& owner = g_tab_owner[h->dev & 63];
    if (owner == h->id) return 0;
    // the tabulations live in __constant__ memory shared by all handles on this device: kernels of the previous
    // owner may still be in flight on another stream, so the (rare) change of owner waits for the device
    if (owner != 0) CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpyToSymbolAsync(c_tabJ, &h->tabJ, sizeof h->tabJ, 0, cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyToSymbolAsync(c_tabR, &h->tabR, sizeof h->tabR, 0, cudaMemcpyHostToDevice, st));
    {   // P2 reference mass matrix of the Jacobian rule (mass block of the solid kernel in pair mode)
        static thread_local double m2[10][10];
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j) {
                double z = 0.0;
                for (int q = 0; q < h->tabJ.nq; ++q) z = fma(h->tabJ.phi2[q][i], h->tabJ.w[q] * h->tabJ.phi2[q][j], z);
                m2[i][j] = z;
            }
        CK(h, cudaMemcpyToSymbolAsync(c_mass2, m2, sizeof m2, 0, cudaMemcpyHostToDevice, st));
    }
    owner = h->id;
    return 0;
}

CellArgs cell_args(const tfsi_handle* h, const Region& r, int nq, const double* x, const double* x1, double* out) {
    CellArgs a;
    a.coords = h->coords; a.cell_vertices = h->cell_vertices; a.cell_dofs = h->cell_dofs;
    a.colidx = h->colidx; a.rowptr = h->rowptr; a.cell_list = r.cell_list; a.rank_blk = r.rank_blk; a.dofs_blk = r.dofs_blk;
    a.geo_blk = r.geo_blk;
    a.ncell_region = r.ncells; a.ndof = h->ndof; a.q0 = 0; a.nq = nq; a.rc = r.rc; a.x = x; a.x1 = x1; a.out = out;
    a.fill_pieces = nullptr; a.n_fill = 0; a.fill_scalar = nullptr; a.n_fill_scalar = 0; a.fill_dst = nullptr;
    a.massJ = nullptr; a.rank_v = nullptr;
    return a;
}

// cells per batch of the Jacobian kernel that runs the region (layout of rank_blk / dofs_blk / geo_blk / massJ)
int region_ncta(const tfsi_handle* h, const Region& r) {
    const bool fl = r.rc.kind == TFSI_KIND_FLUID;
    if (h->gdim == 3) return fl ? FluidCfg<3>::L::NC : SolidCfg<3>::L::NC;
    return fl ? FluidCfg<2>::L::NC : SolidCfg<2>::L::NC;
}

struct KernelTimer {
    tfsi_handle* h; int kind; cudaStream_t st; cudaEvent_t e0, e1; bool on;
    KernelTimer(tfsi_handle* h_, int kind_, cudaStream_t st_) : h(h_), kind(kind_), st(st_), on(h_->profiling != 0) {
        if (on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    }
    ~KernelTimer() { if (on) { cudaEventRecord(e1, st); h->ev[kind].push_back({e0, e1}); } }
};

int grid_for(int64_t n, int block, int sm) {
    int64_t g = (n + block - 1) / block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sm * 16));
}

// largest chunk of quadrature points whose per-point table fits the shared memory of one CTA
template <class SM> int max_points(int nq, size_t extra = 0) {
    int m = nq;
    while (m > 1 && SM::bytes(m) + extra > 227 * 1024) --m;
    return m;
}

// J_nl of one region: one launch per chunk of the quadrature rule (the REDs of the chunks add up). The rule of the
// headline config (Keast 14-point, 3-D) runs in one launch of the compile-time-sized instantiation.
template <int G, int NQ>
int launch_fluid(tfsi_handle* h, const CellArgs& a, int grid, size_t smem, cudaStream_t st) {
    CK(h, cudaFuncSetAttribute(jac_fluid_kernel<G, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jac_fluid_kernel<G, NQ><<<grid, FluidJac<G>::THREADS, smem, st>>>(a);
    return 0;
}

template <int G, int NQ, int MAT, int MASS>
int launch_solid(tfsi_handle* h, const CellArgs& a, int grid, cudaStream_t st) {
    typedef typename SolidJac<G, MAT>::SM SM;
    const size_t smem = MASS ? SM::bytes_with_rank_v(a.nq) : SM::bytes(a.nq);
    if (smem > 227 * 1024)
        return fail(h, TFSI_ERR_UNSUPPORTED, "Jacobian chunk of %d points needs %zu B of shared memory", a.nq, smem);
    CK(h, cudaFuncSetAttribute(jac_solid_kernel<G, NQ, MAT, MASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jac_solid_kernel<G, NQ, MAT, MASS><<<grid, SolidJac<G, MAT>::THREADS, smem, st>>>(a);
    return 0;
}

// what ONE fluid launch of an assembly carries for the matrix pair (see CellArgs)
struct FusedWork { double* fill_dst; bool zeroed; bool used; };

template <int G>
int launch_jacobian_region(tfsi_handle* h, const Region& r, const double* x, const double* x1, double* A, cudaStream_t st,
                           FusedWork* fw = nullptr) {
    const int nq = h->tabJ.nq;
    CellArgs a = cell_args(h, r, nq, x, x1, A);
    const bool is_fluid = r.rc.kind == TFSI_KIND_FLUID;
    if (fw && fw->zeroed) {                  // the matrix holds zeros in the rows J_nl scatters into: + the A_pre part
        if (is_fluid) a.massJ = r.massJ;
        else a.rank_v = r.rank_blk_v;
    }
    const int ncta = region_ncta(h, r);
    const int grid = (int)((r.ncells + ncta - 1) / ncta);
    const bool energy = !is_fluid && r.rc.mat != MAT_SVK;        // energy-based solid model: tangent-table variant
    // (solid: sized as if the second rank table of the pair mode were always staged)
    const int chunk = is_fluid ? max_points<typename FluidJac<G>::SM>(nq)
                               : (energy ? max_points<typename SolidJac<G, 1>::SM>(nq, SolidJac<G, 1>::SM::RANK_BYTES)
                                         : max_points<typename SolidJac<G>::SM>(nq, SolidJac<G>::SM::RANK_BYTES));
    const int nchunk = (nq + chunk - 1) / chunk;
    for (int c = 0; c < nchunk; ++c) {
        a.q0 = (int)((int64_t)nq * c / nchunk);
        a.nq = (int)((int64_t)nq * (c + 1) / nchunk) - a.q0;           // balanced chunks, each <= chunk
        KernelTimer kt(h, is_fluid ? 0 : 1, st);
        if (is_fluid) {
            size_t smem = FluidJac<G>::SM::bytes(a.nq);
            if (smem > 227 * 1024)
                return fail(h, TFSI_ERR_UNSUPPORTED, "Jacobian chunk of %d points needs %zu B of shared memory", a.nq, smem);
            a.fill_dst = nullptr;
            if (fw && !fw->used && fw->fill_dst) {
                if (FluidJac<G>::SM::bytes_with_zero(a.nq) > 227 * 1024)
                    return fail(h, TFSI_ERR_UNSUPPORTED, "internal: fused re-initialisation does not fit beside %d points", a.nq);
                smem = FluidJac<G>::SM::bytes_with_zero(a.nq);
                a.fill_dst = fw->fill_dst; a.fill_pieces = h->fill_pieces; a.n_fill = h->n_fill;
                a.fill_scalar = h->fill_scalar; a.n_fill_scalar = h->n_fill_scalar;
                fw->used = true;
                h->fused_launches++;
            }
            if (int rc = (G == 3 && nq == 14) ? launch_fluid<G, 14>(h, a, grid, smem, st) : launch_fluid<G, 0>(h, a, grid, smem, st)) return rc;
        } else {
            if (c > 0) a.rank_v = nullptr;                       // the mass block goes with the first chunk only
            if (int rc = energy ? (a.rank_v ? launch_solid<G, 0, 1, 1>(h, a, grid, st) : launch_solid<G, 0, 1, 0>(h, a, grid, st))
                       : (G == 3 && nq == 14) ? (a.rank_v ? launch_solid<G, 14, 0, 1>(h, a, grid, st) : launch_solid<G, 14, 0, 0>(h, a, grid, st))
                                              : (a.rank_v ? launch_solid<G, 0, 0, 1>(h, a, grid, st) : launch_solid<G, 0, 0, 0>(h, a, grid, st)))
                return rc;
        }
        h->launches++;
        CK(h, cudaGetLastError());
    }
    if (!is_fluid && r.rc.visco) {
        // viscoelastic part of J_nl (solid.py:49,54): blocks psi x delta d (and psi x delta v for "Form1") through the
        // generic derivative-tensor kernel (binary-search scatter: needs the column array)
        if (!h->colidx) return fail(h, TFSI_ERR_ARG, "the viscoelastic solid needs the column array (tfsi_release_pattern was called)");
        const int off_v = G * Dim<G>::N2;
        LinArgs la;
        la.c = cell_args(h, r, nq, x, x1, A);
        la.nblk = 0;
        auto add = [&](int to, int tr, int gen) {
            LinBlock b; b.test_off = to; b.trial_off = tr; b.mass_kind = 0; b.cm = 0.0; b.ck = 0.0;
            b.ck_per_volume = 0; b.pad_ = 0; b.gen_kind = gen;
            la.blk[la.nblk++] = b;
        };
        if (r.rc.visco == 1) { add(off_v, 0, GEN_VISCO_F1_DD); add(off_v, off_v, GEN_VISCO_F1_DV); }
        else add(off_v, 0, GEN_VISCO_F2_DD);
        const int64_t ntask = r.ncells * la.nblk * 2;
        KernelTimer kt(h, 1, st);
        lin_jacobian_kernel<G><<<(unsigned)((ntask + 127) / 128), 128, 0, st>>>(la);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    return 0;
}

// state of a bound matrix. CLEAN: holds A_pre in every row an assembly scatters into (nothing to do); ZEROED (pair
// mode): the runs the cell kernels scatter into were zero-filled during the previous assembly, the A_pre part of
// them is added by this assembly's cell kernels themselves; USED: assembled into (masked re-initialisation needed); STALE: full
// initialisation needed
enum { BOUND_CLEAN = 0, BOUND_ZEROED = 1, BOUND_USED = 2, BOUND_STALE = 3 };

int find_bound(const tfsi_handle* h, const double* A) {
    for (int i = 0; i < h->nbound; ++i) if (h->bound[i].A == A) return i;
    return -1;
}
void mark_bound(tfsi_handle* h, int state) {          // A_pre / BC set changed: every bound matrix needs a full initialisation
    for (int i = 0; i < h->nbound; ++i) h->bound[i].state = std::max(h->bound[i].state, state);
}

// the first fluid launch of an assembly can carry the fused work: an active fluid region whose first quadrature chunk
// leaves room for the zero block, and no colour-ordered (deterministic) scatter
bool fused_possible(const tfsi_handle* h) {
    if (!h->fill_pieces || !h->apre_own || h->ncolours > 0 || getenv("TFSI_NO_FUSED_INIT")) return false;
    bool fluid = false;
    for (const Region& r : h->regions) {
        if (r.ncells == 0) continue;
        if (r.rc.kind == TFSI_KIND_FLUID) {
            if (!r.rc.fluid_on) continue;
            if (!r.massJ) return false;
            fluid = true;
        } else if (r.rc.solid_on && (!r.rank_blk_v || r.rc.visco == 2)) return false;     // "Form2" has a block in A_pre
    }
    if (!fluid) return false;
    const int nq = h->tabJ.nq;
    if (h->gdim == 3) {
        const int chunk = max_points<FluidJac<3>::SM>(nq), nchunk = (nq + chunk - 1) / chunk;
        return FluidJac<3>::SM::bytes_with_zero(nq / nchunk) <= 227 * 1024;
    }
    const int chunk = max_points<FluidJac<2>::SM>(nq), nchunk = (nq + chunk - 1) / chunk;
    return FluidJac<2>::SM::bytes_with_zero(nq / nchunk) <= 227 * 1024;
}

// A := (A_pre | 0) + J_nl.
// phase 0: everything; 1: only the x-independent initialisation A := A_pre; 2: only the cell kernels
// (the host entry runs phase 1 under the H2D copy of x)
int launch_jacobian(tfsi_handle* h, const double* x, const double* x1, const double* Apre, double* A, cudaStream_t st, int phase = 0) {
    if (Apre == TFSI_APRE_COMPRESSED && !h->cpre_bitmap)
        return fail(h, TFSI_ERR_ARG, "TFSI_APRE_COMPRESSED given but no compressed A_pre is registered (tfsi_compress_A_pre)");
    const bool use_compressed = h->cpre_bitmap && Apre && (Apre == TFSI_APRE_COMPRESSED || Apre == h->cpre_src);
    const int slot = find_bound(h, A);
    const int64_t ngroups = (h->cpre_words + 31) / 32;
    const bool can_fuse = use_compressed && slot >= 0 && h->nbound == 2 && fused_possible(h);
    if (slot >= 0 && h->bound[slot].state == BOUND_ZEROED && !can_fuse) h->bound[slot].state = BOUND_STALE;
    bool zeroed = false;
    if (phase != 2) {
        // A := A_pre streamed ONCE from the bitmap-compressed copy (zero-fill + packed non-zeros); into a bound matrix
        // only the rows that the cell kernels scatter into are rewritten, and nothing at all when it is still CLEAN
        if (use_compressed && slot >= 0 && (h->bound[slot].state == BOUND_CLEAN || h->bound[slot].state == BOUND_ZEROED)) {
            ;
        } else if (use_compressed) {
            const bool partial = slot >= 0 && h->tmask && h->bound[slot].state == BOUND_USED;
            bitmap_stream_kernel<1><<<grid_for(ngroups * 32, 256, h->sm_count), 256, 0, st>>>(
                h->cpre_bitmap, h->cpre_words, h->cpre_base, h->nnz, nullptr, h->cpre_val, A, partial ? h->tmask : nullptr);
            CK(h, cudaGetLastError());
            h->launches++;
            if (slot >= 0) h->bound[slot].state = BOUND_CLEAN;
        } else {
            if (Apre) CK(h, cudaMemcpyAsync(A, Apre, (size_t)h->nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
            else CK(h, cudaMemsetAsync(A, 0, (size_t)h->nnz * sizeof(double), st));
            h->launches++;
            if (slot >= 0) h->bound[slot].state = BOUND_STALE;      // not initialised from the compressed copy
        }
    }
    if (phase == 1) return 0;
    if (slot >= 0) {
        zeroed = h->bound[slot].state == BOUND_ZEROED;
        h->bound[slot].state = (h->bound[slot].state == BOUND_CLEAN || h->bound[slot].state == BOUND_ZEROED) ? BOUND_USED : BOUND_STALE;
    }
    // pair mode: the OTHER matrix of the pair is zero-filled by this assembly's fluid launch, under its FP64 loops
    const int other = (slot >= 0 && h->nbound == 2) ? 1 - slot : -1;
    FusedWork fw;
    fw.fill_dst = (can_fuse && other >= 0 && h->bound[other].state == BOUND_USED && !h->suppress_bg) ? h->bound[other].A : nullptr;
    fw.zeroed = zeroed;
    fw.used = false;
    for (const Region& r : h->regions) {
        if (r.ncells == 0) continue;
        if (r.rc.kind == TFSI_KIND_FLUID ? !r.rc.fluid_on : !r.rc.solid_on) continue;
        int rc = h->gdim == 3 ? launch_jacobian_region<3>(h, r, x, x1, A, st, &fw) : launch_jacobian_region<2>(h, r, x, x1, A, st, &fw);
        if (rc) return rc;
    }
    if (fw.fill_dst && !fw.used) return fail(h, TFSI_ERR_UNSUPPORTED, "internal: no fluid launch carried the fused re-initialisation");
    if (fw.fill_dst) h->bound[other].state = BOUND_ZEROED;
    if (zeroed && h->fo_n) {                 // facet-operator part of A_pre in the zero-filled rows
        coo_add_flagged_kernel<<<grid_for(h->fo_n, 256, h->sm_count), 256, 0, st>>>(h->fo_pos, h->fo_vals, h->fo_rows, h->row_flag, h->fo_n, A);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    return 0;
}

template <int G>
int launch_linear(tfsi_handle* h, const double* x, const double* x1, double* A, cudaStream_t st) {
    constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    const int off_d = 0, off_v = G * N2, off_w = 2 * G * N2 + N1;
    for (const Region& r : h->regions) {
        if (r.ncells == 0) continue;
        LinArgs la;
        la.c = cell_args(h, r, h->tabJ.nq, x, x1, A);
        la.nblk = 0;
        auto add = [&](int to, int tr, int mk, double cm, double ck, int per_vol) {
            LinBlock b; b.test_off = to; b.trial_off = tr; b.mass_kind = mk; b.cm = cm; b.ck = ck;
            b.ck_per_volume = per_vol; b.pad_ = 0; b.gen_kind = 0;
            la.blk[la.nblk++] = b;
        };
        const RegionConst& rc = r.rc;
        if (rc.kind == TFSI_KIND_FLUID) {
            if (rc.fluid_on) add(off_v, off_v, 2, rc.c1lin, 0.0, 0);                  // fluid.py:33
            if (rc.ext == TFSI_EXT_LAPLACE && rc.ext_sub == TFSI_LAPLACE_VOLUME_CHANGE) {  // laplace.py:33-34, frozen at x
                add(off_d, off_d, 0, 0.0, 0.0, 0);
                la.blk[la.nblk - 1].gen_kind = 1;
            } else if (rc.ext == TFSI_EXT_LAPLACE)                                      // laplace.py:45
                add(off_d, off_d, 0, 0.0, rc.ext_sub == TFSI_LAPLACE_VOLUME ? 1.0 : rc.alfa,
                    rc.ext_sub == TFSI_LAPLACE_VOLUME);
            if (rc.ext == TFSI_EXT_ELASTIC) {                                           // elastic.py:24-25, frozen at x
                add(off_d, off_d, 0, 0.0, 0.0, 0);
                la.blk[la.nblk - 1].gen_kind = 2;
            }
            if (rc.ext == TFSI_EXT_BIHARMONIC) {                                        // biharmonic.py:35-36
                add(off_w, off_w, 1, rc.alfa, 0.0, 0);
                add(off_w, off_d, 0, 0.0, -rc.alfa, 0);
                add(off_d, off_w, 0, 0.0, rc.alfa, 0);
            }
        } else if (rc.solid_on) {                                                       // solid.py:41-43
            add(off_v, off_v, 1, rc.rho_s_k, 0.0, 0);
            add(off_d, off_d, 1, rc.delta_rho_k, 0.0, 0);
            add(off_d, off_v, 1, -rc.delta_rho * rc.th0, 0.0, 0);
            if (rc.visco == 2) {                                                        // solid.py:55, frozen at x1
                add(off_v, off_d, 0, 0.0, 0.0, 0);
                la.blk[la.nblk - 1].gen_kind = GEN_VISCO_F2_LIN;
            }
        }
        if (la.nblk == 0) continue;
        const int64_t ntask = r.ncells * la.nblk * 2;
        lin_jacobian_kernel<G><<<(unsigned)((ntask + 127) / 128), 128, 0, st>>>(la);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    return 0;
}

// solid regions with an energy-based material model or viscoelasticity run the LIFT_SOLID_GENERAL instantiation
inline bool general_solid(const RegionConst& rc) { return rc.kind == TFSI_KIND_SOLID && (rc.mat != MAT_SVK || rc.visco != 0); }

template <int G, bool BIH, int NC, int LIFT>
int launch_residual_nc(tfsi_handle* h, const double* x, const double* x1, double* b, cudaStream_t st) {
    typedef ResCfg<G, BIH> R;
    const size_t smem = R::bytes(h->tabR.nq, NC);
    CK(h, cudaFuncSetAttribute(residual_kernel<G, BIH, NC, LIFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (size_t ri = 0; ri < h->regions.size(); ++ri) {
        const Region& r = h->regions[ri];
        if (r.ncells == 0) continue;
        const RegionConst& rc = r.rc;
        if (general_solid(rc) != (LIFT == LIFT_SOLID_GENERAL)) continue;
        if (rc.kind == TFSI_KIND_FLUID && !rc.fluid_on && rc.ext == 0) continue;
        if (rc.kind == TFSI_KIND_SOLID && !rc.solid_on) continue;
        int64_t b0 = 0, b1 = (r.ncells + NC - 1) / NC;
        if (h->res_ranges) { b0 = (*h->res_ranges)[ri].first; b1 = std::min(b1, (*h->res_ranges)[ri].second); }
        if (b0 >= b1) continue;
        CellArgs a = cell_args(h, r, h->tabR.nq, x, x1, b);
        a.dofs_blk = r.res_dofs_blk + (size_t)b0 * h->ndof * NC;         // batches [b0, b1) of the region
        a.geo_blk = r.res_geo_blk ? r.res_geo_blk + (size_t)b0 * (G * G + 1) * NC : nullptr;
        a.cell_list = (r.res_cell_list ? r.res_cell_list : r.cell_list) + b0 * NC;
        a.ncell_region = std::min<int64_t>(r.ncells - b0 * NC, (b1 - b0) * NC);
        KernelTimer kt(h, 2, st);
        residual_kernel<G, BIH, NC, LIFT><<<(unsigned)(b1 - b0), NC * 8, smem, st>>>(a);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    return 0;
}

size_t residual_smem(const tfsi_handle* h, int nc) {
    const int nq = h->tabR.nq;
    if (h->gdim == 3) return h->biharmonic ? ResCfg<3, true>::bytes(nq, nc) : ResCfg<3, false>::bytes(nq, nc);
    return h->biharmonic ? ResCfg<2, true>::bytes(nq, nc) : ResCfg<2, false>::bytes(nq, nc);
}

// cells per residual CTA: 16 (4 warps, two CTAs per SM overlap their phases) when two of them fit into the
// shared memory of one SM, else 32 (8 warps, one CTA per SM), else 16 alone; 0 = the rule does not fit
int residual_nc(const tfsi_handle* h) {
    static const int pref = getenv("TFSI_RES_NC") ? atoi(getenv("TFSI_RES_NC")) : 16;
    if (pref == 16 && residual_smem(h, 16) <= 112 * 1024) return 16;
    if (residual_smem(h, 32) <= 220 * 1024) return 32;
    if (residual_smem(h, 16) <= 220 * 1024) return 16;
    return 0;
}

template <int G, bool BIH>
int launch_residual_t(tfsi_handle* h, const double* x, const double* x1, double* b, cudaStream_t st) {
    bool any_general = false;
    for (const Region& r : h->regions) any_general = any_general || (r.ncells && general_solid(r.rc));
    if (any_general) {
        int rc = h->res_nc == 32 ? launch_residual_nc<G, BIH, 32, LIFT_SOLID_GENERAL>(h, x, x1, b, st)
               : h->res_nc == 16 ? launch_residual_nc<G, BIH, 16, LIFT_SOLID_GENERAL>(h, x, x1, b, st)
               : fail(h, TFSI_ERR_UNSUPPORTED, "residual rule with %d points does not fit in shared memory", h->tabR.nq);
        if (rc) return rc;
    }
    if (!BIH && h->ext == TFSI_EXT_LAPLACE && h->ext_sub == TFSI_LAPLACE_VOLUME_CHANGE) {      // alfa = 1/J(d)
        if (h->res_nc == 32) return launch_residual_nc<G, false, 32, LIFT_VOLUME_CHANGE>(h, x, x1, b, st);
        if (h->res_nc == 16) return launch_residual_nc<G, false, 16, LIFT_VOLUME_CHANGE>(h, x, x1, b, st);
    }
    if (!BIH && h->ext == TFSI_EXT_ELASTIC) {                                                   // elastic.py
        if (h->res_nc == 32) return launch_residual_nc<G, false, 32, LIFT_ELASTIC>(h, x, x1, b, st);
        if (h->res_nc == 16) return launch_residual_nc<G, false, 16, LIFT_ELASTIC>(h, x, x1, b, st);
    }
    if (h->res_nc == 32) return launch_residual_nc<G, BIH, 32, LIFT_DEFAULT>(h, x, x1, b, st);
    if (h->res_nc == 16) return launch_residual_nc<G, BIH, 16, LIFT_DEFAULT>(h, x, x1, b, st);
    return fail(h, TFSI_ERR_UNSUPPORTED, "residual rule with %d points does not fit in shared memory", h->tabR.nq);
}

int launch_residual(tfsi_handle* h, const double* x, const double* x1, double* b, cudaStream_t st) {
    if (h->gdim == 3) return h->biharmonic ? launch_residual_t<3, true>(h, x, x1, b, st) : launch_residual_t<3, false>(h, x, x1, b, st);
    return h->biharmonic ? launch_residual_t<2, true>(h, x, x1, b, st) : launch_residual_t<2, false>(h, x, x1, b, st);
}

// the fixed facet operators and the facet load of the residual (b -= A_f x, b -= A_p x1, b -= p(t) l)
int residual_facet_terms(tfsi_handle* h, const double* x, const double* x1, double* b, cudaStream_t st) {
    if (h->fo_n) {
        coo_sub_matvec_kernel<<<grid_for(h->fo_n, 256, h->sm_count), 256, 0, st>>>(h->fo_rows, h->fo_cols, h->fo_vals, h->fo_n, x, b);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    if (h->fp_n) {
        coo_sub_matvec_kernel<<<grid_for(h->fp_n, 256, h->sm_count), 256, 0, st>>>(h->fp_rows, h->fp_cols, h->fp_vals, h->fp_n, x1, b);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    if (h->fl_n && h->fl_scale != 0.0) {
        sparse_sub_scaled_kernel<<<grid_for(h->fl_n, 256, h->sm_count), 256, 0, st>>>(h->fl_dofs, h->fl_vals, h->fl_n, h->fl_scale, b);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    return 0;
}


}  // namespace

extern "C" {

const char* tfsi_last_error(const tfsi_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int64_t tfsi_nnz(const tfsi_handle* h) { return h ? h->nnz : -1; }
int32_t tfsi_ndof_cell(const tfsi_handle* h) { return h ? h->ndof : -1; }
int64_t tfsi_launch_count(const tfsi_handle* h) { return h ? h->launches : -1; }

void tfsi_destroy(tfsi_handle* h) {
    if (!h) return;
    DeviceGuard dg(h);
    cudaFree(h->coords); cudaFree(h->cell_vertices); cudaFree(h->cell_dofs);
    if (h->own_pattern) { cudaFree((void*)h->colidx); cudaFree((void*)h->rowptr); }
    cudaFree(h->diagpos); cudaFree(h->bc_dofs); cudaFree(h->scratch);
    cudaFree(h->hx); cudaFree(h->hx1); cudaFree(h->hb); cudaFree(h->hg);
    cudaFree(h->fo_rows); cudaFree(h->fo_cols); cudaFree(h->fo_pos); cudaFree(h->fo_vals);
    cudaFree(h->fp_rows); cudaFree(h->fp_cols); cudaFree(h->fp_vals); cudaFree(h->fl_dofs); cudaFree(h->fl_vals);
    cudaFree(h->cpre_bitmap); cudaFree(h->cpre_base); cudaFree(h->cpre_val); cudaFree(h->tmask);
    cudaFree(h->fill_pieces); cudaFree(h->fill_scalar); cudaFree(h->row_flag);
    for (auto& b : h->bound) if (b.ready) cudaEventDestroy(b.ready);
    for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
    for (Region& r : h->regions) { cudaFree(r.cell_list); cudaFree(r.res_cell_list); cudaFree(r.rank_blk); cudaFree(r.dofs_blk); cudaFree(r.res_dofs_blk); cudaFree(r.massJ); cudaFree(r.rank_blk_v); cudaFree(r.geo_blk); cudaFree(r.res_geo_blk); }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->main_stream) cudaStreamDestroy(h->main_stream);
    if (h->ev_entry) cudaEventDestroy(h->ev_entry);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
    for (auto& v : h->ev) for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (g_tab_owner[h->dev & 63] == h->id) g_tab_owner[h->dev & 63] = 0;
    delete h;
}

namespace {

// Greedy colouring of the cells of one region: two cells conflict when they share a mesh VERTEX (then, and only
// then, they share P2 nodes / CSR entries). A cell takes the least-used colour that none of its vertices has seen
// (balanced classes: every class is one kernel launch); a new colour is opened when all are taken.
// Returns the colour of every listed cell; ncol = number of classes.
std::vector<int> colour_cells(const int32_t* cell_vertices, int nvpc, int64_t nv, const std::vector<int32_t>& list, int& ncol) {
    constexpr int W = 4;                                    // up to 256 colours
    std::vector<uint64_t> vmask((size_t)nv * W, 0);
    std::vector<int64_t> used;
    std::vector<int> colour(list.size(), 0);
    for (size_t k = 0; k < list.size(); ++k) {
        const int32_t* v = cell_vertices + (size_t)list[k] * nvpc;
        uint64_t forb[W] = {0, 0, 0, 0};
        for (int i = 0; i < nvpc; ++i) for (int w = 0; w < W; ++w) forb[w] |= vmask[(size_t)v[i] * W + w];
        int best = -1;
        for (int c = 0; c < (int)used.size(); ++c)
            if (!((forb[c >> 6] >> (c & 63)) & 1ull) && (best < 0 || used[c] < used[best])) best = c;
        if (best < 0) {
            if ((int)used.size() >= 64 * W) { ncol = -1; return colour; }
            best = (int)used.size();
            used.push_back(0);
        }
        used[best]++;
        colour[k] = best;
        for (int i = 0; i < nvpc; ++i) vmask[(size_t)v[i] * W + (best >> 6)] |= 1ull << (best & 63);
    }
    ncol = (int)used.size();
    return colour;
}

}  // namespace

int tfsi_create(const tfsi_problem* p, tfsi_handle** out) {
    if (!p || !out) return fail(nullptr, TFSI_ERR_ARG, "null argument");
    *out = nullptr;
    if (p->gdim != 2 && p->gdim != 3) return fail(nullptr, TFSI_ERR_ARG, "gdim must be 2 or 3");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, TFSI_ERR_CUDA, "no CUDA device: turtlefsi_b200 has no CPU fallback");
    tfsi_handle* h = new tfsi_handle();
    h->id = g_next_id++;
    const int g = p->gdim;
    h->gdim = g; h->biharmonic = p->biharmonic ? 1 : 0;
    h->n1 = g + 1; h->n2 = (g == 2) ? 6 : 10;
    h->ndof = (h->biharmonic ? 3 : 2) * g * h->n2 + h->n1;
    h->nv = p->num_vertices; h->nc = p->num_cells; h->N = p->num_dofs;
    h->coords = nullptr; h->cell_vertices = h->cell_dofs = h->bc_dofs = nullptr;
    h->colidx = nullptr; h->rowptr = nullptr; h->own_pattern = 1;
    h->diagpos = nullptr; h->scratch = nullptr;
    h->hx = h->hx1 = h->hb = h->hg = nullptr; h->hx1_valid = 0;
    h->nbc = 0; h->launches = 0; h->profiling = 0; h->ncolours = 0;
    h->copy_stream = h->main_stream = nullptr; h->ev_entry = h->ev_in = h->ev_b = nullptr;
    h->fo_rows = h->fo_cols = nullptr; h->fo_pos = nullptr; h->fo_vals = nullptr; h->fo_n = 0;
    h->fp_rows = h->fp_cols = nullptr; h->fp_vals = nullptr; h->fp_n = 0;
    h->fl_dofs = nullptr; h->fl_vals = nullptr; h->fl_n = 0; h->fl_scale = 1.0;
    h->cpre_src = nullptr; h->cpre_bitmap = nullptr; h->cpre_base = nullptr; h->cpre_val = nullptr; h->cpre_nnz = 0; h->cpre_words = 0;
    h->nbound = 0; h->tmask = nullptr; h->bound_entries = 0; h->fused_launches = 0;
    for (auto& b : h->bound) { b.A = nullptr; b.state = BOUND_STALE; b.pending = 0; b.ready = nullptr; }
    h->fill_pieces = nullptr; h->fill_scalar = nullptr; h->n_fill = h->n_fill_scalar = 0; h->row_flag = nullptr; h->suppress_bg = 0;
    h->apre_last = nullptr; h->apre_sumsq = 0.0; h->apre_own = 0; h->res_ranges = nullptr;
    h->fluid_on = p->fluid_on; h->solid_on = p->solid_on; h->ext = p->extrapolation; h->ext_sub = p->extrapolation_sub_type;
    if (h->ext == TFSI_EXT_BIHARMONIC && !h->biharmonic) { delete h; return fail(nullptr, TFSI_ERR_ARG, "biharmonic lifting needs the 4-field space"); }
    if (h->ext < 0 || h->ext > TFSI_EXT_ELASTIC) { delete h; return fail(nullptr, TFSI_ERR_ARG, "unknown extrapolation code %d", h->ext); }
    h->dev = 0;
    cudaGetDevice(&h->dev);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->dev);

#define CKC(call) do { int rc_ = (call); if (rc_) { g_create_error = h->err; tfsi_destroy(h); return rc_; } } while (0)
    auto cka = [&](cudaError_t e) { if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return TFSI_ERR_CUDA; } return 0; };
    if (tabulate<MAXQ_J>(g, p->nq_jacobian, p->qpts_jacobian, p->qwts_jacobian, h->tabJ) ||
        tabulate<MAXQ_R>(g, p->nq_residual, p->qpts_residual, p->qwts_residual, h->tabR)) {
        tfsi_destroy(h);
        return fail(nullptr, TFSI_ERR_UNSUPPORTED, "quadrature rule too large (Jacobian <= %d, residual <= %d points)", MAXQ_J, MAXQ_R);
    }
    CKC(upload(h, &h->coords, p->coords, (size_t)h->nv * g));
    CKC(upload(h, &h->cell_vertices, p->cell_vertices, (size_t)h->nc * (g + 1)));
    CKC(upload(h, &h->cell_dofs, p->cell_dofs, (size_t)h->nc * h->ndof));
    if (p->borrow_pattern) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, p->colidx) != cudaSuccess || pa.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            tfsi_destroy(h);
            return fail(nullptr, TFSI_ERR_ARG, "borrow_pattern needs DEVICE arrays rowptr / colidx");
        }
        h->own_pattern = 0; h->rowptr = p->rowptr; h->colidx = p->colidx;
    } else {
        int64_t* rp = nullptr;
        CKC(upload(h, &rp, p->rowptr, (size_t)h->N + 1));
        h->rowptr = rp;
    }
    {
        int64_t nnz = 0;
        cudaError_t e = cudaMemcpy(&nnz, h->rowptr + h->N, sizeof nnz, cudaMemcpyDefault);
        if (e != cudaSuccess) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_CUDA, "rowptr readback: %s", cudaGetErrorString(e)); }
        h->nnz = nnz;
    }
    if (!p->borrow_pattern) {
        int32_t* ci = nullptr;
        CKC(upload(h, &ci, p->colidx, (size_t)h->nnz));
        h->colidx = ci;
    }
    CKC(cka(cudaMalloc((void**)&h->diagpos, (size_t)h->N * sizeof(int64_t))));
    CKC(cka(cudaMalloc((void**)&h->scratch, 64)));
    // regions: host pass over the markers (and, for the colour-ordered scatter, over the cell vertices)
    {
        std::vector<int32_t> marker((size_t)h->nc);
        cudaError_t e = cudaMemcpy(marker.data(), p->cell_marker, (size_t)h->nc * sizeof(int32_t), cudaMemcpyDefault);
        if (e != cudaSuccess) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_CUDA, "marker copy: %s", cudaGetErrorString(e)); }
        std::vector<int32_t> cv;
        if (p->deterministic) {
            cv.resize((size_t)h->nc * (g + 1));
            e = cudaMemcpy(cv.data(), p->cell_vertices, cv.size() * sizeof(int32_t), cudaMemcpyDefault);
            if (e != cudaSuccess) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_CUDA, "cell copy: %s", cudaGetErrorString(e)); }
        }
        for (int i = 0; i < p->num_regions; ++i) {
            const tfsi_region& rg = p->regions[i];
            Region r;
            r.marker = rg.marker; r.colour = 0;
            r.rc = make_region_const(rg.kind, rg.rho, rg.mu, rg.lambda, p->dt, p->theta, p->fluid_on, p->solid_on,
                                     p->extrapolation, p->extrapolation_sub_type, p->hmin, p->has_gravity ? p->gravity : 0.0);
            if (rg.kind == TFSI_KIND_SOLID) {
                if (rg.material < TFSI_MAT_SVK || rg.material > TFSI_MAT_GENT || rg.viscoelasticity < 0 || rg.viscoelasticity > 2) {
                    tfsi_destroy(h);
                    return fail(nullptr, TFSI_ERR_ARG, "region %d: unknown material model %d / viscoelasticity %d", i, rg.material, rg.viscoelasticity);
                }
                if (rg.material == TFSI_MAT_GENT && !(rg.Jm > 0.0)) {
                    tfsi_destroy(h);
                    return fail(nullptr, TFSI_ERR_ARG, "region %d: the Gent model needs Jm > 0", i);
                }
                set_solid_model(r.rc, rg.material, rg.C01, rg.C10, rg.C11, rg.Jm, rg.viscoelasticity, rg.mu_visc, rg.lambda_visc);
            }
            r.cell_list = nullptr; r.res_cell_list = nullptr; r.rank_blk = nullptr; r.dofs_blk = nullptr; r.res_dofs_blk = nullptr; r.rank_blk_elems = r.dofs_blk_elems = 0;
            r.massJ = nullptr; r.rank_blk_v = nullptr; r.rank_v_elems = 0; r.geo_blk = nullptr; r.res_geo_blk = nullptr;
            std::vector<int32_t> list;
            for (int64_t c = 0; c < h->nc; ++c) if (marker[(size_t)c] == rg.marker) list.push_back((int32_t)c);
            if (!p->deterministic) {
                r.ncells = (int64_t)list.size();
                CKC(upload(h, &r.cell_list, list.data(), list.size()));
                h->regions.push_back(r);
                continue;
            }
            int ncol = 0;
            std::vector<int> col = colour_cells(cv.data(), g + 1, h->nv, list, ncol);
            if (ncol < 0) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_UNSUPPORTED, "more than 256 colour classes"); }
            h->ncolours = std::max(h->ncolours, ncol);
            std::vector<std::vector<int32_t>> cls((size_t)ncol);
            for (size_t k = 0; k < list.size(); ++k) cls[(size_t)col[k]].push_back(list[k]);
            for (int c = 0; c < ncol; ++c) {
                Region rcpy = r;
                rcpy.colour = c;
                rcpy.ncells = (int64_t)cls[(size_t)c].size();
                CKC(upload(h, &rcpy.cell_list, cls[(size_t)c].data(), cls[(size_t)c].size()));
                h->regions.push_back(rcpy);
            }
        }
    }
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->main_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_entry, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming);
    // pattern checks + diagonal positions
    int* derr = reinterpret_cast<int*>(h->scratch + 1);
    {
        cudaMemset(derr, 0, sizeof(int));
        const int64_t total = h->nc * h->n2 * h->ndof;
        check_pattern_kernel<<<grid_for(total, 256, h->sm_count), 256>>>(h->cell_dofs, h->rowptr, h->colidx, h->nc, h->n2, h->ndof, derr);
        diagpos_kernel<<<grid_for(h->N, 256, h->sm_count), 256>>>(h->rowptr, h->colidx, h->diagpos, h->N, derr);
        int herr = 0;
        cudaError_t e = cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_CUDA, "scatter-map kernels: %s", cudaGetErrorString(e)); }
        if (herr) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_PATTERN, "%d cell-block entries (or diagonals) are missing from the CSR pattern, or a row exceeds 65535 entries", herr); }
        if (p->verify_pattern) {
            std::vector<int32_t> nol((size_t)h->ndof);
            for (int r = 0; r < h->ndof; ++r) {
                const int pofs = 2 * g * h->n2;
                if (r < pofs) nol[r] = r % h->n2;
                else if (r < pofs + h->n1) nol[r] = r - pofs;
                else nol[r] = (r - pofs - h->n1) % h->n2;
            }
            int32_t* dnol = nullptr;
            CKC(upload(h, &dnol, nol.data(), nol.size()));
            const int64_t tot2 = h->nc * h->ndof * h->ndof;
            const int64_t stride = (p->verify_pattern > 1) ? p->verify_pattern : 1;
            verify_rank_kernel<<<grid_for(tot2 / stride + 1, 256, h->sm_count), 256>>>(h->cell_dofs, h->rowptr, h->colidx, dnol, h->nc, h->n2, h->ndof, stride, derr);
            e = cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost);
            cudaFree(dnol);
            if (e != cudaSuccess) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_CUDA, "verify kernel: %s", cudaGetErrorString(e)); }
            if (herr) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_PATTERN, "CSR pattern is not a union of dense cell blocks with node-uniform column sets (%d mismatches)", herr); }
        }
    }
    // per-CTA-batch scatter ranks and blocked dof copies for the cell kernels (bulk-copy / coalesced staging)
    h->res_nc = residual_nc(h);
    for (Region& r : h->regions) {
        if (r.ncells == 0) continue;
        const int g3 = (g == 3);
        const bool fl = r.rc.kind == TFSI_KIND_FLUID;
        const int ncta = region_ncta(h, r);
        const int rks = ncta + 1;
        const int ncols = fl ? (2 * g * h->n2 + h->n1) : g * h->n2;
        const int ncoef = fl ? (2 * g * h->n2 + h->n1) : 2 * g * h->n2;
        size_t rbytes = ((size_t)h->n2 * ncols * rks * sizeof(uint16_t) + 15) / 16 * 16;
        r.rank_blk_elems = rbytes / sizeof(uint16_t);
        r.dofs_blk_elems = (size_t)ncoef * ncta;
        const int64_t nb = (r.ncells + ncta - 1) / ncta;
        CKC(cka(cudaMalloc((void**)&r.rank_blk, (size_t)nb * rbytes)));
        CKC(cka(cudaMemset(r.rank_blk, 0, (size_t)nb * rbytes)));
        CKC(cka(cudaMalloc((void**)&r.dofs_blk, (size_t)nb * r.dofs_blk_elems * sizeof(int32_t))));
        build_rank_blocked_kernel<<<grid_for(nb * h->n2 * ncols * ncta, 256, h->sm_count), 256>>>(
            h->cell_dofs, h->rowptr, h->colidx, r.cell_list, r.ncells, h->ndof, h->n2, ncols, ncta, rks, r.rank_blk_elems, r.rank_blk, derr);
        build_dofs_blocked_kernel<<<grid_for(nb * ncoef * ncta, 256, h->sm_count), 256>>>(
            h->cell_dofs, r.cell_list, r.ncells, h->ndof, ncoef, ncta, r.dofs_blk);
        {                                                    // staged geometry of the Jacobian kernels (80 B per tet)
            double* gb = nullptr;
            CKC(cka(cudaMalloc((void**)&gb, (size_t)nb * (g * g + 1) * ncta * sizeof(double))));
            CellArgs ga = cell_args(h, r, 0, nullptr, nullptr, nullptr);
            if (g3) build_geo_blocked_kernel<3><<<grid_for(nb * ncta, 128, h->sm_count), 128>>>(ga, ncta, gb);
            else build_geo_blocked_kernel<2><<<grid_for(nb * ncta, 128, h->sm_count), 128>>>(ga, ncta, gb);
            r.geo_blk = gb;
        }
        if (h->res_nc) {                                     // residual kernel: [ndof][res_nc] per batch
            const int64_t nbr = (r.ncells + h->res_nc - 1) / h->res_nc;
            if (!p->deterministic) {
                // The residual batches take the region's cells in the order of their LARGEST dof: the host entry uploads the
                // iterate in chunks of consecutive dofs and starts a batch when the chunk holding its largest dof has
                // arrived. In the region's own (Morton) order a single cell across a high-order Morton boundary raises the
                // prefix maximum for every later batch, and the residual then waits for late chunks (6.0 ms of wall time
                // for 5.5 ms of kernels in the end-to-end call). Neighbouring cells still sit in neighbouring batches.
                int32_t *keys = nullptr, *keys_out = nullptr;
                void* tmp = nullptr;
                size_t tmp_bytes = 0;
                CKC(cka(cudaMalloc((void**)&keys, (size_t)r.ncells * sizeof(int32_t))));
                CKC(cka(cudaMalloc((void**)&keys_out, (size_t)r.ncells * sizeof(int32_t))));
                CKC(cka(cudaMalloc((void**)&r.res_cell_list, (size_t)r.ncells * sizeof(int32_t))));
                cell_max_dof_kernel<<<grid_for(r.ncells, 256, h->sm_count), 256>>>(h->cell_dofs, r.cell_list, r.ncells, h->ndof, keys);
                cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, r.cell_list, r.res_cell_list, (int)r.ncells);
                CKC(cka(cudaMalloc(&tmp, tmp_bytes)));
                cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, r.cell_list, r.res_cell_list, (int)r.ncells);
                CKC(cka(cudaDeviceSynchronize()));
                cudaFree(tmp); cudaFree(keys); cudaFree(keys_out);
            }
            const int32_t* rlist = r.res_cell_list ? r.res_cell_list : r.cell_list;
            CKC(cka(cudaMalloc((void**)&r.res_dofs_blk, (size_t)nbr * h->ndof * h->res_nc * sizeof(int32_t))));
            build_dofs_blocked_kernel<<<grid_for(nbr * h->ndof * h->res_nc, 256, h->sm_count), 256>>>(
                h->cell_dofs, rlist, r.ncells, h->ndof, h->ndof, h->res_nc, r.res_dofs_blk);
            {
                double* gb = nullptr;
                CKC(cka(cudaMalloc((void**)&gb, (size_t)nbr * (g * g + 1) * h->res_nc * sizeof(double))));
                CellArgs ga = cell_args(h, r, 0, nullptr, nullptr, nullptr);
                ga.cell_list = rlist;
                if (g3) build_geo_blocked_kernel<3><<<grid_for(nbr * h->res_nc, 128, h->sm_count), 128>>>(ga, h->res_nc, gb);
                else build_geo_blocked_kernel<2><<<grid_for(nbr * h->res_nc, 128, h->sm_count), 128>>>(ga, h->res_nc, gb);
                r.res_geo_blk = gb;
            }
        }
    }
    {
        int herr = 0;
        CKC(cka(cudaDeviceSynchronize()));
        CKC(cka(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost)));
        if (herr) { tfsi_destroy(h); return fail(nullptr, TFSI_ERR_PATTERN, "%d scatter ranks could not be resolved in the CSR pattern", herr); }
    }
#undef CKC
    *out = h;
    return TFSI_OK;
}

int32_t tfsi_num_colours(const tfsi_handle* h) { return h ? h->ncolours : -1; }

int tfsi_release_pattern(tfsi_handle* h) {
    if (!h) return TFSI_ERR_ARG;
    DeviceGuard dg(h);
    CK(h, cudaDeviceSynchronize());
    if (h->own_pattern) cudaFree((void*)h->colidx);
    h->colidx = nullptr;
    return 0;
}

int tfsi_assemble_A_pre(tfsi_handle* h, const double* x, const double* x1, double* Apre, void* stream) {
    if (!h || !x || !x1 || !Apre) return fail(h, TFSI_ERR_ARG, "null argument");
    if (!h->colidx) return fail(h, TFSI_ERR_UNSUPPORTED, "the column array was released (tfsi_release_pattern)");
    DeviceGuard dg(h);
    if (Apre == h->cpre_src) { h->cpre_src = nullptr; mark_bound(h, BOUND_STALE); }     // values change: compressed copy is stale
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = bind_tables(h, st)) return rc;
    CK(h, cudaMemsetAsync(Apre, 0, (size_t)h->nnz * sizeof(double), st));
    if (int rc = (h->gdim == 3 ? launch_linear<3>(h, x, x1, Apre, st) : launch_linear<2>(h, x, x1, Apre, st))) return rc;
    if (h->fo_n) {
        coo_add_matrix_kernel<<<grid_for(h->fo_n, 256, h->sm_count), 256, 0, st>>>(h->fo_pos, h->fo_vals, h->fo_n, Apre);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    // pair mode (tfsi_bind_matrix_pair): the frozen mass factors of fluid.py:33 at THIS state, so that the fluid kernel can
    // add that part of A_pre itself into a zero-filled matrix; and the sum of squares of what was written, by which
    // tfsi_compress_A_pre recognises the array
    for (Region& r : h->regions) {
        if (r.ncells == 0 || r.rc.kind != TFSI_KIND_FLUID || !r.rc.fluid_on) continue;
        const int ncta = region_ncta(h, r);
        const int64_t total = (r.ncells + ncta - 1) / ncta * ncta * h->tabJ.nq;
        if (!r.massJ) CK(h, cudaMalloc((void**)&r.massJ, (size_t)total * sizeof(double)));
        CellArgs a = cell_args(h, r, h->tabJ.nq, x, x1, nullptr);
        if (h->gdim == 3) mass_factor_kernel<3><<<grid_for(total, 128, h->sm_count), 128, 0, st>>>(a, ncta, r.rc.c1lin, r.massJ);
        else mass_factor_kernel<2><<<grid_for(total, 128, h->sm_count), 128, 0, st>>>(a, ncta, r.rc.c1lin, r.massJ);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    if (int rc = tfsi_sumsq(h, Apre, h->nnz, &h->apre_sumsq, stream)) return rc;
    h->apre_last = Apre;
    return 0;
}

int tfsi_assemble_jacobian(tfsi_handle* h, const double* x, const double* x1, const double* Apre, double* A, void* stream) {
    if (!h || !x || !x1 || !A) return fail(h, TFSI_ERR_ARG, "null argument");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = bind_tables(h, st)) return rc;
    return launch_jacobian(h, x, x1, Apre, A, st);
}

int tfsi_assemble_jacobian_phase(tfsi_handle* h, const double* x, const double* x1, const double* Apre, double* A,
                                 void* stream, int32_t phase) {
    if (!h || !A || phase < 1 || phase > 2 || (phase == 2 && (!x || !x1))) return fail(h, TFSI_ERR_ARG, "bad argument");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = bind_tables(h, st)) return rc;
    return launch_jacobian(h, x, x1, Apre, A, st, phase);
}

int tfsi_set_facet_operator(tfsi_handle* h, const int32_t* rows, const int32_t* cols, const double* vals, int64_t n) {
    if (!h || n < 0 || (n && (!rows || !cols || !vals))) return fail(h, TFSI_ERR_ARG, "bad argument");
    DeviceGuard dg(h);
    cudaFree(h->fo_rows); cudaFree(h->fo_cols); cudaFree(h->fo_pos); cudaFree(h->fo_vals);
    h->fo_rows = h->fo_cols = nullptr; h->fo_pos = nullptr; h->fo_vals = nullptr; h->fo_n = 0;
    if (n == 0) return 0;
    if (!h->colidx) return fail(h, TFSI_ERR_UNSUPPORTED, "the column array was released (tfsi_release_pattern)");
    if (int rc = upload(h, &h->fo_rows, rows, (size_t)n)) return rc;
    if (int rc = upload(h, &h->fo_cols, cols, (size_t)n)) return rc;
    if (int rc = upload(h, &h->fo_vals, vals, (size_t)n)) return rc;
    CK(h, cudaMalloc((void**)&h->fo_pos, (size_t)n * sizeof(int64_t)));
    int* derr = reinterpret_cast<int*>(h->scratch + 1);
    CK(h, cudaMemset(derr, 0, sizeof(int)));
    coo_positions_kernel<<<grid_for(n, 256, h->sm_count), 256>>>(h->fo_rows, h->fo_cols, n, h->rowptr, h->colidx, h->fo_pos, derr);
    int herr = 0;
    CK(h, cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) return fail(h, TFSI_ERR_PATTERN, "%d facet-operator entries are outside the CSR pattern", herr);
    h->fo_n = n;
    return 0;
}

int tfsi_set_facet_operator_prev(tfsi_handle* h, const int32_t* rows, const int32_t* cols, const double* vals, int64_t n) {
    if (!h || n < 0 || (n && (!rows || !cols || !vals))) return fail(h, TFSI_ERR_ARG, "bad argument");
    DeviceGuard dg(h);
    cudaFree(h->fp_rows); cudaFree(h->fp_cols); cudaFree(h->fp_vals);
    h->fp_rows = h->fp_cols = nullptr; h->fp_vals = nullptr; h->fp_n = 0;
    if (n == 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        if (rows[i] < 0 || rows[i] >= h->N || cols[i] < 0 || cols[i] >= h->N) return fail(h, TFSI_ERR_ARG, "facet-operator index out of range");
    if (int rc = upload(h, &h->fp_rows, rows, (size_t)n)) return rc;
    if (int rc = upload(h, &h->fp_cols, cols, (size_t)n)) return rc;
    if (int rc = upload(h, &h->fp_vals, vals, (size_t)n)) return rc;
    h->fp_n = n;
    return 0;
}

int tfsi_set_facet_load(tfsi_handle* h, const int32_t* dofs, const double* vals, int64_t n) {
    if (!h || n < 0 || (n && (!dofs || !vals))) return fail(h, TFSI_ERR_ARG, "bad argument");
    DeviceGuard dg(h);
    cudaFree(h->fl_dofs); cudaFree(h->fl_vals);
    h->fl_dofs = nullptr; h->fl_vals = nullptr; h->fl_n = 0;
    if (n == 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        if (dofs[i] < 0 || dofs[i] >= h->N) return fail(h, TFSI_ERR_ARG, "facet-load index out of range");
    if (int rc = upload(h, &h->fl_dofs, dofs, (size_t)n)) return rc;
    if (int rc = upload(h, &h->fl_vals, vals, (size_t)n)) return rc;
    h->fl_n = n;
    return 0;
}

int tfsi_set_facet_load_scale(tfsi_handle* h, double scale) {
    if (!h) return fail(h, TFSI_ERR_ARG, "null argument");
    h->fl_scale = scale;
    return 0;
}

int tfsi_invalidate_A_pre(tfsi_handle* h) {
    if (!h) return TFSI_ERR_ARG;
    DeviceGuard dg(h);
    CK(h, cudaDeviceSynchronize());
    h->cpre_src = nullptr;
    cudaFree(h->cpre_bitmap); cudaFree(h->cpre_base); cudaFree(h->cpre_val);
    h->cpre_bitmap = nullptr; h->cpre_base = nullptr; h->cpre_val = nullptr; h->cpre_nnz = 0; h->cpre_words = 0;
    mark_bound(h, BOUND_STALE);
    for (auto& b : h->bound) b.pending = 0;
    h->apre_own = 0;
    return 0;
}

int tfsi_compress_A_pre(tfsi_handle* h, const double* Apre, void* stream) {
    if (!h || !Apre) return fail(h, TFSI_ERR_ARG, "null argument");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    CK(h, cudaStreamSynchronize(st));
    if (int rc = tfsi_invalidate_A_pre(h)) return rc;
    const int64_t nwords = (h->nnz + 31) / 32, ngroups = (nwords + 31) / 32;
    if (ngroups + 1 > INT32_MAX) return fail(h, TFSI_ERR_UNSUPPORTED, "matrix too large for the bitmap scan");
    h->cpre_words = nwords;
    int64_t* cnt = nullptr;
    CK(h, cudaMalloc((void**)&h->cpre_bitmap, std::max<size_t>((size_t)nwords, 1) * sizeof(uint32_t)));
    CK(h, cudaMalloc((void**)&h->cpre_base, ((size_t)ngroups + 1) * sizeof(int64_t)));
    CK(h, cudaMalloc((void**)&cnt, ((size_t)ngroups + 1) * sizeof(int64_t)));
    CK(h, cudaMemsetAsync(cnt, 0, ((size_t)ngroups + 1) * sizeof(int64_t), st));
    bitmap_build_kernel<<<grid_for(ngroups * 32, 256, h->sm_count), 256, 0, st>>>(Apre, h->nnz, h->cpre_bitmap, nwords, cnt);
    void* tmp = nullptr; size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, h->cpre_base, (int)(ngroups + 1), st);
    CK(h, cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, h->cpre_base, (int)(ngroups + 1), st);
    int64_t total = 0;
    CK(h, cudaMemcpyAsync(&total, h->cpre_base + ngroups, sizeof total, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    cudaFree(tmp); cudaFree(cnt);
    h->cpre_nnz = total;
    CK(h, cudaMalloc((void**)&h->cpre_val, std::max<size_t>((size_t)total, 1) * sizeof(double)));
    bitmap_stream_kernel<0><<<grid_for(ngroups * 32, 256, h->sm_count), 256, 0, st>>>(
        h->cpre_bitmap, nwords, h->cpre_base, h->nnz, Apre, h->cpre_val, nullptr);
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(st));
    h->cpre_src = Apre;
    // is this the array tfsi_assemble_A_pre wrote (same address, same sum of squares up to the summation order)?
    double ss = 0.0;
    if (int rc = tfsi_sumsq(h, Apre, h->nnz, &ss, stream)) return rc;
    h->apre_own = Apre == h->apre_last && fabs(ss - h->apre_sumsq) <= 1e-9 * h->apre_sumsq;
    return 0;
}

int64_t tfsi_compressed_nnz(const tfsi_handle* h) { return h ? h->cpre_nnz : -1; }

namespace {
int build_touched_mask(tfsi_handle* h, cudaStream_t st) {
    if (h->tmask) return 0;
    const int64_t ngroups = (h->cpre_words + 31) / 32;
    // rows that J_nl scatters into: velocity + pressure rows of active fluid cells, velocity rows of active solid cells
    uint8_t* flag = nullptr;
    CK(h, cudaMalloc((void**)&flag, std::max<size_t>((size_t)h->N, 1)));
    CK(h, cudaMemsetAsync(flag, 0, (size_t)h->N, st));
    const int gn2 = h->gdim * h->n2;
    for (const Region& r : h->regions) {
        if (r.ncells == 0) continue;
        const bool fl = r.rc.kind == TFSI_KIND_FLUID;
        if (fl ? !r.rc.fluid_on : !r.rc.solid_on) continue;
        const int lo = gn2, hi = fl ? 2 * gn2 + h->n1 : 2 * gn2;
        mark_rows_kernel<<<grid_for(r.ncells * (hi - lo), 256, h->sm_count), 256, 0, st>>>(h->cell_dofs, r.cell_list, r.ncells, h->ndof, lo, hi, flag);
    }
    CK(h, cudaMalloc((void**)&h->tmask, std::max<size_t>((size_t)ngroups, 1) * sizeof(uint32_t)));
    CK(h, cudaMemsetAsync(h->tmask, 0, (size_t)ngroups * sizeof(uint32_t), st));
    touched_words_kernel<<<grid_for(h->N, 256, h->sm_count), 256, 0, st>>>(h->rowptr, flag, h->N, h->tmask);
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(h->scratch + 2);
    CK(h, cudaMemsetAsync(cnt, 0, sizeof *cnt, st));
    count_pairs_kernel<<<grid_for(ngroups, 256, h->sm_count), 256, 0, st>>>(h->tmask, ngroups, cnt);
    unsigned long long pairs = 0;
    CK(h, cudaMemcpyAsync(&pairs, cnt, sizeof pairs, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    h->row_flag = flag;
    h->bound_entries = std::min<int64_t>((int64_t)pairs * 64, h->nnz);
    return 0;
}

int bind_one(tfsi_handle* h, int slot, double* A, cudaStream_t st) {
    const int64_t ngroups = (h->cpre_words + 31) / 32;
    bitmap_stream_kernel<1><<<grid_for(ngroups * 32, 256, h->sm_count), 256, 0, st>>>(
        h->cpre_bitmap, h->cpre_words, h->cpre_base, h->nnz, nullptr, h->cpre_val, A, nullptr);
    h->launches++;
    CK(h, cudaGetLastError());
    h->bound[slot].A = A; h->bound[slot].state = BOUND_CLEAN; h->bound[slot].pending = 0;
    if (!h->bound[slot].ready) CK(h, cudaEventCreateWithFlags(&h->bound[slot].ready, cudaEventDisableTiming));
    return 0;
}
}  // namespace

int tfsi_bind_matrix(tfsi_handle* h, double* A, void* stream) {
    if (!h || !A) return fail(h, TFSI_ERR_ARG, "null argument");
    if (!h->cpre_bitmap) return fail(h, TFSI_ERR_ARG, "tfsi_bind_matrix needs a compressed A_pre (tfsi_compress_A_pre)");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = build_touched_mask(h, st)) return rc;
    CK(h, cudaDeviceSynchronize());                       // a background initialisation of an earlier pair may be in flight
    h->nbound = 1; h->bound[1].A = nullptr;
    return bind_one(h, 0, A, st);
}

namespace {
// pair mode set-up: (1) the zero-fill lists = the entries of every maximal run of rows J_nl scatters into, as 16-byte
// aligned pieces of <= 4 KB for the TMA bulk stores plus the odd entries at the ends of the runs (host pass over the
// row pointers), (2) the scatter ranks of the velocity columns of the solid regions (mass block).
int build_fused_lists(tfsi_handle* h, cudaStream_t st) {
    if (!h->fill_pieces) {
        std::vector<int64_t> rp((size_t)h->N + 1);
        std::vector<uint8_t> fl((size_t)h->N);
        CK(h, cudaMemcpyAsync(rp.data(), h->rowptr, rp.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CK(h, cudaMemcpyAsync(fl.data(), h->row_flag, fl.size(), cudaMemcpyDeviceToHost, st));
        CK(h, cudaStreamSynchronize(st));
        std::vector<FillPiece> pieces;
        std::vector<int64_t> scalars;
        const int64_t step = FILL_BLOCK / 8;
        int64_t r = 0;
        while (r < h->N) {
            if (!fl[(size_t)r]) { ++r; continue; }
            int64_t r1 = r;
            while (r1 < h->N && fl[(size_t)r1]) ++r1;
            int64_t e0 = rp[(size_t)r], e1 = rp[(size_t)r1];
            if (e0 < e1 && (e0 & 1)) scalars.push_back(e0++);
            if (e0 < e1 && (e1 & 1)) scalars.push_back(--e1);
            for (int64_t e = e0; e < e1; e += step) {
                FillPiece pc; pc.entry = e; pc.nbytes = (int32_t)(std::min(step, e1 - e) * 8); pc.pad_ = 0;
                pieces.push_back(pc);
            }
            r = r1;
        }
        h->n_fill = (int64_t)pieces.size();
        h->n_fill_scalar = (int64_t)scalars.size();
        CK(h, cudaMalloc((void**)&h->fill_pieces, std::max<size_t>(pieces.size(), 1) * sizeof(FillPiece)));
        CK(h, cudaMalloc((void**)&h->fill_scalar, std::max<size_t>(scalars.size(), 1) * sizeof(int64_t)));
        if (!pieces.empty()) CK(h, cudaMemcpy(h->fill_pieces, pieces.data(), pieces.size() * sizeof(FillPiece), cudaMemcpyHostToDevice));
        if (!scalars.empty()) CK(h, cudaMemcpy(h->fill_scalar, scalars.data(), scalars.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    int* derr = reinterpret_cast<int*>(h->scratch + 1);
    for (Region& r : h->regions) {
        if (r.ncells == 0 || r.rc.kind != TFSI_KIND_SOLID || !r.rc.solid_on || r.rank_blk_v) continue;
        if (!h->colidx) return 0;            // column array released: the pair stays correct through the foreground path
        const int ncta = h->gdim == 3 ? SolidCfg<3>::L::NC : SolidCfg<2>::L::NC, rks = ncta + 1, ncols = h->gdim * h->n2;
        const size_t rbytes = ((size_t)h->n2 * ncols * rks * sizeof(uint16_t) + 15) / 16 * 16;
        const int64_t nb = (r.ncells + ncta - 1) / ncta;
        r.rank_v_elems = rbytes / sizeof(uint16_t);
        CK(h, cudaMalloc((void**)&r.rank_blk_v, (size_t)nb * rbytes));
        CK(h, cudaMemsetAsync(r.rank_blk_v, 0, (size_t)nb * rbytes, st));
        CK(h, cudaMemsetAsync(derr, 0, sizeof(int), st));
        build_rank_blocked_kernel<<<grid_for(nb * h->n2 * ncols * ncta, 256, h->sm_count), 256, 0, st>>>(
            h->cell_dofs, h->rowptr, h->colidx, r.cell_list, r.ncells, h->ndof, h->n2, ncols, ncta, rks, r.rank_v_elems,
            r.rank_blk_v, derr, ncols);
        int herr = 0;
        CK(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(h, cudaStreamSynchronize(st));
        if (herr) return fail(h, TFSI_ERR_PATTERN, "%d velocity-column ranks of a solid region are missing from the CSR pattern", herr);
    }
    return 0;
}
}  // namespace

// Two persistent matrices used alternately (A0 in one assembly, A1 in the next): the rows of the idle one that the
// cell kernels scatter into are zero-filled DURING the assembly into the other, by the fluid kernel itself (TMA bulk
// stores between its tasks); the A_pre part of those rows — the mass blocks of fluid.py:33 and solid.py:41, registered
// facet operators — is added by the cell kernels of the next assembly into that matrix. This needs A_pre to be what
// tfsi_assemble_A_pre wrote (checked by address and sum of squares at tfsi_compress_A_pre); with any other A_pre, without
// an active fluid region, with "Form2" viscoelasticity or in the deterministic mode the pair falls back to the
// foreground masked initialisation. Apre: the array given to tfsi_compress_A_pre, or TFSI_APRE_COMPRESSED.
int tfsi_bind_matrix_pair(tfsi_handle* h, double* A0, double* A1, const double* Apre, void* stream) {
    if (!h || !A0 || !A1 || A0 == A1) return fail(h, TFSI_ERR_ARG, "two distinct matrices are needed");
    if (!h->cpre_bitmap) return fail(h, TFSI_ERR_ARG, "tfsi_bind_matrix_pair needs a compressed A_pre (tfsi_compress_A_pre)");
    if (!Apre || (Apre != TFSI_APRE_COMPRESSED && Apre != h->cpre_src))
        return fail(h, TFSI_ERR_ARG, "tfsi_bind_matrix_pair needs the A_pre that was given to tfsi_compress_A_pre (or TFSI_APRE_COMPRESSED)");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = build_touched_mask(h, st)) return rc;
    CK(h, cudaDeviceSynchronize());
    if (int rc = build_fused_lists(h, st)) return rc;
    h->nbound = 2;
    if (int rc = bind_one(h, 0, A0, st)) return rc;
    return bind_one(h, 1, A1, st);
}

// number of fluid launches that carried re-initialisation work for a matrix pair so far
int64_t tfsi_fused_launches(const tfsi_handle* h) { return h ? h->fused_launches : -1; }

int tfsi_unbind_matrix(tfsi_handle* h) {
    if (!h) return TFSI_ERR_ARG;
    DeviceGuard dg(h);
    CK(h, cudaDeviceSynchronize());
    h->nbound = 0;
    for (auto& b : h->bound) { b.A = nullptr; b.state = BOUND_STALE; b.pending = 0; }
    return 0;
}

int64_t tfsi_bound_init_entries(const tfsi_handle* h) { return h ? h->bound_entries : -1; }

int tfsi_assemble_residual(tfsi_handle* h, const double* x, const double* x1, double* b, void* stream) {
    if (!h || !x || !x1 || !b) return fail(h, TFSI_ERR_ARG, "null argument");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = bind_tables(h, st)) return rc;
    CK(h, cudaMemsetAsync(b, 0, (size_t)h->N * sizeof(double), st));
    if (int rc = launch_residual(h, x, x1, b, st)) return rc;
    return residual_facet_terms(h, x, x1, b, st);
}

int tfsi_ident_zeros(tfsi_handle* h, double* A, void* stream) {
    if (!h || !A) return fail(h, TFSI_ERR_ARG, "null argument");
    DeviceGuard dg(h);
    ident_zeros_kernel<<<grid_for(h->N, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(h->rowptr, h->diagpos, A, h->N, 3.0e-16);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_set_bc_dofs(tfsi_handle* h, const int32_t* bc, int64_t n) {
    if (!h || (n && !bc) || n < 0) return fail(h, TFSI_ERR_ARG, "bad argument");
    DeviceGuard dg(h);
    CK(h, cudaDeviceSynchronize());
    cudaFree(h->bc_dofs); h->bc_dofs = nullptr;
    cudaFree(h->hg); h->hg = nullptr;
    h->nbc = n;
    mark_bound(h, BOUND_STALE);    // unit rows of the previous set may sit in the bound matrices
    if (int rc = upload(h, &h->bc_dofs, bc, (size_t)n)) return rc;
    CK(h, cudaMalloc((void**)&h->hg, std::max<size_t>((size_t)n, 1) * sizeof(double)));
    return 0;
}

int tfsi_apply_bc_matrix(tfsi_handle* h, double* A, void* stream) {
    if (!h || !A) return fail(h, TFSI_ERR_ARG, "null argument");
    if (h->nbc == 0) return 0;
    DeviceGuard dg(h);
    bc_matrix_kernel<<<grid_for(h->nbc * 32, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(h->rowptr, h->diagpos, h->bc_dofs, h->nbc, A);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_apply_bc_rhs(tfsi_handle* h, double* b, const double* x, const double* g, int64_t n, void* stream) {
    if (!h || !b || !x) return fail(h, TFSI_ERR_ARG, "null argument");
    if (n != h->nbc) return fail(h, TFSI_ERR_ARG, "%lld BC values for a set of %lld dofs (tfsi_set_bc_dofs)", (long long)n, (long long)h->nbc);
    if (h->nbc == 0) return 0;
    if (!g) return fail(h, TFSI_ERR_ARG, "null BC values");
    DeviceGuard dg(h);
    bc_rhs_kernel<<<grid_for(h->nbc, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(h->bc_dofs, h->nbc, b, x, g);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_apply_bc_vec(tfsi_handle* h, double* x, const double* g, int64_t n, void* stream) {
    if (!h || !x) return fail(h, TFSI_ERR_ARG, "null argument");
    if (n != h->nbc) return fail(h, TFSI_ERR_ARG, "%lld BC values for a set of %lld dofs (tfsi_set_bc_dofs)", (long long)n, (long long)h->nbc);
    if (h->nbc == 0) return 0;
    if (!g) return fail(h, TFSI_ERR_ARG, "null BC values");
    DeviceGuard dg(h);
    bc_vec_kernel<<<grid_for(h->nbc, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(h->bc_dofs, h->nbc, x, g);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_csr_matvec(tfsi_handle* h, const double* A, const double* x, double* y, void* stream) {
    if (!h || !A || !x || !y) return fail(h, TFSI_ERR_ARG, "null argument");
    if (!h->colidx) return fail(h, TFSI_ERR_UNSUPPORTED, "the column array was released (tfsi_release_pattern)");
    DeviceGuard dg(h);
    csr_matvec_kernel<<<grid_for(h->N * 32, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(h->rowptr, h->colidx, A, x, y, h->N);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_sumsq(tfsi_handle* h, const double* v, int64_t n, double* out, void* stream) {
    if (!h || !v || !out) return fail(h, TFSI_ERR_ARG, "null argument");
    DeviceGuard dg(h);
    cudaStream_t st = (cudaStream_t)stream;
    CK(h, cudaMemsetAsync(h->scratch, 0, sizeof(double), st));
    sumsq_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(v, n, h->scratch);
    h->launches++;
    CK(h, cudaMemcpyAsync(out, h->scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    return 0;
}

int tfsi_newton_assemble_host(tfsi_handle* h, const double* x_host, const double* x1_host, const double* g_host,
                              const double* Apre, double* A, double* b_host, double* out_sumsq_b, int32_t recompute,
                              void* stream) {
    if (!h || !x_host || !b_host) return fail(h, TFSI_ERR_ARG, "null argument");
    if (recompute && !A) return fail(h, TFSI_ERR_ARG, "null A");
    if (!x1_host && !h->hx1_valid) return fail(h, TFSI_ERR_ARG, "x1_host is NULL but no previous time level was uploaded yet");
    DeviceGuard dg(h);
    const size_t nb = (size_t)h->N * sizeof(double);
    if (!h->hx) {
        CK(h, cudaMalloc((void**)&h->hx, nb));
        CK(h, cudaMalloc((void**)&h->hx1, nb));
        CK(h, cudaMalloc((void**)&h->hb, nb));
    }
    // Two streams: `io` carries the host<->device copies, `st` the kernels. The x-independent A := A_pre (if this
    // matrix needs one: not in pair mode) runs under the H2D copy of x; the residual is assembled BEFORE the Jacobian
    // so that its D2H copy and norm run under the Jacobian kernels. Both streams start after everything the caller
    // has queued on `stream` (A_pre, A are the caller's buffers).
    cudaStream_t st = h->main_stream, io = h->copy_stream;
    const bool bcs = h->nbc && g_host;
    CK(h, cudaEventRecord(h->ev_entry, (cudaStream_t)stream));
    CK(h, cudaStreamWaitEvent(st, h->ev_entry, 0));
    CK(h, cudaStreamWaitEvent(io, h->ev_entry, 0));
    // The iterate goes up in CHUNKS of consecutive dofs, and the residual batches of every region start as soon as the
    // chunks they read have arrived (prefix maximum of the largest dof per batch: cells and dofs of a locality-ordered
    // mesh run in parallel, so the residual kernel trails the copy by one chunk; on an arbitrary numbering everything
    // simply waits for the last chunk).
    if (bcs) CK(h, cudaMemcpyAsync(h->hg, g_host, (size_t)h->nbc * sizeof(double), cudaMemcpyHostToDevice, io));
    static const int max_chunks = getenv("TFSI_H2D_CHUNKS") ? atoi(getenv("TFSI_H2D_CHUNKS")) : 8;
    const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(max_chunks, h->N / (1 << 20)));
    auto chunk_end = [&](int k) -> int64_t { return h->N * (k + 1) / nchunk; };    // (doubling chunk sizes: no gain)
    if (h->res_nc && nchunk > 1) {
        for (Region& r : h->regions) {
            if (r.ncells == 0 || !r.res_pmax.empty() || !r.res_dofs_blk) continue;
            const int64_t nbr = (r.ncells + h->res_nc - 1) / h->res_nc;
            int32_t* dm = nullptr;
            CK(h, cudaMalloc((void**)&dm, (size_t)nbr * sizeof(int32_t)));
            batch_max_dof_kernel<<<grid_for(nbr, 128, h->sm_count), 128, 0, st>>>(r.res_dofs_blk, nbr, h->ndof * h->res_nc, dm);
            r.res_pmax.resize((size_t)nbr);
            CK(h, cudaMemcpyAsync(r.res_pmax.data(), dm, (size_t)nbr * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            CK(h, cudaStreamSynchronize(st));
            cudaFree(dm);
            for (size_t i = 1; i < r.res_pmax.size(); ++i) r.res_pmax[i] = std::max(r.res_pmax[i], r.res_pmax[i - 1]);
        }
    }
    while ((int)h->ev_chunk.size() < nchunk) {
        cudaEvent_t e;
        CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_chunk.push_back(e);
    }
    for (int k = 0; k < nchunk; ++k) {
        const int64_t lo = k ? chunk_end(k - 1) : 0, hi = chunk_end(k);
        CK(h, cudaMemcpyAsync(h->hx + lo, x_host + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, io));
        if (x1_host) CK(h, cudaMemcpyAsync(h->hx1 + lo, x1_host + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, io));
        CK(h, cudaEventRecord(h->ev_chunk[(size_t)k], io));
    }
    if (x1_host) h->hx1_valid = 1;
    CK(h, cudaEventRecord(h->ev_in, io));
    if (int rc = bind_tables(h, st)) return rc;
    if (recompute) {
        if (int rc = launch_jacobian(h, h->hx, h->hx1, Apre, A, st, 1)) return rc;
    }
    CK(h, cudaMemsetAsync(h->hb, 0, nb, st));
    {
        std::vector<std::pair<int64_t, int64_t>> ranges(h->regions.size(), std::make_pair((int64_t)0, (int64_t)0));
        for (int k = 0; k < nchunk; ++k) {
            const int64_t hi = chunk_end(k);
            bool any = false;
            for (size_t ri = 0; ri < h->regions.size(); ++ri) {
                const Region& r = h->regions[ri];
                const int64_t nbr = h->res_nc ? (r.ncells + h->res_nc - 1) / h->res_nc : 0;
                int64_t upto = nbr;
                if (k + 1 < nchunk)
                    upto = r.res_pmax.empty() ? 0 : (int64_t)(std::lower_bound(r.res_pmax.begin(), r.res_pmax.end(), (int32_t)hi) - r.res_pmax.begin());
                ranges[ri].first = ranges[ri].second;
                ranges[ri].second = std::max(ranges[ri].second, upto);
                any = any || ranges[ri].second > ranges[ri].first;
            }
            if (!any) continue;
            CK(h, cudaStreamWaitEvent(st, h->ev_chunk[(size_t)k], 0));
            h->res_ranges = &ranges;
            const int rc = launch_residual(h, h->hx, h->hx1, h->hb, st);
            h->res_ranges = nullptr;
            if (rc) return rc;
        }
    }
    CK(h, cudaStreamWaitEvent(st, h->ev_in, 0));
    if (int rc = residual_facet_terms(h, h->hx, h->hx1, h->hb, st)) return rc;
    if (bcs) if (int rc = tfsi_apply_bc_rhs(h, h->hb, h->hx, h->hg, h->nbc, st)) return rc;
    CK(h, cudaEventRecord(h->ev_b, st));
    CK(h, cudaStreamWaitEvent(io, h->ev_b, 0));
    CK(h, cudaMemcpyAsync(b_host, h->hb, nb, cudaMemcpyDeviceToHost, io));
    CK(h, cudaMemsetAsync(h->scratch, 0, sizeof(double), io));
    sumsq_kernel<<<grid_for(h->N, 256, h->sm_count), 256, 0, io>>>(h->hb, h->N, h->scratch);
    h->launches++;
    CK(h, cudaGetLastError());
    if (recompute) {
        if (int rc2 = launch_jacobian(h, h->hx, h->hx1, Apre, A, st, 2)) return rc2;
        if (int rc = tfsi_ident_zeros(h, A, st)) return rc;
        if (int rc = tfsi_apply_bc_matrix(h, A, st)) return rc;
    }
    double ss = 0;
    CK(h, cudaMemcpyAsync(&ss, h->scratch, sizeof(double), cudaMemcpyDeviceToHost, io));
    CK(h, cudaStreamSynchronize(io));
    CK(h, cudaStreamSynchronize(st));
    if (out_sumsq_b) *out_sumsq_b = ss;
    return 0;
}

int tfsi_halo_pack(tfsi_handle* h, const double* x, const int32_t* idx, int64_t n, double* buf, void* stream) {
    if (!h || !x || (n && (!idx || !buf))) return fail(h, TFSI_ERR_ARG, "null argument");
    if (n == 0) return 0;
    DeviceGuard dg(h);
    gather_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, idx, n, buf);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_halo_unpack(tfsi_handle* h, double* x, const int32_t* idx, int64_t n, const double* buf, void* stream) {
    if (!h || !x || (n && (!idx || !buf))) return fail(h, TFSI_ERR_ARG, "null argument");
    if (n == 0) return 0;
    DeviceGuard dg(h);
    scatter_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, idx, n, buf);
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

int tfsi_set_profiling(tfsi_handle* h, int32_t on) {
    if (!h) return TFSI_ERR_ARG;
    h->profiling = on;
    for (auto& v : h->ev) { for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } v.clear(); }
    return 0;
}

int tfsi_kernel_time(tfsi_handle* h, int32_t kind, double* total_ms, int64_t* count) {
    if (!h || kind < 0 || kind > 2) return fail(h, TFSI_ERR_ARG, "bad argument");
    DeviceGuard dg(h);
    double tot = 0;
    for (auto& p : h->ev[kind]) {
        CK(h, cudaEventSynchronize(p.second));
        float ms = 0;
        CK(h, cudaEventElapsedTime(&ms, p.first, p.second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = (int64_t)h->ev[kind].size();
    return 0;
}

int tfsi_expand_node_pattern(int64_t num_nodes, int64_t nnz_node, const int64_t* node_rowptr, const int32_t* node_col,
                             const int64_t* node_cum, const int64_t* node_base, const int32_t* node_ncomp,
                             const int64_t* rowptr, int32_t* colidx, void* stream) {
    if (!node_rowptr || !node_col || !node_cum || !node_base || !node_ncomp || !rowptr || !colidx)
        return fail(nullptr, TFSI_ERR_ARG, "null argument");
    int dev = 0, sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, TFSI_ERR_CUDA, "no CUDA device");
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    expand_node_pattern_kernel<<<grid_for(nnz_node, 256, sm), 256, 0, (cudaStream_t)stream>>>(
        num_nodes, node_rowptr, node_col, node_cum, node_base, node_ncomp, rowptr, colidx, nnz_node);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, TFSI_ERR_CUDA, "expand_node_pattern: %s", cudaGetErrorString(e));
    return 0;
}

#ifdef TFSI_PHASE_CLOCKS
// debugging aid (scripts/phase_clocks.py): reads and clears the per-phase cycle sums of the fluid Jacobian kernel
extern "C" int tfsi_debug_phase_clocks(unsigned long long* out16) {
    cudaDeviceSynchronize();
    static unsigned long long z[256 * 16];
    if (cudaMemcpyFromSymbol(z, tfsi::g_phase_clk, sizeof(z)) != cudaSuccess) return -1;
    for (int k = 0; k < 16; ++k) { out16[k] = 0; for (int s = 0; s < 256; ++s) out16[k] += z[s * 16 + k]; }
    for (auto& v : z) v = 0;
    cudaMemcpyToSymbol(tfsi::g_phase_clk, z, sizeof(z));
    return 0;
}
#endif

int tfsi_probe_fp64_peak(double* out_tflops, double* out_ms) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, TFSI_ERR_CUDA, "no CUDA device");
    int dev = 0, sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sm * 8, threads = 256, iters = 4096;
    double* d = nullptr;
    if (cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return fail(nullptr, TFSI_ERR_CUDA, "probe alloc");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_probe_kernel<<<blocks, threads>>>(d, iters, 1.0);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_probe_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return fail(nullptr, TFSI_ERR_CUDA, "probe: %s", cudaGetErrorString(e));
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (out_tflops) *out_tflops = flops / (best * 1e-3) / 1e12;
    if (out_ms) *out_ms = best;
    return 0;
}

}  // extern "C"
