// EXPERIMENT RECORD (not compiled into the library): a persistent, software-pipelined variant of jac_fluid_kernel<3,14>.
// It was written to overlap the FP64-poor phases of a batch (gather, phase 0) with the quadrature loops of the previous
// batch; results are bit-for-bit those of the shipped kernel (smoke + bench parity), but it is SLOWER on B200:
//   2 M-tet tube, fluid launch: 13.1 ms (shipped, single matrix) vs 35.3 -> 29.6 -> 23.8 ms for this kernel
//   (first version -> register spills removed -> REDs restored), two queues bound to SM sub-partitions: 31.7 ms.
// What the ncu captures (gpurun_out/r4_pipe*.ncu-rep, summarised in DESIGN.md section 4) showed:
//   * any real fence (__threadfence_block, fence.acq_rel.cta) BEHIND the scatter in the persistent loop makes ptxas
//     emit ATOMG instead of REDG for every scatter add: the warp then waits for 50 L2 round trips per task;
//   * with the hand-over fence-free, instruction fetch is the limiter (1.8 no-instruction stalls per issue vs 0.06):
//     gather + phase 0 + both half-block loops + their scatters + the small tasks are ~62 KB of code that is live on the
//     SM at all times, against ~20 KB per phase in the lock-step kernel.
// To compile it, paste it into tfsi_kernels.cuh behind jac_fluid_kernel and launch one CTA per SM with
// FluidPipe<3,14>::BYTES(_WITH_ZERO) of shared memory on batches of 10 cells (rank/dof/geometry/mass blocks built with
// nc_cta = 10, rks = 11).
// -------------------------------------------------------------------------------------------
// PIPELINED fluid Jacobian (3-D, whole quadrature rule in shared memory; the headline configuration).
// The kernel above runs one batch of 30 cells per CTA in lock-step phases (gather | phase 0 | tasks), and with one CTA per
// SM nothing overlaps the FP64-poor phases. Here ONE PERSISTENT CTA per SM streams batches of CPW = 10 cells (one warp's
// worth of (cell, b) lanes) through THREE shared-memory slots, and all work is warp tasks drawn from ONE ordered queue:
//   GATHER(m)   one warp: dofs -> x, x1, row starts, geometry of batch m into its slot
//   P0(m, w)    phase 0 of 32 (cell, q) items of batch m (fluid_base); the last one to finish starts the bulk copy
//               of the batch's scatter ranks into the (then dead) coefficient area of the slot
//   BIG(m, r)   one P2xP2 half block column set (field, a, half) x lane = (cell, b): quadrature loop + scatter
//   SMALL(m, k) psi x dp, gamma x dd, gamma x dv
// Queue order of iteration n: BIG/SMALL of batch n with P0 of batch n + 1 and GATHER of batch n + 2 interleaved, so the
// FP64-poor tasks run in the shadow of the quadrature loops of the other warps. A task's prerequisites are always
// EARLIER in the queue (a warp that draws a task whose inputs are not ready yet spins on a shared-memory counter; the
// warps holding the earlier tasks never wait for later ones, hence no deadlock). Slot hand-over: per-slot monotone
// counters (gathers, phase-0 tasks, work tasks done) + one mbarrier per slot for the rank copy (parity = use & 1).
template <int G, int NQ> struct FluidPipe {
    typedef LB<G, 1> L;
    static constexpr int N2 = Dim<G>::N2, N1 = Dim<G>::N1;
    static constexpr int NCS = L::NC, RKS = L::RKS, NSLOT = 3, WARPS = JAC_WARPS;
    static constexpr int NCOEF = FluidJac<G>::NCOEF, NROWS = FluidJac<G>::NROWS;
    static constexpr int STRIDE = base_stride(FluidBase<G>::COUNT, G);
    static constexpr int NP0 = (NQ * NCS + 31) / 32;
    static constexpr int NBIG = 4 * G, NSMALL = 3, NWORK = NBIG + NSMALL;
    static constexpr int NSEQ = NWORK + NP0 + 1;
    static constexpr size_t BASE_BYTES = sizeof(double) * NQ * NCS * STRIDE;
    static constexpr size_t COEF_BYTES = 2 * sizeof(double) * NCOEF * NCS;
    static constexpr size_t RANK_BYTES = (((size_t)N2 * NCOEF * RKS * sizeof(uint16_t)) + 15) / 16 * 16;
    static constexpr size_t UR_BYTES = COEF_BYTES > RANK_BYTES ? COEF_BYTES : RANK_BYTES;
    static constexpr size_t ROWST_BYTES = sizeof(int64_t) * NROWS * NCS;
    static constexpr size_t GEO_BYTES = (sizeof(double) * (G * G + 1) * NCS + 15) / 16 * 16;
    static constexpr size_t OFF_U = BASE_BYTES, OFF_ROWST = OFF_U + UR_BYTES, OFF_GEO = OFF_ROWST + ROWST_BYTES;
    static constexpr size_t SLOT_BYTES = OFF_GEO + GEO_BYTES;
    static constexpr size_t OFF_SYNC = NSLOT * SLOT_BYTES;       // ints: [0] queue A, [1] fill, [2] fix-up, [3] queue B, [4+s] gathers,
    static constexpr size_t OFF_MBAR = OFF_SYNC + 128;           //       [8+s] phase-0 tasks, [12+s] work tasks done,
    static constexpr size_t OFF_ZERO = OFF_SYNC + 160;           //       [16+w] task descriptor of warp w
    static constexpr size_t BYTES = OFF_ZERO, BYTES_WITH_ZERO = OFF_ZERO + FILL_BLOCK;
    static constexpr size_t DOFS_BLK = (size_t)NCOEF * NCS;
    static_assert(BASE_BYTES % 16 == 0 && UR_BYTES % 16 == 0 && ROWST_BYTES % 16 == 0, "slot areas are 16-byte aligned");
    static_assert(BYTES_WITH_ZERO <= 227 * 1024, "three slots + zero block must fit the shared memory of an SM");
};
// queue entries of one iteration: 0..11 BIG r, 12..14 SMALL, 16..20 P0 w (of the NEXT batch), 32 GATHER (next but one).
// TWO queues, bound to the SM sub-partitions: the warps of sub-partitions 0/1 (warp & 3 < 2) draw from queue A, which
// holds the six big tasks of column half 0, the others from queue B with the six of half 1 — the three warps that share
// a scheduler and its 6 KB L0 instruction cache then run the SAME 6 KB quadrature loop (a mix of the two half-block
// variants and the cold tasks on one scheduler starves the instruction fetch: 1.8 no-instruction stalls per issue).
// Prerequisites of a task of iteration n are tasks of iterations < n in either queue, so the no-deadlock argument holds.
constexpr int PIPE_NSEQ_A = 10, PIPE_NSEQ_B = 11;
__constant__ unsigned char c_pipe_seq[2][12] = {{0, 1, 2, 16, 3, 4, 17, 5, 32, 12, 0, 0},
                                                {6, 7, 8, 18, 9, 19, 10, 20, 11, 13, 14, 0}};

#ifndef PIPE_FENCE
#define PIPE_FENCE() asm volatile("" ::: "memory")
#endif
__device__ __forceinline__ void pipe_wait_ge(const int* p, int target) {
    const volatile int* vp = p;
    int spins = 0;
    while (*vp < target) {
        __nanosleep(64);
        if (++spins > (1 << 23)) __trap();       // a broken hand-over becomes an error instead of a hang
    }
    PIPE_FENCE();
}
__device__ __forceinline__ void pipe_signal(int* p, int lane) {
    PIPE_FENCE();
    __syncwarp();
    if (lane == 0) atomicAdd(p, 1);
}

// fused_init_batch handed out only while the hand-out counters are below the gates (a race past a gate only runs ahead
// of the pace by one batch; what is handed out is always cut at the CTA's share)
__device__ __forceinline__ void fused_init_batch_paced(const CellArgs& a, int* counter, const double* zbuf, int64_t fill0, int nfill,
                                                       int gate_fill, int64_t sc0, int nsc, int gate_sc, int lane)
{
    volatile int* vc = counter;
    if (vc[1] < nfill && vc[1] < gate_fill) {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter + 1, FILL_LANES);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < FILL_LANES && base + lane < nfill) {
            const FillPiece pc = a.fill_pieces[fill0 + base + lane];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(a.fill_dst + pc.entry), "r"(smem_u32(zbuf)), "r"(pc.nbytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (vc[2] < nsc && vc[2] < gate_sc) {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter + 2, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + lane < nsc) a.fill_dst[a.fill_scalar[sc0 + base + lane]] = 0.0;
    }
}

// The pipelined kernel lives on the edge of its register allocation like the one above: the quadrature loop of a big
// task keeps 100 accumulators and ptxas spills whatever else lives across it (a local-memory load per point, which
// misses the small L1 left beside 228 KB of shared memory). So NOTHING about the task lives in registers across the loop:
// the warp-uniform part is a packed descriptor in shared memory (one volatile LDS per point and one behind the loop),
// the per-lane part (cell, trial component) is recomputed from %tid.
//   descriptor bits: 0-1 slot, 2 parity of the slot's rank copy, 4 field (d | v), 5-6 test component a,
//                    8-11 number of valid cells of the batch (1..10)
__device__ __forceinline__ int pipe_tid() { int t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); return t; }
__device__ __forceinline__ int pipe_desc(const unsigned char* smem, size_t off_sync, int tid) {
    return reinterpret_cast<const volatile int*>(smem + off_sync)[16 + (tid >> 5)];
}

template <int G, int NQ, int HALF>
__device__ __forceinline__ void fluid_big_task_pipe(const CellArgs& a, unsigned char* smem)
{
    typedef FluidPipe<G, NQ> P;
    constexpr int N2 = Dim<G>::N2;
    constexpr int NJ = HalfCols<G, HALF>::NJ;
    constexpr int QSTR = P::NCS * P::STRIDE;
    double Z[N2][NJ];
#pragma unroll
    for (int i = 0; i < N2; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) Z[i][j] = 0.0;
    // ONE loop-carried vector register: bits 0-1 slot, 4 field, 5-6 a (from the descriptor), 7-8 b, 9-12 cell
    int dq;
    {
        const int tid = pipe_tid(), desc = pipe_desc(smem, P::OFF_SYNC, tid);
        const int l3 = (tid & 31) / G, c = l3 < P::NCS ? l3 : P::NCS - 1, cb = (tid & 31) - l3 * G;
        dq = (desc & 0x73) | (cb << 7) | (c << 9);
    }
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {            // (q stays a separate, warp-UNIFORM counter: the tabulations are uniform loads)
        asm volatile("" : "+r"(dq));          // (keeps the fields packed: unpacked they would be hoisted into registers)
        const double* basec = reinterpret_cast<const double*>(smem)
                              + ((dq & 3) * (int)(P::SLOT_BYTES / 8) + ((dq >> 9) & 15) * P::STRIDE);
        auto bv = [&](int item) { return basec[q * QSTR + item]; };
        double C[G + 1][G + 1];
        const int ca = (dq >> 5) & 3, cb = (dq >> 7) & 3;
        if ((dq & 16) == 0) fluid_C_psi_dd<G>(bv, a.rc, ca, cb, C);
        else fluid_C_psi_dv<G>(bv, a.rc, ca, cb, C);
        block_p2p2<G, HALF, 7>(c_tabJ, q, C, Z);
    }
    const int tid = pipe_tid(), desc = pipe_desc(smem, P::OFF_SYNC, tid);
    const int l3 = (tid & 31) / G, cb = (tid & 31) - l3 * G;
    const int FLD = (desc >> 4) & 1, ca = (desc >> 5) & 3;
    unsigned char* slot = smem + (size_t)(desc & 3) * P::SLOT_BYTES;
    mbar_wait(reinterpret_cast<uint64_t*>(smem + P::OFF_MBAR) + (desc & 3), (uint32_t)((desc >> 2) & 1));
    if (l3 < ((desc >> 8) & 15))
        scatter_p2p2<G, typename P::L, P::NCOEF, HALF>(a, reinterpret_cast<const int64_t*>(slot + P::OFF_ROWST),
                                                       reinterpret_cast<const uint16_t*>(slot + P::OFF_U), l3, ca * N2,
                                                       FLD * G * N2 + cb * N2, Z);
}

template <int G, int NQ>
__global__ void __maxnreg__(TFSI_FLUID_MAXREG) jac_fluid_pipe_kernel(CellArgs a)
{
    typedef FluidPipe<G, NQ> P;
    static_assert(P::NSEQ == PIPE_NSEQ_A + PIPE_NSEQ_B && P::NP0 == 5 && P::NBIG == 12, "c_pipe_seq is written for the 3-D 14-point rule");
    constexpr int N2 = P::N2, N1 = P::N1, NCS = P::NCS, RKS = P::RKS, STRIDE = P::STRIDE, NCOEF = P::NCOEF;
    constexpr int THREADS = P::WARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* sync = reinterpret_cast<int*>(smem_raw + P::OFF_SYNC);
    int* counter = sync;                       // [0] queue A, [1] fill pieces, [2] fix-up entries (fused_init_batch), [3] queue B
    int* gath_done = sync + 4;
    int* p0_done = sync + 8;
    int* work_done = sync + 12;
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem_raw + P::OFF_MBAR);
    double* zbuf = reinterpret_cast<double*>(smem_raw + P::OFF_ZERO);
    const bool fused = a.fill_dst != nullptr;
    if (threadIdx.x < 16) sync[threadIdx.x] = 0;
    if (threadIdx.x == 0)
        for (int s = 0; s < P::NSLOT; ++s) mbar_init(mbars + s, 1);
    if (fused)
        for (int i = threadIdx.x; i < FILL_BLOCK / 8; i += THREADS) zbuf[i] = 0.0;
    __syncthreads();
    // this CTA's share of the zero-fill pieces / fix-up entries (recomputed where needed: nothing of it stays in registers
    // across a task)
    auto share = [&](int64_t& fill0, int& nfill, int64_t& sc0, int& nsc) {
        fill0 = a.n_fill * blockIdx.x / gridDim.x;
        nfill = (int)(a.n_fill * (blockIdx.x + 1) / gridDim.x - fill0);
        sc0 = a.n_fill_scalar * blockIdx.x / gridDim.x;
        nsc = (int)(a.n_fill_scalar * (blockIdx.x + 1) / gridDim.x - sc0);
    };

    for (;;) {
        const int lane = pipe_tid() & 31;
        int64_t ncr = a.ncell_region;
        asm volatile("" : "+l"(ncr));                  // (recomputed per task instead of living in registers)
        const int64_t nb = (ncr + NCS - 1) / NCS;                                              // batches of the region
#ifdef PIPE_CONTIG
        const int64_t gb0 = nb * blockIdx.x / gridDim.x;                                       // a contiguous share per CTA
        const int nloc = (int)(nb * (blockIdx.x + 1) / gridDim.x - gb0);
#else
        const int nloc = nb > blockIdx.x ? (int)((nb - 1 - blockIdx.x) / gridDim.x) + 1 : 0;   // ... of this CTA
#endif
        const int qb = ((pipe_tid() >> 5) & 3) >> 1;   // queue of this warp's SM sub-partition
        const int nseq = qb ? PIPE_NSEQ_B : PIPE_NSEQ_A;
        const int total = (nloc + 2) * nseq;           // iterations n = -2 .. nloc - 1 (the first two only fill the pipe)
        const int cw = (lane / G < NCS) ? lane / G : NCS - 1;
        const int lcomp = lane % G;
        const bool lane_on = lane < NCS * G;
        int T = 0;
        if (lane == 0) T = atomicAdd(counter + 3 * qb, 1);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= total) break;
        if (fused) {      // the CTA's share of the zero-fill, paced with the queue
            int64_t fill0, sc0;
            int nfill, nsc;
            share(fill0, nfill, sc0, nsc);
            const int lf = (int)((int64_t)nfill * (T + 1) / total) + 1;
            const int ls = (int)((int64_t)nsc * (T + 1) / total) + 1;
            fused_init_batch_paced(a, counter, zbuf, fill0, nfill, lf, sc0, nsc, ls, lane);
        }
        const int n = T / nseq - 2;
        const int code = c_pipe_seq[qb][T % nseq];
        // the batch this task belongs to
        const int m = code == 32 ? n + 2 : (code >= 16 ? n + 1 : n);
        if (m < 0 || m >= nloc) continue;
        const int s = m % P::NSLOT, use = m / P::NSLOT;
#ifdef PIPE_CONTIG
        const int64_t gb = gb0 + m;
#else
        const int64_t gb = (int64_t)m * gridDim.x + blockIdx.x;
#endif
        unsigned char* slot = smem_raw + (size_t)s * P::SLOT_BYTES;
        double* base = reinterpret_cast<double*>(slot);
        double* coef = reinterpret_cast<double*>(slot + P::OFF_U);
        double* coef1 = coef + NCOEF * NCS;
        uint16_t* srank = reinterpret_cast<uint16_t*>(slot + P::OFF_U);
        int64_t* rowst = reinterpret_cast<int64_t*>(slot + P::OFF_ROWST);
        double* sgeo = reinterpret_cast<double*>(slot + P::OFF_GEO);

        if (code == 32) {
            // ---- GATHER: the slot is free once every work task of its previous batch has scattered
            pipe_wait_ge(work_done + s, P::NWORK * use);
            const int32_t* dblk = a.dofs_blk + (size_t)gb * P::DOFS_BLK;
            constexpr int NIT = (NCOEF * NCS + 31) / 32, HIT = (NIT + 1) / 2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int dofs[HIT];
                double xv[HIT], x1v[HIT];
                int64_t rp[HIT];
#pragma unroll
                for (int it = 0; it < HIT; ++it) {
                    const int idx = lane + (h * HIT + it) * 32;
                    dofs[it] = (idx < NCOEF * NCS) ? __ldcg(dblk + idx) : 0;
                }
#pragma unroll
                for (int it = 0; it < HIT; ++it) {
                    const int idx = lane + (h * HIT + it) * 32;
                    xv[it] = __ldcg(a.x + dofs[it]);              // (L2 only: the small L1 is left to the register spills)
                    x1v[it] = __ldcg(a.x1 + dofs[it]);
                    rp[it] = (idx >= G * N2 * NCS) ? __ldcg(a.rowptr + dofs[it]) : 0;
                }
#pragma unroll
                for (int it = 0; it < HIT; ++it) {
                    const int idx = lane + (h * HIT + it) * 32;
                    if (idx < NCOEF * NCS) {
                        coef[idx] = xv[it];
                        coef1[idx] = x1v[it];
                        if (idx >= G * N2 * NCS) rowst[idx - G * N2 * NCS] = rp[it];
                    }
                }
            }
            for (int i = lane; i < (G * G + 1) * NCS; i += 32) sgeo[i] = __ldcg(a.geo_blk + (size_t)gb * ((G * G + 1) * NCS) + i);
            pipe_signal(gath_done + s, lane);
        } else if (code >= 16) {
            // ---- P0: one (cell, q) item per lane
            const int t = (code - 16) * 32 + lane;
            const bool on = t < NQ * NCS;
            double mj = 0.0;
            if (a.massJ && on) mj = __ldcg(a.massJ + (size_t)gb * (c_tabJ.nq * NCS) + t);
            pipe_wait_ge(gath_done + s, use + 1);
            if (on) {
                const int c = t % NCS, q = t / NCS;
                double K[G][G];
#pragma unroll
                for (int u = 0; u < G; ++u)
#pragma unroll
                    for (int v = 0; v < G; ++v) K[u][v] = sgeo[(u * G + v) * NCS + c];
                const double adet = sgeo[G * G * NCS + c];
                auto cf = [&](int k) { return coef[k * NCS + c]; };
                auto cf1 = [&](int k) { return coef1[k * NCS + c]; };
                double* dstb = base + ((size_t)q * NCS + c) * STRIDE;
                auto out = [&](int item, double v) { dstb[item] = v; };
                fluid_base<G>(c_tabJ, q, K, adet, a.rc, cf, cf1, out, mj);
            }
            PIPE_FENCE();
            __syncwarp();
            if (lane == 0 && atomicAdd(p0_done + s, 1) + 1 == P::NP0 * (use + 1)) {
                // coefficients are dead: the ranks of the batch land in their place while the first tasks compute
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbars + s, (uint32_t)P::RANK_BYTES);
                bulk_g2s(srank, a.rank_blk + (size_t)gb * (P::RANK_BYTES / sizeof(uint16_t)), (uint32_t)P::RANK_BYTES, mbars + s);
            }
        } else {
            // ---- work tasks of batch m
            const int c = cw;
            const bool valid = lane_on && (gb * NCS + c < a.ncell_region);
            const double* basec = base + (size_t)c * STRIDE;
            const uint32_t parity = (uint32_t)(use & 1);
            {
                const int64_t left = a.ncell_region - gb * NCS;
                const int desc = s | (int)(parity << 2) | ((left < NCS ? (int)left : NCS) << 8)
                                 | (code < P::NBIG ? (((code / G) & 1) << 4) | ((code % G) << 5) : 0);
                if (lane == 0) sync[16 + (threadIdx.x >> 5)] = desc;
                __syncwarp();
            }
            pipe_wait_ge(p0_done + s, P::NP0 * (use + 1));
            if (code < P::NBIG) {
                const int half = code / (2 * G);
                if (half == 0) fluid_big_task_pipe<G, NQ, 0>(a, smem_raw);
                else fluid_big_task_pipe<G, NQ, 1>(a, smem_raw);
            } else {
                const int kind = code - P::NBIG;      // 0: psi x dp (lane comp = a), 1: gamma x dd, 2: gamma x dv
                constexpr int QSTR = NCS * STRIDE;
                if (kind == 0) {
                    const int ca = lcomp;
                    double Z[N2][N1];
#pragma unroll
                    for (int i = 0; i < N2; ++i)
#pragma unroll
                        for (int k = 0; k < N1; ++k) Z[i][k] = 0.0;
#pragma unroll SMALL_UNROLL
                    for (int q = 0; q < NQ; ++q) {
                        auto bv = [&](int item) { return basec[q * QSTR + item]; };
                        double cc[G];
                        fluid_c_psi_dp<G>(bv, ca, cc);
                        block_gradp2_p1<G>(c_tabJ, q, cc, Z);
                    }
                    mbar_wait(mbars + s, parity);
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < N2; ++i) {
                            double* dst = a.out + rowst[(ca * N2 + i) * NCS + c];
                            const uint16_t* rk = srank + ((i * NCOEF) + 2 * G * N2) * RKS + c;
#pragma unroll
                            for (int k = 0; k < N1; ++k) atomicAdd(dst + rk[k * RKS], Z[i][k]);
                        }
                    }
                } else {
                    const int fld = kind - 1, cb = lcomp;
                    double Z[N1][N2];
#pragma unroll
                    for (int k = 0; k < N1; ++k)
#pragma unroll
                        for (int j = 0; j < N2; ++j) Z[k][j] = 0.0;
#pragma unroll SMALL_UNROLL
                    for (int q = 0; q < NQ; ++q) {
                        auto bv = [&](int item) { return basec[q * QSTR + item]; };
                        double cc[G];
                        if (fld == 0) fluid_c_gamma_dd<G>(bv, cb, cc);
                        else fluid_c_gamma_dv<G>(bv, cb, cc);
                        block_p1_gradp2<G>(c_tabJ, q, cc, Z);
                    }
                    mbar_wait(mbars + s, parity);
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < N1; ++k) {
                            double* dst = a.out + rowst[(G * N2 + k) * NCS + c];
                            const uint16_t* rk = srank + ((k * NCOEF) + fld * G * N2 + cb * N2) * RKS + c;
#pragma unroll
                            for (int j = 0; j < N2; ++j) atomicAdd(dst + rk[j * RKS], Z[k][j]);
                        }
                    }
                }
            }
            {   // (no fence: the task only READ the slot, and a fence behind the scatter turns its REDs into ATOMs)
                const int tid = pipe_tid();
                __all_sync(0xffffffffu, 1);
                if ((tid & 31) == 0) atomicAdd(work_done + (pipe_desc(smem_raw, P::OFF_SYNC, tid) & 3), 1);
            }
        }
    }
    const int lane = pipe_tid() & 31;
    if (fused) {      // whatever is left of the CTA's share, then wait for the bulk stores of this thread
        int64_t fill0, sc0;
        int nfill, nsc;
        share(fill0, nfill, sc0, nsc);
        volatile int* vc = counter;
        while (vc[1] < nfill || vc[2] < nsc) fused_init_batch(a, counter, zbuf, fill0, nfill, sc0, nsc, lane);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

