// Block-Jacobi preconditioned conjugate gradient on the float-stored 3x3-BCSR Hessian.
//
// Replaces bsm::solve_pcg (bsm/solve_pcg.h:83-232), BlockedSparseMatrix::{prepare_preconditioning,
// apply_preconditioning, _spmxv} (bsm/BlockedSparseMatrix.h:1147-1361, 988-1138) and the vector kernels of
// bsm/bsm_vector_ops.h.  Same stopping rules (zero RHS, error < abs_tol, error/error_0 < rel_tol, p^T A p <= 0 ->
// "indefinite", max_iter), same float-storage / double-accumulate arithmetic, same block-Jacobi preconditioner.
//
// The ENTIRE solve is ONE persistent cooperative kernel, one 1024-thread CTA per SM:
//   * CTA c owns a contiguous range of block rows holding ~nnzb / #SM blocks and copies ITS SLICE OF THE MATRIX (values,
//     columns, row pointers) and of the vectors INTO SHARED MEMORY ONCE: at the 200k-tet scene the 21.7 MB matrix is
//     spread over the 148 x 227 KB of shared memory of the chip and is never read from L2 / HBM again during the solve.
//     Slices that do not fit (66 k-node cloth, million-tet scenes) are read from L2 every iteration instead (see the MODE
//     comment at pcg_body: typed read-only-path loads with a smaller carve-out, or generic pointers when nothing fits).
//   * Rows are cut into slices of equal  blocks + ROW_COST * rows;  rows with more than LONG_ROW blocks (a rigid body in
//     contact with hundreds of nodes) are cut into segments that queue behind the ordinary rows as further 4-lane groups.
//   * The only vector that crosses CTAs is the preconditioned residual u = M^-1 r (the SpMV operand).  Every iteration
//     each CTA pulls the WINDOW of u around its own rows into shared memory with one TMA bulk copy
//     (cp.async.bulk.shared.global + mbarrier) -- the band of a mesh-ordered matrix -- and gathers from shared memory;
//     the few columns outside the window (rigid bodies, hex-centre nodes, far contacts) are gathered from a padded copy of
//     u in L2 with one 256-bit load each.  The copy of the next product's window is issued right after the barrier and
//     lands under the dot-product reduction.
//   * The recurrence is the Chronopoulos-Gear form of PCG (same iterates as the textbook form in exact arithmetic):
//         p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; w = A u
//         gamma' = r.u ; delta = w.u ; beta = gamma'/gamma ; alpha = gamma' / (delta - beta gamma'/alpha)
//     s = A p is carried by recurrence, so an iteration needs TWO grid-wide barriers (u visible -> SpMV -> dot
//     products visible) instead of three, and p^T A p of the next iteration is the denominator of alpha (its sign is the
//     reference's indefiniteness test).  Phases are separated by a grid barrier (one atomic counter, acquire spin).
//   * Dot products go through per-CTA partial sums that every CTA re-reduces in the same fixed order, so alpha, beta,
//     the error and every stopping decision are computed redundantly but IDENTICALLY everywhere (no broadcast; bitwise
//     reproducible run to run -- the reference's are thread-count dependent, bsm/ParallelNumber.h:39-47).
// The host launches once and reads one record.
#include "internal.h"
#include <algorithm>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>

namespace sb {

int bcsr_view(sb_context* ctx, int* nbr, size_t* nnzb, const unsigned long long** rows, const int32_t** cols, const float** vals);

constexpr int PCG_THREADS = 1024;     // one CTA per SM
constexpr int PCG_MAX_BLOCKS = 1280;  // upper bound of the (virtual) cooperative grid: 8 ranks x 148 CTAs (partial-sum arrays)
constexpr int DIST_MAX_WORLD = 8;     // ranks of one distributed solve (one NVSwitch domain)
constexpr int LANES_PER_ROW = 4;      // lanes cooperating on one block row of the SpMV
constexpr int LONG_ROW = 96;          // rows with more blocks (rigid bodies in contact with many nodes) are cut into segments
constexpr int PCG_TILE_BLOCKS = 768;                // blocks per streamed tile (multiple of 4: 16-byte granularity of the bulk copies)
constexpr unsigned PCG_TILE_BYTES = PCG_TILE_BLOCKS * 40u;   // columns (4 B) then values (36 B) of the tile's blocks
constexpr unsigned PCG_STREAM_SMEM = 112 * 1024;   // dynamic shared memory of a solve whose matrix streams (leaves ~96 KB of L1)
constexpr int MAX_SEGS = 128;         // segments of the long rows of one CTA
constexpr int MIN_SEG = 16;           // blocks per segment (longer when a CTA holds more than MAX_SEGS * MIN_SEG long-row blocks)
constexpr int MAX_LONG_ROWS = 32;     // per CTA; further long rows fall back to the 4-lane path

struct PcgResult {
    double du_dot_grad, du_inf, error, bb;
    int it;             // completed iterations
    int done;           // 1 converged, 2 indefinite, 3 max iterations, 4 aborted (a peer rank never arrived at a barrier)
    int found_indef;
    int pad;
    unsigned long long t_start, t_loaded, t_loop, t_end;   // %globaltimer (ns) of CTA 0: kernel entry, slices resident, first iteration, exit
    long long c_spmv, c_bar, c_red, c_vec, c_win;          // clock64 cycles of CTA 0 inside the loop (only when instrumented)
    unsigned long long barriers;                           // grid barriers this solve executed (the distributed counter never resets)
    unsigned long long ll_uses;                            // flagged all-reduces this solve executed (their flags never repeat either)
    unsigned long long seq;                                // written LAST (after a system fence): the host polls it in pinned memory
};

// Distributed solve (one process per GPU, all ranks hold the same assembled matrix and right-hand side): the W x G CTAs of all
// ranks form ONE virtual grid, CTA (rank, c) takes slice rank * G + c of the same row partition, so every rank keeps 1 / W of the
// matrix in its shared memory.  What crosses GPUs goes through PEER MEMORY inside the persistent kernel (NVLink loads / stores
// on buffers opened with CUDA IPC), no collective library and no kernel boundary:
//   * u = M^-1 r: the owner stores a row into its own copy and into the copy of every rank whose rows reference that column
//     (needmask, built per solve from the replicated pattern: the halo of a slab, plus rigid bodies / far contacts);
//   * dot products: every CTA stores its partial into the partial arrays of ALL ranks; after the barrier every CTA of every rank
//     re-reduces the same W x G values in the same order (identical scalars and decisions everywhere, as on one GPU);
//   * barrier after the vector phase (u and two partials published): one system fence, then one relaxed reduction on every rank's
//     counter, acquire spin on the own one (flat: the fence's round trip + one NVLink hop);
//   * all-reduce of w.u after the product: no fence and no counter -- every CTA stores ONE 16-byte flagged packet
//     {lo, flag, hi, flag} (8-byte halves, each with its flag: the granularity NVLink stores are atomic at) into slot
//     [virtual CTA] of every rank and then polls all W x G slots until they carry this use's flag: one NVLink hop, and the
//     arrival of all packets is the barrier (every CTA has finished its product);
//   * du: every CTA stores its slice of the solution into every rank's copy.
struct DistArgs {
    int world, rank;
    unsigned long long epoch_base;            // barriers completed by earlier solves (monotonic 64-bit counters, never reset)
    unsigned long long timeout_ns;            // a barrier that waits longer aborts the solve (done = 4) instead of hanging the GPU
    unsigned long long* bar[DIST_MAX_WORLD];  // every rank's barrier counter
    double* part[DIST_MAX_WORLD];             // every rank's partial arrays (the set of this solve's parity)
    double* u[DIST_MAX_WORLD];
    double* u4[DIST_MAX_WORLD];
    double* du[DIST_MAX_WORLD];
    uint4* ll[DIST_MAX_WORLD];                // every rank's packet slots [PCG_MAX_BLOCKS]
    unsigned ll_base;                         // flags used by earlier solves
    const unsigned char* needmask;            // [nbr] bit q: rank q reads this block row of u (own bit clear)
    int* abort_flag;                          // local: set by the first CTA that timed out
};

struct PcgArgs {
    const unsigned long long* rows; const int32_t* cols; const float* vals;
    const double* grad;
    float* dinv;                    // global fallbacks of the per-CTA slices
    double *x, *r, *p, *s, *w;
    double* u;                      // preconditioned residual, the one vector every CTA reads (compact: the TMA window source)
    double* u4;                     // the same, one 32-byte (x, y, z, 0) record per block row: far gathers take one request
    double* du;
    double* part;                   // 3 x PCG_MAX_BLOCKS partial sums
    int* rp_scratch;                // [nbr + grid + 1] local row pointers of slices whose row pointers do not fit in shared memory
    unsigned* barrier;              // monotonic arrival counter of the single-GPU barrier (zeroed once, never reset)
    unsigned barrier_base;          // its value when this solve starts
    PcgResult* result;
    int nbr;
    double abs_tol, rel_tol;
    int max_iter, stop_on_indef;
    unsigned long long nnzb;
    unsigned smem_bytes;            // dynamic shared memory of the launch
    int instrument;
    int force_stream;               // test hook: never keep the matrix slice resident
    int tiled;                      // experimental: stream through TMA-filled tile buffers (MODE 3) instead of ordinary loads
    long long* dbg;                 // [5 x grid + 2] per-CTA cycle counters (SB_PCG_DUMP diagnostics, else null)
    DistArgs d;                     // world <= 1: single GPU
    unsigned long long seq;         // number of this solve (PcgResult::seq)
};

struct Pcg {
    DevBuf<double> r, p, s, w, u, u4, x;
    DevBuf<float> dinv;
    DevBuf<double> part;
    DevBuf<int> rp_scratch;
    unsigned* d_barrier = nullptr;
    PcgResult* d_result = nullptr;
    PcgResult* h_result = nullptr;   // pinned, written by the kernel directly
    unsigned long long seq = 0;
    unsigned barrier_base = 0;       // arrivals of the solves so far (mod 2^32)
    int barrier_grid = 0;
    int grid = 0;
    unsigned smem_bytes = 0;
    unsigned smem_launch_last = 0;
};
// Peer-memory state of the distributed solve: ONE communication buffer per rank (cudaMalloc, exported with CUDA IPC), the same
// layout on every rank: [barrier counter | 2 sets of partial sums | u | u4 | du].
struct Dist {
    int world = 1, rank = 0;
    size_t max_dofs = 0;
    unsigned char* base[DIST_MAX_WORLD] = {nullptr};   // base[rank] is the own buffer
    bool opened[DIST_MAX_WORLD] = {false};              // peer mappings opened through IPC (closed at destroy)
    bool connected = false;
    bool enabled = true;                                // sb_dist_set_enabled: off = every rank solves locally (same-run single-GPU baseline)
    size_t bytes = 0, off_part = 0, off_ll = 0, off_u = 0, off_u4 = 0, off_du = 0, off_bcast = 0;
    unsigned long long n_bcasts = 0;                    // broadcasts from rank 0 so far (identical on every rank)
    unsigned ll_base = 0;                               // flags of the flagged all-reduce used so far (identical on every rank)
    unsigned long long epoch_base = 0;                  // barriers completed so far (identical on every rank)
    unsigned long long n_solves = 0, n_local_solves = 0;   // distributed solves / solves the policy kept on the own GPU
    DevBuf<unsigned char> needmask;
    int* d_abort = nullptr;                             // pinned host memory (the kernels write it, the host reads it after its synchronisations)
    int grid_override = 0;                              // test hook (SB_PCG_GRID): smaller grids so that two solves share one GPU
};
void dist_destroy(sb_context* ctx)
{
    Dist* D = ctx->dist;
    if (!D) return;
    for (int q = 0; q < D->world; q++)
        if (q != D->rank && D->opened[q] && D->base[q]) cudaIpcCloseMemHandle(D->base[q]);
    if (D->base[D->rank]) cudaFree(D->base[D->rank]);
    if (D->d_abort) cudaFreeHost(D->d_abort);
    D->needmask.release();
    delete D;
    ctx->dist = nullptr;
}

// Row partition of the virtual grid on the host (the device computes the same bounds: row_lower_bound / ROW_COST below) and
// the halo masks that follow from it; shared by sb_dist_plan (CPU tests) and nothing else on the product path.
static int host_row_lower_bound(const unsigned long long* rows, int nbr, unsigned long long t)
{
    int lo = 0, hi = nbr;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rows[mid] + 4ull * (unsigned long long)mid < t) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static Pcg* get(sb_context* ctx)
{
    if (!ctx->pcg) {
        ctx->pcg = new Pcg();
        cudaMalloc(&ctx->pcg->d_barrier, sizeof(unsigned));
        cudaMalloc(&ctx->pcg->d_result, sizeof(PcgResult));
        cudaMallocHost(&ctx->pcg->h_result, sizeof(PcgResult));
    }
    return ctx->pcg;
}
void pcg_destroy(sb_context* ctx)
{
    Pcg* P = ctx->pcg;
    if (!P) return;
    P->r.release(); P->p.release(); P->s.release(); P->w.release(); P->u.release(); P->u4.release(); P->x.release(); P->dinv.release(); P->part.release(); P->rp_scratch.release();
    if (P->d_barrier) cudaFree(P->d_barrier);
    if (P->d_result) cudaFree(P->d_result);
    if (P->h_result) cudaFreeHost(P->h_result);
    delete P;
    ctx->pcg = nullptr;
}

// ---- grid-wide barrier: monotonic counter (never reset: `base` = its value when the solve started, a multiple of gridDim.x),
//      the epoch-th barrier of a solve completes when it reaches base + epoch * gridDim.x (wrap-safe comparison) ----
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned base, unsigned& epoch)
{
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = base + epoch * gridDim.x;
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

// ---- mbarrier + TMA bulk load (global -> shared) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// ---- block reductions (fixed tree) ----
__device__ __forceinline__ double block_sum(double v, double* s)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < PCG_THREADS / 32) t = s[threadIdx.x];
    if (w == 0) {
        for (int o = PCG_THREADS / 64; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}
__device__ __forceinline__ double block_max(double v, double* s)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < PCG_THREADS / 32) t = s[threadIdx.x];
    if (w == 0) {
        for (int o = PCG_THREADS / 64; o > 0; o >>= 1) t = fmax(t, __shfl_down_sync(0xffffffffu, t, o));
    }
    return t;  // valid in thread 0
}
// two sums at once (one pair of barriers instead of two)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* s)
{
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { s[w] = a; s[PCG_THREADS / 32 + w] = b; }
    __syncthreads();
    double ta = 0.0, tb = 0.0;
    if (threadIdx.x < PCG_THREADS / 32) { ta = s[threadIdx.x]; tb = s[PCG_THREADS / 32 + threadIdx.x]; }
    if (w == 0) {
        for (int o = PCG_THREADS / 64; o > 0; o >>= 1) { ta += __shfl_down_sync(0xffffffffu, ta, o); tb += __shfl_down_sync(0xffffffffu, tb, o); }
    }
    a = ta; b = tb;   // valid in thread 0
}
// every CTA sums the same n (= virtual grid size) partials of two arrays in the same order -> identical values everywhere
__device__ __forceinline__ void all_partials2(const double* partA, const double* partB, int n, double* s, double* bc, double& outA, double& outB)
{
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += PCG_THREADS) { a += __ldcg(partA + i); b += __ldcg(partB + i); }
    block_sum2(a, b, s);
    if (threadIdx.x == 0) { bc[0] = a; bc[1] = b; }
    __syncthreads();
    outA = bc[0]; outB = bc[1];
}
// every CTA sums the same n partials in the same order -> identical value everywhere
__device__ __forceinline__ double all_partials(const double* part, int n, double* s, double* bc)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += PCG_THREADS) v += __ldcg(part + i);
    const double t = block_sum(v, s);
    if (threadIdx.x == 0) *bc = t;
    __syncthreads();
    return *bc;
}
__device__ __forceinline__ double all_partials_max(const double* part, int n, double* s, double* bc)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += PCG_THREADS) v = fmax(v, __ldcg(part + i));
    const double t = block_max(v, s);
    if (threadIdx.x == 0) *bc = t;
    __syncthreads();
    return *bc;
}

__device__ __forceinline__ void apply_dinv(const float* d, double r0, double r1, double r2, double& z0, double& z1, double& z2)
{
    z0 = (double)d[0] * r0 + (double)d[1] * r1 + (double)d[2] * r2;
    z1 = (double)d[3] * r0 + (double)d[4] * r1 + (double)d[5] * r2;
    z2 = (double)d[6] * r0 + (double)d[7] * r1 + (double)d[8] * r2;
}

// ---- flagged all-reduce (sum) over the virtual grid: see DistArgs ----
__device__ __forceinline__ unsigned long long global_ns();
__device__ __forceinline__ void ll_put(const DistArgs& D, int vb, double v, unsigned flag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned lo = (unsigned)b, hi = (unsigned)(b >> 32);
    for (int q = 0; q < D.world; q++)
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(D.ll[q] + vb), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ double ll_all_sum(const DistArgs& D, int n, unsigned flag, double* s, double* bc, int* s_abort)
{
    const uint4* slots = D.ll[D.rank];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += PCG_THREADS) {
        unsigned x, f0, y, f1, spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(f0), "=r"(y), "=r"(f1) : "l"(slots + i) : "memory");
            if (f0 == flag && f1 == flag) break;
            if ((++spins & 1023u) == 0u) {
                const unsigned long long now = global_ns();
                if (!t0) t0 = now;
                if (now - t0 > D.timeout_ns || *(volatile int*)D.abort_flag || *(volatile int*)s_abort) { *(volatile int*)D.abort_flag = 1; *(volatile int*)s_abort = 1; x = y = 0u; break; }
            }
        }
        a += __longlong_as_double((long long)(((unsigned long long)y << 32) | x));
    }
    const double t = block_sum(a, s);
    if (threadIdx.x == 0) *bc = t;
    __syncthreads();
    return *bc;
}

// Row partition: a CTA's iteration costs about one unit per block (product) plus ROW_COST units per block row (vector
// phase, row bookkeeping), so the rows are cut into slices of equal  blocks + ROW_COST * rows.  (Cutting by blocks alone gives
// the slices of sparse rows -- hex-centre nodes: 9 blocks per row instead of 27 -- twice the rows and twice the vector phase.)
constexpr unsigned long long ROW_COST = 4;   // (host_row_lower_bound above restates it)
// first row r in [0, nbr] with rows[r] + ROW_COST * r >= t
__device__ __forceinline__ int row_lower_bound(const unsigned long long* __restrict__ rows, int nbr, unsigned long long t)
{
    int lo = 0, hi = nbr;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rows[mid] + ROW_COST * (unsigned long long)mid < t) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// the same bound found by one warp: 32 probes per round, four rounds instead of sixteen dependent global loads for 37 k rows
__device__ __forceinline__ int row_lower_bound_warp(const unsigned long long* __restrict__ rows, int nbr, unsigned long long t)
{
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = nbr;   // the answer is in [lo, hi]; rows below lo fail the test, hi passes it (or is nbr)
    while (lo < hi) {
        const int step = max(1, (hi - lo) / 32);
        const int p = lo + lane * step;
        const bool ge = (p >= hi) || (rows[p] + ROW_COST * (unsigned long long)p >= t);
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        if (m == 0u) { lo = lo + 31 * step + 1; continue; }
        const int first = __ffs(m) - 1;
        const int nhi = min(lo + first * step, hi);
        lo = (first > 0) ? lo + (first - 1) * step + 1 : lo;
        hi = nhi;
    }
    return lo;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- barrier over the CTAs of ALL ranks of a distributed solve ----
// Everything this CTA stored into peer memory (u halo rows, partial sums, du) is ordered before its arrival by the CTA barrier
// and the system-scope fence; the arrival is one release-reduction on every rank's counter (W NVLink stores, not waited for),
// the wait an acquire spin on the OWN counter.  The counters are monotonic over the life of the context (a rank that is one
// solve ahead may already be arriving at the next solve's first barrier).  A wait longer than the time-out, or an abort seen
// on this GPU, ends the solve (done = 4) instead of hanging the device.
__device__ __forceinline__ bool dist_barrier(const DistArgs& D, unsigned long long& epoch, int* s_abort)
{
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0 && !*s_abort) {
        __threadfence_system();   // (one fence for all W arrivals: a release per reduction would pay the round trip W times)
        for (int q = 0; q < D.world; q++)
            asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(D.bar[q]), "l"(1ull) : "memory");
        const unsigned long long target = (D.epoch_base + epoch) * (unsigned long long)(D.world * (int)gridDim.x);
        unsigned long long v, t0 = 0;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(D.bar[D.rank]) : "memory");
            if (v >= target) break;
            if ((++spins & 1023u) == 0u) {
                const unsigned long long now = global_ns();
                if (!t0) t0 = now;
                if (now - t0 > D.timeout_ns || *(volatile int*)D.abort_flag) { *(volatile int*)D.abort_flag = 1; *s_abort = 1; break; }
            }
        }
    }
    __syncthreads();
    return *s_abort != 0;
}

// The dynamic shared memory of the solver.  A CTA whose slices all fit (the normal case) runs the FAST instance of the solve
// body, in which every slice pointer is derived from this symbol, so that the compiler emits shared-memory loads / stores
// (LDS / STS); pointers that may be either shared or global are generic and every access pays the generic-address path of
// the LSU, which made the SpMV 4x slower.
extern __shared__ __align__(16) unsigned char pcg_smem[];

struct PcgPlan {
    int r0, nr, w0, nwin, n_long;
    bool own_in_win;
    unsigned win_bytes;
    unsigned off_rp, off_r, off_p, off_s, off_w, off_dinv, off_cols, off_vals, off_win;   // byte offsets in pcg_smem (FAST)
    int* rp; double *rs, *ps, *ss, *ws; float* dinv; const int32_t* cols; const float* vals; double* uwin;   // generic pointers
    double* s; double* bc2; int* s_long; unsigned long long* mbar;
    int n_seg; int *s_seg_j0, *s_seg_j1, *s_seg_first; double* s_seg_y;
    unsigned long long b0;                 // first global block of the slice
    unsigned off_tile[2];                  // tile buffers (MODE 2)
    unsigned long long* tbar;              // their mbarriers
    unsigned long long t_start, t_loaded;
    int* s_abort;
};

// MODE 1: every slice (row pointers, vectors, matrix, window) in shared memory.  MODE 2 (the matrix slice does not fit: 66 k-node
// cloth): row pointers, vectors and window in shared memory, the matrix read from global memory / L2 through the read-only
// path every iteration.  MODE 3 (experimental, SB_PCG_TILED=1): as MODE 2, but the matrix slice streams through two
// shared-memory tile buffers filled by TMA bulk copies (double buffered; the first two tiles of the next product are
// prefetched under the vector phase and the barriers).  Measured at C2 with the matrix forced out of shared memory: product
// 5.0 us resident, 11.4 us MODE 2, 19.8 us MODE 3 -- one 30 KB bulk copy in flight per SM does not cover the copy latency;
// it needs a deeper ring of smaller tiles before it can replace MODE 2.  MODE 0: anything may live in global memory (generic
// pointers; million-tet slices, tiny budgets).
template<int MODE, bool DIST>
__device__ __forceinline__ void pcg_body(const PcgArgs& A, const PcgPlan& P)
{
    const int G = gridDim.x;
    const int VG = DIST ? A.d.world * (int)gridDim.x : (int)gridDim.x;               // virtual grid: the CTAs of all ranks
    const int vb = DIST ? A.d.rank * (int)gridDim.x + (int)blockIdx.x : (int)blockIdx.x;   // this CTA in it
    const int nbr = A.nbr;
    const int tid = threadIdx.x;
    unsigned epoch = 0;
    unsigned long long depoch = 0;
    unsigned ll_uses = 0;
    double* part0 = DIST ? A.d.part[A.d.rank] : A.part;
    double* part1 = part0 + PCG_MAX_BLOCKS;
    double* part2 = part0 + 2 * PCG_MAX_BLOCKS;
    // this CTA's partial sum k: into the own array, or into the arrays of all ranks
    auto put_part = [&](int k, double v) {
        if (DIST) { for (int q = 0; q < A.d.world; q++) __stcg(A.d.part[q] + k * PCG_MAX_BLOCKS + vb, v); }
        else __stcg(part0 + k * PCG_MAX_BLOCKS + vb, v);
    };
    // barrier over the virtual grid; true = the distributed solve was aborted
    auto gbar = [&]() -> bool {
        if (DIST) return dist_barrier(A.d, depoch, P.s_abort);
        grid_barrier(A.barrier, A.barrier_base, epoch);
        return false;
    };
    double* s = P.s;
    double* bc2 = P.bc2;
    double& bc = bc2[0];
    int* s_long = P.s_long;
    const int n_seg = P.n_seg;
    const int* s_seg_j0 = P.s_seg_j0; const int* s_seg_j1 = P.s_seg_j1; const int* s_seg_first = P.s_seg_first;
    double* s_seg_y = P.s_seg_y;
    unsigned long long& s_mbar = *P.mbar;
    const int r0 = P.r0, nr = P.nr, w0 = P.w0, nwin = P.nwin, n_long = P.n_long;
    const bool own_in_win = P.own_in_win;
    const unsigned win_bytes = P.win_bytes;
    const unsigned long long t_start = P.t_start, t_loaded = P.t_loaded;
    int* rp = (MODE != 0) ? reinterpret_cast<int*>(pcg_smem + P.off_rp) : P.rp;
    double* rs = (MODE != 0) ? reinterpret_cast<double*>(pcg_smem + P.off_r) : P.rs;
    double* ps = (MODE != 0) ? reinterpret_cast<double*>(pcg_smem + P.off_p) : P.ps;
    double* ss = (MODE != 0) ? reinterpret_cast<double*>(pcg_smem + P.off_s) : P.ss;
    double* ws = (MODE != 0) ? reinterpret_cast<double*>(pcg_smem + P.off_w) : P.ws;
    float* dinv = (MODE != 0) ? reinterpret_cast<float*>(pcg_smem + P.off_dinv) : P.dinv;
    const int32_t* cols = (MODE == 1) ? reinterpret_cast<const int32_t*>(pcg_smem + P.off_cols) : P.cols;
    const float* vals = (MODE == 1) ? reinterpret_cast<const float*>(pcg_smem + P.off_vals) : P.vals;
    auto COL = [&](int j) -> int { return (MODE >= 2) ? __ldg(cols + j) : cols[j]; };
    auto VAL = [&](const float* q) -> float { return (MODE >= 2) ? __ldg(q) : *q; };
    double* uwin = (MODE != 0) ? reinterpret_cast<double*>(pcg_smem + P.off_win) : P.uwin;
    (void)G; (void)nbr; (void)bc; (void)VG; (void)depoch; (void)ll_uses;
    auto is_swept = [&](int lr) {   // long AND listed (every listed row gets its own block reduction)
        if (rp[lr + 1] - rp[lr] <= LONG_ROW) return false;
        for (int k = 0; k < n_long; k++) if (s_long[k] == lr) return true;
        return false;
    };
    double* ug = A.u + 3 * (size_t)r0;                           // own slice of the global u
    auto store_u4 = [&](int br, double z0, double z1, double z2) {   // padded copy for the far gathers
        asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(A.u4 + 4 * (size_t)br), "d"(z0), "d"(z1), "d"(z2), "d"(0.0) : "memory");
    };
    // publish u of own row lr: own copies (compact + padded) and, in a distributed solve, the copies of the ranks that read it
    auto publish_u = [&](int lr, double z0, double z1, double z2) {
        __stcg(ug + 3 * lr, z0); __stcg(ug + 3 * lr + 1, z1); __stcg(ug + 3 * lr + 2, z2);
        store_u4(r0 + lr, z0, z1, z2);
        if (DIST) {
            unsigned m = A.d.needmask[r0 + lr];
            while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                double* pu = A.d.u[q] + 3 * (size_t)(r0 + lr);
                __stcg(pu, z0); __stcg(pu + 1, z1); __stcg(pu + 2, z2);
                asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(A.d.u4[q] + 4 * (size_t)(r0 + lr)), "d"(z0), "d"(z1), "d"(z2), "d"(0.0) : "memory");
            }
        }
    };
    const double* uo_win = own_in_win ? uwin + 3 * (size_t)(r0 - w0) : nullptr;
    auto uo = [&](int i) -> double { return own_in_win ? uo_win[i] : __ldcg(ug + i); };   // own slice of u as this CTA reads it
    double* xg = A.x + 3 * (size_t)r0;                           // x is only ever touched by its owner thread

    // gather of u at scalar column c (first of the three of a block column)
    const int win_lo = 3 * w0, win_hi = 3 * (w0 + nwin);
    auto gather3 = [&](int c, double& a0, double& a1, double& a2) {
        if (c >= win_lo && c < win_hi) { const double* q = uwin + (c - win_lo); a0 = q[0]; a1 = q[1]; a2 = q[2]; }
        else {
            // far column (hex-centre node, rigid body): ONE 32-byte request to the padded copy of u.  The SM's miss path
            // takes about two cycles per request whatever its size, so three 8-byte loads cost three times as much
            double pad;
            asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(pad) : "l"(A.u4 + 4 * (size_t)(c / 3)));
        }
    };
    unsigned win_phase = 0;
    // pull the window of u (all CTAs have published their slices: call after a grid barrier)
    // (issue and wait are separate so that the copy runs under the all-reduce that follows the barrier)
    auto issue_window = [&]() {
        if (!own_in_win) return;
        if (tid == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_expect_tx(&s_mbar, win_bytes);
            tma_load_bulk(uwin, A.u + 3 * (size_t)w0, win_bytes, &s_mbar);
        }
    };
    auto wait_window = [&]() {
        if (!own_in_win) return;
        mbar_wait(&s_mbar, win_phase);
        win_phase ^= 1u;
    };
    auto load_window = [&]() { issue_window(); wait_window(); };
    // w = A u on the own rows; returns this thread's share of w.u
    const int lane = tid % LANES_PER_ROW;
    constexpr int rows_per_pass = PCG_THREADS / LANES_PER_ROW;
    auto spmv = [&]() -> double {
        double wu = 0.0;
        // groups of LANES_PER_ROW lanes take one row each; the segments of the long rows (rigid bodies in contact with many
        // nodes) queue behind the rows as further groups, so a 700-block row costs a few trips of idle groups instead of a
        // sweep and a 1024-thread reduction by the whole CTA
        const int n_groups = nr + n_seg;
        for (int base = 0; base < n_groups; base += rows_per_pass) {
            const int g = base + tid / LANES_PER_ROW;
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
            int j0 = 0, j1 = 0;
            bool is_row = false, is_seg = false;
            if (g < nr) {
                if (!is_swept(g)) { j0 = rp[g]; j1 = rp[g + 1]; is_row = true; }
            } else if (g < n_groups) {
                j0 = s_seg_j0[g - nr]; j1 = s_seg_j1[g - nr]; is_seg = true;
            }
            // four blocks per trip: their gathers of u (shared-memory window, or L2 for far columns -- hex-centre nodes,
            // rigid bodies) are all issued before the first product, so a row costs one memory round trip, not one per block
            constexpr int U = 4;
            for (int j = j0 + lane; j < j1; j += U * LANES_PER_ROW) {
                double a[U][3];
#pragma unroll
                for (int t = 0; t < U; t++) {
                    const int jj = j + t * LANES_PER_ROW;
                    a[t][0] = 0.0; a[t][1] = 0.0; a[t][2] = 0.0;
                    if (jj < j1) gather3(COL(jj), a[t][0], a[t][1], a[t][2]);
                }
#pragma unroll
                for (int t = 0; t < U; t++) {
                    const int jj = j + t * LANES_PER_ROW;
                    if (jj < j1) {
                        const float* m = vals + 9 * (size_t)jj;   // column-major 3x3
                        y0 += (double)VAL(m + 0) * a[t][0] + (double)VAL(m + 3) * a[t][1] + (double)VAL(m + 6) * a[t][2];
                        y1 += (double)VAL(m + 1) * a[t][0] + (double)VAL(m + 4) * a[t][1] + (double)VAL(m + 7) * a[t][2];
                        y2 += (double)VAL(m + 2) * a[t][0] + (double)VAL(m + 5) * a[t][1] + (double)VAL(m + 8) * a[t][2];
                    }
                }
            }
            for (int o = LANES_PER_ROW / 2; o > 0; o >>= 1) {
                y0 += __shfl_down_sync(0xffffffffu, y0, o, LANES_PER_ROW);
                y1 += __shfl_down_sync(0xffffffffu, y1, o, LANES_PER_ROW);
                y2 += __shfl_down_sync(0xffffffffu, y2, o, LANES_PER_ROW);
            }
            if (lane == 0 && is_row) {
                ws[3 * g] = y0; ws[3 * g + 1] = y1; ws[3 * g + 2] = y2;
                wu += uo(3 * g) * y0 + uo(3 * g + 1) * y1 + uo(3 * g + 2) * y2;
            }
            if (lane == 0 && is_seg) {
                s_seg_y[3 * (g - nr)] = y0; s_seg_y[3 * (g - nr) + 1] = y1; s_seg_y[3 * (g - nr) + 2] = y2;
            }
        }
        // long rows: segment sums added in segment order (deterministic)
        if (n_long) {
            __syncthreads();
            if (tid < 3 * n_long) {
                const int q = tid / 3, c = tid % 3;
                double t = 0.0;
                for (int k = s_seg_first[q]; k < s_seg_first[q + 1]; k++) t += s_seg_y[3 * k + c];
                const int lr = s_long[q];
                ws[3 * lr + c] = t;
                wu += uo(3 * lr + c) * t;
            }
        }
        return wu;
    };

    // ---- MODE 3: product over streamed tiles ----
    constexpr int TB = PCG_TILE_BLOCKS;
    const int nb_own = rp[nr];
    const int d_al = (int)(P.b0 & 3ull);                               // the slice starts d_al blocks into its first (4-aligned) tile
    const int n_tiles = (MODE == 3 && nb_own > 0) ? (nb_own + d_al + TB - 1) / TB : 0;
    unsigned long long* tbar = P.tbar;
    unsigned tphase[2] = {0u, 0u};
    bool tloaded[2] = {false, false};
    int tile_use = 0;
    auto tile_cols = [&](int buf) { return reinterpret_cast<int32_t*>(pcg_smem + P.off_tile[buf]); };
    auto tile_vals = [&](int buf) { return reinterpret_cast<float*>(pcg_smem + P.off_tile[buf] + TB * 4); };
    auto issue_tile = [&](int t, int buf) {   // (one thread)
        const int first = t * TB;
        const unsigned cnt = (unsigned)min(TB, (nb_own + d_al - first + 3) & ~3);   // may run up to 3 blocks past the slice (allocation slack)
        const unsigned long long start = (P.b0 - (unsigned long long)d_al) + (unsigned long long)first;
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_expect_tx(&tbar[buf], cnt * 40u);
        tma_load_bulk(tile_cols(buf), A.cols + start, cnt * 4u, &tbar[buf]);
        tma_load_bulk(tile_vals(buf), A.vals + 9 * start, cnt * 36u, &tbar[buf]);
    };
    if (MODE == 3 && tid == 0) {
        if (n_tiles > 0) issue_tile(0, 0);
        if (n_tiles > 1) issue_tile(1, 1);
    }
    auto spmv_tiled = [&]() -> double {
        for (int i = tid; i < 3 * nr; i += PCG_THREADS) ws[i] = 0.0;
        for (int i = tid; i < 3 * n_seg; i += PCG_THREADS) s_seg_y[i] = 0.0;
        __syncthreads();
        for (int tt = 0; tt < n_tiles; tt++) {
            const int buf = (n_tiles > 2) ? (tile_use & 1) : tt;   // (one or two tiles: they stay where they were loaded)
            if (n_tiles > 2 || !tloaded[buf]) {
                mbar_wait(&tbar[buf], tphase[buf]);
                tphase[buf] ^= 1u;
                tloaded[buf] = true;
            }
            const int tb0 = tt * TB - d_al;                        // local block range of the tile (negative start in tile 0)
            const int lo = max(tb0, 0), hi = min(tb0 + TB, nb_own);
            // rows [ra, rb) and segments [sa, sb) with blocks in [lo, hi)
            int ra, rb, sa, sb;
            { int a = 0, b = nr; while (a < b) { const int m = (a + b) >> 1; if (rp[m + 1] <= lo) a = m + 1; else b = m; } ra = a; }
            { int a = ra, b = nr; while (a < b) { const int m = (a + b) >> 1; if (rp[m] < hi) a = m + 1; else b = m; } rb = a; }
            { int a = 0, b = n_seg; while (a < b) { const int m = (a + b) >> 1; if (s_seg_j1[m] <= lo) a = m + 1; else b = m; } sa = a; }
            { int a = sa, b = n_seg; while (a < b) { const int m = (a + b) >> 1; if (s_seg_j0[m] < hi) a = m + 1; else b = m; } sb = a; }
            const int n_rows_t = rb - ra, n_items = n_rows_t + (sb - sa);
            const int32_t* tc = tile_cols(buf) - tb0;              // indexed by the local block number
            const float* tv = tile_vals(buf) - 9 * tb0;
            for (int base = 0; base < n_items; base += rows_per_pass) {
                const int it2 = base + tid / LANES_PER_ROW;
                int j0 = 0, j1 = 0, row = -1, seg = -1;
                if (it2 < n_rows_t) {
                    const int g = ra + it2;
                    if (!is_swept(g)) { j0 = max(rp[g], lo); j1 = min(rp[g + 1], hi); row = g; }
                } else if (it2 < n_items) {
                    seg = sa + (it2 - n_rows_t);
                    j0 = max(s_seg_j0[seg], lo); j1 = min(s_seg_j1[seg], hi);
                }
                double y0 = 0.0, y1 = 0.0, y2 = 0.0;
                constexpr int U = 4;
                for (int j = j0 + lane; j < j1; j += U * LANES_PER_ROW) {
                    double a[U][3];
#pragma unroll
                    for (int t = 0; t < U; t++) {
                        const int jj = j + t * LANES_PER_ROW;
                        a[t][0] = 0.0; a[t][1] = 0.0; a[t][2] = 0.0;
                        if (jj < j1) gather3(tc[jj], a[t][0], a[t][1], a[t][2]);
                    }
#pragma unroll
                    for (int t = 0; t < U; t++) {
                        const int jj = j + t * LANES_PER_ROW;
                        if (jj < j1) {
                            const float* m = tv + 9 * jj;
                            y0 += (double)m[0] * a[t][0] + (double)m[3] * a[t][1] + (double)m[6] * a[t][2];
                            y1 += (double)m[1] * a[t][0] + (double)m[4] * a[t][1] + (double)m[7] * a[t][2];
                            y2 += (double)m[2] * a[t][0] + (double)m[5] * a[t][1] + (double)m[8] * a[t][2];
                        }
                    }
                }
                for (int o = LANES_PER_ROW / 2; o > 0; o >>= 1) {
                    y0 += __shfl_down_sync(0xffffffffu, y0, o, LANES_PER_ROW);
                    y1 += __shfl_down_sync(0xffffffffu, y1, o, LANES_PER_ROW);
                    y2 += __shfl_down_sync(0xffffffffu, y2, o, LANES_PER_ROW);
                }
                if (lane == 0 && row >= 0) { ws[3 * row] += y0; ws[3 * row + 1] += y1; ws[3 * row + 2] += y2; }   // (a row split over two tiles adds twice)
                if (lane == 0 && seg >= 0) { s_seg_y[3 * seg] += y0; s_seg_y[3 * seg + 1] += y1; s_seg_y[3 * seg + 2] += y2; }
            }
            __syncthreads();     // the buffer is free, and this tile's sums are visible to the next tile's
            if (n_tiles > 2 && tid == 0) issue_tile((tt + 2) % max(n_tiles, 1), buf);
            tile_use++;
        }
        if (n_long) {
            if (tid < 3 * n_long) {
                const int q = tid / 3, c = tid % 3;
                double t = 0.0;
                for (int k = s_seg_first[q]; k < s_seg_first[q + 1]; k++) t += s_seg_y[3 * k + c];
                ws[3 * s_long[q] + c] = t;
            }
            __syncthreads();
        }
        double wu = 0.0;
        for (int i = tid; i < 3 * nr; i += PCG_THREADS) wu += uo(i) * ws[i];
        return wu;
    };
    auto product = [&]() -> double { return (MODE == 3) ? spmv_tiled() : spmv(); };
    // no bulk copy may be in flight when the CTA leaves
    auto drain_tiles = [&]() {
        if (MODE != 3) return;
        for (int buf = 0; buf < 2 && buf < n_tiles; buf++)
            if (n_tiles > 2 || !tloaded[buf]) mbar_wait(&tbar[buf], tphase[buf]);
    };

    // ---- phase 0: M^-1 = block-Jacobi inverse; x = 0, r = b = -grad, u = M^-1 r, p = s = 0 ; partials of b.b and r.u ----
    {
        double bb = 0.0, ru = 0.0;
        for (int lr = tid; lr < nr; lr += PCG_THREADS) {
            const int br = r0 + lr;
            // closed-form inverse of the symmetric 3x3 diagonal block, in float like the reference (BlockedSparseMatrix.h:1198-1214)
            float mi[9];
            for (int k = 0; k < 9; k++) mi[k] = 0.0f;
            {   // columns are sorted inside a row: binary search for the diagonal block
                int lo = rp[lr], hi = rp[lr + 1];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (COL(mid) < 3 * br) lo = mid + 1; else hi = mid;
                }
                if (lo < rp[lr + 1] && COL(lo) == 3 * br) {
                    const float* m = vals + 9 * (size_t)lo;
                    const float tmp0 = m[4] * m[8];
                    const float tmp1 = m[5] * m[5];
                    const float tmp2 = m[2] * m[5];
                    const float tmp3 = m[1] * m[1];
                    const float tmp4 = m[2] * m[2];
                    const float tmp5 = (float)(1.0 / (double)(m[0] * tmp0 - m[0] * tmp1 + 2 * m[1] * tmp2 - m[4] * tmp4 - m[8] * tmp3));
                    mi[8] = tmp5 * (m[0] * m[4] - tmp3);
                    mi[4] = tmp5 * (m[0] * m[8] - tmp4);
                    mi[0] = tmp5 * (tmp0 - tmp1);
                    mi[3] = -tmp5 * (m[1] * m[8] - tmp2);
                    mi[1] = mi[3];
                    mi[6] = tmp5 * (m[1] * m[5] - m[4] * m[2]);
                    mi[2] = mi[6];
                    mi[7] = -tmp5 * (m[0] * m[5] - m[1] * m[2]);
                    mi[5] = mi[7];
                }
            }
            for (int k = 0; k < 9; k++) dinv[9 * lr + k] = mi[k];
            const double g0 = -A.grad[3 * br], g1 = -A.grad[3 * br + 1], g2 = -A.grad[3 * br + 2];
            double z0, z1, z2;
            apply_dinv(mi, g0, g1, g2, z0, z1, z2);
            for (int c = 0; c < 3; c++) { xg[3 * lr + c] = 0.0; ps[3 * lr + c] = 0.0; ss[3 * lr + c] = 0.0; }
            rs[3 * lr] = g0; rs[3 * lr + 1] = g1; rs[3 * lr + 2] = g2;
            publish_u(lr, z0, z1, z2);
            bb += g0 * g0 + g1 * g1 + g2 * g2;
            ru += g0 * z0 + g1 * z1 + g2 * z2;
        }
        block_sum2(bb, ru, s);
        if (tid == 0) { put_part(0, bb); put_part(1, ru); }
    }
    const bool aborted0 = gbar();
    double bb, gamma;
    all_partials2(part0, part1, VG, s, bc2, bb, gamma);

    int it = 0, done = 0, found_indef = 0;
    double error = 1.0;               // x0 = 0 -> r = b
    const double error0 = 1.0;
    if (bb < A.abs_tol * A.abs_tol) { done = 1; error = 0.0; }   // zero right-hand side
    else if (1.0 < A.abs_tol) done = 1;
    else if (A.max_iter <= 0) done = 3;
    if (aborted0) done = 4;

    const unsigned long long t_loop = global_ns();
    long long c_spmv = 0, c_bar = 0, c_red = 0, c_vec = 0, c_win = 0, c_t = A.instrument ? clock64() : 0;
#define PCG_TICK(acc) if (A.instrument) { const long long _n = clock64(); acc += _n - c_t; c_t = _n; }
    double alpha = 0.0, beta = 0.0;
    if (!done) {
        // ---- first product: w = A u ; delta = w.u = p^T A p of the first iteration ----
        load_window();
        const double wu = product();
        const double t = block_sum(wu, s);
        bool ab;
        double delta;
        if (DIST) {
            if (tid == 0) ll_put(A.d, vb, t, A.d.ll_base + (++ll_uses));
            else ++ll_uses;
            PCG_TICK(c_spmv);
            delta = ll_all_sum(A.d, VG, A.d.ll_base + ll_uses, s, &bc, P.s_abort);
            ab = *P.s_abort != 0;
            PCG_TICK(c_bar);
        } else {
            if (tid == 0) put_part(2, t);
            PCG_TICK(c_spmv);
            ab = gbar();
            PCG_TICK(c_bar);
            delta = all_partials(part2, VG, s, &bc);
        }
        PCG_TICK(c_red);
        if (ab) done = 4;
        else if (delta <= 0.0) {
            found_indef = 1;
            if (A.stop_on_indef) { it = 1; done = 2; }   // x is returned as is (solve_pcg.h:183-192)
        }
        alpha = gamma / delta;
    }
    while (!done) {
        it++;
        // ---- p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; partials of r.r and r.u ----
        {
            double rr = 0.0, ru = 0.0;
            for (int lr = tid; lr < nr; lr += PCG_THREADS) {
                double q[3], z0, z1, z2;
                for (int c = 0; c < 3; c++) {
                    const double pn = uo(3 * lr + c) + beta * ps[3 * lr + c];
                    const double sn = ws[3 * lr + c] + beta * ss[3 * lr + c];
                    ps[3 * lr + c] = pn; ss[3 * lr + c] = sn;
                    xg[3 * lr + c] += alpha * pn;
                    q[c] = rs[3 * lr + c] - alpha * sn;
                    rs[3 * lr + c] = q[c];
                }
                apply_dinv(dinv + 9 * lr, q[0], q[1], q[2], z0, z1, z2);
                publish_u(lr, z0, z1, z2);
                rr += q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
                ru += q[0] * z0 + q[1] * z1 + q[2] * z2;
            }
            block_sum2(rr, ru, s);
            if (tid == 0) { put_part(0, rr); put_part(1, ru); }
        }
        PCG_TICK(c_vec);
        if (gbar()) { done = 4; break; }
        PCG_TICK(c_bar);
        issue_window();     // every slice of u is published: the window copy runs under the reduction below
        double rr, gamma_new;
        all_partials2(part0, part1, VG, s, bc2, rr, gamma_new);
        error = sqrt(rr / bb);
        PCG_TICK(c_red);
        wait_window();      // (also on the way out: no copy may be in flight when the CTA leaves)
        PCG_TICK(c_win);
        if (error < A.abs_tol || error / error0 < A.rel_tol) { done = 1; break; }
        if (it >= A.max_iter) { done = 3; break; }
        // ---- w = A u ; delta = w.u ----
        {
            const double wu = product();
            const double t = block_sum(wu, s);
            if (DIST) { if (tid == 0) ll_put(A.d, vb, t, A.d.ll_base + (++ll_uses)); else ++ll_uses; }
            else if (tid == 0) put_part(2, t);
        }
        PCG_TICK(c_spmv);
        double delta;
        if (DIST) {
            delta = ll_all_sum(A.d, VG, A.d.ll_base + ll_uses, s, &bc, P.s_abort);
            if (*P.s_abort) { done = 4; break; }
            PCG_TICK(c_bar);
        } else {
            if (gbar()) { done = 4; break; }
            PCG_TICK(c_bar);
            delta = all_partials(part2, VG, s, &bc);
        }
        PCG_TICK(c_red);
        beta = gamma_new / gamma;
        const double pAp = delta - beta * gamma_new / alpha;     // p^T A p of the coming iteration
        gamma = gamma_new;
        if (pAp <= 0.0) {
            found_indef = 1;
            if (A.stop_on_indef) { it++; done = 2; break; }      // the reference counts the iteration that meets it
        }
        alpha = gamma / pAp;
    }

    drain_tiles();
    // ---- du = x ; du.grad and |du|_inf ----
    {
        double dg = 0.0, mx = 0.0;
        for (int lr = tid; lr < nr; lr += PCG_THREADS) {
            for (int c = 0; c < 3; c++) {
                const double v = xg[3 * lr + c];
                if (DIST) { for (int q = 0; q < A.d.world; q++) __stcg(A.d.du[q] + 3 * (size_t)(r0 + lr) + c, v); }
                else A.du[3 * (size_t)(r0 + lr) + c] = v;
                dg += v * A.grad[3 * (size_t)(r0 + lr) + c];
                mx = fmax(mx, fabs(v));
            }
        }
        const double t0 = block_sum(dg, s);
        const double t1 = block_max(mx, s);
        if (gbar()) done = 4;   // CTAs that left the loop one reduction behind may still be reading the partial arrays
        if (tid == 0) { put_part(0, t0); put_part(1, t1); }
    }
    if (gbar()) done = 4;
    if (A.dbg && tid == 0) {
        const int G = gridDim.x;
        A.dbg[blockIdx.x] = c_spmv; A.dbg[G + blockIdx.x] = c_bar + c_red; A.dbg[2 * G + blockIdx.x] = c_vec; A.dbg[3 * G + blockIdx.x] = c_win;
        A.dbg[4 * G + blockIdx.x] = ((long long)nr << 32) | (unsigned)(rp[nr] - rp[0]);
    }
    if (blockIdx.x == 0) {
        const double dg = all_partials(part0, VG, s, &bc);
        const double mx = all_partials_max(part1, VG, s, &bc);
        if (tid == 0) {
            PcgResult R;
            R.du_dot_grad = dg; R.du_inf = mx; R.error = error; R.bb = bb;
            R.it = it; R.done = done; R.found_indef = found_indef; R.pad = 0;
            R.c_spmv = c_spmv; R.c_bar = c_bar; R.c_red = c_red; R.c_vec = c_vec; R.c_win = c_win;
            R.t_start = t_start; R.t_loaded = t_loaded; R.t_loop = t_loop; R.t_end = global_ns();
            R.barriers = DIST ? depoch : (unsigned long long)epoch;
            R.ll_uses = ll_uses;
            R.seq = 0;
            // the record goes straight into pinned host memory; its sequence number follows behind a system fence, so the host,
            // which polls that word, never sees a half-written record (no copy node, no stream synchronisation on the way back)
            *A.result = R;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(&A.result->seq) = A.seq;
        }
    }
}

template<bool DIST>
__global__ void __launch_bounds__(PCG_THREADS, 1) k_pcg_solve(const PcgArgs A)
{
    unsigned char* const smem = pcg_smem;
    __shared__ double s[2 * (PCG_THREADS / 32)];
    __shared__ double bc2[2];
    __shared__ int s_range[2];
    __shared__ int s_long[MAX_LONG_ROWS];
    __shared__ int s_seg_j0[MAX_SEGS], s_seg_j1[MAX_SEGS], s_seg_first[MAX_LONG_ROWS + 1];
    __shared__ int s_n_seg;
    __shared__ double s_seg_y[3 * MAX_SEGS];
    __shared__ int s_n_long;
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ __align__(8) unsigned long long s_tbar[2];
    __shared__ int s_abort;
    const unsigned long long t_start = global_ns();
    const int G = DIST ? A.d.world * (int)gridDim.x : (int)gridDim.x;                      // virtual grid (all ranks)
    const int vb = DIST ? A.d.rank * (int)gridDim.x + (int)blockIdx.x : (int)blockIdx.x;   // this CTA in it
    const int nbr = A.nbr;
    const int tid = threadIdx.x;

    // ---- this CTA's rows: [r0, r1), holding blocks [b0, b0 + nb) ----
    if (tid < 64) {   // (warps 0 and 1: one boundary each)
        const unsigned long long cost = A.nnzb + ROW_COST * (unsigned long long)nbr;
        const int which = tid >> 5;
        int r = row_lower_bound_warp(A.rows, nbr, (cost * (unsigned long long)(vb + which)) / (unsigned long long)G);
        if (which == 1 && vb == G - 1) r = nbr;
        if ((tid & 31) == 0) s_range[which] = r;
    }
    if (tid == 0) {
        s_abort = 0;
        mbar_init(&s_mbar, 1);
        mbar_init(&s_tbar[0], 1);
        mbar_init(&s_tbar[1], 1);
    }
    __syncthreads();
    const int r0 = s_range[0], nr = s_range[1] - s_range[0];
    const unsigned long long b0 = A.rows[r0];
    const int nb = (int)(A.rows[r0 + nr] - b0);

    // ---- shared-memory plan: row pointers; the vector slices r, p, s, w, M^-1; the matrix slice; the window of u gets
    //      what is left.  Whatever does not fit stays in global memory and is reached through the same (generic) pointers ----
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    const bool rp_fit = ((sizeof(int) * (nr + 1) + 15) & ~(size_t)15) <= A.smem_bytes;
    PcgPlan P;
    P.off_rp = P.off_r = P.off_p = P.off_s = P.off_w = P.off_dinv = P.off_cols = P.off_vals = P.off_win = 0;
    if (rp_fit) P.off_rp = (unsigned)off;
    int* rp = rp_fit ? reinterpret_cast<int*>(smem + carve(sizeof(int) * (nr + 1))) : A.rp_scratch + r0 + blockIdx.x;
    const size_t vec_bytes = 4 * ((sizeof(double) * 3 * nr + 15) & ~(size_t)15) + ((sizeof(float) * 9 * nr + 15) & ~(size_t)15);
    const bool vec_fit = rp_fit && off + vec_bytes <= A.smem_bytes;
    double *rs, *ps, *ss, *ws; float* dinv;
    if (vec_fit) {
        P.off_r = (unsigned)off; rs = reinterpret_cast<double*>(smem + carve(sizeof(double) * 3 * nr));
        P.off_p = (unsigned)off; ps = reinterpret_cast<double*>(smem + carve(sizeof(double) * 3 * nr));
        P.off_s = (unsigned)off; ss = reinterpret_cast<double*>(smem + carve(sizeof(double) * 3 * nr));
        P.off_w = (unsigned)off; ws = reinterpret_cast<double*>(smem + carve(sizeof(double) * 3 * nr));
        P.off_dinv = (unsigned)off; dinv = reinterpret_cast<float*>(smem + carve(sizeof(float) * 9 * nr));
    } else {
        rs = A.r + 3 * (size_t)r0; ps = A.p + 3 * (size_t)r0; ss = A.s + 3 * (size_t)r0; ws = A.w + 3 * (size_t)r0; dinv = A.dinv + 9 * (size_t)r0;
    }
    // window of u: as many rows around the own range as fit after the matrix slice (or, when the matrix does not fit, in
    // what is left after the vectors); at least the own rows, else no window at all
    const size_t mat_bytes = ((sizeof(int) * (size_t)nb + 15) & ~(size_t)15) + ((sizeof(float) * 9 * (size_t)nb + 15) & ~(size_t)15);
    const size_t min_win = ((sizeof(double) * 3 * ((size_t)nr + 2) + 15) & ~(size_t)15);
    const bool mat_fit = !A.force_stream && vec_fit && off + mat_bytes + min_win <= A.smem_bytes;
    const int32_t* cols; const float* vals;
    if (mat_fit) {
        P.off_cols = (unsigned)off; int32_t* cs = reinterpret_cast<int32_t*>(smem + carve(sizeof(int) * (size_t)nb));
        P.off_vals = (unsigned)off; float* vs = reinterpret_cast<float*>(smem + carve(sizeof(float) * 9 * (size_t)nb));
        // (eight loads in flight per thread: one load per trip makes the 150 KB copy a chain of ~35 L2 round trips)
        for (int i = tid; i < nb; i += 8 * PCG_THREADS) {
            int32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { const int k = i + q * PCG_THREADS; v[q] = (k < nb) ? __ldg(A.cols + b0 + k) : 0; }
#pragma unroll
            for (int q = 0; q < 8; q++) { const int k = i + q * PCG_THREADS; if (k < nb) cs[k] = v[q]; }
        }
        for (int i = tid; i < 9 * nb; i += 8 * PCG_THREADS) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { const int k = i + q * PCG_THREADS; v[q] = (k < 9 * nb) ? __ldg(A.vals + 9 * b0 + k) : 0.0f; }
#pragma unroll
            for (int q = 0; q < 8; q++) { const int k = i + q * PCG_THREADS; if (k < 9 * nb) vs[k] = v[q]; }
        }
        cols = cs; vals = vs;
    } else {
        cols = A.cols + b0; vals = A.vals + 9 * b0;
    }
    // streamed matrix: two tile buffers, if they fit next to the vectors and a window of at least the own rows
    bool tile_fit = false;
    P.off_tile[0] = P.off_tile[1] = 0;
    if (A.tiled && !mat_fit && vec_fit && off + 2 * (size_t)PCG_TILE_BYTES + min_win <= A.smem_bytes) {
        tile_fit = true;
        P.off_tile[0] = (unsigned)carve(PCG_TILE_BYTES);
        P.off_tile[1] = (unsigned)carve(PCG_TILE_BYTES);
    }
    int w0 = r0, nwin = 0;     // window = block rows [w0, w0 + nwin)
    double* uwin = nullptr;
    if (vec_fit && off + min_win <= A.smem_bytes) {
        const int cap_rows = (int)(((size_t)A.smem_bytes - off) / (sizeof(double) * 3)) & ~1;   // even: 16 B granularity of the bulk copy
        const int half = (cap_rows - nr) / 2;
        w0 = max(0, r0 - half) & ~1;
        int w1 = min(nbr, w0 + cap_rows);
        nwin = w1 - w0;
        P.off_win = (unsigned)off;
        uwin = reinterpret_cast<double*>(smem + off);
    }
    const bool own_in_win = nwin > 0;   // by construction the window then contains [r0, r0 + nr)
    // the bulk copy moves multiples of 16 B: with an odd row count (only possible at the very end of the vector) it also
    // copies the 8 B of padding behind u (the buffer is allocated with that slack)
    const unsigned win_bytes = (unsigned)((sizeof(double) * 3 * (size_t)nwin + 15) & ~(size_t)15);
    for (int i = tid; i <= nr; i += PCG_THREADS) rp[i] = (int)(A.rows[r0 + i] - b0);
    if (tid == 0) s_n_long = 0;
    __syncthreads();
    for (int i = tid; i < nr; i += PCG_THREADS)
        if (rp[i + 1] - rp[i] > LONG_ROW) {
            const int k = atomicAdd(&s_n_long, 1);
            if (k < MAX_LONG_ROWS) s_long[k] = i;
        }
    __syncthreads();
    const int n_long = min(s_n_long, MAX_LONG_ROWS);
    // cut the listed long rows into segments of equal length (at most MAX_SEGS in all); rows in ascending order, so that the
    // segments are ordered by their first block (the streamed product finds a tile's segments by bisection)
    if (tid == 0) {
        for (int i = 1; i < n_long; i++) {
            const int v = s_long[i];
            int k = i - 1;
            while (k >= 0 && s_long[k] > v) { s_long[k + 1] = s_long[k]; k--; }
            s_long[k + 1] = v;
        }
        long long total = 0;
        for (int q = 0; q < n_long; q++) total += rp[s_long[q] + 1] - rp[s_long[q]];
        int seg_len = MIN_SEG;
        if (n_long) seg_len = max(MIN_SEG, (int)((total + (MAX_SEGS - n_long) - 1) / (MAX_SEGS - n_long)));
        int n = 0;
        for (int q = 0; q < n_long; q++) {
            s_seg_first[q] = n;
            const int lr = s_long[q];
            for (int j = rp[lr]; j < rp[lr + 1]; j += seg_len) { s_seg_j0[n] = j; s_seg_j1[n] = min(j + seg_len, rp[lr + 1]); n++; }
        }
        s_seg_first[n_long] = n;
        s_n_seg = n;
    }
    __syncthreads();
    P.n_seg = s_n_seg; P.s_seg_j0 = s_seg_j0; P.s_seg_j1 = s_seg_j1; P.s_seg_first = s_seg_first; P.s_seg_y = s_seg_y;
    P.r0 = r0; P.nr = nr; P.w0 = w0; P.nwin = nwin; P.n_long = n_long; P.own_in_win = own_in_win; P.win_bytes = win_bytes;
    P.rp = rp; P.rs = rs; P.ps = ps; P.ss = ss; P.ws = ws; P.dinv = dinv; P.cols = cols; P.vals = vals; P.uwin = uwin;
    P.s = s; P.bc2 = bc2; P.s_long = s_long; P.mbar = &s_mbar;
    P.t_start = t_start; P.t_loaded = global_ns();
    P.b0 = b0; P.tbar = s_tbar; P.s_abort = &s_abort;
    if (rp_fit && vec_fit && mat_fit && own_in_win) pcg_body<1, DIST>(A, P);
    else if (rp_fit && vec_fit && tile_fit && own_in_win) pcg_body<3, DIST>(A, P);
    else if (rp_fit && vec_fit && own_in_win) pcg_body<2, DIST>(A, P);
    else pcg_body<0, DIST>(A, P);
}

// needmask[c] |= 1 << q for every rank q != rank whose rows reference block column c, for the columns THIS rank owns (the
// pattern is replicated, so every rank derives its send lists locally).  One warp per block row.
__global__ void k_dist_needmask(const unsigned long long* __restrict__ rows, const int32_t* __restrict__ cols, int nbr, unsigned long long nnzb,
                                int world, int rank, int grid, unsigned char* __restrict__ needmask)
{
    __shared__ int bound[DIST_MAX_WORLD + 1];
    if (threadIdx.x <= world) {
        const unsigned long long cost = nnzb + ROW_COST * (unsigned long long)nbr;
        const int VG = world * grid;
        bound[threadIdx.x] = (threadIdx.x == world) ? nbr : row_lower_bound(rows, nbr, (cost * (unsigned long long)(threadIdx.x * grid)) / (unsigned long long)VG);
    }
    __syncthreads();
    const int lo = bound[rank], hi = bound[rank + 1];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nbr; i += warps) {
        if (i >= lo && i < hi) continue;   // own rows read own columns locally
        int q = 0;
        while (q + 1 < world && i >= bound[q + 1]) q++;
        for (unsigned long long j = rows[i] + (threadIdx.x & 31); j < rows[i + 1]; j += 32) {
            const int c = cols[j] / 3;
            if (c >= lo && c < hi) atomicOr(reinterpret_cast<unsigned*>(needmask + (c & ~3)), (1u << q) << (8 * (c & 3)));
        }
    }
}

// The BCSR blocks of this rank's rows: {rows[bound[rank]], rows[bound[rank + 1]]} (same partition rule as the solve).  A rank whose
// next solve is shared only ever reads its own rows of the matrix, so the numeric assembly can skip the others (assembly.cu).
__global__ void k_dist_own_blocks(const unsigned long long* __restrict__ rows, int nbr, unsigned long long nnzb, int world, int rank, int grid,
                                  unsigned long long* __restrict__ out2)
{
    if (threadIdx.x < 2) {
        const unsigned long long cost = nnzb + ROW_COST * (unsigned long long)nbr;
        const int q = rank + (int)threadIdx.x;
        const int r = (q == world) ? nbr : row_lower_bound(rows, nbr, (cost * (unsigned long long)(q * grid)) / (unsigned long long)(world * grid));
        out2[threadIdx.x] = rows[r];
    }
}
// Will the solve that follows an assembly of this pattern be shared by all ranks?  If so, queue the computation of this rank's block
// range into d_range2 and return true (the caller then assembles only that range).
bool dist_own_rows_only(sb_context* ctx, const unsigned long long* rows, int nbr, size_t nnzb, unsigned long long* d_range2)
{
    Dist* D = ctx->dist;
    Pcg* P = ctx->pcg;
    if (!D || !D->connected || !D->enabled || D->world <= 1 || !P || P->grid == 0) return false;
    static const bool off = std::getenv("SB_DIST_FULL_ASSEMBLY") != nullptr;   // diagnostic hook
    if (off) return false;
    static const bool always = std::getenv("SB_DIST_POLICY") && std::string(std::getenv("SB_DIST_POLICY")) == "always";
    const double resident1 = 1.08 * (40.0 * (double)nnzb + 160.0 * (double)nbr) / P->grid;
    if (!always && resident1 <= (double)P->smem_bytes) return false;           // the policy keeps this solve local: it needs the whole matrix
    k_dist_own_blocks<<<1, 32, 0, ctx->stream>>>(rows, nbr, (unsigned long long)nnzb, D->world, D->rank, P->grid, d_range2);
    ctx->launches++;
    return true;
}

// ---- rank 0 is authoritative: broadcast of a vector + a few scalars from rank 0 to every rank, in stream order ----
// Every rank drives the same scene, but FP64 atomics (gradient scatter, contact tables filled in arrival order) make the last bits
// of the gradient and of the energy differ from rank to rank; left alone, the replicas drift apart (measurably within tens of
// time steps on the cloth scenes) and sooner or later take different decisions -- a different number of shared solves is a
// dead-lock.  So after every evaluation all ranks adopt rank 0's gradient, energy and residual (and, after a solve the policy kept
// local, rank 0's du): the replicas stay bitwise identical.  Two buffers alternate; rank 0 waits for the acknowledgements of
// broadcast k - 2 before it overwrites that buffer.
constexpr int BCAST_CTAS = 128;   // (enough CTAs to keep rank 0's NVLink egress busy: it stores the vector W - 1 times)
struct BcastArgs {
    int world, n, n_scal;
    unsigned long long k;                      // number of this broadcast (1, 2, ...)
    unsigned long long timeout_ns;
    double* buf[DIST_MAX_WORLD];               // this broadcast's buffer on every rank: [vector | scalars]
    unsigned long long* arrive[DIST_MAX_WORLD];
    unsigned long long* ack_root;              // on rank 0
    unsigned long long* done_local;
    int* abort_flag;
};
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long target, unsigned long long timeout_ns, int* abort_flag)
{
    unsigned long long v, t0 = 0;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= target) return true;
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = global_ns();
            if (!t0) t0 = now;
            if (now - t0 > timeout_ns || *(volatile int*)abort_flag) { *(volatile int*)abort_flag = 1; return false; }
        }
    }
}
__global__ void __launch_bounds__(256) k_bcast_root(const double* __restrict__ vec, const double* __restrict__ scal, const BcastArgs a)
{
    __shared__ int ok;
    if (threadIdx.x == 0) ok = spin_until(a.ack_root, a.k >= 2 ? (a.k - 2) * (unsigned long long)(a.world - 1) : 0ull, a.timeout_ns, a.abort_flag) ? 1 : 0;
    __syncthreads();
    if (ok) {
        for (int q = 1; q < a.world; q++) {
            for (int i = blockIdx.x * 256 + threadIdx.x; i < a.n; i += gridDim.x * 256) __stcg(a.buf[q] + i, vec[i]);
            if (blockIdx.x == 0 && threadIdx.x < a.n_scal) __stcg(a.buf[q] + a.n + threadIdx.x, scal[threadIdx.x]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && ok) {
        __threadfence_system();
        for (int q = 1; q < a.world; q++) asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(a.arrive[q]), "l"(1ull) : "memory");
    }
}
__global__ void __launch_bounds__(256) k_bcast_peer(double* __restrict__ vec, double* __restrict__ scal, int rank, const BcastArgs a)
{
    __shared__ int ok;
    if (threadIdx.x == 0) ok = spin_until(a.arrive[rank], a.k * (unsigned long long)gridDim.x, a.timeout_ns, a.abort_flag) ? 1 : 0;
    __syncthreads();
    if (ok) {
        const double* src = a.buf[rank];
        for (int i = blockIdx.x * 256 + threadIdx.x; i < a.n; i += gridDim.x * 256) vec[i] = __ldcg(src + i);
        if (blockIdx.x == 0 && threadIdx.x < a.n_scal) scal[threadIdx.x] = __ldcg(src + a.n + threadIdx.x);
    }
    __syncthreads();
    if (threadIdx.x == 0 && ok) {
        __threadfence();
        const unsigned long long done = atomicAdd(a.done_local, 1ull) + 1ull;
        if (done == a.k * (unsigned long long)gridDim.x) {   // the last CTA of this rank: the buffer may be reused
            __threadfence_system();
            asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(a.ack_root), "l"(1ull) : "memory");
        }
    }
}

// a peer did not show up in time (set by any kernel that waits on another GPU): everything received since is void
bool dist_aborted(sb_context* ctx)
{
    return ctx->dist && ctx->dist->d_abort && *(volatile int*)ctx->dist->d_abort != 0;
}

// vec[0 .. n) and scal[0 .. n_scal) of rank 0 replace those of every other rank (queued on the context stream; n may be 0)
int dist_bcast_from_root(sb_context* ctx, double* vec, int n, double* scal, int n_scal)
{
    Dist* D = ctx->dist;
    if (!D || !D->connected || !D->enabled || D->world <= 1) return 0;
    if ((size_t)n > D->max_dofs || n_scal > 8) return fail(ctx, SB_ERR_STATE, "distributed broadcast: vector longer than the peer buffer");
    BcastArgs a;
    a.world = D->world; a.n = n; a.n_scal = n_scal; a.k = ++D->n_bcasts;
    static const double timeout_s = std::getenv("SB_DIST_TIMEOUT_S") ? std::atof(std::getenv("SB_DIST_TIMEOUT_S")) : 30.0;
    a.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    const size_t half = sizeof(double) * (D->max_dofs + 8);
    for (int q = 0; q < D->world; q++) {
        a.buf[q] = reinterpret_cast<double*>(D->base[q] + D->off_bcast + (a.k & 1ull) * half);
        a.arrive[q] = reinterpret_cast<unsigned long long*>(D->base[q] + 64);
    }
    a.ack_root = reinterpret_cast<unsigned long long*>(D->base[0] + 128);
    a.done_local = reinterpret_cast<unsigned long long*>(D->base[D->rank] + 192);
    a.abort_flag = D->d_abort;
    if (D->rank == 0) k_bcast_root<<<BCAST_CTAS, 256, 0, ctx->stream>>>(vec, scal, a);
    else k_bcast_peer<<<BCAST_CTAS, 256, 0, ctx->stream>>>(vec, scal, D->rank, a);
    ctx->launches++;
    return 0;
}

int solve_pcg_internal(sb_context* ctx, double abs_tol, double rel_tol, int max_iter, int stop_on_indef,
                       int* out_iterations, int* out_ok, double* out_du_dot_grad, double* out_du_inf)
{
    int nbr; size_t nnzb; const unsigned long long* rows; const int32_t* cols; const float* vals;
    int r = bcsr_view(ctx, &nbr, &nnzb, &rows, &cols, &vals);
    if (r) return r;
    if (3 * nbr != ctx->ndofs) return fail(ctx, SB_ERR_STATE, "sb_solve_pcg: matrix and DoF vector sizes differ");
    StageTimer timer(ctx, ST_PCG);
    Pcg* P = get(ctx);
    cudaStream_t st = ctx->stream;
    const int n = ctx->ndofs;
    P->r.ensure(n); P->p.ensure(n); P->s.ensure(n); P->w.ensure(n); P->u.ensure(n + 2); P->u4.ensure(4 * (size_t)nbr); P->x.ensure(n); P->dinv.ensure(9 * (size_t)nbr);
    P->part.ensure(3 * PCG_MAX_BLOCKS);
    P->rp_scratch.ensure((size_t)nbr + PCG_MAX_BLOCKS + 1);
    ctx->du.ensure(n);
    if (P->grid == 0) {
        int dev = 0, sms = 0, coop = 0, occ = 0, smem_max = 0;
        SB_CUDA(ctx, cudaGetDevice(&dev));
        SB_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SB_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        SB_CUDA(ctx, cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (!coop) return fail(ctx, SB_ERR_CUDA, "sb_solve_pcg: the device does not support cooperative launches");
        cudaFuncAttributes fa;
        SB_CUDA(ctx, cudaFuncGetAttributes(&fa, k_pcg_solve<false>));
        P->smem_bytes = (unsigned)(smem_max - (int)fa.sharedSizeBytes - 1024);
        // test hook: SB_PCG_SMEM_LIMIT=<bytes> shrinks the budget so that small fixtures exercise the streaming (slices in global
        // memory) and partial-window paths that million-tet scenes take
        if (const char* lim = std::getenv("SB_PCG_SMEM_LIMIT")) {
            const long v = std::atol(lim);
            if (v >= 1024 && (unsigned)v < P->smem_bytes) P->smem_bytes = (unsigned)v;
        }
        SB_CUDA(ctx, cudaFuncSetAttribute(k_pcg_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_bytes));
        SB_CUDA(ctx, cudaFuncSetAttribute(k_pcg_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_bytes));
        SB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pcg_solve<false>, PCG_THREADS, P->smem_bytes));
        if (occ < 1) return fail(ctx, SB_ERR_CUDA, "sb_solve_pcg: the persistent kernel does not fit on an SM");
        P->grid = std::min(sms, PCG_MAX_BLOCKS / DIST_MAX_WORLD);
        // test hook: SB_PCG_GRID=<ctas> shrinks the grid (two distributed solves then fit on ONE GPU side by side)
        if (const char* g = std::getenv("SB_PCG_GRID")) {
            const int v = std::atoi(g);
            if (v >= 1 && v < P->grid) P->grid = v;
        }
    }
    PcgArgs A;
    A.rows = rows; A.cols = cols; A.vals = vals; A.grad = ctx->grad.p; A.dinv = P->dinv.p;
    A.x = P->x.p; A.r = P->r.p; A.p = P->p.p; A.s = P->s.p; A.w = P->w.p; A.u = P->u.p; A.u4 = P->u4.p; A.du = ctx->du.p;
    A.part = P->part.p; A.rp_scratch = P->rp_scratch.p; A.barrier = P->d_barrier; A.result = P->h_result;   // (pinned host memory, UVA)
    A.seq = ++P->seq;
    // Shared memory of this launch.  When the slices fit (per-CTA estimate; the rows are cut by cost, see ROW_COST) the kernel
    // takes everything and the matrix stays resident.  Otherwise the matrix is read with ordinary loads every iteration, and a
    // full carve-out would leave no L1: every 4-byte load of a block would be its own L2 request (the SM's miss path takes ~2
    // cycles per request).  Such solves run with a smaller carve-out, so that the nine loads of a 36-byte block share one or two
    // L1 line fills.  (The experimental tiled mode needs no L1 and takes everything again.)
    Dist* DS = (ctx->dist && ctx->dist->connected && ctx->dist->enabled && ctx->dist->world > 1) ? ctx->dist : nullptr;
    bool kept_local = false;   // connected, but this solve stays on the own GPU: every rank then adopts rank 0's du
    // Policy (SB_DIST_POLICY=auto|always, default auto): a matrix that is resident in ONE GPU's shared memory gains nothing from
    // more GPUs -- its iteration is bound by grid-wide synchronisation, and a synchronisation across NVLink costs about four times
    // an on-chip one -- so such solves stay local (every rank solves its own replica, results identical); systems that do not fit
    // (million-tet bars, 66 k-node cloth) are shared by all ranks.  Every rank takes the same decision from the same pattern sizes.
    if (DS) {
        static const bool always = std::getenv("SB_DIST_POLICY") && std::string(std::getenv("SB_DIST_POLICY")) == "always";
        const double resident1 = 1.08 * (40.0 * (double)nnzb + 160.0 * (double)nbr) / P->grid;
        if (!always && resident1 <= (double)P->smem_bytes) { DS->n_local_solves++; DS = nullptr; kept_local = true; }
    }
    if (DS && (size_t)n > DS->max_dofs) return fail(ctx, SB_ERR_STATE, "sb_solve_pcg: the system has more DoFs than sb_dist_init reserved peer memory for");
    const int vgrid = P->grid * (DS ? DS->world : 1);   // CTAs of all ranks
    unsigned smem_launch = P->smem_bytes;
    {
        const double resident = 1.08 * (40.0 * (double)nnzb + 160.0 * (double)nbr) / vgrid;
        const double rows_max = 1.6 * (double)nbr / vgrid;                                 // (sparse-row slices hold more rows)
        const double tiled = 140.0 * rows_max + 2.0 * PCG_TILE_BYTES + 24.0 * (rows_max + 2.0) + 64.0;
        static const bool tiled_on = std::getenv("SB_PCG_TILED") != nullptr;
        if (resident > (double)P->smem_bytes && (!tiled_on || tiled > (double)P->smem_bytes)) smem_launch = std::min(P->smem_bytes, PCG_STREAM_SMEM);
    }
    if (smem_launch != P->smem_launch_last) {
        const int pct = (int)std::min<long>(100, (100L * (smem_launch + 16 * 1024)) / (228 * 1024) + 1);
        SB_CUDA(ctx, cudaFuncSetAttribute(k_pcg_solve<false>, cudaFuncAttributePreferredSharedMemoryCarveout, smem_launch == P->smem_bytes ? (int)cudaSharedmemCarveoutMaxShared : pct));
        SB_CUDA(ctx, cudaFuncSetAttribute(k_pcg_solve<true>, cudaFuncAttributePreferredSharedMemoryCarveout, smem_launch == P->smem_bytes ? (int)cudaSharedmemCarveoutMaxShared : pct));
        P->smem_launch_last = smem_launch;
    }
    A.nnzb = nnzb; A.smem_bytes = smem_launch; A.instrument = ctx->profile ? 1 : 0;
    // test hook: SB_PCG_FORCE_STREAM=1 sends scenes that would be resident through the streamed-tile product
    static const bool force_stream = std::getenv("SB_PCG_FORCE_STREAM") != nullptr;
    A.force_stream = force_stream ? 1 : 0;
    static const bool tiled_env = std::getenv("SB_PCG_TILED") != nullptr;
    A.tiled = tiled_env ? 1 : 0;
    A.nbr = nbr; A.abs_tol = abs_tol; A.rel_tol = rel_tol; A.max_iter = max_iter; A.stop_on_indef = stop_on_indef;
    if (P->barrier_grid != P->grid) {   // first solve (or a changed grid): start the counter at zero
        SB_CUDA(ctx, cudaMemsetAsync(P->d_barrier, 0, sizeof(unsigned), st));
        P->barrier_base = 0; P->barrier_grid = P->grid;
    }
    A.barrier_base = P->barrier_base;
    // SB_PCG_DUMP=1: per-CTA cycle counters of every solve on stderr (load-balance diagnostics)
    static const bool dump = std::getenv("SB_PCG_DUMP") != nullptr;
    static long long* d_dbg = nullptr;
    if (dump && !d_dbg) cudaMalloc(&d_dbg, (5 * PCG_MAX_BLOCKS + 8) * sizeof(long long));
    if (dump) cudaMemsetAsync(d_dbg, 0, (5 * PCG_MAX_BLOCKS + 8) * sizeof(long long), st);
    A.dbg = dump ? d_dbg : nullptr;
    if (dump) A.instrument = 1;
    A.d.world = 1; A.d.rank = 0;
    if (DS) {
        // distributed solve: every rank runs this same code on the same (replicated) matrix and right-hand side
        DistArgs& d = A.d;
        d.world = DS->world; d.rank = DS->rank; d.epoch_base = DS->epoch_base;
        static const double timeout_s = std::getenv("SB_DIST_TIMEOUT_S") ? std::atof(std::getenv("SB_DIST_TIMEOUT_S")) : 30.0;
        d.timeout_ns = (unsigned long long)(timeout_s * 1e9);
        const size_t set = (size_t)(DS->n_solves & 1ull) * 3 * PCG_MAX_BLOCKS * sizeof(double);
        for (int q = 0; q < DS->world; q++) {
            d.bar[q] = reinterpret_cast<unsigned long long*>(DS->base[q]);
            d.part[q] = reinterpret_cast<double*>(DS->base[q] + DS->off_part + set);
            d.u[q] = reinterpret_cast<double*>(DS->base[q] + DS->off_u);
            d.u4[q] = reinterpret_cast<double*>(DS->base[q] + DS->off_u4);
            d.du[q] = reinterpret_cast<double*>(DS->base[q] + DS->off_du);
            d.ll[q] = reinterpret_cast<uint4*>(DS->base[q] + DS->off_ll);
        }
        d.ll_base = DS->ll_base;
        A.u = d.u[d.rank]; A.u4 = d.u4[d.rank]; A.du = nullptr;
        DS->needmask.ensure((size_t)nbr + 8);
        SB_CUDA(ctx, cudaMemsetAsync(DS->needmask.p, 0, (size_t)nbr + 8, st));
        k_dist_needmask<<<296, 256, 0, st>>>(rows, cols, nbr, (unsigned long long)nnzb, DS->world, DS->rank, P->grid, DS->needmask.p);
        d.needmask = DS->needmask.p; d.abort_flag = DS->d_abort;
        ctx->launches += 1;
        // SB_DIST_TRACE=1: one line per shared solve and rank (the ranks' traces must be identical: the first difference names the
        // host decision that went out of step)
        static const bool trace = std::getenv("SB_DIST_TRACE") != nullptr;
        if (trace) {
            unsigned long long tb; memcpy(&tb, &abs_tol, 8);
            fprintf(stderr, "DISTTRACE rank %d solve %llu nbr %d nnzb %zu abs_tol %016llx max_iter %d epoch %llu ll %u state %llu eval %llu\n", DS->rank, DS->n_solves, nbr, nnzb, tb,
                    max_iter, DS->epoch_base, DS->ll_base, (unsigned long long)ctx->state_version, (unsigned long long)ctx->eval_id);
        }
    }
    void* args[] = {(void*)&A};
    SB_CUDA(ctx, cudaLaunchCooperativeKernel(DS ? (const void*)k_pcg_solve<true> : (const void*)k_pcg_solve<false>, dim3(P->grid), dim3(PCG_THREADS), args, smem_launch, st));
    ctx->launches += 1;
    if (DS) SB_CUDA(ctx, cudaMemcpyAsync(ctx->du.p, DS->base[DS->rank] + DS->off_du, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    if (kept_local) { const int rb = dist_bcast_from_root(ctx, ctx->du.p, n, nullptr, 0); if (rb) return rb; }
    {
        // wait for the kernel's record: poll its sequence number in pinned memory for a while (the usual solve takes 0.1-0.4 ms
        // and the round trip of a copy + stream synchronisation would leave the GPU idle for ~10 us), then block on the stream
        static const bool no_poll = std::getenv("SB_PCG_NO_POLL") != nullptr;
        volatile unsigned long long* seq = &P->h_result->seq;
        bool seen = false;
        if (!no_poll && !DS && !dump && !ctx->profile) {
            const auto t0 = std::chrono::steady_clock::now();
            for (unsigned spins = 0;; spins++) {
                if (*seq == A.seq) { seen = true; break; }
                if ((spins & 255u) == 255u && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        if (!seen) {
            SB_CUDA(ctx, hot_sync(ctx));
            SB_CUDA(ctx, cudaGetLastError());
            if (*seq != A.seq) { P->barrier_grid = 0; return fail(ctx, SB_ERR_CUDA, "sb_solve_pcg: the solve kernel ended without writing its result"); }
        }
    }
    if (!DS) P->barrier_base += (unsigned)P->h_result->barriers * (unsigned)P->grid;
    if (DS) {
        DS->epoch_base += P->h_result->barriers;
        DS->ll_base += (unsigned)P->h_result->ll_uses;
        DS->n_solves++;
        if (P->h_result->done == 4)
            return fail(ctx, SB_ERR_CUDA, "sb_solve_pcg: distributed solve aborted: a peer rank did not arrive at a barrier within SB_DIST_TIMEOUT_S (ranks out of step, or a peer failed)");
    }
    if (dump) {
        const int G = P->grid;
        std::vector<long long> h(5 * (size_t)G + 8);
        cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const int its = P->h_result->it > 0 ? P->h_result->it : 1;
        fprintf(stderr, "PCGDUMP its=%d load_ns=%lld phase0_ns=%lld loop_ns=%lld\n", P->h_result->it, (long long)(P->h_result->t_loaded - P->h_result->t_start),
                (long long)(P->h_result->t_loop - P->h_result->t_loaded), (long long)(P->h_result->t_end - P->h_result->t_loop));
        for (int b = 0; b < G; b++)
            fprintf(stderr, "PCGDUMP cta=%d rows=%lld blocks=%lld spmv=%lld wait=%lld vec=%lld win=%lld\n", b, h[4 * G + b] >> 32, h[4 * G + b] & 0xffffffffll,
                    h[b] / its, h[G + b] / its, h[2 * G + b] / its, h[3 * G + b] / its);
    }
    if (ctx->profile) {
        ctx->stage_calls[ST_CG_ITERATIONS] += P->h_result->it;
        ctx->stage_ms[ST_CG_ITERATIONS] += 1e-6 * (double)(P->h_result->t_end - P->h_result->t_loop);          // iterations + final reduction
        ctx->stage_ms[ST_PCG_SETUP] += 1e-6 * (double)(P->h_result->t_loop - P->h_result->t_start);              // slice load + preconditioner
        ctx->stage_calls[ST_PCG_SETUP]++;
        ctx->stage_calls[ST_PCG_C_SPMV] += P->h_result->c_spmv; ctx->stage_calls[ST_PCG_C_BAR] += P->h_result->c_bar;
        ctx->stage_calls[ST_PCG_C_RED] += P->h_result->c_red; ctx->stage_calls[ST_PCG_C_VEC] += P->h_result->c_vec;
        ctx->stage_calls[ST_PCG_C_WIN] += P->h_result->c_win;
    }
    if (out_iterations) *out_iterations = P->h_result->it;
    if (out_ok) *out_ok = (P->h_result->done == 1) ? 1 : 0;
    if (out_du_dot_grad) *out_du_dot_grad = P->h_result->du_dot_grad;
    if (out_du_inf) *out_du_inf = P->h_result->du_inf;
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_solve_pcg(sb_context* ctx, double abs_tol, double rel_tol, int max_iterations, int stop_on_indefiniteness,
                            int* out_iterations, int* out_ok, double* out_du_dot_grad, double* out_du_inf)
{
    if (!ctx) return SB_ERR_ARG;
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_solve_pcg: no gradient: call sb_eval(SB_EVAL_PGH) first");
    return solve_pcg_internal(ctx, abs_tol, rel_tol, max_iterations, stop_on_indefiniteness, out_iterations, out_ok, out_du_dot_grad, out_du_inf);
}

// ---- distributed solve: set-up over peer memory (include/stark_b200.h "multi-GPU") ----
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int sb_dist_init(sb_context* ctx, int rank, int world, long long max_dofs, unsigned char* out_handle64)
{
    if (!ctx || world < 1 || world > DIST_MAX_WORLD || rank < 0 || rank >= world || max_dofs < 3) return SB_ERR_ARG;
    if (ctx->dist) return fail(ctx, SB_ERR_STATE, "sb_dist_init: already initialised");
    Dist* D = new Dist();
    D->world = world; D->rank = rank; D->max_dofs = (size_t)max_dofs;
    const size_t nbr = ((size_t)max_dofs + 2) / 3;
    D->off_part = 256;
    D->off_ll = align256(D->off_part + 2 * 3 * (size_t)PCG_MAX_BLOCKS * sizeof(double));
    D->off_u = align256(D->off_ll + (size_t)PCG_MAX_BLOCKS * sizeof(uint4));
    D->off_u4 = align256(D->off_u + sizeof(double) * ((size_t)max_dofs + 4));
    D->off_du = align256(D->off_u4 + sizeof(double) * 4 * (nbr + 1));
    D->off_bcast = align256(D->off_du + sizeof(double) * ((size_t)max_dofs + 4));
    D->bytes = align256(D->off_bcast + 2 * sizeof(double) * ((size_t)max_dofs + 8));
    unsigned char* p = nullptr;
    if (cudaMalloc(&p, D->bytes) != cudaSuccess || cudaMemset(p, 0, D->bytes) != cudaSuccess || cudaMallocHost(&D->d_abort, sizeof(int)) != cudaSuccess) {
        delete D;
        return fail(ctx, SB_ERR_CUDA, "sb_dist_init: allocation of the peer buffer failed");
    }
    *D->d_abort = 0;
    D->base[rank] = p;
    if (out_handle64) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaGetLastError(); memset(out_handle64, 0, 64); }   // (single-process use connects by pointer)
        else memcpy(out_handle64, &h, 64);
    }
    cudaDeviceSynchronize();
    ctx->dist = D;
    return 0;
}

// handles: world x 64 bytes (cudaIpcMemHandle_t of every rank, own entry ignored), gathered by the caller's process group
extern "C" int sb_dist_connect(sb_context* ctx, const unsigned char* handles)
{
    if (!ctx || !handles) return SB_ERR_ARG;
    Dist* D = ctx->dist;
    if (!D) return fail(ctx, SB_ERR_STATE, "sb_dist_connect: call sb_dist_init first");
    for (int q = 0; q < D->world; q++) {
        if (q == D->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)q, 64);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(ctx, SB_ERR_CUDA, std::string("sb_dist_connect: cudaIpcOpenMemHandle failed for rank ") + std::to_string(q) + ": " + cudaGetErrorString(e));
        D->base[q] = static_cast<unsigned char*>(p);
        D->opened[q] = true;
    }
    D->connected = true;
    return 0;
}

// the same with plain device pointers (ranks living in ONE process: tests on a single GPU, or threads driving several GPUs
// with peer access enabled by the caller)
extern "C" int sb_dist_connect_ptrs(sb_context* ctx, void* const* bases)
{
    if (!ctx || !bases) return SB_ERR_ARG;
    Dist* D = ctx->dist;
    if (!D) return fail(ctx, SB_ERR_STATE, "sb_dist_connect_ptrs: call sb_dist_init first");
    for (int q = 0; q < D->world; q++)
        if (q != D->rank) { if (!bases[q]) return SB_ERR_ARG; D->base[q] = static_cast<unsigned char*>(bases[q]); }
    D->connected = true;
    return 0;
}

// switch the sharing of solves off / on again (all ranks must do it between the same two solves): with it off every rank solves
// its own replica, which is how bench.py measures the single-GPU rate of the same scene in the same run
extern "C" int sb_dist_set_enabled(sb_context* ctx, int enabled)
{
    if (!ctx) return SB_ERR_ARG;
    if (!ctx->dist) return fail(ctx, SB_ERR_STATE, "sb_dist_set_enabled: call sb_dist_init first");
    ctx->dist->enabled = enabled != 0;
    return 0;
}

extern "C" void* sb_dist_local_base(sb_context* ctx)
{
    return (ctx && ctx->dist) ? ctx->dist->base[ctx->dist->rank] : nullptr;
}

// out4 = { barriers completed, distributed solves, bytes of the peer buffer, solves kept local by the policy }
extern "C" int sb_dist_stats(sb_context* ctx, int* out_rank, int* out_world, double* out3)
{
    if (!ctx) return SB_ERR_ARG;
    const Dist* D = ctx->dist;
    if (out_rank) *out_rank = D ? D->rank : 0;
    if (out_world) *out_world = (D && D->connected) ? D->world : 1;
    if (out3) { out3[0] = D ? (double)D->epoch_base : 0.0; out3[1] = D ? (double)D->n_solves : 0.0; out3[2] = D ? (double)D->bytes : 0.0; out3[3] = D ? (double)D->n_local_solves : 0.0; }
    return 0;
}

// The partition and halo lists of a distributed solve on HOST arrays (no GPU): out_bounds[world + 1] = first block row of every
// rank, out_needmask[nbr] = for the rows of `rank`: bit q set when rank q's rows reference that block column (what the owner pushes
// to q every iteration); rows of other ranks get 0.  `grid` = CTAs per rank (148 on B200).
extern "C" int sb_dist_plan(int nbr, const unsigned long long* rows, const int32_t* cols, int world, int grid, int rank, int32_t* out_bounds, unsigned char* out_needmask)
{
    if (nbr < 0 || !rows || world < 1 || world > DIST_MAX_WORLD || grid < 1 || rank < 0 || rank >= world || !out_bounds) return SB_ERR_ARG;
    const unsigned long long nnzb = rows[nbr];
    const unsigned long long cost = nnzb + 4ull * (unsigned long long)nbr;
    const int VG = world * grid;
    for (int q = 0; q <= world; q++)
        out_bounds[q] = (q == world) ? nbr : host_row_lower_bound(rows, nbr, (cost * (unsigned long long)(q * grid)) / (unsigned long long)VG);
    if (out_needmask) {
        memset(out_needmask, 0, (size_t)nbr);
        const int lo = out_bounds[rank], hi = out_bounds[rank + 1];
        int q = 0;
        for (int i = 0; i < nbr; i++) {
            while (q + 1 < world && i >= out_bounds[q + 1]) q++;
            if (i >= lo && i < hi) continue;
            for (unsigned long long j = rows[i]; j < rows[i + 1]; j++) {
                const int c = cols[j] / 3;
                if (c >= lo && c < hi) out_needmask[c] |= (unsigned char)(1u << q);
            }
        }
    }
    return 0;
}
