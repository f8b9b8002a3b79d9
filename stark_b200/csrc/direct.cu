// Direct solve of the Newton system: sparse Cholesky (L L^T) in FP64 on a tile envelope.
//
// Replaces the DirectLLT branch of NewtonsMethod::_solve_linear_system (symx/solver/NewtonsMethod.cpp:395-418):
// BlockedSparseMatrix::to_triplets (bsm/BlockedSparseMatrix.h:1365-1393) -> Eigen::SimplicialLLT (serial, re-analysed at
// every call) -> solve; "factorisation failed" (matrix not positive definite) is reported as a failed solve, which makes
// the Newton driver project more Hessians, exactly like `solver.info() != Eigen::Success`.
//
// The reference uses this path for its unit tests and small scenes (at 27 k DoFs it already needs 6.3 s per solve,
// BASELINE.md section 2).  Here:
//   * ordering (host, only when the pattern leaves the cached envelope): reverse Cuthill-McKee on the graph of 3x3 blocks
//     from a pseudo-peripheral start; block rows far denser than the rest (a rigid body touched by hundreds of contacts)
//     are taken out of the search and ordered last (an "arrowhead": they cost one full-width tile row each instead of
//     widening the whole band);
//   * storage: 64 x 64 FP64 tiles of the lower triangle, per tile row I the contiguous run of tiles F(I) .. I (row
//     envelope; Cholesky fill stays inside it).  A mesh of n nodes with a separator of s nodes needs ~ n * 6 s doubles
//     (C4, 27 k DoFs: 0.6 GB instead of 5.8 GB dense; C5, 524 k DoFs: ~25 GB -- HBM3e holds it);
//   * numeric (device): right-looking over tile columns k: POTRF of the diagonal tile, TRSM of the tiles (I, k) of the rows
//     whose envelope reaches column k (list built at analysis time), SYRK / GEMM update of the tile pairs of that list
//     (one CTA per tile, 8 x 4 register blocking from transposed shared-memory panels).  FP64 throughout (the FP64 pipe: B200
//     has no FP64 path through tcgen05, and its DMMA rate equals the DFMA rate);
//   * triangular solves tile row by tile row with the same lists; du is returned in the caller's DoF order.
// The analysis is cached: a matrix whose blocks all fall inside the cached envelope (the usual case from one Newton iteration
// to the next, also when a few contact pairs changed) is factorised without any host work but one flag read-back.
#include "internal.h"
#include <algorithm>
#include <cstdlib>
#include <numeric>

namespace sb {

int bcsr_view(sb_context* ctx, int* nbr, size_t* nnzb, const unsigned long long** rows, const int32_t** cols, const float** vals);

constexpr int NB = 64;                 // tile size
constexpr size_t TILE = (size_t)NB * NB;

struct Direct {
    // analysis (cached)
    int nbr = 0, nt = 0;
    bool have_order = false;
    DevBuf<int32_t> perm;              // block row -> position in the elimination order
    DevBuf<int32_t> F, F_new;          // first tile of every tile row (cached envelope / envelope of the current matrix)
    DevBuf<unsigned long long> off;    // tile offset of every tile row's run
    DevBuf<int32_t> act_ptr, act;      // per tile column k: the tile rows I > k with F(I) <= k (ascending)
    std::vector<int32_t> h_F, h_act_ptr, h_act;
    std::vector<int32_t> h_perm;
    size_t n_tiles = 0, n_tiles_at_order = 0;
    // numeric
    DevBuf<double> T;                  // the tiles
    DevBuf<double> W;                  // inverse of every diagonal tile's factor
    DevBuf<double> z;                  // intermediate of the triangular solves
    DevBuf<double> y;                  // [np] right-hand side / solution (elimination order)
    int* d_flags = nullptr;            // [0] factorisation failed, [1] pattern outside the cached envelope
    int* h_flags = nullptr;
    double* h_out = nullptr;           // du.grad, |du|_inf
    // statistics of the last solve (sb_llt_stats)
    double last_order_ms = 0, last_factor_ms = 0;
    int64_t n_orderings = 0, n_analyses = 0;
};
void direct_destroy(sb_context* ctx)
{
    Direct* D = ctx->direct;
    if (!D) return;
    D->perm.release(); D->F.release(); D->F_new.release(); D->off.release(); D->act_ptr.release(); D->act.release();
    D->T.release(); D->W.release(); D->z.release(); D->y.release();
    if (D->d_flags) cudaFree(D->d_flags);
    if (D->h_flags) cudaFreeHost(D->h_flags);
    if (D->h_out) cudaFreeHost(D->h_out);
    delete D;
    ctx->direct = nullptr;
}

// ---------------------------------------------------------------- ordering (host)

// Reverse Cuthill-McKee over the block graph; rows with more than `dense_deg` blocks are ordered last.
static void rcm_order(int nbr, const std::vector<unsigned long long>& rows, const std::vector<int32_t>& cols, std::vector<int32_t>& perm)
{
    std::vector<int32_t> deg(nbr);
    double avg = 0;
    for (int i = 0; i < nbr; i++) { deg[i] = (int32_t)(rows[i + 1] - rows[i]); avg += deg[i]; }
    avg /= std::max(nbr, 1);
    const int dense_deg = std::max(96, (int)(8 * avg));
    std::vector<uint8_t> state(nbr, 0);   // 0 unvisited, 1 visited, 2 dense (ordered last)
    int n_dense = 0;
    for (int i = 0; i < nbr; i++) if (deg[i] > dense_deg) { state[i] = 2; n_dense++; }
    std::vector<int32_t> order; order.reserve(nbr);
    std::vector<int32_t> level, next, nb;
    std::vector<int32_t> mark(nbr, -1);
    int stamp = 0;
    // BFS from s over unvisited sparse rows without committing; returns the last level's minimum-degree node and the depth
    auto probe = [&](int s, int& depth) {
        stamp++;
        level.assign(1, s); mark[s] = stamp; depth = 0;
        int last_best = s;
        while (!level.empty()) {
            next.clear();
            int best = level[0];
            for (int v : level) {
                if (deg[v] < deg[best]) best = v;
                for (unsigned long long j = rows[v]; j < rows[v + 1]; j++) {
                    const int w = cols[j] / 3;
                    if (state[w] == 0 && mark[w] != stamp) { mark[w] = stamp; next.push_back(w); }
                }
            }
            last_best = best;
            depth++;
            level.swap(next);
        }
        return last_best;
    };
    // seeds in ascending degree
    std::vector<int32_t> seeds(nbr);
    std::iota(seeds.begin(), seeds.end(), 0);
    std::stable_sort(seeds.begin(), seeds.end(), [&](int a, int b) { return deg[a] < deg[b]; });
    for (int s0 : seeds) {
        if (state[s0] != 0) continue;
        int s = s0, depth = 0, d2 = 0;
        int e = probe(s, depth);
        for (int it = 0; it < 3; it++) {
            int e2 = probe(e, d2);
            if (d2 <= depth) break;
            s = e; e = e2; depth = d2;
        }
        // Cuthill-McKee from s
        size_t head = order.size();
        order.push_back(s); state[s] = 1;
        while (head < order.size()) {
            const int v = order[head++];
            nb.clear();
            for (unsigned long long j = rows[v]; j < rows[v + 1]; j++) {
                const int w = cols[j] / 3;
                if (state[w] == 0) { state[w] = 1; nb.push_back(w); }
            }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    perm.assign(nbr, 0);
    const int n_sparse = (int)order.size();
    for (int p = 0; p < n_sparse; p++) perm[order[p]] = n_sparse - 1 - p;   // reversed
    int q = n_sparse;
    for (int i = 0; i < nbr; i++) if (state[i] == 2) perm[i] = q++;
    (void)n_dense;
}

// ---------------------------------------------------------------- analysis kernels

// F_new[I] = min tile column of the blocks of tile row I under the ordering (lower triangle of the symmetric matrix)
__global__ void k_envelope(const unsigned long long* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ perm,
                           int32_t* __restrict__ F_new, int nbr)
{
    const int br = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (br >= nbr) return;
    const int pi = perm[br];
    int lo = pi;
    for (unsigned long long j = rows[br] + (threadIdx.x & 31); j < rows[br + 1]; j += 32) {
        const int pj = perm[cols[j] / 3];
        if (pj < lo) lo = pj;
    }
    for (int o = 16; o > 0; o >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if ((threadIdx.x & 31) == 0) {
        // scalar rows 3 pi .. 3 pi + 2 reach scalar column 3 lo: the block (also the diagonal one) may straddle two tile rows
        const int tj = (3 * lo) / NB;
        atomicMin(&F_new[(3 * pi) / NB], tj);
        atomicMin(&F_new[(3 * pi + 2) / NB], tj);
    }
}
__global__ void k_iota_tiles(int32_t* __restrict__ F, int nt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) F[i] = i;
}
__global__ void k_envelope_inside(const int32_t* __restrict__ F_new, const int32_t* __restrict__ F, int nt, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt && F_new[i] < F[i]) flags[1] = 1;
}

// ---------------------------------------------------------------- numeric kernels

__device__ __forceinline__ double* tile_ptr(double* T, const unsigned long long* __restrict__ off, const int32_t* __restrict__ F, int I, int J)
{
    return T + (off[I] + (unsigned long long)(J - F[I])) * TILE;
}

// scatter the BCSR blocks into the tiles (lower triangle in elimination order)
__global__ void k_fill_tiles(const unsigned long long* __restrict__ rows, const int32_t* __restrict__ cols, const float* __restrict__ vals,
                             const int32_t* __restrict__ perm, double* __restrict__ T, const unsigned long long* __restrict__ off,
                             const int32_t* __restrict__ F, int nbr)
{
    const int br = blockIdx.x;
    if (br >= nbr) return;
    const int g = threadIdx.x / 9, k = threadIdx.x % 9;   // 32 groups of nine threads, one BCSR block per group and trip
    const int r = k % 3, c = k / 3;                        // column-major inside the block
    const int pi = perm[br];
    for (unsigned long long j = rows[br] + g; j < rows[br + 1]; j += 32) {
        const int gi = 3 * pi + r, gj = 3 * perm[cols[j] / 3] + c;
        if (gi >= gj) tile_ptr(T, off, F, gi / NB, gj / NB)[(gi % NB) * NB + (gj % NB)] = (double)vals[9 * j + k];
    }
}
__global__ void k_pad_diagonal(double* __restrict__ T, const unsigned long long* __restrict__ off, const int32_t* __restrict__ F, int n, int np)
{
    const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) tile_ptr(T, off, F, i / NB, i / NB)[(i % NB) * NB + (i % NB)] = 1.0;
}

// Cholesky of the diagonal tile k and the inverse of its factor; flags[0] = 1 on a non-positive pivot.
// One CTA of 256 threads.  Phase 1 (factor): the tile lives in registers, thread (ty, tx) of a 16 x 16 grid owns the 4 x 4
// entries (ty + 16 a, tx + 16 b); step j: the owners of column j publish it through a double-buffered shared column (ONE
// barrier per step), everybody scales by 1 / d (rsqrt(d)^2) and applies the rank-1 update to its registers.  Entries above
// the diagonal carry garbage that never reaches a lower entry.  Phase 2 (inverse W = L^-1, needed as a GEMM operand by the
// panel kernel and as a mat-vec operand by the triangular solves -- substitution would be a 2,016-long dependent chain per
// row): column c of W by forward substitution in a 4-lane group, the column's entries in the lanes' registers (m = 4 s + q
// in lane q), dot products split over the four lanes and closed with two shuffles.
__global__ void __launch_bounds__(256) k_potrf_inv(double* __restrict__ T, const unsigned long long* __restrict__ off, const int32_t* __restrict__ F,
                                                   double* __restrict__ Winv, int k, int* __restrict__ flags)
{
    __shared__ double L[NB][NB + 1];
    __shared__ double colbuf[2][NB];
    __shared__ double invd[NB];
    if (flags[0]) return;
    double* base = tile_ptr(T, off, F, k, k);
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    double s[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int row = ty + 16 * a, col = tx + 16 * b;
            s[a][b] = (col <= row) ? base[row * NB + col] : 0.0;
        }
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int bj = j / 16, txj = j % 16;
        double* cb = colbuf[j & 1];
        if (tx == txj) {
#pragma unroll
            for (int a = 0; a < 4; a++) cb[ty + 16 * a] = s[a][bj];
        }
        __syncthreads();
        const double d = cb[j];
        if (!(d > 0.0)) { if (threadIdx.x == 0) flags[0] = 1; return; }   // shared value: uniform exit
        const double rs = rsqrt(d);
        const double rinv = rs * rs;
        if (tx == txj) {
#pragma unroll
            for (int a = 0; a < 4; a++) { const int row = ty + 16 * a; L[row][j] = (row >= j) ? s[a][bj] * rs : 0.0; }
            if (ty == 0) invd[j] = rs;
        }
        double ra[4], cbv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) ra[a] = cb[ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; b++) cbv[b] = cb[tx + 16 * b] * rinv;
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (tx + 16 * b > j) {
#pragma unroll
                for (int a = 0; a < 4; a++) s[a][b] -= ra[a] * cbv[b];
            }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NB * NB; t += 256) { const int i = t / NB, j = t % NB; if (j <= i) base[t] = L[i][j]; }
    // inverse: column c in the 4-lane group c
    const int c = threadIdx.x / 4, q = threadIdx.x % 4;
    double w[NB / 4];
#pragma unroll
    for (int u = 0; u < NB / 4; u++) w[u] = 0.0;
#pragma unroll
    for (int i = 0; i < NB; i++) {
        double sum = 0.0;
#pragma unroll
        for (int u = 0; u < (i + 3) / 4; u++) {
            const int m = 4 * u + q;
            if (m < i) sum += L[i][m] * w[u];      // w[u] = 0 for m < c
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const double val = (i < c) ? 0.0 : ((i == c) ? invd[i] : -sum * invd[i]);
        if (q == i % 4) w[i / 4] = val;
    }
    double* W = Winv + (size_t)k * TILE;
#pragma unroll
    for (int u = 0; u < NB / 4; u++) W[(4 * u + q) * NB + c] = w[u];
}

// One kernel for the two tile products of the factorisation (one CTA of 128 threads per output tile; both operands are
// transposed into shared memory ([q][row], pitch 66 doubles) so that a thread reads its 8 + 4 operands of one q with six
// 128-bit loads and issues 32 DFMA on them):
//   PANEL:  T(I, k) <- T(I, k) W_k^T            for the active rows I of column k   (the TRSM, as a product with L_kk^-1)
//   UPDATE: T(I, J) -= T(I, k) T(J, k)^T        for the pairs J <= I of the active rows of column k   (SYRK / GEMM)
constexpr int SP = NB + 2;
template<bool PANEL>
__global__ void __launch_bounds__(128) k_tile_product(double* __restrict__ T, const unsigned long long* __restrict__ off, const int32_t* __restrict__ F,
                                                      const double* __restrict__ Winv, const int32_t* __restrict__ act, int m, int k, const int* __restrict__ flags)
{
    if (flags[0]) return;
    extern __shared__ __align__(16) double sm[];
    double* Pa = sm;                 // [NB][SP]
    double* Pb = sm + NB * SP;
    int I, J;
    const double *pa, *pb;
    if (PANEL) {
        I = act[blockIdx.x]; J = k;
        pa = tile_ptr(T, off, F, I, k);
        pb = Winv + (size_t)k * TILE;
    } else {
        // linear index -> (a, b), b <= a, in the lower triangle of the m x m pair matrix
        const int idx = blockIdx.x;
        int a = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= idx) a++;
        while (a * (a + 1) / 2 > idx) a--;
        const int b = idx - a * (a + 1) / 2;
        if (a >= m) return;
        I = act[a]; J = act[b];
        pa = tile_ptr(T, off, F, I, k);
        pb = tile_ptr(T, off, F, J, k);
    }
    {
        // 16 loads in flight per thread and operand, then the transposing stores
        double va[8], vb[8];
#pragma unroll
        for (int h = 0; h < 4; h++) {
#pragma unroll
            for (int u = 0; u < 8; u++) { const int t = threadIdx.x + 128 * (8 * h + u); va[u] = pa[t]; vb[u] = pb[t]; }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int t = threadIdx.x + 128 * (8 * h + u);
                const int r = t / NB, q = t % NB;
                Pa[q * SP + r] = va[u];
                Pb[q * SP + r] = vb[u];
            }
        }
    }
    __syncthreads();
    const int tr = (threadIdx.x / 16) * 8, tc = (threadIdx.x % 16) * 4;
    double acc[8][4];
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
#pragma unroll 4
    for (int q = 0; q < NB; q++) {
        double av[8], bv[4];
        const double2* ap = reinterpret_cast<const double2*>(Pa + q * SP + tr);
        const double2* bp = reinterpret_cast<const double2*>(Pb + q * SP + tc);
#pragma unroll
        for (int u = 0; u < 4; u++) { const double2 t2 = ap[u]; av[2 * u] = t2.x; av[2 * u + 1] = t2.y; }
#pragma unroll
        for (int u = 0; u < 2; u++) { const double2 t2 = bp[u]; bv[2 * u] = t2.x; bv[2 * u + 1] = t2.y; }
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u][v] += av[u] * bv[v];
    }
    double* out = tile_ptr(T, off, F, I, J);
#pragma unroll
    for (int u = 0; u < 8; u++) {
        double* o = out + (tr + u) * NB + tc;
        if (PANEL) {
            *reinterpret_cast<double2*>(o) = make_double2(acc[u][0], acc[u][1]);
            *reinterpret_cast<double2*>(o + 2) = make_double2(acc[u][2], acc[u][3]);
        } else if (I != J) {
            double2 o0 = *reinterpret_cast<double2*>(o), o1 = *reinterpret_cast<double2*>(o + 2);
            o0.x -= acc[u][0]; o0.y -= acc[u][1]; o1.x -= acc[u][2]; o1.y -= acc[u][3];
            *reinterpret_cast<double2*>(o) = o0; *reinterpret_cast<double2*>(o + 2) = o1;
        } else {
#pragma unroll
            for (int v = 0; v < 4; v++) if (tc + v <= tr + u) o[v] -= acc[u][v];
        }
    }
}

// The same two products on the FP64 TENSOR path: mma.sync m8n8k4 (DMMA).  One CTA of 128 threads per output tile, each warp a
// 32 x 32 quadrant = 4 x 4 fragments of 8 x 8; both operands stay row-major in shared memory (pitch 68 doubles: the fragment
// loads of a warp -- 8 rows x 4 consecutive doubles -- then need the minimum two wavefronts), because C = A B^T takes its A
// fragment as A[r0 + lane / 4][q0 + lane % 4] and its B fragment (col-major 4 x 8) as B[c0 + lane / 4][q0 + lane % 4]: the same
// access.  Per 4-wide k step a warp issues 8 fragment loads and 16 DMMA (256 FMA each) instead of 64 x 4 DFMA per lane: 8 x fewer
// issue slots for the same flops, which is what the latency-bound DFMA version (FP64 pipe 22 % active) was short of.
constexpr int MP = NB + 4;
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// MODE 0 (panel):    T(I, k) <- T(I, k) W_k^T                                   I = act[blockIdx.x]
// MODE 1 (narrow):   T(I, J) -= T(I, k) T(J, k)^T                               I = act[blockIdx.x] >= J = act[blockIdx.y]  (J inside the column block)
// MODE 2 (trailing): T(I, J) -= sum_{kk = k .. k_end - 1} T(I, kk) T(J, kk)^T    pairs J <= I of act, one pass over the output tile for the
//                    whole column block (a tile row whose envelope starts after kk has no tile there and contributes nothing)
template<int MODE>
__global__ void __launch_bounds__(128) k_tile_product_mma(double* __restrict__ T, const unsigned long long* __restrict__ off, const int32_t* __restrict__ F,
                                                          const double* __restrict__ Winv, const int32_t* __restrict__ act, int m, int k, int k_end,
                                                          const int* __restrict__ flags)
{
    if (flags[0]) return;
    extern __shared__ __align__(16) double sm[];
    double* Pa = sm;                 // [NB][MP] row-major
    double* Pb = sm + NB * MP;
    int I, J;
    if (MODE == 0) { I = act[blockIdx.x]; J = k; }
    else if (MODE == 1) {
        if (blockIdx.x < blockIdx.y) return;
        I = act[blockIdx.x]; J = act[blockIdx.y];
    } else {
        const int idx = blockIdx.x;
        int a = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= idx) a++;
        while (a * (a + 1) / 2 > idx) a--;
        const int b = idx - a * (a + 1) / 2;
        if (a >= m) return;
        I = act[a]; J = act[b];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;     // this warp's quadrant
    const int fr = lane >> 2, fq = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const int FI = F[I], FJ = (MODE == 0) ? 0 : F[J];
    bool first = true;
    for (int kk = k; kk < ((MODE == 2) ? k_end : k + 1); kk++) {
        if (MODE == 2 && (FI > kk || FJ > kk)) continue;       // (uniform in the CTA)
        const double* pa = T + (off[I] + (unsigned long long)(kk - FI)) * TILE;
        const double* pb = (MODE == 0) ? Winv + (size_t)kk * TILE : T + (off[J] + (unsigned long long)(kk - FJ)) * TILE;
        if (!first) __syncthreads();                           // the previous column's fragments have been read
        first = false;
        {
            // 2 x 8 x 16-byte loads in flight per thread, then the stores (rows keep their order: no transposition)
            const double2* ga = reinterpret_cast<const double2*>(pa);
            const double2* gb = reinterpret_cast<const double2*>(pb);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double2 va[8], vb[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { const int t = threadIdx.x + 128 * (8 * h + u); va[u] = ga[t]; vb[u] = gb[t]; }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int t = threadIdx.x + 128 * (8 * h + u);      // pair index: row t / 32, columns 2 (t % 32), +1
                    const int r = t / (NB / 2), q = 2 * (t % (NB / 2));
                    *reinterpret_cast<double2*>(Pa + r * MP + q) = va[u];
                    *reinterpret_cast<double2*>(Pb + r * MP + q) = vb[u];
                }
            }
        }
        __syncthreads();
        const double* arow = Pa + (wr + fr) * MP + fq;
        const double* brow = Pb + (wc + fr) * MP + fq;
#pragma unroll 4
        for (int q0 = 0; q0 < NB; q0 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { af[i] = arow[i * 8 * MP + q0]; bf[i] = brow[i * 8 * MP + q0]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    if (first) return;                                         // no column of the block reaches this pair
    // fragment (i, j): rows wr + 8 i + lane / 4, columns wc + 8 j + 2 (lane % 4), + 1
    double* out = T + (off[I] + (unsigned long long)(J - FI)) * TILE;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = wr + 8 * i + fr, c = wc + 8 * j + 2 * fq;
            double* o = out + r * NB + c;
            if (MODE == 0) *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (I != J) {
                double2 v = *reinterpret_cast<double2*>(o);
                v.x -= acc[i][j][0]; v.y -= acc[i][j][1];
                *reinterpret_cast<double2*>(o) = v;
            } else {
                if (c <= r) o[0] -= acc[i][j][0];
                if (c + 1 <= r) o[1] -= acc[i][j][1];
            }
        }
}

// ---- triangular solves with one right-hand side, one launch per tile column ----
// forward  (k ascending):  z_k = W_k y_k ;  y_I -= T(I, k) z_k for the active rows I of column k
// backward (k descending): x_k = W_k^T z_k ; z_J -= T(k, J)^T x_k for J = F(k) .. k - 1
// Every CTA (256 threads) recomputes the 64-vector of its column from the inverse diagonal factor (a mat-vec, no substitution
// chain), CTA 0 stores it, and each CTA applies one tile.  `in` is only read at tile row k and updated at other rows; `out` is a
// different array.
template<bool BACKWARD>
__global__ void __launch_bounds__(256) k_solve_step(const double* __restrict__ T, const unsigned long long* __restrict__ off, const int32_t* __restrict__ F,
                                                    const double* __restrict__ Winv, const int32_t* __restrict__ act, double* __restrict__ in,
                                                    double* __restrict__ out, int k)
{
    __shared__ double Ws[NB][NB + 1];
    __shared__ double vin[NB], v[NB];
    __shared__ double part[4][NB];
    const double* W = Winv + (size_t)k * TILE;
    {
        double tmp[16];
#pragma unroll
        for (int u = 0; u < 16; u++) tmp[u] = W[threadIdx.x + 256 * u];
#pragma unroll
        for (int u = 0; u < 16; u++) { const int t = threadIdx.x + 256 * u; Ws[t / NB][t % NB] = tmp[u]; }
    }
    if (threadIdx.x < NB) vin[threadIdx.x] = in[k * NB + threadIdx.x];
    // this CTA's tile, prefetched into registers while the mat-vec runs
    const bool has_tile = BACKWARD ? (k - F[k] > 0) : (act != nullptr);
    const double* tile = nullptr;
    int target = 0;
    if (has_tile) {
        if (BACKWARD) { target = F[k] + blockIdx.x; tile = T + (off[k] + (unsigned long long)blockIdx.x) * TILE; }
        else { target = act[blockIdx.x]; tile = T + (off[target] + (unsigned long long)(k - F[target])) * TILE; }
    }
    double tv[16];
    if (has_tile) {
#pragma unroll
        for (int u = 0; u < 16; u++) tv[u] = tile[threadIdx.x + 256 * u];   // element (row 4 u + tid / 64, col tid % 64)
    }
    __syncthreads();
    {
        // v = W vin (forward) or W^T vin (backward): four threads per output entry
        const int e = threadIdx.x / 4, q = threadIdx.x % 4;
        double sum = 0.0;
#pragma unroll
        for (int u = 0; u < 16; u++) { const int j = 4 * u + q; sum += (BACKWARD ? Ws[j][e] : Ws[e][j]) * vin[j]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (q == 0) { v[e] = sum; if (blockIdx.x == 0) out[k * NB + e] = sum; }
    }
    __syncthreads();
    if (!has_tile) return;
    const int col = threadIdx.x % NB, rq = threadIdx.x / NB;   // this thread holds tile rows 4 u + rq of column col
    if (BACKWARD) {
        // in_J[col] -= sum_j tile[j][col] v[j]
        double sum = 0.0;
#pragma unroll
        for (int u = 0; u < 16; u++) sum += tv[u] * v[4 * u + rq];
        part[rq][col] = sum;
        __syncthreads();
        if (threadIdx.x < NB) in[target * NB + threadIdx.x] -= part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
    } else {
        // in_I[row] -= sum_c tile[row][c] v[c]: reduce every row over the 64 threads that hold its entries (two warps)
        __shared__ double rowsum[NB][2];
        const double vc = v[col];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            double p = tv[u] * vc;
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            if ((threadIdx.x & 31) == 0) rowsum[4 * u + rq][(threadIdx.x >> 5) & 1] = p;
        }
        __syncthreads();
        if (threadIdx.x < NB) in[target * NB + threadIdx.x] -= rowsum[threadIdx.x][0] + rowsum[threadIdx.x][1];
    }
}

// y (elimination order) = -grad ; du = y back in DoF order
__global__ void k_rhs(const double* __restrict__ grad, const int32_t* __restrict__ perm, double* __restrict__ y, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[3 * perm[i / 3] + i % 3] = -grad[i];
}
__global__ void k_unpermute(double* __restrict__ du, const double* __restrict__ y, const int32_t* __restrict__ perm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) du[i] = y[3 * perm[i / 3] + i % 3];
}
// out[0] = du.grad, out[1] = |du|_inf (single CTA, fixed tree)
__global__ void __launch_bounds__(1024) k_du_stats(const double* __restrict__ du, const double* __restrict__ grad, int n, double* __restrict__ out)
{
    __shared__ double s0[32], s1[32];
    double dg = 0.0, mx = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) { dg += du[i] * grad[i]; mx = fmax(mx, fabs(du[i])); }
    for (int o = 16; o > 0; o >>= 1) { dg += __shfl_down_sync(0xffffffffu, dg, o); mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = dg; s1[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; w++) { a += s0[w]; b = fmax(b, s1[w]); }
        out[0] = a; out[1] = b;
    }
}

// ---------------------------------------------------------------- host driver

static size_t llt_max_bytes()
{
    const char* e = getenv("SB_LLT_MAX_BYTES");
    if (e && *e) return (size_t)strtoull(e, nullptr, 10);
    return (size_t)96 << 30;
}

// envelope of the current matrix under the current ordering -> F_new (device); flags[1] = it leaves the cached envelope
static void envelope_of_current(sb_context* ctx, Direct& D, const unsigned long long* rows, const int32_t* cols, int nbr, int nt, bool compare)
{
    cudaStream_t st = ctx->stream;
    k_iota_tiles<<<(nt + 255) / 256, 256, 0, st>>>(D.F_new.p, nt);
    k_envelope<<<(nbr + 7) / 8, 256, 0, st>>>(rows, cols, D.perm.p, D.F_new.p, nbr);
    if (compare) k_envelope_inside<<<(nt + 255) / 256, 256, 0, st>>>(D.F_new.p, D.F.p, nt, D.d_flags);
    ctx->launches += compare ? 3 : 2;
}

// offsets and active lists from the envelope on the host; uploads them
static int analyse(sb_context* ctx, Direct& D, int nt)
{
    cudaStream_t st = ctx->stream;
    D.h_F.resize(nt);
    SB_CUDA(ctx, cudaMemcpyAsync(D.h_F.data(), D.F_new.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<unsigned long long> off(nt + 1);
    off[0] = 0;
    for (int I = 0; I < nt; I++) off[I + 1] = off[I] + (unsigned long long)(I - D.h_F[I] + 1);
    D.n_tiles = (size_t)off[nt];
    // active rows of column k: the rows I > k with F(I) <= k
    D.h_act_ptr.assign(nt + 1, 0);
    for (int I = 0; I < nt; I++)
        for (int k = D.h_F[I]; k < I; k++) D.h_act_ptr[k + 1]++;
    for (int k = 0; k < nt; k++) D.h_act_ptr[k + 1] += D.h_act_ptr[k];
    std::vector<int32_t> act((size_t)D.h_act_ptr[nt]);
    {
        std::vector<int32_t> fill(D.h_act_ptr.begin(), D.h_act_ptr.end() - 1);
        for (int I = 0; I < nt; I++)   // ascending I: every column's list comes out sorted
            for (int k = D.h_F[I]; k < I; k++) act[(size_t)fill[k]++] = I;
    }
    D.h_act = act;
    D.off.ensure(nt + 1); D.F.ensure(nt); D.act_ptr.ensure(nt + 1); D.act.ensure(std::max<size_t>(act.size(), 1));
    SB_CUDA(ctx, cudaMemcpyAsync(D.off.p, off.data(), sizeof(unsigned long long) * (nt + 1), cudaMemcpyHostToDevice, st));
    SB_CUDA(ctx, cudaMemcpyAsync(D.F.p, D.h_F.data(), sizeof(int32_t) * nt, cudaMemcpyHostToDevice, st));
    if (!act.empty()) SB_CUDA(ctx, cudaMemcpyAsync(D.act.p, act.data(), sizeof(int32_t) * act.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));   // the host vectors go out of scope
    D.n_analyses++;
    return 0;
}

static int reorder(sb_context* ctx, Direct& D, const unsigned long long* rows, const int32_t* cols, int nbr, size_t nnzb)
{
    cudaStream_t st = ctx->stream;
    std::vector<unsigned long long> h_rows(nbr + 1);
    std::vector<int32_t> h_cols(nnzb);
    SB_CUDA(ctx, cudaMemcpyAsync(h_rows.data(), rows, sizeof(unsigned long long) * (nbr + 1), cudaMemcpyDeviceToHost, st));
    if (nnzb) SB_CUDA(ctx, cudaMemcpyAsync(h_cols.data(), cols, sizeof(int32_t) * nnzb, cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    rcm_order(nbr, h_rows, h_cols, D.h_perm);
    D.perm.ensure(nbr);
    SB_CUDA(ctx, cudaMemcpyAsync(D.perm.p, D.h_perm.data(), sizeof(int32_t) * nbr, cudaMemcpyHostToDevice, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    D.n_orderings++;
    return 0;
}

int solve_llt_internal(sb_context* ctx, int* out_ok, double* out_du_dot_grad, double* out_du_inf)
{
    int nbr; size_t nnzb; const unsigned long long* rows; const int32_t* cols; const float* vals;
    int r = bcsr_view(ctx, &nbr, &nnzb, &rows, &cols, &vals);
    if (r) return r;
    const int n = 3 * nbr;
    if (n != ctx->ndofs) return fail(ctx, SB_ERR_STATE, "sb_solve_llt: matrix and DoF vector sizes differ");
    StageTimer timer(ctx, ST_PCG);
    if (!ctx->direct) {
        ctx->direct = new Direct();
        cudaMalloc(&ctx->direct->d_flags, 2 * sizeof(int));
        cudaMallocHost(&ctx->direct->h_flags, 2 * sizeof(int));
        cudaMallocHost(&ctx->direct->h_out, 2 * sizeof(double));
        cudaFuncSetAttribute(k_tile_product<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NB * SP * sizeof(double)));
        cudaFuncSetAttribute(k_tile_product<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NB * SP * sizeof(double)));
        cudaFuncSetAttribute(k_tile_product_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NB * MP * sizeof(double)));
        cudaFuncSetAttribute(k_tile_product_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NB * MP * sizeof(double)));
        cudaFuncSetAttribute(k_tile_product_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NB * MP * sizeof(double)));
    }
    Direct& D = *ctx->direct;
    cudaStream_t st = ctx->stream;
    const int nt = (n + NB - 1) / NB, np = nt * NB;
    D.F_new.ensure(nt);
    SB_CUDA(ctx, cudaMemsetAsync(D.d_flags, 0, 2 * sizeof(int), st));

    // ---- analysis: reuse the cached ordering + envelope when the matrix fits inside
    bool need_analysis = !D.have_order || D.nbr != nbr || D.nt != nt;
    if (need_analysis) {
        if ((r = reorder(ctx, D, rows, cols, nbr, nnzb))) return r;
        envelope_of_current(ctx, D, rows, cols, nbr, nt, false);
        if ((r = analyse(ctx, D, nt))) return r;
        D.n_tiles_at_order = D.n_tiles;
        D.have_order = true; D.nbr = nbr; D.nt = nt;
    } else {
        envelope_of_current(ctx, D, rows, cols, nbr, nt, true);
        SB_CUDA(ctx, cudaMemcpyAsync(D.h_flags, D.d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SB_CUDA(ctx, cudaStreamSynchronize(st));
        if (D.h_flags[1]) {
            // outside the cached envelope: new envelope under the old ordering; when that costs much more storage than the
            // ordering was made for (contacts between far-apart nodes), order again
            if ((r = analyse(ctx, D, nt))) return r;
            if (D.n_tiles > D.n_tiles_at_order + D.n_tiles_at_order / 4 + 8) {
                if ((r = reorder(ctx, D, rows, cols, nbr, nnzb))) return r;
                envelope_of_current(ctx, D, rows, cols, nbr, nt, false);
                if ((r = analyse(ctx, D, nt))) return r;
                D.n_tiles_at_order = D.n_tiles;
            }
            SB_CUDA(ctx, cudaMemsetAsync(D.d_flags, 0, 2 * sizeof(int), st));
        }
    }
    if (D.n_tiles * TILE * sizeof(double) > llt_max_bytes())
        return fail(ctx, SB_ERR_STATE, "sb_solve_llt: the factor's tile envelope needs " + std::to_string((D.n_tiles * TILE * sizeof(double)) >> 20) +
                                           " MiB (limit SB_LLT_MAX_BYTES); use the block-Jacobi PCG for this system");
    D.T.ensure(D.n_tiles * TILE);
    if (!D.T.p) return fail(ctx, SB_ERR_CUDA, "sb_solve_llt: out of device memory for the factor");
    D.W.ensure((size_t)nt * TILE);
    D.y.ensure(np); D.z.ensure(np);
    ctx->du.ensure(n);

    // ---- numeric factorisation
    static const bool dump = getenv("SB_LLT_DUMP") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    if (dump) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventRecord(e0, st); }
    SB_CUDA(ctx, cudaMemsetAsync(D.T.p, 0, sizeof(double) * D.n_tiles * TILE, st));
    SB_CUDA(ctx, cudaMemsetAsync(D.y.p, 0, sizeof(double) * np, st));
    k_fill_tiles<<<nbr, 288, 0, st>>>(rows, cols, vals, D.perm.p, D.T.p, D.off.p, D.F.p, nbr);
    if (np > n) k_pad_diagonal<<<(np - n + 63) / 64, 64, 0, st>>>(D.T.p, D.off.p, D.F.p, n, np);
    ctx->launches += 2;
    const size_t syrk_smem = 2 * NB * SP * sizeof(double), mma_smem = 2 * NB * MP * sizeof(double);
    static const bool use_mma = getenv("SB_LLT_NO_MMA") == nullptr;   // (A/B hook: the DFMA version of the tile products)
    if (use_mma) {
        // blocked right-looking: inside a block of `pw` tile columns every column updates only the block's own columns (narrow
        // update), and the tiles to the right of the block are read and written ONCE per block for all its columns (trailing
        // update) -- at the million-tet bar the active window (180 MB of tiles) does not fit the L2, and one pass per column made
        // the factorisation HBM-bound
        static const int pw = std::max(1, getenv("SB_LLT_PANEL") ? atoi(getenv("SB_LLT_PANEL")) : 4);
        for (int k0 = 0; k0 < nt; k0 += pw) {
            const int k1 = std::min(k0 + pw, nt);
            for (int j = k0; j < k1; j++) {
                k_potrf_inv<<<1, 256, 0, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, j, D.d_flags);
                ctx->launches++;
                const int m = D.h_act_ptr[j + 1] - D.h_act_ptr[j];
                if (m <= 0) continue;
                const int32_t* act = D.act.p + D.h_act_ptr[j];
                k_tile_product_mma<0><<<m, 128, mma_smem, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, act, m, j, j + 1, D.d_flags);
                ctx->launches++;
                int nb = 0;   // active rows of column j that are columns of this block (the list is ascending)
                while (nb < m && D.h_act[(size_t)D.h_act_ptr[j] + nb] < k1) nb++;
                if (nb > 0) {
                    k_tile_product_mma<1><<<dim3(m, nb), 128, mma_smem, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, act, m, j, j + 1, D.d_flags);
                    ctx->launches++;
                }
            }
            const int mt = D.h_act_ptr[k1] - D.h_act_ptr[k1 - 1];
            if (mt > 0) {
                k_tile_product_mma<2><<<(unsigned)((size_t)mt * (mt + 1) / 2), 128, mma_smem, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, D.act.p + D.h_act_ptr[k1 - 1], mt, k0, k1, D.d_flags);
                ctx->launches++;
            }
        }
    } else
    for (int k = 0; k < nt; k++) {
        k_potrf_inv<<<1, 256, 0, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, k, D.d_flags);
        ctx->launches++;
        const int m = D.h_act_ptr[k + 1] - D.h_act_ptr[k];
        if (m > 0) {
            const int32_t* act = D.act.p + D.h_act_ptr[k];
            k_tile_product<true><<<m, 128, syrk_smem, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, act, m, k, D.d_flags);
            k_tile_product<false><<<(unsigned)((size_t)m * (m + 1) / 2), 128, syrk_smem, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, act, m, k, D.d_flags);
            ctx->launches += 2;
        }
    }
    SB_CUDA(ctx, cudaMemcpyAsync(D.h_flags, D.d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (dump) cudaEventRecord(e1, st);
    // ---- triangular solves (queued behind the factorisation; their results are ignored when it failed)
    k_rhs<<<(n + 255) / 256, 256, 0, st>>>(ctx->grad.p, D.perm.p, D.y.p, n);
    for (int k = 0; k < nt; k++) {
        const int m = D.h_act_ptr[k + 1] - D.h_act_ptr[k];
        k_solve_step<false><<<std::max(m, 1), 256, 0, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, m > 0 ? D.act.p + D.h_act_ptr[k] : nullptr, D.y.p, D.z.p, k);
    }
    for (int k = nt - 1; k >= 0; k--) {
        const int m = k - D.h_F[k];
        k_solve_step<true><<<std::max(m, 1), 256, 0, st>>>(D.T.p, D.off.p, D.F.p, D.W.p, nullptr, D.z.p, D.y.p, k);
    }
    k_unpermute<<<(n + 255) / 256, 256, 0, st>>>(ctx->du.p, D.y.p, D.perm.p, n);
    k_du_stats<<<1, 1024, 0, st>>>(ctx->du.p, ctx->grad.p, n, ctx->d_scalars + 4);
    ctx->launches += 2 * nt + 3;
    SB_CUDA(ctx, cudaMemcpyAsync(D.h_out, ctx->d_scalars + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dump) cudaEventRecord(e2, st);
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    SB_CUDA(ctx, cudaGetLastError());
    if (dump) {
        float f = 0, s2 = 0;
        cudaEventElapsedTime(&f, e0, e1); cudaEventElapsedTime(&s2, e1, e2);
        size_t pairs = 0; int mmax = 0;
        for (int k = 0; k < nt; k++) { const size_t m = D.h_act_ptr[k + 1] - D.h_act_ptr[k]; pairs += m * (m + 1) / 2; mmax = std::max(mmax, (int)m); }
        fprintf(stderr, "LLT n %d tile rows %d tiles %zu (%.1f MiB) syrk tiles %zu (%.2f GFLOP) max active %d | factor %.3f ms solve %.3f ms ok %d | orderings %lld analyses %lld\n",
                n, nt, D.n_tiles, D.n_tiles * TILE * 8.0 / 1048576.0, pairs, pairs * 2.0 * NB * NB * NB * 1e-9, mmax, f, s2, !D.h_flags[0],
                (long long)D.n_orderings, (long long)D.n_analyses);
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    }
    if (D.h_flags[0]) {   // not positive definite: Eigen::SimplicialLLT::info() != Success
        if (out_ok) *out_ok = 0;
        if (out_du_dot_grad) *out_du_dot_grad = 0.0;
        if (out_du_inf) *out_du_inf = 0.0;
        return 0;
    }
    if (out_ok) *out_ok = 1;
    if (out_du_dot_grad) *out_du_dot_grad = D.h_out[0];
    if (out_du_inf) *out_du_inf = D.h_out[1];
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_solve_llt(sb_context* ctx, int* out_ok, double* out_du_dot_grad, double* out_du_inf)
{
    if (!ctx) return SB_ERR_ARG;
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_solve_llt: no gradient: call sb_eval(SB_EVAL_PGH) first");
    return solve_llt_internal(ctx, out_ok, out_du_dot_grad, out_du_inf);
}

// n_tiles, n_tile_rows, factor bytes, orderings made, analyses made (diagnostic; tests and bench.py)
extern "C" int sb_llt_stats(sb_context* ctx, double* out5)
{
    if (!ctx || !out5) return SB_ERR_ARG;
    for (int i = 0; i < 5; i++) out5[i] = 0.0;
    if (!ctx->direct) return 0;
    const Direct& D = *ctx->direct;
    out5[0] = (double)D.n_tiles; out5[1] = (double)D.nt; out5[2] = (double)(D.n_tiles * TILE * sizeof(double));
    out5[3] = (double)D.n_orderings; out5[4] = (double)D.n_analyses;
    return 0;
}

// the ordering alone, on host arrays (no GPU involved): block rows / first scalar columns as sb_bcsr_get returns them
extern "C" int sb_llt_order(int nbr, const unsigned long long* rows, const int32_t* cols, int32_t* out_perm)
{
    if (nbr < 0 || !rows || !out_perm || (rows[nbr] && !cols)) return SB_ERR_ARG;
    std::vector<unsigned long long> r(rows, rows + nbr + 1);
    std::vector<int32_t> c(cols, cols + rows[nbr]);
    std::vector<int32_t> perm;
    rcm_order(nbr, r, c, perm);
    std::copy(perm.begin(), perm.end(), out_perm);
    return 0;
}
