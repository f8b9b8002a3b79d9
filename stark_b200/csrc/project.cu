// Projection of element Hessians to positive definiteness.
//
// Replaces project_to_PD_inplace (symx/solver/second_order/project_to_PD.cpp:13-82, Eigen::SelfAdjointEigenSolver on
// n in {3,6,9,12,15} or dynamic) and the selection logic of ElementHessians::{project_to_PD_inplace__all,
// _project_to_PD_for_update} (ElementHessians.cpp:48-182) together with the PPN block selection of
// NewtonsMethod::_project_and_assemble (symx/solver/NewtonsMethod.cpp:316-327).
//
// One warp per selected element: cyclic Jacobi eigen-decomposition in shared memory, eigenvalues below eps clamped to
// eps (or mirrored), H = V diag(l) V^T rebuilt only if something changed -- the projected matrix replaces the original
// in the element-Hessian store and the next numeric assembly re-sums every BCSR block (instead of the reference's
// "add (projected - original)" update pass; same matrix up to float rounding).
#include "internal.h"
#include <algorithm>

namespace sb {

constexpr int PROJ_MAX_N = 24;
constexpr int PROJ_MAX_POTS = 128;

struct ProjTable {
    int n_pots;
    unsigned long long E_off[PROJ_MAX_POTS + 1];   // first global element id of each potential (+ total)
    unsigned long long H_off[PROJ_MAX_POTS];
    unsigned long long rows_off[PROJ_MAX_POTS];
    unsigned long long blk_off[PROJ_MAX_POTS];     // first source (element block) of each potential in assembly numbering
    int nb[PROJ_MAX_POTS];
};

__device__ __forceinline__ int find_pot(const ProjTable& T, unsigned long long e)
{
    int lo = 0, hi = T.n_pots - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (T.E_off[mid] <= e) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// active[b] = |g_b|_inf >= threshold ; counts inactive blocks (all_projected <=> none inactive)
__global__ void k_active_blocks(const double* __restrict__ grad, uint8_t* __restrict__ active, int nbr, double threshold, int* __restrict__ n_inactive)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbr) return;
    const double m = fmax(fabs(grad[3 * b]), fmax(fabs(grad[3 * b + 1]), fabs(grad[3 * b + 2])));
    const uint8_t a = (m >= threshold) ? 1 : 0;
    active[b] = a;
    if (!a) atomicAdd(n_inactive, 1);
}

// select not-yet-projected elements (all of them, or those touching an active block) into a compact list
__global__ void k_select(const ProjTable* __restrict__ Tp, const int32_t* __restrict__ rows_all, const uint8_t* __restrict__ active, int use_active,
                         uint8_t* __restrict__ projected, uint32_t* __restrict__ list, int* __restrict__ n_list, unsigned long long n_elem_total)
{
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elem_total) return;
    if (projected[e]) return;
    bool sel = true;
    if (use_active) {
        const ProjTable& T = *Tp;
        const int pi = find_pot(T, e);
        const int nb = T.nb[pi];
        const int32_t* r = rows_all + T.rows_off[pi] + (e - T.E_off[pi]) * nb;
        sel = false;
        for (int b = 0; b < nb; b++) sel = sel || active[r[b]];
    }
    if (sel) {
        projected[e] = 1;
        list[atomicAdd(n_list, 1)] = (uint32_t)e;
    }
}

// One 128-thread CTA per selected element.
//  1. cheap exit: lambda_min(A) > eps  <=>  A - eps I has an LDL^T factorisation with positive pivots (n steps);
//  2. otherwise a PARALLEL-ORDER two-sided Jacobi eigen-solve: the n/2 disjoint rotations of one round-robin step are
//     computed from the same matrix and applied together (columns of A and V, then rows of A), n-1 steps per sweep --
//     the dependent chain of a sweep is n-1 steps instead of the n(n-1)/2 rotations of the cyclic order;
//  3. clamp / mirror the eigenvalues below eps and rebuild H = V diag(l) V^T.
constexpr int PROJ_THREADS = 128;
__global__ void __launch_bounds__(PROJ_THREADS) k_project(const ProjTable* __restrict__ Tp, double* __restrict__ H_all, const uint32_t* __restrict__ list,
                                                           const int* __restrict__ n_list, double eps, int mirror, int* __restrict__ n_changed,
                                                           const DirtyView dv)
{
    __shared__ double A[PROJ_MAX_N * PROJ_MAX_N];
    __shared__ double V[PROJ_MAX_N * PROJ_MAX_N];
    __shared__ double s_c[PROJ_MAX_N / 2], s_s[PROJ_MAX_N / 2];
    __shared__ int s_p[PROJ_MAX_N / 2], s_q[PROJ_MAX_N / 2];
    __shared__ double s_red[2][PROJ_THREADS / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ProjTable& T = *Tp;
    const int total = *n_list;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const unsigned long long e = list[item];
        const int pi = find_pot(T, e);
        const int n = 3 * T.nb[pi];
        double* H = H_all + T.H_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(n * n);
        __syncthreads();   // previous item fully done
        // load (symmetrised); V = A - eps I for the factorisation test
        for (int k = tid; k < n * n; k += PROJ_THREADS) {
            const int i = k / n, j = k - i * n;
            const double a = 0.5 * (H[i * n + j] + H[j * n + i]);
            A[k] = a;
            V[k] = a - ((i == j) ? eps : 0.0);
        }
        __syncthreads();
        bool pd = true;
        for (int j = 0; j < n; j++) {
            const double d = V[j * n + j];
            if (!(d > 0.0)) { pd = false; break; }   // shared value: uniform across the CTA
            const double inv = 1.0 / d;
            const int m = n - 1 - j;                  // trailing block: rows / cols j+1 .. n-1 (lower triangle incl. diagonal)
            for (int t = tid; t < m * m; t += PROJ_THREADS) {
                const int i = j + 1 + t / m, k2 = j + 1 + t % m;
                if (k2 <= i) V[i * n + k2] -= V[i * n + j] * V[k2 * n + j] * inv;
            }
            __syncthreads();
        }
        if (pd) continue;
        __syncthreads();
        for (int k = tid; k < n * n; k += PROJ_THREADS) {
            const int i = k / n, j = k - i * n;
            V[k] = (i == j) ? 1.0 : 0.0;
        }
        const int ne = (n + 1) & ~1;       // even number of players (a dummy index n when n is odd)
        const int np = ne / 2;             // pairs per step
        __syncthreads();
        for (int sweep = 0; sweep < 30; sweep++) {
            // convergence: off-diagonal mass vs total
            double off = 0.0, diag = 0.0;
            for (int k = tid; k < n * n; k += PROJ_THREADS) {
                const int i = k / n, j = k - i * n;
                const double v = A[k] * A[k];
                if (i == j) diag += v; else off += v;
            }
            for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); diag += __shfl_xor_sync(0xffffffffu, diag, o); }
            if (lane == 0) { s_red[0][warp] = off; s_red[1][warp] = diag; }
            __syncthreads();
            off = 0.0; diag = 0.0;
            for (int w = 0; w < PROJ_THREADS / 32; w++) { off += s_red[0][w]; diag += s_red[1][w]; }
            __syncthreads();
            if (off <= 1e-25 * (diag + off) || off == 0.0) break;   // |off| / |A| <= 3e-13: eigenvalue error ~ |off|^2 / gap, far below 1e-10 parity
            for (int step = 0; step < ne - 1; step++) {
                // round-robin pairing: player ne-1 is fixed, the others rotate
                if (tid < np) {
                    int p, q;
                    if (tid == 0) { p = ne - 1; q = step; }
                    else { p = (step + tid) % (ne - 1); q = (step - tid + (ne - 1)) % (ne - 1); }
                    if (p > q) { const int t = p; p = q; q = t; }
                    double c = 1.0, sn = 0.0;
                    if (q < n) {   // (a pair with the dummy index does nothing)
                        const double apq = A[p * n + q];
                        if (apq != 0.0) {
                            const double app = A[p * n + p], aqq = A[q * n + q];
                            const double tau = (aqq - app) / (2.0 * apq);
                            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                            c = 1.0 / sqrt(1.0 + t * t);
                            sn = t * c;
                        }
                    }
                    s_p[tid] = p; s_q[tid] = (q < n) ? q : -1; s_c[tid] = c; s_s[tid] = sn;
                }
                __syncthreads();
                // columns p, q of A and V (all rows)
                for (int t = tid; t < np * n; t += PROJ_THREADS) {
                    const int pr = t / n, k = t - pr * n;
                    const int p = s_p[pr], q = s_q[pr];
                    if (q < 0) continue;
                    const double c = s_c[pr], sn = s_s[pr];
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - sn * akq;
                    A[k * n + q] = sn * akp + c * akq;
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - sn * vkq;
                    V[k * n + q] = sn * vkp + c * vkq;
                }
                __syncthreads();
                // rows p, q of A (all columns)
                for (int t = tid; t < np * n; t += PROJ_THREADS) {
                    const int pr = t / n, k = t - pr * n;
                    const int p = s_p[pr], q = s_q[pr];
                    if (q < 0) continue;
                    const double c = s_c[pr], sn = s_s[pr];
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - sn * aqk;
                    A[q * n + k] = sn * apk + c * aqk;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        // clamp / mirror
        if (tid == 0) {
            int changed = 0;
            for (int i = 0; i < n; i++) if (A[i * n + i] < eps) changed = 1;
            s_flag = changed;
        }
        __syncthreads();
        if (s_flag) {
            if (tid < n) {
                const double l = A[tid * n + tid];
                A[tid * n + tid] = (l < eps) ? (mirror ? -l : eps) : l;
            }
            __syncthreads();
            for (int k = tid; k < n * n; k += PROJ_THREADS) {
                const int i = k / n, j = k - i * n;
                double acc = 0.0;
                for (int m = 0; m < n; m++) acc += V[i * n + m] * A[m * n + m] * V[j * n + m];
                H[k] = acc;
            }
            if (tid == 0) atomicAdd(n_changed, 1);
            if (dv.dirty) {   // the BCSR blocks this element contributes to must be re-summed
                const int nb = T.nb[pi];
                const unsigned long long src0 = T.blk_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(nb * nb);
                for (int k = tid; k < nb * nb; k += PROJ_THREADS) {
                    const unsigned long long src = src0 + k;
                    const uint32_t f = (src < dv.n_static) ? dv.s_final[dv.s_blk_of_src[src]] : dv.d_final[dv.d_blk_of_src[src - dv.n_static]];
                    dv.dirty[f] = 1;
                }
            }
        }
    }
}

struct Projector {
    DevBuf<uint8_t> active;
    DevBuf<uint32_t> list;
    ProjTable* d_table = nullptr;
    int* d_counts = nullptr;   // [0] n_list, [1] n_changed, [2] n_inactive
    int* h_counts = nullptr;
};
void projector_destroy(sb_context* ctx)
{
    Projector* P = ctx->projector;
    if (!P) return;
    P->active.release(); P->list.release();
    if (P->d_table) cudaFree(P->d_table);
    if (P->d_counts) cudaFree(P->d_counts);
    if (P->h_counts) cudaFreeHost(P->h_counts);
    delete P;
    ctx->projector = nullptr;
}

int project_internal(sb_context* ctx, double grad_threshold, double eps, int mirror, int64_t* out_n_projected, int64_t* out_n_hessians, int* out_all_projected)
{
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_project_to_pd: call sb_eval(SB_EVAL_PGH) first");
    if (!ctx->projector) ctx->projector = new Projector();
    Projector& P = *ctx->projector;
    cudaStream_t st = ctx->stream;
    if (!P.d_table) {
        SB_CUDA(ctx, cudaMalloc(&P.d_table, sizeof(ProjTable)));
        SB_CUDA(ctx, cudaMalloc(&P.d_counts, 4 * sizeof(int)));
        SB_CUDA(ctx, cudaMallocHost(&P.h_counts, 4 * sizeof(int)));
    }
    const size_t n_elem = ctx->n_hessians;
    if (out_n_hessians) *out_n_hessians = (int64_t)n_elem;
    if (grad_threshold < 0.0 || n_elem == 0) {
        if (out_n_projected) *out_n_projected = ctx->n_projected;
        if (out_all_projected) *out_all_projected = 0;
        return 0;
    }
    StageTimer timer(ctx, ST_PROJECT);
    ProjTable T;
    T.n_pots = 0;
    unsigned long long blk_off = 0;
    for (int pidx : layout_order(ctx)) {
        Potential& p = ctx->potentials[pidx];
        if (p.n_elem == 0) continue;
        if (T.n_pots >= PROJ_MAX_POTS) return fail(ctx, SB_ERR_STATE, "sb_project_to_pd: too many active potentials");
        if (p.k->n_dof > PROJ_MAX_N) return fail(ctx, SB_ERR_STATE, "sb_project_to_pd: element size above 24 DoFs is not supported");
        T.E_off[T.n_pots] = p.E_off; T.H_off[T.n_pots] = p.H_off; T.rows_off[T.n_pots] = p.rows_off; T.nb[T.n_pots] = p.k->nb;
        T.blk_off[T.n_pots] = blk_off;
        blk_off += (unsigned long long)p.n_elem * p.k->nb * p.k->nb;
        T.n_pots++;
    }
    T.E_off[T.n_pots] = n_elem;
    SB_CUDA(ctx, cudaMemcpyAsync(P.d_table, &T, sizeof(ProjTable), cudaMemcpyHostToDevice, st));
    SB_CUDA(ctx, cudaMemsetAsync(P.d_counts, 0, 4 * sizeof(int), st));
    const int nbr = ctx->ndofs / 3;
    P.active.ensure(nbr + 1);
    P.list.ensure(n_elem + 1);
    const int use_active = (grad_threshold > 0.0) ? 1 : 0;
    if (use_active) {
        k_active_blocks<<<(nbr + 255) / 256, 256, 0, st>>>(ctx->grad.p, P.active.p, nbr, grad_threshold, P.d_counts + 2);
        ctx->launches++;
    }
    k_select<<<(unsigned)((n_elem + 255) / 256), 256, 0, st>>>(P.d_table, ctx->rows.p, P.active.p, use_active, ctx->projected.p, P.list.p, P.d_counts, n_elem);
    const int grid = (int)std::min<size_t>(n_elem, 148 * 8);
    DirtyView dv;
    if (!assembly_dirty_view(ctx, &dv)) dv.dirty = nullptr;
    k_project<<<grid, PROJ_THREADS, 0, st>>>(P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv);
    ctx->launches += 2;
    SB_CUDA(ctx, cudaMemcpyAsync(P.h_counts, P.d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));   // also protects the stack-resident table T
    SB_CUDA(ctx, cudaGetLastError());
    ctx->n_projected += P.h_counts[0];
    if (ctx->profile) { ctx->stage_calls[ST_PROJ_SELECTED] += P.h_counts[0]; ctx->stage_calls[ST_PROJ_CHANGED] += P.h_counts[1]; }
    if (out_n_projected) *out_n_projected = ctx->n_projected;
    if (out_all_projected) *out_all_projected = use_active ? (P.h_counts[2] == 0) : 1;
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_project_to_pd(sb_context* ctx, double grad_threshold, double eps, int mirror, int64_t* out_n_projected, int64_t* out_n_hessians, int* out_all_projected)
{
    if (!ctx) return SB_ERR_ARG;
    return project_internal(ctx, grad_threshold, eps, mirror, out_n_projected, out_n_hessians, out_all_projected);
}
