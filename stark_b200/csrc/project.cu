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

// One WARP per selected element (everything between the lanes of a warp is __syncwarp / shuffles), eight elements per CTA,
// one kernel instance per element size N (compile-time loop bounds and index arithmetic):
//  1. cheap exit: lambda_min(A) > eps  <=>  A - eps I has an LDL^T factorisation with positive pivots (N steps);
//  2. otherwise a PARALLEL-ORDER two-sided Jacobi eigen-solve: the N/2 disjoint rotations of one round-robin step are
//     computed from the same matrix and applied together (columns of A and V, then rows of A), N-1 steps per sweep --
//     the dependent chain of a sweep is N-1 steps instead of the N(N-1)/2 rotations of the cyclic order;
//  3. clamp / mirror the eigenvalues below eps and rebuild H = V diag(l) V^T.
// Contact-heavy steps project thousands of 12 x 12 tet Hessians per call; the kernel is instruction-bound, so the per-size
// instances (no runtime divisions) and 2.4 KB of shared memory per 12 x 12 element (64 warps resident per SM) matter.
constexpr int PROJ_WARPS = 8;
constexpr int PROJ_THREADS = 32 * PROJ_WARPS;
template<int N> constexpr size_t proj_smem_per_warp() { return sizeof(double) * (2 * N * N + 2 * ((N + 1) / 2)) + sizeof(int) * 2 * ((N + 1) / 2 + 1); }

template<int N>
__global__ void __launch_bounds__(PROJ_THREADS) k_project(const ProjTable* __restrict__ Tp, double* __restrict__ H_all, const uint32_t* __restrict__ list,
                                                           const int* __restrict__ n_list, double eps, int mirror, int* __restrict__ n_changed,
                                                           const DirtyView dv)
{
    constexpr int NE = (N + 1) & ~1;   // even number of players (a dummy index N when N is odd)
    constexpr int NP = NE / 2;         // pairs per step
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smem_raw + (size_t)warp * ((proj_smem_per_warp<N>() + 15) & ~(size_t)15);
    double* A = reinterpret_cast<double*>(base);
    double* V = A + N * N;
    double* cs_c = V + N * N;
    double* cs_s = cs_c + NP;
    int* pp = reinterpret_cast<int*>(cs_s + NP);
    int* pq = pp + NP + 1;
    const ProjTable& T = *Tp;
    const int total = *n_list;
    for (int item = blockIdx.x * PROJ_WARPS + warp; item < total; item += gridDim.x * PROJ_WARPS) {
        const unsigned long long e = list[item];
        const int pi = find_pot(T, e);
        if (3 * T.nb[pi] != N) continue;   // another size class: handled by its own kernel instance
        double* H = H_all + T.H_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(N * N);
        __syncwarp();   // previous item fully done
        // load (symmetrised); V = A - eps I for the factorisation test
        for (int k = lane; k < N * N; k += 32) {
            const int i = k / N, j = k - i * N;
            const double a = 0.5 * (H[i * N + j] + H[j * N + i]);
            A[k] = a;
            V[k] = a - ((i == j) ? eps : 0.0);
        }
        __syncwarp();
        bool pd = true;
        for (int j = 0; j < N; j++) {
            const double d = V[j * N + j];
            if (!(d > 0.0)) { pd = false; break; }   // shared value: uniform across the warp
            const double inv = 1.0 / d;
            const int m = N - 1 - j;                  // trailing block: rows / cols j+1 .. N-1 (lower triangle incl. diagonal)
            for (int t = lane; t < m * m; t += 32) {
                const int i = j + 1 + t / m, k2 = j + 1 + t % m;
                if (k2 <= i) V[i * N + k2] -= V[i * N + j] * V[k2 * N + j] * inv;
            }
            __syncwarp();
        }
        if (pd) continue;
        __syncwarp();
        for (int k = lane; k < N * N; k += 32) {
            const int i = k / N, j = k - i * N;
            V[k] = (i == j) ? 1.0 : 0.0;
        }
        __syncwarp();
        int sweeps_done = 0;
        for (int sweep = 0; sweep < 30; sweep++) {
            sweeps_done = sweep;
            // convergence: off-diagonal mass vs total
            double off = 0.0, diag = 0.0;
            for (int k = lane; k < N * N; k += 32) {
                const int i = k / N, j = k - i * N;
                const double v = A[k] * A[k];
                if (i == j) diag += v; else off += v;
            }
            for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); diag += __shfl_xor_sync(0xffffffffu, diag, o); }
            if (off <= 1e-25 * (diag + off) || off == 0.0) break;   // |off| / |A| <= 3e-13: eigenvalue error ~ |off|^2 / gap, far below 1e-10 parity
            for (int step = 0; step < NE - 1; step++) {
                // round-robin pairing: player NE-1 is fixed, the others rotate
                if (lane < NP) {
                    int p, q;
                    if (lane == 0) { p = NE - 1; q = step; }
                    else { p = (step + lane) % (NE - 1); q = (step - lane + (NE - 1)) % (NE - 1); }
                    if (p > q) { const int t = p; p = q; q = t; }
                    double c = 1.0, sn = 0.0;
                    if (q < N) {   // (a pair with the dummy index does nothing)
                        const double apq = A[p * N + q];
                        if (apq != 0.0) {
                            const double app = A[p * N + p], aqq = A[q * N + q];
                            const double tau = (aqq - app) / (2.0 * apq);
                            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                            c = rsqrt(1.0 + t * t);
                            sn = t * c;
                        }
                    }
                    pp[lane] = p; pq[lane] = (q < N) ? q : -1; cs_c[lane] = c; cs_s[lane] = sn;
                }
                __syncwarp();
                // columns p, q of A and V (all rows)
                for (int t = lane; t < NP * N; t += 32) {
                    const int pr = t / N, k = t - pr * N;
                    const int p = pp[pr], q = pq[pr];
                    if (q < 0) continue;
                    const double c = cs_c[pr], sn = cs_s[pr];
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - sn * akq;
                    A[k * N + q] = sn * akp + c * akq;
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - sn * vkq;
                    V[k * N + q] = sn * vkp + c * vkq;
                }
                __syncwarp();
                // rows p, q of A (all columns)
                for (int t = lane; t < NP * N; t += 32) {
                    const int pr = t / N, k = t - pr * N;
                    const int p = pp[pr], q = pq[pr];
                    if (q < 0) continue;
                    const double c = cs_c[pr], sn = cs_s[pr];
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - sn * aqk;
                    A[q * N + k] = sn * apk + c * aqk;
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0) atomicAdd(n_changed + 2, sweeps_done);   // diagnostic: total Jacobi sweeps (d_counts[3])
        // clamp / mirror
        bool changed = false;
        for (int i = 0; i < N; i++) if (A[i * N + i] < eps) changed = true;   // shared values: uniform across the warp
        if (changed) {
            __syncwarp();
            if (lane < N) {
                const double l = A[lane * N + lane];
                A[lane * N + lane] = (l < eps) ? (mirror ? -l : eps) : l;
            }
            __syncwarp();
            for (int k = lane; k < N * N; k += 32) {
                const int i = k / N, j = k - i * N;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < N; m++) acc += V[i * N + m] * A[m * N + m] * V[j * N + m];
                H[k] = acc;
            }
            if (lane == 0) atomicAdd(n_changed, 1);
            if (dv.dirty) {   // the BCSR blocks this element contributes to must be re-summed
                constexpr int nb = N / 3;
                const unsigned long long src0 = T.blk_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(nb * nb);
                for (int k = lane; k < nb * nb; k += 32) {
                    const unsigned long long src = src0 + k;
                    const uint32_t f = (src < dv.n_static) ? dv.s_final[dv.s_blk_of_src[src]] : dv.d_final[dv.d_blk_of_src[src - dv.n_static]];
                    dv.dirty[f] = 1;
                }
            }
        }
    }
}

template<int N> static cudaError_t launch_project(int grid, cudaStream_t st, const ProjTable* Tp, double* H, const uint32_t* list, const int* n_list,
                                                  double eps, int mirror, int* n_changed, const DirtyView& dv)
{
    const size_t smem = PROJ_WARPS * ((proj_smem_per_warp<N>() + 15) & ~(size_t)15);
    static bool configured = false;
    if (!configured) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_project<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        configured = true;
    }
    k_project<N><<<grid, PROJ_THREADS, smem, st>>>(Tp, H, list, n_list, eps, mirror, n_changed, dv);
    return cudaGetLastError();
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
    const int grid = (int)std::min<size_t>((n_elem + PROJ_WARPS - 1) / PROJ_WARPS, 148 * 8);
    DirtyView dv;
    if (!assembly_dirty_view(ctx, &dv)) dv.dirty = nullptr;
    bool sizes[PROJ_MAX_N / 3 + 1] = {false};
    for (auto& p : ctx->potentials) if (p.n_elem > 0) sizes[p.k->nb] = true;
    for (int nb = 1; nb <= PROJ_MAX_N / 3; nb++) {
        if (!sizes[nb]) continue;
        cudaError_t e = cudaSuccess;
        switch (nb) {
        case 1: e = launch_project<3>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 2: e = launch_project<6>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 3: e = launch_project<9>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 4: e = launch_project<12>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 5: e = launch_project<15>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 6: e = launch_project<18>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 7: e = launch_project<21>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        case 8: e = launch_project<24>(grid, st, P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv); break;
        }
        SB_CUDA(ctx, e);
        ctx->launches++;
    }
    ctx->launches += 1;
    SB_CUDA(ctx, cudaMemcpyAsync(P.h_counts, P.d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));   // also protects the stack-resident table T
    SB_CUDA(ctx, cudaGetLastError());
    ctx->n_projected += P.h_counts[0];
    if (ctx->profile) { ctx->stage_calls[ST_PROJ_SELECTED] += P.h_counts[0]; ctx->stage_calls[ST_PROJ_CHANGED] += P.h_counts[1]; ctx->stage_calls[ST_PROJ_SWEEPS] += P.h_counts[3]; }
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
