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
    int invariant[PROJ_MAX_POTS];                  // energy invariant under a common translation of all its nodes
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
//  0. TRANSLATION-INVARIANT potentials (strain, bending, deformable-deformable contact / friction / attachments: the energy
//     depends on differences of the nodal DoFs only) have three exact null vectors (the rigid translations), so A - eps I is
//     never positive definite and the cheap exit below could never fire.  Their Hessian is first compressed to the
//     3 (nn - 1)-dimensional complement, M = Q^T H Q with Q = W (x) I3 and W the fixed Helmert basis of the nn nodes
//     (orthonormal, orthogonal to (1..1)); eig(H) = eig(M) + three zeros, so H_proj = Q M_proj Q^T + eps (I - Q Q^T) is the
//     same matrix the reference's 12 x 12 eigen-solve produces (its three null eigenvalues, +-1e-17 |H|, are clamped to
//     eps = 1e-10 as well), up to rounding.  A tet becomes a 9 x 9 problem and, when the tet is not inverted / buckled,
//     exits after the factorisation;
//  1. cheap exit: lambda_min(A) > eps  <=>  A - eps I has an LDL^T factorisation with positive pivots (R steps);
//  2. otherwise a PARALLEL-ORDER two-sided Jacobi eigen-solve: the R/2 disjoint rotations of one round-robin step are
//     computed from the same matrix and applied together (columns of A and V, then rows of A), R-1 steps per sweep --
//     the dependent chain of a sweep is R-1 steps instead of the R(R-1)/2 rotations of the cyclic order;
//  3. clamp / mirror the eigenvalues below eps and rebuild V diag(l) V^T.
constexpr int PROJ_WARPS = 8;
constexpr int PROJ_THREADS = 32 * PROJ_WARPS;
template<int N> constexpr size_t proj_smem_per_warp() { return sizeof(double) * (2 * N * N + N + 2 * ((N + 1) / 2)) + sizeof(int) * 2 * ((N + 1) / 2 + 1); }

struct ProjScratch {
    double* A; double* V; double* lam; double* cs_c; double* cs_s; int* pp; int* pq;
};

// PD test + eigen-projection of the symmetric R x R matrix in S.A (row-major, pitch R).  Returns true when eigenvalues were
// clamped; S.A then holds the projected matrix.  Returns false when the matrix is left as it is.
template<int R>
__device__ bool project_small(const ProjScratch& S, double eps, int mirror, int lane, int& sweeps_done)
{
    constexpr int NE = (R + 1) & ~1;   // even number of players (a dummy index R when R is odd)
    constexpr int NP = NE / 2;         // pairs per step
    double* A = S.A; double* V = S.V;
    for (int k = lane; k < R * R; k += 32) {
        const int i = k / R, j = k - i * R;
        V[k] = A[k] - ((i == j) ? eps : 0.0);
    }
    __syncwarp();
    bool pd = true;
    for (int j = 0; j < R; j++) {
        const double d = V[j * R + j];
        if (!(d > 0.0)) { pd = false; break; }   // shared value: uniform across the warp
        const double inv = 1.0 / d;
        const int m = R - 1 - j;                  // trailing block: rows / cols j+1 .. R-1 (lower triangle incl. diagonal)
        for (int t = lane; t < m * m; t += 32) {
            const int i = j + 1 + t / m, k2 = j + 1 + t % m;
            if (k2 <= i) V[i * R + k2] -= V[i * R + j] * V[k2 * R + j] * inv;
        }
        __syncwarp();
    }
    if (pd) return false;
    __syncwarp();
    for (int k = lane; k < R * R; k += 32) {
        const int i = k / R, j = k - i * R;
        V[k] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 30; sweep++) {
        sweeps_done = sweep;
        // convergence: off-diagonal mass vs total
        double off = 0.0, diag = 0.0;
        for (int k = lane; k < R * R; k += 32) {
            const int i = k / R, j = k - i * R;
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
                if (q < R) {   // (a pair with the dummy index does nothing)
                    const double apq = A[p * R + q];
                    if (apq != 0.0) {
                        const double app = A[p * R + p], aqq = A[q * R + q];
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = rsqrt(1.0 + t * t);
                        sn = t * c;
                    }
                }
                S.pp[lane] = p; S.pq[lane] = (q < R) ? q : -1; S.cs_c[lane] = c; S.cs_s[lane] = sn;
            }
            __syncwarp();
            // columns p, q of A and V (all rows)
            for (int t = lane; t < NP * R; t += 32) {
                const int pr = t / R, k = t - pr * R;
                const int p = S.pp[pr], q = S.pq[pr];
                if (q < 0) continue;
                const double c = S.cs_c[pr], sn = S.cs_s[pr];
                const double akp = A[k * R + p], akq = A[k * R + q];
                A[k * R + p] = c * akp - sn * akq;
                A[k * R + q] = sn * akp + c * akq;
                const double vkp = V[k * R + p], vkq = V[k * R + q];
                V[k * R + p] = c * vkp - sn * vkq;
                V[k * R + q] = sn * vkp + c * vkq;
            }
            __syncwarp();
            // rows p, q of A (all columns)
            for (int t = lane; t < NP * R; t += 32) {
                const int pr = t / R, k = t - pr * R;
                const int p = S.pp[pr], q = S.pq[pr];
                if (q < 0) continue;
                const double c = S.cs_c[pr], sn = S.cs_s[pr];
                const double apk = A[p * R + k], aqk = A[q * R + k];
                A[p * R + k] = c * apk - sn * aqk;
                A[q * R + k] = sn * apk + c * aqk;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    // clamp / mirror
    bool changed = false;
    for (int i = 0; i < R; i++) if (A[i * R + i] < eps) changed = true;   // shared values: uniform across the warp
    if (!changed) return false;
    if (lane < R) {
        const double l = A[lane * R + lane];
        S.lam[lane] = (l < eps) ? (mirror ? -l : eps) : l;
    }
    __syncwarp();
    for (int k = lane; k < R * R; k += 32) {
        const int i = k / R, j = k - i * R;
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < R; m++) acc += V[i * R + m] * S.lam[m] * V[j * R + m];
        A[k] = acc;   // (reads V and lam only)
    }
    __syncwarp();
    return true;
}

// Helmert basis of nn nodes: column j (0 <= j < nn-1) = (1, .., 1, -(j+1), 0, ..) / sqrt((j+1)(j+2)) with j+1 leading ones
__device__ __forceinline__ double helmert(int a, int j)
{
    const double s = rsqrt((double)((j + 1) * (j + 2)));
    return (a <= j) ? s : ((a == j + 1) ? -(double)(j + 1) * s : 0.0);
}

template<int N>
__device__ void project_item(unsigned char* base, const ProjTable& T, int pi, unsigned long long e, double* __restrict__ H_all, double eps, int mirror,
                             int* __restrict__ n_changed, const DirtyView& dv, int lane)
{
    constexpr int NN = N / 3;                       // nodes
    constexpr int R = (N > 3) ? N - 3 : N;          // size after removing the translations
    ProjScratch S;
    S.A = reinterpret_cast<double*>(base);
    S.V = S.A + N * N;
    S.lam = S.V + N * N;
    S.cs_c = S.lam + N;
    S.cs_s = S.cs_c + (N + 1) / 2;
    S.pp = reinterpret_cast<int*>(S.cs_s + (N + 1) / 2);
    S.pq = S.pp + (N + 1) / 2 + 1;
    double* A = S.A; double* V = S.V;
    {
        double* H = H_all + T.H_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(N * N);
        const bool invariant = (N > 3) && T.invariant[pi];
        __syncwarp();   // previous item fully done
        for (int k = lane; k < N * N; k += 32) {   // load (symmetrised)
            const int i = k / N, j = k - i * N;
            A[k] = 0.5 * (H[i * N + j] + H[j * N + i]);
        }
        __syncwarp();
        int sweeps = 0;
        bool changed;
        if (invariant) {
            // T1 = H Q  (N x R):  T1[(a r), (j s)] = sum_b H[(a r), (b s)] W[b][j]
            for (int k = lane; k < N * R; k += 32) {
                const int row = k / R, col = k - row * R;
                const int j = col / 3, sc = col - 3 * j;
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < NN; b++) acc += A[row * N + 3 * b + sc] * helmert(b, j);
                V[k] = acc;
            }
            __syncwarp();
            // M = Q^T T1  (R x R):  M[(i r), c] = sum_a W[a][i] T1[(a r), c]
            for (int k = lane; k < R * R; k += 32) {
                const int row = k / R, col = k - row * R;
                const int i = row / 3, rc = row - 3 * i;
                double acc = 0.0;
#pragma unroll
                for (int a = 0; a < NN; a++) acc += helmert(a, i) * V[(3 * a + rc) * R + col];
                A[k] = acc;
            }
            __syncwarp();
            changed = project_small<R>(S, eps, mirror, lane, sweeps);
            if (changed) {
                // T1' = Q M'  (N x R), then H = T1' Q^T + eps (I - Q Q^T);  I - Q Q^T = (1/nn) ones (x) I3
                for (int k = lane; k < N * R; k += 32) {
                    const int row = k / R, col = k - row * R;
                    const int a = row / 3, rc = row - 3 * a;
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NN - 1; i++) acc += helmert(a, i) * A[(3 * i + rc) * R + col];
                    V[k] = acc;
                }
                __syncwarp();
                for (int k = lane; k < N * N; k += 32) {
                    const int row = k / N, col = k - row * N;
                    const int b = col / 3, sc = col - 3 * b;
                    double acc = ((row % 3) == sc) ? eps / (double)NN : 0.0;
#pragma unroll
                    for (int j = 0; j < NN - 1; j++) acc += V[row * R + 3 * j + sc] * helmert(b, j);
                    H[k] = acc;
                }
            }
        } else {
            changed = project_small<N>(S, eps, mirror, lane, sweeps);
            if (changed)
                for (int k = lane; k < N * N; k += 32) H[k] = A[k];
        }
        if (lane == 0 && sweeps) atomicAdd(n_changed + 2, sweeps);   // diagnostic: total Jacobi sweeps (d_counts[3])
        if (changed) {
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

// one launch for all element sizes: the warp dispatches on the size of its element (uniform across the warp)
__global__ void __launch_bounds__(PROJ_THREADS) k_project(const ProjTable* __restrict__ Tp, double* __restrict__ H_all, const uint32_t* __restrict__ list,
                                                           const int* __restrict__ n_list, double eps, int mirror, int* __restrict__ n_changed,
                                                           const DirtyView dv, int smem_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smem_raw + (size_t)warp * smem_per_warp;
    const ProjTable& T = *Tp;
    const int total = *n_list;
    for (int item = blockIdx.x * PROJ_WARPS + warp; item < total; item += gridDim.x * PROJ_WARPS) {
        const unsigned long long e = list[item];
        const int pi = find_pot(T, e);
        switch (T.nb[pi]) {
        case 1: project_item<3>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        case 2: project_item<6>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        case 3: project_item<9>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        case 4: project_item<12>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        case 5: project_item<15>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        case 6: project_item<18>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        case 7: project_item<21>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        default: project_item<24>(base, T, pi, e, H_all, eps, mirror, n_changed, dv, lane); break;
        }
    }
}

static size_t proj_smem_for(int nb)
{
    switch (nb) {
    case 1: return proj_smem_per_warp<3>();
    case 2: return proj_smem_per_warp<6>();
    case 3: return proj_smem_per_warp<9>();
    case 4: return proj_smem_per_warp<12>();
    case 5: return proj_smem_per_warp<15>();
    case 6: return proj_smem_per_warp<18>();
    case 7: return proj_smem_per_warp<21>();
    default: return proj_smem_per_warp<24>();
    }
}

// potentials whose DoF blocks are all deformable points and whose energy depends on their differences only
static bool translation_invariant(const std::string& name)
{
    static const char* prefixes[] = {"EnergyTetStrain", "EnergyTriangleStrain", "EnergySegmentStrain", "EnergyDiscreteShells", "EnergyBendingFlat",
                                     "EnergyAttachments_d_d", "contact_d_d", "friction_d_d"};
    for (const char* p : prefixes) if (name.rfind(p, 0) == 0) return true;
    return false;
}

struct Projector {
    DevBuf<uint8_t> active;
    DevBuf<uint32_t> list;
    ProjTable* d_table = nullptr;
    int* d_counts = nullptr;   // [0] n_list, [1] n_changed, [2] n_inactive
    int* h_counts = nullptr;
    size_t smem_configured = 0;
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
        T.invariant[T.n_pots] = translation_invariant(p.name) ? 1 : 0;
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
    int nb_max = 1;
    for (auto& p : ctx->potentials) if (p.n_elem > 0) nb_max = std::max(nb_max, p.k->nb);
    const size_t per_warp = (proj_smem_for(nb_max) + 15) & ~(size_t)15;   // shared memory sized for the largest element present
    const size_t smem = PROJ_WARPS * per_warp;
    if (smem > P.smem_configured) {
        SB_CUDA(ctx, cudaFuncSetAttribute(k_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        P.smem_configured = smem;
    }
    k_project<<<grid, PROJ_THREADS, smem, st>>>(P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv, (int)per_warp);
    SB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
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
