// Projection of element Hessians to positive definiteness.
//
// Replaces project_to_PD_inplace (symx/solver/second_order/project_to_PD.cpp:13-82, Eigen::SelfAdjointEigenSolver on
// n in {3,6,9,12,15} or dynamic) and the selection logic of ElementHessians::{project_to_PD_inplace__all,
// _project_to_PD_for_update} (ElementHessians.cpp:48-182) together with the PPN block selection of
// NewtonsMethod::_project_and_assemble (symx/solver/NewtonsMethod.cpp:316-327).
//
// A group of lanes per selected element (a warp for short lists, 8 lanes for long ones): PD test by LDL^T, else a
// parallel-order Jacobi eigen-decomposition in shared memory, eigenvalues below eps clamped to eps (or mirrored),
// H = V diag(l) V^T rebuilt only if something changed -- the projected matrix replaces the original in the element-Hessian
// store and the next numeric assembly re-sums the BCSR blocks it touches (instead of the reference's "add (projected -
// original)" update pass; same matrix up to float rounding).
#include "internal.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdlib>

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

// How far does the PPN threshold have to fall before the selection changes?  m = max over not-yet-projected elements of the largest
// |g_b|_inf of their blocks (a threshold above m selects nothing); g_min = min over all blocks of |g_b|_inf (at or below it every block
// is active: "all projected").  Non-negative doubles compare like their bit patterns, so integer atomics do.
__global__ void k_selection_bounds(const ProjTable* __restrict__ Tp, const int32_t* __restrict__ rows_all, const double* __restrict__ grad, const uint8_t* __restrict__ projected,
                                   unsigned long long n_elem_total, int nbr, unsigned long long* __restrict__ out /* [0] m bits, [1] g_min bits */)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_elem_total && !projected[i]) {
        const ProjTable& T = *Tp;
        const int pi = find_pot(T, i);
        const int nb = T.nb[pi];
        const int32_t* r = rows_all + T.rows_off[pi] + (i - T.E_off[pi]) * nb;
        double m = 0.0;
        for (int b = 0; b < nb; b++) { const int row = r[b]; m = fmax(m, fmax(fabs(grad[3 * row]), fmax(fabs(grad[3 * row + 1]), fabs(grad[3 * row + 2])))); }
        atomicMax(out, (unsigned long long)__double_as_longlong(m));
    }
    if (i < (unsigned long long)nbr) {
        const double m = fmax(fabs(grad[3 * i]), fmax(fabs(grad[3 * i + 1]), fabs(grad[3 * i + 2])));
        atomicMin(out + 1, (unsigned long long)__double_as_longlong(m));
    }
}

// One GROUP of G lanes per selected element, 32 / G elements per warp, all groups of a warp in LOCKSTEP (one control flow per
// warp: groups that are done or idle run on with identity rotations, so __syncwarp() stays legal and nothing serialises):
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
//     computed from the same matrix and applied together, R-1 steps per sweep -- the dependent chain of a sweep is R-1
//     steps instead of the R(R-1)/2 rotations of the cyclic order.  A step is: rotation parameters (two rsqrt, no
//     division), then ONE pass in which every 2x2 block (pair i, pair j) of A takes both of its rotations and the columns of
//     V take theirs;
//  3. clamp / mirror the eigenvalues below eps and rebuild V diag(l) V^T.
// Two instances.  G = 32 (a warp per element) has the shortest dependent chain and serves short lists, which are bound by
// the latency of one element.  Long lists (a projection of most of the mesh) are bound by instruction issue -- a 9 x 9 step
// has 5 rotations, 25 blocks and 45 column pairs of V, so a warp per element issues most instructions for a handful of live
// lanes -- and go to G = 8, where four elements share every instruction.  Both are launched; each looks at the list length
// and one of them returns at once.
constexpr int PROJ_WARPS = 4;
constexpr int PROJ_THREADS = 32 * PROJ_WARPS;
constexpr int PROJ_LONG_LIST = 6000;   // lists longer than this take the G = 8 instance
template<int N> constexpr size_t proj_smem_per_group() { return sizeof(double) * (2 * N * N + N + 2 * ((N + 1) / 2)) + sizeof(int) * 2 * ((N + 1) / 2 + 1); }

struct ProjScratch {
    double* A; double* V; double* lam; double* cs_c; double* cs_s; int* pp; int* pq;
};

// 1 / sqrt((j + 1)(j + 2)): scale of column j of the Helmert basis
__constant__ double c_helmert_scale[8] = {0.70710678118654752440, 0.40824829046386301637, 0.28867513459481288225, 0.22360679774997896964,
                                          0.18257418583505537115, 0.15430334996209191026, 0.13363062095621219234, 0.11785113019775792073};

// PD test + eigen-projection of the symmetric R x R matrix in S.A (row-major, pitch R) of every group of the warp (groups
// with valid == false hold a zero matrix and only keep step).  Returns true when eigenvalues were clamped; S.A then holds
// the projected matrix.  Returns false when the matrix is left as it is.
template<int R, int G>
__device__ bool project_small(const ProjScratch& S, double eps, int mirror, int lane, bool valid, int& sweeps_done)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NE = (R + 1) & ~1;   // even number of players (a dummy index R when R is odd)
    constexpr int NP = NE / 2;         // pairs per step
    double* A = S.A; double* V = S.V;
    for (int k = lane; k < R * R; k += G) {
        const int i = k / R, j = k - i * R;
        V[k] = A[k] - ((i == j) ? eps : 0.0);
    }
    __syncwarp();
    bool pd = true;
    for (int j = 0; j < R; j++) {
        const double d = V[j * R + j];            // shared value: uniform across the group
        if (!(d > 0.0)) pd = false;               // (the factorisation runs on: its results are not used any more)
        const double inv = 1.0 / d;
        const int m = R - 1 - j;                  // trailing block: rows / cols j+1 .. R-1 (lower triangle incl. diagonal)
        for (int t = lane; t < m * m; t += G) {
            const int i = j + 1 + t / m, k2 = j + 1 + t % m;
            if (k2 <= i) V[i * R + k2] -= V[i * R + j] * V[k2 * R + j] * inv;
        }
        __syncwarp();
    }
    const bool need = valid && !pd;
    if (!__any_sync(FULL, need)) return false;
    for (int k = lane; k < R * R; k += G) {
        const int i = k / R, j = k - i * R;
        V[k] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();
    bool conv = !need;
    for (int sweep = 0; sweep < 30; sweep++) {
        // convergence: off-diagonal mass vs total
        double off = 0.0, diag = 0.0;
        for (int k = lane; k < R * R; k += G) {
            const int i = k / R, j = k - i * R;
            const double v = A[k] * A[k];
            if (i == j) diag += v; else off += v;
        }
        for (int o = G / 2; o > 0; o >>= 1) { off += __shfl_xor_sync(FULL, off, o); diag += __shfl_xor_sync(FULL, diag, o); }
        if (off <= 1e-25 * (diag + off) || off == 0.0) conv = true;   // |off| / |A| <= 3e-13: eigenvalue error ~ |off|^2 / gap, far below 1e-10 parity
        if (__all_sync(FULL, conv)) break;
        if (!conv) sweeps_done = sweep + 1;
        for (int step = 0; step < NE - 1; step++) {
            // round-robin pairing: player NE-1 is fixed, the others rotate
            for (int pr = lane; pr < NP; pr += G) {
                int p, q;
                if (pr == 0) { p = NE - 1; q = step; }
                else { p = (step + pr) % (NE - 1); q = (step - pr + (NE - 1)) % (NE - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, sn = 0.0;
                if (q < R && !conv) {   // (a pair with the dummy index does nothing; a converged matrix is left alone)
                    const double apq = A[p * R + q];
                    if (apq != 0.0) {
                        // rotation by the angle |theta| <= pi/4 with tan(2 theta) = 2 apq / (aqq - app):
                        //   cos(2 theta) = |a| / h,  a = aqq - app, h = hypot(a, 2 apq)
                        //   c = sqrt((1 + |a| / h) / 2),  s = sign(a) apq / (h c)
                        const double a = A[q * R + q] - A[p * R + p], b = 2.0 * apq;
                        const double x = a * a + b * b;
                        if (x > 1e-290 && x < 1e290) {
                            const double rh = rsqrt(x);
                            const double w = 0.5 + 0.5 * fabs(a) * rh;
                            const double rc = rsqrt(w);
                            c = w * rc;
                            sn = (a >= 0.0 ? 0.5 : -0.5) * b * rh * rc;
                        } else {   // out of the range of the squared form: the classical formulas
                            const double tau = a / b;
                            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                            c = rsqrt(1.0 + t * t);
                            sn = t * c;
                        }
                    }
                }
                S.pp[pr] = p; S.pq[pr] = (q < R) ? q : -1; S.cs_c[pr] = c; S.cs_s[pr] = sn;
            }
            __syncwarp();
            // A <- J^T A J, block (pair i, pair j) at a time: columns by pair j's rotation, then rows by pair i's
            for (int t = lane; t < NP * NP; t += G) {
                const int bi = t / NP, bj = t - bi * NP;
                const int pi = S.pp[bi], qi = S.pq[bi], pj = S.pp[bj], qj = S.pq[bj];
                const double ci = S.cs_c[bi], si = S.cs_s[bi], cj = S.cs_c[bj], sj = S.cs_s[bj];
                const double a_pp = A[pi * R + pj];
                const double a_pq = (qj >= 0) ? A[pi * R + qj] : 0.0;
                const double a_qp = (qi >= 0) ? A[qi * R + pj] : 0.0;
                const double a_qq = (qi >= 0 && qj >= 0) ? A[qi * R + qj] : 0.0;
                const double t_pp = cj * a_pp - sj * a_pq, t_pq = sj * a_pp + cj * a_pq;
                const double t_qp = cj * a_qp - sj * a_qq, t_qq = sj * a_qp + cj * a_qq;
                A[pi * R + pj] = ci * t_pp - si * t_qp;
                if (qj >= 0) A[pi * R + qj] = ci * t_pq - si * t_qq;
                if (qi >= 0) A[qi * R + pj] = si * t_pp + ci * t_qp;
                if (qi >= 0 && qj >= 0) A[qi * R + qj] = si * t_pq + ci * t_qq;
            }
            // V <- V J (all rows)
            for (int t = lane; t < NP * R; t += G) {
                const int pr = t / R, k = t - pr * R;
                const int p = S.pp[pr], q = S.pq[pr];
                if (q < 0) continue;
                const double c = S.cs_c[pr], sn = S.cs_s[pr];
                const double vkp = V[k * R + p], vkq = V[k * R + q];
                V[k * R + p] = c * vkp - sn * vkq;
                V[k * R + q] = sn * vkp + c * vkq;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    // clamp / mirror
    bool changed = false;
    if (need)
        for (int i = 0; i < R; i++) if (A[i * R + i] < eps) changed = true;   // shared values: uniform across the group
    if (!__any_sync(FULL, changed)) return false;
    for (int i = lane; i < R; i += G) {
        const double l = A[i * R + i];
        S.lam[i] = (l < eps) ? (mirror ? -l : eps) : l;
    }
    __syncwarp();
    for (int k = lane; k < R * R; k += G) {
        const int i = k / R, j = k - i * R;
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < R; m++) acc += V[i * R + m] * S.lam[m] * V[j * R + m];
        if (changed) A[k] = acc;   // (reads V and lam only)
    }
    __syncwarp();
    return changed;
}

// Helmert basis of nn nodes: column j (0 <= j < nn-1) = (1, .., 1, -(j+1), 0, ..) / sqrt((j+1)(j+2)) with j+1 leading ones
__device__ __forceinline__ double helmert(int a, int j)
{
    const double s = c_helmert_scale[j];
    return (a <= j) ? s : ((a == j + 1) ? -(double)(j + 1) * s : 0.0);
}

// every group of the warp calls this together with the same N; groups with valid == false have nothing to do in this pass
template<int N, int G>
__device__ void project_item(unsigned char* base, const ProjTable& T, int pi, unsigned long long e, bool valid, double* __restrict__ H_all, double eps, int mirror,
                             int* __restrict__ n_changed, const DirtyView& dv, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
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
    double* H = valid ? H_all + T.H_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(N * N) : nullptr;
    const bool invariant = (N > 3) && valid && T.invariant[pi];
    // the deflated and the plain path are taken in turn when the warp holds both kinds (same N, e.g. 4-node contact pairs with
    // and without a rigid body never share N, so in practice one pass)
    for (int pass = 0; pass < 2; pass++) {
        const bool mine = valid && (invariant == (pass == 0));
        if (!__any_sync(FULL, mine)) continue;
        __syncwarp();   // previous item fully done
        for (int k = lane; k < N * N; k += G) {   // load (symmetrised)
            const int i = k / N, j = k - i * N;
            A[k] = mine ? 0.5 * (H[i * N + j] + H[j * N + i]) : 0.0;
        }
        __syncwarp();
        int sweeps = 0;
        bool changed;
        if (pass == 0) {
            // T1 = H Q  (N x R):  T1[(a r), (j s)] = sum_b H[(a r), (b s)] W[b][j]
            for (int k = lane; k < N * R; k += G) {
                const int row = k / R, col = k - row * R;
                const int j = col / 3, sc = col - 3 * j;
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < NN; b++) acc += A[row * N + 3 * b + sc] * helmert(b, j);
                V[k] = acc;
            }
            __syncwarp();
            // M = Q^T T1  (R x R):  M[(i r), c] = sum_a W[a][i] T1[(a r), c]
            for (int k = lane; k < R * R; k += G) {
                const int row = k / R, col = k - row * R;
                const int i = row / 3, rc = row - 3 * i;
                double acc = 0.0;
#pragma unroll
                for (int a = 0; a < NN; a++) acc += helmert(a, i) * V[(3 * a + rc) * R + col];
                A[k] = acc;
            }
            __syncwarp();
            changed = project_small<R, G>(S, eps, mirror, lane, mine, sweeps);
            if (__any_sync(FULL, changed)) {
                // T1' = Q M'  (N x R), then H = T1' Q^T + eps (I - Q Q^T);  I - Q Q^T = (1/nn) ones (x) I3
                for (int k = lane; k < N * R; k += G) {
                    const int row = k / R, col = k - row * R;
                    const int a = row / 3, rc = row - 3 * a;
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NN - 1; i++) acc += helmert(a, i) * A[(3 * i + rc) * R + col];
                    V[k] = acc;
                }
                __syncwarp();
                for (int k = lane; k < N * N; k += G) {
                    const int row = k / N, col = k - row * N;
                    const int b = col / 3, sc = col - 3 * b;
                    double acc = ((row % 3) == sc) ? eps / (double)NN : 0.0;
#pragma unroll
                    for (int j = 0; j < NN - 1; j++) acc += V[row * R + 3 * j + sc] * helmert(b, j);
                    if (changed) H[k] = acc;
                }
            }
        } else {
            changed = project_small<N, G>(S, eps, mirror, lane, mine, sweeps);
            if (changed)
                for (int k = lane; k < N * N; k += G) H[k] = A[k];
        }
        if (lane == 0 && sweeps) atomicAdd(n_changed + 2, sweeps);   // diagnostic: total Jacobi sweeps (d_counts[3])
        if (changed) {
            if (lane == 0) atomicAdd(n_changed, 1);
            if (dv.dirty) {   // the BCSR blocks this element contributes to must be re-summed
                constexpr int nb = N / 3;
                const unsigned long long src0 = T.blk_off[pi] + (e - T.E_off[pi]) * (unsigned long long)(nb * nb);
                for (int k = lane; k < nb * nb; k += G) {
                    const unsigned long long src = src0 + k;
                    const uint32_t f = (src < dv.n_static) ? dv.s_final[dv.s_blk_of_src[src]] : dv.d_final_of_src[src - dv.n_static];
                    dv.dirty[f] = 1;
                }
            }
        }
        __syncwarp();
    }
}

// one launch for all element sizes.  The groups of a warp take consecutive list items; the warp then works through the
// element sizes present among them, one size at a time (usually one: neighbours in the list come from the same potential).
template<int G>
__global__ void __launch_bounds__(PROJ_THREADS) k_project(const ProjTable* __restrict__ Tp, double* __restrict__ H_all, const uint32_t* __restrict__ list,
                                                           const int* __restrict__ n_list, double eps, int mirror, int* __restrict__ n_changed,
                                                           const DirtyView dv, int smem_per_group)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int ITEMS = PROJ_THREADS / G;   // elements in flight per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int group = threadIdx.x / G, lane = threadIdx.x % G;
    unsigned char* base = smem_raw + (size_t)group * smem_per_group;
    const ProjTable& T = *Tp;
    const int total = *n_list;
    if ((G == 32) != (total <= PROJ_LONG_LIST)) return;   // the other instance takes this list
    const int warp_first = (threadIdx.x / 32) * (32 / G);  // first group of this warp
    for (int item0 = blockIdx.x * ITEMS + warp_first; item0 < total; item0 += gridDim.x * ITEMS) {   // (warp-uniform loop)
        const int item = item0 + (group - warp_first);
        unsigned long long e = 0;
        int pi = 0, nb = 0;
        if (item < total) { e = list[item]; pi = find_pot(T, e); nb = T.nb[pi]; }
        bool pending = nb > 0;
        while (true) {
            const unsigned m = __ballot_sync(FULL, pending);
            if (!m) break;
            const int nb_cur = __shfl_sync(FULL, nb, __ffs(m) - 1);
            const bool valid = pending && nb == nb_cur;
            switch (nb_cur) {
            case 1: project_item<3, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            case 2: project_item<6, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            case 3: project_item<9, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            case 4: project_item<12, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            case 5: project_item<15, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            case 6: project_item<18, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            case 7: project_item<21, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            default: project_item<24, G>(base, T, pi, e, valid, H_all, eps, mirror, n_changed, dv, lane); break;
            }
            if (valid) pending = false;
        }
    }
}

static size_t proj_smem_for(int nb)
{
    switch (nb) {
    case 1: return proj_smem_per_group<3>();
    case 2: return proj_smem_per_group<6>();
    case 3: return proj_smem_per_group<9>();
    case 4: return proj_smem_per_group<12>();
    case 5: return proj_smem_per_group<15>();
    case 6: return proj_smem_per_group<18>();
    case 7: return proj_smem_per_group<21>();
    default: return proj_smem_per_group<24>();
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
    ProjTable table_host;
    bool table_valid = false;
    int* d_counts = nullptr;   // [0] n_list, [1] n_changed, [2] n_inactive
    int* h_counts = nullptr;
    size_t smem_configured = 0;
    unsigned long long* d_bounds = nullptr;   // project_selection_bounds
    unsigned long long* h_bounds = nullptr;
};
void projector_destroy(sb_context* ctx)
{
    Projector* P = ctx->projector;
    if (!P) return;
    P->active.release(); P->list.release();
    if (P->d_table) cudaFree(P->d_table);
    if (P->d_counts) cudaFree(P->d_counts);
    if (P->h_counts) cudaFreeHost(P->h_counts);
    if (P->d_bounds) cudaFree(P->d_bounds);
    if (P->h_bounds) cudaFreeHost(P->h_bounds);
    delete P;
    ctx->projector = nullptr;
}

void preload_project_kernels()
{
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_active_blocks); cudaFuncGetAttributes(&fa, k_select);
    cudaFuncGetAttributes(&fa, k_project<32>); cudaFuncGetAttributes(&fa, k_project<8>);
    cudaGetLastError();
}

// buffers of the projector, sized for the current evaluation (called from the P+G+H evaluation so that the first projection
// of a run -- typically the first iteration in contact -- does not pay for pinned-memory and device allocations)
void projector_prepare(sb_context* ctx)
{
    if (!ctx->projector) ctx->projector = new Projector();
    Projector& P = *ctx->projector;
    if (!P.d_table) {
        cudaMalloc(&P.d_table, sizeof(ProjTable));
        cudaMalloc(&P.d_counts, 4 * sizeof(int));
        cudaMallocHost(&P.h_counts, 4 * sizeof(int));
    }
    P.active.ensure(ctx->ndofs / 3 + 1);
    P.list.ensure(ctx->n_hessians + 1);
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
    ctx->pgh_cache_ok = false;   // the element Hessians are about to be modified in place
    const auto t_begin = std::chrono::steady_clock::now();
    ProjTable T;
    std::memset(&T, 0, sizeof(T));   // (compared bytewise with the uploaded copy)
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
    // (uploaded only when it changed: a copy from pageable memory blocks the host, and the PPN loop calls this several times
    //  per Newton iteration with the same layout)
    if (!P.table_valid || std::memcmp(&P.table_host, &T, sizeof(ProjTable)) != 0) {
        SB_CUDA(ctx, cudaMemcpyAsync(P.d_table, &T, sizeof(ProjTable), cudaMemcpyHostToDevice, st));
        P.table_host = T;
        P.table_valid = true;
    }
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
    DirtyView dv;
    if (!assembly_dirty_view(ctx, &dv)) dv.dirty = nullptr;
    int nb_max = 1;
    for (auto& p : ctx->potentials) if (p.n_elem > 0) nb_max = std::max(nb_max, p.k->nb);
    const size_t per_group = (proj_smem_for(nb_max) + 15) & ~(size_t)15;   // shared memory (per group) sized for the largest element present
    const size_t smem32 = (PROJ_THREADS / 32) * per_group, smem8 = (PROJ_THREADS / 8) * per_group;
    if (smem8 > P.smem_configured) {
        SB_CUDA(ctx, cudaFuncSetAttribute(k_project<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem32, 48 * 1024)));
        SB_CUDA(ctx, cudaFuncSetAttribute(k_project<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem8, 48 * 1024)));
        P.smem_configured = smem8;
    }
    // both instances; the list length (known on the device only) decides which one works
    const int grid32 = (int)std::min<size_t>((std::min<size_t>(n_elem, PROJ_LONG_LIST) + PROJ_THREADS / 32 - 1) / (PROJ_THREADS / 32), 148 * 16);
    const int grid8 = (int)std::min<size_t>((n_elem + PROJ_THREADS / 8 - 1) / (PROJ_THREADS / 8), 148 * 8);
    const int grid = grid32;
    const size_t smem = smem8;
    k_project<32><<<grid32, PROJ_THREADS, smem32, st>>>(P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv, (int)per_group);
    if (n_elem > (size_t)PROJ_LONG_LIST) {
        k_project<8><<<grid8, PROJ_THREADS, smem8, st>>>(P.d_table, ctx->H.p, P.list.p, P.d_counts, eps, mirror, P.d_counts + 1, dv, (int)per_group);
        ctx->launches++;
    }
    SB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->launches += 1;
    SB_CUDA(ctx, cudaMemcpyAsync(P.h_counts, P.d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, hot_sync(ctx));   // also protects the stack-resident table T
    SB_CUDA(ctx, cudaGetLastError());
    ctx->n_projected += P.h_counts[0];
    static const bool dump = std::getenv("SB_PROJ_DUMP") != nullptr;   // per-call diagnostics on stderr
    if (dump) fprintf(stderr, "PROJDUMP selected=%d changed=%d sweeps=%d inactive_blocks=%d threshold=%g grid=%d smem=%zu us=%.1f\n", P.h_counts[0], P.h_counts[1], P.h_counts[3], P.h_counts[2], grad_threshold, grid, smem,
                      1e6 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    if (ctx->profile) { ctx->stage_calls[ST_PROJ_SELECTED] += P.h_counts[0]; ctx->stage_calls[ST_PROJ_CHANGED] += P.h_counts[1]; ctx->stage_calls[ST_PROJ_SWEEPS] += P.h_counts[3]; }
    if (out_n_projected) *out_n_projected = ctx->n_projected;
    if (out_all_projected) *out_all_projected = use_active ? (P.h_counts[2] == 0) : 1;
    return 0;
}

// See k_selection_bounds.  Valid for the projection flags as they are now (call it right after a projection that selected nothing).
// out_m < 0: every element is projected already.
int project_selection_bounds(sb_context* ctx, double* out_m, double* out_gmin)
{
    Projector* Pp = ctx->projector;
    if (!Pp || !Pp->table_valid || !ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "project_selection_bounds: no projection has run on this evaluation");
    Projector& P = *Pp;
    cudaStream_t st = ctx->stream;
    if (!P.d_bounds) { SB_CUDA(ctx, cudaMalloc(&P.d_bounds, 2 * sizeof(unsigned long long))); SB_CUDA(ctx, cudaMallocHost(&P.h_bounds, 2 * sizeof(unsigned long long))); }
    P.h_bounds[0] = 0ull; P.h_bounds[1] = ~0ull;   // sentinels the kernel can only raise / lower (pinned: the copy is asynchronous)
    SB_CUDA(ctx, cudaMemcpyAsync(P.d_bounds, P.h_bounds, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    const unsigned long long n_elem = ctx->n_hessians;
    const int nbr = ctx->ndofs / 3;
    const unsigned long long n = std::max<unsigned long long>(n_elem, (unsigned long long)nbr);
    // "no unprojected element" must be told apart from m == 0: count them through the sign of a separate launch-free trick --
    // the kernel raises out[0] from 0 only for unprojected elements with m > 0; elements with a zero gradient cannot be told from
    // none, and both mean the same for the caller (no positive threshold selects anything new)
    k_selection_bounds<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.d_table, ctx->rows.p, ctx->grad.p, ctx->projected.p, n_elem, nbr, P.d_bounds);
    ctx->launches++;
    SB_CUDA(ctx, cudaMemcpyAsync(P.h_bounds, P.d_bounds, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    SB_CUDA(ctx, cudaGetLastError());
    long long mb = (long long)P.h_bounds[0], gb = (long long)P.h_bounds[1];
    double m, g;
    std::memcpy(&m, &mb, sizeof(double)); std::memcpy(&g, &gb, sizeof(double));
    if (P.h_bounds[1] == ~0ull) g = 0.0;
    *out_m = m;
    *out_gmin = g;
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_project_to_pd(sb_context* ctx, double grad_threshold, double eps, int mirror, int64_t* out_n_projected, int64_t* out_n_hessians, int* out_all_projected)
{
    if (!ctx) return SB_ERR_ARG;
    return project_internal(ctx, grad_threshold, eps, mirror, out_n_projected, out_n_hessians, out_all_projected);
}
