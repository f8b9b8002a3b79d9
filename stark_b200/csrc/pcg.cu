// Block-Jacobi preconditioned conjugate gradient on the float-stored 3x3-BCSR Hessian.
//
// Replaces bsm::solve_pcg (bsm/solve_pcg.h:83-232), BlockedSparseMatrix::{prepare_preconditioning,
// apply_preconditioning, _spmxv} (bsm/BlockedSparseMatrix.h:1147-1361, 988-1138) and the vector kernels of
// bsm/bsm_vector_ops.h.  Same recurrence, same stopping rules (zero RHS, error < abs_tol, error/error_0 < rel_tol,
// p^T A p <= 0 -> "indefinite", max_iter), same float-storage / double-accumulate arithmetic.
//
// The whole working set (matrix ~20 MB at the 200k-tet scene + 6 vectors) lives in the 126 MB L2, so an iteration is
// bound by launch + reduction latency, not HBM.  Every iteration is therefore three fused kernels whose scalars
// (alpha, beta, error, flags) never leave the device: each CTA re-reduces the fixed-size partial-sum arrays it needs,
// and the host only reads the status word once per batch of iterations.  The fixed grid makes all dot products
// bitwise reproducible run to run (the reference's are thread-count dependent, bsm/ParallelNumber.h:39-47).
#include "internal.h"

namespace sb {

int bcsr_view(sb_context* ctx, int* nbr, size_t* nnzb, const unsigned long long** rows, const int32_t** cols, const float** vals);

constexpr int PCG_BLOCKS = 296;    // 2 CTAs per SM on 148 SMs
constexpr int PCG_THREADS = 256;
constexpr int LANES_PER_ROW = 8;   // lanes cooperating on one block row of the SpMV
constexpr int PCG_BATCH = 6;       // iterations launched between two host reads of the status word

// device-resident solver state
struct PcgState {
    double bb;          // ||b||^2
    double rz;          // r.z of the current iteration
    double error, error0;
    double abs_tol, rel_tol;
    int it;             // completed iterations
    int max_iter;
    int stop_on_indef;
    int done;           // 0 running, 1 converged, 2 indefinite, 3 max iterations
    int found_indef;
    int pad;
};

struct Pcg {
    DevBuf<double> r, z, p, Ap, x;
    DevBuf<float> dinv;
    DevBuf<double> part;       // 3 x PCG_BLOCKS partial sums
    PcgState* d_state = nullptr;
    PcgState* h_state = nullptr;
};
static Pcg* get(sb_context* ctx)
{
    if (!ctx->pcg) {
        ctx->pcg = new Pcg();
        cudaMalloc(&ctx->pcg->d_state, sizeof(PcgState));
        cudaMallocHost(&ctx->pcg->h_state, sizeof(PcgState));
    }
    return ctx->pcg;
}
void pcg_destroy(sb_context* ctx)
{
    Pcg* P = ctx->pcg;
    if (!P) return;
    P->r.release(); P->z.release(); P->p.release(); P->Ap.release(); P->x.release(); P->dinv.release(); P->part.release();
    if (P->d_state) cudaFree(P->d_state);
    if (P->h_state) cudaFreeHost(P->h_state);
    delete P;
    ctx->pcg = nullptr;
}

// ---- block reductions ------------------------------------------------------------------------------------------
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
        for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}
// every CTA sums the same PCG_BLOCKS partials in the same order -> identical value everywhere
__device__ __forceinline__ double all_partials(const double* __restrict__ part, double* s)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < PCG_BLOCKS; i += PCG_THREADS) v += part[i];
    const double t = block_sum(v, s);
    __shared__ double bc;
    if (threadIdx.x == 0) bc = t;
    __syncthreads();
    return bc;
}

// closed-form inverse of the symmetric 3x3 diagonal block, in float like the reference (BlockedSparseMatrix.h:1198-1214)
__global__ void k_block_jacobi(const unsigned long long* __restrict__ rows, const int32_t* __restrict__ cols, const float* __restrict__ vals,
                               float* __restrict__ dinv, int nbr)
{
    const int br = blockIdx.x * blockDim.x + threadIdx.x;
    if (br >= nbr) return;
    float* mi = dinv + 9 * (size_t)br;
    for (int k = 0; k < 9; k++) mi[k] = 0.0f;
    for (unsigned long long j = rows[br]; j < rows[br + 1]; j++) {
        if (cols[j] == 3 * br) {
            const float* m = vals + 9 * j;
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
            break;
        }
    }
}

__device__ __forceinline__ void apply_dinv(const float* __restrict__ d, double r0, double r1, double r2, double& z0, double& z1, double& z2)
{
    z0 = (double)d[0] * r0 + (double)d[1] * r1 + (double)d[2] * r2;
    z1 = (double)d[3] * r0 + (double)d[4] * r1 + (double)d[5] * r2;
    z2 = (double)d[6] * r0 + (double)d[7] * r1 + (double)d[8] * r2;
}

// x = 0, r = b = -grad, z = M^-1 r, p = z ; partials of b.b and r.z
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_init(const double* __restrict__ grad, const float* __restrict__ dinv,
                                                            double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                                                            double* __restrict__ part, int nbr)
{
    __shared__ double s[PCG_THREADS / 32];
    double bb = 0.0, rz = 0.0;
    for (int br = blockIdx.x * PCG_THREADS + threadIdx.x; br < nbr; br += PCG_BLOCKS * PCG_THREADS) {
        const double r0 = -grad[3 * br], r1 = -grad[3 * br + 1], r2 = -grad[3 * br + 2];
        double z0, z1, z2;
        apply_dinv(dinv + 9 * (size_t)br, r0, r1, r2, z0, z1, z2);
        x[3 * br] = 0.0; x[3 * br + 1] = 0.0; x[3 * br + 2] = 0.0;
        r[3 * br] = r0; r[3 * br + 1] = r1; r[3 * br + 2] = r2;
        z[3 * br] = z0; z[3 * br + 1] = z1; z[3 * br + 2] = z2;
        p[3 * br] = z0; p[3 * br + 1] = z1; p[3 * br + 2] = z2;
        bb += r0 * r0 + r1 * r1 + r2 * r2;
        rz += r0 * z0 + r1 * z1 + r2 * z2;
    }
    const double t0 = block_sum(bb, s);
    if (threadIdx.x == 0) part[blockIdx.x] = t0;
    const double t1 = block_sum(rz, s);
    if (threadIdx.x == 0) part[PCG_BLOCKS + blockIdx.x] = t1;
}
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_init_state(PcgState* st, const double* __restrict__ part, double abs_tol, double rel_tol, int max_iter, int stop_on_indef)
{
    __shared__ double s[PCG_THREADS / 32];
    const double bb = all_partials(part, s);
    const double rz = all_partials(part + PCG_BLOCKS, s);
    if (threadIdx.x == 0) {
        st->bb = bb; st->rz = rz;
        st->abs_tol = abs_tol; st->rel_tol = rel_tol; st->max_iter = max_iter; st->stop_on_indef = stop_on_indef;
        st->it = 0; st->found_indef = 0; st->pad = 0;
        st->error = 1.0; st->error0 = 1.0;   // x0 = 0 -> r = b
        st->done = 0;
        if (bb < abs_tol * abs_tol) { st->done = 1; st->error = 0.0; }   // zero right-hand side
        else if (1.0 < abs_tol) st->done = 1;
        else if (max_iter <= 0) st->done = 3;
    }
}

// Ap = A p with LANES_PER_ROW lanes per block row; partial of p.Ap
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_spmv(const PcgState* __restrict__ st, const unsigned long long* __restrict__ rows,
                                                            const int32_t* __restrict__ cols, const float* __restrict__ vals,
                                                            const double* __restrict__ p, double* __restrict__ Ap, double* __restrict__ part, int nbr)
{
    __shared__ double s[PCG_THREADS / 32];
    if (st->done) return;
    const int lane = threadIdx.x % LANES_PER_ROW;
    const int rows_per_cta = PCG_THREADS / LANES_PER_ROW;
    double pAp = 0.0;
    for (int base = blockIdx.x * rows_per_cta; base < nbr; base += PCG_BLOCKS * rows_per_cta) {
        const int br = base + threadIdx.x / LANES_PER_ROW;
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
        if (br < nbr) {
            const unsigned long long j1 = rows[br + 1];
            for (unsigned long long j = rows[br] + lane; j < j1; j += LANES_PER_ROW) {
                const float* m = vals + 9 * j;   // column-major 3x3
                const int c = cols[j];
                const double x0 = p[c], x1 = p[c + 1], x2 = p[c + 2];
                y0 += (double)m[0] * x0 + (double)m[3] * x1 + (double)m[6] * x2;
                y1 += (double)m[1] * x0 + (double)m[4] * x1 + (double)m[7] * x2;
                y2 += (double)m[2] * x0 + (double)m[5] * x1 + (double)m[8] * x2;
            }
        }
        for (int o = LANES_PER_ROW / 2; o > 0; o >>= 1) {
            y0 += __shfl_down_sync(0xffffffffu, y0, o, LANES_PER_ROW);
            y1 += __shfl_down_sync(0xffffffffu, y1, o, LANES_PER_ROW);
            y2 += __shfl_down_sync(0xffffffffu, y2, o, LANES_PER_ROW);
        }
        if (lane == 0 && br < nbr) {
            Ap[3 * br] = y0; Ap[3 * br + 1] = y1; Ap[3 * br + 2] = y2;
            pAp += p[3 * br] * y0 + p[3 * br + 1] * y1 + p[3 * br + 2] * y2;
        }
    }
    const double t = block_sum(pAp, s);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// alpha = rz / pAp ; x += alpha p ; r -= alpha Ap ; z = M^-1 r ; partials of r.r and r.z
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_update(const PcgState* __restrict__ st, const float* __restrict__ dinv,
                                                              const double* __restrict__ p, const double* __restrict__ Ap,
                                                              double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                              double* __restrict__ part, int nbr)
{
    __shared__ double s[PCG_THREADS / 32];
    if (st->done) return;
    const double pAp = all_partials(part, s);
    if (pAp <= 0.0 && st->stop_on_indef) return;   // x is returned as is (solve_pcg.h:183-192); k_pcg_direction records the status
    const double alpha = st->rz / pAp;
    double rr = 0.0, rz = 0.0;
    for (int br = blockIdx.x * PCG_THREADS + threadIdx.x; br < nbr; br += PCG_BLOCKS * PCG_THREADS) {
        double r0 = r[3 * br], r1 = r[3 * br + 1], r2 = r[3 * br + 2];
        x[3 * br] += alpha * p[3 * br]; x[3 * br + 1] += alpha * p[3 * br + 1]; x[3 * br + 2] += alpha * p[3 * br + 2];
        r0 -= alpha * Ap[3 * br]; r1 -= alpha * Ap[3 * br + 1]; r2 -= alpha * Ap[3 * br + 2];
        r[3 * br] = r0; r[3 * br + 1] = r1; r[3 * br + 2] = r2;
        double z0, z1, z2;
        apply_dinv(dinv + 9 * (size_t)br, r0, r1, r2, z0, z1, z2);
        z[3 * br] = z0; z[3 * br + 1] = z1; z[3 * br + 2] = z2;
        rr += r0 * r0 + r1 * r1 + r2 * r2;
        rz += r0 * z0 + r1 * z1 + r2 * z2;
    }
    const double t0 = block_sum(rr, s);
    if (threadIdx.x == 0) part[PCG_BLOCKS + blockIdx.x] = t0;
    const double t1 = block_sum(rz, s);
    if (threadIdx.x == 0) part[2 * PCG_BLOCKS + blockIdx.x] = t1;
}

// convergence tests, beta = rz_new / rz_old, p = z + beta p.  CTA 0 publishes the new state AFTER every CTA has read
// the old one: the state update is deferred to a tiny follow-up kernel so there is no intra-kernel race.
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_direction(const PcgState* __restrict__ st, const double* __restrict__ z, double* __restrict__ p,
                                                                 const double* __restrict__ part, int nbr)
{
    __shared__ double s[PCG_THREADS / 32];
    if (st->done) return;
    const double pAp = all_partials(part, s);
    if (pAp <= 0.0 && st->stop_on_indef) return;
    const double rr = all_partials(part + PCG_BLOCKS, s);
    const double error = sqrt(rr / st->bb);
    if (error < st->abs_tol || error / st->error0 < st->rel_tol) return;   // converged: p is not needed any more
    const double rz_new = all_partials(part + 2 * PCG_BLOCKS, s);
    const double beta = rz_new / st->rz;
    for (int i = blockIdx.x * PCG_THREADS + threadIdx.x; i < 3 * nbr; i += PCG_BLOCKS * PCG_THREADS) p[i] = z[i] + beta * p[i];
}
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_advance(PcgState* st, const double* __restrict__ part)
{
    __shared__ double s[PCG_THREADS / 32];
    if (st->done) return;
    const double pAp = all_partials(part, s);
    const double rr = all_partials(part + PCG_BLOCKS, s);
    const double rz_new = all_partials(part + 2 * PCG_BLOCKS, s);
    if (threadIdx.x != 0) return;
    const int it = st->it + 1;
    st->it = it;
    if (pAp <= 0.0) {
        st->found_indef = 1;
        if (st->stop_on_indef) { st->done = 2; return; }
    }
    const double error = sqrt(rr / st->bb);
    st->error = error;
    if (error < st->abs_tol || error / st->error0 < st->rel_tol) { st->done = 1; return; }
    st->rz = rz_new;
    if (it >= st->max_iter) st->done = 3;
}

// du = x ; partials of du.grad and |du|_inf
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_finish(const double* __restrict__ x, const double* __restrict__ grad, double* __restrict__ du,
                                                              double* __restrict__ part, int n)
{
    __shared__ double s[PCG_THREADS / 32];
    double dg = 0.0, mx = 0.0;
    for (int i = blockIdx.x * PCG_THREADS + threadIdx.x; i < n; i += PCG_BLOCKS * PCG_THREADS) {
        const double v = x[i];
        du[i] = v;
        dg += v * grad[i];
        mx = fmax(mx, fabs(v));
    }
    const double t = block_sum(dg, s);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
    // max via the same tree (values are non-negative)
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < PCG_THREADS / 32; w++) m = fmax(m, s[w]);
        part[PCG_BLOCKS + blockIdx.x] = m;
    }
}
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_finish2(const double* __restrict__ part, double* __restrict__ out)
{
    __shared__ double s[PCG_THREADS / 32];
    const double dg = all_partials(part, s);
    double mx = 0.0;
    for (int i = threadIdx.x; i < PCG_BLOCKS; i += PCG_THREADS) mx = fmax(mx, part[PCG_BLOCKS + i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < PCG_THREADS / 32; w++) m = fmax(m, s[w]);
        out[0] = dg;
        out[1] = m;
    }
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
    P->r.ensure(n); P->z.ensure(n); P->p.ensure(n); P->Ap.ensure(n); P->x.ensure(n); P->dinv.ensure(9 * (size_t)nbr);
    P->part.ensure(3 * PCG_BLOCKS);
    ctx->du.ensure(n);

    k_block_jacobi<<<(nbr + 255) / 256, 256, 0, st>>>(rows, cols, vals, P->dinv.p, nbr);
    k_pcg_init<<<PCG_BLOCKS, PCG_THREADS, 0, st>>>(ctx->grad.p, P->dinv.p, P->x.p, P->r.p, P->z.p, P->p.p, P->part.p, nbr);
    k_pcg_init_state<<<1, PCG_THREADS, 0, st>>>(P->d_state, P->part.p, abs_tol, rel_tol, max_iter, stop_on_indef);
    ctx->launches += 3;
    int launched = 0;
    while (true) {
        for (int b = 0; b < PCG_BATCH && launched < max_iter; b++, launched++) {
            k_pcg_spmv<<<PCG_BLOCKS, PCG_THREADS, 0, st>>>(P->d_state, rows, cols, vals, P->p.p, P->Ap.p, P->part.p, nbr);
            k_pcg_update<<<PCG_BLOCKS, PCG_THREADS, 0, st>>>(P->d_state, P->dinv.p, P->p.p, P->Ap.p, P->x.p, P->r.p, P->z.p, P->part.p, nbr);
            k_pcg_direction<<<PCG_BLOCKS, PCG_THREADS, 0, st>>>(P->d_state, P->z.p, P->p.p, P->part.p, nbr);
            k_pcg_advance<<<1, PCG_THREADS, 0, st>>>(P->d_state, P->part.p);
            ctx->launches += 4;
        }
        SB_CUDA(ctx, cudaMemcpyAsync(P->h_state, P->d_state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
        SB_CUDA(ctx, cudaStreamSynchronize(st));
        if (P->h_state->done || launched >= max_iter) break;
    }
    k_pcg_finish<<<PCG_BLOCKS, PCG_THREADS, 0, st>>>(P->x.p, ctx->grad.p, ctx->du.p, P->part.p, n);
    k_pcg_finish2<<<1, PCG_THREADS, 0, st>>>(P->part.p, ctx->d_scalars + 2);
    ctx->launches += 2;
    SB_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 2, ctx->d_scalars + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    SB_CUDA(ctx, cudaGetLastError());
    if (out_iterations) *out_iterations = P->h_state->it;
    if (out_ok) *out_ok = (P->h_state->done == 1) ? 1 : 0;
    if (out_du_dot_grad) *out_du_dot_grad = ctx->h_scalars[2];
    if (out_du_inf) *out_du_inf = ctx->h_scalars[3];
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
