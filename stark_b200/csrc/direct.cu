// Direct solve of the Newton system: dense blocked Cholesky (L L^T) in FP64.
//
// Replaces the DirectLLT branch of NewtonsMethod::_solve_linear_system (symx/solver/NewtonsMethod.cpp:395-418):
// BlockedSparseMatrix::to_triplets (bsm/BlockedSparseMatrix.h:1365-1393) -> Eigen::SimplicialLLT (serial, re-analysed at
// every call) -> solve; "factorisation failed" (matrix not positive definite) is reported as a failed solve, which makes
// the Newton driver project more Hessians, exactly like `solver.info() != Eigen::Success`.
//
// The reference uses this path for its unit tests and small scenes (at 27 k DoFs it already needs 6.3 s per solve,
// BASELINE.md section 2).  Here the float-stored BCSR matrix is expanded to a dense FP64 lower triangle and factorised
// right-looking in 64 x 64 tiles: POTRF of the diagonal tile (one CTA), TRSM of the panel below it (one CTA per tile),
// SYRK / GEMM update of the trailing tiles (one CTA per tile, 4 x 4 register blocking) -- n^3 / 3 FP64 flops on the
// FP64 pipe (FP64 has no tcgen05 path).  Dense storage bounds the size: n <= 32,768 DoFs
// (8.6 GB); larger systems are the block-Jacobi PCG's domain and are rejected with an error, never silently re-routed.
#include "internal.h"
#include <algorithm>

namespace sb {

int bcsr_view(sb_context* ctx, int* nbr, size_t* nnzb, const unsigned long long** rows, const int32_t** cols, const float** vals);

constexpr int NB = 64;                 // tile size
constexpr int LLT_MAX_N = 32768;

struct Direct {
    DevBuf<double> A;                  // [np x np] row-major, lower triangle used
    DevBuf<double> y;                  // [np] right-hand side / solution
    int* d_fail = nullptr;
    int* h_fail = nullptr;
    double* h_out = nullptr;           // du.grad, |du|_inf
};
void direct_destroy(sb_context* ctx)
{
    Direct* D = ctx->direct;
    if (!D) return;
    D->A.release(); D->y.release();
    if (D->d_fail) cudaFree(D->d_fail);
    if (D->h_fail) cudaFreeHost(D->h_fail);
    if (D->h_out) cudaFreeHost(D->h_out);
    delete D;
    ctx->direct = nullptr;
}

// A = dense(BCSR) on the lower triangle (i >= j); padding rows get a unit diagonal
__global__ void k_dense_from_bcsr(const unsigned long long* __restrict__ rows, const int32_t* __restrict__ cols, const float* __restrict__ vals,
                                  double* __restrict__ A, int nbr, int np)
{
    const int br = blockIdx.x;
    if (br >= nbr) return;
    const int g = threadIdx.x / 9, k = threadIdx.x % 9;   // 32 groups of nine threads, one BCSR block per group and trip
    const int r = k % 3, c = k / 3;                        // column-major inside the block
    for (unsigned long long j = rows[br] + g; j < rows[br + 1]; j += 32) {
        const int gi = 3 * br + r, gj = cols[j] + c;
        if (gi >= gj) A[(size_t)gi * np + gj] = (double)vals[9 * j + k];
    }
}
__global__ void k_pad_diagonal(double* __restrict__ A, int n, int np)
{
    const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) A[(size_t)i * np + i] = 1.0;
}

// Cholesky of the diagonal tile k (lower, in place); *fail = 1 on a non-positive pivot
__global__ void __launch_bounds__(256) k_potrf_tile(double* __restrict__ A, int np, int k, int* __restrict__ fail)
{
    __shared__ double T[NB][NB + 1];
    double* base = A + (size_t)(k * NB) * np + k * NB;
    for (int t = threadIdx.x; t < NB * NB; t += 256) { const int i = t / NB, j = t % NB; T[i][j] = (j <= i) ? base[(size_t)i * np + j] : 0.0; }
    __syncthreads();
    for (int j = 0; j < NB; j++) {
        const double d = T[j][j];
        if (!(d > 0.0)) { if (threadIdx.x == 0) *fail = 1; return; }   // shared value: uniform exit
        const double l = sqrt(d);
        __syncthreads();
        if (threadIdx.x == 0) T[j][j] = l;
        for (int i = j + 1 + threadIdx.x; i < NB; i += 256) T[i][j] /= l;
        __syncthreads();
        // trailing update: T[i][c] -= T[i][j] * T[c][j] for j < c <= i
        for (int t = threadIdx.x; t < NB * NB; t += 256) {
            const int i = t / NB, c = t % NB;
            if (c > j && c <= i) T[i][c] -= T[i][j] * T[c][j];
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < NB * NB; t += 256) { const int i = t / NB, j = t % NB; if (j <= i) base[(size_t)i * np + j] = T[i][j]; }
}

// panel: A[i][k] <- A[i][k] L_kk^-T for every tile row i > k (one CTA per tile, one thread per tile row)
__global__ void __launch_bounds__(NB) k_trsm_panel(double* __restrict__ A, int np, int k, const int* __restrict__ fail)
{
    if (*fail) return;
    __shared__ double L[NB][NB + 1];
    const int it = k + 1 + blockIdx.x;
    const double* Lkk = A + (size_t)(k * NB) * np + k * NB;
    for (int t = threadIdx.x; t < NB * NB; t += NB) { const int i = t / NB, j = t % NB; L[i][j] = Lkk[(size_t)i * np + j]; }
    __syncthreads();
    double* row = A + (size_t)(it * NB + threadIdx.x) * np + k * NB;
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) x[j] = row[j];
#pragma unroll
    for (int j = 0; j < NB; j++) {
        double acc = x[j];
#pragma unroll
        for (int m = 0; m < j; m++) acc -= x[m] * L[j][m];
        x[j] = acc / L[j][j];
    }
#pragma unroll
    for (int j = 0; j < NB; j++) row[j] = x[j];
}

// trailing update: A[i][j] -= A[i][k] A[j][k]^T for k < j <= i (one CTA per tile; 16 x 16 threads, 4 x 4 outputs each)
__global__ void __launch_bounds__(256) k_syrk_update(double* __restrict__ A, int np, int k, int nt, const int* __restrict__ fail)
{
    if (*fail) return;
    // linear tile index -> (i, j) in the lower triangle of the trailing (nt - k - 1)^2 tile matrix
    const int m = nt - k - 1;
    int idx = blockIdx.x;
    int ti = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= idx) ti++;
    while (ti * (ti + 1) / 2 > idx) ti--;
    const int tj = idx - ti * (ti + 1) / 2;
    if (ti >= m) return;
    const int I = k + 1 + ti, J = k + 1 + tj;
    constexpr int KH = NB / 2;   // the two 64 x 64 panels are streamed through shared memory in two K-halves (2 x 16.5 KB)
    __shared__ double Pa[NB][KH + 1], Pb[NB][KH + 1];
    const double* pa = A + (size_t)(I * NB) * np + k * NB;
    const double* pb = A + (size_t)(J * NB) * np + k * NB;
    const int tr = (threadIdx.x / 16) * 4, tc = (threadIdx.x % 16) * 4;
    double acc[4][4] = {{0}};
    for (int half = 0; half < 2; half++) {
        __syncthreads();
        for (int t = threadIdx.x; t < NB * KH; t += 256) {
            const int r = t / KH, c = t % KH;
            Pa[r][c] = pa[(size_t)r * np + half * KH + c];
            Pb[r][c] = pb[(size_t)r * np + half * KH + c];
        }
        __syncthreads();
        for (int q = 0; q < KH; q++) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { a[u] = Pa[tr + u][q]; b[u] = Pb[tc + u][q]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] += a[u] * b[v];
        }
    }
    double* out = A + (size_t)(I * NB) * np + J * NB;
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++)
            if (I != J || tc + v <= tr + u) out[(size_t)(tr + u) * np + tc + v] -= acc[u][v];
}

// ---- triangular solves with one right-hand side ----
// y_k <- L_kk^-1 y_k (forward) or L_kk^-T y_k (backward), one CTA
__global__ void __launch_bounds__(NB) k_solve_diag(const double* __restrict__ A, double* __restrict__ y, int np, int k, int transposed)
{
    __shared__ double L[NB][NB + 1];
    __shared__ double v[NB];
    const double* Lkk = A + (size_t)(k * NB) * np + k * NB;
    for (int t = threadIdx.x; t < NB * NB; t += NB) { const int i = t / NB, j = t % NB; L[i][j] = Lkk[(size_t)i * np + j]; }
    v[threadIdx.x] = y[k * NB + threadIdx.x];
    __syncthreads();
    if (!transposed) {
        for (int j = 0; j < NB; j++) {
            if (threadIdx.x == j) v[j] /= L[j][j];
            __syncthreads();
            if (threadIdx.x > j) v[threadIdx.x] -= L[threadIdx.x][j] * v[j];
            __syncthreads();
        }
    } else {
        for (int j = NB - 1; j >= 0; j--) {
            if (threadIdx.x == j) v[j] /= L[j][j];
            __syncthreads();
            if (threadIdx.x < j) v[threadIdx.x] -= L[j][threadIdx.x] * v[j];
            __syncthreads();
        }
    }
    y[k * NB + threadIdx.x] = v[threadIdx.x];
}
// forward: y_i -= A[i][k] y_k for tile rows i > k ; backward: y_i -= A[k][i]^T y_k for tile rows i < k
__global__ void __launch_bounds__(NB) k_solve_update(const double* __restrict__ A, double* __restrict__ y, int np, int k, int transposed)
{
    __shared__ double yk[NB];
    yk[threadIdx.x] = y[k * NB + threadIdx.x];
    __syncthreads();
    double acc = 0.0;
    if (!transposed) {
        const int it = k + 1 + blockIdx.x;
        const double* row = A + (size_t)(it * NB + threadIdx.x) * np + k * NB;
        for (int j = 0; j < NB; j++) acc += row[j] * yk[j];
        y[it * NB + threadIdx.x] -= acc;
    } else {
        const int it = blockIdx.x;   // < k
        const double* col = A + (size_t)(k * NB) * np + it * NB + threadIdx.x;   // A[k*NB + j][it*NB + t]
        for (int j = 0; j < NB; j++) acc += col[(size_t)j * np] * yk[j];
        y[it * NB + threadIdx.x] -= acc;
    }
}

__global__ void k_rhs(const double* __restrict__ grad, double* __restrict__ y, int n, int np)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) y[i] = (i < n) ? -grad[i] : 0.0;
}
__global__ void k_copy_n(double* __restrict__ dst, const double* __restrict__ src, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// out[0] = du.grad, out[1] = |du|_inf (single CTA, fixed tree: the systems this path handles are small)
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

int solve_llt_internal(sb_context* ctx, int* out_ok, double* out_du_dot_grad, double* out_du_inf)
{
    int nbr; size_t nnzb; const unsigned long long* rows; const int32_t* cols; const float* vals;
    int r = bcsr_view(ctx, &nbr, &nnzb, &rows, &cols, &vals);
    if (r) return r;
    const int n = 3 * nbr;
    if (n != ctx->ndofs) return fail(ctx, SB_ERR_STATE, "sb_solve_llt: matrix and DoF vector sizes differ");
    if (n > LLT_MAX_N) return fail(ctx, SB_ERR_STATE, "sb_solve_llt: the direct solver is dense (n <= 32768 DoFs); use the block-Jacobi PCG for larger systems");
    StageTimer timer(ctx, ST_PCG);
    if (!ctx->direct) {
        ctx->direct = new Direct();
        cudaMalloc(&ctx->direct->d_fail, sizeof(int));
        cudaMallocHost(&ctx->direct->h_fail, sizeof(int));
        cudaMallocHost(&ctx->direct->h_out, 2 * sizeof(double));
    }
    Direct& D = *ctx->direct;
    cudaStream_t st = ctx->stream;
    const int nt = (n + NB - 1) / NB, np = nt * NB;
    D.A.ensure((size_t)np * np);
    D.y.ensure(np);
    ctx->du.ensure(n);
    SB_CUDA(ctx, cudaMemsetAsync(D.A.p, 0, sizeof(double) * (size_t)np * np, st));
    SB_CUDA(ctx, cudaMemsetAsync(D.d_fail, 0, sizeof(int), st));
    k_dense_from_bcsr<<<nbr, 288, 0, st>>>(rows, cols, vals, D.A.p, nbr, np);
    if (np > n) k_pad_diagonal<<<(np - n + 63) / 64, 64, 0, st>>>(D.A.p, n, np);
    ctx->launches += 2;
    for (int k = 0; k < nt; k++) {
        k_potrf_tile<<<1, 256, 0, st>>>(D.A.p, np, k, D.d_fail);
        ctx->launches++;
        const int m = nt - k - 1;
        if (m > 0) {
            k_trsm_panel<<<m, NB, 0, st>>>(D.A.p, np, k, D.d_fail);
            k_syrk_update<<<m * (m + 1) / 2, 256, 0, st>>>(D.A.p, np, k, nt, D.d_fail);
            ctx->launches += 2;
        }
    }
    SB_CUDA(ctx, cudaMemcpyAsync(D.h_fail, D.d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    SB_CUDA(ctx, cudaGetLastError());
    if (*D.h_fail) {   // not positive definite: Eigen::SimplicialLLT::info() != Success
        if (out_ok) *out_ok = 0;
        if (out_du_dot_grad) *out_du_dot_grad = 0.0;
        if (out_du_inf) *out_du_inf = 0.0;
        return 0;
    }
    k_rhs<<<(np + 255) / 256, 256, 0, st>>>(ctx->grad.p, D.y.p, n, np);
    for (int k = 0; k < nt; k++) {
        k_solve_diag<<<1, NB, 0, st>>>(D.A.p, D.y.p, np, k, 0);
        if (nt - k - 1 > 0) k_solve_update<<<nt - k - 1, NB, 0, st>>>(D.A.p, D.y.p, np, k, 0);
    }
    for (int k = nt - 1; k >= 0; k--) {
        k_solve_diag<<<1, NB, 0, st>>>(D.A.p, D.y.p, np, k, 1);
        if (k > 0) k_solve_update<<<k, NB, 0, st>>>(D.A.p, D.y.p, np, k, 1);
    }
    k_copy_n<<<(n + 255) / 256, 256, 0, st>>>(ctx->du.p, D.y.p, n);
    k_du_stats<<<1, 1024, 0, st>>>(ctx->du.p, ctx->grad.p, n, ctx->d_scalars + 4);
    ctx->launches += 4 * nt + 3;
    SB_CUDA(ctx, cudaMemcpyAsync(D.h_out, ctx->d_scalars + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    SB_CUDA(ctx, cudaGetLastError());
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
