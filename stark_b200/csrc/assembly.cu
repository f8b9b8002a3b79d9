// Global Hessian assembly: element Hessians (COO of 3x3 blocks) -> 3x3-block CSR, float storage.
//
// Replaces ElementHessians::{assemble_global, update_global} (symx/solver/second_order/ElementHessians.cpp:224-294)
// and BlockedSparseMatrix::{start_insertion, add_block, end_insertion} (bsm/BlockedSparseMatrix.h:332-593, 782-895).
// The reference inserts every block under a per-row mutex after a binary search; here assembly is split into
//   * a SYMBOLIC phase, run only when some connectivity changed: one 64-bit key (block_row, block_col) per element
//     block, radix sort, run-length heads -> BCSR pattern + for every BCSR block the contiguous list of its sources;
//   * a NUMERIC phase: a segmented reduction -- nine threads per BCSR block sum their sources in FP64 and store float
//     (the reference accumulates in float in thread order, BlockedSparseMatrix.h:899-952).
// Output layout is the reference's (BlockedSparseMatrix.h:266-271): rows = offsets per block row, cols = first scalar
// column of the block, vals = 9 floats per block, column-major inside the block.
#include "internal.h"
#include <cub/cub.cuh>

namespace sb {

struct PotDesc {
    unsigned long long H_off;     // offset of the potential's element Hessians in ctx->H
    unsigned long long rows_off;  // offset of its block rows in ctx->rows
    unsigned long long blk_off;   // offset of its element blocks in the source numbering
    int n_elem, nb;
};

struct Assembly {
    uint64_t built_version = 0;
    size_t built_n_src = 0;
    size_t n_src = 0;
    int nbr = 0;
    size_t nnzb = 0;
    DevBuf<uint64_t> keys, keys_sorted;
    DevBuf<uint32_t> ids, ids_sorted;
    DevBuf<uint32_t> src_off;        // per source (unsorted numbering): offset of its (0,0) entry in ctx->H
    DevBuf<uint8_t> src_pitch;       // per source: row pitch n of its element Hessian
    DevBuf<uint32_t> sorted_off;     // per sorted source
    DevBuf<uint8_t> sorted_pitch;
    DevBuf<uint32_t> head, blk_of;   // head flags / BCSR block of every sorted source
    DevBuf<uint32_t> seg;            // [nnzb + 1] first sorted source of every BCSR block
    DevBuf<int32_t> blk_row;         // [nnzb] block row of every BCSR block
    DevBuf<unsigned long long> rows; // [nbr + 1]
    DevBuf<int32_t> cols;            // [nnzb] scalar column
    DevBuf<float> vals;              // [9 nnzb]
    DevBuf<uint8_t> temp;            // cub scratch
    DevBuf<uint32_t> blk_of_src;     // per source (unsorted numbering): its BCSR block
    DevBuf<uint8_t> dirty;           // per BCSR block: a source changed since the last numeric pass (PD projection)
    DevBuf<uint32_t> long_blocks;    // BCSR blocks with more than LONG_SEG sources (rigid bodies touched by many contacts)
    int* d_n_long = nullptr;
    uint64_t assembled_eval = 0;     // evaluation the values belong to
    bool numeric_valid = false;
};

static Assembly* get(sb_context* ctx)
{
    if (!ctx->assembly) ctx->assembly = new Assembly();
    return ctx->assembly;
}
void assembly_destroy(sb_context* ctx)
{
    Assembly* A = ctx->assembly;
    if (!A) return;
    A->keys.release(); A->keys_sorted.release(); A->ids.release(); A->ids_sorted.release(); A->src_off.release();
    A->src_pitch.release(); A->sorted_off.release(); A->sorted_pitch.release(); A->head.release(); A->blk_of.release();
    A->seg.release(); A->blk_row.release(); A->rows.release(); A->cols.release(); A->vals.release(); A->temp.release(); A->blk_of_src.release(); A->dirty.release(); A->long_blocks.release();
    if (A->d_n_long) cudaFree(A->d_n_long);
    delete A;
    ctx->assembly = nullptr;
}

__global__ void k_make_keys(PotDesc d, const int32_t* __restrict__ rows_all, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids,
                            uint32_t* __restrict__ src_off, uint8_t* __restrict__ src_pitch)
{
    const int nb2 = d.nb * d.nb;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)d.n_elem * nb2) return;
    const int e = (int)(t / nb2);
    const int k = (int)(t - (size_t)e * nb2);
    const int bi = k / d.nb, bj = k - bi * d.nb;
    const int32_t* r = rows_all + d.rows_off + (size_t)e * d.nb;
    const size_t id = d.blk_off + t;
    const int n = 3 * d.nb;
    keys[id] = ((uint64_t)(uint32_t)r[bi] << 32) | (uint32_t)r[bj];
    ids[id] = (uint32_t)id;
    src_off[id] = (uint32_t)(d.H_off + (size_t)e * n * n + (size_t)(3 * bi) * n + 3 * bj);
    src_pitch[id] = (uint8_t)n;
}

__global__ void k_heads(const uint64_t* __restrict__ keys, uint32_t* __restrict__ head, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// blk_of = inclusive scan of head (1-based); fill the pattern at heads, gather the per-source offsets in sorted order
__global__ void k_fill_pattern(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ head, const uint32_t* __restrict__ blk_of,
                               const uint32_t* __restrict__ ids_sorted, const uint32_t* __restrict__ src_off, const uint8_t* __restrict__ src_pitch,
                               uint32_t* __restrict__ sorted_off, uint8_t* __restrict__ sorted_pitch,
                               uint32_t* __restrict__ seg, int32_t* __restrict__ blk_row, int32_t* __restrict__ cols, uint32_t* __restrict__ blk_of_src, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids_sorted[i];
    sorted_off[i] = src_off[id];
    sorted_pitch[i] = src_pitch[id];
    blk_of_src[id] = blk_of[i] - 1;
    if (head[i]) {
        const uint32_t b = blk_of[i] - 1;
        const uint64_t k = keys[i];
        seg[b] = (uint32_t)i;
        blk_row[b] = (int32_t)(k >> 32);
        cols[b] = 3 * (int32_t)(k & 0xffffffffu);
    }
}

// rows[r] = first BCSR block whose block row is >= r (blocks are sorted by (row, col))
__global__ void k_row_ptr(const int32_t* __restrict__ blk_row, unsigned long long* __restrict__ rows, int nbr, size_t nnzb)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nbr) return;
    size_t lo = 0, hi = nnzb;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (blk_row[mid] < r) lo = mid + 1; else hi = mid;
    }
    rows[r] = lo;
}

// Segmented reduction: 9 consecutive threads own one BCSR block (thread k -> entry (r = k % 3, c = k / 3), column-major).
// Sources are summed in their sorted order (deterministic), four loads in flight per thread.  ONLY_DIRTY: re-sum just the
// blocks a PD projection touched (the reference's update_global, ElementHessians.cpp:262-294, adds projected - original).
// a long block is left to k_assemble_long only if it made it into the (capped) list
__device__ __forceinline__ bool long_rank_ok(size_t b, const uint32_t* __restrict__ long_blocks, const int* __restrict__ n_long)
{
    const int n = *n_long;
    if (n <= 4096) return true;          // nothing was dropped
    for (int i = 0; i < 4096; i++) if (long_blocks[i] == (uint32_t)b) return true;
    return false;
}
constexpr int LONG_SEG = 64;
constexpr int LONG_CAP = 4096;      // long blocks beyond this many are summed by the plain path
constexpr int LONG_CTAS = 64;

__global__ void k_find_long(const uint32_t* __restrict__ seg, uint32_t* __restrict__ long_blocks, int* __restrict__ n_long, size_t nnzb)
{
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nnzb) return;
    if (seg[b + 1] - seg[b] > LONG_SEG) {
        const int k = atomicAdd(n_long, 1);
        if (k < LONG_CAP) long_blocks[k] = (uint32_t)b;
    }
}

// One CTA per long block: 32 strided partial sums per entry, then a fixed shared-memory tree (deterministic).
template<bool ONLY_DIRTY>
__global__ void __launch_bounds__(288) k_assemble_long(const double* __restrict__ H, const uint32_t* __restrict__ seg,
                                                        const uint32_t* __restrict__ sorted_off, const uint8_t* __restrict__ sorted_pitch,
                                                        float* __restrict__ vals, const uint8_t* __restrict__ dirty,
                                                        const uint32_t* __restrict__ long_blocks, const int* __restrict__ n_long)
{
    __shared__ double sm[32][9];
    const int total = min(*n_long, LONG_CAP);
    const int k = threadIdx.x % 9, lane = threadIdx.x / 9;
    const int r = k % 3, c = k / 3;
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
        const uint32_t b = long_blocks[i];
        if (ONLY_DIRTY && !dirty[b]) continue;   // uniform across the CTA
        const uint32_t s0 = seg[b], s1 = seg[b + 1];
        double acc = 0.0;
        for (uint32_t s = s0 + lane; s < s1; s += 32) acc += H[(size_t)sorted_off[s] + (size_t)r * sorted_pitch[s] + c];
        sm[lane][k] = acc;
        __syncthreads();
        for (int w = 16; w > 0; w >>= 1) {
            if (lane < w) sm[lane][k] += sm[lane + w][k];
            __syncthreads();
        }
        if (lane == 0) vals[9 * (size_t)b + k] = (float)sm[0][k];
        __syncthreads();
    }
}

template<bool ONLY_DIRTY>
__global__ void __launch_bounds__(288) k_assemble_numeric(const double* __restrict__ H, const uint32_t* __restrict__ seg,
                                                           const uint32_t* __restrict__ sorted_off, const uint8_t* __restrict__ sorted_pitch,
                                                           float* __restrict__ vals, uint8_t* __restrict__ dirty, size_t nnzb,
                                                           const uint32_t* __restrict__ long_blocks, const int* __restrict__ n_long)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t b = t / 9;
    if (b >= nnzb) return;
    if (ONLY_DIRTY && !dirty[b]) return;
    const int k = (int)(t - b * 9);
    const int r = k % 3, c = k / 3;
    const uint32_t s0 = seg[b], s1 = seg[b + 1];
    if (s1 - s0 > LONG_SEG && long_rank_ok(b, long_blocks, n_long)) return;   // summed by k_assemble_long
    double acc = 0.0;
    uint32_t s = s0;
    for (; s + 4 <= s1; s += 4) {
        const size_t o0 = (size_t)sorted_off[s] + (size_t)r * sorted_pitch[s] + c;
        const size_t o1 = (size_t)sorted_off[s + 1] + (size_t)r * sorted_pitch[s + 1] + c;
        const size_t o2 = (size_t)sorted_off[s + 2] + (size_t)r * sorted_pitch[s + 2] + c;
        const size_t o3 = (size_t)sorted_off[s + 3] + (size_t)r * sorted_pitch[s + 3] + c;
        const double v0 = H[o0], v1 = H[o1], v2 = H[o2], v3 = H[o3];
        acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; s < s1; s++) acc += H[(size_t)sorted_off[s] + (size_t)r * sorted_pitch[s] + c];
    vals[t] = (float)acc;
}
// the dirty flags are cleared by a separate pass (the nine threads of a block must all have seen the flag)
__global__ void k_clear_dirty(uint8_t* __restrict__ dirty, size_t nnzb)
{
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nnzb) dirty[b] = 0;
}

static int build_symbolic(sb_context* ctx, Assembly* A)
{
    StageTimer timer(ctx, ST_ASM_SYMBOLIC);
    cudaStream_t st = ctx->stream;
    const size_t n = ctx->n_blocks_total;
    if (ctx->H_total >= (1ull << 32)) return fail(ctx, SB_ERR_STATE, "sb_assemble: element Hessian storage exceeds 32-bit offsets");
    A->n_src = n;
    A->nbr = ctx->ndofs / 3;
    A->keys.ensure(n + 1); A->keys_sorted.ensure(n + 1); A->ids.ensure(n + 1); A->ids_sorted.ensure(n + 1);
    A->src_off.ensure(n + 1); A->src_pitch.ensure(n + 1); A->sorted_off.ensure(n + 1); A->sorted_pitch.ensure(n + 1);
    A->head.ensure(n + 1); A->blk_of.ensure(n + 1); A->blk_of_src.ensure(n + 1);

    size_t blk_off = 0;
    for (auto& p : ctx->potentials) {
        if (p.n_elem == 0) continue;
        PotDesc d;
        d.H_off = p.H_off; d.rows_off = p.rows_off; d.blk_off = blk_off; d.n_elem = p.n_elem; d.nb = p.k->nb;
        const size_t cnt = (size_t)p.n_elem * d.nb * d.nb;
        k_make_keys<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d, ctx->rows.p, A->keys.p, A->ids.p, A->src_off.p, A->src_pitch.p);
        ctx->launches++;
        blk_off += cnt;
    }
    int row_bits = 1;
    while ((1ll << row_bits) < A->nbr + 1) row_bits++;
    size_t temp_bytes = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, A->keys.p, A->keys_sorted.p, A->ids.p, A->ids_sorted.p, (int)n, 0, 32 + row_bits, st);
    cub::DeviceScan::InclusiveSum(nullptr, tb2, A->head.p, A->blk_of.p, (int)n, st);
    A->temp.ensure(std::max(temp_bytes, tb2) + 16);
    temp_bytes = A->temp.cap;
    SB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(A->temp.p, temp_bytes, A->keys.p, A->keys_sorted.p, A->ids.p, A->ids_sorted.p, (int)n, 0, 32 + row_bits, st));
    k_heads<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A->keys_sorted.p, A->head.p, n);
    temp_bytes = A->temp.cap;
    SB_CUDA(ctx, cub::DeviceScan::InclusiveSum(A->temp.p, temp_bytes, A->head.p, A->blk_of.p, (int)n, st));
    ctx->launches += 6;
    uint32_t nnzb32 = 0;
    SB_CUDA(ctx, cudaMemcpyAsync(&nnzb32, A->blk_of.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    A->nnzb = nnzb32;
    A->seg.ensure(A->nnzb + 2); A->blk_row.ensure(A->nnzb + 1); A->cols.ensure(A->nnzb + 1); A->vals.ensure(9 * A->nnzb + 9);
    A->rows.ensure(A->nbr + 2);
    A->dirty.ensure(A->nnzb + 1);
    SB_CUDA(ctx, cudaMemsetAsync(A->dirty.p, 0, A->nnzb + 1, st));
    k_fill_pattern<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A->keys_sorted.p, A->head.p, A->blk_of.p, A->ids_sorted.p, A->src_off.p, A->src_pitch.p,
                                                                 A->sorted_off.p, A->sorted_pitch.p, A->seg.p, A->blk_row.p, A->cols.p, A->blk_of_src.p, n);
    const uint32_t n32 = (uint32_t)n;
    SB_CUDA(ctx, cudaMemcpyAsync(A->seg.p + A->nnzb, &n32, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    k_row_ptr<<<(A->nbr + 1 + 255) / 256, 256, 0, st>>>(A->blk_row.p, A->rows.p, A->nbr, A->nnzb);
    if (!A->d_n_long) SB_CUDA(ctx, cudaMalloc(&A->d_n_long, sizeof(int)));
    A->long_blocks.ensure(LONG_CAP);
    SB_CUDA(ctx, cudaMemsetAsync(A->d_n_long, 0, sizeof(int), st));
    k_find_long<<<(unsigned)((A->nnzb + 255) / 256), 256, 0, st>>>(A->seg.p, A->long_blocks.p, A->d_n_long, A->nnzb);
    ctx->launches += 3;
    SB_CUDA(ctx, cudaStreamSynchronize(st));   // n32 is a stack variable
    SB_CUDA(ctx, cudaGetLastError());
    A->built_version = ctx->pattern_version;
    A->built_n_src = n;
    return 0;
}

int assemble_internal(sb_context* ctx)
{
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_assemble: call sb_eval(SB_EVAL_PGH) first");
    Assembly* A = get(ctx);
    if (ctx->n_blocks_total == 0) return fail(ctx, SB_ERR_STATE, "sb_assemble: no element Hessians");
    bool rebuilt = false;
    if (A->built_version != ctx->pattern_version || A->built_n_src != ctx->n_blocks_total || A->nbr != ctx->ndofs / 3) {
        int r = build_symbolic(ctx, A);
        if (r) return r;
        rebuilt = true;
    }
    StageTimer timer(ctx, ST_ASM_NUMERIC);
    const size_t nt = 9 * A->nnzb;
    const unsigned grid = (unsigned)((nt + 287) / 288);
    if (!rebuilt && A->numeric_valid && A->assembled_eval == ctx->eval_id) {
        // same evaluation, same pattern: only PD-projected elements changed since the last pass
        k_assemble_numeric<true><<<grid, 288, 0, ctx->stream>>>(ctx->H.p, A->seg.p, A->sorted_off.p, A->sorted_pitch.p, A->vals.p, A->dirty.p, A->nnzb, A->long_blocks.p, A->d_n_long);
        k_assemble_long<true><<<LONG_CTAS, 288, 0, ctx->stream>>>(ctx->H.p, A->seg.p, A->sorted_off.p, A->sorted_pitch.p, A->vals.p, A->dirty.p, A->long_blocks.p, A->d_n_long);
    } else {
        k_assemble_numeric<false><<<grid, 288, 0, ctx->stream>>>(ctx->H.p, A->seg.p, A->sorted_off.p, A->sorted_pitch.p, A->vals.p, A->dirty.p, A->nnzb, A->long_blocks.p, A->d_n_long);
        k_assemble_long<false><<<LONG_CTAS, 288, 0, ctx->stream>>>(ctx->H.p, A->seg.p, A->sorted_off.p, A->sorted_pitch.p, A->vals.p, A->dirty.p, A->long_blocks.p, A->d_n_long);
    }
    k_clear_dirty<<<(unsigned)((A->nnzb + 255) / 256), 256, 0, ctx->stream>>>(A->dirty.p, A->nnzb);
    ctx->launches += 3;
    SB_CUDA(ctx, cudaGetLastError());
    A->assembled_eval = ctx->eval_id;
    A->numeric_valid = true;
    return 0;
}

// for project.cu: where a changed element's blocks land (null when the pattern is stale -> the next assembly is a full one anyway)
bool assembly_dirty_view(sb_context* ctx, const uint32_t** blk_of_src, uint8_t** dirty)
{
    Assembly* A = ctx->assembly;
    if (!A || A->built_version != ctx->pattern_version || A->built_n_src != ctx->n_blocks_total || A->nbr != ctx->ndofs / 3) return false;
    *blk_of_src = A->blk_of_src.p; *dirty = A->dirty.p;
    return true;
}

// accessors for pcg.cu
int bcsr_view(sb_context* ctx, int* nbr, size_t* nnzb, const unsigned long long** rows, const int32_t** cols, const float** vals)
{
    Assembly* A = ctx->assembly;
    if (!A || !A->numeric_valid) return fail(ctx, SB_ERR_STATE, "no assembled matrix: call sb_assemble first");
    *nbr = A->nbr; *nnzb = A->nnzb; *rows = A->rows.p; *cols = A->cols.p; *vals = A->vals.p;
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_assemble(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    return assemble_internal(ctx);
}

int sb_bcsr_info(sb_context* ctx, int* n_block_rows, int64_t* nnzb)
{
    if (!ctx) return SB_ERR_ARG;
    Assembly* A = ctx->assembly;
    if (!A || !A->numeric_valid) return fail(ctx, SB_ERR_STATE, "sb_bcsr_info: call sb_assemble first");
    if (n_block_rows) *n_block_rows = A->nbr;
    if (nnzb) *nnzb = (int64_t)A->nnzb;
    return SB_OK;
}

int sb_bcsr_get(sb_context* ctx, int64_t* host_rows, int32_t* host_cols, float* host_vals)
{
    if (!ctx) return SB_ERR_ARG;
    Assembly* A = ctx->assembly;
    if (!A || !A->numeric_valid) return fail(ctx, SB_ERR_STATE, "sb_bcsr_get: call sb_assemble first");
    if (host_rows) SB_CUDA(ctx, cudaMemcpyAsync(host_rows, A->rows.p, sizeof(int64_t) * (A->nbr + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (host_cols) SB_CUDA(ctx, cudaMemcpyAsync(host_cols, A->cols.p, sizeof(int32_t) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_vals) SB_CUDA(ctx, cudaMemcpyAsync(host_vals, A->vals.p, sizeof(float) * 9 * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

}  // extern "C"
