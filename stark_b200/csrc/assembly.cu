// Global Hessian assembly: element Hessians (COO of 3x3 blocks) -> 3x3-block CSR, float storage.
//
// Replaces ElementHessians::{assemble_global, update_global} (symx/solver/second_order/ElementHessians.cpp:224-294)
// and BlockedSparseMatrix::{start_insertion, add_block, update_block, end_insertion} (bsm/BlockedSparseMatrix.h:332-593,
// 782-895).  The reference inserts every block under a per-row mutex after a binary search; here assembly is
//   * a STATIC SYMBOLIC phase, run only when a mesh / joint connectivity changes (practically once): one 64-bit key
//     (block_row, block_col) per element block of the static potentials (3.4 M at the 200k-tet scene), radix sort,
//     run-length heads -> the static block list and, per block, the contiguous list of its sources;
//   * a DYNAMIC SYMBOLIC phase, run whenever the contact / friction tables changed (every evaluation in contact): the few
//     thousand blocks of the contact elements are sorted and reduced the same way, then MERGED into the static list --
//     a binary search per dynamic block, a rank per static block -- which yields the final pattern (rows / cols) and, per
//     final block, its static and its dynamic source segment.  Cost: O(#contact blocks log) + two passes over the
//     pattern, instead of re-sorting millions of unchanged keys;
//   * a NUMERIC phase: a segmented reduction -- nine threads per BCSR block sum their sources in FP64 and store float
//     (the reference accumulates in float in thread order, BlockedSparseMatrix.h:899-952); blocks with very many sources
//     (a rigid body touched by thousands of contacts) get a CTA each; after a PD projection only the blocks of the
//     projected elements are re-summed (the reference's update_global adds projected - original).
// Output layout is the reference's (BlockedSparseMatrix.h:266-271): rows = offsets per block row, cols = first scalar
// column of the block, vals = 9 floats per block, column-major inside the block.
#include "internal.h"
#include <cstdlib>
#include <algorithm>
#include <cub/cub.cuh>

namespace sb {

// one sorted-and-reduced class of sources (static or dynamic)
struct SourceSet {
    size_t n = 0;                    // sources
    size_t n_blocks = 0;             // distinct (row, col) keys
    DevBuf<uint64_t> keys, keys_sorted;
    DevBuf<uint32_t> ids, ids_sorted;
    DevBuf<uint32_t> src_off;        // per source (unsorted numbering): offset of its (0,0) entry in ctx->H
    DevBuf<uint8_t> src_pitch;       // per source: row pitch n of its element Hessian
    DevBuf<uint32_t> sorted_off;     // per sorted source
    DevBuf<uint8_t> sorted_pitch;
    DevBuf<uint32_t> head, blk_of;   // head flags / 1-based block of every sorted source
    DevBuf<uint64_t> blk_key;        // [n_blocks] key of every distinct block
    DevBuf<uint32_t> seg;            // [n_blocks + 1] first sorted source of every block
    DevBuf<uint32_t> blk_of_src;     // per source (unsorted numbering): its block in THIS set
    void release()
    {
        keys.release(); keys_sorted.release(); ids.release(); ids_sorted.release(); src_off.release(); src_pitch.release();
        sorted_off.release(); sorted_pitch.release(); head.release(); blk_of.release(); blk_key.release(); seg.release(); blk_of_src.release();
    }
};

struct Assembly {
    uint64_t built_static = 0, built_dynamic = 0;
    size_t built_n_src = 0;
    int nbr = 0;
    size_t nnzb = 0;
    SourceSet S, D;                  // static / dynamic sources
    // merge
    DevBuf<uint32_t> d_pos;          // per dynamic block: lower bound in the static block list
    DevBuf<uint32_t> d_isnew, d_newrank;   // per dynamic block: not in the static list / inclusive rank among the new ones
    DevBuf<uint32_t> newpos;         // per new block: its d_pos (ascending)
    DevBuf<uint32_t> s_final;        // per static block: final BCSR block
    DevBuf<uint32_t> d_final;        // per dynamic block: final BCSR block
    DevBuf<int4> seg4;               // per final block: static [x, y) and dynamic [z, w) sorted-source ranges
    DevBuf<int32_t> blk_row;         // [nnzb] block row of every BCSR block
    DevBuf<unsigned long long> rows; // [nbr + 1]
    DevBuf<int32_t> cols;            // [nnzb] scalar column
    DevBuf<float> vals;              // [9 nnzb]
    DevBuf<uint8_t> temp;            // cub scratch
    DevBuf<uint8_t> dirty;           // per BCSR block: a source changed since the last numeric pass (PD projection)
    DevBuf<uint32_t> long_blocks;    // BCSR blocks with more than LONG_SEG sources
    int key_shift = 20;              // key = block_row << key_shift | block_col
    DevBuf<PotDesc> descs;           // per-class potential descriptors of the key kernel
    int* d_counts = nullptr;         // [0] long blocks [1] new blocks
    int* h_counts = nullptr;
    uint64_t assembled_eval = 0;     // evaluation the values belong to
    bool numeric_valid = false;
    // symbolic phase launched ahead on a side stream by the P+G+H evaluation (see assembly_prefetch_symbolic)
    bool pf_pending = false;
    uint64_t pf_static = 0, pf_dynamic = 0;
    size_t pf_n_src = 0;
    bool pf_failed = false;
    cudaEvent_t ev_sym = nullptr;
    // SCATTER MODE.  When the contact tables change but every block their elements touch already exists in the pattern (contact
    // that persists: pairs come and go inside the same neighbourhoods), nothing is sorted or merged: a kernel behind the
    // dynamic potentials looks every dynamic source's block up in the current BCSR rows (k_dyn_locate_src, one binary search
    // in a row), and the numeric phase adds the dynamic contributions through a small FP64 hash table keyed by the BCSR
    // block (k_scatter_dynamic) which the segmented reduction then picks up.  Blocks of pairs that are gone stay in the
    // pattern as explicit zeros until the next sort-based rebuild (sb_bcsr_get drops them: the reference drops zero blocks at
    // end_insertion).  A source whose block is missing raises a flag that travels to the host with the evaluation's scalars,
    // and the sort-based symbolic phase above runs as before.
    bool scatter_mode = false;
    int n_long_static = 0;             // static blocks with more than LONG_SEG sources
    DevBuf<uint32_t> d_final_of_src;   // per dynamic source: its BCSR block (both modes; the projection's dirty marking uses it)
    DevBuf<uint8_t> has_dyn;           // per BCSR block: receives dynamic contributions in this numeric pass
    DevBuf<unsigned long long> own_range;   // several GPUs: the block range of this rank's rows
    DevBuf<int32_t> hkeys;             // hash table: BCSR block or -1
    DevBuf<double> hacc;               // 9 FP64 accumulators per slot
    size_t hcap = 0;                   // slots (power of two)
    PotDesc* h_descs = nullptr;        // pinned staging of the dynamic potentials' descriptors
    size_t loc_n = 0;                  // dynamic sources of the last k_dyn_locate_src
    size_t loc_cap = 0;                // fused path: sources the per-source arrays can hold
    uint64_t loc_dynamic = 0;          // ... and the table version it looked at
    long long n_scatter_hits = 0, n_scatter_misses = 0;
    cudaEvent_t ev_loc = nullptr, ev_scatter = nullptr;
    bool miss_flag_dirty = true;        // ctx->d_scalars[2] must be cleared before the next lookup
    bool scatter_in_flight = false;     // a speculative pass has been issued on the side stream and not yet been ordered before the context stream
    uint64_t scatter_done_eval = 0;     // evaluation whose dynamic contributions already sit in the hash table (speculative pass)
};

static int scatter_pass(sb_context* ctx, Assembly* A, cudaStream_t st);
__global__ void k_scatter_clear(uint8_t* __restrict__ has_dyn, size_t n_blocks, int32_t* __restrict__ hkeys, double* __restrict__ hacc, size_t cap);

static Assembly* get(sb_context* ctx)
{
    if (!ctx->assembly) {
        ctx->assembly = new Assembly();
        cudaMalloc(&ctx->assembly->d_counts, 4 * sizeof(int));
        cudaMallocHost(&ctx->assembly->h_counts, 4 * sizeof(int));
        cudaEventCreateWithFlags(&ctx->assembly->ev_sym, cudaEventDisableTiming);
    }
    return ctx->assembly;
}
void assembly_destroy(sb_context* ctx)
{
    Assembly* A = ctx->assembly;
    if (!A) return;
    A->S.release(); A->D.release();
    A->d_pos.release(); A->d_isnew.release(); A->d_newrank.release(); A->newpos.release(); A->s_final.release(); A->d_final.release();
    A->seg4.release(); A->blk_row.release(); A->rows.release(); A->own_range.release(); A->cols.release(); A->vals.release(); A->temp.release();
    A->dirty.release(); A->long_blocks.release(); A->descs.release();
    A->d_final_of_src.release(); A->has_dyn.release(); A->hkeys.release(); A->hacc.release();
    if (A->ev_loc) { cudaEventDestroy(A->ev_loc); cudaEventDestroy(A->ev_scatter); }
    if (A->h_descs) cudaFreeHost(A->h_descs);
    if (std::getenv("SB_ASM_DUMP")) fprintf(stderr, "[stark_b200 assembly] contact-table changes absorbed without a symbolic phase: %lld, with a sort-based rebuild: %lld\n", A->n_scatter_hits, A->n_scatter_misses);
    if (A->d_counts) cudaFree(A->d_counts);
    if (A->h_counts) cudaFreeHost(A->h_counts);
    if (ctx->issuer) ctx->issuer->wait();
    if (A->ev_sym) { cudaEventSynchronize(A->ev_sym); cudaEventDestroy(A->ev_sym); }
    delete A;
    ctx->assembly = nullptr;
}

// keys of the element blocks of ALL potentials of one class in one launch (descs sorted by blk_off; key = row << shift | col)
__global__ void k_make_keys(const PotDesc* __restrict__ descs, int n_descs, size_t n_total, int shift, const int32_t* __restrict__ rows_all,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ ids, uint32_t* __restrict__ src_off, uint8_t* __restrict__ src_pitch)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_total) return;
    int lo = 0, hi = n_descs - 1;   // last desc with blk_off <= g
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].blk_off <= g) lo = mid; else hi = mid - 1;
    }
    const PotDesc d = descs[lo];
    const size_t t = g - d.blk_off;
    const int nb2 = d.nb * d.nb;
    const int e = (int)(t / nb2);
    const int k = (int)(t - (size_t)e * nb2);
    const int bi = k / d.nb, bj = k - bi * d.nb;
    const int32_t* r = rows_all + d.rows_off + (size_t)e * d.nb;
    const size_t id = d.blk_off + t;
    const int n = 3 * d.nb;
    keys[id] = ((uint64_t)(uint32_t)r[bi] << shift) | (uint32_t)r[bj];
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

// blk_of = inclusive scan of head (1-based); block keys and segment starts at heads; per-source offsets in sorted order
__global__ void k_fill_set(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ head, const uint32_t* __restrict__ blk_of,
                           const uint32_t* __restrict__ ids_sorted, const uint32_t* __restrict__ src_off, const uint8_t* __restrict__ src_pitch,
                           uint32_t* __restrict__ sorted_off, uint8_t* __restrict__ sorted_pitch,
                           uint32_t* __restrict__ seg, uint64_t* __restrict__ blk_key, uint32_t* __restrict__ blk_of_src, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids_sorted[i];
    sorted_off[i] = src_off[id];
    sorted_pitch[i] = src_pitch[id];
    const uint32_t b = blk_of[i] - 1;
    blk_of_src[id] = b;
    if (head[i]) {
        seg[b] = (uint32_t)i;
        blk_key[b] = keys[i];
    }
    if (i == n - 1) seg[b + 1] = (uint32_t)n;
}

// ---- merge of the dynamic block list into the static one ----
__global__ void k_dyn_locate(const uint64_t* __restrict__ dkey, const uint32_t* __restrict__ ndb_p, size_t n_upper, const uint64_t* __restrict__ skey, size_t nsb,
                             uint32_t* __restrict__ d_pos, uint32_t* __restrict__ d_isnew)
{
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_upper) return;
    if (j >= *ndb_p) { d_pos[j] = 0; d_isnew[j] = 0; return; }   // padding up to the host-side upper bound
    const uint64_t k = dkey[j];
    size_t lo = 0, hi = nsb;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (skey[mid] < k) lo = mid + 1; else hi = mid;
    }
    d_pos[j] = (uint32_t)lo;
    d_isnew[j] = (lo < nsb && skey[lo] == k) ? 0u : 1u;
}
// compacted list of the positions of the new blocks; their total
__global__ void k_dyn_compact(const uint32_t* __restrict__ d_pos, const uint32_t* __restrict__ d_isnew, const uint32_t* __restrict__ d_newrank_incl,
                              uint32_t* __restrict__ newpos, size_t n_upper, int* __restrict__ n_new)
{
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_upper) return;
    if (d_isnew[j]) newpos[d_newrank_incl[j] - 1] = d_pos[j];
    if (j == n_upper - 1) *n_new = (int)d_newrank_incl[j];
}
// final index of every static block = own index + number of new blocks inserted before it (new keys with pos <= i)
__global__ void k_static_final(const uint64_t* __restrict__ skey, const uint32_t* __restrict__ sseg, size_t nsb,
                               const uint32_t* __restrict__ newpos, const int* __restrict__ n_new_p,
                               uint32_t* __restrict__ s_final, int4* __restrict__ seg4, int32_t* __restrict__ blk_row, int32_t* __restrict__ cols, int shift)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsb) return;
    const int n_new = *n_new_p;
    int lo = 0, hi = n_new;       // upper_bound(newpos, i)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (newpos[mid] <= (uint32_t)i) lo = mid + 1; else hi = mid;
    }
    const uint32_t f = (uint32_t)i + (uint32_t)lo;
    s_final[i] = f;
    seg4[f] = make_int4((int)sseg[i], (int)sseg[i + 1], 0, 0);
    const uint64_t k = skey[i];
    blk_row[f] = (int32_t)(k >> shift);
    cols[f] = 3 * (int32_t)(k & ((1ull << shift) - 1));
}
__global__ void k_dyn_final(const uint64_t* __restrict__ dkey, const uint32_t* __restrict__ dseg, const uint32_t* __restrict__ ndb_p,
                            const uint32_t* __restrict__ d_pos, const uint32_t* __restrict__ d_isnew, const uint32_t* __restrict__ d_newrank_incl,
                            const uint32_t* __restrict__ s_final, uint32_t* __restrict__ d_final,
                            int4* __restrict__ seg4, int32_t* __restrict__ blk_row, int32_t* __restrict__ cols, int shift)
{
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *ndb_p) return;
    uint32_t f;
    if (d_isnew[j]) {
        f = d_pos[j] + (d_newrank_incl[j] - 1);
        seg4[f] = make_int4(0, 0, (int)dseg[j], (int)dseg[j + 1]);
        const uint64_t k = dkey[j];
        blk_row[f] = (int32_t)(k >> shift);
        cols[f] = 3 * (int32_t)(k & ((1ull << shift) - 1));
    } else {
        f = s_final[d_pos[j]];
        // the static thread wrote (x, y, 0, 0); only z, w are touched here (k_static_final has completed: stream order)
        seg4[f].z = (int)dseg[j];
        seg4[f].w = (int)dseg[j + 1];
    }
    d_final[j] = f;
}

// sort-based build: per dynamic source, its BCSR block (the scatter mode's k_dyn_locate_src produces the same table)
__global__ void k_src_final(const uint32_t* __restrict__ blk_of_src, const uint32_t* __restrict__ d_final, uint32_t* __restrict__ d_final_of_src, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d_final_of_src[i] = d_final[blk_of_src[i]];
}

// Scatter mode, step 1 (behind the dynamic potentials' kernels of an evaluation): BCSR block of every dynamic source by a
// binary search for its column inside its block row; a missing block raises *miss (a double: it rides with the evaluation's
// scalars).  Also records where the source's 3x3 block lies in the element-Hessian store.
__global__ void k_dyn_locate_src(const PotDesc* __restrict__ descs, int n_descs, size_t n_total, const DynTotals* __restrict__ tot, const int32_t* __restrict__ rows_all,
                                 const unsigned long long* __restrict__ row_ptr, const int32_t* __restrict__ cols, int nbr,
                                 uint32_t* __restrict__ d_final_of_src, uint32_t* __restrict__ src_off, uint8_t* __restrict__ src_pitch, double* __restrict__ miss)
{
    if (tot) { n_descs = tot->n_descs; n_total = tot->n_dyn_src; }   // (fused path: counts known on the device only)
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_total; g += (size_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_descs - 1;   // last desc with blk_off <= g
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].blk_off <= g) lo = mid; else hi = mid - 1;
    }
    const PotDesc d = descs[lo];
    const size_t t = g - d.blk_off;
    const int nb2 = d.nb * d.nb;
    const int e = (int)(t / nb2);
    const int k = (int)(t - (size_t)e * nb2);
    const int bi = k / d.nb, bj = k - bi * d.nb;
    const int32_t* r = rows_all + d.rows_off + (size_t)e * d.nb;
    const int n = 3 * d.nb;
    src_off[g] = (uint32_t)(d.H_off + (size_t)e * n * n + (size_t)(3 * bi) * n + 3 * bj);
    src_pitch[g] = (uint8_t)n;
    const int row = r[bi], col = 3 * r[bj];
    uint32_t f = 0xffffffffu;
    if (row >= 0 && row < nbr) {
        size_t a = row_ptr[row], b = row_ptr[row + 1];
        while (a < b) {
            const size_t mid = (a + b) >> 1;
            if (cols[mid] < col) a = mid + 1; else b = mid;
        }
        if (a < row_ptr[row + 1] && cols[a] == col) f = (uint32_t)a;
    }
    d_final_of_src[g] = f;
    if (f == 0xffffffffu) *miss = 1.0;
  }
}

// Scatter mode, step 2 (numeric phase): every dynamic source adds its 3x3 block to the FP64 accumulators of its BCSR block in a
// lock-free open-addressing table (9 threads per source).
__device__ __forceinline__ uint32_t blk_hash(uint32_t f) { f *= 0x9E3779B1u; return f ^ (f >> 15); }
// A persistent grid; every CTA first combines the contributions of its sources per BCSR block in a small shared-memory table
// (the rigid-body blocks receive a contribution from EVERY contact element: without this, thousands of FP64 atomics queue on
// the same nine addresses), then flushes the combined blocks into the global table.
constexpr int SC_SLOTS = 256;      // shared-memory slots per CTA (power of two); overflow goes straight to the global table
constexpr int SC_THREADS = 288;    // 32 sources x 9 entries per pass
__device__ __forceinline__ void scatter_global(uint32_t f, int k, double v, int32_t* __restrict__ hkeys, double* __restrict__ hacc, uint32_t hmask)
{
    uint32_t slot = blk_hash(f) & hmask;
    while (true) {
        const int32_t prev = atomicCAS(hkeys + slot, -1, (int32_t)f);
        if (prev == -1 || prev == (int32_t)f) break;
        slot = (slot + 1) & hmask;
    }
    atomicAdd(hacc + 9 * (size_t)slot + k, v);
}
__global__ void __launch_bounds__(SC_THREADS) k_scatter_dynamic(const double* __restrict__ H, const uint32_t* __restrict__ d_final_of_src, const uint32_t* __restrict__ src_off,
                                  const uint8_t* __restrict__ src_pitch, size_t n_src, const DynTotals* __restrict__ tot, int32_t* __restrict__ hkeys, double* __restrict__ hacc,
                                  uint32_t hmask, uint8_t* __restrict__ has_dyn)
{
    if (tot) { n_src = tot->n_dyn_src; if (2 * n_src > (size_t)hmask + 1) return; }   // (fused path; a table too small for the actual count: the host redoes the pass)
    __shared__ int32_t s_key[SC_SLOTS];
    __shared__ double s_acc[SC_SLOTS][9];
    for (int i = threadIdx.x; i < SC_SLOTS; i += SC_THREADS) s_key[i] = -1;
    for (int i = threadIdx.x; i < SC_SLOTS * 9; i += SC_THREADS) (&s_acc[0][0])[i] = 0.0;
    __syncthreads();
    const int k = threadIdx.x % 9, lane_src = threadIdx.x / 9;
    const int r = k % 3, c = k / 3;
    for (size_t g = (size_t)blockIdx.x * 32 + lane_src; g < n_src; g += (size_t)gridDim.x * 32) {
        const uint32_t f = d_final_of_src[g];
        if (f == 0xffffffffu) continue;   // (a speculative pass behind a lookup that found a block missing: its result is discarded)
        const double v = H[(size_t)src_off[g] + (size_t)r * src_pitch[g] + c];
        if (k == 0) has_dyn[f] = 1;
        uint32_t slot = blk_hash(f) & (SC_SLOTS - 1);
        bool placed = false;
        for (int probe = 0; probe < 8; probe++) {
            const int32_t prev = atomicCAS(s_key + slot, -1, (int32_t)f);
            if (prev == -1 || prev == (int32_t)f) { placed = true; break; }
            slot = (slot + 1) & (SC_SLOTS - 1);
        }
        if (placed) atomicAdd(&s_acc[slot][k], v);
        else scatter_global(f, k, v, hkeys, hacc, hmask);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SC_SLOTS * 9; i += SC_THREADS) {
        const int slot = i / 9, kk = i - 9 * slot;
        const int32_t f = s_key[slot];
        if (f >= 0) scatter_global((uint32_t)f, kk, s_acc[slot][kk], hkeys, hacc, hmask);
    }
}
__device__ __forceinline__ double dyn_lookup(uint32_t f, int k, const int32_t* __restrict__ hkeys, const double* __restrict__ hacc, uint32_t hmask)
{
    uint32_t slot = blk_hash(f) & hmask;
    while (hkeys[slot] != (int32_t)f) slot = (slot + 1) & hmask;   // (present: has_dyn[f] was set by the thread that inserted it)
    return hacc[9 * (size_t)slot + k];
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

constexpr int LONG_SEG = 64;
constexpr int LONG_CAP = 4096;      // long blocks beyond this many are summed by the plain path
constexpr int LONG_CTAS = 64;

// a long block is left to k_assemble_long only if it made it into the (capped) list
__device__ __forceinline__ bool long_rank_ok(size_t b, const uint32_t* __restrict__ long_blocks, const int* __restrict__ n_long)
{
    const int n = *n_long;
    if (n <= LONG_CAP) return true;          // nothing was dropped
    for (int i = 0; i < LONG_CAP; i++) if (long_blocks[i] == (uint32_t)b) return true;
    return false;
}

__global__ void k_find_long(const int4* __restrict__ seg4, uint32_t* __restrict__ long_blocks, int* __restrict__ n_long, size_t nnzb)
{
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nnzb) return;
    const int4 g = seg4[b];
    if ((g.y - g.x) + (g.w - g.z) > LONG_SEG) {
        const int k = atomicAdd(n_long, 1);
        if (k < LONG_CAP) long_blocks[k] = (uint32_t)b;
    }
}

// static blocks with more than LONG_SEG sources (counted once per static build: when there are none, scatter-mode passes do not
// launch k_assemble_long at all)
__global__ void k_count_long_static(const uint32_t* __restrict__ seg, const uint32_t* __restrict__ n_blocks_p, int* __restrict__ out)
{
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= *n_blocks_p) return;
    if (seg[b + 1] - seg[b] > LONG_SEG) atomicAdd(out, 1);
}

struct NumericArgs {
    const double* H;
    const int4* seg4;
    const uint32_t* s_off; const uint8_t* s_pitch;   // static sources, sorted
    const uint32_t* d_off; const uint8_t* d_pitch;   // dynamic sources, sorted
    float* vals;
    uint8_t* dirty;
    const uint32_t* long_blocks; const int* n_long;
    size_t nnzb;
    // scatter mode: the dynamic ranges of seg4 are stale; dynamic contributions come from the hash table
    int scatter;
    const uint8_t* has_dyn; const int32_t* hkeys; const double* hacc; uint32_t hmask;
    // several GPUs sharing the solve: only the blocks [own_range[0], own_range[1]) of this rank's rows are summed (null: all)
    const unsigned long long* own_range;
};

// One CTA per long block: 32 strided partial sums per entry, then a fixed shared-memory tree (deterministic).
template<bool ONLY_DIRTY>
__global__ void __launch_bounds__(288) k_assemble_long(const NumericArgs a)
{
    __shared__ double sm[32][9];
    const int total = min(*a.n_long, LONG_CAP);
    const int k = threadIdx.x % 9, lane = threadIdx.x / 9;
    const int r = k % 3, c = k / 3;
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
        const uint32_t b = a.long_blocks[i];
        if (ONLY_DIRTY && !a.dirty[b]) continue;   // uniform across the CTA
        if (a.own_range && (b < a.own_range[0] || b >= a.own_range[1])) continue;
        const int4 g = a.seg4[b];
        if (a.scatter && g.y - g.x <= LONG_SEG) continue;   // (long through its dynamic sources at the last rebuild only: the plain kernel sums it; uniform across the CTA)
        double acc = 0.0;
        for (int s = g.x + lane; s < g.y; s += 32) acc += a.H[(size_t)a.s_off[s] + (size_t)r * a.s_pitch[s] + c];
        if (!a.scatter) { for (int s = g.z + lane; s < g.w; s += 32) acc += a.H[(size_t)a.d_off[s] + (size_t)r * a.d_pitch[s] + c]; }
        else if (lane == 0 && a.has_dyn[b]) acc += dyn_lookup(b, k, a.hkeys, a.hacc, a.hmask);
        sm[lane][k] = acc;
        __syncthreads();
        for (int w = 16; w > 0; w >>= 1) {
            if (lane < w) sm[lane][k] += sm[lane + w][k];
            __syncthreads();
        }
        if (lane == 0) a.vals[9 * (size_t)b + k] = (float)sm[0][k];
        __syncthreads();
    }
}

// Segmented reduction: 9 consecutive threads own one BCSR block (thread k -> entry (r = k % 3, c = k / 3), column-major).
// Sources are summed in their sorted order (deterministic): static ones first, then dynamic ones; four loads in flight.
template<bool ONLY_DIRTY>
__global__ void __launch_bounds__(288) k_assemble_numeric(const NumericArgs a)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t b = t / 9;
    if (b >= a.nnzb) return;
    if (ONLY_DIRTY && !a.dirty[b]) return;
    if (a.own_range && (b < a.own_range[0] || b >= a.own_range[1])) return;
    const int k = (int)(t - b * 9);
    const int r = k % 3, c = k / 3;
    const int4 g = a.seg4[b];
    if ((a.scatter ? (g.y - g.x) : (g.y - g.x) + (g.w - g.z)) > LONG_SEG && long_rank_ok(b, a.long_blocks, a.n_long)) return;   // summed by k_assemble_long
    double acc = 0.0;
    int s = g.x;
    for (; s + 4 <= g.y; s += 4) {
        const size_t o0 = (size_t)a.s_off[s] + (size_t)r * a.s_pitch[s] + c;
        const size_t o1 = (size_t)a.s_off[s + 1] + (size_t)r * a.s_pitch[s + 1] + c;
        const size_t o2 = (size_t)a.s_off[s + 2] + (size_t)r * a.s_pitch[s + 2] + c;
        const size_t o3 = (size_t)a.s_off[s + 3] + (size_t)r * a.s_pitch[s + 3] + c;
        const double v0 = a.H[o0], v1 = a.H[o1], v2 = a.H[o2], v3 = a.H[o3];
        acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; s < g.y; s++) acc += a.H[(size_t)a.s_off[s] + (size_t)r * a.s_pitch[s] + c];
    if (!a.scatter) { for (s = g.z; s < g.w; s++) acc += a.H[(size_t)a.d_off[s] + (size_t)r * a.d_pitch[s] + c]; }
    else if (a.has_dyn[b]) acc += dyn_lookup((uint32_t)b, k, a.hkeys, a.hacc, a.hmask);
    a.vals[t] = (float)acc;
}
// the dirty flags are cleared by a separate pass (the nine threads of a block must all have seen the flag)
__global__ void k_clear_dirty(uint8_t* __restrict__ dirty, size_t nnzb)
{
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nnzb) dirty[b] = 0;
}

// sort + reduce one class of sources
static int build_set(sb_context* ctx, Assembly* A, SourceSet& X, bool dynamic, cudaStream_t st)
{
    size_t n = 0;
    for (auto& p : ctx->potentials) if (p.dynamic == dynamic) n += (size_t)p.n_elem * p.k->nb * p.k->nb;
    X.n = n;
    X.n_blocks = 0;
    if (n == 0) return 0;
    X.keys.ensure(n + 1); X.keys_sorted.ensure(n + 1); X.ids.ensure(n + 1); X.ids_sorted.ensure(n + 1);
    X.src_off.ensure(n + 1); X.src_pitch.ensure(n + 1); X.sorted_off.ensure(n + 1); X.sorted_pitch.ensure(n + 1);
    X.head.ensure(n + 1); X.blk_of.ensure(n + 1); X.blk_of_src.ensure(n + 1);
    X.blk_key.ensure(n + 1); X.seg.ensure(n + 2);   // upper bounds: every source its own block
    std::vector<PotDesc> descs;
    size_t blk_off = 0;
    for (int pi : layout_order(ctx)) {
        Potential& p = ctx->potentials[pi];
        if (p.dynamic != dynamic || p.n_elem == 0) continue;
        PotDesc d;
        d.H_off = p.H_off; d.rows_off = p.rows_off; d.blk_off = blk_off; d.n_elem = p.n_elem; d.nb = p.k->nb;
        descs.push_back(d);
        blk_off += (size_t)p.n_elem * d.nb * d.nb;
    }
    A->descs.ensure(descs.size() + 1);
    SB_CUDA(ctx, cudaMemcpyAsync(A->descs.p, descs.data(), descs.size() * sizeof(PotDesc), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    k_make_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A->descs.p, (int)descs.size(), n, A->key_shift, ctx->rows.p, X.keys.p, X.ids.p, X.src_off.p, X.src_pitch.p);
    ctx->launches++;
    const int key_bits = 2 * A->key_shift;
    size_t temp_bytes = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, X.keys.p, X.keys_sorted.p, X.ids.p, X.ids_sorted.p, (int)n, 0, key_bits, st);
    cub::DeviceScan::InclusiveSum(nullptr, tb2, X.head.p, X.blk_of.p, (int)n, st);
    A->temp.ensure(std::max(temp_bytes, tb2) + 16);
    temp_bytes = A->temp.cap;
    SB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(A->temp.p, temp_bytes, X.keys.p, X.keys_sorted.p, X.ids.p, X.ids_sorted.p, (int)n, 0, key_bits, st));
    k_heads<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(X.keys_sorted.p, X.head.p, n);
    temp_bytes = A->temp.cap;
    SB_CUDA(ctx, cub::DeviceScan::InclusiveSum(A->temp.p, temp_bytes, X.head.p, X.blk_of.p, (int)n, st));
    k_fill_set<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(X.keys_sorted.p, X.head.p, X.blk_of.p, X.ids_sorted.p, X.src_off.p, X.src_pitch.p,
                                                             X.sorted_off.p, X.sorted_pitch.p, X.seg.p, X.blk_key.p, X.blk_of_src.p, n);
    ctx->launches += 7;
    if (dynamic) { X.n_blocks = n; return 0; }   // upper bound; the exact count stays on the device (blk_of[n - 1])
    uint32_t nb32 = 0;
    int n_long_static = 0;
    SB_CUDA(ctx, cudaMemsetAsync(A->d_counts + 3, 0, sizeof(int), st));
    k_count_long_static<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(X.seg.p, X.blk_of.p + (n - 1), A->d_counts + 3);
    SB_CUDA(ctx, cudaMemcpyAsync(&nb32, X.blk_of.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaMemcpyAsync(&n_long_static, A->d_counts + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    X.n_blocks = nb32;
    A->n_long_static = n_long_static;
    return 0;
}

// Phase A of the symbolic build, on stream st, without a host synchronisation: sort the sources, merge the dynamic blocks into
// the static list, and start the copy of the block count to the host.  Phase B (below) needs that count.
static int symbolic_phase_a(sb_context* ctx, Assembly* A, bool rebuild_static_in, cudaStream_t st)
{
    bool rebuild_static = rebuild_static_in;
    if (ctx->H_total >= (1ull << 32)) return fail(ctx, SB_ERR_STATE, "sb_assemble: element Hessian storage exceeds 32-bit offsets");
    A->nbr = ctx->ndofs / 3;
    {
        int bits = 1;
        while ((1ll << bits) < A->nbr + 1) bits++;
        if (bits != A->key_shift) { A->key_shift = bits; rebuild_static = true; }   // the static keys use the same encoding
    }
    int r;
    if (rebuild_static && (r = build_set(ctx, A, A->S, false, st))) return r;
    if ((r = build_set(ctx, A, A->D, true, st))) return r;
    const size_t nsb = A->S.n_blocks, ndb = A->D.n_blocks;   // ndb: host-side UPPER bound (= number of dynamic sources)
    const uint32_t* ndb_dev = (A->D.n > 0) ? A->D.blk_of.p + (A->D.n - 1) : nullptr;
    if (nsb + ndb == 0) return fail(ctx, SB_ERR_STATE, "sb_assemble: no element Hessians");
    const size_t cap = nsb + ndb;   // upper bound of the merged pattern
    A->seg4.ensure(cap + 1); A->blk_row.ensure(cap + 1); A->cols.ensure(cap + 8); A->vals.ensure(9 * cap + 72);   // (slack: the PCG streams 4-block-aligned tiles that may run 3 blocks past the end)
    A->rows.ensure(A->nbr + 2); A->dirty.ensure(cap + 1); A->long_blocks.ensure(LONG_CAP);
    A->s_final.ensure(nsb + 1); A->d_final.ensure(ndb + 1);
    A->d_pos.ensure(ndb + 1); A->d_isnew.ensure(ndb + 1); A->d_newrank.ensure(ndb + 1); A->newpos.ensure(ndb + 1);
    SB_CUDA(ctx, cudaMemsetAsync(A->d_counts, 0, 4 * sizeof(int), st));
    if (ndb > 0) {
        k_dyn_locate<<<(unsigned)((ndb + 255) / 256), 256, 0, st>>>(A->D.blk_key.p, ndb_dev, ndb, A->S.blk_key.p, nsb, A->d_pos.p, A->d_isnew.p);
        size_t tb = 0;
        cub::DeviceScan::InclusiveSum(nullptr, tb, A->d_isnew.p, A->d_newrank.p, (int)ndb, st);
        A->temp.ensure(tb + 16);
        tb = A->temp.cap;
        SB_CUDA(ctx, cub::DeviceScan::InclusiveSum(A->temp.p, tb, A->d_isnew.p, A->d_newrank.p, (int)ndb, st));
        k_dyn_compact<<<(unsigned)((ndb + 255) / 256), 256, 0, st>>>(A->d_pos.p, A->d_isnew.p, A->d_newrank.p, A->newpos.p, ndb, A->d_counts + 1);
        ctx->launches += 3;
    }
    if (nsb > 0) {
        k_static_final<<<(unsigned)((nsb + 255) / 256), 256, 0, st>>>(A->S.blk_key.p, A->S.seg.p, nsb, A->newpos.p, A->d_counts + 1,
                                                                       A->s_final.p, A->seg4.p, A->blk_row.p, A->cols.p, A->key_shift);
        ctx->launches++;
    }
    if (ndb > 0) {
        k_dyn_final<<<(unsigned)((ndb + 255) / 256), 256, 0, st>>>(A->D.blk_key.p, A->D.seg.p, ndb_dev, A->d_pos.p, A->d_isnew.p, A->d_newrank.p,
                                                                    A->s_final.p, A->d_final.p, A->seg4.p, A->blk_row.p, A->cols.p, A->key_shift);
        A->d_final_of_src.ensure(A->D.n + 1);
        k_src_final<<<(unsigned)((A->D.n + 255) / 256), 256, 0, st>>>(A->D.blk_of_src.p, A->d_final.p, A->d_final_of_src.p, A->D.n);
        ctx->launches += 2;
    }
    SB_CUDA(ctx, cudaMemcpyAsync(A->h_counts, A->d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    return 0;
}
// Phase B, on the context stream, once phase A has completed (the caller has synchronised with it)
static int symbolic_phase_b(sb_context* ctx, Assembly* A)
{
    cudaStream_t st = ctx->stream;
    const size_t nsb = A->S.n_blocks;
    A->nnzb = nsb + (size_t)A->h_counts[1];
    k_row_ptr<<<(A->nbr + 1 + 255) / 256, 256, 0, st>>>(A->blk_row.p, A->rows.p, A->nbr, A->nnzb);
    k_find_long<<<(unsigned)((A->nnzb + 255) / 256), 256, 0, st>>>(A->seg4.p, A->long_blocks.p, A->d_counts, A->nnzb);
    SB_CUDA(ctx, cudaMemsetAsync(A->dirty.p, 0, A->nnzb + 1, st));
    ctx->launches += 2;
    SB_CUDA(ctx, cudaGetLastError());
    A->built_static = ctx->static_version;
    A->built_dynamic = ctx->dynamic_version;
    A->built_n_src = ctx->n_blocks_total;
    A->scatter_mode = false;   // (seg4 carries this evaluation's dynamic source ranges)
    return 0;
}
static int build_symbolic(sb_context* ctx, Assembly* A, bool rebuild_static)
{
    StageTimer timer(ctx, ST_ASM_SYMBOLIC);
    int r = symbolic_phase_a(ctx, A, rebuild_static, ctx->stream);
    if (r) return r;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return symbolic_phase_b(ctx, A);
}

static bool pattern_current(sb_context* ctx, Assembly* A)
{
    return A->built_static == ctx->static_version && A->built_dynamic == ctx->dynamic_version && A->built_n_src == ctx->n_blocks_total && A->nbr == ctx->ndofs / 3;
}

void preload_assembly_kernels()
{
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_make_keys); cudaFuncGetAttributes(&fa, k_heads); cudaFuncGetAttributes(&fa, k_fill_set);
    cudaFuncGetAttributes(&fa, k_dyn_locate); cudaFuncGetAttributes(&fa, k_dyn_compact); cudaFuncGetAttributes(&fa, k_static_final);
    cudaFuncGetAttributes(&fa, k_dyn_final); cudaFuncGetAttributes(&fa, k_row_ptr); cudaFuncGetAttributes(&fa, k_find_long);
    cudaFuncGetAttributes(&fa, k_assemble_numeric<true>); cudaFuncGetAttributes(&fa, k_assemble_numeric<false>);
    cudaFuncGetAttributes(&fa, k_assemble_long<true>); cudaFuncGetAttributes(&fa, k_assemble_long<false>);
    cudaFuncGetAttributes(&fa, k_clear_dirty);
    cudaGetLastError();
}

static bool static_part_stale(sb_context* ctx, Assembly* A)
{
    return A->built_static != ctx->static_version || A->nbr != ctx->ndofs / 3 || A->S.n != ctx->n_static_blocks;
}

// Called by the P+G+H evaluation once the kernels of the dynamic (contact / friction) potentials are enqueued on the side
// streams: if only the dynamic part of the pattern is stale -- every Newton iteration in contact -- the symbolic phase (a
// chain of ~20 small kernels, ~120 us) is launched on a side stream NOW, behind those kernels, and runs under the volume
// elements' evaluation, the reductions and the PD projection instead of after them.  assemble_internal picks the result up.
void assembly_prefetch_symbolic(sb_context* ctx, unsigned side_mask)
{
    Assembly* A = ctx->assembly;
    static const bool disabled = std::getenv("SB_NO_PREFETCH") != nullptr;   // diagnostic hook
    if (disabled) return;
    if (!A || !A->numeric_valid || A->pf_pending || ctx->n_blocks_total == 0) return;
    if (pattern_current(ctx, A) || static_part_stale(ctx, A)) return;
    {
        int bits = 1;
        while ((1ll << bits) < ctx->ndofs / 3 + 1) bits++;
        if (bits != A->key_shift) return;
    }
    A->pf_pending = true;
    A->pf_failed = false;
    A->pf_static = ctx->static_version; A->pf_dynamic = ctx->dynamic_version; A->pf_n_src = ctx->n_blocks_total;
    // SB_HELPER_THREAD=1: issued by the helper thread while this thread goes on launching the rest of the evaluation (measured:
    // no gain -- launches from two threads serialise in the driver); default: issued here
    static const bool helper = std::getenv("SB_HELPER_THREAD") != nullptr;
    auto job = [ctx, A, side_mask] {
        cudaStream_t st = ctx->sym_stream;
        for (int k = 0; k < sb_context::N_SIDE; k++) if (side_mask & (1u << k)) cudaStreamWaitEvent(st, ctx->ev_dyn[k], 0);   // the dynamic potentials' block rows are written
        timeline_point(st, "symbolic: dependencies met");
        if (symbolic_phase_a(ctx, A, false, st)) { cudaStreamSynchronize(st); A->pf_failed = true; }
        cudaEventRecord(A->ev_sym, st);
        timeline_point(st, "symbolic: phase A done");
    };
    if (helper) ctx->issuer->post(job); else job();
}
// host-side join with the helper thread, then device-side: before anything the prefetched phase reads (block rows, layout) is
// rewritten
void assembly_prefetch_drain(sb_context* ctx)
{
    Assembly* A = ctx->assembly;
    if (A && A->pf_pending) { ctx->issuer->wait(); cudaEventSynchronize(A->ev_sym); }
}

// Scatter mode, called by the P+G+H evaluation once the dynamic potentials' kernels are enqueued and joined into the context
// stream: can the changed contact tables be absorbed by the current pattern?  Launches k_dyn_locate_src; its miss flag lands in
// ctx->d_scalars[2] and travels to the host with the evaluation's scalars.  Returns true when the kernel was launched.
bool assembly_locate_possible(sb_context* ctx)
{
    Assembly* A = ctx->assembly;
    static const bool disabled = std::getenv("SB_NO_SCATTER") != nullptr;   // diagnostic hook
    if (disabled || !A || !A->numeric_valid || A->pf_pending || ctx->n_blocks_total == 0) return false;
    if (pattern_current(ctx, A) || static_part_stale(ctx, A)) return false;
    if (A->nnzb == 0 || A->nnzb >= (1ull << 31) || ctx->H_total >= (1ull << 32)) return false;
    int bits = 1;
    while ((1ll << bits) < ctx->ndofs / 3 + 1) bits++;
    return bits == A->key_shift;
}
bool assembly_locate_dynamic(sb_context* ctx)
{
    Assembly* A = ctx->assembly;
    if (!assembly_locate_possible(ctx)) return false;
    // descriptors of the dynamic potentials (same numbering as the sort-based path: layout order, blk_off within the class)
    if (!A->h_descs) cudaMallocHost(&A->h_descs, 128 * sizeof(PotDesc));
    int nd = 0;
    size_t blk_off = 0;
    for (int pi : layout_order(ctx)) {
        Potential& p = ctx->potentials[pi];
        if (!p.dynamic || p.n_elem == 0) continue;
        if (nd >= 128) return false;
        PotDesc d;
        d.H_off = p.H_off; d.rows_off = p.rows_off; d.blk_off = blk_off; d.n_elem = p.n_elem; d.nb = p.k->nb;
        A->h_descs[nd++] = d;
        blk_off += (size_t)p.n_elem * d.nb * d.nb;
    }
    const size_t n = blk_off;
    cudaStream_t st = ctx->stream;
    if (A->miss_flag_dirty) { cudaMemsetAsync(ctx->d_scalars + 2, 0, sizeof(double), st); A->miss_flag_dirty = false; }   // (only a lookup that missed leaves it set)
    A->loc_n = n;
    A->loc_dynamic = ctx->dynamic_version;
    if (n == 0) return true;   // no dynamic element left: the pattern trivially holds every block
    if (ctx->H_total >= (1ull << 32)) return false;
    A->descs.ensure(nd + 1);
    A->D.src_off.ensure(n + 1); A->D.src_pitch.ensure(n + 1); A->d_final_of_src.ensure(n + 1);
    cudaMemcpyAsync(A->descs.p, A->h_descs, nd * sizeof(PotDesc), cudaMemcpyHostToDevice, st);
    k_dyn_locate_src<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A->descs.p, nd, n, nullptr, ctx->rows.p, A->rows.p, A->cols.p, A->nbr, A->d_final_of_src.p,
                                                                   A->D.src_off.p, A->D.src_pitch.p, ctx->d_scalars + 2);
    ctx->launches++;
    // Speculatively, right behind the lookup: the numeric phase's scatter pass of this evaluation (valid if no block is missing,
    // which is the common case; the host learns that with the evaluation's scalars).  On the context stream: this stretch of the
    // evaluation is bound by the host's API calls, the GPU has time to spare, and a side stream would cost three more calls.
    {
        static const bool on_main = std::getenv("SB_SCATTER_MAIN") != nullptr;   // (experiment hook)
        const size_t keep_n = A->D.n;
        A->D.n = n;
        int rs;
        if (on_main) rs = scatter_pass(ctx, A, st);
        else {
            if (!A->ev_loc) { cudaEventCreateWithFlags(&A->ev_loc, cudaEventDisableTiming); cudaEventCreateWithFlags(&A->ev_scatter, cudaEventDisableTiming); }
            cudaEventRecord(A->ev_loc, st);
            cudaStreamWaitEvent(ctx->sym_stream, A->ev_loc, 0);
            rs = scatter_pass(ctx, A, ctx->sym_stream);
            cudaEventRecord(A->ev_scatter, ctx->sym_stream);
            A->scatter_in_flight = true;
        }
        A->D.n = keep_n;
        A->scatter_done_eval = rs ? 0 : ctx->eval_id;
    }
    return true;
}
// ---- fused detection + evaluation: the same two steps with the sources' count and descriptors known on the device only ----
bool assembly_locate_ready(sb_context* ctx)
{
    Assembly* A = ctx->assembly;
    static const bool disabled = std::getenv("SB_NO_SCATTER") != nullptr;
    if (disabled || !A || !A->numeric_valid || A->pf_pending || static_part_stale(ctx, A)) return false;
    if (A->nnzb == 0 || A->nnzb >= (1ull << 31) || ctx->H.cap >= (1ull << 32)) return false;
    int bits = 1;
    while ((1ll << bits) < ctx->ndofs / 3 + 1) bits++;
    return bits == A->key_shift;
}
PotDesc* assembly_descs_dev(sb_context* ctx, int n)
{
    Assembly* A = ctx->assembly;
    A->descs.ensure(n + 1);
    return A->descs.p;
}
size_t assembly_scatter_capacity(sb_context* ctx, size_t n_src_estimate)
{
    (void)ctx;
    size_t cap = 1024;
    while (cap < 2 * n_src_estimate) cap <<= 1;
    return cap;
}
int assembly_locate_dynamic_dev(sb_context* ctx, const DynTotals* d_tot, size_t n_src_estimate)
{
    Assembly* A = ctx->assembly;
    cudaStream_t st = ctx->stream;
    if (A->miss_flag_dirty) { SB_CUDA(ctx, cudaMemsetAsync(ctx->d_scalars + 2, 0, sizeof(double), st)); A->miss_flag_dirty = false; }
    const size_t n = std::max<size_t>(n_src_estimate, 1024);
    A->D.src_off.ensure(n + 1); A->D.src_pitch.ensure(n + 1); A->d_final_of_src.ensure(n + 1);
    A->loc_cap = std::min(std::min(A->D.src_off.cap, A->D.src_pitch.cap), A->d_final_of_src.cap);
    k_dyn_locate_src<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(A->descs.p, 0, 0, d_tot, ctx->rows.p, A->rows.p, A->cols.p, A->nbr, A->d_final_of_src.p,
                                                                                           A->D.src_off.p, A->D.src_pitch.p, ctx->d_scalars + 2);
    ctx->launches++;
    // speculative scatter pass on the side stream (capacity from the estimate; the kernel declines if the actual count does not fit)
    if (!A->ev_loc) { cudaEventCreateWithFlags(&A->ev_loc, cudaEventDisableTiming); cudaEventCreateWithFlags(&A->ev_scatter, cudaEventDisableTiming); }
    const size_t cap = assembly_scatter_capacity(ctx, n);
    A->hkeys.ensure(cap); A->hacc.ensure(9 * cap); A->has_dyn.ensure(A->nnzb + 1);
    A->hcap = cap;
    SB_CUDA(ctx, cudaEventRecord(A->ev_loc, st));
    SB_CUDA(ctx, cudaStreamWaitEvent(ctx->sym_stream, A->ev_loc, 0));
    {
        const size_t n_clear = std::max((A->nnzb + 1 + 15) / 16, cap);
        k_scatter_clear<<<(unsigned)((n_clear + 255) / 256), 256, 0, ctx->sym_stream>>>(A->has_dyn.p, A->nnzb + 1, A->hkeys.p, A->hacc.p, cap);
        k_scatter_dynamic<<<(unsigned)std::min<size_t>((n + 31) / 32, 148 * 4), SC_THREADS, 0, ctx->sym_stream>>>(ctx->H.p, A->d_final_of_src.p, A->D.src_off.p, A->D.src_pitch.p, 0, d_tot,
                                                                                                               A->hkeys.p, A->hacc.p, (uint32_t)(cap - 1), A->has_dyn.p);
        ctx->launches += 2;
    }
    SB_CUDA(ctx, cudaEventRecord(A->ev_scatter, ctx->sym_stream));
    A->scatter_in_flight = true;
    A->loc_dynamic = 0;   // (set by assembly_locate_result_dev: the table version is bumped when the tables are published)
    return 0;
}
void assembly_locate_result_dev(sb_context* ctx, bool miss, size_t n_src, bool scatter_valid)
{
    Assembly* A = ctx->assembly;
    A->scatter_done_eval = 0;
    if (miss || n_src > A->loc_cap) { A->n_scatter_misses++; A->miss_flag_dirty = miss; return; }   // (sort-based symbolic phase at the next assembly)
    A->n_scatter_hits++;
    A->D.n = n_src;
    A->built_dynamic = ctx->dynamic_version;
    A->built_n_src = ctx->n_blocks_total;
    A->scatter_mode = true;
    if (scatter_valid && 2 * n_src <= A->hcap) A->scatter_done_eval = ctx->eval_id;
}
// ... and the answer, once the evaluation's scalars are on the host
void assembly_locate_result(sb_context* ctx, bool miss)
{
    Assembly* A = ctx->assembly;
    if (!A || A->loc_dynamic != ctx->dynamic_version) return;
    if (miss) { A->n_scatter_misses++; A->scatter_done_eval = 0; A->miss_flag_dirty = true; return; }   // (the sort-based symbolic phase runs at the next assembly)
    A->n_scatter_hits++;
    A->D.n = A->loc_n;
    A->built_dynamic = ctx->dynamic_version;
    A->built_n_src = ctx->n_blocks_total;
    A->scatter_mode = true;
}

// dynamic contributions of one numeric pass into the FP64 hash table (all sources, every pass: a pass after a PD projection
// re-sums only dirty blocks, and those see the projected elements' new values)
__global__ void k_scatter_clear(uint8_t* __restrict__ has_dyn, size_t n_blocks, int32_t* __restrict__ hkeys, double* __restrict__ hacc, size_t cap)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 16 flags, then (if in range) one key and its nine accumulators per thread
    if (16 * i < n_blocks) {
        if (16 * i + 16 <= n_blocks) *reinterpret_cast<uint4*>(has_dyn + 16 * i) = make_uint4(0, 0, 0, 0);
        else for (size_t b = 16 * i; b < n_blocks; b++) has_dyn[b] = 0;
    }
    if (i < cap) {
        hkeys[i] = -1;
        for (int k = 0; k < 9; k++) hacc[9 * i + k] = 0.0;
    }
}
static int scatter_pass(sb_context* ctx, Assembly* A, cudaStream_t st)
{
    const size_t n_src = A->D.n;
    size_t cap = 1024;
    while (cap < 2 * n_src) cap <<= 1;
    A->hkeys.ensure(cap); A->hacc.ensure(9 * cap); A->has_dyn.ensure(A->nnzb + 1);
    A->hcap = cap;
    {
        const size_t n_clear = std::max((A->nnzb + 1 + 15) / 16, n_src > 0 ? cap : (size_t)0);
        k_scatter_clear<<<(unsigned)((n_clear + 255) / 256), 256, 0, st>>>(A->has_dyn.p, A->nnzb + 1, A->hkeys.p, A->hacc.p, n_src > 0 ? cap : 0);
        ctx->launches++;
    }
    if (n_src > 0) {
        k_scatter_dynamic<<<(unsigned)std::min<size_t>((n_src + 31) / 32, 148 * 4), SC_THREADS, 0, st>>>(ctx->H.p, A->d_final_of_src.p, A->D.src_off.p, A->D.src_pitch.p, n_src, nullptr,
                                                                                                       A->hkeys.p, A->hacc.p, (uint32_t)(cap - 1), A->has_dyn.p);
        ctx->launches++;
    }
    return 0;
}

int assemble_internal(sb_context* ctx)
{
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_assemble: call sb_eval(SB_EVAL_PGH) first");
    Assembly* A = get(ctx);
    if (ctx->n_blocks_total == 0) return fail(ctx, SB_ERR_STATE, "sb_assemble: no element Hessians");
    bool rebuilt = false;
    if (A->scatter_in_flight) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, A->ev_scatter, 0)); A->scatter_in_flight = false; }   // (it reads buffers a rebuild below rewrites)
    if (A->pf_pending) {
        ctx->issuer->wait();
        if (A->pf_failed) { cudaEventSynchronize(A->ev_sym); A->pf_pending = false; }   // (falls back to the in-place build below)
    }
    if (!pattern_current(ctx, A)) {
        int r;
        if (A->pf_pending && A->pf_static == ctx->static_version && A->pf_dynamic == ctx->dynamic_version && A->pf_n_src == ctx->n_blocks_total &&
            !static_part_stale(ctx, A)) {
            StageTimer timer(ctx, ST_ASM_SYMBOLIC);
            SB_CUDA(ctx, cudaEventSynchronize(A->ev_sym));
            r = symbolic_phase_b(ctx, A);
        } else {
            if (A->pf_pending) cudaEventSynchronize(A->ev_sym);   // (a stale prefetch: let it finish before its buffers are reused)
            r = build_symbolic(ctx, A, static_part_stale(ctx, A));
        }
        A->pf_pending = false;
        if (r) return r;
        rebuilt = true;
    } else if (A->pf_pending) {
        cudaEventSynchronize(A->ev_sym);
        A->pf_pending = false;
    }
    StageTimer timer(ctx, ST_ASM_NUMERIC);
    NumericArgs a;
    a.H = ctx->H.p; a.seg4 = A->seg4.p; a.s_off = A->S.sorted_off.p; a.s_pitch = A->S.sorted_pitch.p; a.d_off = A->D.sorted_off.p; a.d_pitch = A->D.sorted_pitch.p;
    a.vals = A->vals.p; a.dirty = A->dirty.p; a.long_blocks = A->long_blocks.p; a.n_long = A->d_counts; a.nnzb = A->nnzb;
    a.scatter = A->scatter_mode ? 1 : 0;
    a.has_dyn = nullptr; a.hkeys = nullptr; a.hacc = nullptr; a.hmask = 0;
    a.own_range = nullptr;
    if (ctx->assemble_own_rows) {
        A->own_range.ensure(2);
        if (dist_own_rows_only(ctx, A->rows.p, A->nbr, A->nnzb, A->own_range.p)) a.own_range = A->own_range.p;
    }
    if (A->scatter_mode) {
        if (A->scatter_done_eval == ctx->eval_id && ctx->n_projected == 0) {
            // the evaluation already ran this pass on a side stream (behind its pattern lookup, under its reductions and its sync;
            // ordered before the context stream above)
        } else {
            int rs = scatter_pass(ctx, A, ctx->stream);
            if (rs) return rs;
        }
        A->scatter_done_eval = 0;
        a.has_dyn = A->has_dyn.p; a.hkeys = A->hkeys.p; a.hacc = A->hacc.p; a.hmask = (uint32_t)(A->hcap - 1);
    }
    const size_t nt = 9 * A->nnzb;
    const unsigned grid = (unsigned)((nt + 287) / 288);
    // the few long blocks (a rigid body's diagonal: thousands of sources, one CTA each, latency-bound) are summed on a side
    // stream while the bulk kernel streams the element Hessians; the two write disjoint blocks
    cudaStream_t side = ctx->side[0];
    const bool with_long = !(A->scatter_mode && A->n_long_static == 0);   // (scatter mode: only statically long blocks are left to it)
    if (with_long) {
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        SB_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev_fork, 0));
    }
    const bool update = !rebuilt && A->numeric_valid && A->assembled_eval == ctx->eval_id;
    if (update) {
        // same evaluation, same pattern: only PD-projected elements changed since the last pass
        if (with_long) k_assemble_long<true><<<LONG_CTAS, 288, 0, side>>>(a);
        k_assemble_numeric<true><<<grid, 288, 0, ctx->stream>>>(a);
    } else {
        if (with_long) k_assemble_long<false><<<LONG_CTAS, 288, 0, side>>>(a);
        k_assemble_numeric<false><<<grid, 288, 0, ctx->stream>>>(a);
    }
    if (with_long) {
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_join[0], side));
        SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0));
    }
    // (the dirty flags are only ever set by a PD projection of this evaluation: nothing to clear after a full pass of a fresh one)
    if (update || ctx->n_projected > 0) { SB_CUDA(ctx, cudaMemsetAsync(A->dirty.p, 0, A->nnzb + 1, ctx->stream)); }
    ctx->launches += with_long ? 2 : 1;
    SB_CUDA(ctx, cudaGetLastError());
    A->assembled_eval = ctx->eval_id;
    A->numeric_valid = true;
    return 0;
}

// for project.cu: where a changed element's blocks land (false when the pattern is stale -> the next assembly is a full one anyway)
bool assembly_dirty_view(sb_context* ctx, DirtyView* v)
{
    Assembly* A = ctx->assembly;
    if (!A || !pattern_current(ctx, A)) return false;
    v->n_static = (unsigned long long)A->S.n;
    v->s_blk_of_src = A->S.blk_of_src.p; v->s_final = A->s_final.p;
    v->d_final_of_src = A->d_final_of_src.p;
    v->dirty = A->dirty.p;
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

// Scatter mode keeps the blocks of contact pairs that are gone as explicit zeros until the next sort-based rebuild; the
// reference drops zero blocks at end_insertion (BlockedSparseMatrix.h:782-895), so the matrix handed out is compacted: blocks
// without a static source that received no dynamic contribution in the last numeric pass are left out.
static int bcsr_kept_blocks(sb_context* ctx, Assembly* A, std::vector<uint32_t>& keep)
{
    keep.clear();
    if (!A->scatter_mode) return 0;
    std::vector<int4> seg(A->nnzb);
    std::vector<uint8_t> dyn(A->nnzb);
    SB_CUDA(ctx, cudaMemcpyAsync(seg.data(), A->seg4.p, sizeof(int4) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(dyn.data(), A->has_dyn.p, A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t b = 0; b < A->nnzb; b++) if (seg[b].y > seg[b].x || dyn[b]) keep.push_back((uint32_t)b);
    return 0;
}

int sb_bcsr_info(sb_context* ctx, int* n_block_rows, int64_t* nnzb)
{
    if (!ctx) return SB_ERR_ARG;
    Assembly* A = ctx->assembly;
    if (!A || !A->numeric_valid) return fail(ctx, SB_ERR_STATE, "sb_bcsr_info: call sb_assemble first");
    if (n_block_rows) *n_block_rows = A->nbr;
    if (nnzb) {
        *nnzb = (int64_t)A->nnzb;
        if (A->scatter_mode) {
            std::vector<uint32_t> keep;
            int r = bcsr_kept_blocks(ctx, A, keep); if (r) return r;
            *nnzb = (int64_t)keep.size();
        }
    }
    return SB_OK;
}

int sb_bcsr_get(sb_context* ctx, int64_t* host_rows, int32_t* host_cols, float* host_vals)
{
    if (!ctx) return SB_ERR_ARG;
    Assembly* A = ctx->assembly;
    if (!A || !A->numeric_valid) return fail(ctx, SB_ERR_STATE, "sb_bcsr_get: call sb_assemble first");
    if (!A->scatter_mode) {
        if (host_rows) SB_CUDA(ctx, cudaMemcpyAsync(host_rows, A->rows.p, sizeof(int64_t) * (A->nbr + 1), cudaMemcpyDeviceToHost, ctx->stream));
        if (host_cols) SB_CUDA(ctx, cudaMemcpyAsync(host_cols, A->cols.p, sizeof(int32_t) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
        if (host_vals) SB_CUDA(ctx, cudaMemcpyAsync(host_vals, A->vals.p, sizeof(float) * 9 * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SB_OK;
    }
    std::vector<uint32_t> keep;
    int r = bcsr_kept_blocks(ctx, A, keep); if (r) return r;
    std::vector<int64_t> rows(A->nbr + 1);
    std::vector<int32_t> cols(A->nnzb);
    std::vector<float> vals(9 * A->nnzb);
    SB_CUDA(ctx, cudaMemcpyAsync(rows.data(), A->rows.p, sizeof(int64_t) * (A->nbr + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(cols.data(), A->cols.p, sizeof(int32_t) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(vals.data(), A->vals.p, sizeof(float) * 9 * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t k = 0;
    int row = 0;
    if (host_rows) host_rows[0] = 0;
    for (size_t i = 0; i < keep.size(); i++) {
        const uint32_t b = keep[i];
        while (row < A->nbr && (int64_t)b >= rows[row + 1]) { row++; if (host_rows) host_rows[row] = (int64_t)k; }
        if (host_cols) host_cols[k] = cols[b];
        if (host_vals) for (int c = 0; c < 9; c++) host_vals[9 * k + c] = vals[9 * (size_t)b + c];
        k++;
    }
    while (row < A->nbr) { row++; if (host_rows) host_rows[row] = (int64_t)k; }
    return SB_OK;
}

}  // extern "C"
