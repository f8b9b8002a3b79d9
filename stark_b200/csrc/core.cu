// Context, arrays, DoF sets, potentials and the evaluation driver behind the C-ABI (include/stark_b200.h).
#include "internal.h"
#include <algorithm>
#include <chrono>
#include <cstring>

namespace sb {

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// SB_TIMELINE=1: no synchronisation; every stage records a CUDA event pair on the context stream and host timestamps, and
// sb_newton_solve prints the timeline of its last call (where the GPU idles between stages, how long the host takes to issue
// a stage).  Diagnostic only.
struct TimelineRec { int stage; double h0, h1; cudaEvent_t e0, e1; };
static std::vector<TimelineRec> g_timeline;
static std::vector<cudaEvent_t> g_event_pool;
static const bool g_timeline_on = std::getenv("SB_TIMELINE") != nullptr;
static cudaEvent_t pool_event()
{
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
bool timeline_enabled() { return g_timeline_on; }
// a completion point on any stream ("everything issued so far on `st` is done at t"), labelled
struct TimelinePoint { const char* label; cudaEvent_t e; double h; };
static std::vector<TimelinePoint> g_points;
static std::mutex g_points_mutex;   // (the helper thread records points too)
void timeline_point(cudaStream_t st, const char* label)
{
    if (!g_timeline_on) return;
    std::lock_guard<std::mutex> lk(g_points_mutex);
    TimelinePoint p; p.label = label; p.e = pool_event(); p.h = now_ms();
    cudaEventRecord(p.e, st);
    g_points.push_back(p);
}
void timeline_mark(sb_context* ctx, int stage)   // a zero-length record (e.g. begin / end of the solve)
{
    if (!g_timeline_on) return;
    TimelineRec r; r.stage = stage; r.h0 = r.h1 = now_ms(); r.e0 = pool_event(); r.e1 = nullptr;
    cudaEventRecord(r.e0, ctx->stream);
    g_timeline.push_back(r);
}
void timeline_dump(sb_context* ctx, const char* const* names)
{
    if (!g_timeline_on || g_timeline.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    if (!names) {   // discard
        for (TimelineRec& r : g_timeline) { g_event_pool.push_back(r.e0); if (r.e1) g_event_pool.push_back(r.e1); }
        g_timeline.clear();
        for (TimelinePoint& q : g_points) g_event_pool.push_back(q.e);
        g_points.clear();
        return;
    }
    cudaDeviceSynchronize();
    for (TimelinePoint& q : g_points) {
        float g = 0.f;
        cudaEventElapsedTime(&g, g_timeline.front().e0, q.e);
        fprintf(stderr, "TIMEPOINT %-44s issued %9.3f done %9.3f\n", q.label, q.h - g_timeline.front().h0, g);
        g_event_pool.push_back(q.e);
    }
    g_points.clear();
    const TimelineRec& base = g_timeline.front();
    double prev_gpu_end = 0.0;
    fprintf(stderr, "TIMELINE %-18s %9s %9s | %9s %9s %9s (ms; gpu gap = idle before the stage's first kernel)\n", "stage", "host_beg", "host_dur", "gpu_beg", "gpu_dur", "gpu_gap");
    for (const TimelineRec& r : g_timeline) {
        float g0 = 0.f, g1 = 0.f;
        cudaEventElapsedTime(&g0, base.e0, r.e0);
        if (r.e1) cudaEventElapsedTime(&g1, base.e0, r.e1); else g1 = g0;
        fprintf(stderr, "TIMELINE %-18s %9.3f %9.3f | %9.3f %9.3f %9.3f\n", r.stage >= 0 ? names[r.stage] : "mark", r.h0 - base.h0, r.h1 - r.h0, g0, g1 - g0, g0 - prev_gpu_end);
        prev_gpu_end = g1;
    }
    for (TimelineRec& r : g_timeline) { g_event_pool.push_back(r.e0); if (r.e1) g_event_pool.push_back(r.e1); }
    g_timeline.clear();
}
StageTimer::StageTimer(sb_context* c, int s) : ctx(c), stage(s), t0(0.0)
{
    if (g_timeline_on) {
        TimelineRec r; r.stage = s; r.h0 = now_ms(); r.h1 = 0.0; r.e0 = pool_event(); r.e1 = pool_event();
        cudaEventRecord(r.e0, ctx->stream);
        g_timeline.push_back(r);
        t0 = (double)(g_timeline.size() - 1);
        return;
    }
    if (!ctx->profile) return;
    cudaStreamSynchronize(ctx->stream);
    t0 = now_ms();
}
StageTimer::~StageTimer()
{
    if (g_timeline_on) {
        TimelineRec& r = g_timeline[(size_t)t0];
        cudaEventRecord(r.e1, ctx->stream);
        r.h1 = now_ms();
        return;
    }
    if (!ctx->profile) return;
    cudaStreamSynchronize(ctx->stream);
    ctx->stage_ms[stage] += now_ms() - t0;
    ctx->stage_calls[stage]++;
}

static const char* const g_stage_names[ST_COUNT] = {"contact_update", "intersections", "eval_pgh", "eval_p", "project_to_pd", "assembly_symbolic", "assembly_numeric", "pcg", "line_search_misc", "cg_iterations", "pcg_kernel_setup", "project_selected", "project_changed", "pcg_cycles_spmv", "pcg_cycles_barrier", "pcg_cycles_reduce", "pcg_cycles_vector",
                                            "tile_pairs_pt", "tile_pairs_ee", "tile_pairs_et", "candidates_pt", "candidates_ee", "candidates_et", "project_sweeps", "pcg_cycles_window"};
const char* const* stage_names() { return g_stage_names; }

int g_tet_grid_cap = 148 * 3;

void Issuer::start(int dev)
{
    device = dev;
    th = std::thread([this] {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(m);
        while (true) {
            cv.wait(lk, [this] { return quit || job; });
            if (quit) return;
            std::function<void()> f = std::move(job);
            job = nullptr;
            lk.unlock();
            f();
            state.store(0, std::memory_order_release);
            lk.lock();
        }
    });
}
void Issuer::post(std::function<void()> f)
{
    wait();
    state.store(1, std::memory_order_release);
    { std::lock_guard<std::mutex> lk(m); job = std::move(f); }
    cv.notify_one();
}
void Issuer::stop()
{
    wait();
    { std::lock_guard<std::mutex> lk(m); quit = true; }
    cv.notify_one();
    if (th.joinable()) th.join();
}

// The hot path's host synchronisations (detection counts, evaluation scalars, projection counts, PCG result): SB_SPIN_SYNC=1
// polls an event instead of calling cudaStreamSynchronize (which may block / yield depending on the context's scheduling flags)
cudaError_t hot_sync(sb_context* ctx)
{
    static const bool spin = std::getenv("SB_SPIN_SYNC") != nullptr;
    if (!spin) return cudaStreamSynchronize(ctx->stream);
    if (!ctx->ev_sync) cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming);
    cudaError_t e = cudaEventRecord(ctx->ev_sync, ctx->stream);
    if (e != cudaSuccess) return e;
    while ((e = cudaEventQuery(ctx->ev_sync)) == cudaErrorNotReady) {}
    return e;
}

// writers of device arrays are ordered behind asynchronous downloads still in flight (sb_array_download_async)
void order_after_async_downloads(sb_context* ctx)
{
    if (!ctx->copy_dev_pending) return;
    cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done, 0);
    ctx->copy_dev_pending = false;
}

int fail(sb_context* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->error = msg;
    return code;
}
int check_cuda(sb_context* ctx, cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return 0;
    return fail(ctx, SB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------------
// deterministic reductions: fixed grid, fixed tree -> bitwise reproducible run to run
// ---------------------------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 296;   // 2 CTAs per SM on a 148-SM B200
constexpr int RED_THREADS = 256;

template<bool ABSMAX>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_stage1(const double* __restrict__ in, size_t n, double* __restrict__ partial)
{
    __shared__ double s[RED_THREADS];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (size_t)RED_BLOCKS * RED_THREADS) {
        const double v = in[i];
        acc = ABSMAX ? fmax(acc, fabs(v)) : acc + v;
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int w = RED_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] = ABSMAX ? fmax(s[threadIdx.x], s[threadIdx.x + w]) : s[threadIdx.x] + s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
template<bool ABSMAX>
__global__ void __launch_bounds__(512) k_reduce_stage2(const double* __restrict__ partial, double* __restrict__ out)
{
    __shared__ double s[512];
    s[threadIdx.x] = (threadIdx.x < RED_BLOCKS) ? partial[threadIdx.x] : 0.0;
    __syncthreads();
    for (int w = 256; w > 0; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] = ABSMAX ? fmax(s[threadIdx.x], s[threadIdx.x + w]) : s[threadIdx.x] + s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}
// energy sum and gradient inf-norm of one P+G+H evaluation in two launches instead of four (same trees, same results: the
// first RED_BLOCKS blocks are the sum's, the others the maximum's)
__global__ void __launch_bounds__(RED_THREADS) k_reduce_pair_stage1(const double* __restrict__ e, size_t ne, const unsigned long long* __restrict__ ne_dev, const double* __restrict__ g, size_t ng, double* __restrict__ partial)
{
    __shared__ double s[RED_THREADS];
    if (ne_dev) ne = (size_t)*ne_dev;   // (fused detection + evaluation: the element count is known on the device only)
    const bool mx = blockIdx.x >= RED_BLOCKS;
    const int b = mx ? blockIdx.x - RED_BLOCKS : blockIdx.x;
    const double* in = mx ? g : e;
    const size_t n = mx ? ng : ne;
    double acc = 0.0;
    for (size_t i = (size_t)b * RED_THREADS + threadIdx.x; i < n; i += (size_t)RED_BLOCKS * RED_THREADS) {
        const double v = in[i];
        acc = mx ? fmax(acc, fabs(v)) : acc + v;
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int w = RED_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] = mx ? fmax(s[threadIdx.x], s[threadIdx.x + w]) : s[threadIdx.x] + s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(mx ? 512 : 0) + b] = s[0];
}
__global__ void __launch_bounds__(512) k_reduce_pair_stage2(const double* __restrict__ partial, double* __restrict__ out)
{
    __shared__ double s[512];
    const bool mx = blockIdx.x == 1;
    s[threadIdx.x] = (threadIdx.x < RED_BLOCKS) ? partial[(mx ? 512 : 0) + threadIdx.x] : 0.0;
    __syncthreads();
    for (int w = 256; w > 0; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] = mx ? fmax(s[threadIdx.x], s[threadIdx.x + w]) : s[threadIdx.x] + s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[mx ? 1 : 0] = s[0];
}
static void reduce_sum_and_absmax(sb_context* ctx, const double* e, size_t ne, const double* g, size_t ng, double* d_out2, const unsigned long long* ne_dev = nullptr)
{
    ctx->scratch.ensure(1024);
    k_reduce_pair_stage1<<<2 * RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(e, ne, ne_dev, g, ng, ctx->scratch.p);
    k_reduce_pair_stage2<<<2, 512, 0, ctx->stream>>>(ctx->scratch.p, d_out2);
    ctx->launches += 2;
}
void reduce_sum(sb_context* ctx, const double* d_in, size_t n, double* d_out)
{
    ctx->scratch.ensure(1024);
    k_reduce_stage1<false><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(d_in, n, ctx->scratch.p);
    k_reduce_stage2<false><<<1, 512, 0, ctx->stream>>>(ctx->scratch.p, d_out);
    ctx->launches += 2;
}
void reduce_absmax(sb_context* ctx, const double* d_in, size_t n, double* d_out)
{
    ctx->scratch.ensure(1024);
    k_reduce_stage1<true><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(d_in, n, ctx->scratch.p + 512);
    k_reduce_stage2<true><<<1, 512, 0, ctx->stream>>>(ctx->scratch.p + 512, d_out);
    ctx->launches += 2;
}

std::vector<int> layout_order(const sb_context* ctx)
{
    std::vector<int> order;
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < (int)ctx->potentials.size(); i++)
            if ((int)ctx->potentials[i].dynamic == pass) order.push_back(i);
    return order;
}

int recompute_dof_offsets(sb_context* ctx)
{
    int off = 0;
    for (auto& s : ctx->dof_sets) {
        s.offset = off;
        const Array& a = ctx->arrays[s.array];
        off += a.n_rows * a.stride;
    }
    ctx->ndofs = off;
    return 0;
}

// Build the per-slot gather table of a potential from its fetch list (device pointers may have moved).
int refresh_slots(sb_context* ctx, Potential& p)
{
    std::vector<FetchSlot> h(p.k->n_in);
    for (auto& s : h) { s.base = nullptr; s.conn_col = -1; s.stride = 0; s.off = 0; s.pad = 0; }
    for (const sb_fetch& f : p.fetch) {
        const Array& a = ctx->arrays[f.array];
        for (int c = 0; c < f.stride; c++) {
            FetchSlot& s = h[f.first_slot + c];
            s.base = a.d.p;
            s.conn_col = f.conn_col;
            s.stride = a.stride;
            s.off = c;
        }
    }
    // upload only when a binding moved (arrays are re-allocated rarely); pageable-source async copies are staged by the
    // driver before the call returns, and slots_host outlives the call anyway
    if (h.size() != p.slots_host.size() || std::memcmp(h.data(), p.slots_host.data(), h.size() * sizeof(FetchSlot)) != 0) {
        p.slots.ensure(h.size());
        p.slots_host = h;
        SB_CUDA(ctx, cudaMemcpyAsync(p.slots.p, p.slots_host.data(), h.size() * sizeof(FetchSlot), cudaMemcpyHostToDevice, ctx->stream));
    }
    // DoF block offsets
    for (int b = 0; b < p.k->nb; b++) p.blocks[b].dof_offset = ctx->dof_sets[p.block_set[b]].offset;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// evaluation
// ---------------------------------------------------------------------------------------------------
// large potentials stay on the context stream; the small ones are spread over the side streams (fork / join by events)
constexpr int SMALL_POTENTIAL = 100000;   // (everything but the volume elements of a large mesh: the per-node inertia terms overlap with them too)

static EvalArgs make_args(sb_context* ctx, Potential& p)
{
    EvalArgs a;
    a.dyn = nullptr;
    a.slots = p.slots.p;
    a.slots_host = p.slots_host.data();
    a.conn = p.conn_ext ? p.conn_ext : p.conn.p;
    a.conn_stride = p.conn_stride;
    a.n_elem = p.n_elem;
    for (int b = 0; b < MAX_BLOCKS; b++) a.blocks[b] = p.blocks[b];
    a.grad = ctx->grad.p;
    a.H = ctx->H.p + p.H_off;
    a.rows = ctx->rows.p + p.rows_off;
    a.E_elem = ctx->E_elem.p + p.E_off;
    a.g_elem = nullptr;
    a.user = p.k->user;
    return a;
}

// grow an element-output buffer; `keep` > 0: the first `keep` entries are already written by kernels in flight (the static
// potentials of a pre-launched evaluation) and must survive -- rare (the buffers carry slack for the dynamic potentials)
template<class T> static void grow_output(sb_context* ctx, DevBuf<T>& b, size_t n, size_t keep)
{
    if (n <= b.cap) return;
    const size_t want = n + n / 16 + (1u << 20);
    if (keep) { cudaDeviceSynchronize(); b.ensure_keep(want, keep, ctx->stream); }
    else b.ensure(want);
}

// P+G+H, first half: everything that does not depend on the contact tables.  Lays out the STATIC potentials (their offsets
// never depend on the dynamic ones, which follow them), clears gradient / projection flags and launches the static potentials'
// kernels.  sb_newton_solve calls this right after a line-search step has been applied and BEFORE the collision detection of
// the trial state: the volume elements' kernel (70 us at the 200k-tet scene) then runs while the host issues the detection,
// instead of after the detection's synchronisation.  eval_internal picks the result up if the state has not changed since.
static int pgh_static_part(sb_context* ctx, bool beside_detection)
{
    assembly_prefetch_drain(ctx);   // (a prefetched symbolic phase still reads the previous evaluation's block rows)
    if (ctx->bulk_pending) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_bulk, 0)); ctx->bulk_pending = false; }   // (an abandoned pre-launch still writing the buffers)
    ctx->pgh_cache_ok = false;
    ctx->have_pgh = false;
    recompute_dof_offsets(ctx);
    if (ctx->ndofs <= 0) return fail(ctx, SB_ERR_STATE, "sb_eval: no degrees of freedom");
    if (ctx->ndofs % 3 != 0) return fail(ctx, SB_ERR_STATE, "sb_eval: ndofs must be divisible by 3");
    size_t H_total = 0, rows_total = 0, E_total = 0, n_blocks = 0;
    for (auto& p : ctx->potentials) {
        if (p.dynamic) continue;
        H_total = (H_total + 15) & ~(size_t)15;   // every potential's Hessian block starts 128 B aligned (bulk stores)
        p.H_off = H_total; p.rows_off = rows_total; p.E_off = E_total;
        const size_t n = p.k->n_dof;
        H_total += (size_t)p.n_elem * n * n;
        rows_total += (size_t)p.n_elem * p.k->nb;
        E_total += (size_t)p.n_elem;
        n_blocks += (size_t)p.n_elem * p.k->nb * p.k->nb;
    }
    ctx->st_H = H_total; ctx->st_rows = rows_total; ctx->st_E = E_total; ctx->st_blocks = n_blocks;
    grow_output(ctx, ctx->E_elem, E_total + 1, 0);
    grow_output(ctx, ctx->H, H_total + 1, 0);
    grow_output(ctx, ctx->rows, rows_total + 1, 0);
    grow_output(ctx, ctx->projected, E_total + 1, 0);
    ctx->grad.ensure(ctx->ndofs);
    SB_CUDA(ctx, cudaMemsetAsync(ctx->grad.p, 0, sizeof(double) * ctx->ndofs, ctx->stream));
    SB_CUDA(ctx, cudaMemsetAsync(ctx->projected.p, 0, ctx->projected.cap, ctx->stream));   // (whole buffer: the dynamic element count is not known yet)
    for (auto& p : ctx->potentials) {
        if (p.dynamic || p.n_elem == 0) continue;
        int r = refresh_slots(ctx, p);
        if (r) return r;
    }
    // the big potentials first: they are the critical path of the evaluation.  (The fork event is recorded BEFORE them: the small
    // potentials on the side streams wait for the memsets above, not for the volume kernel.)
    ctx->st_side_mask = 0;
    bool forked = false;
    int next_side = 0;
    for (auto& p : ctx->potentials)
        if (!p.dynamic && p.n_elem > 0 && (beside_detection || p.n_elem < SMALL_POTENTIAL) && !forked) { SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream)); forked = true; }
    // pre-launched ahead of the collision detection: the big kernels go to their own low-priority stream with a grid that leaves
    // room on every SM (2 instead of 3 CTAs of the volume kernel), so that the detection's kernels -- issued next on the context
    // stream -- run BESIDE them instead of behind them
    bool bulk_used = false;
    g_tet_grid_cap = beside_detection ? 148 * 2 : 148 * 3;
    for (int pass = 0; pass < 2; pass++)
        for (auto& p : ctx->potentials) {
            if (p.dynamic || p.n_elem == 0) continue;
            const bool big = p.n_elem >= SMALL_POTENTIAL;
            if (big != (pass == 0)) continue;
            cudaStream_t st = ctx->stream;
            if (!big) {
                // (measured: one side stream for the pre-launched small potentials -- fewer joins -- is slower than spreading them)
                const int k = next_side++ % sb_context::N_SIDE;
                if (!(ctx->st_side_mask & (1u << k))) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[k], ctx->ev_fork, 0)); ctx->st_side_mask |= 1u << k; }
                st = ctx->side[k];
            } else if (beside_detection) {
                if (!bulk_used) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->bulk_stream, ctx->ev_fork, 0)); bulk_used = true; }
                st = ctx->bulk_stream;
            }
            p.k->launch_pgh(make_args(ctx, p), st);
            ctx->launches++;
            timeline_point(st, p.k->name);
        }
    g_tet_grid_cap = 148 * 3;
    if (bulk_used) { SB_CUDA(ctx, cudaEventRecord(ctx->ev_bulk, ctx->bulk_stream)); ctx->bulk_pending = true; }
    SB_CUDA(ctx, cudaGetLastError());
    ctx->st_next_side = next_side;
    return 0;
}
int eval_prelaunch_static(sb_context* ctx)
{
    // Measured at the 200k-tet scene: 921 it/s with the pre-launch against 934 without (the dynamic potentials are then issued
    // after the detection's synchronisation with no kernel to hide behind, instead of under the volume kernel) -- off unless
    // SB_PRELAUNCH is set; kept for scenes whose detection is long compared with the volume kernel.
    static const bool off = std::getenv("SB_NO_PRELAUNCH") != nullptr;
    if (off) return 0;
    if (ctx->pre_valid && ctx->pre_state == ctx->state_version && ctx->pre_static == ctx->static_version) return 0;   // already under way for this state
    StageTimer timer(ctx, ST_EVAL_PGH);
    int r = pgh_static_part(ctx, true);
    if (r) return r;
    ctx->pre_valid = true;
    ctx->pre_state = ctx->state_version; ctx->pre_static = ctx->static_version;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Fused collision detection + P+G+H evaluation of a line-search trial state: ONE host synchronisation
// ---------------------------------------------------------------------------------------------------
// The plain path is detection -> sync (table sizes) -> contact / friction potentials -> reductions -> sync: between the two syncs
// the GPU waits for the host to size and issue a dozen launches.  Here everything behind the detection is queued at once: a
// one-thread kernel lays the tables' elements out on the device (same formulas as the host layout), the table potentials' kernels
// (one multi-potential launch sized by the previous counts, grid-stride) read their element count and offsets from that layout, the
// assembly's pattern lookup, the speculative scatter pass and the reductions take their counts from it too, and the host
// synchronises once for the detection's counters, the layout totals and the evaluation's scalars.  Whatever does not fit the
// assumptions (buffer overflow, a table whose potential is not in the multi-potential kernel, no assembled pattern yet) falls back
// to the plain path, which redoes the work.
struct DynMeta { int n; int table_nb[MULTI_G_MAX]; int table_ndof[MULTI_G_MAX]; const int* count[MULTI_G_MAX]; };
__global__ void k_dyn_layout(const __grid_constant__ DynMeta M, int table_cap, unsigned long long st_H, unsigned long long st_rows, unsigned long long st_E,
                             unsigned long long cap_H, unsigned long long cap_rows, unsigned long long cap_E,
                             DynLayout* __restrict__ layout, PotDesc* __restrict__ descs, DynTotals* __restrict__ tot)
{
    __shared__ int s_n[MULTI_G_MAX];
    if ((int)threadIdx.x < M.n) s_n[threadIdx.x] = min(*M.count[threadIdx.x], table_cap);   // (the counters' loads in parallel: 35 dependent round trips otherwise)
    __syncthreads();
    if (threadIdx.x != 0) return;
    unsigned long long H = st_H, R = st_rows, E = st_E, blk = 0;
    int nd = 0;
    for (int d = 0; d < M.n; d++) {
        const int n = s_n[d];
        const unsigned long long nd2 = (unsigned long long)M.table_ndof[d] * M.table_ndof[d];
        H = (H + 15ull) & ~15ull;
        layout[d].n_elem = n; layout[d].pad = 0; layout[d].H_off = H; layout[d].rows_off = R; layout[d].E_off = E;
        if (n > 0) {
            PotDesc pd; pd.H_off = H; pd.rows_off = R; pd.blk_off = blk; pd.n_elem = n; pd.nb = M.table_nb[d];
            descs[nd++] = pd;
        }
        H += (unsigned long long)n * nd2;
        R += (unsigned long long)n * M.table_nb[d];
        E += (unsigned long long)n;
        blk += (unsigned long long)n * M.table_nb[d] * M.table_nb[d];
    }
    const bool overflow = H + 1 > cap_H || R + 1 > cap_rows || E + 1 > cap_E || H >= (1ull << 32);
    if (overflow) { for (int d = 0; d < M.n; d++) layout[d].n_elem = 0; nd = 0; blk = 0; }
    tot->E_total = overflow ? st_E : E; tot->H_total = H; tot->rows_total = R; tot->n_dyn_src = blk; tot->n_descs = nd; tot->overflow = overflow ? 1 : 0;
}

// counters (64 ints) | table digests | scalars (3 doubles) | layout totals -> one contiguous block, one device-to-host copy
struct Mailbox { int counters[64]; unsigned long long hash[64]; double scalars[4]; DynTotals totals; };
__global__ void k_pack_mailbox(const int* __restrict__ counters, const unsigned long long* __restrict__ hash, int n_hash, const double* __restrict__ scalars,
                               const DynTotals* __restrict__ tot, Mailbox* __restrict__ out)
{
    const int t = threadIdx.x;
    if (t < 64) out->counters[t] = counters[t];
    if (t < n_hash) out->hash[t] = hash[t];
    if (t < 3) out->scalars[t] = scalars[t];
    if (t == 0) out->totals = *tot;
}

void eval_discard(sb_context* ctx)
{
    ctx->have_pgh = false; ctx->pgh_cache_ok = false; ctx->pre_valid = false;
}

int eval_fused(sb_context* ctx, int* out_intersections, double* out_E, double* out_res, bool* out_done)
{
    *out_done = false;
    // Measured at the 200k-tet scene (tools/ab.py, medians of 3): 1105 it/s fused against 1107 plain -- the host-bound stretch
    // behind the detection (~100 us) turns into ~95 us of queued GPU work (layout kernel, table kernels, lookup, reductions, the
    // packed read-back), so nothing is gained there; kept behind SB_FUSED=1 for scenes with many populated tables, where the
    // plain path's per-table host work grows and this one's does not.
    static const bool on = std::getenv("SB_FUSED") != nullptr;
    if (!on || ctx->profile || !contact_fusable(ctx) || !assembly_locate_ready(ctx)) return 0;
    // the table potentials, in layout (= registration) order; all of them must have a body in the multi-potential kernel or be
    // empty now AND after this detection (checked below)
    std::vector<int> dyn;
    for (int i = 0; i < (int)ctx->potentials.size(); i++) if (ctx->potentials[i].dynamic) dyn.push_back(i);
    if (dyn.empty() || (int)dyn.size() > MULTI_G_MAX) return 0;
    int rc;
    if (!(ctx->pre_valid && ctx->pre_state == ctx->state_version && ctx->pre_static == ctx->static_version)) {
        if ((rc = eval_prelaunch_static(ctx))) return rc;
        if (!ctx->pre_valid) return 0;   // (pre-launch switched off)
    }
    StageTimer timer(ctx, ST_EVAL_PGH);
    cudaStream_t st = ctx->stream;
    if (!ctx->d_dyn_layout) {
        SB_CUDA(ctx, cudaMalloc(&ctx->d_dyn_layout, MULTI_G_MAX * sizeof(DynLayout)));
        SB_CUDA(ctx, cudaMalloc(&ctx->d_dyn_totals, sizeof(DynTotals)));
        SB_CUDA(ctx, cudaMallocHost(&ctx->h_dyn_totals, sizeof(DynTotals)));
        SB_CUDA(ctx, cudaMalloc(&ctx->d_mailbox, sizeof(Mailbox)));
        SB_CUDA(ctx, cudaMallocHost(&ctx->h_mailbox, sizeof(Mailbox)));
    }
    // ---- detection: all launches, counters on their way to the host ----
    if ((rc = contact_fused_issue(ctx))) return rc;
    // ---- layout on the device ----
    DynMeta M;
    M.n = (int)dyn.size();
    int fam_ctas[MULTI_G_FAMILIES] = {0, 0, 0, 0, 0, 0};
    for (int f = 0; f < MULTI_G_FAMILIES; f++) { ctx->multi_g[f].n = 0; ctx->multi_g[f].pad = 0; }
    int table_cap = 0;
    size_t est_src = 0;
    std::vector<int> ineligible;
    std::vector<std::pair<int, EvalArgs>> own;
    for (int d = 0; d < M.n; d++) {
        Potential& p = ctx->potentials[dyn[d]];
        const int32_t* conn = nullptr;
        if (contact_issue_table(ctx, dyn[d], &conn, &M.count[d], &table_cap) < 0) { eval_discard(ctx); return 0; }   // (a dynamic potential that is not a contact table)
        M.table_nb[d] = p.k->nb; M.table_ndof[d] = p.k->n_dof;
        const int est = std::max(64, 2 * p.n_elem);
        est_src += (size_t)est * p.k->nb * p.k->nb;
        // Tables that held elements at the last detection get their own kernel (sized by twice that count, grid-stride); the
        // others -- usually empty again -- share the multi-potential launch with one CTA each.  (All of them in the multi-potential
        // kernel was measured: 49 us against ~10, its code is the sum of 35 differentiated energies and every CTA runs a different one.)
        if ((rc = refresh_slots(ctx, p))) return rc;
        EvalArgs a = make_args(ctx, p);
        a.dyn = ctx->d_dyn_layout + d;
        a.conn = conn;
        a.H = ctx->H.p; a.rows = ctx->rows.p; a.E_elem = ctx->E_elem.p;
        if (p.n_elem > 0) { a.n_elem = est; own.push_back(std::make_pair(dyn[d], a)); continue; }
        const int ctas = (p.k->p_kind >= 0) ? multi_g_ctas(p.k->p_kind, 1) : 0;
        if (ctas <= 0) { ineligible.push_back(dyn[d]); continue; }
        const int fam = multi_g_family(p.k->p_kind);
        MultiGArgs& MG = ctx->multi_g[fam];
        MG.kind[MG.n] = p.k->p_kind; MG.cta0[MG.n] = fam_ctas[fam]; MG.it[MG.n] = a;
        MG.n++;
        fam_ctas[fam] += ctas;
    }
    PotDesc* d_descs = assembly_descs_dev(ctx, M.n);
    k_dyn_layout<<<1, 64, 0, st>>>(M, table_cap, ctx->st_H, ctx->st_rows, ctx->st_E, ctx->H.cap, ctx->rows.cap, std::min(ctx->E_elem.cap, ctx->projected.cap),
                                   ctx->d_dyn_layout, d_descs, ctx->d_dyn_totals);
    // ---- the table potentials in one launch, reading that layout ----
    timeline_point(st, "fused: layout");
    unsigned own_mask = 0;
    if (!own.empty()) {
        SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        int k = 0;
        for (auto& oa : own) {
            const int sidx = k++ % sb_context::N_SIDE;
            if (!(own_mask & (1u << sidx))) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[sidx], ctx->ev_fork, 0)); own_mask |= 1u << sidx; }
            ctx->potentials[oa.first].k->launch_pgh(oa.second, ctx->side[sidx]);
            ctx->launches++;
        }
    }
    for (int f = 0; f < MULTI_G_FAMILIES; f++) if (ctx->multi_g[f].n > 0) { launch_pgh_multi(f, ctx->multi_g[f], fam_ctas[f], st); ctx->launches++; }
    timeline_point(st, "fused: table potentials");
    ctx->launches += 2;
    // ---- joins (static part on the side / bulk streams), pattern lookup + speculative scatter, reductions ----
    if (ctx->bulk_pending) { SB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_bulk, 0)); ctx->bulk_pending = false; }
    for (int k = 0; k < sb_context::N_SIDE; k++)
        if ((ctx->st_side_mask | own_mask) & (1u << k)) {
            SB_CUDA(ctx, cudaEventRecord(ctx->ev_join[k], ctx->side[k]));
            SB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[k], 0));
        }
    timeline_point(st, "fused: joined");
    if ((rc = assembly_locate_dynamic_dev(ctx, ctx->d_dyn_totals, est_src))) return rc;
    timeline_point(st, "fused: located");
    reduce_sum_and_absmax(ctx, ctx->E_elem.p, 0, ctx->grad.p, ctx->ndofs, ctx->d_scalars, &ctx->d_dyn_totals->E_total);
    timeline_point(st, "fused: reduced");
    {   // everything the host needs, in one copy
        const int* d_counters; const unsigned long long* d_hash;
        contact_readback_sources(ctx, &d_counters, &d_hash);
        k_pack_mailbox<<<1, 64, 0, st>>>(d_counters, d_hash, contact_n_digest_words(), ctx->d_scalars, ctx->d_dyn_totals, reinterpret_cast<Mailbox*>(ctx->d_mailbox));
        ctx->launches++;
        SB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mailbox, ctx->d_mailbox, sizeof(Mailbox), cudaMemcpyDeviceToHost, st));
    }
    // ---- the one synchronisation ----
    SB_CUDA(ctx, hot_sync(ctx));
    SB_CUDA(ctx, cudaGetLastError());
    {
        const Mailbox& mb = *reinterpret_cast<const Mailbox*>(ctx->h_mailbox);
        contact_readback_deliver(ctx, mb.counters, mb.hash);
        for (int k = 0; k < 3; k++) ctx->h_scalars[k] = mb.scalars[k];
        *ctx->h_dyn_totals = mb.totals;
    }
    bool retry = false;
    if ((rc = contact_fused_finish(ctx, out_intersections, &retry))) return rc;
    const DynTotals& T = *ctx->h_dyn_totals;
    bool fallback = retry || T.overflow != 0;
    for (int pi : ineligible) if (ctx->potentials[pi].n_elem > 0) fallback = true;
    static const bool fdump = std::getenv("SB_FUSED_DUMP") != nullptr;
    if (fdump) fprintf(stderr, "FUSED retry=%d overflow=%d fallback=%d n_int=%d miss=%g n_src=%llu est_src=%zu E_total=%llu\n", (int)retry, T.overflow, (int)fallback, *out_intersections, ctx->h_scalars[2], T.n_dyn_src, est_src, T.E_total);
    if (fallback) { eval_discard(ctx); return 0; }   // (the plain path redoes detection / evaluation; the gradient already holds these tables' contributions, hence the full discard)
    // ---- host copy of the layout (same formulas as k_dyn_layout and eval_internal) ----
    size_t H_total = ctx->st_H, rows_total = ctx->st_rows, E_total = ctx->st_E, n_blocks = ctx->st_blocks;
    for (int pi : dyn) {
        Potential& p = ctx->potentials[pi];
        H_total = (H_total + 15) & ~(size_t)15;
        p.H_off = H_total; p.rows_off = rows_total; p.E_off = E_total;
        const size_t n = p.k->n_dof;
        H_total += (size_t)p.n_elem * n * n;
        rows_total += (size_t)p.n_elem * p.k->nb;
        E_total += (size_t)p.n_elem;
        n_blocks += (size_t)p.n_elem * p.k->nb * p.k->nb;
    }
    if (E_total != T.E_total || H_total != T.H_total || rows_total != T.rows_total) { eval_discard(ctx); return fail(ctx, SB_ERR_STATE, "eval_fused: device and host layouts differ"); }
    ctx->pre_valid = false;
    ctx->n_hessians = E_total; ctx->n_blocks_total = n_blocks; ctx->n_static_blocks = ctx->st_blocks; ctx->n_rows_total = rows_total; ctx->H_total = H_total;
    ctx->n_projected = 0;
    ctx->eval_id++;
    projector_prepare(ctx);
    ctx->have_pgh = true;
    assembly_locate_result_dev(ctx, ctx->h_scalars[2] != 0.0, (size_t)T.n_dyn_src, true);
    ctx->pgh_cache_ok = true;
    ctx->pgh_state = ctx->state_version; ctx->pgh_dynamic = ctx->dynamic_version; ctx->pgh_static = ctx->static_version;
    ctx->pgh_E = ctx->h_scalars[0]; ctx->pgh_residual = ctx->h_scalars[1];
    *out_E = ctx->h_scalars[0];
    *out_res = ctx->h_scalars[1];
    *out_done = true;
    return 0;
}

int eval_internal(sb_context* ctx, int mode, double* out_E, double* out_grad_inf, bool sync_scalars)
{
    if (mode != SB_EVAL_P && mode != SB_EVAL_PGH) return fail(ctx, SB_ERR_ARG, "sb_eval: unknown mode");
    // The line search evaluates its first trial state with gradient and Hessians (newton.cu): when the trial is accepted, the
    // next iteration's evaluation is this one again -- same DoFs, same tables, element Hessians not yet projected.
    if (mode == SB_EVAL_PGH && sync_scalars && ctx->have_pgh && ctx->pgh_cache_ok && ctx->pgh_state == ctx->state_version &&
        ctx->pgh_dynamic == ctx->dynamic_version && ctx->pgh_static == ctx->static_version) {
        if (out_E) *out_E = ctx->pgh_E;
        if (out_grad_inf) *out_grad_inf = ctx->pgh_residual;
        return 0;
    }
    StageTimer timer(ctx, mode == SB_EVAL_PGH ? ST_EVAL_PGH : ST_EVAL_P);
    static const bool eval_dump = std::getenv("SB_EVAL_DUMP") != nullptr;   // diagnostic: host time of the phases of every evaluation
    const double td0 = eval_dump ? now_ms() : 0.0;
    const bool pre = ctx->pre_valid && ctx->pre_state == ctx->state_version && ctx->pre_static == ctx->static_version;
    if (ctx->pre_valid && !pre) {
        // a pre-launched static part of another state (the trial was rejected before it was evaluated): its kernels are still
        // writing the output buffers -- they are in stream order before anything launched below
        ctx->have_pgh = false;
    }
    ctx->pre_valid = false;
    double td1 = td0;

    if (mode == SB_EVAL_PGH) {
        if (!pre) { int r = pgh_static_part(ctx, false); if (r) return r; }
        td1 = eval_dump ? now_ms() : 0.0;
        // ---- second half: the dynamic potentials (contact / friction tables), laid out behind the static ones ----
        size_t H_total = ctx->st_H, rows_total = ctx->st_rows, E_total = ctx->st_E, n_blocks = ctx->st_blocks;
        for (auto& p : ctx->potentials) {
            if (!p.dynamic) continue;
            H_total = (H_total + 15) & ~(size_t)15;
            p.H_off = H_total; p.rows_off = rows_total; p.E_off = E_total;
            const size_t n = p.k->n_dof;
            H_total += (size_t)p.n_elem * n * n;
            rows_total += (size_t)p.n_elem * p.k->nb;
            E_total += (size_t)p.n_elem;
            n_blocks += (size_t)p.n_elem * p.k->nb * p.k->nb;
        }
        if (E_total + 1 > ctx->projected.cap) {   // (the flags of the new tail must be cleared as well)
            grow_output(ctx, ctx->projected, E_total + 1, ctx->st_E);
            SB_CUDA(ctx, cudaMemsetAsync(ctx->projected.p + ctx->st_E, 0, ctx->projected.cap - ctx->st_E, ctx->stream));
        }
        grow_output(ctx, ctx->E_elem, E_total + 1, ctx->st_E);
        grow_output(ctx, ctx->H, H_total + 1, ctx->st_H);
        grow_output(ctx, ctx->rows, rows_total + 1, ctx->st_rows);
        ctx->n_hessians = E_total;
        ctx->n_blocks_total = n_blocks;
        ctx->n_static_blocks = ctx->st_blocks;
        ctx->n_rows_total = rows_total;
        ctx->H_total = H_total;
        ctx->n_projected = 0;
        ctx->eval_id++;
        projector_prepare(ctx);
        const double te1 = eval_dump ? now_ms() : 0.0;
        unsigned side_mask = ctx->st_side_mask;
        int next_side = ctx->st_next_side;
        bool forked = false;
        unsigned dyn_mask = 0;
        bool dyn_all_small = true;
        // the contact / friction tables go into ONE launch on the context stream (k_eval_pgh_multi); what is not eligible for
        // it (24-DoF elements, very large tables) is launched on its own as before
        static const bool no_multi = std::getenv("SB_NO_MULTI_PGH") != nullptr;   // diagnostic hook
        bool MG = false;   // any table in a multi-potential launch
        int fam_ctas[MULTI_G_FAMILIES] = {0, 0, 0, 0, 0, 0};
        for (int f = 0; f < MULTI_G_FAMILIES; f++) { ctx->multi_g[f].n = 0; ctx->multi_g[f].pad = 0; }
        for (auto& p : ctx->potentials) {
            if (!p.dynamic || p.n_elem == 0) continue;
            int r = refresh_slots(ctx, p);
            if (r) return r;
            if (!no_multi && p.n_elem < SMALL_POTENTIAL && p.k->p_kind >= 0) {
                const int ctas = multi_g_ctas(p.k->p_kind, p.n_elem);
                const int fam = multi_g_family(p.k->p_kind);
                if (ctas > 0 && fam >= 0 && ctx->multi_g[fam].n < MULTI_G_MAX) {
                    MultiGArgs& M = ctx->multi_g[fam];
                    M.kind[M.n] = p.k->p_kind; M.cta0[M.n] = fam_ctas[fam]; M.it[M.n] = make_args(ctx, p);
                    M.n++;
                    fam_ctas[fam] += ctas;
                    MG = true;
                    continue;
                }
            }
            cudaStream_t st = ctx->stream;
            if (p.n_elem < SMALL_POTENTIAL) {
                if (!forked) { SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream)); forked = true; }   // (behind the detection that wrote the tables)
                // (two side streams for the contact / friction tables: this stretch of the evaluation is bound by the HOST's API
                //  calls -- every further stream costs a wait, a record and a join -- not by these few-microsecond kernels)
                const int k = next_side++ % 2;
                if (!(dyn_mask & (1u << k))) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[k], ctx->ev_fork, 0)); dyn_mask |= 1u << k; }
                st = ctx->side[k];
            } else dyn_all_small = false;
            p.k->launch_pgh(make_args(ctx, p), st);
            ctx->launches++;
            timeline_point(st, p.k->name);
        }
        // The symbolic phase of the coming assembly depends on the dynamic potentials' block rows only: the helper thread issues
        // it (its own stream, behind these kernels) while this thread goes on with the reductions.
        if (MG) {
            // one launch per family of tables with elements (contact d_d / rb_rb / rb_d, friction likewise), on the two side streams
            int kf = 0;
            for (int f = 0; f < MULTI_G_FAMILIES; f++) {
                if (ctx->multi_g[f].n == 0) continue;
                if (!forked) { SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream)); forked = true; }
                const int k = kf++ % 2;
                if (!(dyn_mask & (1u << k))) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[k], ctx->ev_fork, 0)); dyn_mask |= 1u << k; }
                launch_pgh_multi(f, ctx->multi_g[f], fam_ctas[f], ctx->side[k]);
                ctx->launches++;
            }
            timeline_point(ctx->stream, "contact / friction tables (one launch per family)");
        }
        const double te2 = eval_dump ? now_ms() : 0.0;
        // (the sort-based symbolic phase is only prefetched when the scatter-mode lookup cannot run: first assembly, stale pattern)
        const bool prefetch = (dyn_mask || MG) && dyn_all_small && !(sync_scalars && assembly_locate_possible(ctx));
        unsigned pf_mask = dyn_mask;
        if (prefetch) {
            for (int k = 0; k < sb_context::N_SIDE; k++)
                if (dyn_mask & (1u << k)) SB_CUDA(ctx, cudaEventRecord(ctx->ev_dyn[k], ctx->side[k]));

        }
        side_mask |= dyn_mask;
        if (ctx->bulk_pending) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_bulk, 0)); ctx->bulk_pending = false; }
        for (int k = 0; k < sb_context::N_SIDE; k++)
            if (side_mask & (1u << k)) {
                SB_CUDA(ctx, cudaEventRecord(ctx->ev_join[k], ctx->side[k]));
                SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0));
            }
        SB_CUDA(ctx, cudaGetLastError());
        timeline_point(ctx->stream, "eval: joined");
        const double te3 = eval_dump ? now_ms() : 0.0;
        // changed contact tables: do all their blocks exist in the assembled pattern?  (answer rides with the scalars below)
        const bool located = sync_scalars && assembly_locate_dynamic(ctx);
        const double te4 = eval_dump ? now_ms() : 0.0;
        reduce_sum_and_absmax(ctx, ctx->E_elem.p, E_total, ctx->grad.p, ctx->ndofs, ctx->d_scalars);
        // several GPUs on one scene: rank 0's gradient, energy and residual are everybody's (the replicas stay bitwise identical)
        { const int rb = dist_bcast_from_root(ctx, ctx->grad.p, ctx->ndofs, ctx->d_scalars, 2); if (rb) return rb; }
        timeline_point(ctx->stream, "eval: reduced");
        if (eval_dump) fprintf(stderr, "EVALDUMP2 layout=%.1f dyn_launch=%.1f join=%.1f locate=%.1f reduce=%.1f us\n", 1e3 * (te1 - td1), 1e3 * (te2 - te1), 1e3 * (te3 - te2), 1e3 * (te4 - te3), 1e3 * (now_ms() - te4));
        ctx->have_pgh = true;
        if (sync_scalars) SB_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->locate_pending = located;
        // issued LAST: the ~20 launches of the symbolic phase take the host a while, and the reductions must already be queued
        // (not when the scatter-mode lookup is under way: the symbolic phase is only needed if that reports a missing block)
        if (prefetch && !located) assembly_prefetch_symbolic(ctx, pf_mask);
    } else {
        // ---- energy only (line-search trials after a backtrack) ----
        recompute_dof_offsets(ctx);
        if (ctx->ndofs <= 0) return fail(ctx, SB_ERR_STATE, "sb_eval: no degrees of freedom");
        size_t E_total = 0;
        for (int pi : layout_order(ctx)) { Potential& p = ctx->potentials[pi]; p.E_off = E_total; E_total += (size_t)p.n_elem; }
        // (a P evaluation overwrites the element energies only; the offsets of Hessians / rows of the last P+G+H stay valid)
        grow_output(ctx, ctx->E_elem, E_total + 1, 0);
        for (auto& p : ctx->potentials) {
            if (p.n_elem == 0) continue;
            int r = refresh_slots(ctx, p);
            if (r) return r;
        }
        int n_small = 0;
        for (auto& p : ctx->potentials) if (p.n_elem > 0 && p.n_elem < SMALL_POTENTIAL) n_small++;
        const bool fork = n_small >= 2;
        bool side_used[sb_context::N_SIDE] = {false, false, false, false};
        if (fork) SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        int next_side = 0;
        // every small potential goes into one multi-potential launch on a side stream
        MultiPArgs M;
        M.n = 0; M.pad = 0;
        int multi_ctas = 0;
        const bool use_multi = n_small >= 2;
        for (auto& p : ctx->potentials) {
            if (p.n_elem == 0) continue;
            if (use_multi && p.n_elem < SMALL_POTENTIAL && p.k->p_kind >= 0 && M.n < MULTI_P_MAX) {
                MultiPItem& it = M.it[M.n++];
                it.slots = p.slots.p; it.conn = p.conn_ext ? p.conn_ext : p.conn.p; it.E_elem = ctx->E_elem.p + p.E_off;
                it.conn_stride = p.conn_stride; it.n_elem = p.n_elem; it.kind = p.k->p_kind; it.cta0 = multi_ctas;
                multi_ctas += multi_p_ctas(p.k->p_kind, p.n_elem);
                continue;
            }
            cudaStream_t st = ctx->stream;
            if (fork && p.n_elem < SMALL_POTENTIAL) {
                const int k = next_side++ % sb_context::N_SIDE;
                if (!side_used[k]) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[k], ctx->ev_fork, 0)); side_used[k] = true; }
                st = ctx->side[k];
            }
            p.k->launch_p(make_args(ctx, p), st);
            ctx->launches++;
        }
        if (M.n > 0) {
            const int k = next_side++ % sb_context::N_SIDE;
            if (!side_used[k]) { SB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[k], ctx->ev_fork, 0)); side_used[k] = true; }
            launch_p_multi(M, multi_ctas, ctx->side[k]);
            ctx->launches++;
        }
        for (int k = 0; k < sb_context::N_SIDE; k++)
            if (side_used[k]) {
                SB_CUDA(ctx, cudaEventRecord(ctx->ev_join[k], ctx->side[k]));
                SB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[k], 0));
            }
        SB_CUDA(ctx, cudaGetLastError());
        reduce_sum(ctx, ctx->E_elem.p, E_total, ctx->d_scalars + 0);
        { const int rb = dist_bcast_from_root(ctx, nullptr, 0, ctx->d_scalars, 1); if (rb) return rb; }
    }
    if (sync_scalars && mode != SB_EVAL_PGH) SB_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    const double td2 = eval_dump ? now_ms() : 0.0;
    const double td3 = td2;
    if (!sync_scalars) ctx->issuer->wait();   // (the helper thread's job reads the potentials: it ends inside this call)
    if (sync_scalars) {
        SB_CUDA(ctx, hot_sync(ctx));
        if (dist_aborted(ctx)) return fail(ctx, SB_ERR_CUDA, "sb_eval: a peer rank did not deliver its broadcast within SB_DIST_TIMEOUT_S (ranks out of step, or a peer failed)");
        ctx->issuer->wait();
        if (eval_dump) fprintf(stderr, "EVALDUMP mode=%d static=%.1f issue=%.1f prefetch=%.1f sync=%.1f us\n", mode, 1e3 * (td1 - td0), 1e3 * (td2 - td1), 1e3 * (td3 - td2), 1e3 * (now_ms() - td3));
        if (out_E) *out_E = ctx->h_scalars[0];
        if (out_grad_inf && mode == SB_EVAL_PGH) *out_grad_inf = ctx->h_scalars[1];
        if (mode == SB_EVAL_PGH && ctx->locate_pending) { assembly_locate_result(ctx, ctx->h_scalars[2] != 0.0); ctx->locate_pending = false; }
        if (mode == SB_EVAL_PGH) {
            ctx->pgh_cache_ok = true;
            ctx->pgh_state = ctx->state_version; ctx->pgh_dynamic = ctx->dynamic_version; ctx->pgh_static = ctx->static_version;
            ctx->pgh_E = ctx->h_scalars[0]; ctx->pgh_residual = ctx->h_scalars[1];
        }
    }
    return 0;
}

// scatter / gather between the flat DoF vector and the DoF arrays
__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_axpy_set(double* __restrict__ dst, const double* __restrict__ base, const double* __restrict__ d, double s, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = base[i] + s * d[i];
}
__global__ void k_axpy(double* __restrict__ y, const double* __restrict__ x, double a, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}
__global__ void k_fill(double* __restrict__ x, double v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void k_scale(double* __restrict__ x, double s, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= s;
}

}  // namespace sb

using namespace sb;

// =====================================================================================================
// C-ABI
// =====================================================================================================
extern "C" {

int sb_create(sb_context** out, int device, void* stream)
{
    if (!out) return SB_ERR_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return SB_ERR_CUDA;  // no CPU fallback
    if (device < 0 || device >= n_dev) return SB_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SB_ERR_CUDA;
    sb_context* ctx = new sb_context();
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
        ctx->own_stream = true;
    }
    for (int k = 0; k < sb_context::N_SIDE; k++) {
        // (high priority: when an SM slot frees up under a large kernel of the context stream, the small potentials' CTAs go first)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&ctx->side[k], cudaStreamNonBlocking, std::getenv("SB_NO_PRIORITY") ? prio_lo : prio_hi) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
        if (cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    }
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->sym_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    for (int k = 0; k < sb_context::N_SIDE; k++)
        if (cudaEventCreateWithFlags(&ctx->ev_dyn[k], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&ctx->bulk_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
        if (cudaEventCreateWithFlags(&ctx->ev_bulk, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    }
    ctx->issuer = new Issuer();
    ctx->issuer->start(device);
    if (cudaMallocHost(&ctx->h_scalars, 64 * sizeof(double)) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    if (cudaMalloc(&ctx->d_scalars, 64 * sizeof(double)) != cudaSuccess) { delete ctx; return SB_ERR_CUDA; }
    cudaMemset(ctx->d_scalars, 0, 64 * sizeof(double));
    {
        static bool preloaded = false;
        if (!preloaded) { preload_eval_kernels(); preload_project_kernels(); preload_assembly_kernels(); preloaded = true; }
    }
    *out = ctx;
    return SB_OK;
}

void sb_destroy(sb_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->issuer) { ctx->issuer->stop(); delete ctx->issuer; ctx->issuer = nullptr; }
    cudaStreamSynchronize(ctx->stream);
    if (ctx->sym_stream) cudaStreamSynchronize(ctx->sym_stream);
    for (auto& r : ctx->host_regions) cudaHostUnregister(r.first);
    ctx->host_regions.clear();
    assembly_destroy(ctx);
    pcg_destroy(ctx);
    dist_destroy(ctx);
    user_kernels_destroy(ctx);
    contact_destroy(ctx);
    projector_destroy(ctx);
    direct_destroy(ctx);
    for (auto& a : ctx->arrays) a.d.release();
    for (auto& p : ctx->potentials) { p.conn.release(); p.slots.release(); }
    ctx->H.release(); ctx->rows.release(); ctx->E_elem.release(); ctx->grad.release(); ctx->du.release();
    ctx->dofs_saved.release(); ctx->scratch.release(); ctx->projected.release();
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    for (int k = 0; k < sb_context::N_SIDE; k++) {
        if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
        if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int k = 0; k < sb_context::N_SIDE; k++) if (ctx->ev_dyn[k]) cudaEventDestroy(ctx->ev_dyn[k]);
    if (ctx->sym_stream) cudaStreamDestroy(ctx->sym_stream);
    if (ctx->bulk_stream) { cudaStreamSynchronize(ctx->bulk_stream); cudaStreamDestroy(ctx->bulk_stream); }
    if (ctx->ev_bulk) cudaEventDestroy(ctx->ev_bulk);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_copy_src); cudaEventDestroy(ctx->ev_copy_done); }
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sb_last_error(const sb_context* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
void* sb_get_stream(sb_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int sb_synchronize(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int64_t sb_launch_count(const sb_context* ctx) { return ctx ? ctx->launches.load() : 0; }

int sb_profile_stages(sb_context* ctx, int enable)
{
    if (!ctx) return SB_ERR_ARG;
    ctx->profile = enable != 0;
    for (int i = 0; i < 32; i++) { ctx->stage_ms[i] = 0.0; ctx->stage_calls[i] = 0; }
    return SB_OK;
}
const char* sb_profile_report(sb_context* ctx)
{
    if (!ctx) return "";
    const char* const* names = g_stage_names;
    ctx->profile_report.clear();
    for (int i = 0; i < ST_COUNT; i++)
        ctx->profile_report += std::string(names[i]) + " " + std::to_string(ctx->stage_ms[i]) + " " + std::to_string(ctx->stage_calls[i]) + "\n";
    return ctx->profile_report.c_str();
}

int sb_array_create(sb_context* ctx, const char* label, int stride, int* out_array)
{
    if (!ctx || stride <= 0 || !out_array) return fail(ctx, SB_ERR_ARG, "sb_array_create: bad argument");
    Array a;
    a.label = label ? label : "";
    a.stride = stride;
    ctx->arrays.push_back(a);
    *out_array = (int)ctx->arrays.size() - 1;
    return SB_OK;
}
static int check_array(sb_context* ctx, int array, const char* where)
{
    if (!ctx) return SB_ERR_ARG;
    if (array < 0 || array >= (int)ctx->arrays.size()) return fail(ctx, SB_ERR_ARG, std::string(where) + ": unknown array handle");
    return 0;
}
int sb_array_upload(sb_context* ctx, int array, const double* host, int n_rows)
{
    int r = check_array(ctx, array, "sb_array_upload"); if (r) return r;
    if (n_rows < 0 || (n_rows > 0 && !host)) return fail(ctx, SB_ERR_ARG, "sb_array_upload: bad argument");
    Array& a = ctx->arrays[array];
    order_after_async_downloads(ctx);
    a.d.ensure((size_t)std::max(n_rows, 1) * a.stride);
    a.n_rows = n_rows;
    ctx->state_version++;
    if (n_rows > 0) {
        // pageable source: the driver stages the data before the call returns; registered (pinned) source: truly asynchronous,
        // the caller keeps the buffer untouched until the next synchronising call (sb_host_register contract)
        SB_CUDA(ctx, cudaMemcpyAsync(a.d.p, host, sizeof(double) * (size_t)n_rows * a.stride, cudaMemcpyHostToDevice, ctx->stream));
        cudaPointerAttributes at;
        const bool pinned = (cudaPointerGetAttributes(&at, host) == cudaSuccess) && at.type == cudaMemoryTypeHost;
        cudaGetLastError();   // an unregistered pointer may leave a sticky-free error code behind on old drivers
        if (pinned) {
            bool ours = false;
            for (auto& r : ctx->host_regions) if ((const char*)host >= (const char*)r.first && (const char*)host < (const char*)r.first + r.second) ours = true;
            if (!ours) SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // pinned by someone else: keep the borrow-for-the-call contract
        }
    }
    return SB_OK;
}
int sb_array_fill(sb_context* ctx, int array, int n_rows, double value)
{
    int r = check_array(ctx, array, "sb_array_fill"); if (r) return r;
    if (n_rows < 0) return fail(ctx, SB_ERR_ARG, "sb_array_fill: bad argument");
    Array& a = ctx->arrays[array];
    order_after_async_downloads(ctx);
    a.d.ensure((size_t)std::max(n_rows, 1) * a.stride);
    a.n_rows = n_rows;
    ctx->state_version++;
    const int n = n_rows * a.stride;
    if (n > 0) {
        if (value == 0.0) SB_CUDA(ctx, cudaMemsetAsync(a.d.p, 0, sizeof(double) * (size_t)n, ctx->stream));
        else { k_fill<<<(n + 255) / 256, 256, 0, ctx->stream>>>(a.d.p, value, n); ctx->launches++; }
    }
    return SB_OK;
}
// state roll of a time step on the device (replaces the host loops of PointDynamics::_on_time_step_accepted,
// S/models/deformables/PointDynamics.cpp:64-78: x0 += dt v1, v0 = v1)
int sb_array_axpy(sb_context* ctx, int y, int x, double alpha, int n_rows)
{
    int r = check_array(ctx, y, "sb_array_axpy"); if (r) return r;
    r = check_array(ctx, x, "sb_array_axpy"); if (r) return r;
    Array& ay = ctx->arrays[y]; Array& ax = ctx->arrays[x];
    if (n_rows < 0 || ay.stride != ax.stride || n_rows > ay.n_rows || n_rows > ax.n_rows) return fail(ctx, SB_ERR_ARG, "sb_array_axpy: shapes do not match");
    order_after_async_downloads(ctx);
    ctx->state_version++;
    const int n = n_rows * ay.stride;
    if (n > 0) { k_axpy<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ay.d.p, ax.d.p, alpha, n); ctx->launches++; }
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}
int sb_array_copy(sb_context* ctx, int dst, int src, int n_rows)
{
    int r = check_array(ctx, dst, "sb_array_copy"); if (r) return r;
    r = check_array(ctx, src, "sb_array_copy"); if (r) return r;
    Array& ad = ctx->arrays[dst]; Array& as = ctx->arrays[src];
    if (n_rows < 0 || ad.stride != as.stride || n_rows > as.n_rows) return fail(ctx, SB_ERR_ARG, "sb_array_copy: shapes do not match");
    order_after_async_downloads(ctx);
    ad.d.ensure((size_t)std::max(n_rows, 1) * ad.stride);
    ad.n_rows = n_rows;
    ctx->state_version++;
    if (n_rows > 0) SB_CUDA(ctx, cudaMemcpyAsync(ad.d.p, as.d.p, sizeof(double) * (size_t)n_rows * ad.stride, cudaMemcpyDeviceToDevice, ctx->stream));
    return SB_OK;
}
// Asynchronous read-back into a REGISTERED (sb_host_register) host buffer: ordered after everything submitted so far, runs on
// the library's copy stream beside whatever is submitted next, and is complete after sb_download_wait (or any synchronising
// download).  Later calls that write arrays are ordered behind it on the device.
int sb_array_download_async(sb_context* ctx, int array, double* host, int n_rows)
{
    int r = check_array(ctx, array, "sb_array_download_async"); if (r) return r;
    Array& a = ctx->arrays[array];
    if (n_rows < 0 || n_rows > a.n_rows || (n_rows > 0 && !host)) return fail(ctx, SB_ERR_ARG, "sb_array_download_async: bad argument");
    bool ours = false;
    for (auto& rg : ctx->host_regions) if ((const char*)host >= (const char*)rg.first && (const char*)host + sizeof(double) * (size_t)n_rows * a.stride <= (const char*)rg.first + rg.second) ours = true;
    if (!ours) return fail(ctx, SB_ERR_ARG, "sb_array_download_async: the host buffer is not registered (sb_host_register)");
    if (n_rows == 0) return SB_OK;
    if (!ctx->copy_stream) {
        SB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        SB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_src, cudaEventDisableTiming));
        SB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming));
    }
    SB_CUDA(ctx, cudaEventRecord(ctx->ev_copy_src, ctx->stream));
    SB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_src, 0));
    SB_CUDA(ctx, cudaMemcpyAsync(host, a.d.p, sizeof(double) * (size_t)n_rows * a.stride, cudaMemcpyDeviceToHost, ctx->copy_stream));
    SB_CUDA(ctx, cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    ctx->copy_dev_pending = true;
    ctx->copy_host_pending = true;
    return SB_OK;
}
int sb_download_wait(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    if (ctx->copy_host_pending) { SB_CUDA(ctx, cudaEventSynchronize(ctx->ev_copy_done)); ctx->copy_host_pending = false; ctx->copy_dev_pending = false; }
    return SB_OK;
}
int sb_host_register(sb_context* ctx, void* host, uint64_t bytes)
{
    if (!ctx || !host || bytes == 0) return fail(ctx, SB_ERR_ARG, "sb_host_register: bad argument");
    SB_CUDA(ctx, cudaHostRegister(host, (size_t)bytes, cudaHostRegisterDefault));
    ctx->host_regions.push_back({host, (size_t)bytes});
    return SB_OK;
}
int sb_host_unregister(sb_context* ctx, void* host)
{
    if (!ctx || !host) return fail(ctx, SB_ERR_ARG, "sb_host_unregister: bad argument");
    for (size_t i = 0; i < ctx->host_regions.size(); i++)
        if (ctx->host_regions[i].first == host) {
            SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            SB_CUDA(ctx, cudaHostUnregister(host));
            ctx->host_regions.erase(ctx->host_regions.begin() + i);
            return SB_OK;
        }
    return fail(ctx, SB_ERR_ARG, "sb_host_unregister: not a registered buffer");
}
int sb_array_download(sb_context* ctx, int array, double* host, int n_rows)
{
    int r = check_array(ctx, array, "sb_array_download"); if (r) return r;
    Array& a = ctx->arrays[array];
    if (n_rows < 0 || n_rows > a.n_rows || (n_rows > 0 && !host)) return fail(ctx, SB_ERR_ARG, "sb_array_download: bad size");
    if (n_rows > 0) {
        SB_CUDA(ctx, cudaMemcpyAsync(host, a.d.p, sizeof(double) * (size_t)n_rows * a.stride, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SB_OK;
}
int sb_array_rows(sb_context* ctx, int array, int* out_rows)
{
    int r = check_array(ctx, array, "sb_array_rows"); if (r) return r;
    if (out_rows) *out_rows = ctx->arrays[array].n_rows;
    return SB_OK;
}

int sb_dof_add(sb_context* ctx, int array, int* out_set)
{
    int r = check_array(ctx, array, "sb_dof_add"); if (r) return r;
    if (ctx->arrays[array].stride != 3) return fail(ctx, SB_ERR_ARG, "sb_dof_add: expected DoF stride of 3");
    ctx->dof_sets.push_back({array, 0});
    if (out_set) *out_set = (int)ctx->dof_sets.size() - 1;
    recompute_dof_offsets(ctx);
    return SB_OK;
}
int sb_dof_total(sb_context* ctx, int* out_ndofs)
{
    if (!ctx) return SB_ERR_ARG;
    recompute_dof_offsets(ctx);
    if (out_ndofs) *out_ndofs = ctx->ndofs;
    return SB_OK;
}
int sb_dofs_get(sb_context* ctx, double* host_u)
{
    if (!ctx || !host_u) return fail(ctx, SB_ERR_ARG, "sb_dofs_get: bad argument");
    recompute_dof_offsets(ctx);
    for (auto& s : ctx->dof_sets) {
        Array& a = ctx->arrays[s.array];
        const size_t n = (size_t)a.n_rows * a.stride;
        if (n) SB_CUDA(ctx, cudaMemcpyAsync(host_u + s.offset, a.d.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int sb_dofs_set(sb_context* ctx, const double* host_u)
{
    if (!ctx || !host_u) return fail(ctx, SB_ERR_ARG, "sb_dofs_set: bad argument");
    order_after_async_downloads(ctx);
    ctx->state_version++;
    recompute_dof_offsets(ctx);
    for (auto& s : ctx->dof_sets) {
        Array& a = ctx->arrays[s.array];
        const size_t n = (size_t)a.n_rows * a.stride;
        if (n) SB_CUDA(ctx, cudaMemcpyAsync(a.d.p, host_u + s.offset, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

const char* sb_kernel_names(void)
{
    static std::string names;
    if (names.empty())
        for (const auto& k : all_kernels()) { names += k.name; names += "\n"; }
    return names.c_str();
}

int sb_potential_create(sb_context* ctx, const char* kernel_name, int conn_stride, const sb_fetch* fetch, int n_fetch, int* out_potential)
{
    if (!ctx || !kernel_name || !fetch || n_fetch <= 0 || conn_stride <= 0) return fail(ctx, SB_ERR_ARG, "sb_potential_create: bad argument");
    const KernelInfo* k = find_kernel(kernel_name);
    if (!k) return fail(ctx, SB_ERR_NO_KERNEL, std::string("sb_potential_create: no kernel named '") + kernel_name + "'");
    return sb::potential_create_with_kernel(ctx, k, kernel_name, conn_stride, fetch, n_fetch, out_potential);
}
}  // extern "C"

// the checks and bookkeeping shared by built-in kernels and generated ones (user.cu)
int sb::potential_create_with_kernel(sb_context* ctx, const KernelInfo* k, const char* kernel_name, int conn_stride, const sb_fetch* fetch, int n_fetch, int* out_potential)
{
    if (!ctx || !k || !kernel_name || !fetch || n_fetch <= 0 || conn_stride <= 0) return fail(ctx, SB_ERR_ARG, "sb_potential_create: bad argument");
    Potential p;
    p.k = k;
    p.name = kernel_name;
    p.conn_stride = conn_stride;
    std::vector<int> covered(k->n_in, 0);
    for (int i = 0; i < n_fetch; i++) {
        const sb_fetch& f = fetch[i];
        if (f.array < 0 || f.array >= (int)ctx->arrays.size()) return fail(ctx, SB_ERR_ARG, "sb_potential_create: unknown array in fetch table");
        if (f.stride != ctx->arrays[f.array].stride) return fail(ctx, SB_ERR_LAYOUT, "sb_potential_create: fetch stride differs from the array stride");
        if (f.conn_col < -1 || f.conn_col >= conn_stride) return fail(ctx, SB_ERR_LAYOUT, "sb_potential_create: connectivity column out of range");
        if (f.first_slot < 0 || f.first_slot + f.stride > k->n_in) return fail(ctx, SB_ERR_LAYOUT, std::string("sb_potential_create: in[] slot out of range for ") + kernel_name);
        for (int c = 0; c < f.stride; c++) covered[f.first_slot + c]++;
        p.fetch.push_back(f);
    }
    for (int s = 0; s < k->n_in; s++)
        if (covered[s] != 1) return fail(ctx, SB_ERR_LAYOUT, std::string("sb_potential_create: in[] slots of ") + kernel_name + " are not covered exactly once (expected " + std::to_string(k->n_in) + " inputs)");
    // DoF blocks: the kernel's DOF_SLOT[b] must be bound to a DoF array, in DoF-set order then slot order
    int prev_set = -1, prev_slot = -1;
    for (int b = 0; b < k->nb; b++) {
        const int slot = k->dof_slot[b];
        const sb_fetch* src = nullptr;
        for (const auto& f : p.fetch)
            if (f.first_slot == slot && f.stride == 3) src = &f;
        if (!src) return fail(ctx, SB_ERR_LAYOUT, std::string("sb_potential_create: DoF block slot not bound by a stride-3 fetch in ") + kernel_name);
        int set = -1;
        for (int s = 0; s < (int)ctx->dof_sets.size(); s++)
            if (ctx->dof_sets[s].array == src->array) set = s;
        if (set < 0) return fail(ctx, SB_ERR_LAYOUT, std::string("sb_potential_create: DoF slot of ") + kernel_name + " is bound to an array that is not a DoF set");
        if (src->conn_col < 0) return fail(ctx, SB_ERR_LAYOUT, "sb_potential_create: DoF fetch must be connectivity-indexed");
        if (set < prev_set || (set == prev_set && slot < prev_slot)) return fail(ctx, SB_ERR_LAYOUT, std::string("sb_potential_create: DoF block order of ") + kernel_name + " differs from the reference's (DoF set, then slot)");
        prev_set = set; prev_slot = slot;
        p.block_set[b] = set;
        p.blocks[b].conn_col = src->conn_col;
        p.blocks[b].dof_offset = 0;
    }
    // every other binding of a DoF array would be a DoF the kernel does not differentiate
    for (const auto& f : p.fetch) {
        bool is_dof_array = false;
        for (const auto& s : ctx->dof_sets) if (s.array == f.array) is_dof_array = true;
        if (!is_dof_array) continue;
        bool known = false;
        for (int b = 0; b < k->nb; b++) if (k->dof_slot[b] == f.first_slot) known = true;
        if (!known) return fail(ctx, SB_ERR_LAYOUT, std::string("sb_potential_create: ") + kernel_name + " binds a DoF array at a slot the kernel treats as constant");
    }
    for (int b = k->nb; b < MAX_BLOCKS; b++) { p.blocks[b].conn_col = 0; p.blocks[b].dof_offset = 0; p.block_set[b] = 0; }
    ctx->potentials.push_back(std::move(p));
    if (out_potential) *out_potential = (int)ctx->potentials.size() - 1;
    ctx->static_version++;
    return SB_OK;
}

extern "C" {

static int check_pot(sb_context* ctx, int pot, const char* where)
{
    if (!ctx) return SB_ERR_ARG;
    if (pot < 0 || pot >= (int)ctx->potentials.size()) return fail(ctx, SB_ERR_ARG, std::string(where) + ": unknown potential handle");
    return 0;
}
int sb_potential_set_connectivity(sb_context* ctx, int potential, const int32_t* conn, int n_elements)
{
    int r = check_pot(ctx, potential, "sb_potential_set_connectivity"); if (r) return r;
    if (n_elements < 0 || (n_elements > 0 && !conn)) return fail(ctx, SB_ERR_ARG, "sb_potential_set_connectivity: bad argument");
    Potential& p = ctx->potentials[potential];
    if (p.conn_ext) return fail(ctx, SB_ERR_STATE, "sb_potential_set_connectivity: connectivity of this potential is owned by the contact module");
    p.conn.ensure((size_t)std::max(n_elements, 1) * p.conn_stride);
    p.n_elem = n_elements;
    if (n_elements > 0) {
        SB_CUDA(ctx, cudaMemcpyAsync(p.conn.p, conn, sizeof(int32_t) * (size_t)n_elements * p.conn_stride, cudaMemcpyHostToDevice, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->static_version++;
    ctx->have_pgh = false;
    return SB_OK;
}
int sb_potential_info(sb_context* ctx, int potential, int* n_in, int* n_dofs, int* n_elements)
{
    int r = check_pot(ctx, potential, "sb_potential_info"); if (r) return r;
    const Potential& p = ctx->potentials[potential];
    if (n_in) *n_in = p.k->n_in;
    if (n_dofs) *n_dofs = p.k->n_dof;
    if (n_elements) *n_elements = p.n_elem;
    return SB_OK;
}

int sb_eval(sb_context* ctx, int mode, double* out_E, double* out_grad_inf)
{
    if (!ctx) return SB_ERR_ARG;
    return eval_internal(ctx, mode, out_E, out_grad_inf, true);
}
int sb_eval_prelaunch(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    return eval_prelaunch_static(ctx);
}
int sb_grad_get(sb_context* ctx, double* host_grad)
{
    if (!ctx || !host_grad) return fail(ctx, SB_ERR_ARG, "sb_grad_get: bad argument");
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_grad_get: no gradient evaluated yet");
    SB_CUDA(ctx, cudaMemcpyAsync(host_grad, ctx->grad.p, sizeof(double) * ctx->ndofs, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// [E | grad | hess] per element, gradient recomputed from a dedicated pass so that it is the element's own share
int sb_potential_get_element_output(sb_context* ctx, int potential, double* host_sol)
{
    int r = check_pot(ctx, potential, "sb_potential_get_element_output"); if (r) return r;
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_potential_get_element_output: call sb_eval(SB_EVAL_PGH) first");
    Potential& p = ctx->potentials[potential];
    if (p.n_elem == 0) return SB_OK;
    const int n = p.k->n_dof, n_out = 1 + n + n * n;
    // re-run this potential alone with a per-element gradient sink (values identical, the flat gradient gets a scratch copy)
    DevBuf<double> g_elem, grad_tmp;
    g_elem.ensure((size_t)p.n_elem * n);
    grad_tmp.ensure(ctx->ndofs);
    SB_CUDA(ctx, cudaMemsetAsync(grad_tmp.p, 0, sizeof(double) * ctx->ndofs, ctx->stream));
    r = refresh_slots(ctx, p); if (r) return r;
    EvalArgs a;
    a.dyn = nullptr;
    a.slots = p.slots.p;
    a.slots_host = p.slots_host.data();
    a.conn = p.conn_ext ? p.conn_ext : p.conn.p;
    a.conn_stride = p.conn_stride;
    a.n_elem = p.n_elem;
    for (int b = 0; b < MAX_BLOCKS; b++) a.blocks[b] = p.blocks[b];
    // outputs go to scratch buffers: the live element Hessians (possibly PD-projected, with dirty flags and an assembled
    // matrix that refer to them) are not touched by this diagnostic call
    DevBuf<double> H_tmp, E_tmp;
    DevBuf<int32_t> rows_tmp;
    H_tmp.ensure((size_t)p.n_elem * n * n + 32);   // (+ slack: the bulk stores of the tet kernel want 128-byte aligned tiles)
    E_tmp.ensure(p.n_elem);
    rows_tmp.ensure((size_t)p.n_elem * p.k->nb);
    a.grad = grad_tmp.p;
    a.H = H_tmp.p;
    a.rows = rows_tmp.p;
    a.E_elem = E_tmp.p;
    a.g_elem = g_elem.p;
    a.user = p.k->user;
    p.k->launch_pgh(a, ctx->stream);
    ctx->launches++;
    std::vector<double> E(p.n_elem), g((size_t)p.n_elem * n), H((size_t)p.n_elem * n * n);
    SB_CUDA(ctx, cudaMemcpyAsync(E.data(), E_tmp.p, E.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(g.data(), g_elem.p, g.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(H.data(), H_tmp.p, H.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    H_tmp.release(); E_tmp.release(); rows_tmp.release();
    for (int e = 0; e < p.n_elem; e++) {
        double* o = host_sol + (size_t)e * n_out;
        o[0] = E[e];
        std::memcpy(o + 1, g.data() + (size_t)e * n, n * sizeof(double));
        std::memcpy(o + 1 + n, H.data() + (size_t)e * n * n, (size_t)n * n * sizeof(double));
    }
    g_elem.release(); grad_tmp.release();
    return SB_OK;
}
int sb_potential_get_block_rows(sb_context* ctx, int potential, int32_t* host_rows)
{
    int r = check_pot(ctx, potential, "sb_potential_get_block_rows"); if (r) return r;
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_potential_get_block_rows: call sb_eval(SB_EVAL_PGH) first");
    Potential& p = ctx->potentials[potential];
    if (p.n_elem == 0) return SB_OK;
    SB_CUDA(ctx, cudaMemcpyAsync(host_rows, ctx->rows.p + p.rows_off, sizeof(int32_t) * (size_t)p.n_elem * p.k->nb, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int sb_potential_get_hessians(sb_context* ctx, int potential, double* host_hessians)
{
    int r = check_pot(ctx, potential, "sb_potential_get_hessians"); if (r) return r;
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_potential_get_hessians: call sb_eval(SB_EVAL_PGH) first");
    Potential& p = ctx->potentials[potential];
    if (p.n_elem == 0) return SB_OK;
    const size_t n = p.k->n_dof;
    SB_CUDA(ctx, cudaMemcpyAsync(host_hessians, ctx->H.p + p.H_off, sizeof(double) * (size_t)p.n_elem * n * n, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// ---- line-search support ----
int sb_dofs_save(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    recompute_dof_offsets(ctx);
    ctx->dofs_saved.ensure(ctx->ndofs);
    for (auto& s : ctx->dof_sets) {
        Array& a = ctx->arrays[s.array];
        const size_t n = (size_t)a.n_rows * a.stride;
        if (n) SB_CUDA(ctx, cudaMemcpyAsync(ctx->dofs_saved.p + s.offset, a.d.p, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return SB_OK;
}
int sb_dofs_apply_step(sb_context* ctx, double step)
{
    if (!ctx) return SB_ERR_ARG;
    if (!ctx->du.p || !ctx->dofs_saved.p) return fail(ctx, SB_ERR_STATE, "sb_dofs_apply_step: no saved DoFs / direction");
    order_after_async_downloads(ctx);
    ctx->state_version++;
    for (auto& s : ctx->dof_sets) {
        Array& a = ctx->arrays[s.array];
        const int n = a.n_rows * a.stride;
        if (n) { k_axpy_set<<<(n + 255) / 256, 256, 0, ctx->stream>>>(a.d.p, ctx->dofs_saved.p + s.offset, ctx->du.p + s.offset, step, n); ctx->launches++; }
    }
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}
int sb_du_scale(sb_context* ctx, double factor)
{
    if (!ctx || !ctx->du.p) return fail(ctx, SB_ERR_STATE, "sb_du_scale: no direction");
    k_scale<<<(ctx->ndofs + 255) / 256, 256, 0, ctx->stream>>>(ctx->du.p, factor, ctx->ndofs);
    ctx->launches++;
    SB_CUDA(ctx, cudaGetLastError());
    return SB_OK;
}
int sb_du_get(sb_context* ctx, double* host_du)
{
    if (!ctx || !host_du || !ctx->du.p) return fail(ctx, SB_ERR_STATE, "sb_du_get: no direction");
    SB_CUDA(ctx, cudaMemcpyAsync(host_du, ctx->du.p, sizeof(double) * ctx->ndofs, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

}  // extern "C"

extern "C" int sb_profile_potential(sb_context* ctx, int potential, int mode, int reps, double* out_avg_ms)
{
    int r = check_pot(ctx, potential, "sb_profile_potential"); if (r) return r;
    if (reps <= 0 || !out_avg_ms) return fail(ctx, SB_ERR_ARG, "sb_profile_potential: bad argument");
    if (!ctx->have_pgh) return fail(ctx, SB_ERR_STATE, "sb_profile_potential: call sb_eval(SB_EVAL_PGH) first (output buffers)");
    Potential& p = ctx->potentials[potential];
    if (p.n_elem == 0) { *out_avg_ms = 0.0; return SB_OK; }
    r = refresh_slots(ctx, p); if (r) return r;
    EvalArgs a;
    a.dyn = nullptr;
    a.slots = p.slots.p;
    a.slots_host = p.slots_host.data();
    a.conn = p.conn_ext ? p.conn_ext : p.conn.p;
    a.conn_stride = p.conn_stride;
    a.n_elem = p.n_elem;
    for (int b = 0; b < MAX_BLOCKS; b++) a.blocks[b] = p.blocks[b];
    DevBuf<double> grad_tmp;   // scratch gradient so that the solver state is not disturbed
    grad_tmp.ensure(ctx->ndofs);
    SB_CUDA(ctx, cudaMemsetAsync(grad_tmp.p, 0, sizeof(double) * ctx->ndofs, ctx->stream));
    // scratch outputs: the solver's element Hessians / block rows / energies stay as the last evaluation left them
    DevBuf<double> H_tmp, E_tmp;
    DevBuf<int32_t> rows_tmp;
    const size_t nd = p.k->n_dof;
    H_tmp.ensure((size_t)p.n_elem * nd * nd + 32);
    E_tmp.ensure(p.n_elem);
    rows_tmp.ensure((size_t)p.n_elem * p.k->nb);
    a.grad = grad_tmp.p;
    a.H = H_tmp.p;
    a.rows = rows_tmp.p;
    a.E_elem = E_tmp.p;
    a.g_elem = nullptr;
    a.user = p.k->user;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < reps; i++) {
        if (mode == SB_EVAL_PGH) p.k->launch_pgh(a, ctx->stream); else p.k->launch_p(a, ctx->stream);
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    grad_tmp.release(); H_tmp.release(); E_tmp.release(); rows_tmp.release();
    SB_CUDA(ctx, cudaGetLastError());
    *out_avg_ms = (double)ms / reps;
    return SB_OK;
}
