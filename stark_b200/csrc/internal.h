// Internal state of a stark_b200 context (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>
#include "../../include/stark_b200.h"

namespace sb {

constexpr int MAX_BLOCKS = 8;     // 3-DoF blocks per element (reference potentials use at most 8)
constexpr int EVAL_THREADS = 256;

// One in[] slot of a potential: value = base[row * stride + off], row = conn[e * conn_stride + conn_col] (or 0)
struct FetchSlot {
    const double* base;
    int32_t conn_col;
    int32_t stride;
    int32_t off;
    int32_t pad;
};

struct DofBlock {
    int32_t dof_offset;  // offset of the DoF set in the flat DoF vector
    int32_t conn_col;    // connectivity column that indexes the set
};

// Element layout of the dynamic potentials computed ON THE DEVICE (fused detection + evaluation, core.cu): the contact / friction
// tables' sizes are only on the device when their potentials' kernels are queued, so those kernels read their element count and
// their offsets in the element-output buffers from here.
struct DynLayout {
    int32_t n_elem;
    int32_t pad;
    unsigned long long H_off, rows_off, E_off;   // offsets (in elements of the respective buffer) from the buffers' bases
};
// descriptor of one potential's element blocks for the assembly's key / lookup kernels (assembly.cu)
struct PotDesc {
    unsigned long long H_off;     // offset of the potential's element Hessians in ctx->H
    unsigned long long rows_off;  // offset of its block rows in ctx->rows
    unsigned long long blk_off;   // offset of its element blocks in the source numbering of its class (static / dynamic)
    int n_elem, nb;
};
// totals written by the device-side layout kernel (read back with the evaluation's scalars)
struct DynTotals {
    unsigned long long E_total, H_total, rows_total, n_dyn_src;
    int32_t n_descs;      // dynamic potentials with elements
    int32_t overflow;     // the element-output buffers (or the scatter table) are too small for this layout: nothing was evaluated
};

struct EvalArgs {
    const DynLayout* dyn;          // non-null: n_elem / H / rows / E_elem come from the device-side layout (H, rows, E_elem are then the buffers' bases)
    const FetchSlot* slots;
    const FetchSlot* slots_host;   // HOST copy of the same table (for launchers that pass it as a kernel parameter)
    const int32_t* conn;
    int32_t conn_stride;
    int32_t n_elem;
    DofBlock blocks[MAX_BLOCKS];
    double* grad;      // flat gradient (atomic accumulation)
    double* H;         // element Hessians of this potential, n*n per element, row-major
    int32_t* rows;     // global block rows, NB per element
    double* E_elem;    // per-element energy
    double* g_elem;    // optional per-element gradient (n per element) for parity dumps, may be null
    const void* user;  // generated kernel of a user potential (user.cu), else null
};

struct KernelInfo {
    const char* name;
    int n_in, n_dof, nb;
    const int* dof_slot;
    void (*launch_pgh)(const EvalArgs&, cudaStream_t);
    void (*launch_p)(const EvalArgs&, cudaStream_t);
    int p_kind;     // index into the multi-potential energy kernel's dispatch (-1: own kernel only)
    const void* user = nullptr;   // user.cu: the generated kernel behind launch_pgh / launch_p
};
// one launch for the energy-only evaluation of many small potentials (eval.cu)
struct MultiPItem { const FetchSlot* slots; const int32_t* conn; double* E_elem; int conn_stride, n_elem, kind, cta0; };
constexpr int MULTI_P_MAX = 64;
struct MultiPArgs { int n; int pad; MultiPItem it[MULTI_P_MAX]; };
// one launch for the P+G+H evaluation of the contact / friction tables (eval.cu)
constexpr int MULTI_G_MAX = 40;
constexpr int MULTI_G_FAMILIES = 6;   // contact d_d / rb_rb / rb_d, friction d_d / rb_rb / rb_d: one kernel each
struct MultiGArgs { int n; int pad; int kind[MULTI_G_MAX]; int cta0[MULTI_G_MAX]; EvalArgs it[MULTI_G_MAX]; };
int multi_g_ctas(int p_kind, int n_elem);
int multi_g_family(int p_kind);
void launch_pgh_multi(int family, const MultiGArgs& M, int total_ctas, cudaStream_t s);
int multi_p_ctas(int p_kind, int n_elem);
void launch_p_multi(const MultiPArgs& M, int total_ctas, cudaStream_t s);
const KernelInfo* find_kernel(const char* name);
// load the kernels' code now instead of at their first launch (called once from sb_create)
void preload_eval_kernels();
void preload_project_kernels();
void projector_prepare(sb_context* ctx);
void preload_assembly_kernels();
void assembly_prefetch_symbolic(sb_context* ctx, unsigned side_mask);   // assembly.cu: symbolic phase ahead of time, issued by the helper thread behind ev_dyn[k] of the side streams in the mask
void assembly_prefetch_drain(sb_context* ctx);
bool assembly_locate_possible(sb_context* ctx);
bool assembly_locate_dynamic(sb_context* ctx);              // assembly.cu, scatter mode: can the current pattern absorb the changed contact tables?
void assembly_locate_result(sb_context* ctx, bool miss);
const std::vector<KernelInfo>& all_kernels();

template<class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    void ensure(size_t n)
    {
        if (n <= cap) return;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap = ncap + ncap / 2 + 256;
        T* np = nullptr;
        cudaMalloc(&np, ncap * sizeof(T));
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
    }
    void ensure_keep(size_t n, size_t keep, cudaStream_t s)
    {
        if (n <= cap) return;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap = ncap + ncap / 2 + 256;
        T* np = nullptr;
        cudaMalloc(&np, ncap * sizeof(T));
        if (p) {
            if (keep) cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
            cudaStreamSynchronize(s);
            cudaFree(p);
        }
        p = np;
        cap = ncap;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Array {
    std::string label;
    int stride = 0;
    int n_rows = 0;
    DevBuf<double> d;
};

struct DofSet {
    int array;
    int offset;  // in the flat DoF vector (recomputed when sizes change)
};

struct Potential {
    const KernelInfo* k = nullptr;
    std::string name;
    int conn_stride = 0;
    std::vector<sb_fetch> fetch;
    DofBlock blocks[MAX_BLOCKS];
    int block_set[MAX_BLOCKS];  // DoF set of each block
    int n_elem = 0;
    DevBuf<int32_t> conn;       // owned connectivity (host-provided)
    const int32_t* conn_ext = nullptr;  // or device-resident table owned by the contact module
    bool dynamic = false;               // element set changes at every evaluation (contact / friction tables)
    const int32_t* n_elem_dev = nullptr;
    DevBuf<FetchSlot> slots;
    std::vector<FetchSlot> slots_host;  // what `slots` holds on the device (re-uploaded only when a binding moved)
    // offsets into the shared element-output buffers (recomputed each evaluation)
    size_t H_off = 0, rows_off = 0, E_off = 0;
};

// Stage profiling (sb_profile_stages): host wall time between two stream synchronisations, accumulated per stage.
enum Stage { ST_CONTACT_UPDATE, ST_INTERSECTIONS, ST_EVAL_PGH, ST_EVAL_P, ST_PROJECT, ST_ASM_SYMBOLIC, ST_ASM_NUMERIC, ST_PCG, ST_LINE_SEARCH_MISC, ST_CG_ITERATIONS /* calls = iterations, ms = in-kernel time of the iteration loop */, ST_PCG_SETUP /* in-kernel: slice load + preconditioner */, ST_PROJ_SELECTED /* calls = elements selected */, ST_PROJ_CHANGED /* calls = elements modified */,
             ST_PCG_C_SPMV, ST_PCG_C_BAR, ST_PCG_C_RED, ST_PCG_C_VEC /* calls = SM cycles of CTA 0 in the PCG loop */,
             ST_TILE_PAIRS_PT, ST_TILE_PAIRS_EE, ST_TILE_PAIRS_ET, ST_CAND_PT, ST_CAND_EE, ST_CAND_ET /* calls = totals over the detections */, ST_PROJ_SWEEPS /* calls = Jacobi sweeps */, ST_PCG_C_WIN /* cycles: TMA window load */, ST_COUNT };
struct StageTimer {
    sb_context* ctx;
    int stage;
    double t0;
    StageTimer(sb_context* c, int s);
    ~StageTimer();
};

bool timeline_enabled();
void timeline_mark(sb_context* ctx, int stage);
void timeline_point(cudaStream_t st, const char* label);
void timeline_dump(sb_context* ctx, const char* const* names);
const char* const* stage_names();

// A helper host thread per context: issues an independent chain of launches (the symbolic phase of the assembly: ~20 small
// kernels on its own stream) while the calling thread keeps issuing the evaluation.  A Newton iteration at the 200k-tet scene
// is bound by the HOST's launch rate, not by the kernels; two issuing threads halve that.  One job at a time; wait() is the
// host-side join (the job's launches are then all enqueued; device-side ordering is by events as usual).
struct Issuer {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    std::atomic<int> state{0};   // 0 idle, 1 job posted / running
    bool quit = false;
    int device = 0;
    void start(int dev);
    void post(std::function<void()> f);
    void wait() { while (state.load(std::memory_order_acquire) != 0) { /* the job is a few tens of microseconds of launches */ } }
    void stop();
};

struct Assembly;   // assembly.cu
struct Pcg;        // pcg.cu
struct Dist;       // pcg.cu: peer-memory state of a distributed solve
struct Contact;    // contact.cu
struct Projector;  // project.cu
struct Direct;     // direct.cu

}  // namespace sb

struct sb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // side streams: the many small potentials (joints, contact and friction tables) of one evaluation run concurrently
    static constexpr int N_SIDE = 4;
    cudaStream_t side[N_SIDE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[N_SIDE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // sb_newton_solve's timing events
    cudaEvent_t ev_sync = nullptr;                  // hot_sync
    bool timer_started = false;                     // sb_newton_timer_begin has recorded ev_t0 for the coming solve
    cudaStream_t sym_stream = nullptr;              // the helper thread's stream (symbolic phase of the assembly)
    cudaEvent_t ev_dyn[N_SIDE] = {nullptr, nullptr, nullptr, nullptr};   // "the dynamic potentials' kernels on this side stream are done"
    sb::Issuer* issuer = nullptr;
    cudaStream_t bulk_stream = nullptr;             // pre-launched volume kernels (low priority: they run beside the collision detection)
    cudaEvent_t ev_bulk = nullptr;
    bool bulk_pending = false;                      // ev_bulk has been recorded and not yet joined into the context stream
    // sb_array_download_async
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy_src = nullptr, ev_copy_done = nullptr;
    bool copy_dev_pending = false, copy_host_pending = false;
    std::string error;
    std::atomic<int64_t> launches{0};   // (kernels are also issued by the helper thread)

    std::vector<std::pair<void*, size_t>> host_regions;   // sb_host_register: pinned mirrors
    std::vector<sb::Array> arrays;
    std::vector<sb::DofSet> dof_sets;
    int ndofs = 0;
    std::vector<sb::Potential> potentials;

    // element outputs of the last PGH evaluation
    sb::DevBuf<double> H;        // all element Hessians
    sb::DevBuf<int32_t> rows;    // all element block rows
    sb::DevBuf<double> E_elem;   // all element energies
    sb::DevBuf<double> grad;     // flat gradient
    sb::DevBuf<double> du;       // Newton direction
    sb::DevBuf<double> dofs_saved;
    sb::DevBuf<double> scratch;  // reductions
    sb::DevBuf<uint8_t> projected;  // per element Hessian: already projected
    size_t n_hessians = 0, n_blocks_total = 0, n_rows_total = 0, H_total = 0;
    int64_t n_projected = 0;
    // Connectivity versions.  STATIC potentials (meshes, joints) change rarely; DYNAMIC ones (contact / friction tables)
    // change at every detection.  The element-output buffers are laid out static-first so that static offsets never move.
    uint64_t static_version = 1, dynamic_version = 1;
    size_t n_static_blocks = 0;     // element blocks of the static potentials (they are numbered first)
    uint64_t state_version = 1;     // bumped whenever an array / the DoFs / the contact set-up change (caches keyed on the state)
    uint64_t eval_id = 0;           // bumped by every PGH evaluation (the element Hessians are rewritten)
    bool have_pgh = false;
    // the last P+G+H evaluation, reusable while nothing it depends on has changed (see eval_internal)
    bool pgh_cache_ok = false;
    uint64_t pgh_state = 0, pgh_dynamic = 0, pgh_static = 0;
    double pgh_E = 0.0, pgh_residual = 0.0;
    // first half of a P+G+H evaluation (the static potentials) launched ahead of the collision detection (eval_prelaunch_static)
    unsigned char* d_mailbox = nullptr;      // everything the host reads after a fused detection + evaluation, packed for ONE copy
    unsigned char* h_mailbox = nullptr;
    sb::DynLayout* d_dyn_layout = nullptr;   // fused detection + evaluation (core.cu: eval_fused)
    sb::DynTotals* d_dyn_totals = nullptr;
    sb::DynTotals* h_dyn_totals = nullptr;
    sb::MultiGArgs multi_g[sb::MULTI_G_FAMILIES];        // staging of the multi-potential P+G+H launch (6 KB: not on the stack of every evaluation)
    bool locate_pending = false;   // a scatter-mode pattern lookup rides with this evaluation's scalars
    bool pre_valid = false;
    uint64_t pre_state = 0, pre_static = 0;
    size_t st_H = 0, st_rows = 0, st_E = 0, st_blocks = 0;   // totals of the static potentials in the element-output buffers
    unsigned st_side_mask = 0;
    int st_next_side = 0;

    bool profile = false;           // stage profiling on (adds a stream synchronisation at every stage boundary)
    double stage_ms[32] = {0};
    int64_t stage_calls[32] = {0};
    std::string profile_report;

    double* h_scalars = nullptr;    // pinned host scratch (16 doubles)
    double* d_scalars = nullptr;    // device scratch (16 doubles)

    sb::Assembly* assembly = nullptr;
    sb::Pcg* pcg = nullptr;
    sb::Contact* contact = nullptr;
    sb::Projector* projector = nullptr;
    sb::Direct* direct = nullptr;
    sb::Dist* dist = nullptr;
    std::vector<void*> user_kernels;   // user.cu: generated kernels owned by this context
    bool assemble_own_rows = false;    // set by the Newton driver around its assemblies: a shared PCG solve follows, other ranks' rows are not needed
};

namespace sb {
int fail(sb_context* ctx, int code, const std::string& msg);
void order_after_async_downloads(sb_context* ctx);
cudaError_t hot_sync(sb_context* ctx);
int check_cuda(sb_context* ctx, cudaError_t e, const char* what);
#define SB_CUDA(ctx, call) do { int _r = sb::check_cuda((ctx), (call), #call); if (_r) return _r; } while (0)
int recompute_dof_offsets(sb_context* ctx);
// potential indices in buffer-layout order: static potentials first, then dynamic ones
std::vector<int> layout_order(const sb_context* ctx);
int refresh_slots(sb_context* ctx, Potential& p);
// reductions (core.cu): deterministic sum / inf-norm into d_out[0]
void reduce_sum(sb_context* ctx, const double* d_in, size_t n, double* d_out);
void reduce_absmax(sb_context* ctx, const double* d_in, size_t n, double* d_out);
void assembly_destroy(sb_context* ctx);
void pcg_destroy(sb_context* ctx);
void contact_destroy(sb_context* ctx);
void projector_destroy(sb_context* ctx);
void direct_destroy(sb_context* ctx);
void dist_destroy(sb_context* ctx);
void user_kernels_destroy(sb_context* ctx);
bool dist_aborted(sb_context* ctx);
bool dist_own_rows_only(sb_context* ctx, const unsigned long long* rows, int nbr, size_t nnzb, unsigned long long* d_range2);   // pcg.cu
int dist_bcast_from_root(sb_context* ctx, double* vec, int n, double* scal, int n_scal);   // pcg.cu: rank 0's values replace every rank's (no-op without peers)
int potential_create_with_kernel(sb_context* ctx, const KernelInfo* k, const char* kernel_name, int conn_stride, const sb_fetch* fetch, int n_fetch, int* out_potential);
int solve_llt_internal(sb_context* ctx, int* out_ok, double* out_du_dot_grad, double* out_du_inf);
int assemble_internal(sb_context* ctx);
// where the blocks of an element land in the assembled matrix: sources are numbered static-first; a static source maps
// through its static block, a dynamic one through its dynamic block
struct DirtyView {
    unsigned long long n_static;
    const uint32_t* s_blk_of_src; const uint32_t* s_final;
    const uint32_t* d_final_of_src;   // per dynamic source: its BCSR block
    uint8_t* dirty;
};
bool assembly_dirty_view(sb_context* ctx, DirtyView* v);
int project_internal(sb_context* ctx, double grad_threshold, double eps, int mirror, int64_t* out_n_projected, int64_t* out_n_hessians, int* out_all_projected);
int project_selection_bounds(sb_context* ctx, double* out_m, double* out_gmin);   // project.cu: when does a falling PPN threshold select something new?
int solve_pcg_internal(sb_context* ctx, double abs_tol, double rel_tol, int max_iter, int stop_on_indef, int* out_iterations, int* out_ok, double* out_du_dot_grad, double* out_du_inf);
// contact hooks used by the Newton driver (contact.cu)
bool contact_fusable(sb_context* ctx);
int contact_fused_issue(sb_context* ctx);
int contact_issue_table(sb_context* ctx, int pot, const int32_t** conn, const int** count_dev, int* cap);
int contact_fused_finish(sb_context* ctx, int* out_intersections, bool* retry);
void contact_readback_sources(sb_context* ctx, const int** counters, const unsigned long long** hash);
void contact_readback_deliver(sb_context* ctx, const int* counters, const unsigned long long* hash);
int contact_n_digest_words();
int eval_fused(sb_context* ctx, int* out_intersections, double* out_E, double* out_res, bool* out_done);   // core.cu
void eval_discard(sb_context* ctx);
// assembly.cu, fused path: lookup + speculative scatter with the dynamic sources' count / descriptors on the device
bool assembly_locate_ready(sb_context* ctx);
PotDesc* assembly_descs_dev(sb_context* ctx, int n);
size_t assembly_scatter_capacity(sb_context* ctx, size_t n_src_estimate);
int assembly_locate_dynamic_dev(sb_context* ctx, const DynTotals* d_tot, size_t n_src_estimate);
void assembly_locate_result_dev(sb_context* ctx, bool miss, size_t n_src, bool scatter_valid);
int contact_update_internal(sb_context* ctx);
int contact_intersections_internal(sb_context* ctx, int* out_count);
bool contact_active(sb_context* ctx);
int eval_internal(sb_context* ctx, int mode, double* out_E, double* out_grad_inf, bool sync_scalars);
int eval_prelaunch_static(sb_context* ctx);
extern int g_tet_grid_cap;   // persistent grid of the volume kernel (tet_analytic.cuh): 3 CTAs per SM alone, 2 when it shares the SMs with the detection   // static potentials of the coming P+G+H evaluation, before the collision detection
}  // namespace sb
