/*
 * stark_b200 -- C-ABI of the B200-native Newton hot path of STARK.
 *
 * Plain pointers and sizes only; every function returns 0 on success or a negative sb_status and never throws.
 * `sb_last_error()` gives the message of the last failure.  One context per Simulation, one host thread per
 * context, calls are synchronous from the caller's view (SURVEY.md section 8(b) "Threading").
 *
 * Each group of entry points names the reference interface it replaces (paths relative to /root/reference;
 * symx/ = stark/extern/symx/src/, tmcd/ = stark/extern/TriangleMeshCollisionDetection/src/, S/ = stark/src/).
 */
#ifndef STARK_B200_H
#define STARK_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SB_API __attribute__((visibility("default")))
#else
#define SB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_context sb_context;

enum sb_status {
    SB_OK = 0,
    SB_ERR_CUDA = -1,       /* a CUDA runtime call failed */
    SB_ERR_ARG = -2,        /* invalid argument / unknown handle */
    SB_ERR_LAYOUT = -3,     /* fetch table does not match the kernel's in[] layout */
    SB_ERR_NO_KERNEL = -4,  /* no hand-written kernel for this potential name */
    SB_ERR_STATE = -5       /* call made in the wrong order */
};

/* ---- context -------------------------------------------------------------------------------------------------
 * replaces: symx::Context (symx/solver/Context.h:11-22) as far as execution resources go.
 * `stream` is a cudaStream_t (0 = the library creates its own non-blocking stream). */
SB_API int sb_create(sb_context** out, int device, void* stream);
SB_API void sb_destroy(sb_context* ctx);
SB_API const char* sb_last_error(const sb_context* ctx);
SB_API void* sb_get_stream(sb_context* ctx);
SB_API int sb_synchronize(sb_context* ctx);
/* number of kernels launched by this context since creation (bench.py `gpu_launches`) */
SB_API int64_t sb_launch_count(const sb_context* ctx);

/* ---- arrays ----------------------------------------------------------------------------------------------------
 * replaces: symx::DataMap<const double> (symx/compile/data_maps.h:86-105): an array of `n_rows` x `stride` doubles,
 * row-major, identified by a small integer handle.  Host buffers are borrowed for the duration of the call only. */
SB_API int sb_array_create(sb_context* ctx, const char* label, int stride, int* out_array);
SB_API int sb_array_upload(sb_context* ctx, int array, const double* host, int n_rows);
SB_API int sb_array_download(sb_context* ctx, int array, double* host, int n_rows);
SB_API int sb_array_rows(sb_context* ctx, int array, int* out_rows);
/* every entry of the first n_rows rows = value (the reference zeroes soft.v1 / rigid.v1 / rigid.w1 before every time step,
 * S/models/deformables/PointDynamics.cpp:58-62, S/models/rigidbodies/RigidBodyDynamics.cpp:136-147) */
SB_API int sb_array_fill(sb_context* ctx, int array, int n_rows, double value);
/* State roll of an accepted time step on the device: y += alpha * x (x0 += dt v1) and dst = src (v0 = v1) over the first n_rows
 * rows; replaces the host loops of PointDynamics::_on_time_step_accepted (S/models/deformables/PointDynamics.cpp:64-78). */
SB_API int sb_array_axpy(sb_context* ctx, int y, int x, double alpha, int n_rows);
SB_API int sb_array_copy(sb_context* ctx, int dst, int src, int n_rows);
/* Asynchronous read-back into a buffer registered with sb_host_register: ordered behind everything submitted so far, runs beside
 * whatever is submitted next, complete after sb_download_wait.  (The host mirrors of x0 / v0 are refreshed this way while the
 * next time step is already being solved.) */
SB_API int sb_array_download_async(sb_context* ctx, int array, double* host, int n_rows);
SB_API int sb_download_wait(sb_context* ctx);
/* Pinned mirrors (SURVEY.md 8(b) "Data ownership"): a host buffer registered here is page-locked in place, and uploads
 * FROM it are asynchronous -- the buffer must stay untouched until the next call that synchronises (any download, sb_eval,
 * sb_newton_solve, sb_synchronize).  Unregistered (pageable) buffers keep the borrow-for-the-call contract. */
SB_API int sb_host_register(sb_context* ctx, void* host, uint64_t bytes);
SB_API int sb_host_unregister(sb_context* ctx, void* host);

/* ---- degrees of freedom ---------------------------------------------------------------------------------------
 * replaces: GlobalPotential::add_dof / get_dofs / set_dofs / get_dofs_offsets (symx/solver/GlobalPotential.cpp:86-144).
 * DoF sets are concatenated in registration order (soft.v1, rigid.v1, rigid.w1 in STARK). */
SB_API int sb_dof_add(sb_context* ctx, int array, int* out_set);
SB_API int sb_dof_total(sb_context* ctx, int* out_ndofs);
SB_API int sb_dofs_get(sb_context* ctx, double* host_u);       /* flat, length ndofs */
SB_API int sb_dofs_set(sb_context* ctx, const double* host_u);

/* ---- potentials --------------------------------------------------------------------------------------------------
 * replaces: GlobalPotential::add_potential (symx/solver/GlobalPotential.cpp:22-84) + SecondOrderCompiledPotential
 * (symx/solver/second_order/SecondOrderCompiledPotential.cpp:6-87) + the BlockFetch table of CompiledInLoop::run
 * (symx/compile/CompiledInLoop_run.h:170-190).
 * `kernel_name` is the reference's potential name ("EnergyTetStrain", "contact_rb_d_pt_tp_cubic", ...).
 * `fetch[i]` binds in[first_slot .. first_slot+stride) to row conn[elem][conn_col] of `array` (conn_col = -1: row 0). */
typedef struct sb_fetch {
    int32_t array;
    int32_t conn_col;
    int32_t first_slot;
    int32_t stride;
} sb_fetch;
SB_API int sb_potential_create(sb_context* ctx, const char* kernel_name, int conn_stride, const sb_fetch* fetch, int n_fetch, int* out_potential);
SB_API int sb_potential_set_connectivity(sb_context* ctx, int potential, const int32_t* conn, int n_elements);
SB_API int sb_potential_info(sb_context* ctx, int potential, int* n_in, int* n_dofs, int* n_elements);
/* User potentials (GlobalPotential::add_potential with an energy no built-in kernel covers, e.g. examples/main.cpp:666-690):
 * replaces the reference's code generator + host-compiler JIT (symx/compile/Compilation.cpp:381-469 `_add_instructions_scalar`,
 * one C statement per symx::core::Op of a Sequence -- compile/FixedBranchSequence.h:44-83 -- then g++ and dlopen) by CUDA source
 * + NVRTC + a cubin cache ($SB_CACHE_DIR, default ~/.cache/stark_b200, keyed by a hash of the generated source).
 * The caller differentiates the energy with symx itself (SecondOrderCompiledPotential.cpp:62-80: gradient, symmetric Hessian)
 * and passes the operation sequences of [E] and of [E | grad(n) | hess(n x n, row-major)] with n = 3 n_blocks:
 *   sb_op mirrors symx::core::Op: type = symx::ExprType value (symbol/Expr.h:12-44; Symbol = 5 means out[dst] = value a),
 *   dst / a / b / cond = indices into [in[0 .. n_in) | temporaries], constant = value of a ConstantFloat; Branch ops use the
 *   reference's encoding (cond == -2: end-if; a == 0: `if (value[cond] > 0) {`; a == 1: `} else {`).
 * dof_block_slots[b] = in[] slot of the first of the three DoF symbols of block b (DoF-set order, then slot order, as
 * SecondOrderCompiledPotential builds `dofs`).  Fails with SB_ERR_NO_KERNEL when NVRTC cannot be loaded. */
typedef struct sb_op {
    int32_t type, dst, a, b, cond, pad;
    double constant;
} sb_op;
SB_API int sb_potential_create_user(sb_context* ctx, const char* name, int conn_stride, const sb_fetch* fetch, int n_fetch, int n_in, int n_blocks,
                                    const int32_t* dof_block_slots, const sb_op* ops_p, int n_ops_p, const sb_op* ops_pgh, int n_ops_pgh, int* out_potential);
/* the two halves of the above without a GPU: the generated CUDA source (out_length = its size; out_source may be NULL), and
 * its compilation for sm_100a (NVRTC cross-compiles; out_log receives the compiler's messages on failure) */
SB_API int sb_user_codegen(const char* name, int n_in, int n_blocks, const sb_op* ops_p, int n_ops_p, const sb_op* ops_pgh, int n_ops_pgh,
                           char* out_source, long long capacity, long long* out_length);
SB_API int sb_user_compile(const char* source, long long* out_cubin_bytes, int* out_was_cached, char* out_log, int log_capacity);
/* names of all built-in kernels, '\n' separated */
SB_API const char* sb_kernel_names(void);

/* ---- evaluation ---------------------------------------------------------------------------------------------------
 * replaces: SecondOrderCompiledGlobal::{evaluate_P, evaluate_P__dP_du__local_d2P_du2}
 * (symx/solver/second_order/SecondOrderCompiledGlobal.cpp:72-93, 119-142). */
enum sb_eval_mode { SB_EVAL_P = 0, SB_EVAL_PGH = 2 };
/* A SB_EVAL_PGH call at a state that has just been evaluated with SB_EVAL_PGH -- no upload, DoF change, connectivity or
 * contact-table change and no PD projection in between -- returns the stored energy and gradient norm without launching
 * anything: gradient, element Hessians and block rows on the device are still that evaluation's.  (sb_newton_solve uses this:
 * the first line-search trial is evaluated with SB_EVAL_PGH and becomes the next iteration's evaluation when accepted.) */
SB_API int sb_eval(sb_context* ctx, int mode, double* out_E, double* out_grad_inf);
/* Optional hint: start the kernels of the next SB_EVAL_PGH evaluation that do not depend on the contact tables (volume / shell /
 * inertia / joint potentials) NOW, at the current state.  A collision detection issued next (sb_contact_begin_time_step,
 * sb_contact_update) then runs beside them; sb_eval / sb_newton_solve pick the result up if the state has not changed since. */
SB_API int sb_eval_prelaunch(sb_context* ctx);
SB_API int sb_grad_get(sb_context* ctx, double* host_grad);   /* flat, length ndofs */
/* per-element output of one potential as the reference lays it out: [E | grad(n) | hess(n*n) row-major] per element */
SB_API int sb_potential_get_element_output(sb_context* ctx, int potential, double* host_sol);
SB_API int sb_potential_get_block_rows(sb_context* ctx, int potential, int32_t* host_rows);
/* the stored element Hessians of one potential (n*n per element, row-major) as they are now, i.e. after any projection */
SB_API int sb_potential_get_hessians(sb_context* ctx, int potential, double* host_hessians);

/* ---- PD projection + assembly -----------------------------------------------------------------------------------
 * replaces: ElementHessians::{project_to_PD_*, assemble_global, update_global}
 * (symx/solver/second_order/ElementHessians.cpp:48-294), project_to_PD_inplace (project_to_PD.cpp:13-82) and
 * BlockedSparseMatrix insertion (bsm/BlockedSparseMatrix.h:332-593, 782-895).
 * project: grad_threshold < 0 projects nothing, 0 projects every element, > 0 projects elements touching a DoF block
 * with |g|_inf >= threshold (PPN selection, NewtonsMethod.cpp:316-327).  Returns counts for the `ph` statistic. */
SB_API int sb_project_to_pd(sb_context* ctx, double grad_threshold, double eps, int mirror, int64_t* out_n_projected, int64_t* out_n_hessians, int* out_all_projected);
SB_API int sb_assemble(sb_context* ctx);
SB_API int sb_bcsr_info(sb_context* ctx, int* n_block_rows, int64_t* nnzb);
/* reference layout: rows u64[nbr+1], cols = first scalar column of the block, vals float[9*nnzb] column-major per block */
SB_API int sb_bcsr_get(sb_context* ctx, int64_t* host_rows, int32_t* host_cols, float* host_vals);

/* ---- linear solve ---------------------------------------------------------------------------------------------------
 * replaces: NewtonsMethod::_solve_linear_system (symx/solver/NewtonsMethod.cpp:388-457), bsm::solve_pcg
 * (bsm/solve_pcg.h:83-232) with the block-Jacobi preconditioner (bsm/BlockedSparseMatrix.h:1147-1361).
 * Solves H du = -grad.  out_ok = 0 when p^T A p <= 0 was met (indefiniteness) or max_iterations was hit. */
SB_API int sb_solve_pcg(sb_context* ctx, double abs_tol, double rel_tol, int max_iterations, int stop_on_indefiniteness,
                 int* out_iterations, int* out_ok, double* out_du_dot_grad, double* out_du_inf);
/* DirectLLT branch of the same function (NewtonsMethod.cpp:395-418: to_triplets -> Eigen::SimplicialLLT): sparse FP64
 * Cholesky of the assembled matrix -- reverse Cuthill-McKee ordering of the 3x3-block graph (dense rows last), factor stored
 * as a row envelope of 64x64 tiles, tile POTRF / TRSM / SYRK on the device; ordering and envelope are cached while the
 * pattern stays inside them.  The factor must fit SB_LLT_MAX_BYTES (environment, default 96 GiB) or the call fails with
 * SB_ERR_STATE.  out_ok = 0 when the matrix is not positive definite (Eigen's info() != Success). */
SB_API int sb_solve_llt(sb_context* ctx, int* out_ok, double* out_du_dot_grad, double* out_du_inf);
/* diagnostics of the last sb_solve_llt: out5 = { tiles in the envelope, tile rows, factor bytes, orderings made so far,
 * analyses made so far } */
SB_API int sb_llt_stats(sb_context* ctx, double* out5);
/* the fill-reducing ordering sb_solve_llt uses, on HOST arrays in sb_bcsr_get's layout (rows: nbr + 1 block offsets, cols:
 * first scalar column of every block); out_perm[block row] = position in the elimination order.  Needs no GPU. */
SB_API int sb_llt_order(int nbr, const unsigned long long* rows, const int32_t* cols, int32_t* out_perm);
SB_API int sb_du_get(sb_context* ctx, double* host_du);

/* ---- line search support ----------------------------------------------------------------------------------------------
 * replaces: the apply_scaled_du lambda of NewtonsMethod::_line_search_inplace (symx/solver/NewtonsMethod.cpp:469-476).
 * save: remember the current DoFs; apply: dofs = saved + step * scale * du. */
SB_API int sb_dofs_save(sb_context* ctx);
SB_API int sb_dofs_apply_step(sb_context* ctx, double step);
SB_API int sb_du_scale(sb_context* ctx, double factor);

/* ---- collision detection + contact lists ---------------------------------------------------------------------------------
 * replaces: tmcd::ProximityDetection / IntersectionDetection (tmcd/ProximityDetection.cpp:74-220,
 * tmcd/IntersectionDetection.cpp:54-108, tmcd/BroadPhasePTEEBase.cpp, tmcd/AABBs.cpp, tmcd/Octree.cpp) and the vertex
 * update + list building of EnergyFrictionalContact (S/models/interactions/EnergyFrictionalContact.cpp:219-262, 368-799).
 * Meshes are registered once; per-mesh vertex positions are produced on the device from the DoF arrays. */
typedef struct sb_contact_mesh {
    int32_t physical_system;   /* 0 = deformable, 1 = rigid body */
    int32_t rigid_body;        /* body index when rigid */
    int32_t n_vertices;
    int32_t n_triangles;
    int32_t n_edges;
    const int32_t* vertex_global;   /* deformable: global node index of every collision vertex; rigid: NULL */
    const double* vertices_local;   /* rigid: local coordinates (3 per vertex); deformable: NULL */
    const int32_t* triangles;       /* 3 per triangle, local vertex ids */
    const int32_t* edges;           /* 2 per edge, local vertex ids */
    double contact_thickness;
} sb_contact_mesh;
typedef struct sb_contact_bindings {
    /* arrays the contact module reads / binds into the contact + friction potentials */
    int32_t soft_v1, soft_x0, soft_X;          /* deformable DoF, x0, rest positions */
    int32_t rb_v1, rb_w1, rb_t0, rb_q0;        /* rigid DoFs and state (q0 as w,x,y,z) */
    int32_t dt;                                /* 1x1 array holding the time step */
} sb_contact_bindings;
SB_API int sb_contact_init(sb_context* ctx, const sb_contact_bindings* bindings);
SB_API int sb_contact_add_mesh(sb_context* ctx, const sb_contact_mesh* mesh, int* out_group);
SB_API int sb_contact_blacklist(sb_context* ctx, int group_a, int group_b);
SB_API int sb_contact_set_friction(sb_context* ctx, int group_a, int group_b, double mu);
SB_API int sb_contact_set_params(sb_context* ctx, double contact_stiffness, double friction_stick_slide_threshold,
                          int enable_point_triangle, int enable_edge_edge, int enable_friction);
/* before_energy_evaluation: vertex update + proximity + the 21 contact tables (dt < 0: use the bound dt array) */
SB_API int sb_contact_update(sb_context* ctx);
/* before_time_step: proximity at dt = 0 + friction tables with T / bary / fn / mu */
SB_API int sb_contact_update_friction(sb_context* ctx);
/* before_time_step in one detection: friction tables (dt = 0) and -- when every DoF array is zero, as in the reference's callback
 * order (PointDynamics.cpp:58-62 and RigidBodyDynamics.cpp:136-147 zero v1 / w1 before EnergyFrictionalContact.cpp:531 runs) --
 * also the contact tables and the intersection count of the initial Newton state, which sb_newton_solve then finds cached.
 * dofs_are_zero = 0: same as sb_contact_update_friction. */
SB_API int sb_contact_begin_time_step(sb_context* ctx, int dofs_are_zero);
/* is_intermediate_state_valid: number of edge-triangle intersections at the current DoFs.  When the contact tables are not
 * current at these DoFs, the same detection also rebuilds them (what sb_contact_update would do next at this state: the
 * evaluation that follows a valid state starts with it), so the two share one vertex update and one synchronisation. */
SB_API int sb_contact_count_intersections(sb_context* ctx, int* out_count);
/* raw results for parity tests: kind 0..5 = pt_pp, pt_pe, pt_pt, ee_pp, ee_pe, ee_ee; 6 = intersections.
 * ids rows as in tests/golden (see oracle/ref_driver.cpp); call with NULL buffers to query the count. */
SB_API int sb_contact_get_proximity(sb_context* ctx, int kind, int32_t* host_ids, double* host_dist, int capacity, int* out_count, int* out_width);
SB_API int sb_contact_get_vertices(sb_context* ctx, int group, double* host_xyz);
/* feed externally computed vertex positions (tests of the detection stage alone) */
SB_API int sb_contact_set_vertices(sb_context* ctx, int group, const double* host_xyz);
SB_API int sb_contact_detect(sb_context* ctx, double enlargement, int with_intersections);
/* contact / friction potential handle by reference name (e.g. "contact_rb_d_pt_tp_cubic"), -1 if absent */
SB_API int sb_contact_potential(sb_context* ctx, const char* name, int* out_potential);

/* ---- Newton driver -----------------------------------------------------------------------------------------------------------
 * replaces: symx::NewtonsMethod::solve (symx/solver/NewtonsMethod.cpp:28-252) with NewtonSettings
 * (symx/solver/solver_utils.h:170-258) and SolveStats (symx/solver/NewtonsMethod.h:32-43). */
typedef struct sb_newton_settings {
    int32_t max_iterations, min_iterations;
    double residual_tolerance_abs, residual_tolerance_rel, step_tolerance;
    int32_t max_iterations_as_success;
    double step_cap;
    int32_t enable_armijo_backtracking;
    double line_search_armijo_beta;
    int32_t max_backtracking_armijo_iterations, max_backtracking_invalid_state_iterations;
    int32_t projection_mode;   /* 0 Newton, 1 ProjectedNewton, 2 ProjectOnDemand, 3 Progressive */
    double projection_eps;
    int32_t project_to_pd_use_mirroring, project_on_demand_countdown;
    double ppn_tightening_factor, ppn_release_factor;
    int32_t linear_solver;     /* 0 DirectLLT (sparse tile-envelope Cholesky), 1 BDPCG */
    int32_t cg_max_iterations;
    double cg_abs_tolerance, cg_rel_tolerance;
    int32_t cg_stop_on_indefiniteness;
    double bailout_residual;
    int32_t contact_enabled;   /* run the built-in contact callbacks (update / intersection test) */
    int32_t skip_converged_state_check;   /* the caller runs the is_converged_state_valid callbacks itself */
    /* EnergyFrictionalContact::GlobalParams::intersection_test_enabled (S/models/interactions/EnergyFrictionalContact.cpp:774-799):
     * 0 = is_initial / is_intermediate / is_converged_state_valid always answer "valid" (no edge-triangle test, no step halving) */
    int32_t intersection_test_enabled;
} sb_newton_settings;
typedef struct sb_newton_stats {
    int32_t result;            /* symx::SolverReturn value (solver_utils.h:15-26) */
    int32_t newton_iterations, cg_iterations;
    int32_t ls_cap_iterations, ls_max_iterations, ls_inv_iterations, ls_bt_iterations;
    int64_t n_hessians, n_projected_hessians;
    int32_t n_evaluations;
    double last_residual, last_energy;
    double residuals[64];      /* residual of every evaluation (first 64) */
    double gpu_ms;             /* device time of the solve: CUDA events on the context stream around the whole call */
} sb_newton_stats;
/* Optional: start the device clock of the next sb_newton_solve NOW (stats.gpu_ms then also covers what is submitted between this
 * call and the solve: the start-of-step collision detection, a pre-launched evaluation).  Without it the clock starts at the
 * solve's entry. */
SB_API int sb_newton_timer_begin(sb_context* ctx);
SB_API void sb_newton_default_settings(sb_newton_settings* s);   /* STARK's defaults (S/core/Settings.cpp:40-54) */
SB_API int sb_newton_solve(sb_context* ctx, const sb_newton_settings* settings, sb_newton_stats* stats);

/* ---- multi-GPU: distributed linear solve over peer memory ---------------------------------------------------------------------
 * replaces: nothing in the reference (it is single-process OpenMP); shards what SURVEY.md 8(e) names -- the SpMV
 * (bsm/BlockedSparseMatrix.h:988-1138) and the PCG body (bsm/solve_pcg.h:174-222) -- over the GPUs of one NVSwitch domain.
 * One process (or thread) per GPU, every rank drives the SAME scene (replicated state, evaluation and assembly); after
 * sb_dist_init + sb_dist_connect every sb_solve_pcg / BDPCG solve of sb_newton_solve is ONE solve shared by all ranks: the
 * persistent kernels of the ranks form one virtual grid, each keeps 1 / world of the matrix in shared memory, and the halo of
 * u, the dot-product partials, the barrier and du cross GPUs through NVLink loads / stores issued inside the kernel on
 * buffers mapped with CUDA IPC.  All ranks must issue the same sequence of solves: to that end every sb_eval of a connected
 * context ends with rank 0's gradient, energy and residual replacing every rank's (FP64 atomics make their last bits differ
 * between replicas, and replicas that drift apart end up taking different decisions), so connect before the first evaluation.
 * A rank that waits longer than SB_DIST_TIMEOUT_S (environment, default 30 s) for its peers fails with SB_ERR_CUDA instead of
 * hanging.
 * sb_dist_init allocates the rank's peer buffer (sized for max_dofs) and returns its cudaIpcMemHandle_t (64 bytes);
 * sb_dist_connect takes the world x 64 bytes of all ranks' handles (gathered by the caller, e.g. torch.distributed). */
SB_API int sb_dist_init(sb_context* ctx, int rank, int world, long long max_dofs, unsigned char* out_handle64);
SB_API int sb_dist_connect(sb_context* ctx, const unsigned char* handles);
/* ranks inside one process connect by pointer instead (bases[q] = sb_dist_local_base of rank q's context) */
SB_API int sb_dist_connect_ptrs(sb_context* ctx, void* const* bases);
SB_API void* sb_dist_local_base(sb_context* ctx);
/* sharing off / on again (every rank between the same two solves): off = each rank solves its own replica locally */
SB_API int sb_dist_set_enabled(sb_context* ctx, int enabled);
/* out4 = { barriers completed, distributed solves, bytes of the peer buffer, solves the policy kept on the own GPU }.
 * Policy (environment SB_DIST_POLICY = auto | always, default auto): a system whose matrix is resident in ONE GPU's shared memory
 * is solved locally by every rank (its iteration is bound by grid-wide synchronisation, which costs more across NVLink than on
 * chip); systems that do not fit are shared by all ranks.  All ranks decide alike (same pattern sizes). */
SB_API int sb_dist_stats(sb_context* ctx, int* out_rank, int* out_world, double* out4);
/* The row partition and the halo send lists of a distributed solve, on HOST arrays in sb_bcsr_get's layout (no GPU needed):
 * out_bounds[world + 1] = first block row of every rank; out_needmask[nbr]: for the block rows `rank` owns, bit q is set when
 * rank q's rows reference that column of u (the owner pushes the row to q every iteration).  grid = CTAs per rank. */
SB_API int sb_dist_plan(int nbr, const unsigned long long* rows, const int32_t* cols, int world, int grid, int rank,
                        int32_t* out_bounds, unsigned char* out_needmask);

/* ---- measurement ----------------------------------------------------------------------------------------------------------
 * Launches the element kernel of ONE potential `reps` times on the current state (mode as sb_eval) and returns the average
 * launch duration measured with CUDA events on the context stream (bench.py roofline). */
SB_API int sb_profile_potential(sb_context* ctx, int potential, int mode, int reps, double* out_avg_ms);
/* Stage profiling of the Newton path (the analogue of the reference's symx::Logger scoped timers, symx/solver/Logger.h):
 * when enabled every stage is bracketed by stream synchronisations and its host wall time accumulated.  The report is
 * one line per stage: "<name> <total ms> <calls>".  Enabling resets the totals.  Off by default (it serialises the path). */
SB_API int sb_profile_stages(sb_context* ctx, int enable);
SB_API const char* sb_profile_report(sb_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* STARK_B200_H */
