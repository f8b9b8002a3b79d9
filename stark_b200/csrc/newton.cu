// Newton driver on device-resident state.
//
// Replaces symx::NewtonsMethod::solve and its helpers (symx/solver/NewtonsMethod.cpp:28-641): same control flow, same
// projection policies (Newton / ProjectedNewton / ProjectOnDemand / Progressive), same Eisenstat-Walker forcing
// tolerance for the PCG, same four-stage line search (cap / max / invalid-state halving / Armijo).  The host only sees
// scalars (E, |g|_inf, du.g, |du|_inf, counts); DoFs, gradient, element Hessians, BCSR and contact tables stay on the GPU.
// The model callbacks the reference runs on the host are the built-in contact stages here:
//   before_energy_evaluation  -> contact_update_internal       (EnergyFrictionalContact.cpp:368-530)
//   is_*_state_valid          -> contact_intersections_internal (EnergyFrictionalContact.cpp:774-799)
#include "internal.h"
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <limits>

using namespace sb;

namespace {
// symx::SolverReturn (symx/solver/solver_utils.h:15-26)
enum Ret { Successful = 0, Running, InvalidInitialState, TooManyIterations, TooManyArmijoIterations, LinearSystemSolveFailure,
           TooManyInvalidIntermediateIterations, StepDoesNotDescend, InvalidConvergedState };
enum Proj { PNewton = 0, PProjectedNewton = 1, PProjectOnDemand = 2, PProgressive = 3 };
}

extern "C" void sb_newton_default_settings(sb_newton_settings* s)
{
    if (!s) return;
    // symx defaults (solver_utils.h:170-258) with STARK's overrides (S/core/Settings.cpp:45-49)
    s->max_iterations = std::numeric_limits<int>::max();
    s->min_iterations = 0;
    s->residual_tolerance_abs = 1e-6;
    s->residual_tolerance_rel = 0.0;
    s->step_tolerance = 1e-3;
    s->max_iterations_as_success = 0;
    s->step_cap = std::numeric_limits<double>::infinity();
    s->enable_armijo_backtracking = 1;
    s->line_search_armijo_beta = 1e-4;
    s->max_backtracking_armijo_iterations = 20;
    s->max_backtracking_invalid_state_iterations = 8;
    s->projection_mode = PProgressive;
    s->projection_eps = 1e-10;
    s->project_to_pd_use_mirroring = 0;
    s->project_on_demand_countdown = 4;
    s->ppn_tightening_factor = 0.5;
    s->ppn_release_factor = 2.0;
    s->linear_solver = 1;
    s->cg_max_iterations = 10000;
    s->cg_abs_tolerance = 1e-12;
    s->cg_rel_tolerance = 1e-4;
    s->cg_stop_on_indefiniteness = 1;
    s->bailout_residual = 1e-10;
    s->contact_enabled = 1;
    s->skip_converged_state_check = 0;
    s->intersection_test_enabled = 1;
}

extern "C" int sb_newton_timer_begin(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    if (!ctx->ev_t0) { cudaEventCreate(&ctx->ev_t0); cudaEventCreate(&ctx->ev_t1); }
    SB_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    ctx->timer_started = true;
    return SB_OK;
}

extern "C" int sb_newton_solve(sb_context* ctx, const sb_newton_settings* S, sb_newton_stats* stats)
{
    if (!ctx || !S || !stats) return fail(ctx, SB_ERR_ARG, "sb_newton_solve: bad argument");
    if (S->linear_solver != 0 && S->linear_solver != 1) return fail(ctx, SB_ERR_ARG, "sb_newton_solve: unknown linear solver");
    recompute_dof_offsets(ctx);
    const int ndofs = ctx->ndofs;
    if (ndofs <= 0) return fail(ctx, SB_ERR_STATE, "sb_newton_solve: no degrees of freedom");
    if (ndofs % 3 != 0) return fail(ctx, SB_ERR_STATE, "sb_newton_solve: ndofs must be divisible by 3");
    *stats = sb_newton_stats();
    const bool contact = S->contact_enabled && contact_active(ctx);
    int rc;
    auto record = [&](double r) { if (stats->n_evaluations < 64) stats->residuals[stats->n_evaluations] = r; stats->n_evaluations++; };
    auto state_valid = [&](bool& valid) -> int {
        valid = true;
        if (!contact || !S->intersection_test_enabled) return 0;
        int n = 0;
        int r = contact_intersections_internal(ctx, &n);
        if (r) return r;
        valid = (n == 0);
        return 0;
    };

    double E0 = 0.0, du_dot_grad = 0.0, res_0 = std::numeric_limits<double>::max();
    int result = Running;
    int pdn_countdown = 0;
    double ppn_threshold = -1.0;

    // (the two timing events live in the context: an error return below leaks nothing; `result` stays "Running" for a
    //  caller that ignores the return code)
    stats->result = Running;
    if (!ctx->ev_t0) { cudaEventCreate(&ctx->ev_t0); cudaEventCreate(&ctx->ev_t1); }
    cudaEvent_t ev0 = ctx->ev_t0, ev1 = ctx->ev_t1;
    if (!ctx->timer_started) cudaEventRecord(ev0, ctx->stream);   // (sb_newton_timer_begin started the clock earlier)
    ctx->timer_started = false;
    timeline_mark(ctx, -1);
    static const bool no_spec = std::getenv("SB_NO_SPECULATION") != nullptr;   // diagnostic hook
    // the static potentials of the first evaluation do not depend on the contact tables: their kernels are launched before the
    // collision detection of the initial state (eval_internal picks them up)
    if (!no_spec && (rc = eval_prelaunch_static(ctx))) return rc;
    bool valid = true;
    if ((rc = state_valid(valid))) return rc;
    if (!valid) result = InvalidInitialState;

    int it = -1;
    while (result == Running) {
        it++;
        if (it == S->max_iterations) {
            result = S->max_iterations_as_success ? Successful : TooManyIterations;
            break;
        }
        if (contact && (rc = contact_update_internal(ctx))) return rc;
        double residual = 0.0;
        if ((rc = eval_internal(ctx, SB_EVAL_PGH, &E0, &residual, true))) return rc;
        record(residual);
        stats->last_residual = residual;
        stats->last_energy = E0;
        if (it == 0) res_0 = residual;
        if (residual < S->bailout_residual) { result = Successful; break; }
        if (it >= S->min_iterations) {
            if (residual < S->residual_tolerance_abs) { result = Successful; break; }
            if (it > 0 && residual / res_0 < S->residual_tolerance_rel) { result = Successful; break; }
        }

        // ---- project + assemble + solve until the direction descends (NewtonsMethod.cpp:137-182)
        bool assembled = false;
        bool solved = false;
        double du_inf = 0.0;
        int64_t n_proj = 0, n_hess = 0;
        bool have_prev = false;      // a solve of this iteration's current matrix has already been made (and did not give a descent direction)
        int prev_ok = 0, prev_cg_it = 0;
        bool have_bounds = false;    // bound_m / bound_gmin describe the current projection flags (see project_selection_bounds)
        double bound_m = 0.0, bound_gmin = 0.0;
        while (!solved) {
            bool all_projected = false;
            const int64_t n_proj_before = n_proj;
            // _project_and_assemble (NewtonsMethod.cpp:254-352).  The reference assembles the unprojected matrix first and
            // then adds (projected - original) blocks; here the projected Hessians replace the originals in the element
            // store and one numeric assembly follows, so projection simply runs first.
            bool projected_now = false;
            int allp = 0;
            switch (S->projection_mode) {
            case PNewton: break;
            case PProjectedNewton:
                if ((rc = project_internal(ctx, 0.0, S->projection_eps, S->project_to_pd_use_mirroring, &n_proj, &n_hess, &allp))) return rc;
                all_projected = true; projected_now = true;
                break;
            case PProjectOnDemand:
                if (pdn_countdown > 0) {
                    if ((rc = project_internal(ctx, 0.0, S->projection_eps, S->project_to_pd_use_mirroring, &n_proj, &n_hess, &allp))) return rc;
                    all_projected = true; projected_now = true;
                }
                break;
            case PProgressive:
                if (ppn_threshold > 0.0) {
                    if (ppn_threshold < 1e-12) ppn_threshold = 0.0;
                    if (have_bounds && ppn_threshold > bound_m) {
                        // a threshold that is still above every unprojected element's largest block gradient selects nothing: the
                        // reference projects, assembles and solves again with an unchanged matrix; here the round is only counted
                        allp = (ppn_threshold <= bound_gmin) ? 1 : 0;
                    } else {
                        if ((rc = project_internal(ctx, ppn_threshold, S->projection_eps, S->project_to_pd_use_mirroring, &n_proj, &n_hess, &allp))) return rc;
                        if (n_proj != n_proj_before) have_bounds = false;
                        else if (!have_bounds && ppn_threshold > 0.0 && assembled) {
                            // nothing selected: find out how far the threshold has to fall (instead of one launch + sync per halving)
                            if ((rc = project_selection_bounds(ctx, &bound_m, &bound_gmin))) return rc;
                            have_bounds = true;
                        }
                    }
                    all_projected = (allp != 0); projected_now = true;
                }
                break;
            default: return fail(ctx, SB_ERR_ARG, "sb_newton_solve: unknown projection mode");
            }
            // A tightened threshold that selects no new element leaves the matrix as it is, and the (deterministic) solve would
            // fail exactly as it just did: the reference assembles and solves again; here the previous outcome is reused (its CG
            // iterations are counted again, so the statistics stay comparable) and the threshold is tightened further.
            const bool same_matrix = have_prev && assembled && projected_now && n_proj == n_proj_before;
            int cg_it = 0, ok = 0;
            if (same_matrix) {
                ok = prev_ok; cg_it = prev_cg_it;
            } else {
                if (!assembled || projected_now) {
                    ctx->assemble_own_rows = (S->linear_solver == 1);   // (several GPUs: a shared PCG solve only reads this rank's rows)
                    rc = assemble_internal(ctx);
                    ctx->assemble_own_rows = false;
                    if (rc) return rc;
                    assembled = true;
                }
                // _solve_linear_system (NewtonsMethod.cpp:420-451): forcing sequence
                const double forcing = std::min(1e-2, residual * std::min(0.5, std::sqrt(residual)));
                const double abs_tol = std::max(forcing, S->cg_abs_tolerance);
                if (S->linear_solver == 0) {
                    if ((rc = solve_llt_internal(ctx, &ok, &du_dot_grad, &du_inf))) return rc;
                } else if ((rc = solve_pcg_internal(ctx, abs_tol, S->cg_rel_tolerance, S->cg_max_iterations, S->cg_stop_on_indefiniteness, &cg_it, &ok, &du_dot_grad, &du_inf))) return rc;
            }
            have_prev = true; prev_ok = ok; prev_cg_it = cg_it;
            stats->cg_iterations += cg_it;
            const bool can_project_more = (S->projection_mode != PNewton) && !all_projected;
            if (!ok) {
                if (!can_project_more) { result = LinearSystemSolveFailure; break; }
            } else {
                if (du_dot_grad < 0.0) { solved = true; break; }
                if (!can_project_more) { result = StepDoesNotDescend; break; }
            }
            // _increase_projection (NewtonsMethod.cpp:354-371)
            if (S->projection_mode == PProjectOnDemand) pdn_countdown = S->project_on_demand_countdown;
            else if (S->projection_mode == PProgressive) {
                if (ppn_threshold < 0.0) ppn_threshold = residual;   // grad.cwiseAbs().maxCoeff()
                ppn_threshold *= S->ppn_tightening_factor;
            }
        }
        if (result != Running) break;
        // _decrease_projection
        if (S->projection_mode == PProjectOnDemand) pdn_countdown--;
        else if (S->projection_mode == PProgressive) ppn_threshold *= S->ppn_release_factor;
        stats->n_hessians += (int64_t)ctx->n_hessians;
        stats->n_projected_hessians += ctx->n_projected;

        if (it >= S->min_iterations && du_inf < S->step_tolerance) { result = Successful; break; }

        // ---- line search (NewtonsMethod.cpp:459-641)
        if ((rc = sb_dofs_save(ctx))) return rc;
        double retraction = 1.0;
        if (du_inf > S->step_cap) {
            retraction *= S->step_cap / du_inf;
            if ((rc = sb_du_scale(ctx, retraction))) return rc;
            du_inf = S->step_cap;
            stats->ls_cap_iterations++;
        }
        // [max] no max_allowed_step callback is registered by STARK
        double step = 1.0;
        if ((rc = sb_dofs_apply_step(ctx, step))) return rc;
        // The first Armijo trial is evaluated with gradient and Hessians (below); the volume elements of that evaluation do not
        // depend on the contact tables, so their kernels go out now, ahead of the trial state's collision detection.
        if (S->enable_armijo_backtracking && !no_spec && (rc = eval_prelaunch_static(ctx))) return rc;
        // First trial: collision detection (validity + contact tables) and the evaluation with gradient and Hessians are queued
        // together and synchronised once (eval_fused); anything it cannot handle goes the plain way below.
        bool have_E1 = false;
        double E1_fused = 0.0;
        int ls_inv = 0;
        for (; ls_inv < S->max_backtracking_invalid_state_iterations; ++ls_inv) {
            bool fused = false;
            if (ls_inv == 0 && contact && S->intersection_test_enabled && S->enable_armijo_backtracking && !no_spec) {
                int n_int = 0;
                double res_unused = 0.0;
                if ((rc = eval_fused(ctx, &n_int, &E1_fused, &res_unused, &fused))) return rc;
                if (fused) {
                    valid = (n_int == 0);
                    if (valid) have_E1 = true; else eval_discard(ctx);   // (an invalid state: the evaluation queued behind its detection is of no use)
                }
            }
            if (!fused && (rc = state_valid(valid))) return rc;
            if (valid) break;
            step *= 0.5;
            if ((rc = sb_dofs_apply_step(ctx, step))) return rc;
            stats->ls_inv_iterations++;
        }
        if (ls_inv == S->max_backtracking_invalid_state_iterations) { result = TooManyInvalidIntermediateIterations; break; }
        if (S->enable_armijo_backtracking) {
            const double expected = S->line_search_armijo_beta * du_dot_grad * retraction;
            double E_threshold = E0 + expected * step;
            double E1 = 0.0;
            int k = 0;
            for (; k < S->max_backtracking_armijo_iterations; ++k) {
                if (contact && (rc = contact_update_internal(ctx))) return rc;
                // The first trial is almost always accepted, and the next iteration then evaluates energy, gradient and
                // Hessians at this very state: evaluate them now (eval_internal hands the result out again) instead of the
                // energy alone.  Later trials (after a backtrack) are energy-only.
                if (k == 0 && have_E1) E1 = E1_fused;   // (evaluated together with the detection above; eval_internal hands it out again to the next iteration)
                else if (k == 0 && !no_spec) { double res_unused = 0.0; if ((rc = eval_internal(ctx, SB_EVAL_PGH, &E1, &res_unused, true))) return rc; }
                else if ((rc = eval_internal(ctx, SB_EVAL_P, &E1, nullptr, true))) return rc;
                if (E1 < E_threshold) break;
                step *= 0.5;
                if ((rc = sb_dofs_apply_step(ctx, step))) return rc;
                E_threshold = E0 + expected * step;
                stats->ls_bt_iterations++;
            }
            if (k == S->max_backtracking_armijo_iterations) { result = TooManyArmijoIterations; break; }
        }
    }

    if (result == Successful && !S->skip_converged_state_check) {
        if ((rc = state_valid(valid))) return rc;
        if (!valid) result = InvalidConvergedState;
    }
    stats->result = result;
    cudaEventRecord(ev1, ctx->stream);
    cudaEventSynchronize(ev1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    stats->gpu_ms = ms;
    stats->newton_iterations = it;
    timeline_mark(ctx, -1);
    static const char* tl = std::getenv("SB_TIMELINE");   // "1": every solve, "N": only solves with at least N iterations
    if (tl && it >= std::atoi(tl)) timeline_dump(ctx, stage_names()); else if (tl) timeline_dump(ctx, nullptr);
    return SB_OK;
}
