// Drop-in replacement of symx/src/solver/NewtonsMethod.cpp over the stark_b200 C-ABI (include/stark_b200.h).
//
// Compiled AGAINST THE REFERENCE'S OWN HEADERS and linked INSTEAD OF the reference's NewtonsMethod.cpp: same class, same public
// members (`create`, `solve`, `settings`, `callbacks`, `get_last_solve_stats`, `print_summary` -- symx/src/solver/
// NewtonsMethod.h:79-87), so core::Stark (stark/src/core/Stark.cpp:158, 176, 278, 295-296), every model and every user of the
// reference recompile unchanged.  integration/Makefile.shim builds the reference's unmodified sources with this file in place of
// that one translation unit (outputs under oracle/_ref/shim/, git-ignored) and runs the reference's own Catch2 suite
// (tests/rb_constraints.cpp) and scenes on the GPU path.
//
// What happens where:
//   * constructor: walks GlobalPotential exactly as SecondOrderCompiledGlobal does (SecondOrderCompiledGlobal.cpp:9-70):
//     DoF maps -> sb_dof_add in order; every potential's MappedWorkspace::maps -> one device array per distinct bound container
//     (DataMap::id) and the potential's fetch table {array, connectivity_index, first_symbol_idx, stride}; the kernel is looked
//     up by the potential's NAME (sb_potential_create).  A potential without a built-in kernel (a user's add_potential) is
//     differentiated here with symx's own symbolic engine and its operation sequences go to the library's Sequence -> CUDA -> NVRTC
//     back-end (sb_potential_create_user).
//   * solve(): the reference's control flow (NewtonsMethod.cpp:28-252, 254-371, 388-457, 459-641) on device-resident state, stage
//     by stage through the C-ABI.  User / model callbacks run on the host exactly where the reference runs them; around them the
//     DoFs are written back into the model's arrays (GlobalPotential::set_dofs) and every bound array / connectivity the callbacks
//     may have touched is compared with what the device holds and re-sent if it changed (the DataMap lambdas are re-evaluated
//     every time, as CompiledInLoop::run does).  Conditional potentials (`return {E, cond}`) are filtered on the host with the
//     reference's own expression interpreter (Scalar::eval).
//   * errors: C-ABI status codes become the reference's "print and exit(-1)"; numerical failures become SolverReturn values.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <vector>

// (read access to SolverCallbacks' private lists: empty lists let the shim skip the host round trip around an evaluation; a
//  maintainer would add `bool empty() const` accessors instead)
#include <algorithm>
#include <functional>
#include <iomanip>
#include <sstream>
#include <Eigen/Dense>
#include "Context.h"   // (everything solver_utils.h includes is included first: the access widening below must not reach other headers)
#define private public
#include "solver_utils.h"
#undef private
#include "NewtonsMethod.h"
#include "../symbol/diff.h"            // gradient / symmetric Hessian of a user potential's energy (symx's own symbolic engine)
#include "../symbol/utils.h"           // collect_scalars
#include "../compile/Sequence.h"       // the operation sequence the reference's code generator would print
#include <fmt/format.h>

#include "stark_b200.h"

using namespace symx;

namespace {

[[noreturn]] void die(const std::string& msg)
{
    std::cout << "symx error (stark_b200 shim): " << msg << std::endl;
    exit(-1);
}

struct ArrayRec {
    std::uintptr_t id = 0;
    int handle = -1;
    int stride = 1;
    std::function<const double*()> data;
    std::function<int32_t()> n_elements;
    std::vector<double> shadow;   // what the device holds
    bool is_dof = false;
    bool uploaded = false;
};
struct PotRec {
    const Potential* pot = nullptr;
    int handle = -1;
    int conn_stride = 0;
    std::vector<int32_t> conn_shadow;     // what the device holds (after filtering)
    std::vector<int32_t> conn_filtered;   // scratch
    bool uploaded = false;
};
struct Gpu {
    sb_context* ctx = nullptr;
    std::vector<ArrayRec> arrays;
    std::unordered_map<std::uintptr_t, int> array_of_id;
    std::vector<PotRec> pots;
    std::vector<double> u;   // flat DoF scratch
    long long h2d_bytes = 0, d2h_bytes = 0;

    void check(int status, const char* what) const
    {
        if (status != SB_OK) die(std::string(what) + ": " + sb_last_error(ctx));
    }
    int array_for(const DataMap<const double>& m)
    {
        const std::uintptr_t id = m.id();
        auto it = array_of_id.find(id);
        if (it != array_of_id.end()) {
            if (arrays[it->second].stride != m.stride) die("one container is bound with two different strides");
            return it->second;
        }
        ArrayRec a;
        a.id = id; a.stride = m.stride; a.data = m.data; a.n_elements = m.n_elements;
        check(sb_array_create(ctx, ("symx_" + std::to_string(arrays.size())).c_str(), m.stride, &a.handle), "sb_array_create");
        arrays.push_back(a);
        array_of_id[id] = (int)arrays.size() - 1;
        return (int)arrays.size() - 1;
    }
    // host -> device for everything that differs from what the device holds (the DataMap lambdas are evaluated now)
    void push_arrays(bool include_dofs)
    {
        for (ArrayRec& a : arrays) {
            if (a.is_dof && !include_dofs && a.uploaded) continue;
            const int n = a.n_elements();
            const size_t len = (size_t)n * a.stride;
            const double* p = a.data();
            if (a.uploaded && a.shadow.size() == len && (len == 0 || std::memcmp(a.shadow.data(), p, len * sizeof(double)) == 0)) continue;
            a.shadow.assign(p, p + len);
            check(sb_array_upload(ctx, a.handle, p, n), "sb_array_upload");
            a.uploaded = true;
            h2d_bytes += (long long)len * 8;
        }
    }
    void push_connectivity()
    {
        for (PotRec& r : pots) {
            auto mws = r.pot->get_mws();
            const int n = mws->conn.n_elements();
            const int32_t* c = mws->conn.data();
            const int32_t* use = c;
            int n_use = n;
            if (r.pot->has_conditional() && n > 0) {
                // SecondOrderCompiledPotential::_evaluate_element_condition with the reference's expression interpreter
                const Scalar cond = r.pot->get_condition();
                r.conn_filtered.clear();
                for (int e = 0; e < n; e++) {
                    for (const auto& m : mws->maps) {
                        const int row = (m.connectivity_index >= 0) ? c[(size_t)e * r.conn_stride + m.connectivity_index] : 0;
                        const double* d = m.data() + (size_t)row * m.stride;
                        for (int k = 0; k < m.stride; k++) mws->ws.get_scalar(m.first_symbol_idx + k).set_value(d[k]);
                    }
                    if (cond.eval() > 0.0) r.conn_filtered.insert(r.conn_filtered.end(), c + (size_t)e * r.conn_stride, c + (size_t)(e + 1) * r.conn_stride);
                }
                use = r.conn_filtered.data();
                n_use = (int)(r.conn_filtered.size() / r.conn_stride);
            }
            const size_t len = (size_t)n_use * r.conn_stride;
            if (r.uploaded && r.conn_shadow.size() == len && (len == 0 || std::memcmp(r.conn_shadow.data(), use, len * sizeof(int32_t)) == 0)) continue;
            r.conn_shadow.assign(use, use + len);
            check(sb_potential_set_connectivity(ctx, r.handle, use, n_use), "sb_potential_set_connectivity");
            r.uploaded = true;
            h2d_bytes += (long long)len * 4;
        }
    }
};

std::unordered_map<const NewtonsMethod*, std::unique_ptr<Gpu>>& registry()
{
    static std::unordered_map<const NewtonsMethod*, std::unique_ptr<Gpu>> r;
    return r;
}

bool no_callbacks(const SolverCallbacks& c)
{
    return c.before_energy_evaluation.empty() && c.is_initial_state_valid.empty() && c.is_intermediate_state_valid.empty() && c.on_intermediate_state_invalid.empty() &&
           c.on_armijo_fail.empty() && c.is_converged.empty() && c.is_converged_state_valid.empty() && c.max_allowed_step.empty();
}
bool default_residual_in_use(const SolverCallbacks& c) { return c.residual.target_type() == typeid(default_residual); }

}  // namespace

symx::NewtonsMethod::NewtonsMethod(spGlobalPotential global_potential, spContext context, spSolverCallbacks callbacks)
    : global_potential(global_potential), context(context), callbacks(callbacks)
{
    this->output = context->output;
    this->logger = context->logger;
    if (callbacks == nullptr) this->callbacks = std::make_shared<SolverCallbacks>(context);

    auto gpu = std::make_unique<Gpu>();
    int device = 0;
    if (const char* d = std::getenv("STARK_B200_DEVICE")) device = std::atoi(d);
    if (sb_create(&gpu->ctx, device, nullptr) != SB_OK) die("sb_create failed: no usable CUDA device (this back-end has no CPU fallback)");

    // one device array per distinct container bound by any potential (strides come from these bindings: DoF maps are flat)
    for (const auto& pot : global_potential->get_potentials())
        for (const auto& m : pot->get_mws()->maps) gpu->array_for(m);
    // DoFs, in registration order (GlobalPotential::get_dofs_offsets); a DoF map is the flat view (stride 1) of a container of 3-vectors
    for (const auto& m : global_potential->get_dof_maps()) {
        int a;
        auto it = gpu->array_of_id.find(m.id());
        if (it != gpu->array_of_id.end()) a = it->second;
        else {   // a DoF set no potential reads: still part of the DoF vector
            DataMap<const double> cm(m.id, [m]() { return (const double*)m.data(); }, [m]() { return m.n_elements() * m.stride / 3; }, 3, -1, 0);
            a = gpu->array_for(cm);
        }
        if (gpu->arrays[a].stride != 3) die("a DoF container must hold 3-vectors (block size 3)");
        gpu->arrays[a].is_dof = true;
        int set = -1;
        gpu->check(sb_dof_add(gpu->ctx, gpu->arrays[a].handle, &set), "sb_dof_add");
    }
    // potentials
    for (const auto& pot : global_potential->get_potentials()) {
        auto mws = pot->get_mws();
        std::vector<sb_fetch> fetch;
        for (const auto& m : mws->maps) {
            sb_fetch f;
            f.array = gpu->arrays[gpu->array_for(m)].handle;
            f.conn_col = m.connectivity_index;
            f.first_slot = m.first_symbol_idx;
            f.stride = m.stride;
            fetch.push_back(f);
        }
        PotRec r;
        r.pot = pot.get();
        r.conn_stride = mws->conn.stride;
        int status = sb_potential_create(gpu->ctx, pot->get_name().c_str(), r.conn_stride, fetch.data(), (int)fetch.size(), &r.handle);
        if (status == SB_ERR_NO_KERNEL) {
            // a user potential: differentiate it as SecondOrderCompiledPotential does (SecondOrderCompiledPotential.cpp:9-80) and hand
            // the operation sequences of [E] and [E | grad | hess] to the library's Sequence -> CUDA -> NVRTC back-end
            std::vector<Scalar> dofs;
            std::vector<int32_t> block_slots;
            for (const auto& dof_map : global_potential->get_dof_maps()) {
                const std::vector<Scalar> set_dofs = mws->get_symbols(dof_map);
                if (set_dofs.size() % 3 != 0) die("user potential '" + pot->get_name() + "': DoF symbols do not come in 3-vectors");
                for (size_t k = 0; k < set_dofs.size(); k += 3) block_slots.push_back(set_dofs[k].get_symbol_idx());
                dofs.insert(dofs.end(), set_dofs.begin(), set_dofs.end());
            }
            if (dofs.empty()) die("user potential '" + pot->get_name() + "' has no degrees of freedom");
            const Scalar v = pot->get_expression();
            DiffCache diff_cache;
            const Vector g = gradient(v, dofs, diff_cache);
            const Matrix h = gradient(g, dofs, /*symmetric=*/true, diff_cache);
            Sequence seq_p({v});
            Sequence seq_pgh(collect_scalars({{v}, g.values(), h.values()}));
            auto to_ops = [](const Sequence& seq) {
                std::vector<sb_op> ops(seq.ops.size());
                for (size_t k = 0; k < seq.ops.size(); k++) {
                    const core::Op& o = seq.ops[k];
                    ops[k].type = (int32_t)o.type; ops[k].dst = o.dst; ops[k].a = o.a; ops[k].b = o.b; ops[k].cond = o.cond; ops[k].pad = 0; ops[k].constant = o.constant;
                }
                return ops;
            };
            const std::vector<sb_op> ops_p = to_ops(seq_p), ops_pgh = to_ops(seq_pgh);
            status = sb_potential_create_user(gpu->ctx, pot->get_name().c_str(), r.conn_stride, fetch.data(), (int)fetch.size(), seq_pgh.get_n_inputs(),
                                              (int)block_slots.size(), block_slots.data(), ops_p.data(), (int)ops_p.size(), ops_pgh.data(), (int)ops_pgh.size(), &r.handle);
        }
        gpu->check(status, ("sb_potential_create(" + pot->get_name() + ")").c_str());
        gpu->pots.push_back(r);
    }
    registry()[this] = std::move(gpu);
}

spNewtonsMethod symx::NewtonsMethod::create(spGlobalPotential global_potential, spContext context, spSolverCallbacks callbacks)
{
    return std::make_shared<NewtonsMethod>(global_potential, context, callbacks);
}

SolverReturn NewtonsMethod::solve()
{
    Gpu& G = *registry().at(this);
    sb_context* ctx = G.ctx;
    const int ndofs = global_potential->get_total_n_dofs();
    if (ndofs <= 0) { std::cout << "symx error NewtonsMethod::solve(): No degrees of freedom." << std::endl; exit(1); }
    if (ndofs % 3 != 0) { std::cout << "symx error NewtonsMethod::solve(): ndofs must be divisible by 3." << std::endl; exit(1); }
    const NewtonSettings& S = this->settings;
    SolverCallbacks& cb = *this->callbacks;
    const bool host_hooks = !no_callbacks(cb);
    const bool plain_residual = default_residual_in_use(cb);

    this->stats = SolveStats();
    this->grad.resize(ndofs);
    G.u.resize(ndofs);
    double E0 = 0.0, du_dot_grad = 0.0, res_0 = std::numeric_limits<double>::max();
    SolverReturn result = SolverReturn::Running;
    this->pdn_countdown = 0;
    this->ppn_threshold = -1.0;

    // the model's arrays as they are now (initial guess, parameters, connectivity) -> device
    G.push_arrays(true);
    G.push_connectivity();

    // DoFs device -> model arrays, so that host callbacks see the state the device is at
    auto dofs_to_host = [&]() {
        if (!host_hooks) return;
        G.check(sb_dofs_get(ctx, G.u.data()), "sb_dofs_get");
        G.d2h_bytes += (long long)ndofs * 8;
        global_potential->set_dofs(G.u.data());
        // (the DoF arrays' shadows follow, so that push_arrays does not send them back)
        for (ArrayRec& a : G.arrays) if (a.is_dof) { const double* p = a.data(); a.shadow.assign(p, p + (size_t)a.n_elements() * a.stride); }
    };
    // everything a callback may have changed -> device
    auto host_to_device = [&]() {
        if (!host_hooks) return;
        G.push_arrays(false);
        G.push_connectivity();
    };
    auto evaluate = [&](bool with_derivatives, double& E, double& residual) {
        dofs_to_host();
        cb.run_before_energy_evaluation();
        host_to_device();
        double r_inf = 0.0;
        auto _t = this->logger->time(with_derivatives ? "evaluate_P_grad_hess" : "evaluate_P");
        G.check(sb_eval(ctx, with_derivatives ? SB_EVAL_PGH : SB_EVAL_P, &E, &r_inf), "sb_eval");
        if (!with_derivatives) return;
        if (plain_residual) residual = r_inf;
        else {
            G.check(sb_grad_get(ctx, this->grad.data()), "sb_grad_get");
            G.d2h_bytes += (long long)ndofs * 8;
            residual = cb.compute_residual(this->grad);
        }
    };

    dofs_to_host();
    if (!cb.run_is_initial_state_valid()) {
        this->output->print_with_new_line("Newton failure: Invalid initial state.", Verbosity::Medium);
        result = SolverReturn::InvalidInitialState;
    }

    int it = -1;
    while (result == SolverReturn::Running) {
        it++;
        if (it == S.max_iterations) {
            this->output->print_with_new_line("Newton failure: Too many iterations.", Verbosity::Medium);
            result = S.max_iterations_as_success ? SolverReturn::Successful : SolverReturn::TooManyIterations;
            break;
        }
        this->output->print_with_new_line(fmt::format("{:2d}. ", it), Verbosity::Medium);
        double residual = 0.0;
        evaluate(true, E0, residual);
        this->output->print(fmt::format("r0: {:.2e} | ", residual), Verbosity::Medium);
        if (it == 0) res_0 = residual;
        if (residual < S.bailout_residual) { result = SolverReturn::Successful; break; }
        if (it >= S.min_iterations) {
            if (residual < S.residual_tolerance_abs) { result = SolverReturn::Successful; break; }
            if (it > 0 && residual / res_0 < S.residual_tolerance_rel) { result = SolverReturn::Successful; break; }
        }

        // ---- project + assemble + solve until the direction descends (NewtonsMethod.cpp:137-182, 254-371) ----
        bool assembled = false, solved = false;
        double du_inf = 0.0;
        int64_t n_proj = 0, n_hess = 0;
        const int cg_before = this->stats.cg_iterations;
        while (!solved) {
            bool all_projected = false, projected_now = false;
            int allp = 0;
            {
                auto _t = this->logger->time("project_to_PD");
                switch (S.projection_mode) {
                case ProjectionToPD::Newton: break;
                case ProjectionToPD::ProjectedNewton:
                    G.check(sb_project_to_pd(ctx, 0.0, S.projection_eps, S.project_to_pd_use_mirroring, &n_proj, &n_hess, &allp), "sb_project_to_pd");
                    all_projected = true; projected_now = true;
                    break;
                case ProjectionToPD::ProjectOnDemand:
                    if (this->pdn_countdown > 0) {
                        G.check(sb_project_to_pd(ctx, 0.0, S.projection_eps, S.project_to_pd_use_mirroring, &n_proj, &n_hess, &allp), "sb_project_to_pd");
                        all_projected = true; projected_now = true;
                    }
                    break;
                case ProjectionToPD::Progressive:
                    if (this->ppn_threshold > 0.0) {
                        if (this->ppn_threshold < 1e-12) this->ppn_threshold = 0.0;
                        G.check(sb_project_to_pd(ctx, this->ppn_threshold, S.projection_eps, S.project_to_pd_use_mirroring, &n_proj, &n_hess, &allp), "sb_project_to_pd");
                        all_projected = (allp != 0); projected_now = true;
                    }
                    break;
                default: std::cout << "Error: Unknown projection mode." << std::endl; exit(1);
                }
            }
            if (!assembled || projected_now) {
                auto _t = this->logger->time("assembly");
                G.check(sb_assemble(ctx), "sb_assemble");
                assembled = true;
            }
            // _solve_linear_system (NewtonsMethod.cpp:388-457)
            int cg_it = 0, ok = 0;
            {
                auto _t = this->logger->time("linear_system_solve");
                if (S.linear_solver == LinearSolver::DirectLLT) G.check(sb_solve_llt(ctx, &ok, &du_dot_grad, &du_inf), "sb_solve_llt");
                else {
                    const double forcing = std::min(1e-2, residual * std::min(0.5, std::sqrt(residual)));
                    const double abs_tol = std::max(forcing, S.cg_abs_tolerance);
                    G.check(sb_solve_pcg(ctx, abs_tol, S.cg_rel_tolerance, S.cg_max_iterations, S.cg_stop_on_indefiniteness, &cg_it, &ok, &du_dot_grad, &du_inf), "sb_solve_pcg");
                }
            }
            this->last_cg_iterations = cg_it;
            this->stats.cg_iterations += cg_it;
            const bool can_project_more = (S.projection_mode != ProjectionToPD::Newton) && !all_projected;
            if (!ok) {
                if (!can_project_more) { this->output->print("Linear system failed. ", Verbosity::Summary); result = SolverReturn::LinearSystemSolveFailure; break; }
            } else {
                if (du_dot_grad < 0.0) { solved = true; break; }
                if (!can_project_more) { this->output->print("Step does not descend. ", Verbosity::Summary); result = SolverReturn::StepDoesNotDescend; break; }
            }
            // _increase_projection
            if (S.projection_mode == ProjectionToPD::ProjectOnDemand) this->pdn_countdown = S.project_on_demand_countdown;
            else if (S.projection_mode == ProjectionToPD::Progressive) {
                if (this->ppn_threshold < 0.0) this->ppn_threshold = residual;   // grad.cwiseAbs().maxCoeff() (the default residual)
                this->ppn_threshold *= S.ppn_tightening_factor;
            }
        }
        if (result != SolverReturn::Running) {
            this->output->print_with_new_line("Newton failure: Could not solve the linear system or find a descend direction.", Verbosity::Summary);
            break;
        }
        // _decrease_projection
        if (S.projection_mode == ProjectionToPD::ProjectOnDemand) this->pdn_countdown--;
        else if (S.projection_mode == ProjectionToPD::Progressive) this->ppn_threshold *= S.ppn_release_factor;
        {
            int64_t np = 0, nh = 0; int dummy = 0;
            G.check(sb_project_to_pd(ctx, -1.0, S.projection_eps, 0, &np, &nh, &dummy), "sb_project_to_pd");   // (threshold < 0: counts only)
            this->stats.n_hessians += (uint64_t)nh;
            this->stats.n_projected_hessians += (uint64_t)np;
            this->output->print(fmt::format("ph: {:4.1f}% | #CG: {:4d} | ", nh ? 100.0 * (double)np / (double)nh : 0.0, this->stats.cg_iterations - cg_before), Verbosity::Medium);
            logger->add_and_append("n_hessians", (double)nh);
            logger->add_and_append("n_projected_hessians", (double)np);
            logger->add_and_append("cg_iterations", this->last_cg_iterations);
        }
        this->output->print(fmt::format("du: {:.1e} | ", du_inf), Verbosity::Medium);
        if (it >= S.min_iterations && du_inf < S.step_tolerance) { result = SolverReturn::Successful; break; }

        // ---- line search (NewtonsMethod.cpp:459-641): cap / max / invalid-state halving / Armijo ----
        G.check(sb_dofs_save(ctx), "sb_dofs_save");
        double retraction = 1.0;
        if (du_inf > S.step_cap) {
            retraction *= S.step_cap / du_inf;
            G.check(sb_du_scale(ctx, S.step_cap / du_inf), "sb_du_scale");
            du_inf = S.step_cap;
            this->stats.ls_cap_iterations++;
        }
        const double max_step = cb.run_max_allowed_step();
        if (max_step < 1.0) {
            retraction *= max_step;
            G.check(sb_du_scale(ctx, max_step), "sb_du_scale");
            du_inf *= max_step;
            this->stats.ls_max_iterations++;
        }
        double step = 1.0;
        G.check(sb_dofs_apply_step(ctx, step), "sb_dofs_apply_step");
        int ls_inv = 0;
        for (; ls_inv < S.max_backtracking_invalid_state_iterations; ++ls_inv) {
            dofs_to_host();
            if (cb.run_is_intermediate_state_valid()) break;
            step *= 0.5;
            G.check(sb_dofs_apply_step(ctx, step), "sb_dofs_apply_step");
            this->stats.ls_inv_iterations++;
        }
        if (ls_inv == S.max_backtracking_invalid_state_iterations) {
            this->output->print_with_new_line("Newton failure: Too many invalid intermediate states in the line search.", Verbosity::Medium);
            cb.run_on_intermediate_state_invalid();
            result = SolverReturn::TooManyInvalidIntermediateIterations;
            break;
        }
        if (S.enable_armijo_backtracking) {
            const double expected = S.line_search_armijo_beta * du_dot_grad * retraction;
            double E_threshold = E0 + expected * step, E1 = 0.0, unused = 0.0;
            int k = 0;
            for (; k < S.max_backtracking_armijo_iterations; ++k) {
                evaluate(false, E1, unused);
                if (E1 < E_threshold) break;
                step *= 0.5;
                G.check(sb_dofs_apply_step(ctx, step), "sb_dofs_apply_step");
                E_threshold = E0 + expected * step;
                this->stats.ls_bt_iterations++;
            }
            if (k == S.max_backtracking_armijo_iterations) {
                this->output->print_with_new_line("Newton failure: Too many Armijo iterations.", Verbosity::Medium);
                cb.run_on_armijo_fail();
                result = SolverReturn::TooManyArmijoIterations;
                break;
            }
        }
        if (it >= S.min_iterations && cb.run_is_converged()) { result = SolverReturn::Successful; break; }
    }

    // the solution goes back into the model's arrays (the reference leaves it there through set_dofs)
    G.check(sb_dofs_get(ctx, G.u.data()), "sb_dofs_get");
    G.d2h_bytes += (long long)ndofs * 8;
    global_potential->set_dofs(G.u.data());
    for (ArrayRec& a : G.arrays) if (a.is_dof) { const double* p = a.data(); a.shadow.assign(p, p + (size_t)a.n_elements() * a.stride); }

    if (result == SolverReturn::Successful && !cb.run_is_converged_state_valid()) {
        this->output->print_with_new_line("Newton failure: Invalid converged state.", Verbosity::Medium);
        result = SolverReturn::InvalidConvergedState;
    }
    this->stats.newton_iterations = it;
    if (this->stats.n_hessians > 0) this->stats.projected_hessians_ratio = (double)this->stats.n_projected_hessians / (double)this->stats.n_hessians;
    logger->add_and_append("newton_iterations", stats.newton_iterations);
    return result;
}

void NewtonsMethod::print_summary(double total_time) const
{
    const Gpu& G = *registry().at(this);
    this->output->print_with_new_line(fmt::format("NewtonsMethod (stark_b200 back-end): {} kernels launched, {:.1f} MB host->device, {:.1f} MB device->host{}",
        (long long)sb_launch_count(G.ctx), 1e-6 * (double)G.h2d_bytes, 1e-6 * (double)G.d2h_bytes, total_time >= 0.0 ? fmt::format(", total {:.3f} s", total_time) : ""), Verbosity::Summary);
}
