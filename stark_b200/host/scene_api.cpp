// extern "C" scene-level entry points over the C++ host layer, for bench.py and the end-to-end tests (ctypes).
// The scenes are the BASELINE.json configurations, built exactly like oracle/ref_driver.cpp builds them through the
// reference's public API (same generators, parameters and order of add_* calls).
#include "stark_b200.hpp"

#include <cmath>
#include <cstring>
#include <memory>

using namespace stark_b200;

namespace {
struct Scene {
    std::unique_ptr<Simulation> sim;
    std::string name;
};

// C2: n^3 tet grid (Soft_Rubber) dropped onto a fixed rigid floor box, IPC contact + friction (oracle/ref_driver.cpp scene_tetdrop)
void build_tetdrop(Scene& sc, int n, double dt, double drop, double vz, int device, void* stream)
{
    Settings s;
    s.simulation.max_time_step_size = dt;
    s.device = device; s.stream = stream;
    sc.sim = std::make_unique<Simulation>(s);
    Simulation& sim = *sc.sim;
    EnergyFrictionalContact::GlobalParams cp;
    cp.default_contact_thickness = 0.001;
    cp.min_contact_stiffness = 1e8;
    sim.contact.set_global_params(cp);
    auto H = sim.add_volume_grid({1.0, 1.0, 1.0}, {n, n, n}, VolumeParams::Soft_Rubber());
    sim.dyn.add_displacement(H.point_set, {0.0, 0.0, 0.5 + drop});
    if (vz != 0.0) sim.dyn.set_velocity(H.point_set, {0.0, 0.0, -vz});
    auto floor = sim.add_box(1.0, {4.0, 4.0, 0.1});
    sim.set_translation(floor.body, {0.0, 0.0, -0.05});
    sim.rb_constraints.add_fix(sim.rb, floor.body);
    sim.contact.set_friction(floor.contact_group, H.contact_group, 0.5);
}

// C5: tet bar, both end caps prescribed, one cap rotating 90 deg/s, no contact (oracle/ref_driver.cpp scene_tetbar)
void build_tetbar(Scene& sc, int nx, int ny, int nz, double dt, int device, void* stream)
{
    Settings s;
    s.simulation.max_time_step_size = dt;
    s.simulation.init_frictional_contact = false;
    s.device = device; s.stream = stream;
    sc.sim = std::make_unique<Simulation>(s);
    Simulation& sim = *sc.sim;
    if (ny <= 0) ny = nx;
    if (nz <= 0) nz = 8 * nx;
    const double h = 1.0 / 22.0;
    const Vec3 dim = {nx * h, ny * h, nz * h};
    auto H = sim.add_volume_grid(dim, {nx, ny, nz}, VolumeParams::Soft_Rubber());
    const double hz = 0.5 * dim[2];
    const double inf = std::numeric_limits<double>::max();
    sim.prescribed_positions.add_inside_aabb(sim.dyn, H.point_set, {0.0, 0.0, -hz}, {dim[0], dim[1], 0.001}, 1e3, inf);
    const int g1 = sim.prescribed_positions.add_inside_aabb(sim.dyn, H.point_set, {0.0, 0.0, hz}, {dim[0], dim[1], 0.001}, 1e3, inf);
    Simulation* ps = &sim;
    sim.add_time_event([ps, g1](double t) { ps->prescribed_positions.set_transformation(g1, {0.0, 0.0, 0.0}, 90.0 * t, {0.0, 0.0, 1.0}); });
}

// C4: n^3 tet grid with a prescribed bottom face under a chain of nb hinged rigid boxes, first box fixed
// (oracle/ref_driver.cpp scene_tetchain)
void build_tetchain(Scene& sc, int n, int nb, double dt, double drop, int device, void* stream)
{
    Settings s;
    s.simulation.max_time_step_size = dt;
    s.device = device; s.stream = stream;
    sc.sim = std::make_unique<Simulation>(s);
    Simulation& sim = *sc.sim;
    EnergyFrictionalContact::GlobalParams cp;
    cp.default_contact_thickness = 0.001;
    cp.min_contact_stiffness = 1e7;
    sim.contact.set_global_params(cp);
    VolumeParams material = VolumeParams::Soft_Rubber();
    material.density = 50.0;   // light foam: the block keeps its shape under its own weight
    auto H = sim.add_volume_grid({1.0, 1.0, 1.0}, {n, n, n}, material);
    sim.dyn.add_displacement(H.point_set, {0.0, 0.0, 0.5});
    sim.prescribed_positions.add_inside_aabb(sim.dyn, H.point_set, {0.0, 0.0, 0.0}, {1.0, 1.0, 0.001}, 1e7, std::numeric_limits<double>::max());
    if (nb <= 0) nb = 10;
    const double z = 1.0 + 0.04 + drop + 0.04;
    Simulation::BoxHandle prev{-1, -1};
    for (int i = 0; i < nb; i++) {
        auto b = sim.add_box(1.0, {0.08, 0.08, 0.08});
        sim.set_translation(b.body, {-0.2 + 0.1 * i, 0.0, z});
        sim.contact.set_friction(b.contact_group, H.contact_group, 0.3);
        if (i == 0) sim.rb_constraints.add_fix(sim.rb, b.body);
        else {
            sim.rb_constraints.add_hinge(sim.rb, prev.body, b.body, {-0.25 + 0.1 * i, 0.0, z}, {0.0, 1.0, 0.0});
            sim.contact.disable_collision(prev.contact_group, b.contact_group);
        }
        prev = b;
    }
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) void* sbh_scene_create(const char* name, int n, int ny, int nz, double dt, double drop, double vz, int device, void* stream)
{
    auto* sc = new Scene();
    sc->name = name;
    if (sc->name == "tetdrop") build_tetdrop(*sc, n, dt, drop, vz, device, stream);
    else if (sc->name == "tetbar") build_tetbar(*sc, n, ny, nz, dt, device, stream);
    else if (sc->name == "tetchain") build_tetchain(*sc, n, ny, dt, drop, device, stream);
    else { delete sc; return nullptr; }
    return sc;
}
__attribute__((visibility("default"))) void sbh_scene_destroy(void* h) { delete static_cast<Scene*>(h); }

// out[0] keep_going, [1] accepted, [2] result, [3] newton its, [4] cg its, [5] evaluations, [6] dt, [7] runtime s, [8] solve s,
// [9] first residual, [10] ls_inv, [11] ls_bt, [12] time, [13] ndofs, [14] contact stiffness, [15] device ms of the Newton solve
__attribute__((visibility("default"))) int sbh_scene_step(void* h, double* out)
{
    Simulation& sim = *static_cast<Scene*>(h)->sim;
    const bool keep = sim.run_one_time_step();
    const StepStats& s = sim.last_step();
    if (out) {
        out[0] = keep; out[1] = s.accepted; out[2] = s.result; out[3] = s.newton_iterations; out[4] = s.cg_iterations; out[5] = s.n_evaluations;
        out[6] = s.dt; out[7] = s.runtime_s; out[8] = s.solve_s; out[9] = s.first_residual; out[10] = s.ls_inv; out[11] = s.ls_bt; out[12] = sim.current_time;
        out[13] = sim.ndofs(); out[14] = sim.contact.contact_stiffness; out[15] = s.solve_gpu_ms;
    }
    return keep ? 1 : 0;
}
__attribute__((visibility("default"))) int sbh_scene_residuals(void* h, double* out, int cap)
{
    const StepStats& s = static_cast<Scene*>(h)->sim->last_step();
    const int n = std::min<int>(cap, (int)s.residuals.size());
    for (int i = 0; i < n; i++) out[i] = s.residuals[i];
    return n;
}
// out[0] nodes, [1] tets, [2] ndofs, [3] h2d bytes so far, [4] d2h bytes so far, [5] kernel launches so far, [6] total newton its, [7] total solve s
__attribute__((visibility("default"))) void sbh_scene_totals(void* h, double* out)
{
    Simulation& sim = *static_cast<Scene*>(h)->sim;
    out[0] = sim.dyn.size(); out[1] = (double)(sim.tet_strain.conn_complete.size() + sim.tet_strain.conn_elasticity_only.size()); out[2] = sim.ndofs();
    out[3] = (double)sim.h2d_bytes; out[4] = (double)sim.d2h_bytes; out[5] = (double)sb_launch_count(sim.context());
    out[6] = (double)sim.total_newton_iterations; out[7] = sim.total_solve_s;
}
__attribute__((visibility("default"))) int sbh_scene_positions(void* h, double* x0, int n_nodes)
{
    Simulation& sim = *static_cast<Scene*>(h)->sim;
    if (n_nodes != sim.dyn.size()) return -1;
    std::memcpy(x0, sim.dyn.x0.data.data(), sizeof(double) * 3 * (size_t)n_nodes);
    return 0;
}
__attribute__((visibility("default"))) int sbh_scene_potential(void* h, const char* name)
{
    Simulation& sim = *static_cast<Scene*>(h)->sim;
    const std::string n = name;
    if (n == "EnergyTetStrain") return sim.tet_strain.potential_complete;
    if (n == "EnergyTetStrain_Elasticity_Only") return sim.tet_strain.potential_elasticity_only;
    if (n == "EnergyLumpedInertia") return sim.lumped_inertia.potential;
    return -1;
}
__attribute__((visibility("default"))) void* sbh_scene_context(void* h) { return static_cast<Scene*>(h)->sim->context(); }

}  // extern "C"
