// extern "C" scene-level entry points over the C++ host layer, for bench.py and the end-to-end tests (ctypes).
// The scenes are the BASELINE.json configurations, built exactly like oracle/ref_driver.cpp builds them through the
// reference's public API (same generators, parameters and order of add_* calls).
#include "stark_b200.hpp"

#include <algorithm>
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

// C1 / C3: n x n cloth grid (Cotton_Fabric) over a fixed, scripted rigid box (oracle/ref_driver.cpp scene_cloth;
// examples/main.cpp:371-414).  discrete_shells = true uses the dihedral-angle bending energy with friction mu.
void build_cloth(Scene& sc, int n, double dt, double drop, bool discrete_shells, double mu, int device, void* stream)
{
    Settings s;
    s.simulation.max_time_step_size = dt;
    s.device = device; s.stream = stream;
    sc.sim = std::make_unique<Simulation>(s);
    Simulation& sim = *sc.sim;
    EnergyFrictionalContact::GlobalParams cp;
    // below the mesh spacing (0.4 m / n): 2 mm up to n = 64, 0.3 edge lengths on finer grids (same rule as oracle/ref_driver.cpp)
    cp.default_contact_thickness = (n > 64) ? 0.3 * 0.4 / n : 0.002;
    sim.contact.set_global_params(cp);
    SurfaceParams material = SurfaceParams::Cotton_Fabric();
    if (discrete_shells) material.flat_rest_angle = false;
    auto cloth = sim.add_surface_grid({0.4, 0.4}, {n, n}, material);
    if (drop != 0.003) sim.dyn.add_displacement(cloth.point_set, {0.0, 0.0, drop});
    auto box = sim.add_box(1.0, {0.08, 0.08, 0.08});
    sim.set_translation(box.body, {0.0, 0.0, -0.08});
    const auto fix = sim.rb_constraints.add_fix(sim.rb, box.body);
    if (mu > 0.0) sim.contact.set_friction(box.contact_group, cloth.contact_group, mu);
    Simulation* ps = &sim;
    sim.add_time_event([ps, fix](double t) { ps->rb_constraints.set_fix_transformation(fix, {0.0, 0.0, -0.08 - 0.1 * std::sin(t)}, 90.0 * t, {0.0, 0.0, 1.0}); });
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
    else if (sc->name == "cloth") build_cloth(*sc, n, dt, drop, false, 0.0, device, stream);
    else if (sc->name == "cloth_shells") build_cloth(*sc, n, dt, drop, true, 0.3, device, stream);
    else { delete sc; return nullptr; }
    return sc;
}
__attribute__((visibility("default"))) void sbh_scene_destroy(void* h) { delete static_cast<Scene*>(h); }
// settings.newton.linear_solver: 0 DirectLLT, 1 BDPCG (symx::LinearSolver, solver_utils.h)
__attribute__((visibility("default"))) void sbh_scene_set_linear_solver(void* h, int kind) { static_cast<Scene*>(h)->sim->settings.newton.linear_solver = kind; }

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
// wait for the asynchronous read-backs of the last step (the host mirrors of x0 / v0 are then current)
__attribute__((visibility("default"))) void sbh_scene_sync(void* h) { static_cast<Scene*>(h)->sim->sync_host(); }
__attribute__((visibility("default"))) int sbh_scene_positions(void* h, double* x0, int n_nodes)
{
    Simulation& sim = *static_cast<Scene*>(h)->sim;
    if (n_nodes != sim.dyn.size()) return -1;
    sim.sync_host();
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
// host copy of a bound array by its label (e.g. "shells.rest_angle"); returns the number of doubles, -1 if unknown
__attribute__((visibility("default"))) int sbh_scene_array(void* h, const char* label, double* out, int cap)
{
    static_cast<Scene*>(h)->sim->sync_host();
    const DeviceArray* a = static_cast<Scene*>(h)->sim->find_array(label);
    if (!a) return -1;
    const int n = (int)a->data.size();
    if (out) std::memcpy(out, a->data.data(), sizeof(double) * (size_t)std::min(n, cap));
    return n;
}
// connectivity table of a deformable potential by name; returns the number of int32 entries, -1 if unknown
__attribute__((visibility("default"))) int sbh_scene_connectivity(void* h, const char* name, int32_t* out, int cap)
{
    Simulation& sim = *static_cast<Scene*>(h)->sim;
    const std::string n = name;
    const int32_t* p = nullptr;
    int len = 0;
    auto take = [&](const auto& conn) { len = (int)(conn.size() * sizeof(conn[0]) / sizeof(int32_t)); p = conn.empty() ? nullptr : conn[0].data(); };
    if (n == "EnergyDiscreteShells") take(sim.discrete_shells.conn_complete);
    else if (n == "EnergyBendingFlat") take(sim.discrete_shells.conn_flat_rest);
    else if (n == "EnergyTriangleStrain") take(sim.triangle_strain.conn_complete);
    else if (n == "EnergyTriangleStrain_Elasticity_Only") take(sim.triangle_strain.conn_elasticity_only);
    else if (n == "EnergyLumpedInertia") take(sim.lumped_inertia.conn);
    else if (n == "EnergyTetStrain") take(sim.tet_strain.conn_complete);
    else return -1;
    if (out && p) std::memcpy(out, p, sizeof(int32_t) * (size_t)std::min(len, cap));
    return len;
}
// ---- mesh helpers of the presets, callable without a GPU (CPU tests) ----
// triangle grid (generate_triangle_grid): returns the number of triangles; V gets 3 (n0+1)(n1+1) doubles, T 3 ints per triangle
__attribute__((visibility("default"))) int sbh_mesh_triangle_grid(int n0, int n1, double dx, double dy, double* V, int32_t* T)
{
    std::vector<Vec3> v;
    std::vector<std::array<int, 3>> t;
    generate_triangle_grid(v, t, {0.0, 0.0}, {dx, dy}, {n0, n1});
    if (V) for (size_t i = 0; i < v.size(); i++) for (int c = 0; c < 3; c++) V[3 * i + c] = v[i][c];
    if (T) for (size_t i = 0; i < t.size(); i++) for (int c = 0; c < 3; c++) T[3 * i + c] = t[i][c];
    return (int)t.size();
}
// hinges of a triangle mesh (find_internal_angles): returns their number; out gets 4 ints each {edge_0, edge_1, opp_0, opp_1}
__attribute__((visibility("default"))) int sbh_mesh_internal_angles(const int32_t* T, int n_tri, int n_nodes, int32_t* out, int cap)
{
    std::vector<std::array<int, 3>> t(n_tri);
    for (int i = 0; i < n_tri; i++) t[i] = {T[3 * i], T[3 * i + 1], T[3 * i + 2]};
    std::vector<std::array<int, 4>> a;
    find_internal_angles(a, t, n_nodes);
    if (out) for (size_t i = 0; i < a.size() && (int)i < cap; i++) for (int c = 0; c < 4; c++) out[4 * i + c] = a[i][c];
    return (int)a.size();
}
// tet grid (generate_tet_grid) and its boundary (find_surface): returns the number of tets; n_surface gets the triangle count
__attribute__((visibility("default"))) int sbh_mesh_tet_grid(int n0, int n1, int n2, double dx, double dy, double dz, double* V, int32_t* T, int* n_vertices, int* n_surface)
{
    std::vector<Vec3> v;
    std::vector<std::array<int, 4>> t;
    generate_tet_grid(v, t, {0.0, 0.0, 0.0}, {dx, dy, dz}, {n0, n1, n2});
    if (V) for (size_t i = 0; i < v.size(); i++) for (int c = 0; c < 3; c++) V[3 * i + c] = v[i][c];
    if (T) for (size_t i = 0; i < t.size(); i++) for (int c = 0; c < 4; c++) T[4 * i + c] = t[i][c];
    if (n_vertices) *n_vertices = (int)v.size();
    if (n_surface) {
        std::vector<std::array<int, 3>> tris;
        std::vector<int> map;
        find_surface(tris, map, v, t);
        *n_surface = (int)tris.size();
    }
    return (int)t.size();
}
__attribute__((visibility("default"))) void* sbh_scene_context(void* h) { return static_cast<Scene*>(h)->sim->context(); }

}  // extern "C"
