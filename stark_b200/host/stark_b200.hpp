// Host-side mirror (C++) of the part of STARK's public interface that drives the Newton hot path.
//
// Everything here sits ABOVE the C-ABI (include/stark_b200.h): it owns host state, binds arrays / connectivity /
// potentials to the device library exactly the way the reference's models bind symbols through MappedWorkspace, and
// runs the time-step loop of stark::core::Stark.  Class and member names follow the reference so that the parity tests
// read like the reference's own scenes:
//   stark::Settings            S/core/Settings.h:10-49            -> stark_b200::Settings
//   stark::core::Stark         S/core/Stark.{h,cpp}               -> stark_b200::Stark (run_one_step, adaptive dt, retries)
//   stark::PointDynamics       S/models/deformables/PointDynamics.{h,cpp}
//   stark::EnergyLumpedInertia / EnergyTetStrain / EnergyPrescribedPositions   S/models/deformables/**
//   stark::RigidBodyDynamics / EnergyRigidBodyInertia / EnergyRigidBodyConstraints   S/models/rigidbodies/**
//   stark::EnergyFrictionalContact   S/models/interactions/EnergyFrictionalContact.{h,cpp}
//   stark::Simulation + presets      S/models/Simulation.h, S/models/presets/**
// Error convention of the reference is kept: fatal misuse prints a message and exit(-1)s.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/stark_b200.h"

namespace stark_b200 {

using Vec3 = std::array<double, 3>;
using Mat3 = std::array<double, 9>;   // row-major
using Quat = std::array<double, 4>;   // w, x, y, z

struct Settings {
    struct Simulation {
        Vec3 gravity = {0.0, 0.0, -9.81};
        bool init_frictional_contact = true;
        double max_time_step_size = 1.0 / 30.0;
        bool use_adaptive_time_step = true;
        double time_step_size_success_multiplier = 1.05;
        double time_step_size_lower_bound = 1e-6;
    } simulation;
    sb_newton_settings newton;
    int device = 0;
    void* stream = nullptr;
    Settings() { sb_newton_default_settings(&newton); }
};

// A host array mirrored on the device through the C-ABI (symx::DataMap analogue)
struct DeviceArray {
    std::vector<double> data;
    std::vector<double> shadow;   // what the device holds (uploads of unchanged arrays are skipped)
    bool has_shadow = false;
    bool volatile_data = false;   // rewritten every step: no shadow bookkeeping, always uploaded
    bool pinned = false;          // registered as a pinned mirror (asynchronous uploads)
    bool device_current = false;  // the device copy is the truth (rolled there at the end of a step); the host mirror follows asynchronously
    int stride = 1;
    int id = -1;
    std::string label;
    int rows() const { return (int)(data.size() / stride); }
};

class Stark;

// ---- deformables ------------------------------------------------------------------------------------------------
class PointDynamics {
public:
    DeviceArray X, x0, v0, v1, a, f;   // stride 3
    std::vector<int> set_begin;         // offsets of the point sets (IntervalVector)
    int add(const std::vector<Vec3>& x);
    int size() const { return X.rows(); }
    int get_global_index(int set, int local) const { return set_begin[set] + local; }
    int get_set_size(int set) const { return ((set + 1 < (int)set_begin.size()) ? set_begin[set + 1] : size()) - set_begin[set]; }
    void add_displacement(int set, const Vec3& d);
    void set_velocity(int set, const Vec3& v);
};

struct VolumeParams {   // stark::Volume::Params (S/models/presets/deformables_preset_types.h)
    double density = 1000.0, inertia_damping = 0.1;
    bool elasticity_only = false;
    double scale = 1.0, youngs_modulus = 1e4, poissons_ratio = 0.3, strain_limit = 1.0, strain_limit_stiffness = 1e2, strain_damping = 0.0;
    double contact_thickness = 0.0;
    static VolumeParams Soft_Rubber() { return VolumeParams(); }
};

struct SurfaceParams {   // stark::Surface::Params (S/models/presets/deformables_preset_types.{h,cpp})
    double density = 0.2, inertia_damping = 0.1;   // density in kg/m^2
    bool elasticity_only = false;
    double scale = 1.0, thickness = 0.001, youngs_modulus = 5e3, poissons_ratio = 0.3, strain_damping = 0.1 * 0.001 * 5e3, strain_limit = 0.1, strain_limit_stiffness = 1e6, inflation = 0.0;
    double bending_scale = 1.0, bending_stiffness = 1e-6, bending_damping = 0.1 * 1e-6;
    bool flat_rest_angle = true;
    double contact_thickness = 0.0;
    static SurfaceParams Cotton_Fabric() { return SurfaceParams(); }
};

class EnergyLumpedInertia {
public:
    std::vector<std::array<int32_t, 3>> conn;   // {idx, glob, group}
    DeviceArray lumped_volume, density, damping, is_quasistatic;
    int potential = -1;
    int add(PointDynamics& dyn, int set, const std::vector<std::array<int, 4>>& tets, double density, double damping);
    int add(PointDynamics& dyn, int set, const std::vector<std::array<int, 3>>& triangles, double density, double damping);
private:
    int add_lumped(PointDynamics& dyn, int set, const std::vector<double>& lumped, double density, double damping);
};

class EnergyTriangleStrain {
public:
    std::vector<std::array<int32_t, 5>> conn_complete, conn_elasticity_only;   // {idx, group, i, j, k}
    DeviceArray scale, thickness, youngs_modulus, poissons_ratio, strain_damping, strain_limit, strain_limit_stiffness, inflation;
    int potential_complete = -1, potential_elasticity_only = -1;
    int add(PointDynamics& dyn, int set, const std::vector<std::array<int, 3>>& triangles, const SurfaceParams& p);
};

class EnergyDiscreteShells {
public:
    std::vector<std::array<int32_t, 6>> conn_complete, conn_flat_rest;   // {idx, group, v_edge_0, v_edge_1, v_opp_0, v_opp_1}
    DeviceArray rest_dihedral_angle_rad, rest_edge_length, rest_height, bergou_K, bergou_coef;   // per hinge (bergou_K stride 4)
    DeviceArray scale, bending_stiffness, bending_damping;                                       // per group
    int potential_complete = -1, potential_flat_rest = -1;
    int add(PointDynamics& dyn, int set, const std::vector<std::array<int, 3>>& triangles, const SurfaceParams& p);
};

class EnergyTetStrain {
public:
    std::vector<std::array<int32_t, 6>> conn_complete, conn_elasticity_only;   // {idx, group, i, j, k, l}
    DeviceArray scale, youngs_modulus, poissons_ratio, strain_limit, strain_limit_stiffness, strain_damping;
    int potential_complete = -1, potential_elasticity_only = -1;
    int add(PointDynamics& dyn, int set, const std::vector<std::array<int, 4>>& tets, const VolumeParams& p);
};

class EnergyPrescribedPositions {
public:
    std::vector<std::array<int32_t, 3>> conn;   // {idx, point, group}
    DeviceArray target_positions, stiffness;
    std::vector<Vec3> rest_positions;
    std::vector<double> tolerance;
    std::vector<std::array<int, 2>> group_begin_end;
    int potential = -1;
    int add_inside_aabb(PointDynamics& dyn, int set, const Vec3& center, const Vec3& dim, double stiffness, double tolerance);
    // does is_converged_state_valid have anything to test? (the default tolerance is infinite, S/models/types.h:56-57)
    bool checks_tolerance() const { if (conn.empty()) return false; for (double t : tolerance) if (t < 1e300) return true; return false; }
    void set_transformation(int group, const Vec3& t, double angle_deg, const Vec3& axis);
    bool is_converged_state_valid(const PointDynamics& dyn, double dt);
};

// ---- rigid bodies ---------------------------------------------------------------------------------------------------
class RigidBodyDynamics {
public:
    DeviceArray t0, q0_, v0, v1, w0, w1, a, aa, force, torque;   // q0_ stride 4 (w,x,y,z), others stride 3
    std::vector<Quat> q0;
    std::vector<Mat3> R0;
    int add();
    int get_n_bodies() const { return (int)q0.size(); }
    Vec3 get_x1(int rb, const Vec3& x_loc, double dt) const;
    Vec3 get_d1(int rb, const Vec3& d_loc, double dt) const;
    Vec3 position_at(int rb, const Vec3& x_loc) const;
    Vec3 direction(int rb, const Vec3& d_loc) const;
};

class EnergyRigidBodyInertia {
public:
    std::vector<std::array<int32_t, 1>> conn;
    DeviceArray mass, linear_damping, angular_damping, is_quasistatic, J0_glob;   // J0_glob stride 9
    std::vector<Mat3> J_loc;
    int potential_linear = -1, potential_angular = -1;
    void add(int rb, double mass, const Mat3& inertia_local);
    void before_time_step(const RigidBodyDynamics& rb);
};

class EnergyRigidBodyConstraints {
public:
    struct GlobalPoints { std::vector<std::array<int32_t, 2>> conn; DeviceArray loc, target_glob, stiffness, is_active; std::vector<double> tolerance_in_m; int potential = -1; } global_points;
    struct GlobalDirections { std::vector<std::array<int32_t, 2>> conn; DeviceArray d_loc, target_d_glob, stiffness, is_active; std::vector<double> tolerance_in_deg; std::vector<Vec3> d_loc_rest; int potential = -1; } global_directions;
    struct Fix { int anchor_point, z_lock, x_lock; };   // RBCFixHandler (rigidbody_constraints_ui.h:333-380)
    struct Points { std::vector<std::array<int32_t, 3>> conn; DeviceArray a_loc, b_loc, stiffness, is_active; std::vector<double> tolerance_in_m; int potential = -1; } points;
    struct Directions { std::vector<std::array<int32_t, 3>> conn; DeviceArray da_loc, db_loc, stiffness, is_active; std::vector<double> tolerance_in_deg; int potential = -1; } directions;
    double default_stiffness = 1e6, default_tolerance_in_m = 0.001, default_tolerance_in_deg = 1.0;
    double stiffness_hard_multiplier = 2.0, stiffness_soft_multiplier = 1.05, soft_constraint_capacity_hardening_point = 0.75;
    Fix add_fix(const RigidBodyDynamics& rb, int body);
    void set_fix_transformation(const Fix& fix, const Vec3& translation, double angle_deg, const Vec3& axis);
    void add_hinge(const RigidBodyDynamics& rb, int body_a, int body_b, const Vec3& p_glob, const Vec3& d_glob);
    bool adjust_stiffness(const RigidBodyDynamics& rb, double dt, double cap, double multiplier, bool positions_set);
};

// ---- interactions ------------------------------------------------------------------------------------------------------
class EnergyFrictionalContact {
public:
    struct GlobalParams {
        double default_contact_thickness = -1.0, min_contact_stiffness = 1e6, max_contact_stiffness = 1e20, friction_stick_slide_threshold = 0.1;
        bool collisions_enabled = true, friction_enabled = true, triangle_point_enabled = true, edge_edge_enabled = true, intersection_test_enabled = true;
    } global_params;
    struct Mesh { int ps, idx_in_ps; std::vector<int32_t> vertex_global; std::vector<double> vertices_local; std::vector<int32_t> triangles, edges; double thickness; };
    std::vector<Mesh> meshes;
    std::vector<std::array<int, 2>> disabled, friction_pairs_idx;
    std::vector<double> friction_pairs_mu;
    double contact_stiffness = 1e3;
    void set_global_params(const GlobalParams& p) { global_params = p; contact_stiffness = p.min_contact_stiffness; }
    int add_triangles_deformable(int set, const std::vector<int32_t>& vertex_global, const std::vector<std::array<int, 3>>& triangles, double thickness);
    int add_triangles_rigid(int body, const std::vector<Vec3>& vertices, const std::vector<std::array<int, 3>>& triangles, double thickness);
    void set_friction(int a, int b, double mu);
    void disable_collision(int a, int b);
    bool is_empty() const { return meshes.empty(); }
};

// ---- core --------------------------------------------------------------------------------------------------------------
struct StepStats {
    int result = 0;   // symx::SolverReturn
    int newton_iterations = 0, cg_iterations = 0, ls_inv = 0, ls_bt = 0, n_evaluations = 0;
    double dt = 0.0, runtime_s = 0.0, solve_s = 0.0, solve_gpu_ms = 0.0, first_residual = 0.0;
    std::vector<double> residuals;
    bool accepted = false;
};

class Simulation {
public:
    explicit Simulation(const Settings& settings);
    ~Simulation();
    Simulation(const Simulation&) = delete;

    Settings settings;
    double dt, current_time = 0.0;
    int current_time_step = 0;
    Vec3 gravity;

    PointDynamics dyn;
    EnergyLumpedInertia lumped_inertia;
    EnergyTetStrain tet_strain;
    EnergyTriangleStrain triangle_strain;
    EnergyDiscreteShells discrete_shells;
    EnergyPrescribedPositions prescribed_positions;
    RigidBodyDynamics rb;
    EnergyRigidBodyInertia rb_inertia;
    EnergyRigidBodyConstraints rb_constraints;
    EnergyFrictionalContact contact;

    // presets (S/models/presets/DeformablesPresets.cpp:73-85, RigidBodyPresets.cpp:47-53)
    struct VolumeHandle { int point_set, contact_group; int n_vertices, n_tets; };
    struct BoxHandle { int body, contact_group; };
    struct SurfaceHandle { int point_set, contact_group; int n_vertices, n_triangles; };
    VolumeHandle add_volume_grid(const Vec3& dim, const std::array<int, 3>& subdivisions, const VolumeParams& params);
    SurfaceHandle add_surface_grid(const std::array<double, 2>& dim, const std::array<int, 2>& subdivisions, const SurfaceParams& params);
    BoxHandle add_box(double mass, const Vec3& size, double contact_thickness = 0.0);
    void set_translation(int body, const Vec3& t);
    void add_time_event(std::function<void(double)> f) { time_events.push_back(f); }

    // stark::Simulation::run_one_time_step (S/models/Simulation.cpp:73-77) -> Stark::run_one_step (S/core/Stark.cpp:133-244)
    bool run_one_time_step();
    const StepStats& last_step() const { return stats; }
    sb_context* context() { return ctx; }
    const DeviceArray* find_array(const std::string& label) const { for (const DeviceArray* a : all_arrays) if (a->label == label) return a; return nullptr; }
    int ndofs();
    // totals over all steps taken so far
    long long total_newton_iterations = 0, total_evaluations = 0, total_cg_iterations = 0, h2d_bytes = 0, d2h_bytes = 0;
    double total_solve_s = 0.0;
    bool host_mirror_pending = false;   // asynchronous read-backs of x0 / v0 in flight (sync_host waits for them)
    void sync_host();
    double phase_s[4] = {0, 0, 0, 0};   // host wall time: before the solve, sb_newton_solve, converged-state callbacks + downloads, state roll (SB_HOST_DUMP=1 prints them)
    long long phase_steps = 0;

private:
    sb_context* ctx = nullptr;
    bool is_init = false;
    DeviceArray dt_arr, gravity_arr;
    StepStats stats;
    std::vector<std::function<void(double)>> time_events;
    std::vector<DeviceArray*> all_arrays;
    void initialize();
    void reg(DeviceArray& a, const char* label, int stride);
    void upload(DeviceArray& a);
    void pin(DeviceArray& a);
    void download(DeviceArray& a);
    void check(int status, const char* what);
};

// mesh helpers (S/utils/mesh_generators.cpp:264-377, S/utils/mesh_utils.cpp:278-327, S/utils/mesh_utils.h:153-166)
void generate_triangle_grid(std::vector<Vec3>& vertices, std::vector<std::array<int, 3>>& triangles, const std::array<double, 2>& center, const std::array<double, 2>& dim, const std::array<int, 2>& n, double z = 0.0);
void find_internal_angles(std::vector<std::array<int, 4>>& internal_angles, const std::vector<std::array<int, 3>>& triangles, int n_nodes);
void generate_tet_grid(std::vector<Vec3>& vertices, std::vector<std::array<int, 4>>& tets, const Vec3& center, const Vec3& dim, const std::array<int, 3>& n);
void find_surface(std::vector<std::array<int, 3>>& triangles, std::vector<int>& triangle_to_tet_node_map, const std::vector<Vec3>& vertices, const std::vector<std::array<int, 4>>& tets);
std::vector<std::array<int, 2>> find_edges_from_triangles(const std::vector<std::array<int, 3>>& triangles, int n_nodes);

}  // namespace stark_b200
