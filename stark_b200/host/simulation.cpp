// Host layer above the C-ABI: scene state, binding of arrays / potentials, and the time-step loop.
// See stark_b200.hpp for the map to the reference's classes.
#include "stark_b200.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <iterator>
#include <map>

namespace stark_b200 {

// ---------------------------------------------------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------------------------------------------------
static inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
static inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
static inline Vec3 operator*(double s, const Vec3& a) { return {s * a[0], s * a[1], s * a[2]}; }
static inline double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline Vec3 cross(const Vec3& a, const Vec3& b) { return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; }
static inline double norm(const Vec3& a) { return std::sqrt(dot(a, a)); }
static inline Vec3 matvec(const Mat3& R, const Vec3& x) { return {R[0] * x[0] + R[1] * x[1] + R[2] * x[2], R[3] * x[0] + R[4] * x[1] + R[5] * x[2], R[6] * x[0] + R[7] * x[1] + R[8] * x[2]}; }
static inline Vec3 matTvec(const Mat3& R, const Vec3& x) { return {R[0] * x[0] + R[3] * x[1] + R[6] * x[2], R[1] * x[0] + R[4] * x[1] + R[7] * x[2], R[2] * x[0] + R[5] * x[1] + R[8] * x[2]}; }
static Mat3 quat_to_rotation(const Quat& q)
{   // Eigen::Quaterniond::toRotationMatrix
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    return {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
}
static Quat quat_time_integration(const Quat& q, const Vec3& w, double dt)
{   // S/models/rigidbodies/rigidbody_transformations.cpp:33-40
    const double pw = -w[0] * q[1] - w[1] * q[2] - w[2] * q[3];
    const double px = w[0] * q[0] + w[1] * q[3] - w[2] * q[2];
    const double py = w[1] * q[0] + w[2] * q[1] - w[0] * q[3];
    const double pz = w[2] * q[0] + w[0] * q[2] - w[1] * q[1];
    Quat e = {q[0] + 0.5 * dt * pw, q[1] + 0.5 * dt * px, q[2] + 0.5 * dt * py, q[3] + 0.5 * dt * pz};
    const double n = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]);
    for (double& c : e) c /= n;
    return e;
}
static Vec3 get3(const DeviceArray& a, int i) { return {a.data[3 * i], a.data[3 * i + 1], a.data[3 * i + 2]}; }
static void set3(DeviceArray& a, int i, const Vec3& v) { for (int c = 0; c < 3; c++) a.data[3 * i + c] = v[c]; }
static void push3(DeviceArray& a, const Vec3& v) { a.stride = 3; for (int c = 0; c < 3; c++) a.data.push_back(v[c]); }
static void push1(DeviceArray& a, double v) { a.stride = 1; a.data.push_back(v); }
[[noreturn]] static void die(const std::string& msg) { std::cout << "stark_b200 error: " << msg << std::endl; exit(-1); }

// ---------------------------------------------------------------------------------------------------------------------
// mesh helpers
// ---------------------------------------------------------------------------------------------------------------------
void generate_tet_grid(std::vector<Vec3>& V, std::vector<std::array<int, 4>>& T, const Vec3& center, const Vec3& dim, const std::array<int, 3>& n)
{   // 12 tets per hexahedron around a centre node, alternating diagonals (S/utils/mesh_generators.cpp:264-377)
    const Vec3 bottom = center - 0.5 * dim;
    const int nx = n[0] + 1, ny = n[1] + 1, nz = n[2] + 1;
    const int n_points = nx * ny * nz, n_hex = n[0] * n[1] * n[2];
    const double dx = dim[0] / n[0], dy = dim[1] / n[1], dz = dim[2] / n[2];
    V.assign(n_points + n_hex, Vec3{0, 0, 0});
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++)
        V[nz * ny * i + nz * j + k] = {bottom[0] + i * dx, bottom[1] + j * dy, bottom[2] + k * dz};
    const int co = n_points;
    for (int i = 0; i < n[0]; i++) for (int j = 0; j < n[1]; j++) for (int k = 0; k < n[2]; k++)
        V[co + n[2] * n[1] * i + n[2] * j + k] = {bottom[0] + i * dx + 0.5 * dx, bottom[1] + j * dy + 0.5 * dy, bottom[2] + k * dz + 0.5 * dz};
    T.clear();
    T.reserve(12 * (size_t)n_hex);
    for (int ei = 0; ei < n[0]; ei++) for (int ej = 0; ej < n[1]; ej++) for (int ek = 0; ek < n[2]; ek++) {
        int c[9];
        for (int a = 0; a < 8; a++) c[a] = nz * ny * (ei + ((a >> 2) & 1)) + nz * (ej + ((a >> 1) & 1)) + (ek + (a & 1));
        c[8] = co + n[2] * n[1] * ei + n[2] * ej + ek;
        static const int A[12][3] = {{0, 1, 4}, {1, 5, 4}, {0, 2, 1}, {1, 2, 3}, {0, 4, 6}, {0, 6, 2}, {3, 2, 7}, {2, 6, 7}, {4, 5, 7}, {4, 7, 6}, {1, 3, 5}, {3, 7, 5}};
        static const int B[12][3] = {{0, 1, 5}, {0, 5, 4}, {0, 3, 1}, {0, 2, 3}, {0, 4, 2}, {2, 4, 6}, {3, 2, 6}, {3, 6, 7}, {5, 7, 6}, {5, 6, 4}, {1, 3, 7}, {1, 7, 5}};
        const bool first = ((ek % 2 == 0) && (ei % 2 == ej % 2)) || ((ek % 2 == 1) && (ei % 2 != ej % 2));
        const int (*P)[3] = first ? A : B;
        for (int t = 0; t < 12; t++) T.push_back({c[P[t][0]], c[P[t][1]], c[P[t][2]], c[8]});
    }
}

void find_surface(std::vector<std::array<int, 3>>& out_tris, std::vector<int>& tri_to_tet_node, const std::vector<Vec3>& V, const std::vector<std::array<int, 4>>& tets)
{   // faces that occur once, wound outwards, renumbered to a compact vertex set (S/utils/mesh_utils.cpp:278-327)
    std::map<std::array<int, 3>, int> face_tet;
    for (int t = 0; t < (int)tets.size(); t++) {
        const auto& q = tets[t];
        const std::array<std::array<int, 3>, 4> faces = {{{q[0], q[1], q[2]}, {q[0], q[1], q[3]}, {q[0], q[2], q[3]}, {q[1], q[2], q[3]}}};
        for (auto f : faces) {
            std::sort(f.begin(), f.end());
            auto it = face_tet.find(f);
            if (it == face_tet.end()) face_tet[f] = t; else face_tet.erase(it);
        }
    }
    std::vector<std::array<int, 3>> tris;
    for (const auto& kv : face_tet) {
        std::array<int, 3> f = kv.first;
        const auto& q = tets[kv.second];
        const Vec3 c = 0.25 * (V[q[0]] + V[q[1]] + V[q[2]] + V[q[3]]);
        const Vec3 nrm = cross(V[f[1]] - V[f[0]], V[f[2]] - V[f[0]]);
        if (dot(nrm, c - V[f[0]]) > 0.0) std::swap(f[0], f[1]);   // normal must point away from the tet centre
        tris.push_back(f);
    }
    std::vector<int> used(V.size(), 0);
    for (auto& f : tris) for (int v : f) used[v] = 1;
    std::vector<int> old_to_new(V.size(), -1);
    tri_to_tet_node.clear();
    for (int v = 0; v < (int)V.size(); v++) if (used[v]) { old_to_new[v] = (int)tri_to_tet_node.size(); tri_to_tet_node.push_back(v); }
    out_tris.clear();
    for (auto& f : tris) out_tris.push_back({old_to_new[f[0]], old_to_new[f[1]], old_to_new[f[2]]});
}

std::vector<std::array<int, 2>> find_edges_from_triangles(const std::vector<std::array<int, 3>>& tris, int n_nodes)
{   // S/utils/mesh_utils.h:153-166
    std::vector<std::array<int, 2>> e;
    for (const auto& t : tris)
        for (int i = 0; i < 3; i++) for (int j = i + 1; j < 3; j++) e.push_back({std::min(t[i], t[j]), std::max(t[i], t[j])});
    std::sort(e.begin(), e.end(), [&](const std::array<int, 2>& a, const std::array<int, 2>& b) { return (long long)a[0] * n_nodes + a[1] < (long long)b[0] * n_nodes + b[1]; });
    e.erase(std::unique(e.begin(), e.end()), e.end());
    return e;
}

void generate_triangle_grid(std::vector<Vec3>& V, std::vector<std::array<int, 3>>& T, const std::array<double, 2>& center, const std::array<double, 2>& dim, const std::array<int, 2>& n, double z)
{   // two triangles per quad, diagonals alternating in a checkerboard (S/utils/mesh_generators.cpp:100-167)
    const double bx = center[0] - 0.5 * dim[0], by = center[1] - 0.5 * dim[1];
    const int nx = n[0] + 1, ny = n[1] + 1;
    const double dx = dim[0] / n[0], dy = dim[1] / n[1];
    V.assign((size_t)nx * ny, Vec3{0, 0, 0});
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) V[(size_t)ny * i + j] = {bx + i * dx, by + j * dy, z};
    T.clear();
    for (int ei = 0; ei < n[0]; ei++)
        for (int ej = 0; ej < n[1]; ej++) {
            const int q[4] = {ny * ei + ej, ny * ei + ej + 1, ny * (ei + 1) + ej, ny * (ei + 1) + ej + 1};
            if (ei % 2 == ej % 2) { T.push_back({q[0], q[2], q[3]}); T.push_back({q[0], q[3], q[1]}); }
            else { T.push_back({q[0], q[2], q[1]}); T.push_back({q[2], q[3], q[1]}); }
        }
}
void find_internal_angles(std::vector<std::array<int, 4>>& out, const std::vector<std::array<int, 3>>& tris, int n_nodes)
{   // for every edge, the two nodes adjacent to both end points (S/utils/mesh_utils.cpp:217-252)
    out.clear();
    if (tris.empty()) return;
    std::vector<std::vector<int>> nn(n_nodes);
    auto add = [&](int a, int b) { if (std::find(nn[a].begin(), nn[a].end(), b) == nn[a].end()) nn[a].push_back(b); };
    for (const auto& t : tris)
        for (int i = 0; i < 3; i++) for (int j = i + 1; j < 3; j++) { add(t[i], t[j]); add(t[j], t[i]); }
    for (auto& v : nn) std::sort(v.begin(), v.end());
    std::vector<int> buf;
    for (const auto& e : find_edges_from_triangles(tris, n_nodes)) {
        buf.clear();
        std::set_intersection(nn[e[0]].begin(), nn[e[0]].end(), nn[e[1]].begin(), nn[e[1]].end(), std::back_inserter(buf));
        if (buf.size() == 2) out.push_back({e[0], e[1], buf[0], buf[1]});
        else if (buf.size() > 2) die("triangle mesh has edges with more than two incident triangles.");
    }
}
static Mat3 angle_axis_rotation(double angle_deg, const Vec3& axis)
{   // Eigen::AngleAxisd(deg2rad(angle), axis).toRotationMatrix() (Rodrigues)
    const double th = angle_deg * M_PI / 180.0, c = std::cos(th), sn = std::sin(th);
    const Vec3 u = (1.0 / norm(axis)) * axis;
    return {c + u[0] * u[0] * (1 - c), u[0] * u[1] * (1 - c) - u[2] * sn, u[0] * u[2] * (1 - c) + u[1] * sn,
            u[1] * u[0] * (1 - c) + u[2] * sn, c + u[1] * u[1] * (1 - c), u[1] * u[2] * (1 - c) - u[0] * sn,
            u[2] * u[0] * (1 - c) - u[1] * sn, u[2] * u[1] * (1 - c) + u[0] * sn, c + u[2] * u[2] * (1 - c)};
}

// ---------------------------------------------------------------------------------------------------------------------
// deformables
// ---------------------------------------------------------------------------------------------------------------------
int PointDynamics::add(const std::vector<Vec3>& x)
{
    const int set = (int)set_begin.size();
    set_begin.push_back(size());
    for (const Vec3& p : x) { push3(X, p); push3(x0, p); push3(v0, {0, 0, 0}); push3(v1, {0, 0, 0}); push3(a, {0, 0, 0}); push3(f, {0, 0, 0}); }
    return set;
}
void PointDynamics::add_displacement(int set, const Vec3& d)
{   // PointSetHandler::add_displacement(..., also_at_rest_pose = true)
    for (int i = 0; i < get_set_size(set); i++) { const int g = get_global_index(set, i); set3(x0, g, get3(x0, g) + d); set3(X, g, get3(X, g) + d); }
    x0.device_current = false;
}
void PointDynamics::set_velocity(int set, const Vec3& v)
{
    for (int i = 0; i < get_set_size(set); i++) set3(v0, get_global_index(set, i), v);
    v0.device_current = false;
}

int EnergyLumpedInertia::add(PointDynamics& dyn, int set, const std::vector<std::array<int, 4>>& tets, double rho, double damp)
{   // S/models/deformables/point/EnergyLumpedInertia.cpp:137-159 (quarter of every incident tet volume), :52-71
    const int group = density.rows();
    push1(density, rho); push1(damping, damp); push1(is_quasistatic, 0.0);
    std::vector<double> lumped(dyn.get_set_size(set), 0.0);
    for (const auto& t : tets) {
        const Vec3 A = get3(dyn.X, dyn.get_global_index(set, t[0])), B = get3(dyn.X, dyn.get_global_index(set, t[1])), C = get3(dyn.X, dyn.get_global_index(set, t[2])), D = get3(dyn.X, dyn.get_global_index(set, t[3]));
        const double vol = std::abs(dot(cross(B - A, C - A), D - A)) / 6.0;
        for (int k = 0; k < 4; k++) lumped[t[k]] += vol / 4.0;
    }
    for (int i = 0; i < (int)lumped.size(); i++) {
        if (lumped[i] > 0.0) {
            conn.push_back({(int32_t)conn.size(), dyn.get_global_index(set, i), group});
            push1(lumped_volume, lumped[i]);
        }
    }
    return group;
}
int EnergyLumpedInertia::add(PointDynamics& dyn, int set, const std::vector<std::array<int, 3>>& tris, double rho, double damp)
{   // S/models/deformables/point/EnergyLumpedInertia.cpp:116-138 (a third of every incident triangle area; density per m^2)
    const int group = density.rows();
    push1(density, rho); push1(damping, damp); push1(is_quasistatic, 0.0);
    std::vector<double> lumped(dyn.get_set_size(set), 0.0);
    for (const auto& t : tris) {
        const Vec3 A = get3(dyn.X, dyn.get_global_index(set, t[0])), B = get3(dyn.X, dyn.get_global_index(set, t[1])), C = get3(dyn.X, dyn.get_global_index(set, t[2]));
        const double area = 0.5 * norm(cross(A - C, B - C));
        for (int k = 0; k < 3; k++) lumped[t[k]] += area / 3.0;
    }
    for (int i = 0; i < (int)lumped.size(); i++) {
        if (lumped[i] > 0.0) {
            conn.push_back({(int32_t)conn.size(), dyn.get_global_index(set, i), group});
            push1(lumped_volume, lumped[i]);
        }
    }
    return group;
}

int EnergyTriangleStrain::add(PointDynamics& dyn, int set, const std::vector<std::array<int, 3>>& tris, const SurfaceParams& p)
{   // S/models/deformables/surface/EnergyTriangleStrain.cpp:131-155
    const int group = youngs_modulus.rows();
    push1(scale, p.scale); push1(thickness, p.thickness); push1(youngs_modulus, p.youngs_modulus); push1(poissons_ratio, p.poissons_ratio);
    push1(strain_damping, p.strain_damping); push1(strain_limit, p.strain_limit); push1(strain_limit_stiffness, p.strain_limit_stiffness); push1(inflation, p.inflation);
    auto& conn = p.elasticity_only ? conn_elasticity_only : conn_complete;
    for (const auto& t : tris)
        conn.push_back({(int32_t)conn.size(), group, dyn.get_global_index(set, t[0]), dyn.get_global_index(set, t[1]), dyn.get_global_index(set, t[2])});
    return group;
}

int EnergyDiscreteShells::add(PointDynamics& dyn, int set, const std::vector<std::array<int, 3>>& tris, const SurfaceParams& p)
{   // S/models/deformables/surface/EnergyDiscreteShells.cpp:93-166
    constexpr double EPSILON = 1e-12;
    const int group = bending_stiffness.rows();
    push1(scale, p.bending_scale); push1(bending_stiffness, p.bending_stiffness); push1(bending_damping, p.bending_damping);
    if (p.flat_rest_angle && p.bending_scale != 1.0) die("EnergyDiscreteShells::add(): scale cannot be different from 1.0 if flat_rest_angle == true");
    std::vector<std::array<int, 4>> angles;
    find_internal_angles(angles, tris, dyn.get_set_size(set));
    auto cot = [](const Vec3& v, const Vec3& w) { return dot(v, w) / norm(cross(v, w)); };
    auto& conn = p.flat_rest_angle ? conn_flat_rest : conn_complete;
    bergou_K.stride = 4;
    for (const auto& a : angles) {
        const int g[4] = {dyn.get_global_index(set, a[0]), dyn.get_global_index(set, a[1]), dyn.get_global_index(set, a[2]), dyn.get_global_index(set, a[3])};
        conn.push_back({(int32_t)conn.size(), group, g[0], g[1], g[2], g[3]});
        const Vec3 X0 = get3(dyn.X, g[0]), X1 = get3(dyn.X, g[1]), X2 = get3(dyn.X, g[2]), X3 = get3(dyn.X, g[3]);
        const Vec3 e0 = X1 - X0, e1 = X2 - X0, e2 = X3 - X0, e3 = X2 - X1, e4 = X3 - X1;
        const double len = norm(e0);
        push1(rest_edge_length, len);
        const Vec3 n0 = cross(e0, e1), n1 = -1.0 * cross(e0, e2);
        push1(rest_dihedral_angle_rad, std::acos((1.0 - EPSILON) * dot((1.0 / norm(n0)) * n0, (1.0 / norm(n1)) * n1)));
        const double A0 = 0.5 * norm(n0), A1 = 0.5 * norm(n1);
        push1(rest_height, 1.0 / 6.0 * (2.0 * A0 / len + 2.0 * A1 / len));
        const Vec3 me0 = -1.0 * e0;
        const double c01 = cot(e0, e1), c02 = cot(e0, e2), c03 = cot(me0, e3), c04 = cot(me0, e4);
        push1(bergou_coef, 3.0 / (A0 + A1) * 0.5);
        for (double k : {c03 + c04, c01 + c02, -c01 - c03, -c02 - c04}) bergou_K.data.push_back(k);
    }
    return group;
}

int EnergyTetStrain::add(PointDynamics& dyn, int set, const std::vector<std::array<int, 4>>& tets, const VolumeParams& p)
{   // S/models/deformables/volume/EnergyTetStrain.cpp:125-146
    const int group = youngs_modulus.rows();
    push1(scale, p.scale); push1(youngs_modulus, p.youngs_modulus); push1(poissons_ratio, p.poissons_ratio);
    push1(strain_damping, p.strain_damping); push1(strain_limit, p.strain_limit); push1(strain_limit_stiffness, p.strain_limit_stiffness);
    auto& conn = p.elasticity_only ? conn_elasticity_only : conn_complete;
    for (const auto& t : tets)
        conn.push_back({(int32_t)conn.size(), group, dyn.get_global_index(set, t[0]), dyn.get_global_index(set, t[1]), dyn.get_global_index(set, t[2]), dyn.get_global_index(set, t[3])});
    return group;
}

int EnergyPrescribedPositions::add_inside_aabb(PointDynamics& dyn, int set, const Vec3& c, const Vec3& dim, double k, double tol)
{   // S/models/deformables/point/EnergyPrescribedPositions.cpp:36-68
    const int group = stiffness.rows();
    push1(stiffness, k);
    tolerance.push_back(tol);
    const int begin = target_positions.rows();
    for (int i = 0; i < dyn.get_set_size(set); i++) {
        const int g = dyn.get_global_index(set, i);
        const Vec3 x = get3(dyn.x0, g);
        bool inside = true;
        for (int d = 0; d < 3; d++) inside = inside && (c[d] - 0.5 * dim[d] <= x[d]) && (x[d] <= c[d] + 0.5 * dim[d]);
        if (!inside) continue;
        conn.push_back({(int32_t)conn.size(), g, group});
        push3(target_positions, x);
        rest_positions.push_back(x);
    }
    group_begin_end.push_back({begin, target_positions.rows()});
    return group;
}
void EnergyPrescribedPositions::set_transformation(int group, const Vec3& t, double angle_deg, const Vec3& axis)
{   // :125-139, Eigen::AngleAxisd -> rotation matrix (Rodrigues)
    const double th = angle_deg * M_PI / 180.0, c = std::cos(th), s = std::sin(th);
    const Vec3 u = (1.0 / norm(axis)) * axis;
    const Mat3 R = {c + u[0] * u[0] * (1 - c), u[0] * u[1] * (1 - c) - u[2] * s, u[0] * u[2] * (1 - c) + u[1] * s,
                    u[1] * u[0] * (1 - c) + u[2] * s, c + u[1] * u[1] * (1 - c), u[1] * u[2] * (1 - c) - u[0] * s,
                    u[2] * u[0] * (1 - c) - u[1] * s, u[2] * u[1] * (1 - c) + u[0] * s, c + u[2] * u[2] * (1 - c)};
    for (int i = group_begin_end[group][0]; i < group_begin_end[group][1]; i++) set3(target_positions, i, matvec(R, rest_positions[i]) + t);
}
bool EnergyPrescribedPositions::is_converged_state_valid(const PointDynamics& dyn, double dt)
{   // :141-165
    for (const auto& row : conn) {
        const Vec3 x1 = get3(dyn.x0, row[1]) + dt * get3(dyn.v1, row[1]);
        const Vec3 d = x1 - get3(target_positions, row[0]);
        const double tol = tolerance[row[2]];
        if (dot(d, d) > tol * tol) { stiffness.data[row[2]] *= 2.0; return false; }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// rigid bodies
// ---------------------------------------------------------------------------------------------------------------------
int RigidBodyDynamics::add()
{
    const int id = get_n_bodies();
    q0.push_back({1, 0, 0, 0});
    R0.push_back({1, 0, 0, 0, 1, 0, 0, 0, 1});
    q0_.stride = 4;
    for (double c : {1.0, 0.0, 0.0, 0.0}) q0_.data.push_back(c);
    for (DeviceArray* a : {&t0, &v0, &v1, &w0, &w1, &this->a, &aa, &force, &torque}) push3(*a, {0, 0, 0});
    return id;
}
Vec3 RigidBodyDynamics::get_x1(int b, const Vec3& x_loc, double dt) const
{
    const Mat3 R1 = quat_to_rotation(quat_time_integration(q0[b], get3(w1, b), dt));
    return (get3(t0, b) + dt * get3(v1, b)) + matvec(R1, x_loc);
}
Vec3 RigidBodyDynamics::get_d1(int b, const Vec3& d_loc, double dt) const { return matvec(quat_to_rotation(quat_time_integration(q0[b], get3(w1, b), dt)), d_loc); }
Vec3 RigidBodyDynamics::position_at(int b, const Vec3& x_loc) const { return get3(t0, b) + matvec(R0[b], x_loc); }
Vec3 RigidBodyDynamics::direction(int b, const Vec3& d_loc) const { return matvec(R0[b], d_loc); }

void EnergyRigidBodyInertia::add(int rb, double m, const Mat3& J)
{
    if (rb != mass.rows()) die("EnergyRigidBodyInertia::add() found non-consecutive rigid body added.");
    conn.push_back({(int32_t)rb});
    push1(mass, m); push1(linear_damping, 0.0); push1(angular_damping, 0.0); push1(is_quasistatic, 0.0);
    J_loc.push_back(J);
    J0_glob.stride = 9;
    J0_glob.data.resize(9 * (size_t)(rb + 1), 0.0);
}
void EnergyRigidBodyInertia::before_time_step(const RigidBodyDynamics& rb)
{   // J0_glob = R0 J R0^T (S/models/rigidbodies/EnergyRigidBodyInertia.cpp:85-104)
    for (int b = 0; b < rb.get_n_bodies(); b++) {
        const Mat3& R = rb.R0[b];
        const Mat3& J = J_loc[b];
        Mat3 RJ;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { RJ[3 * i + j] = 0; for (int k = 0; k < 3; k++) RJ[3 * i + j] += R[3 * i + k] * J[3 * k + j]; }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += RJ[3 * i + k] * R[3 * j + k]; J0_glob.data[9 * b + 3 * i + j] = s; }
    }
}

EnergyRigidBodyConstraints::Fix EnergyRigidBodyConstraints::add_fix(const RigidBodyDynamics& rb, int body)
{   // RigidBodies::add_constraint_fix (S/models/rigidbodies/RigidBodies.cpp:205-212): anchor point + z lock + x lock
    const Vec3 t = get3(rb.t0, body);
    Fix fix;
    fix.anchor_point = (int)global_points.conn.size();
    fix.z_lock = (int)global_directions.conn.size();
    fix.x_lock = fix.z_lock + 1;
    {
        auto& d = global_points;
        d.conn.push_back({(int32_t)d.conn.size(), body});
        push3(d.loc, matTvec(rb.R0[body], t - get3(rb.t0, body))); push3(d.target_glob, t);
        push1(d.stiffness, default_stiffness); push1(d.is_active, 1.0); d.tolerance_in_m.push_back(default_tolerance_in_m);
    }
    for (const Vec3& dir : {Vec3{0, 0, 1}, Vec3{1, 0, 0}}) {
        auto& d = global_directions;
        d.conn.push_back({(int32_t)d.conn.size(), body});
        push3(d.d_loc, matTvec(rb.R0[body], dir)); push3(d.target_d_glob, dir);
        d.d_loc_rest.push_back(matTvec(rb.R0[body], dir));
        push1(d.stiffness, default_stiffness); push1(d.is_active, 1.0); d.tolerance_in_deg.push_back(default_tolerance_in_deg);
    }
    return fix;
}
void EnergyRigidBodyConstraints::set_fix_transformation(const Fix& fix, const Vec3& translation, double angle_deg, const Vec3& axis)
{   // RBCFixHandler::set_transformation (rigidbody_constraints_ui.h:369-379): new anchor target; the locks' LOCAL directions become R d_loc_rest
    set3(global_points.target_glob, fix.anchor_point, translation);
    const Mat3 R = angle_axis_rotation(angle_deg, axis);
    for (int k : {fix.z_lock, fix.x_lock}) set3(global_directions.d_loc, k, matvec(R, global_directions.d_loc_rest[k]));
}
void EnergyRigidBodyConstraints::add_hinge(const RigidBodyDynamics& rb, int a, int b, const Vec3& p, const Vec3& dg)
{   // RigidBodies::add_constraint_hinge (RigidBodies.cpp:245-252): point + direction
    {
        auto& d = points;
        d.conn.push_back({(int32_t)d.conn.size(), a, b});
        push3(d.a_loc, matTvec(rb.R0[a], p - get3(rb.t0, a))); push3(d.b_loc, matTvec(rb.R0[b], p - get3(rb.t0, b)));
        push1(d.stiffness, default_stiffness); push1(d.is_active, 1.0); d.tolerance_in_m.push_back(default_tolerance_in_m);
    }
    {
        auto& d = directions;
        d.conn.push_back({(int32_t)d.conn.size(), a, b});
        push3(d.da_loc, matTvec(rb.R0[a], dg)); push3(d.db_loc, matTvec(rb.R0[b], dg));
        push1(d.stiffness, default_stiffness); push1(d.is_active, 1.0); d.tolerance_in_deg.push_back(default_tolerance_in_deg);
    }
}
bool EnergyRigidBodyConstraints::adjust_stiffness(const RigidBodyDynamics& rb, double dt, double, double mult, bool set)
{   // S/models/rigidbodies/EnergyRigidBodyConstraints.cpp:268-398 (the four constraint kinds built here)
    auto x1 = [&](int b, const Vec3& l) { return set ? rb.position_at(b, l) : rb.get_x1(b, l, dt); };
    auto d1 = [&](int b, const Vec3& l) { return set ? rb.direction(b, l) : rb.get_d1(b, l, dt); };
    auto deg = [](double C) { return std::asin(C) * 180.0 / M_PI; };
    bool valid = true;
    for (auto& row : global_points.conn) {
        const double C = norm(x1(row[1], get3(global_points.loc, row[0])) - get3(global_points.target_glob, row[0]));
        if (C > global_points.tolerance_in_m[row[0]]) { valid = false; global_points.stiffness.data[row[0]] *= mult; }
    }
    for (auto& row : global_directions.conn) {
        const double C = norm(d1(row[1], get3(global_directions.d_loc, row[0])) - get3(global_directions.target_d_glob, row[0]));
        if (deg(C) > global_directions.tolerance_in_deg[row[0]]) { valid = false; global_directions.stiffness.data[row[0]] *= mult; }
    }
    for (auto& row : points.conn) {
        const double C = norm(x1(row[2], get3(points.b_loc, row[0])) - x1(row[1], get3(points.a_loc, row[0])));
        if (C > points.tolerance_in_m[row[0]]) { valid = false; points.stiffness.data[row[0]] *= mult; }
    }
    for (auto& row : directions.conn) {
        const double C = norm(d1(row[2], get3(directions.db_loc, row[0])) - d1(row[1], get3(directions.da_loc, row[0])));
        if (deg(C) > directions.tolerance_in_deg[row[0]]) { valid = false; directions.stiffness.data[row[0]] *= mult; }
    }
    return valid;
}

// ---------------------------------------------------------------------------------------------------------------------
// contact
// ---------------------------------------------------------------------------------------------------------------------
static double init_thickness(const EnergyFrictionalContact& c, double t)
{   // EnergyFrictionalContact::_init_contact_thickness (S/models/interactions/EnergyFrictionalContact.cpp:131-148)
    if (t == 0.0) {
        if (c.global_params.default_contact_thickness > 0.0) return c.global_params.default_contact_thickness;
        die("Undefined contact thickness found. Explicitly declare per-object contact thickness or set a global default.");
    }
    return t;
}
int EnergyFrictionalContact::add_triangles_deformable(int set, const std::vector<int32_t>& vg, const std::vector<std::array<int, 3>>& tris, double th)
{
    Mesh m;
    m.ps = 0; m.idx_in_ps = set; m.vertex_global = vg; m.thickness = init_thickness(*this, th);
    for (auto& t : tris) for (int v : t) m.triangles.push_back(v);
    for (auto& e : find_edges_from_triangles(tris, (int)vg.size())) { m.edges.push_back(e[0]); m.edges.push_back(e[1]); }
    meshes.push_back(m);
    return (int)meshes.size() - 1;
}
int EnergyFrictionalContact::add_triangles_rigid(int body, const std::vector<Vec3>& V, const std::vector<std::array<int, 3>>& tris, double th)
{
    Mesh m;
    m.ps = 1; m.idx_in_ps = body; m.thickness = init_thickness(*this, th);
    for (auto& v : V) for (double c : v) m.vertices_local.push_back(c);
    for (auto& t : tris) for (int v : t) m.triangles.push_back(v);
    for (auto& e : find_edges_from_triangles(tris, (int)V.size())) { m.edges.push_back(e[0]); m.edges.push_back(e[1]); }
    meshes.push_back(m);
    return (int)meshes.size() - 1;
}
void EnergyFrictionalContact::set_friction(int a, int b, double mu) { friction_pairs_idx.push_back({a, b}); friction_pairs_mu.push_back(mu); }
void EnergyFrictionalContact::disable_collision(int a, int b) { disabled.push_back({a, b}); }

// ---------------------------------------------------------------------------------------------------------------------
// Simulation
// ---------------------------------------------------------------------------------------------------------------------
Simulation::Simulation(const Settings& s) : settings(s)
{
    dt = s.simulation.max_time_step_size;
    gravity = s.simulation.gravity;
    const int r = sb_create(&ctx, s.device, s.stream);
    if (r != SB_OK) die("sb_create failed (no CUDA device? this library has no CPU fallback), status " + std::to_string(r));
}
Simulation::~Simulation()
{
    if (std::getenv("SB_HOST_DUMP") && phase_steps > 0)
        std::fprintf(stderr, "[stark_b200 host] %lld steps: pre %.3f ms, solve %.3f ms, post %.3f ms, roll %.3f ms per step\n", phase_steps,
                     1e3 * phase_s[0] / phase_steps, 1e3 * phase_s[1] / phase_steps, 1e3 * phase_s[2] / phase_steps, 1e3 * phase_s[3] / phase_steps);
    sb_destroy(ctx);
}

void Simulation::check(int status, const char* what)
{
    if (status != SB_OK) die(std::string(what) + ": " + sb_last_error(ctx));
}
void Simulation::reg(DeviceArray& a, const char* label, int stride)
{
    a.stride = stride;
    a.label = label;
    check(sb_array_create(ctx, label, stride, &a.id), "sb_array_create");
    all_arrays.push_back(&a);
}
void Simulation::upload(DeviceArray& a)
{
    if (a.device_current) return;   // rolled on the device: nothing to send
    if (!a.volatile_data) {
        if (a.has_shadow && a.shadow == a.data) return;   // the device already holds these values
        a.shadow = a.data;
        a.has_shadow = true;
    }
    check(sb_array_upload(ctx, a.id, a.data.data(), a.rows()), "sb_array_upload");
    h2d_bytes += (long long)a.data.size() * 8;
    // per-node accelerations / forces have no setter after initialisation in this layer: one upload, no 0.9 MB compare per step
    if (&a == &dyn.a || &a == &dyn.f) a.device_current = true;
}
void Simulation::pin(DeviceArray& a)
{   // large per-step arrays become pinned mirrors: their uploads are asynchronous DMA without a staging copy
    if (a.pinned || a.data.empty()) return;
    check(sb_host_register(ctx, a.data.data(), (uint64_t)a.data.size() * 8), "sb_host_register");
    a.pinned = true;
}
void Simulation::download(DeviceArray& a)
{
    check(sb_array_download(ctx, a.id, a.data.data(), a.rows()), "sb_array_download");
    d2h_bytes += (long long)a.data.size() * 8;
}
void Simulation::sync_host()
{
    if (!host_mirror_pending) return;
    check(sb_download_wait(ctx), "sb_download_wait");
    host_mirror_pending = false;
}
int Simulation::ndofs()
{
    int n = 0;
    sb_dof_total(ctx, &n);
    return n;
}

Simulation::VolumeHandle Simulation::add_volume_grid(const Vec3& dim, const std::array<int, 3>& sub, const VolumeParams& p)
{   // DeformablesPresets::add_volume_grid -> add_volume (S/models/presets/DeformablesPresets.cpp:60-85)
    std::vector<Vec3> V;
    std::vector<std::array<int, 4>> T;
    generate_tet_grid(V, T, {0, 0, 0}, dim, sub);
    std::vector<std::array<int, 3>> tris;
    std::vector<int> tri_to_tet;
    find_surface(tris, tri_to_tet, V, T);
    const int set = dyn.add(V);
    lumped_inertia.add(dyn, set, T, p.density, p.inertia_damping);
    tet_strain.add(dyn, set, T, p);
    int group = -1;
    if (settings.simulation.init_frictional_contact) {
        std::vector<int32_t> vg;
        for (int v : tri_to_tet) vg.push_back(dyn.get_global_index(set, v));
        group = contact.add_triangles_deformable(set, vg, tris, p.contact_thickness);
    }
    return {set, group, (int)V.size(), (int)T.size()};
}

Simulation::SurfaceHandle Simulation::add_surface_grid(const std::array<double, 2>& dim, const std::array<int, 2>& sub, const SurfaceParams& p)
{   // DeformablesPresets::add_surface_grid -> add_surface (S/models/presets/DeformablesPresets.cpp:31-50)
    std::vector<Vec3> V;
    std::vector<std::array<int, 3>> T;
    generate_triangle_grid(V, T, {0.0, 0.0}, dim, sub);
    const int set = dyn.add(V);
    lumped_inertia.add(dyn, set, T, p.density, p.inertia_damping);
    triangle_strain.add(dyn, set, T, p);
    discrete_shells.add(dyn, set, T, p);
    int group = -1;
    if (settings.simulation.init_frictional_contact) {
        std::vector<int32_t> vg;
        for (int v = 0; v < (int)V.size(); v++) vg.push_back(dyn.get_global_index(set, v));
        group = contact.add_triangles_deformable(set, vg, T, p.contact_thickness);
    }
    return {set, group, (int)V.size(), (int)T.size()};
}

Simulation::BoxHandle Simulation::add_box(double mass, const Vec3& size, double thickness)
{   // RigidBodyPresets::add_box (S/models/presets/RigidBodyPresets.cpp:47-53): par_shapes unit cube, centred, scaled
    static const double C[8][3] = {{0, 0, 0}, {0, 1, 0}, {1, 1, 0}, {1, 0, 0}, {0, 0, 1}, {0, 1, 1}, {1, 1, 1}, {1, 0, 1}};
    static const int F[12][3] = {{7, 6, 5}, {5, 4, 7}, {0, 1, 2}, {2, 3, 0}, {6, 7, 3}, {3, 2, 6}, {5, 6, 2}, {2, 1, 5}, {4, 5, 1}, {1, 0, 4}, {7, 4, 0}, {0, 3, 7}};
    std::vector<Vec3> V;
    for (auto& c : C) V.push_back({(c[0] - 0.5) * size[0], (c[1] - 0.5) * size[1], (c[2] - 0.5) * size[2]});
    std::vector<std::array<int, 3>> T;
    for (auto& f : F) T.push_back({f[0], f[1], f[2]});
    const int body = rb.add();
    // inertia_tensor_box (S/models/rigidbodies/inertia_tensors.cpp)
    const double Ix = mass / 12.0 * (size[1] * size[1] + size[2] * size[2]), Iy = mass / 12.0 * (size[0] * size[0] + size[2] * size[2]), Iz = mass / 12.0 * (size[0] * size[0] + size[1] * size[1]);
    rb_inertia.add(body, mass, {Ix, 0, 0, 0, Iy, 0, 0, 0, Iz});
    int group = -1;
    if (settings.simulation.init_frictional_contact) {
        group = contact.add_triangles_rigid(body, V, T, thickness);
        contact.disable_collision(group, group);
    }
    return {body, group};
}
void Simulation::set_translation(int body, const Vec3& t) { set3(rb.t0, body, t); }

static void push_fetch(std::vector<sb_fetch>& f, const DeviceArray& a, int col, int slot) { f.push_back({a.id, col, slot, a.stride}); }

void Simulation::initialize()
{
    is_init = true;
    dt_arr.data = {dt};
    gravity_arr.data = {gravity[0], gravity[1], gravity[2]};
    // ---- arrays (symx::DataMap bindings) ----
    reg(dyn.v1, "soft.v1", 3); reg(dyn.X, "dyn.X", 3); reg(dyn.x0, "dyn.x0", 3); reg(dyn.v0, "dyn.v0", 3); reg(dyn.a, "dyn.a", 3); reg(dyn.f, "dyn.f", 3);
    reg(rb.v1, "rigid.v1", 3); reg(rb.w1, "rigid.w1", 3); reg(rb.t0, "rb.t0", 3); reg(rb.q0_, "rb.q0_", 4); reg(rb.v0, "rb.v0", 3); reg(rb.w0, "rb.w0", 3);
    reg(rb.a, "rb.a", 3); reg(rb.aa, "rb.aa", 3); reg(rb.force, "rb.force", 3); reg(rb.torque, "rb.torque", 3);
    reg(dt_arr, "dt", 1); reg(gravity_arr, "gravity", 3);
    auto& li = lumped_inertia;
    reg(li.lumped_volume, "lumped_volume", 1); reg(li.density, "density", 1); reg(li.damping, "damping", 1); reg(li.is_quasistatic, "is_quasistatic", 1);
    auto& ts = tet_strain;
    reg(ts.scale, "tet.scale", 1); reg(ts.youngs_modulus, "tet.E", 1); reg(ts.poissons_ratio, "tet.nu", 1); reg(ts.strain_limit, "tet.strain_limit", 1);
    reg(ts.strain_limit_stiffness, "tet.sl_stiffness", 1); reg(ts.strain_damping, "tet.damping", 1);
    auto& tri = triangle_strain;
    reg(tri.scale, "tri.scale", 1); reg(tri.thickness, "tri.thickness", 1); reg(tri.youngs_modulus, "tri.E", 1); reg(tri.poissons_ratio, "tri.nu", 1);
    reg(tri.strain_damping, "tri.damping", 1); reg(tri.strain_limit, "tri.strain_limit", 1); reg(tri.strain_limit_stiffness, "tri.sl_stiffness", 1); reg(tri.inflation, "tri.inflation", 1);
    auto& ds = discrete_shells;
    reg(ds.rest_dihedral_angle_rad, "shells.rest_angle", 1); reg(ds.rest_edge_length, "shells.rest_edge_length", 1); reg(ds.rest_height, "shells.rest_height", 1);
    reg(ds.bergou_K, "shells.bergou_K", 4); reg(ds.bergou_coef, "shells.bergou_coef", 1);
    reg(ds.scale, "shells.scale", 1); reg(ds.bending_stiffness, "shells.stiffness", 1); reg(ds.bending_damping, "shells.damping", 1);
    auto& pp = prescribed_positions;
    reg(pp.target_positions, "prescribed.target", 3); reg(pp.stiffness, "prescribed.stiffness", 1);
    auto& ri = rb_inertia;
    reg(ri.mass, "rb.mass", 1); reg(ri.linear_damping, "rb.linear_damping", 1); reg(ri.angular_damping, "rb.angular_damping", 1); reg(ri.is_quasistatic, "rb.is_quasistatic", 1); reg(ri.J0_glob, "rb.J0_glob", 9);
    auto& gp = rb_constraints.global_points;
    reg(gp.loc, "gp.loc", 3); reg(gp.target_glob, "gp.target", 3); reg(gp.stiffness, "gp.stiffness", 1); reg(gp.is_active, "gp.is_active", 1);
    auto& gd = rb_constraints.global_directions;
    reg(gd.d_loc, "gd.d_loc", 3); reg(gd.target_d_glob, "gd.target", 3); reg(gd.stiffness, "gd.stiffness", 1); reg(gd.is_active, "gd.is_active", 1);
    auto& cp = rb_constraints.points;
    reg(cp.a_loc, "points.a_loc", 3); reg(cp.b_loc, "points.b_loc", 3); reg(cp.stiffness, "points.stiffness", 1); reg(cp.is_active, "points.is_active", 1);
    auto& cd = rb_constraints.directions;
    reg(cd.da_loc, "directions.da_loc", 3); reg(cd.db_loc, "directions.db_loc", 3); reg(cd.stiffness, "directions.stiffness", 1); reg(cd.is_active, "directions.is_active", 1);
    for (DeviceArray* a : {&dyn.x0, &dyn.v0, &dyn.v1}) { a->volatile_data = true; pin(*a); }
    for (DeviceArray* a : {&rb.v1, &rb.w1}) a->volatile_data = true;   // DoF arrays: the solve rewrites them on the device
    for (DeviceArray* a : all_arrays) upload(*a);

    // ---- DoF sets in the reference's registration order: soft.v1, rigid.v1, rigid.w1 ----
    check(sb_dof_add(ctx, dyn.v1.id, nullptr), "sb_dof_add");
    check(sb_dof_add(ctx, rb.v1.id, nullptr), "sb_dof_add");
    check(sb_dof_add(ctx, rb.w1.id, nullptr), "sb_dof_add");

    // ---- potentials: fetch tables in the order the reference's lambdas create their symbols ----
    auto create = [&](const char* name, int conn_stride, const std::vector<sb_fetch>& f, const int32_t* conn, int n) {
        int pot = -1;
        check(sb_potential_create(ctx, name, conn_stride, f.data(), (int)f.size(), &pot), name);
        check(sb_potential_set_connectivity(ctx, pot, conn, n), name);
        return pot;
    };
    {   // EnergyLumpedInertia (S/models/deformables/point/EnergyLumpedInertia.cpp:16-26)
        std::vector<sb_fetch> f;
        push_fetch(f, dyn.v1, 1, 0); push_fetch(f, dyn.x0, 1, 3); push_fetch(f, dyn.v0, 1, 6); push_fetch(f, dyn.a, 1, 9); push_fetch(f, dyn.f, 1, 12);
        push_fetch(f, li.lumped_volume, 0, 15); push_fetch(f, li.density, 2, 16); push_fetch(f, li.damping, 2, 17); push_fetch(f, li.is_quasistatic, 2, 18);
        push_fetch(f, dt_arr, -1, 19); push_fetch(f, gravity_arr, -1, 20);
        li.potential = create("EnergyLumpedInertia", 3, f, li.conn.empty() ? nullptr : li.conn[0].data(), (int)li.conn.size());
    }
    {   // EnergyPrescribedPositions (point/EnergyPrescribedPositions.cpp:19-23)
        std::vector<sb_fetch> f;
        push_fetch(f, dyn.v1, 1, 0); push_fetch(f, dyn.x0, 1, 3); push_fetch(f, pp.target_positions, 0, 6); push_fetch(f, pp.stiffness, 2, 9); push_fetch(f, dt_arr, -1, 10);
        pp.potential = create("EnergyPrescribedPositions", 3, f, pp.conn.empty() ? nullptr : pp.conn[0].data(), (int)pp.conn.size());
    }
    for (int complete = 1; complete >= 0; complete--) {   // EnergyTriangleStrain (surface/EnergyTriangleStrain.cpp:13-34, 82-98)
        std::vector<sb_fetch> f;
        for (int k = 0; k < 3; k++) push_fetch(f, dyn.v1, 2 + k, 3 * k);
        for (int k = 0; k < 3; k++) push_fetch(f, dyn.x0, 2 + k, 9 + 3 * k);
        for (int k = 0; k < 3; k++) push_fetch(f, dyn.X, 2 + k, 18 + 3 * k);
        push_fetch(f, tri.scale, 1, 27); push_fetch(f, tri.thickness, 1, 28); push_fetch(f, tri.youngs_modulus, 1, 29); push_fetch(f, tri.poissons_ratio, 1, 30);
        if (complete) {
            push_fetch(f, tri.strain_damping, 1, 31); push_fetch(f, tri.strain_limit, 1, 32); push_fetch(f, tri.strain_limit_stiffness, 1, 33);
            push_fetch(f, tri.inflation, 1, 34); push_fetch(f, dt_arr, -1, 35);
        } else { push_fetch(f, tri.inflation, 1, 31); push_fetch(f, dt_arr, -1, 32); }
        auto& conn = complete ? tri.conn_complete : tri.conn_elasticity_only;
        const int pot = create(complete ? "EnergyTriangleStrain" : "EnergyTriangleStrain_Elasticity_Only", 5, f, conn.empty() ? nullptr : conn[0].data(), (int)conn.size());
        (complete ? tri.potential_complete : tri.potential_elasticity_only) = pot;
    }
    {   // EnergyDiscreteShells / EnergyBendingFlat (surface/EnergyDiscreteShells.cpp:26-62, 64-92)
        std::vector<sb_fetch> f;
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.v1, 2 + k, 3 * k);
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.x0, 2 + k, 12 + 3 * k);
        push_fetch(f, ds.rest_dihedral_angle_rad, 0, 24); push_fetch(f, ds.rest_edge_length, 0, 25); push_fetch(f, ds.rest_height, 0, 26);
        push_fetch(f, ds.scale, 1, 27); push_fetch(f, ds.bending_stiffness, 1, 28); push_fetch(f, ds.bending_damping, 1, 29); push_fetch(f, dt_arr, -1, 30);
        ds.potential_complete = create("EnergyDiscreteShells", 6, f, ds.conn_complete.empty() ? nullptr : ds.conn_complete[0].data(), (int)ds.conn_complete.size());
        f.clear();
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.v1, 2 + k, 3 * k);
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.x0, 2 + k, 12 + 3 * k);
        push_fetch(f, ds.bergou_K, 0, 24); push_fetch(f, ds.bergou_coef, 0, 28); push_fetch(f, ds.bending_stiffness, 1, 29); push_fetch(f, dt_arr, -1, 30);
        ds.potential_flat_rest = create("EnergyBendingFlat", 6, f, ds.conn_flat_rest.empty() ? nullptr : ds.conn_flat_rest[0].data(), (int)ds.conn_flat_rest.size());
    }
    for (int complete = 1; complete >= 0; complete--) {   // EnergyTetStrain (volume/EnergyTetStrain.cpp:16-28, 84-93)
        std::vector<sb_fetch> f;
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.v1, 2 + k, 3 * k);
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.x0, 2 + k, 12 + 3 * k);
        for (int k = 0; k < 4; k++) push_fetch(f, dyn.X, 2 + k, 24 + 3 * k);
        push_fetch(f, ts.scale, 1, 36); push_fetch(f, ts.youngs_modulus, 1, 37); push_fetch(f, ts.poissons_ratio, 1, 38);
        if (complete) { push_fetch(f, ts.strain_limit, 1, 39); push_fetch(f, ts.strain_limit_stiffness, 1, 40); push_fetch(f, ts.strain_damping, 1, 41); push_fetch(f, dt_arr, -1, 42); }
        else push_fetch(f, dt_arr, -1, 39);
        auto& conn = complete ? ts.conn_complete : ts.conn_elasticity_only;
        const int pot = create(complete ? "EnergyTetStrain" : "EnergyTetStrain_Elasticity_Only", 6, f, conn.empty() ? nullptr : conn[0].data(), (int)conn.size());
        (complete ? ts.potential_complete : ts.potential_elasticity_only) = pot;
    }
    {   // EnergyRigidBodyInertia (S/models/rigidbodies/EnergyRigidBodyInertia.cpp:16-24, 45-52)
        std::vector<sb_fetch> f;
        push_fetch(f, rb.v1, 0, 0); push_fetch(f, rb.v0, 0, 3); push_fetch(f, rb.a, 0, 6); push_fetch(f, rb.force, 0, 9); push_fetch(f, ri.mass, 0, 12);
        push_fetch(f, ri.linear_damping, 0, 13); push_fetch(f, ri.is_quasistatic, 0, 14); push_fetch(f, dt_arr, -1, 15); push_fetch(f, gravity_arr, -1, 16);
        ri.potential_linear = create("EnergyRigidBodyInertia_Linear", 1, f, ri.conn.empty() ? nullptr : ri.conn[0].data(), (int)ri.conn.size());
        f.clear();
        push_fetch(f, rb.w1, 0, 0); push_fetch(f, rb.w0, 0, 3); push_fetch(f, rb.aa, 0, 6); push_fetch(f, rb.torque, 0, 9); push_fetch(f, ri.J0_glob, 0, 12);
        push_fetch(f, ri.angular_damping, 0, 21); push_fetch(f, ri.is_quasistatic, 0, 22); push_fetch(f, dt_arr, -1, 23);
        ri.potential_angular = create("EnergyRigidBodyInertia_Angular", 1, f, ri.conn.empty() ? nullptr : ri.conn[0].data(), (int)ri.conn.size());
    }
    {   // rigid body constraints (S/models/rigidbodies/EnergyRigidBodyConstraints.cpp:30-156); every listed constraint is active
        std::vector<sb_fetch> f;
        push_fetch(f, gp.loc, 0, 0); push_fetch(f, gp.target_glob, 0, 3); push_fetch(f, gp.stiffness, 0, 6); push_fetch(f, gp.is_active, 0, 7); push_fetch(f, dt_arr, -1, 8);
        push_fetch(f, rb.v1, 1, 9); push_fetch(f, rb.w1, 1, 12); push_fetch(f, rb.t0, 1, 15); push_fetch(f, rb.q0_, 1, 18);
        gp.potential = create("rb_constraint_global_points", 2, f, gp.conn.empty() ? nullptr : gp.conn[0].data(), (int)gp.conn.size());
        f.clear();
        push_fetch(f, gd.d_loc, 0, 0); push_fetch(f, gd.target_d_glob, 0, 3); push_fetch(f, gd.stiffness, 0, 6); push_fetch(f, gd.is_active, 0, 7); push_fetch(f, dt_arr, -1, 8);
        push_fetch(f, rb.w1, 1, 9); push_fetch(f, rb.q0_, 1, 12);
        gd.potential = create("rb_constraint_global_directions", 2, f, gd.conn.empty() ? nullptr : gd.conn[0].data(), (int)gd.conn.size());
        f.clear();
        push_fetch(f, cp.a_loc, 0, 0); push_fetch(f, cp.b_loc, 0, 3); push_fetch(f, cp.stiffness, 0, 6); push_fetch(f, cp.is_active, 0, 7); push_fetch(f, dt_arr, -1, 8);
        push_fetch(f, rb.v1, 1, 9); push_fetch(f, rb.w1, 1, 12); push_fetch(f, rb.t0, 1, 15); push_fetch(f, rb.q0_, 1, 18);
        push_fetch(f, rb.v1, 2, 22); push_fetch(f, rb.w1, 2, 25); push_fetch(f, rb.t0, 2, 28); push_fetch(f, rb.q0_, 2, 31);
        cp.potential = create("rb_constraint_points", 3, f, cp.conn.empty() ? nullptr : cp.conn[0].data(), (int)cp.conn.size());
        f.clear();
        push_fetch(f, cd.da_loc, 0, 0); push_fetch(f, cd.db_loc, 0, 3); push_fetch(f, cd.stiffness, 0, 6); push_fetch(f, cd.is_active, 0, 7); push_fetch(f, dt_arr, -1, 8);
        push_fetch(f, rb.w1, 1, 9); push_fetch(f, rb.q0_, 1, 12); push_fetch(f, rb.w1, 2, 16); push_fetch(f, rb.q0_, 2, 19);
        cd.potential = create("rb_constraint_directions", 3, f, cd.conn.empty() ? nullptr : cd.conn[0].data(), (int)cd.conn.size());
    }
    // ---- contact (S/models/interactions/EnergyFrictionalContact.cpp:14-40) ----
    if (settings.simulation.init_frictional_contact && !contact.is_empty()) {
        sb_contact_bindings b;
        b.soft_v1 = dyn.v1.id; b.soft_x0 = dyn.x0.id; b.soft_X = dyn.X.id; b.rb_v1 = rb.v1.id; b.rb_w1 = rb.w1.id; b.rb_t0 = rb.t0.id; b.rb_q0 = rb.q0_.id; b.dt = dt_arr.id;
        check(sb_contact_init(ctx, &b), "sb_contact_init");
        for (auto& m : contact.meshes) {
            sb_contact_mesh cm;
            cm.physical_system = m.ps; cm.rigid_body = m.idx_in_ps;
            cm.n_vertices = (m.ps == 0) ? (int)m.vertex_global.size() : (int)m.vertices_local.size() / 3;
            cm.n_triangles = (int)m.triangles.size() / 3; cm.n_edges = (int)m.edges.size() / 2;
            cm.vertex_global = m.vertex_global.data(); cm.vertices_local = m.vertices_local.data(); cm.triangles = m.triangles.data(); cm.edges = m.edges.data();
            cm.contact_thickness = m.thickness;
            check(sb_contact_add_mesh(ctx, &cm, nullptr), "sb_contact_add_mesh");
        }
        for (auto& p : contact.disabled) check(sb_contact_blacklist(ctx, p[0], p[1]), "sb_contact_blacklist");
        for (size_t i = 0; i < contact.friction_pairs_idx.size(); i++)
            check(sb_contact_set_friction(ctx, contact.friction_pairs_idx[i][0], contact.friction_pairs_idx[i][1], contact.friction_pairs_mu[i]), "sb_contact_set_friction");
    }
    settings.newton.contact_enabled = (settings.simulation.init_frictional_contact && !contact.is_empty() && contact.global_params.collisions_enabled) ? 1 : 0;
    // Stark::_initialize: the initial state must be valid (S/core/Stark.cpp:307-311)
    if (settings.newton.contact_enabled && contact.global_params.intersection_test_enabled) {
        check(sb_contact_set_params(ctx, contact.contact_stiffness, contact.global_params.friction_stick_slide_threshold, contact.global_params.triangle_point_enabled,
                                    contact.global_params.edge_edge_enabled, contact.global_params.friction_enabled), "sb_contact_set_params");
        int n = 0;
        check(sb_contact_count_intersections(ctx, &n), "sb_contact_count_intersections");
        if (n > 0) die("Initial state is not valid. Exiting simulation.");
    }
}

bool Simulation::run_one_time_step()
{
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    for (auto& ev : time_events) ev(current_time);   // EventDrivenScript::run_a_cycle (S/models/Simulation.cpp:75)
    if (!is_init) initialize();
    stats = StepStats();
    stats.dt = dt;

    // should_continue_execution (EnergyFrictionalContact.cpp:812-823)
    if (settings.newton.contact_enabled && contact.contact_stiffness > contact.global_params.max_contact_stiffness) {
        std::cout << "Contact stiffness exceeded maximum value." << std::endl;
        return false;
    }

    // ---- before_time_step callbacks, in registration order ----
    // v1 <- 0 (PointDynamics.cpp:58-62) happens on the device below; the host mirror of v1 is only ever read after a download
    // (prescribed positions with a finite tolerance), so the 0.9 MB host fill per step is kept for that case only
    if (prescribed_positions.checks_tolerance()) std::fill(dyn.v1.data.begin(), dyn.v1.data.end(), 0.0);
    for (int b = 0; b < rb.get_n_bodies(); b++) for (int c = 0; c < 4; c++) rb.q0_.data[4 * b + c] = rb.q0[b][c];   // RigidBodyDynamics.cpp:136-147
    std::fill(rb.v1.data.begin(), rb.v1.data.end(), 0.0);
    std::fill(rb.w1.data.begin(), rb.w1.data.end(), 0.0);
    rb_inertia.before_time_step(rb);                                                          // EnergyRigidBodyInertia.cpp:85-104
    // per-step state to the device (SURVEY.md Appendix B "once per time step" + "scalars that may change between retries")
    dt_arr.data[0] = dt;
    gravity_arr.data = {gravity[0], gravity[1], gravity[2]};
    check(sb_array_fill(ctx, dyn.v1.id, dyn.v1.rows(), 0.0), "sb_array_fill");   // v1 <- 0 on the device (no transfer)
    check(sb_array_fill(ctx, rb.v1.id, rb.v1.rows(), 0.0), "sb_array_fill");
    check(sb_array_fill(ctx, rb.w1.id, rb.w1.rows(), 0.0), "sb_array_fill");
    for (DeviceArray* a : {&dyn.x0, &dyn.v0, &dyn.a, &dyn.f, &rb.t0, &rb.q0_, &rb.v0, &rb.w0, &rb.a, &rb.aa, &rb.force, &rb.torque, &dt_arr, &gravity_arr,
                           &rb_inertia.J0_glob, &prescribed_positions.target_positions, &prescribed_positions.stiffness, &rb_constraints.global_points.target_glob,
                           &rb_constraints.global_points.stiffness, &rb_constraints.global_directions.target_d_glob, &rb_constraints.global_directions.stiffness,
                           &rb_constraints.global_directions.d_loc,   // (scripted fix constraints rotate the LOCAL lock directions)
                           &rb_constraints.points.stiffness, &rb_constraints.directions.stiffness})
        upload(*a);
    check(sb_newton_timer_begin(ctx), "sb_newton_timer_begin");   // (the step's device time covers the start-of-step detection and the pre-launched evaluation)
    if (settings.newton.contact_enabled) {
        check(sb_contact_set_params(ctx, contact.contact_stiffness, contact.global_params.friction_stick_slide_threshold, contact.global_params.triangle_point_enabled,
                                    contact.global_params.edge_edge_enabled, contact.global_params.friction_enabled), "sb_contact_set_params");
        // friction tables (EnergyFrictionalContact.cpp:531-773); v1 / w1 are zero here, so the same detection also serves the
        // solve's initial validity test and first contact update -- and the volume elements of the first evaluation, which do not
        // depend on it, are started first and run beside it
        check(sb_eval_prelaunch(ctx), "sb_eval_prelaunch");
        check(sb_contact_begin_time_step(ctx, 1), "sb_contact_begin_time_step");
    }

    // ---- Newton solve on the device ----
    sb_newton_settings ns = settings.newton;
    ns.intersection_test_enabled = contact.global_params.intersection_test_enabled ? 1 : 0;
    ns.skip_converged_state_check = 1;   // the host runs the converged-state callbacks below, in the reference's order
    sb_newton_stats st;
    const auto t_solve = clk::now();
    check(sb_newton_solve(ctx, &ns, &st), "sb_newton_solve");
    const auto t_solved = clk::now();
    // the rigid bodies (a handful) are rolled on the host; the deformable state stays on the device unless a converged-state
    // callback needs it here (prescribed positions with a finite tolerance)
    download(rb.v1); download(rb.w1);
    const bool need_v1 = prescribed_positions.checks_tolerance();
    if (need_v1) { sync_host(); download(dyn.v1); }
    int result = st.result;
    if (result == 0) {
        // is_converged_state_valid callbacks, short-circuiting like SolverCallbacks::_run_bool (solver_utils.h:55-62):
        // prescribed positions -> rigid body constraints -> contact
        bool valid = !need_v1 || prescribed_positions.is_converged_state_valid(dyn, dt);
        valid = valid && rb_constraints.adjust_stiffness(rb, dt, 1.0, rb_constraints.stiffness_hard_multiplier, false);
        if (valid && settings.newton.contact_enabled && contact.global_params.intersection_test_enabled) {
            int n = 0;
            check(sb_contact_count_intersections(ctx, &n), "sb_contact_count_intersections");
            valid = (n == 0);
        }
        if (!valid) result = 8;   // InvalidConvergedState
    }
    if (result == 6 && settings.newton.contact_enabled) contact.contact_stiffness *= 2.0;    // on_intermediate_state_invalid (EnergyFrictionalContact.cpp:800-806)
    stats.solve_s = std::chrono::duration<double>(clk::now() - t_solve).count();
    stats.result = result;
    stats.newton_iterations = st.newton_iterations; stats.cg_iterations = st.cg_iterations; stats.ls_inv = st.ls_inv_iterations; stats.ls_bt = st.ls_bt_iterations;
    stats.n_evaluations = st.n_evaluations;
    stats.solve_gpu_ms = st.gpu_ms;
    stats.first_residual = st.residuals[0];
    for (int i = 0; i < std::min(64, st.n_evaluations); i++) stats.residuals.push_back(st.residuals[i]);
    total_newton_iterations += st.newton_iterations; total_evaluations += st.n_evaluations; total_cg_iterations += st.cg_iterations;
    total_solve_s += stats.solve_s;

    const auto t_post = clk::now();
    bool keep_going = true;
    if (result == 0) {
        // ---- on_time_step_accepted (S/core/Stark.cpp:161-170) ----
        // PointDynamics.cpp:64-78 on the device: x0 += dt v1, v0 = v1; the host mirrors follow asynchronously (the copy runs under
        // the next step's solve; sync_host() before any host access)
        if (dyn.size() > 0) {
            sync_host();   // (a previous read-back still writing the mirrors)
            check(sb_array_axpy(ctx, dyn.x0.id, dyn.v1.id, dt, dyn.size()), "sb_array_axpy");
            check(sb_array_copy(ctx, dyn.v0.id, dyn.v1.id, dyn.size()), "sb_array_copy");
            dyn.x0.device_current = dyn.v0.device_current = true;
            check(sb_array_download_async(ctx, dyn.x0.id, dyn.x0.data.data(), dyn.size()), "sb_array_download_async");
            check(sb_array_download_async(ctx, dyn.v0.id, dyn.v0.data.data(), dyn.size()), "sb_array_download_async");
            d2h_bytes += 2ll * (long long)dyn.x0.data.size() * 8;
            host_mirror_pending = true;
        }
        for (int b = 0; b < rb.get_n_bodies(); b++) {                                          // RigidBodyDynamics.cpp:149-166
            set3(rb.t0, b, get3(rb.t0, b) + dt * get3(rb.v1, b));
            rb.q0[b] = quat_time_integration(rb.q0[b], get3(rb.w1, b), dt);
            rb.R0[b] = quat_to_rotation(rb.q0[b]);
            set3(rb.v0, b, get3(rb.v1, b)); set3(rb.w0, b, get3(rb.w1, b));
        }
        rb_constraints.adjust_stiffness(rb, dt, rb_constraints.soft_constraint_capacity_hardening_point, rb_constraints.stiffness_soft_multiplier, true);
        if (settings.newton.contact_enabled) contact.contact_stiffness = std::max(contact.global_params.min_contact_stiffness, 0.99 * contact.contact_stiffness);
        current_time += dt;
        current_time_step++;
        dt = std::min(settings.simulation.max_time_step_size, dt * settings.simulation.time_step_size_success_multiplier);
        stats.accepted = true;
    } else if (result == 8 || result == 6) {
        // InvalidConvergedState / TooManyInvalidIntermediateIterations: retry the same dt with the hardened parameters
    } else {
        if (!settings.simulation.use_adaptive_time_step) keep_going = false;
        else {
            dt /= 2.0;
            if (dt < settings.simulation.time_step_size_lower_bound) keep_going = false;
        }
    }
    const auto t_end = clk::now();
    phase_s[0] += std::chrono::duration<double>(t_solve - t_begin).count(); phase_s[1] += std::chrono::duration<double>(t_solved - t_solve).count();
    phase_s[2] += std::chrono::duration<double>(t_post - t_solved).count(); phase_s[3] += std::chrono::duration<double>(t_end - t_post).count();
    phase_steps++;
    static const char* hd = std::getenv("SB_HOST_DUMP");
    if (hd && hd[0] == '2')
        std::fprintf(stderr, "[stark_b200 host] step %d: pre %.3f solve %.3f (gpu %.3f) post %.3f roll %.3f ms, %d its\n", current_time_step, 1e3 * std::chrono::duration<double>(t_solve - t_begin).count(),
                     1e3 * std::chrono::duration<double>(t_solved - t_solve).count(), st.gpu_ms, 1e3 * std::chrono::duration<double>(t_post - t_solved).count(), 1e3 * std::chrono::duration<double>(t_end - t_post).count(), st.newton_iterations);
    stats.runtime_s = std::chrono::duration<double>(t_end - t_begin).count();
    return keep_going;
}

}  // namespace stark_b200
