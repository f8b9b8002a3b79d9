// Hand-written element energies for every potential STARK registers on the Newton hot path.
//
// Each struct restates ONE `global_potential->add_potential(...)` lambda of the reference as a templated
// device function `energy<T>(in, seed)`; T = double gives the value, T = sbad::D2 gives the lane's share of
// the gradient/Hessian (ad.cuh).  `in[]` has exactly the layout the reference's MappedWorkspace produces
// (concatenation of the make_* calls in the order the lambda executes them -- SURVEY.md Appendix A,
// docs/potential_layouts.txt), so a host that walks MappedWorkspace::maps can bind arrays 1:1.
// DoF blocks are numbered in the reference's order: DoF set (soft.v1, rigid.v1, rigid.w1), then map creation
// order (symx/solver/second_order/SecondOrderCompiledPotential.cpp:11-33).  `DOF_SLOT[b]` is the in[] slot where
// the b-th 3-DoF block lives; the host checks it against the fetch table at registration.
#pragma once
#include "ad.cuh"

namespace sbpot {
using namespace sbad;

// =====================================================================================================
// shared building blocks
// =====================================================================================================

// x1 = x0 + dt * v1   (S/models/time_integration.cpp:5)
template<class T> SB_HD V3<T> soft_x1(const Seed<T>& s, int k0, const double* v1, const double* x0, double dt)
{
    const V3<T> v = s.dof3(k0, v1);
    return V3<T>(x0[0] + dt * v.x, x0[1] + dt * v.y, x0[2] + dt * v.z);
}

// Rigid body frame at t_{n+1}: q1 = normalize(q0 + dt/2 (0,w) (x) q0), R1 = R(q1), t1 = t0 + dt v
// (S/models/rigidbodies/rigidbody_transformations.cpp:54-160, RigidBodyDynamics.cpp:46-61)
template<class T> struct RigidFrame {
    M3<T> R;
    V3<T> t, v, w;
};
template<class T> SB_HD M3<T> quat_to_rotation(const T& qw, const T& qx, const T& qy, const T& qz)
{
    const T tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const T twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const T txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const T tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    M3<T> R;
    R(0, 0) = 1.0 - (tyy + tzz); R(0, 1) = txy - twz;         R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;         R(1, 1) = 1.0 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;         R(2, 1) = tyz + twx;         R(2, 2) = 1.0 - (txx + tyy);
    return R;
}
template<class T> SB_HD M3<T> rotation_q1(const V3<T>& w, const double* q0, double dt)
{
    // (0, w) (x) q0 with (a,b,c,d) = (0,w) and (e,f,g,h) = q0
    const double e = q0[0], f = q0[1], g = q0[2], h = q0[3];
    const T p0 = -(w.x * f) - w.y * g - w.z * h;
    const T p1 = w.x * e + w.y * h - w.z * g;
    const T p2 = w.y * e - w.x * h + w.z * f;
    const T p3 = w.x * g - w.y * f + w.z * e;
    const double hdt = 0.5 * dt;
    T qw = e + hdt * p0, qx = f + hdt * p1, qy = g + hdt * p2, qz = h + hdt * p3;
    const T rn = inv(Sqrt(qw * qw + qx * qx + qy * qy + qz * qz));
    qw = qw * rn; qx = qx * rn; qy = qy * rn; qz = qz * rn;
    return quat_to_rotation(qw, qx, qy, qz);
}
// kv / kw: element DoF index of the body's linear / angular velocity block (-1: not a DoF of this element)
template<class T> SB_HD RigidFrame<T> rigid_frame(const Seed<T>& s, int kv, int kw, const double* v, const double* w, const double* t0, const double* q0, double dt)
{
    RigidFrame<T> f;
    f.v = s.dof3(kv, v);
    f.w = s.dof3(kw, w);
    f.R = rotation_q1(f.w, q0, dt);
    f.t = V3<T>(t0[0] + dt * f.v.x, t0[1] + dt * f.v.y, t0[2] + dt * f.v.z);
    return f;
}
template<class T> SB_HD V3<T> rigid_x1(const RigidFrame<T>& f, const double* Xloc) { return f.t + mul(f.R, ld3(Xloc)); }
template<class T> SB_HD V3<T> rigid_v1(const RigidFrame<T>& f, const double* Xloc) { return f.v + cross(f.w, mul(f.R, ld3(Xloc))); }

// distances (S/models/distances.cpp:57-109)
template<class T> SB_HD T distance_point_point(const V3<T>& p, const V3<T>& q) { return Sqrt(norm2(p - q)); }
template<class T> SB_HD T distance_point_line(const V3<T>& p, const V3<T>& a, const V3<T>& b)
{
    const V3<T> ab = b - a, ap = p - a;
    const T e = dot(ap, ab);
    return Sqrt(dot(ap, ap) - e * e / dot(ab, ab));
}
template<class T> SB_HD T distance_point_plane(const V3<T>& p, const V3<T>& a, const V3<T>& b, const V3<T>& c)
{
    const V3<T> n = normalized(cross(a - c, b - c));
    const T d = dot(p - a, n);
    return Sqrt(d * d);
}
template<class T> SB_HD T distance_line_line(const V3<T>& a, const V3<T>& b, const V3<T>& p, const V3<T>& q)
{
    const V3<T> n = cross(b - a, q - p);
    const T l = dot(p - a, n);
    return Sqrt(sq(l) / norm2(n));
}

// IPC cubic barrier  k (dhat - d)^3 / 3   (EnergyFrictionalContact.cpp:1225-1237)
template<class T> SB_HD T barrier_cubic(const T& d, double dhat, double k) { return k * cube(dhat - d) / 3.0; }

// Edge-edge mollifier (EnergyFrictionalContact.cpp:1251-1259)
template<class T> SB_HD T ee_mollifier(const V3<T>& ea0, const V3<T>& ea1, const V3<T>& eb0, const V3<T>& eb1,
                                       const double* ear0, const double* ear1, const double* ebr0, const double* ebr1)
{
    const double eps_x = 1e-3 * norm2(ld3(ear0) - ld3(ear1)) * norm2(ld3(ebr0) - ld3(ebr1));
    const T x = norm2(cross(ea1 - ea0, eb1 - eb0));
    if (val(x) - eps_x > 0.0) return T(1.0);
    const T r = x / eps_x;
    return (2.0 - r) * r;
}

// IPC friction, C0 model (EnergyFrictionalContact.cpp:1260-1289); T2x3 row-major tangent basis
template<class T> SB_HD T friction_C0(const V3<T>& v, const double* Tm, double mu, double fn, double epsv, double dt)
{
    constexpr double PERTURBATION = 1e-9;
    T u0 = (Tm[0] * v.x + Tm[1] * v.y + Tm[2] * v.z) * dt + 1.13 * PERTURBATION;
    T u1 = (Tm[3] * v.x + Tm[4] * v.y + Tm[5] * v.z) * dt - 1.07 * PERTURBATION;
    const T u = Sqrt(u0 * u0 + u1 * u1);
    const double epsu = dt * epsv;
    const double k = mu * fn / epsu;
    const double eps = mu * fn / (2.0 * k);
    if (epsu - val(u) > 0.0) return 0.5 * k * sq(u);
    return mu * fn * (u - eps);
}

// =====================================================================================================
// deformables
// =====================================================================================================

// S/models/deformables/point/EnergyLumpedInertia.cpp:12-50
struct EnergyLumpedInertia {
    static constexpr int N_IN = 23, N_DOF = 3, NB = 1;
    static constexpr int DOF_SLOT[NB] = {0};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[19];
        const V3<T> x1 = soft_x1(s, 0, in + 0, in + 3, dt);
        const V3<double> x0 = ld3(in + 3), v0 = ld3(in + 6), a = ld3(in + 9), f = ld3(in + 12), g = ld3(in + 20);
        const double mass = in[15] * in[16], damping = in[17], quasi = in[18];
        const V3<double> f_ext = mass * (a + g) + f;
        T E = -dot(f_ext, x1);
        if (!(quasi - 0.5 > 0.0)) {
            const V3<T> dev = x1 - (x0 + dt * v0);
            const V3<T> dev2 = x1 - x0;
            E += 0.5 * mass * (dot(dev, dev) / (dt * dt) + dot(dev2, dev2) * damping / dt);
        }
        return E;
    }
};

// S/models/deformables/point/EnergyPrescribedPositions.cpp:15-33
struct EnergyPrescribedPositions {
    static constexpr int N_IN = 11, N_DOF = 3, NB = 1;
    static constexpr int DOF_SLOT[NB] = {0};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> x1 = soft_x1(s, 0, in + 0, in + 3, in[10]);
        return 0.5 * in[9] * norm2(x1 - ld3(in + 6));
    }
};

// S/models/deformables/line/EnergySegmentStrain.cpp:11-56 (COMPLETE) / :57-90 (elasticity only)
template<bool COMPLETE> struct EnergySegmentStrainT {
    static constexpr int N_IN = COMPLETE ? 25 : 22, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {0, 3};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[N_IN - 1];
        const V3<T> xa = soft_x1(s, 0, in + 0, in + 6, dt), xb = soft_x1(s, 3, in + 3, in + 9, dt);
        const double scale = in[18], radius = in[19], E = in[20];
        const double l_rest = norm(scale * ld3(in + 12) - scale * ld3(in + 15));
        const double volume = M_PI * radius * radius * l_rest;
        const T l = norm(xa - xb);
        const T e = (l - l_rest) / l_rest;
        T En = volume * E * sq(e) / 2.0;
        if (COMPLETE) {
            const double damping = in[21], limit = in[22], k_sl = in[23];
            const T over = e - limit;
            if (val(over) > 0.0) En += volume * k_sl * cube(over) / 3.0;
            const double l0 = norm(ld3(in + 9) - ld3(in + 6));
            const double e0 = (l0 - l_rest) / l_rest;
            En += dt * damping * sq((e - e0) / dt) / 2.0;
        }
        return En;
    }
};
using EnergySegmentStrain = EnergySegmentStrainT<true>;
using EnergySegmentStrain_Elasticity_Only = EnergySegmentStrainT<false>;

// S/models/deformables/surface/EnergyTriangleStrain.cpp:13-81 (COMPLETE) / :82-125 (elasticity only)
template<bool COMPLETE> struct EnergyTriangleStrainT {
    static constexpr int N_IN = COMPLETE ? 36 : 33, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[N_IN - 1];
        const double scale = in[27], thickness = in[28], e = in[29], nu = in[30];
        const double inflation = COMPLETE ? in[34] : in[31];
        V3<T> x1[3];
        for (int a = 0; a < 3; a++) x1[a] = soft_x1(s, 3 * a, in + 3 * a, in + 9 + 3 * a, dt);
        const V3<double> X0 = scale * ld3(in + 18), X1 = scale * ld3(in + 21), X2 = scale * ld3(in + 24);
        const double rest_area = 0.5 * norm(cross(X0 - X2, X1 - X2));
        // triangle_jacobian (deformable_tools.cpp:7-22): project rest triangle on its own plane
        const V3<double> u = normalized(X1 - X0);
        const V3<double> n = cross(u, X2 - X0);
        const V3<double> v = normalized(cross(u, n));
        const double p0x = dot(u, X0), p0y = dot(v, X0), p1x = dot(u, X1), p1y = dot(v, X1), p2x = dot(u, X2), p2y = dot(v, X2);
        const double a00 = p1x - p0x, a01 = p2x - p0x, a10 = p1y - p0y, a11 = p2y - p0y;  // DX (2x2)
        const double rdet = 1.0 / (a00 * a11 - a01 * a10);
        const double i00 = a11 * rdet, i01 = -a01 * rdet, i10 = -a10 * rdet, i11 = a00 * rdet;  // DXinv
        // F (3x2) = [x1[1]-x1[0], x1[2]-x1[0]] DXinv
        const V3<T> d1 = x1[1] - x1[0], d2 = x1[2] - x1[0];
        const V3<T> F0 = d1 * i00 + d2 * i10, F1 = d1 * i01 + d2 * i11;
        const T C00 = dot(F0, F0), C01 = dot(F0, F1), C11 = dot(F1, F1);
        const double mu = e / (2.0 * (1.0 + nu));
        const double lambda = (e * nu) / ((1.0 + nu) * (1.0 - nu));
        const T area = 0.5 * norm(cross(x1[0] - x1[2], x1[1] - x1[2]));
        const T logJ = Log(area / rest_area);
        T density = 0.5 * mu * (C00 + C11 - 2.0) - mu * logJ + 0.5 * lambda * sq(logJ);
        if (COMPLETE) {
            const double damping = in[31], limit = in[32], k_sl = in[33];
            const T E00 = 0.5 * (C00 - 1.0), E01 = 0.5 * C01, E11 = 0.5 * (C11 - 1.0);
            {   // strain-rate damping
                const V3<double> q1 = ld3(in + 12) - ld3(in + 9), q2 = ld3(in + 15) - ld3(in + 9);
                const V3<double> G0 = q1 * i00 + q2 * i10, G1 = q1 * i01 + q2 * i11;
                const double e00 = 0.5 * (dot(G0, G0) - 1.0), e01 = 0.5 * dot(G0, G1), e11 = 0.5 * (dot(G1, G1) - 1.0);
                const T r00 = (E00 - e00) / dt, r01 = (E01 - e01) / dt, r11 = (E11 - e11) / dt;
                density += 0.5 * damping * (sq(r00) + 2.0 * sq(r01) + sq(r11));
            }
            {   // strain limiting on the eigenvalues of the 2x2 Green strain (deformable_tools.cpp:27-36)
                const T delta = Sqrt(4.0 * sq(E01) + sq(E00 - E11));
                const T s0 = 0.5 * (E00 + E11 + delta), s1 = 0.5 * (E00 + E11 - delta);
                if (val(s0) - limit > 0.0) density += k_sl * cube(s0 - limit) / 3.0;
                if (val(s1) - limit > 0.0) density += k_sl * cube(s1 - limit) / 3.0;
            }
        }
        {   // inflation
            const V3<double> x00 = ld3(in + 9), x01 = ld3(in + 12), x02 = ld3(in + 15);
            const V3<double> n0 = -normalized(cross(x01 - x00, x02 - x00));
            density += inflation * dot(n0, x1[0] + x1[1] + x1[2]) / 3.0;
        }
        return thickness * rest_area * density;
    }
};
using EnergyTriangleStrain = EnergyTriangleStrainT<true>;
using EnergyTriangleStrain_Elasticity_Only = EnergyTriangleStrainT<false>;

// dihedral angle (EnergyDiscreteShells.cpp:12-23), EPSILON = 1e-12
template<class T> SB_HD T dihedral_angle(const V3<T>* x)
{
    const V3<T> e0 = x[1] - x[0], e1 = x[2] - x[0], e2 = x[3] - x[0];
    const V3<T> n0 = cross(e0, e1);
    const V3<T> n1 = -cross(e0, e2);
    return Acos((1.0 - 1e-12) * dot(normalized(n0), normalized(n1)));
}

// S/models/deformables/surface/EnergyDiscreteShells.cpp:26-62
struct EnergyDiscreteShells {
    static constexpr int N_IN = 31, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[30];
        V3<T> x1[4];
        V3<double> x0[4];
        for (int a = 0; a < 4; a++) { x1[a] = soft_x1(s, 3 * a, in + 3 * a, in + 12 + 3 * a, dt); x0[a] = ld3(in + 12 + 3 * a); }
        const double rest_angle = in[24], scale = in[27], stiffness = in[28], damping = in[29];
        const double ratio = (in[25] * scale) / (in[26] * scale);
        const T da_1 = dihedral_angle(x1);
        const T delta = da_1 - rest_angle;
        T E = stiffness * (delta * delta) * ratio;
        const double da_0 = dihedral_angle(x0);
        E += damping * 1.0 / dt * (0.5 * sq(da_1) - da_0 * da_1) * ratio;
        return E;
    }
};

// S/models/deformables/surface/EnergyDiscreteShells.cpp:64-92 (quadratic Bergou bending)
struct EnergyBendingFlat {
    static constexpr int N_IN = 31, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[30], coef = in[28], stiffness = in[29];
        const double* K = in + 24;
        V3<T> x1[4];
        for (int a = 0; a < 4; a++) x1[a] = soft_x1(s, 3 * a, in + 3 * a, in + 12 + 3 * a, dt);
        // x^T (coef K K^T) x = coef (K.x)^2 per coordinate
        const V3<T> kx = x1[0] * K[0] + x1[1] * K[1] + x1[2] * K[2] + x1[3] * K[3];
        return 0.5 * stiffness * coef * norm2(kx);
    }
};

// S/models/deformables/volume/EnergyTetStrain.cpp:12-79 (COMPLETE) / :80-123 (elasticity only)
template<bool COMPLETE> struct EnergyTetStrainT {
    static constexpr int N_IN = COMPLETE ? 43 : 40, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[N_IN - 1];
        const double scale = in[36], e = in[37], nu = in[38];
        V3<T> x1[4];
        for (int a = 0; a < 4; a++) x1[a] = soft_x1(s, 3 * a, in + 3 * a, in + 12 + 3 * a, dt);
        M3<double> DX;
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) DX(r, c) = scale * in[24 + 3 * (c + 1) + r] - scale * in[24 + r];
        double detDX;
        const M3<double> DXinv = inv3(DX, detDX);
        const double rest_volume = detDX / 6.0;
        // F = Dx1 * DXinv
        V3<T> dx[3] = {x1[1] - x1[0], x1[2] - x1[0], x1[3] - x1[0]};
        M3<T> F;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) F(r, c) = dx[0][r] * DXinv(0, c) + dx[1][r] * DXinv(1, c) + dx[2][r] * DXinv(2, c);
        const double mu = e / (2.0 * (1.0 + nu));
        const double lambda = (e * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu));
        const double mu_ = 4.0 / 3.0 * mu, lambda_ = lambda + 5.0 / 6.0 * mu;
        const double alpha = 1.0 + mu_ / lambda_ - mu_ / (4.0 * lambda_);
        const T detF = det3(F);
        T Ic = sq(F.m[0]);
        for (int k = 1; k < 9; k++) Ic += sq(F.m[k]);
        T density = 0.5 * mu_ * (Ic - 3.0) + 0.5 * lambda_ * sq(detF - alpha) - 0.5 * mu_ * Log(Ic + 1.0);
        if (COMPLETE) {
            const double limit = in[39], k_sl = in[40], damping = in[41];
            // Green strain (upper triangle): 00 01 02 11 12 22
            T E1[6];
            {
                int q = 0;
                for (int a = 0; a < 3; a++)
                    for (int b = a; b < 3; b++) {
                        T c = F(0, a) * F(0, b) + F(1, a) * F(1, b) + F(2, a) * F(2, b);
                        E1[q++] = 0.5 * (c - (a == b ? 1.0 : 0.0));
                    }
            }
            if (damping != 0.0) {
                M3<double> F0;
                V3<double> d0[3] = {ld3(in + 15) - ld3(in + 12), ld3(in + 18) - ld3(in + 12), ld3(in + 21) - ld3(in + 12)};
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) F0(r, c) = d0[0][r] * DXinv(0, c) + d0[1][r] * DXinv(1, c) + d0[2][r] * DXinv(2, c);
                T acc(0.0);
                int q = 0;
                for (int a = 0; a < 3; a++)
                    for (int b = a; b < 3; b++) {
                        const double c0 = F0(0, a) * F0(0, b) + F0(1, a) * F0(1, b) + F0(2, a) * F0(2, b);
                        const double e0 = 0.5 * (c0 - (a == b ? 1.0 : 0.0));
                        const T r = (E1[q++] - e0) / dt;
                        acc += (a == b ? 1.0 : 2.0) * sq(r);
                    }
                density += 0.5 * damping * acc;
            }
            {   // smooth upper bound of the largest Green-strain eigenvalue, cubic penalty above the limit
                const double tr = val(E1[0]) + val(E1[3]) + val(E1[5]);
                const double m = tr / 3.0;
                const double d00 = val(E1[0]) - m, d11 = val(E1[3]) - m, d22 = val(E1[5]) - m;
                const double dn = ::sqrt(d00 * d00 + d11 * d11 + d22 * d22 + 2.0 * (val(E1[1]) * val(E1[1]) + val(E1[2]) * val(E1[2]) + val(E1[4]) * val(E1[4])));
                const double largest = m + ::sqrt(2.0 / 3.0) * dn;
                if (largest - limit > 0.0) {
                    const T trT = E1[0] + E1[3] + E1[5];
                    const T mT = trT / 3.0;
                    const T e00 = E1[0] - mT, e11 = E1[3] - mT, e22 = E1[5] - mT;
                    const T dnT = Sqrt(sq(e00) + sq(e11) + sq(e22) + 2.0 * (sq(E1[1]) + sq(E1[2]) + sq(E1[4])));
                    const T dl = mT + ::sqrt(2.0 / 3.0) * dnT - limit;
                    density += k_sl * cube(dl) / 3.0;
                }
            }
        }
        return rest_volume * density;
    }
};
using EnergyTetStrain = EnergyTetStrainT<true>;
using EnergyTetStrain_Elasticity_Only = EnergyTetStrainT<false>;

// =====================================================================================================
// rigid bodies
// =====================================================================================================

// S/models/rigidbodies/EnergyRigidBodyInertia.cpp:13-40
struct EnergyRigidBodyInertia_Linear {
    static constexpr int N_IN = 19, N_DOF = 3, NB = 1;
    static constexpr int DOF_SLOT[NB] = {0};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> v1 = s.dof3(0, in);
        const V3<double> v0 = ld3(in + 3), a = ld3(in + 6), f = ld3(in + 9), g = ld3(in + 16);
        const double m = in[12], damping = in[13], quasi = in[14], dt = in[15];
        const V3<double> f_ext = m * (a + g) + f;
        T E = -dt * dot(f_ext, v1);
        if (!(quasi - 0.5 > 0.0)) {
            const V3<T> dev = v1 - v0;
            E += 0.5 * m * dot(dev, dev) + 0.5 * m * dot(v1, v1) * damping * dt;
        }
        return E;
    }
};

// S/models/rigidbodies/EnergyRigidBodyInertia.cpp:42-67
struct EnergyRigidBodyInertia_Angular {
    static constexpr int N_IN = 24, N_DOF = 3, NB = 1;
    static constexpr int DOF_SLOT[NB] = {0};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> w1 = s.dof3(0, in);
        const V3<double> w0 = ld3(in + 3), aa = ld3(in + 6), t = ld3(in + 9);
        M3<double> J;
        for (int k = 0; k < 9; k++) J.m[k] = in[12 + k];
        const double damping = in[21], quasi = in[22], dt = in[23];
        const V3<double> t_ext = mul(J, aa) + t;
        T E = -dt * dot(t_ext, w1);
        if (!(quasi - 0.5 > 0.0)) {
            const V3<T> dev = w1 - w0;
            E += 0.5 * (dot(dev, mul(J, dev)) + dot(w1, mul(J, w1)) * damping * dt);
        }
        return E;
    }
};

// S/models/rigidbodies/EnergyRigidBodyConstraints.cpp:30-45 + RigidBodyConstraints.h:110-113
// (the `is_active` condition is resolved by the host, which only lists active constraints)
struct rb_constraint_global_points {
    static constexpr int N_IN = 22, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {9, 12};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 0, 3, in + 9, in + 12, in + 15, in + 18, in[8]);
        return 0.5 * in[6] * norm2(ld3(in + 3) - rigid_x1(f, in + 0));
    }
};

// EnergyRigidBodyConstraints.cpp:47-62 + RigidBodyConstraints.h:150-153
struct rb_constraint_global_directions {
    static constexpr int N_IN = 16, N_DOF = 3, NB = 1;
    static constexpr int DOF_SLOT[NB] = {9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> w = s.dof3(0, in + 9);
        const M3<T> R = rotation_q1(w, in + 12, in[8]);
        return 0.5 * in[6] * norm2(ld3(in + 3) - mul(R, ld3(in + 0)));
    }
};

// EnergyRigidBodyConstraints.cpp:64-80 + RigidBodyConstraints.h:191-194 ; DoF order: va, vb, wa, wb
struct rb_constraint_points {
    static constexpr int N_IN = 35, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {9, 22, 12, 25};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[8];
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 9, in + 12, in + 15, in + 18, dt);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 22, in + 25, in + 28, in + 31, dt);
        return 0.5 * in[6] * norm2(rigid_x1(fb, in + 3) - rigid_x1(fa, in + 0));
    }
};

// EnergyRigidBodyConstraints.cpp:82-99 + RigidBodyConstraints.h:228-231 ; DoF order: va, vb, wa, wb
struct rb_constraint_point_on_axis {
    static constexpr int N_IN = 38, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {12, 25, 15, 28};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[11];
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 12, in + 15, in + 18, in + 21, dt);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 25, in + 28, in + 31, in + 34, dt);
        const V3<T> a = rigid_x1(fa, in + 0);
        const V3<T> da = mul(fa.R, ld3(in + 3));
        const V3<T> b = rigid_x1(fb, in + 6);
        const V3<T> a2 = a + da;
        const V3<T> ab = a2 - a, ap = b - a;
        const T e = dot(ap, ab);
        return 0.5 * in[9] * (dot(ap, ap) - e * e / dot(ab, ab));
    }
};

// EnergyRigidBodyConstraints.cpp:101-118 + RigidBodyConstraints.h:265-268
struct rb_constraint_distances {
    static constexpr int N_IN = 36, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {10, 23, 13, 26};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[9];
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 10, in + 13, in + 16, in + 19, dt);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 23, in + 26, in + 29, in + 32, dt);
        return 0.5 * in[7] * sq(in[6] - norm(rigid_x1(fb, in + 3) - rigid_x1(fa, in + 0)));
    }
};

// EnergyRigidBodyConstraints.cpp:120-138 + RigidBodyConstraints.h:305-311
struct rb_constraint_distance_limits {
    static constexpr int N_IN = 37, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {11, 24, 14, 27};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[10], dmin = in[6], dmax = in[7], k = in[8];
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 11, in + 14, in + 17, in + 20, dt);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 24, in + 27, in + 30, in + 33, dt);
        const T length = norm(rigid_x1(fb, in + 3) - rigid_x1(fa, in + 0));
        T E(0.0);
        if (dmin - val(length) > 0.0) E += k * sq(dmin - length) / 2.0;
        if (val(length) - dmax > 0.0) E += k * sq(length - dmax) / 2.0;
        return E;
    }
};

// EnergyRigidBodyConstraints.cpp:140-156 + RigidBodyConstraints.h:357-360 ; DoF order: wa, wb
struct rb_constraint_directions {
    static constexpr int N_IN = 23, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {9, 16};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[8];
        const M3<T> Ra = rotation_q1(s.dof3(0, in + 9), in + 12, dt);
        const M3<T> Rb = rotation_q1(s.dof3(3, in + 16), in + 19, dt);
        return 0.5 * in[6] * norm2(mul(Rb, ld3(in + 3)) - mul(Ra, ld3(in + 0)));
    }
};

// EnergyRigidBodyConstraints.cpp:158-175 + RigidBodyConstraints.h:409-413
struct rb_constraint_angle_limits {
    static constexpr int N_IN = 24, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {10, 17};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[9], dmax = in[6], k = in[7];
        const M3<T> Ra = rotation_q1(s.dof3(0, in + 10), in + 13, dt);
        const M3<T> Rb = rotation_q1(s.dof3(3, in + 17), in + 20, dt);
        const T length = norm(mul(Rb, ld3(in + 3)) - mul(Ra, ld3(in + 0)));
        if (val(length) - dmax > 0.0) return k * cube(length - dmax) / 3.0;
        return T(0.0);
    }
};

// EnergyRigidBodyConstraints.cpp:177-196 + RigidBodyConstraints.h:455-466
struct rb_constraint_damped_spring {
    static constexpr int N_IN = 37, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {11, 24, 14, 27};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[10], rest = in[6], stiffness = in[7], damping = in[8];
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 11, in + 14, in + 17, in + 20, dt);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 24, in + 27, in + 30, in + 33, dt);
        // x0 = t0 + R(q0) x_loc
        const double* qa = in + 20; const double* qb = in + 33;
        const M3<double> Ra0 = quat_to_rotation<double>(qa[0], qa[1], qa[2], qa[3]);
        const M3<double> Rb0 = quat_to_rotation<double>(qb[0], qb[1], qb[2], qb[3]);
        const V3<double> a0 = ld3(in + 17) + mul(Ra0, ld3(in + 0));
        const V3<double> b0 = ld3(in + 30) + mul(Rb0, ld3(in + 3));
        const T l1 = norm(rigid_x1(fb, in + 3) - rigid_x1(fa, in + 0));
        const double l0 = norm(b0 - a0);
        return 0.5 * stiffness * sq(l1 - rest) + 0.5 * damping * sq((l1 - l0) / dt);
    }
};

// =====================================================================================================
// frictional contact: one energy per (physical-system pair, primitive pair) as in
// S/models/interactions/EnergyFrictionalContact.cpp:829-1218.  SOFT... helpers read a deformable point's
// [v1 | x0] slots; the rigid side reads [X_loc..., v, w, t0, q0].
// =====================================================================================================

// ---- deformable - deformable contact (EnergyFrictionalContact.cpp:833-898) ----
struct contact_d_d_pt_pp {
    static constexpr int N_IN = 17, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {0, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> p = soft_x1(s, 0, in + 0, in + 3, in[6]);
        const V3<T> q = soft_x1(s, 3, in + 7, in + 10, in[13]);
        return barrier_cubic(distance_point_point(p, q), in[14] + in[15], in[16]);
    }
};
struct contact_d_d_pt_pe {
    static constexpr int N_IN = 23, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {0, 7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> p = soft_x1(s, 0, in + 0, in + 3, in[6]);
        const V3<T> e0 = soft_x1(s, 3, in + 7, in + 13, in[19]), e1 = soft_x1(s, 6, in + 10, in + 16, in[19]);
        return barrier_cubic(distance_point_line(p, e0, e1), in[20] + in[21], in[22]);
    }
};
struct contact_d_d_pt_pt {
    static constexpr int N_IN = 29, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 7, 10, 13};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> p = soft_x1(s, 0, in + 0, in + 3, in[6]);
        const V3<T> t0 = soft_x1(s, 3, in + 7, in + 16, in[25]), t1 = soft_x1(s, 6, in + 10, in + 19, in[25]), t2 = soft_x1(s, 9, in + 13, in + 22, in[25]);
        return barrier_cubic(distance_point_plane(p, t0, t1, t2), in[26] + in[27], in[28]);
    }
};
struct contact_d_d_ee_pp {
    static constexpr int N_IN = 55, N_DOF = 18, NB = 6;
    static constexpr int DOF_SLOT[NB] = {0, 3, 19, 26, 29, 45};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> ea0 = soft_x1(s, 0, in + 0, in + 6, in[12]), ea1 = soft_x1(s, 3, in + 3, in + 9, in[12]);
        const V3<T> p = soft_x1(s, 6, in + 19, in + 22, in[25]);
        const V3<T> eb0 = soft_x1(s, 9, in + 26, in + 32, in[38]), eb1 = soft_x1(s, 12, in + 29, in + 35, in[38]);
        const V3<T> q = soft_x1(s, 15, in + 45, in + 48, in[51]);
        const T d = distance_point_point(p, q);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 13, in + 16, in + 39, in + 42) * barrier_cubic(d, in[53] + in[54], in[52]);
    }
};
struct contact_d_d_ee_pe {
    static constexpr int N_IN = 48, N_DOF = 15, NB = 5;
    static constexpr int DOF_SLOT[NB] = {0, 3, 19, 26, 29};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> ea0 = soft_x1(s, 0, in + 0, in + 6, in[12]), ea1 = soft_x1(s, 3, in + 3, in + 9, in[12]);
        const V3<T> p = soft_x1(s, 6, in + 19, in + 22, in[25]);
        const V3<T> eb0 = soft_x1(s, 9, in + 26, in + 32, in[38]), eb1 = soft_x1(s, 12, in + 29, in + 35, in[38]);
        const T d = distance_point_line(p, eb0, eb1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 13, in + 16, in + 39, in + 42) * barrier_cubic(d, in[46] + in[47], in[45]);
    }
};
struct contact_d_d_ee_ee {
    static constexpr int N_IN = 41, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 19, 22};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> ea0 = soft_x1(s, 0, in + 0, in + 6, in[12]), ea1 = soft_x1(s, 3, in + 3, in + 9, in[12]);
        const V3<T> eb0 = soft_x1(s, 6, in + 19, in + 25, in[31]), eb1 = soft_x1(s, 9, in + 22, in + 28, in[31]);
        const T d = distance_line_line(ea0, ea1, eb0, eb1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 13, in + 16, in + 32, in + 35) * barrier_cubic(d, in[39] + in[40], in[38]);
    }
};

// ---- rigid - deformable contact (EnergyFrictionalContact.cpp:974-1072); DoF order: soft blocks, rb v, rb w ----
struct contact_rb_d_pt_pp {
    static constexpr int N_IN = 27, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {17, 4, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const V3<T> p = rigid_x1(f, in + 1);
        const V3<T> q = soft_x1(s, 0, in + 17, in + 20, in[23]);
        return barrier_cubic(distance_point_point(p, q), in[24] + in[25], in[26]);
    }
};
struct contact_rb_d_pt_pe {
    static constexpr int N_IN = 33, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {17, 20, 4, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 6, 9, in + 4, in + 7, in + 10, in + 13, in[0]);
        const V3<T> p = rigid_x1(f, in + 1);
        const V3<T> e0 = soft_x1(s, 0, in + 17, in + 23, in[29]), e1 = soft_x1(s, 3, in + 20, in + 26, in[29]);
        return barrier_cubic(distance_point_line(p, e0, e1), in[30] + in[31], in[32]);
    }
};
struct contact_rb_d_pt_pt {
    static constexpr int N_IN = 39, N_DOF = 15, NB = 5;
    static constexpr int DOF_SLOT[NB] = {17, 20, 23, 4, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 9, 12, in + 4, in + 7, in + 10, in + 13, in[0]);
        const V3<T> p = rigid_x1(f, in + 1);
        const V3<T> t0 = soft_x1(s, 0, in + 17, in + 26, in[35]), t1 = soft_x1(s, 3, in + 20, in + 29, in[35]), t2 = soft_x1(s, 6, in + 23, in + 32, in[35]);
        return barrier_cubic(distance_point_plane(p, t0, t1, t2), in[36] + in[37], in[38]);
    }
};
struct contact_rb_d_pt_ep {
    static constexpr int N_IN = 30, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {20, 7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 7, in + 10, in + 13, in + 16, in[0]);
        const V3<T> e0 = rigid_x1(f, in + 1), e1 = rigid_x1(f, in + 4);
        const V3<T> p = soft_x1(s, 0, in + 20, in + 23, in[26]);
        return barrier_cubic(distance_point_line(p, e0, e1), in[27] + in[28], in[29]);
    }
};
struct contact_rb_d_pt_tp {
    static constexpr int N_IN = 33, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {23, 10, 13};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 10, in + 13, in + 16, in + 19, in[0]);
        const V3<T> t0 = rigid_x1(f, in + 1), t1 = rigid_x1(f, in + 4), t2 = rigid_x1(f, in + 7);
        const V3<T> p = soft_x1(s, 0, in + 23, in + 26, in[29]);
        return barrier_cubic(distance_point_plane(p, t0, t1, t2), in[30] + in[31], in[32]);
    }
};
// The two rigid-side symbol sets of ee_pp / ee_pe (edge and point of the SAME body, two get_x1 calls) are separate
// DoF blocks in the reference (n = 21 / 18): soft eb0, eb1, [q], rb v (edge), rb v (point), rb w (edge), rb w (point).
struct contact_rb_d_ee_pp {
    static constexpr int N_IN = 72, N_DOF = 21, NB = 7;
    static constexpr int DOF_SLOT[NB] = {43, 46, 62, 7, 30, 10, 33};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fe = rigid_frame(s, 9, 15, in + 7, in + 10, in + 13, in + 16, in[0]);
        const RigidFrame<T> fp = rigid_frame(s, 12, 18, in + 30, in + 33, in + 36, in + 39, in[26]);
        const V3<T> ea0 = rigid_x1(fe, in + 1), ea1 = rigid_x1(fe, in + 4);
        const V3<T> p = rigid_x1(fp, in + 27);
        const V3<T> eb0 = soft_x1(s, 0, in + 43, in + 49, in[55]), eb1 = soft_x1(s, 3, in + 46, in + 52, in[55]);
        const V3<T> q = soft_x1(s, 6, in + 62, in + 65, in[68]);
        const T d = distance_point_point(p, q);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 56, in + 59) * barrier_cubic(d, in[70] + in[71], in[69]);
    }
};
struct contact_rb_d_ee_pe {
    static constexpr int N_IN = 65, N_DOF = 18, NB = 6;
    static constexpr int DOF_SLOT[NB] = {43, 46, 7, 30, 10, 33};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fe = rigid_frame(s, 6, 12, in + 7, in + 10, in + 13, in + 16, in[0]);
        const RigidFrame<T> fp = rigid_frame(s, 9, 15, in + 30, in + 33, in + 36, in + 39, in[26]);
        const V3<T> ea0 = rigid_x1(fe, in + 1), ea1 = rigid_x1(fe, in + 4);
        const V3<T> p = rigid_x1(fp, in + 27);
        const V3<T> eb0 = soft_x1(s, 0, in + 43, in + 49, in[55]), eb1 = soft_x1(s, 3, in + 46, in + 52, in[55]);
        const T d = distance_point_line(p, eb0, eb1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 56, in + 59) * barrier_cubic(d, in[63] + in[64], in[62]);
    }
};
struct contact_rb_d_ee_ee {
    static constexpr int N_IN = 48, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {26, 29, 7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 6, 9, in + 7, in + 10, in + 13, in + 16, in[0]);
        const V3<T> ea0 = rigid_x1(f, in + 1), ea1 = rigid_x1(f, in + 4);
        const V3<T> eb0 = soft_x1(s, 0, in + 26, in + 32, in[38]), eb1 = soft_x1(s, 3, in + 29, in + 35, in[38]);
        const T d = distance_line_line(ea0, ea1, eb0, eb1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 39, in + 42) * barrier_cubic(d, in[46] + in[47], in[45]);
    }
};
struct contact_rb_d_ee_ep {
    static constexpr int N_IN = 55, N_DOF = 15, NB = 5;
    static constexpr int DOF_SLOT[NB] = {26, 29, 45, 7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 9, 12, in + 7, in + 10, in + 13, in + 16, in[0]);
        const V3<T> ea0 = rigid_x1(f, in + 1), ea1 = rigid_x1(f, in + 4);
        const V3<T> eb0 = soft_x1(s, 0, in + 26, in + 32, in[38]), eb1 = soft_x1(s, 3, in + 29, in + 35, in[38]);
        const V3<T> q = soft_x1(s, 6, in + 45, in + 48, in[51]);
        const T d = distance_point_line(q, ea0, ea1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 39, in + 42) * barrier_cubic(d, in[53] + in[54], in[52]);
    }
};

// ---- deformable - deformable friction (EnergyFrictionalContact.cpp:1078-1118) ----
struct friction_d_d_pp {
    static constexpr int N_IN = 16, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {0, 3};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> v = s.dof3(3, in + 3) - s.dof3(0, in + 0);
        return friction_C0(v, in + 6, in[12], in[13], in[14], in[15]);
    }
};
struct friction_d_d_pe {
    static constexpr int N_IN = 21, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> vb = in[9] * s.dof3(3, in + 3) + in[10] * s.dof3(6, in + 6);
        const V3<T> v = vb - s.dof3(0, in + 0);
        return friction_C0(v, in + 11, in[17], in[18], in[19], in[20]);
    }
};
struct friction_d_d_pt {
    static constexpr int N_IN = 25, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> vb = in[12] * s.dof3(3, in + 3) + in[13] * s.dof3(6, in + 6) + in[14] * s.dof3(9, in + 9);
        const V3<T> v = vb - s.dof3(0, in + 0);
        return friction_C0(v, in + 15, in[21], in[22], in[23], in[24]);
    }
};
struct friction_d_d_ee {
    static constexpr int N_IN = 24, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const V3<T> a0 = s.dof3(0, in + 0), a1 = s.dof3(3, in + 3), b0 = s.dof3(6, in + 6), b1 = s.dof3(9, in + 9);
        const V3<T> va = a0 + in[12] * (a1 - a0);
        const V3<T> vb = b0 + in[13] * (b1 - b0);
        return friction_C0(vb - va, in + 14, in[20], in[21], in[22], in[23]);
    }
};

// ---- rigid - deformable friction (EnergyFrictionalContact.cpp:1160-1218) ----
struct friction_rb_d_pp {
    static constexpr int N_IN = 30, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {17, 4, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const V3<T> v = s.dof3(0, in + 17) - rigid_v1(f, in + 1);
        return friction_C0(v, in + 20, in[26], in[27], in[28], in[29]);
    }
};
struct friction_rb_d_pe {
    static constexpr int N_IN = 35, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {17, 20, 4, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 6, 9, in + 4, in + 7, in + 10, in + 13, in[0]);
        const V3<T> vb = in[23] * s.dof3(0, in + 17) + in[24] * s.dof3(3, in + 20);
        return friction_C0(vb - rigid_v1(f, in + 1), in + 25, in[31], in[32], in[33], in[34]);
    }
};
struct friction_rb_d_pt {
    static constexpr int N_IN = 39, N_DOF = 15, NB = 5;
    static constexpr int DOF_SLOT[NB] = {17, 20, 23, 4, 7};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 9, 12, in + 4, in + 7, in + 10, in + 13, in[0]);
        const V3<T> vb = in[26] * s.dof3(0, in + 17) + in[27] * s.dof3(3, in + 20) + in[28] * s.dof3(6, in + 23);
        return friction_C0(vb - rigid_v1(f, in + 1), in + 29, in[35], in[36], in[37], in[38]);
    }
};
struct friction_rb_d_ee {
    static constexpr int N_IN = 38, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {20, 23, 7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 6, 9, in + 7, in + 10, in + 13, in + 16, in[0]);
        const V3<T> a0 = rigid_v1(f, in + 1), a1 = rigid_v1(f, in + 4);
        const V3<T> b0 = s.dof3(0, in + 20), b1 = s.dof3(3, in + 23);
        const V3<T> va = a0 + in[26] * (a1 - a0);
        const V3<T> vb = b0 + in[27] * (b1 - b0);
        return friction_C0(vb - va, in + 28, in[34], in[35], in[36], in[37]);
    }
};
struct friction_rb_d_ep {
    static constexpr int N_IN = 35, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {20, 7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 7, in + 10, in + 13, in + 16, in[0]);
        const V3<T> vb = in[23] * rigid_v1(f, in + 1) + in[24] * rigid_v1(f, in + 4);
        return friction_C0(vb - s.dof3(0, in + 20), in + 25, in[31], in[32], in[33], in[34]);
    }
};
struct friction_rb_d_tp {
    static constexpr int N_IN = 39, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {23, 10, 13};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 10, in + 13, in + 16, in + 19, in[0]);
        const V3<T> vb = in[26] * rigid_v1(f, in + 1) + in[27] * rigid_v1(f, in + 4) + in[28] * rigid_v1(f, in + 7);
        return friction_C0(vb - s.dof3(0, in + 23), in + 29, in[35], in[36], in[37], in[38]);
    }
};


// ---- velocity controllers (EnergyRigidBodyConstraints.cpp:198-239 + RigidBodyConstraints.h:54-69 c1_controller_energy) ----
template<class T> SB_HD T c1_controller(const V3<T>& da1, const V3<T>& va1, const V3<T>& vb1, double target, double max_force, double delay, double dt)
{
    const T v = dot(da1, vb1 - va1);
    const double k = max_force / delay;
    const double eps = delay / 2.0;
    const T dv = v - target;
    if (val(dv) < -delay) return -(max_force * (dv - eps) * dt);
    if (val(dv) < delay) return 0.5 * k * sq(dv) * dt;
    return max_force * (dv - eps) * dt;
}
// DoF order: va, vb, wa
struct rb_constraint_linear_velocity {
    static constexpr int N_IN = 21, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {7, 10, 13};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[20];
        const M3<T> Ra = rotation_q1(s.dof3(6, in + 13), in + 16, dt);
        const V3<T> da1 = mul(Ra, ld3(in + 0));
        return c1_controller(da1, s.dof3(0, in + 7), s.dof3(3, in + 10), in[3], in[4], in[5], dt);
    }
};
// DoF order: wa, wb
struct rb_constraint_angular_velocity {
    static constexpr int N_IN = 18, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {7, 10};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[17];
        const V3<T> wa = s.dof3(0, in + 7), wb = s.dof3(3, in + 10);
        const M3<T> Ra = rotation_q1(wa, in + 13, dt);
        const V3<T> da1 = mul(Ra, ld3(in + 0));
        return c1_controller(da1, wa, wb, in[3], in[4], in[5], dt);
    }
};

// =====================================================================================================
// attachments (S/models/interactions/EnergyAttachments.cpp:17-135): 0.5 k |q - p|^2 between (barycentric) points
// =====================================================================================================
struct EnergyAttachments_d_d_p_p {
    static constexpr int N_IN = 14, N_DOF = 6, NB = 2;
    static constexpr int DOF_SLOT[NB] = {0, 3};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[13];
        const V3<T> a = soft_x1(s, 0, in + 0, in + 6, dt), b = soft_x1(s, 3, in + 3, in + 9, dt);
        return 0.5 * in[12] * norm2(b - a);
    }
};
struct EnergyAttachments_d_d_p_e {
    static constexpr int N_IN = 22, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[21];
        const V3<T> p = soft_x1(s, 0, in + 0, in + 9, dt);
        const V3<T> q = in[18] * soft_x1(s, 3, in + 3, in + 12, dt) + in[19] * soft_x1(s, 6, in + 6, in + 15, dt);
        return 0.5 * in[20] * norm2(q - p);
    }
};
struct EnergyAttachments_d_d_p_t {
    static constexpr int N_IN = 29, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[28];
        const V3<T> p = soft_x1(s, 0, in + 0, in + 12, dt);
        const V3<T> q = in[24] * soft_x1(s, 3, in + 3, in + 15, dt) + in[25] * soft_x1(s, 6, in + 6, in + 18, dt) + in[26] * soft_x1(s, 9, in + 9, in + 21, dt);
        return 0.5 * in[27] * norm2(q - p);
    }
};
struct EnergyAttachments_d_d_e_e {
    static constexpr int N_IN = 30, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {0, 3, 6, 9};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[29];
        const V3<T> p = in[24] * soft_x1(s, 0, in + 0, in + 12, dt) + in[25] * soft_x1(s, 3, in + 3, in + 15, dt);
        const V3<T> q = in[26] * soft_x1(s, 6, in + 6, in + 18, dt) + in[27] * soft_x1(s, 9, in + 9, in + 21, dt);
        return 0.5 * in[28] * norm2(q - p);
    }
};
// DoF order: soft point, rb v, rb w
struct EnergyAttachments_rb_d {
    static constexpr int N_IN = 24, N_DOF = 9, NB = 3;
    static constexpr int DOF_SLOT[NB] = {2, 11, 14};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const double dt = in[1];
        const V3<T> xd = soft_x1(s, 0, in + 2, in + 5, dt);
        const RigidFrame<T> f = rigid_frame(s, 3, 6, in + 11, in + 14, in + 17, in + 20, dt);
        return 0.5 * in[0] * norm2(xd - rigid_x1(f, in + 8));
    }
};

// =====================================================================================================
// rigid - rigid contact (EnergyFrictionalContact.cpp:900-968) ; DoF order: v_a, v_b, w_a, w_b (each symbol set of a body
// that the reference creates separately -- edge and edge-point of ee_pp / ee_pe -- is its own block)
// =====================================================================================================
struct contact_rb_rb_pt_pp {
    static constexpr int N_IN = 37, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {4, 21, 7, 24};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 21, in + 24, in + 27, in + 30, in[17]);
        return barrier_cubic(distance_point_point(rigid_x1(fa, in + 1), rigid_x1(fb, in + 18)), in[34] + in[35], in[36]);
    }
};
struct contact_rb_rb_pt_pe {
    static constexpr int N_IN = 40, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {4, 24, 7, 27};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 24, in + 27, in + 30, in + 33, in[17]);
        return barrier_cubic(distance_point_line(rigid_x1(fa, in + 1), rigid_x1(fb, in + 18), rigid_x1(fb, in + 21)), in[37] + in[38], in[39]);
    }
};
struct contact_rb_rb_pt_pt {
    static constexpr int N_IN = 43, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {4, 27, 7, 30};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 27, in + 30, in + 33, in + 36, in[17]);
        return barrier_cubic(distance_point_plane(rigid_x1(fa, in + 1), rigid_x1(fb, in + 18), rigid_x1(fb, in + 21), rigid_x1(fb, in + 24)), in[40] + in[41], in[42]);
    }
};
// blocks: v_a (edge), v_a (point), v_b (edge), v_b (point), w_a (edge), w_a (point), w_b (edge), w_b (point)
struct contact_rb_rb_ee_pp {
    static constexpr int N_IN = 89, N_DOF = 24, NB = 8;
    static constexpr int DOF_SLOT[NB] = {7, 30, 50, 73, 10, 33, 53, 76};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fea = rigid_frame(s, 0, 12, in + 7, in + 10, in + 13, in + 16, in[0]);
        const RigidFrame<T> fpa = rigid_frame(s, 3, 15, in + 30, in + 33, in + 36, in + 39, in[26]);
        const RigidFrame<T> feb = rigid_frame(s, 6, 18, in + 50, in + 53, in + 56, in + 59, in[43]);
        const RigidFrame<T> fpb = rigid_frame(s, 9, 21, in + 73, in + 76, in + 79, in + 82, in[69]);
        const V3<T> ea0 = rigid_x1(fea, in + 1), ea1 = rigid_x1(fea, in + 4), eb0 = rigid_x1(feb, in + 44), eb1 = rigid_x1(feb, in + 47);
        const T d = distance_point_point(rigid_x1(fpa, in + 27), rigid_x1(fpb, in + 70));
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 63, in + 66) * barrier_cubic(d, in[87] + in[88], in[86]);
    }
};
// blocks: v_a (edge), v_a (point), v_b, w_a (edge), w_a (point), w_b
struct contact_rb_rb_ee_pe {
    static constexpr int N_IN = 72, N_DOF = 18, NB = 6;
    static constexpr int DOF_SLOT[NB] = {7, 30, 50, 10, 33, 53};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fea = rigid_frame(s, 0, 9, in + 7, in + 10, in + 13, in + 16, in[0]);
        const RigidFrame<T> fpa = rigid_frame(s, 3, 12, in + 30, in + 33, in + 36, in + 39, in[26]);
        const RigidFrame<T> feb = rigid_frame(s, 6, 15, in + 50, in + 53, in + 56, in + 59, in[43]);
        const V3<T> ea0 = rigid_x1(fea, in + 1), ea1 = rigid_x1(fea, in + 4), eb0 = rigid_x1(feb, in + 44), eb1 = rigid_x1(feb, in + 47);
        const T d = distance_point_line(rigid_x1(fpa, in + 27), eb0, eb1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 63, in + 66) * barrier_cubic(d, in[70] + in[71], in[69]);
    }
};
struct contact_rb_rb_ee_ee {
    static constexpr int N_IN = 55, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {7, 33, 10, 36};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 7, in + 10, in + 13, in + 16, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 33, in + 36, in + 39, in + 42, in[26]);
        const V3<T> ea0 = rigid_x1(fa, in + 1), ea1 = rigid_x1(fa, in + 4), eb0 = rigid_x1(fb, in + 27), eb1 = rigid_x1(fb, in + 30);
        const T d = distance_line_line(ea0, ea1, eb0, eb1);
        return ee_mollifier(ea0, ea1, eb0, eb1, in + 20, in + 23, in + 46, in + 49) * barrier_cubic(d, in[53] + in[54], in[52]);
    }
};

// ---- rigid - rigid friction (EnergyFrictionalContact.cpp:1119-1158) ; DoF order: v_a, v_b, w_a, w_b ----
struct friction_rb_rb_pp {
    static constexpr int N_IN = 44, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {4, 21, 7, 24};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 21, in + 24, in + 27, in + 30, in[17]);
        return friction_C0(rigid_v1(fb, in + 18) - rigid_v1(fa, in + 1), in + 34, in[40], in[41], in[42], in[43]);
    }
};
struct friction_rb_rb_pe {
    static constexpr int N_IN = 49, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {4, 24, 7, 27};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 24, in + 27, in + 30, in + 33, in[17]);
        const V3<T> vb = in[37] * rigid_v1(fb, in + 18) + in[38] * rigid_v1(fb, in + 21);
        return friction_C0(vb - rigid_v1(fa, in + 1), in + 39, in[45], in[46], in[47], in[48]);
    }
};
struct friction_rb_rb_pt {
    static constexpr int N_IN = 53, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {4, 27, 7, 30};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 4, in + 7, in + 10, in + 13, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 27, in + 30, in + 33, in + 36, in[17]);
        const V3<T> vb = in[40] * rigid_v1(fb, in + 18) + in[41] * rigid_v1(fb, in + 21) + in[42] * rigid_v1(fb, in + 24);
        return friction_C0(vb - rigid_v1(fa, in + 1), in + 43, in[49], in[50], in[51], in[52]);
    }
};
struct friction_rb_rb_ee {
    static constexpr int N_IN = 52, N_DOF = 12, NB = 4;
    static constexpr int DOF_SLOT[NB] = {7, 27, 10, 30};
    template<class T> SB_HD static T energy(const double* in, const Seed<T>& s)
    {
        const RigidFrame<T> fa = rigid_frame(s, 0, 6, in + 7, in + 10, in + 13, in + 16, in[0]);
        const RigidFrame<T> fb = rigid_frame(s, 3, 9, in + 27, in + 30, in + 33, in + 36, in[20]);
        const V3<T> a0 = rigid_v1(fa, in + 1), a1 = rigid_v1(fa, in + 4), b0 = rigid_v1(fb, in + 21), b1 = rigid_v1(fb, in + 24);
        const V3<T> va = a0 + in[40] * (a1 - a0);
        const V3<T> vb = b0 + in[41] * (b1 - b0);
        return friction_C0(vb - va, in + 42, in[48], in[49], in[50], in[51]);
    }
};

}  // namespace sbpot
