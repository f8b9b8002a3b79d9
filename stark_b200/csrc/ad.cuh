// Lane-distributed second-order forward-mode automatic differentiation for per-element energies.
//
// The reference differentiates each potential symbolically on the host (symx/symbol/diff.cpp) and JIT-compiles
// one scalar CPU function per potential (symx/compile/Compilation.cpp:381-469).  Here the derivative work is
// spread across the lanes of a warp group instead: an element with n DoFs has n(n+1)/2 distinct Hessian
// entries, and every lane owns exactly one (i, j) pair.  A lane carries, for every intermediate quantity q,
//     q.v = q,   q.gi = dq/du_i,   q.gj = dq/du_j,   q.h = d2q/(du_i du_j)
// so the whole [E | grad | hess] output of the reference (SecondOrderCompiledPotential.cpp:138-181) falls out
// of one templated evaluation of the energy with T = D2.  With T = double the same source is the energy-only
// kernel used by the Armijo line search (evaluate_P, SecondOrderCompiledGlobal.cpp:72-93).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include <utility>

#define SB_HD __host__ __device__ __forceinline__

namespace sbad {

struct D2 {
    double v, gi, gj, h;
    SB_HD D2() {}
    SB_HD D2(double v_) : v(v_), gi(0.0), gj(0.0), h(0.0) {}
    SB_HD D2(double v_, double gi_, double gj_, double h_) : v(v_), gi(gi_), gj(gj_), h(h_) {}
};

// ---- value access (uniform for double and D2) -------------------------------------------------------
SB_HD double val(double a) { return a; }
SB_HD double val(const D2& a) { return a.v; }

// ---- addition / subtraction ---------------------------------------------------------------------------
SB_HD D2 operator+(const D2& a, const D2& b) { return D2(a.v + b.v, a.gi + b.gi, a.gj + b.gj, a.h + b.h); }
SB_HD D2 operator-(const D2& a, const D2& b) { return D2(a.v - b.v, a.gi - b.gi, a.gj - b.gj, a.h - b.h); }
SB_HD D2 operator+(const D2& a, double b) { return D2(a.v + b, a.gi, a.gj, a.h); }
SB_HD D2 operator+(double a, const D2& b) { return D2(a + b.v, b.gi, b.gj, b.h); }
SB_HD D2 operator-(const D2& a, double b) { return D2(a.v - b, a.gi, a.gj, a.h); }
SB_HD D2 operator-(double a, const D2& b) { return D2(a - b.v, -b.gi, -b.gj, -b.h); }
SB_HD D2 operator-(const D2& a) { return D2(-a.v, -a.gi, -a.gj, -a.h); }
SB_HD D2& operator+=(D2& a, const D2& b) { a = a + b; return a; }
SB_HD D2& operator-=(D2& a, const D2& b) { a = a - b; return a; }
SB_HD D2& operator+=(D2& a, double b) { a.v += b; return a; }
SB_HD D2& operator-=(D2& a, double b) { a.v -= b; return a; }

// ---- multiplication -------------------------------------------------------------------------------------
SB_HD D2 operator*(const D2& a, const D2& b)
{
    return D2(a.v * b.v,
              a.gi * b.v + a.v * b.gi,
              a.gj * b.v + a.v * b.gj,
              a.h * b.v + a.gi * b.gj + a.gj * b.gi + a.v * b.h);
}
SB_HD D2 operator*(const D2& a, double b) { return D2(a.v * b, a.gi * b, a.gj * b, a.h * b); }
SB_HD D2 operator*(double a, const D2& b) { return D2(a * b.v, a * b.gi, a * b.gj, a * b.h); }
SB_HD D2& operator*=(D2& a, const D2& b) { a = a * b; return a; }
SB_HD D2& operator*=(D2& a, double b) { a = a * b; return a; }

// ---- generic unary chain rule:  f(a) with f' = d1, f'' = d2 ----------------------------------------------
SB_HD D2 chain(const D2& a, double f, double d1, double d2)
{
    return D2(f, d1 * a.gi, d1 * a.gj, d2 * a.gi * a.gj + d1 * a.h);
}

SB_HD D2 inv(const D2& a)
{
    const double r = 1.0 / a.v;
    return chain(a, r, -r * r, 2.0 * r * r * r);
}
SB_HD double inv(double a) { return 1.0 / a; }

SB_HD D2 operator/(const D2& a, const D2& b) { return a * inv(b); }
SB_HD D2 operator/(const D2& a, double b) { return a * (1.0 / b); }
SB_HD D2 operator/(double a, const D2& b) { return a * inv(b); }
SB_HD D2& operator/=(D2& a, const D2& b) { a = a / b; return a; }
SB_HD D2& operator/=(D2& a, double b) { a = a / b; return a; }

SB_HD D2 Sqrt(const D2& a)
{
    const double s = ::sqrt(a.v);
    const double d1 = 0.5 / s;
    return chain(a, s, d1, -0.5 * d1 / a.v);
}
SB_HD double Sqrt(double a) { return ::sqrt(a); }

SB_HD D2 Log(const D2& a)
{
    const double r = 1.0 / a.v;
    return chain(a, ::log(a.v), r, -r * r);
}
SB_HD double Log(double a) { return ::log(a); }

SB_HD D2 Acos(const D2& a)
{
    const double s2 = 1.0 - a.v * a.v;
    const double rs = 1.0 / ::sqrt(s2);
    return chain(a, ::acos(a.v), -rs, -a.v * rs / s2);
}
SB_HD double Acos(double a) { return ::acos(a); }

SB_HD D2 sq(const D2& a) { return D2(a.v * a.v, 2.0 * a.v * a.gi, 2.0 * a.v * a.gj, 2.0 * (a.gi * a.gj + a.v * a.h)); }
SB_HD double sq(double a) { return a * a; }

SB_HD D2 cube(const D2& a)
{
    const double a2 = a.v * a.v;
    return chain(a, a2 * a.v, 3.0 * a2, 6.0 * a.v);
}
SB_HD double cube(double a) { return a * a * a; }

// ---- select on the sign of a condition value (symx::branch, symbol/Scalar.h:119-130: `if (cond > 0.0)`) ----
template<class T> SB_HD T select_pos(double cond, const T& pos, const T& neg) { return (cond > 0.0) ? pos : neg; }

// ---- result type of mixed arithmetic ------------------------------------------------------------------
template<class A, class B> using Mix = typename std::conditional<std::is_same<A, D2>::value || std::is_same<B, D2>::value, D2, double>::type;

// ---- tiny fixed-size vector / matrix helpers -------------------------------------------------------------
template<class T> struct V3 {
    T x, y, z;
    SB_HD V3() {}
    SB_HD V3(const T& x_, const T& y_, const T& z_) : x(x_), y(y_), z(z_) {}
    template<class U> SB_HD V3(const V3<U>& o) : x(o.x), y(o.y), z(o.z) {}
    SB_HD T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    SB_HD const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template<class A, class B> SB_HD V3<Mix<A, B>> operator+(const V3<A>& a, const V3<B>& b) { return V3<Mix<A, B>>(a.x + b.x, a.y + b.y, a.z + b.z); }
template<class A, class B> SB_HD V3<Mix<A, B>> operator-(const V3<A>& a, const V3<B>& b) { return V3<Mix<A, B>>(a.x - b.x, a.y - b.y, a.z - b.z); }
template<class A> SB_HD V3<A> operator-(const V3<A>& a) { return V3<A>(-a.x, -a.y, -a.z); }
template<class A> SB_HD V3<A> operator*(const V3<A>& a, double s) { return V3<A>(a.x * s, a.y * s, a.z * s); }
template<class A> SB_HD V3<A> operator*(double s, const V3<A>& a) { return V3<A>(a.x * s, a.y * s, a.z * s); }
template<class A> SB_HD V3<D2> operator*(const V3<A>& a, const D2& s) { return V3<D2>(a.x * s, a.y * s, a.z * s); }
template<class A> SB_HD V3<D2> operator*(const D2& s, const V3<A>& a) { return V3<D2>(a.x * s, a.y * s, a.z * s); }
template<class A, class B> SB_HD Mix<A, B> dot(const V3<A>& a, const V3<B>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template<class A, class B> SB_HD V3<Mix<A, B>> cross(const V3<A>& a, const V3<B>& b)
{
    return V3<Mix<A, B>>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template<class A> SB_HD A norm2(const V3<A>& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
template<class A> SB_HD A norm(const V3<A>& a) { return Sqrt(norm2(a)); }
template<class A> SB_HD V3<A> normalized(const V3<A>& a)
{
    const A r = inv(norm(a));
    return V3<A>(a.x * r, a.y * r, a.z * r);
}
SB_HD V3<double> ld3(const double* p) { return V3<double>(p[0], p[1], p[2]); }

// row-major 3x3
template<class T> struct M3 {
    T m[9];
    SB_HD T& operator()(int r, int c) { return m[3 * r + c]; }
    SB_HD const T& operator()(int r, int c) const { return m[3 * r + c]; }
};
template<class A, class B> SB_HD V3<Mix<A, B>> mul(const M3<A>& M, const V3<B>& v)
{
    return V3<Mix<A, B>>(M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z,
                         M.m[3] * v.x + M.m[4] * v.y + M.m[5] * v.z,
                         M.m[6] * v.x + M.m[7] * v.y + M.m[8] * v.z);
}
template<class A> SB_HD A det3(const M3<A>& M)
{
    return M.m[0] * (M.m[4] * M.m[8] - M.m[5] * M.m[7])
         - M.m[1] * (M.m[3] * M.m[8] - M.m[5] * M.m[6])
         + M.m[2] * (M.m[3] * M.m[7] - M.m[4] * M.m[6]);
}
SB_HD M3<double> inv3(const M3<double>& M, double& det_out)
{
    const double d = det3(M);
    det_out = d;
    const double r = 1.0 / d;
    M3<double> I;
    I.m[0] = (M.m[4] * M.m[8] - M.m[5] * M.m[7]) * r;
    I.m[1] = (M.m[2] * M.m[7] - M.m[1] * M.m[8]) * r;
    I.m[2] = (M.m[1] * M.m[5] - M.m[2] * M.m[4]) * r;
    I.m[3] = (M.m[5] * M.m[6] - M.m[3] * M.m[8]) * r;
    I.m[4] = (M.m[0] * M.m[8] - M.m[2] * M.m[6]) * r;
    I.m[5] = (M.m[2] * M.m[3] - M.m[0] * M.m[5]) * r;
    I.m[6] = (M.m[3] * M.m[7] - M.m[4] * M.m[6]) * r;
    I.m[7] = (M.m[1] * M.m[6] - M.m[0] * M.m[7]) * r;
    I.m[8] = (M.m[0] * M.m[4] - M.m[1] * M.m[3]) * r;
    return I;
}

// ---- DoF seeding -----------------------------------------------------------------------------------------
// Seed<double>: plain value.  Seed<D2>: lane (i, j) marks the k-th element DoF.
template<class T> struct Seed;
template<> struct Seed<double> {
    SB_HD double dof(int, double v) const { return v; }
    SB_HD V3<double> dof3(int, const double* p) const { return V3<double>(p[0], p[1], p[2]); }
};
template<> struct Seed<D2> {
    int i, j;
    SB_HD D2 dof(int k, double v) const { return D2(v, (k == i) ? 1.0 : 0.0, (k == j) ? 1.0 : 0.0, 0.0); }
    // three consecutive DoFs starting at element DoF index k0
    SB_HD V3<D2> dof3(int k0, const double* p) const { return V3<D2>(dof(k0, p[0]), dof(k0 + 1, p[1]), dof(k0 + 2, p[2])); }
};

}  // namespace sbad
