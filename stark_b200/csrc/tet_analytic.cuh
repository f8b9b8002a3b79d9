// Hand-derived P + grad + Hessian kernel for EnergyTetStrain / EnergyTetStrain_Elasticity_Only
// (S/models/deformables/volume/EnergyTetStrain.cpp:12-123): stable Neo-Hookean (Smith et al. 2022, eq. 49) + Green-strain
// rate damping + cubic strain limiting on  tr(E)/3 + sqrt(2/3) |dev E|,  times the rest volume.
//
// The reference differentiates the energy symbolically into a ~12.7 k-operation straight-line program per tet; here the
// derivatives are taken in F-space by hand.  With x_a = x0_a + dt v_a, F = sum_a x_a (x) g_a (g_a = shape-function
// gradients, rows of G) and W = vol * Psi(F):
//     dW/dv_a      = dt   * vol * P g_a                         P = dPsi/dF
//     d2W/dv_a dv_b = dt^2 * vol * sum_cd g_ac A_(r c)(s d) g_bd      A = d2Psi/dF2
// and A is a sum of isotropic (delta_rs M_cd), rank-one (u_rc u_sd) and one skew term (the Hessian of det F), so every
// 3x3 block (a, b) of the 12x12 element Hessian costs a handful of 3-vector operations (formulas in DESIGN.md).
//
// Mapping (HBM-bound design: 1,152 of the 1,632 algorithmic bytes per tet are the dense Hessian write; FP64 work is kept
// at ~1.1 k instructions per tet so the FP64 pipe stays far below the store stream):
//   * ONE THREAD PER ELEMENT.  The F-space state (G, F, cof F, S, the per-node vectors F g_n, C g_n, S g_n) is computed
//     exactly once per tet and only the 10 upper-triangular 3x3 blocks are evaluated; the lower ones are their transposes.
//     (A lane-per-block mapping repeats the ~400-instruction set-up in every lane: 4x the FP64 work, which then bounds
//     the kernel instead of HBM.)
//   * one warp = a tile of 32 consecutive elements = 36,864 contiguous bytes of the Hessian store.  Each thread stages its
//     element in shared memory (row pitch 1,168 B = 1,152 + 16 so that the 8 B stores of a half-warp spread over the
//     banks) and sends it to global memory with its OWN bulk asynchronous copy (cp.async.bulk.global.shared::cta, the
//     TMA engine): full 128 B lines, no LSU store traffic, and the copy drains while the warp gathers and sets up its
//     next tile; the buffer is reclaimed with cp.async.bulk.wait_group.read;
//   * the fetch table travels as a __grid_constant__ kernel parameter (constant bank, uniform loads).  When it has STARK's
//     shape (v1 / x0 / X each one array indexed by the same four node columns, which the launcher verifies) the gather is
//     4 node ids + 12 runs of 3 consecutive doubles, all issued back to back, with the NEXT tile's node ids prefetched;
//     any other binding takes the generic per-slot path of the same kernel;
//   * damping and strain-limit terms (zero for most materials / states) are added in a separate, rarely taken pass over
//     the staged element, which keeps the common path's register footprint small;
//   * persistent grid (3 two-warp CTAs per SM, 73 KB of staging each), tiles strided over the grid;
//   * gradient: FP64 atomics into the flat gradient; block rows and energy: direct stores.
#pragma once

namespace sb {

constexpr int TET_WARPS = 2;                     // warps per CTA
constexpr int TET_THREADS = 32 * TET_WARPS;
constexpr int TET_TILE = 32;                     // elements per warp tile (one per lane)
constexpr int TET_PITCH = 146;                   // doubles between two elements in the staging buffer (1,168 B)
constexpr int TET_MAX_NIN = 43;
constexpr int TET_SMEM_BYTES = TET_WARPS * TET_TILE * TET_PITCH * 8;

struct TetParams {
    EvalArgs a;
    FetchSlot slot[TET_MAX_NIN];
};

// is the fetch table STARK's? (three arrays, each indexed by the same four node columns, components contiguous)
static bool tet_layout_is_canonical(const FetchSlot* s)
{
    for (int k = 0; k < 3; k++)
        for (int n = 0; n < 4; n++)
            for (int c = 0; c < 3; c++) {
                const FetchSlot& f = s[12 * k + 3 * n + c];
                const FetchSlot& f0 = s[12 * k];
                if (f.base != f0.base || f.stride != f0.stride || f.off != f0.off + c) return false;
                if (f.conn_col < 0 || f.conn_col != s[3 * n].conn_col) return false;
            }
    return true;
}

template<bool COMPLETE, bool CANON>
__global__ void __launch_bounds__(TET_THREADS) k_tet_analytic(const __grid_constant__ TetParams P)
{
    constexpr int NIN = COMPLETE ? 43 : 40;
    extern __shared__ __align__(128) double s_H[];
    const EvalArgs& a = P.a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sH = s_H + (warp * TET_TILE + lane) * TET_PITCH;          // this thread's element
    const unsigned sH_addr = (unsigned)__cvta_generic_to_shared(sH);
    const int n_tiles = (a.n_elem + TET_TILE - 1) / TET_TILE;
    const int tile_stride = gridDim.x * TET_WARPS;
    const bool bulk_ok = ((reinterpret_cast<unsigned long long>(a.H) & 15ull) == 0ull);
    bool store_pending = false;

    int tile = blockIdx.x * TET_WARPS + warp;
    // node ids of the tile being set up (CANON: prefetched one tile ahead)
    int nid[4] = {0, 0, 0, 0};
    auto load_nodes = [&](int t) {
        const int e = min(t * TET_TILE + lane, a.n_elem - 1);
        const int32_t* ce = a.conn + (size_t)e * a.conn_stride;
#pragma unroll
        for (int n = 0; n < 4; n++) nid[n] = ce[P.slot[3 * n].conn_col];
    };
    if (CANON && tile < n_tiles) load_nodes(tile);

    for (; tile < n_tiles; tile += tile_stride) {
        const int e = tile * TET_TILE + lane;
        const bool live = e < a.n_elem;
        const int32_t* ce = a.conn + (size_t)(live ? e : a.n_elem - 1) * a.conn_stride;   // dead lanes redo the last element
        auto in = [&](int slot) -> double {
            const FetchSlot& fs = P.slot[slot];
            const int row = (fs.conn_col >= 0) ? ce[fs.conn_col] : 0;
            return fs.base[(size_t)row * fs.stride + fs.off];
        };
        // ---- gather ----
        double vv[12], xx[12], XX[12];
        int node[4];
        if (CANON) {
#pragma unroll
            for (int n = 0; n < 4; n++) node[n] = nid[n];
            const double* bv = P.slot[0].base + P.slot[0].off;
            const double* bx = P.slot[12].base + P.slot[12].off;
            const double* bX = P.slot[24].base + P.slot[24].off;
            const int sv = P.slot[0].stride, sx = P.slot[12].stride, sX = P.slot[24].stride;
#pragma unroll
            for (int n = 0; n < 4; n++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    vv[3 * n + c] = bv[(size_t)node[n] * sv + c];
                    xx[3 * n + c] = bx[(size_t)node[n] * sx + c];
                    XX[3 * n + c] = bX[(size_t)node[n] * sX + c];
                }
        } else {
#pragma unroll
            for (int k = 0; k < 12; k++) { vv[k] = in(k); xx[k] = in(12 + k); XX[k] = in(24 + k); }
#pragma unroll
            for (int n = 0; n < 4; n++) node[n] = ce[a.blocks[n].conn_col];
        }
        const double dt = in(NIN - 1), scale = in(36), ym = in(37), nu = in(38);
        if (CANON && tile + tile_stride < n_tiles) load_nodes(tile + tile_stride);   // prefetch the next tile's node ids

        // rest shape: B = DX^-1, vol = det(DX) / 6, G = shape-function gradients (4 x 3)
        double G[12];
        double vol;
        {
#pragma unroll
            for (int k = 0; k < 12; k++) XX[k] = scale * XX[k];
            double DX[9];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < 3; r++) DX[3 * r + c] = XX[3 * (c + 1) + r] - XX[r];
            const double c00 = DX[4] * DX[8] - DX[5] * DX[7], c01 = DX[5] * DX[6] - DX[3] * DX[8], c02 = DX[3] * DX[7] - DX[4] * DX[6];
            const double det = DX[0] * c00 + DX[1] * c01 + DX[2] * c02;
            const double rd = 1.0 / det;
            double B[9];
            B[0] = c00 * rd; B[1] = (DX[2] * DX[7] - DX[1] * DX[8]) * rd; B[2] = (DX[1] * DX[5] - DX[2] * DX[4]) * rd;
            B[3] = c01 * rd; B[4] = (DX[0] * DX[8] - DX[2] * DX[6]) * rd; B[5] = (DX[2] * DX[3] - DX[0] * DX[5]) * rd;
            B[6] = c02 * rd; B[7] = (DX[1] * DX[6] - DX[0] * DX[7]) * rd; B[8] = (DX[0] * DX[4] - DX[1] * DX[3]) * rd;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                G[c] = -(B[c] + B[3 + c] + B[6 + c]);
                G[3 + c] = B[c]; G[6 + c] = B[3 + c]; G[9 + c] = B[6 + c];
            }
            vol = det / 6.0;
        }
        // F = sum_n x_n (x) g_n,  x_n = x0_n + dt v_n
        double F[9];
#pragma unroll
        for (int k = 0; k < 9; k++) F[k] = 0.0;
#pragma unroll
        for (int n = 0; n < 4; n++)
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double x = xx[3 * n + r] + dt * vv[3 * n + r];
#pragma unroll
                for (int c = 0; c < 3; c++) F[3 * r + c] += x * G[3 * n + c];
            }
        double Ic = 0.0;
#pragma unroll
        for (int k = 0; k < 9; k++) Ic += F[k] * F[k];
        double Cf[9];   // cofactor matrix = dJ/dF
        Cf[0] = F[4] * F[8] - F[5] * F[7]; Cf[1] = F[5] * F[6] - F[3] * F[8]; Cf[2] = F[3] * F[7] - F[4] * F[6];
        Cf[3] = F[2] * F[7] - F[1] * F[8]; Cf[4] = F[0] * F[8] - F[2] * F[6]; Cf[5] = F[1] * F[6] - F[0] * F[7];
        Cf[6] = F[1] * F[5] - F[2] * F[4]; Cf[7] = F[2] * F[3] - F[0] * F[5]; Cf[8] = F[0] * F[4] - F[1] * F[3];
        const double J = F[0] * Cf[0] + F[1] * Cf[1] + F[2] * Cf[2];
        const double mu = ym / (2.0 * (1.0 + nu));
        const double lambda = (ym * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu));
        const double mu_ = 4.0 / 3.0 * mu, lambda_ = lambda + 5.0 / 6.0 * mu;
        const double alpha = 1.0 + mu_ / lambda_ - mu_ / (4.0 * lambda_);
        const double ri = 1.0 / (Ic + 1.0);
        const double c1 = mu_ * (1.0 - ri);
        const double cFF = 2.0 * mu_ * ri * ri;     // coefficient of (F (x) F), elastic part
        const double c4 = lambda_ * (J - alpha);
        double Psi = 0.5 * mu_ * (Ic - 3.0) + 0.5 * lambda_ * (J - alpha) * (J - alpha) - 0.5 * mu_ * log(Ic + 1.0);

        // ---- does this element have damping / strain-limit terms?  (cheap test; the terms themselves are the rare pass below) ----
        bool extra = false;
        if (COMPLETE) {
            const double limit = in(39), damping = in(41);
            double E1d[3], E1o[3];
#pragma unroll
            for (int i = 0; i < 3; i++) E1d[i] = 0.5 * (F[i] * F[i] + F[3 + i] * F[3 + i] + F[6 + i] * F[6 + i] - 1.0);
            E1o[0] = 0.5 * (F[0] * F[1] + F[3] * F[4] + F[6] * F[7]);
            E1o[1] = 0.5 * (F[0] * F[2] + F[3] * F[5] + F[6] * F[8]);
            E1o[2] = 0.5 * (F[1] * F[2] + F[4] * F[5] + F[7] * F[8]);
            const double m = (E1d[0] + E1d[1] + E1d[2]) / 3.0;
            const double d0 = E1d[0] - m, d1 = E1d[1] - m, d2 = E1d[2] - m;
            const double dn = sqrt(d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * (E1o[0] * E1o[0] + E1o[1] * E1o[1] + E1o[2] * E1o[2]));
            const double dl = m + sqrt(2.0 / 3.0) * dn - limit;
            extra = (damping != 0.0) || (dl > 0.0);
        }

        // ---- per-node vectors: w_n = F g_n, c_n = C g_n ----
        double w[12], cc[12];
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const double* gn = G + 3 * n;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                w[3 * n + r] = F[3 * r] * gn[0] + F[3 * r + 1] * gn[1] + F[3 * r + 2] * gn[2];
                cc[3 * n + r] = Cf[3 * r] * gn[0] + Cf[3 * r + 1] * gn[1] + Cf[3 * r + 2] * gn[2];
            }
        }

        // ---- elastic gradient dW/dv_n = dt vol P g_n, P g_n = c1 w_n + c4 c_n; block rows ----
        const double gs = vol * dt;
        if (live && !extra) {
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const int off = a.blocks[n].dof_offset;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const double pg = gs * (c1 * w[3 * n + r] + c4 * cc[3 * n + r]);
                    atomicAdd(a.grad + off + 3 * node[n] + r, pg);
                    if (a.g_elem) a.g_elem[(size_t)e * 12 + 3 * n + r] = pg;
                }
            }
        }
        if (live) {
#pragma unroll
            for (int n = 0; n < 4; n++) a.rows[(size_t)e * 4 + n] = a.blocks[n].dof_offset / 3 + node[n];
        }

        // the previous tile's bulk copy must have finished READING this thread's staging slab before it is overwritten
        if (store_pending) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            store_pending = false;
        }

        // ---- elastic Hessian: 10 upper blocks (na <= nb), each stored with its transpose; pre-scaled by vol dt^2 ----
        const double hs = vol * dt * dt;
        {
            const double A1 = hs * cFF, A3 = hs * lambda_, Aiso = hs * c1, A4 = hs * c4;
#pragma unroll
            for (int na = 0; na < 4; na++) {
                const double* ga = G + 3 * na;
                const double a1w[3] = {A1 * w[3 * na], A1 * w[3 * na + 1], A1 * w[3 * na + 2]};
                const double a3c[3] = {A3 * cc[3 * na], A3 * cc[3 * na + 1], A3 * cc[3 * na + 2]};
#pragma unroll
                for (int nb = na; nb < 4; nb++) {
                    const double* gb = G + 3 * nb;
                    const double iso = Aiso * (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]);
                    double blk[9];
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int s = 0; s < 3; s++) blk[3 * r + s] = a1w[r] * w[3 * nb + s] + a3c[r] * cc[3 * nb + s];
                    blk[0] += iso; blk[4] += iso; blk[8] += iso;
                    if (nb != na) {   // A4 * skew(F (g_a x g_b)); vanishes on the diagonal blocks
                        const double xg[3] = {ga[1] * gb[2] - ga[2] * gb[1], ga[2] * gb[0] - ga[0] * gb[2], ga[0] * gb[1] - ga[1] * gb[0]};
                        const double q0 = A4 * (F[0] * xg[0] + F[1] * xg[1] + F[2] * xg[2]);
                        const double q1 = A4 * (F[3] * xg[0] + F[4] * xg[1] + F[5] * xg[2]);
                        const double q2 = A4 * (F[6] * xg[0] + F[7] * xg[1] + F[8] * xg[2]);
                        blk[1] += q2; blk[2] -= q1; blk[3] -= q2; blk[5] += q0; blk[6] += q1; blk[7] -= q0;
                    }
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int s = 0; s < 3; s++) {
                            sH[(3 * na + r) * 12 + 3 * nb + s] = blk[3 * r + s];
                            if (nb != na) sH[(3 * nb + s) * 12 + 3 * na + r] = blk[3 * r + s];
                        }
                }
            }
        }

        // ---- rare pass: damping + strain-limit terms on top of the staged element (EnergyTetStrain.cpp:62-72) ----
        if (COMPLETE && extra) {
            const double limit = in(39), k_sl = in(40), damping = in(41);
            double E1[6], FFt[6], S[6] = {0, 0, 0, 0, 0, 0};
            {
                int q = 0;
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = i; j < 3; j++) {
                        E1[q] = 0.5 * (F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j] - (i == j ? 1.0 : 0.0));
                        FFt[q] = F[3 * i] * F[3 * j] + F[3 * i + 1] * F[3 * j + 1] + F[3 * i + 2] * F[3 * j + 2];
                        q++;
                    }
            }
            double kappa = 0.0;
            if (damping != 0.0) {
                double F0[9];
#pragma unroll
                for (int k = 0; k < 9; k++) F0[k] = 0.0;
#pragma unroll
                for (int n = 0; n < 4; n++)
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int c = 0; c < 3; c++) F0[3 * r + c] += xx[3 * n + r] * G[3 * n + c];
                const double k2 = damping / (dt * dt);
                int q = 0;
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = i; j < 3; j++) {
                        const double e0 = 0.5 * (F0[i] * F0[j] + F0[3 + i] * F0[3 + j] + F0[6 + i] * F0[6 + j] - (i == j ? 1.0 : 0.0));
                        const double D = E1[q] - e0;
                        S[q] += k2 * D;
                        acc += (i == j ? 1.0 : 2.0) * D * D;
                        q++;
                    }
                Psi += 0.5 * k2 * acc;
                kappa += k2;
            }
            const double m = (E1[0] + E1[3] + E1[5]) / 3.0;
            const double dv[6] = {E1[0] - m, E1[1], E1[2], E1[3] - m, E1[4], E1[5] - m};
            const double dn = sqrt(dv[0] * dv[0] + dv[3] * dv[3] + dv[5] * dv[5] + 2.0 * (dv[1] * dv[1] + dv[2] * dv[2] + dv[4] * dv[4]));
            const double s23 = sqrt(2.0 / 3.0);
            const double dl = m + s23 * dn - limit;
            double dcFF = 0.0, cN = 0.0, cD = 0.0, s23d = 0.0;
            double FD[9];
#pragma unroll
            for (int k = 0; k < 9; k++) FD[k] = 0.0;
            if (dl > 0.0) {
                const double dn_inv = 1.0 / dn;
                Psi += k_sl * dl * dl * dl / 3.0;
                const double sN = k_sl * dl * dl;
                // N = I/3 + sqrt(2/3) dev/|dev|
                S[0] += sN * (1.0 / 3.0 + s23 * dv[0] * dn_inv); S[1] += sN * s23 * dv[1] * dn_inv; S[2] += sN * s23 * dv[2] * dn_inv;
                S[3] += sN * (1.0 / 3.0 + s23 * dv[3] * dn_inv); S[4] += sN * s23 * dv[4] * dn_inv; S[5] += sN * (1.0 / 3.0 + s23 * dv[5] * dn_inv);
                const double beta = sN * s23;
                kappa += beta * dn_inv;
                dcFF = -beta * dn_inv / 3.0;
                cN = 2.0 * k_sl * dl;
                cD = -beta * dn_inv * dn_inv * dn_inv;
                s23d = s23 * dn_inv;
                const double dm[9] = {dv[0], dv[1], dv[2], dv[1], dv[3], dv[4], dv[2], dv[4], dv[5]};
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) FD[3 * r + c] = F[3 * r] * dm[c] + F[3 * r + 1] * dm[3 + c] + F[3 * r + 2] * dm[6 + c];
            }
            // per-node vectors t_n = S g_n, d_n = (F dev) g_n, nn_n = (F N) g_n
            double t[12], d[12], nn[12];
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const double* gn = G + 3 * n;
                t[3 * n + 0] = S[0] * gn[0] + S[1] * gn[1] + S[2] * gn[2];
                t[3 * n + 1] = S[1] * gn[0] + S[3] * gn[1] + S[4] * gn[2];
                t[3 * n + 2] = S[2] * gn[0] + S[4] * gn[1] + S[5] * gn[2];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    d[3 * n + r] = FD[3 * r] * gn[0] + FD[3 * r + 1] * gn[1] + FD[3 * r + 2] * gn[2];
                    nn[3 * n + r] = w[3 * n + r] / 3.0 + s23d * d[3 * n + r];
                }
            }
            // gradient with the E-based stress: P g_n = c1 w_n + c4 c_n + F (S g_n)
            if (live) {
#pragma unroll
                for (int n = 0; n < 4; n++) {
                    const int off = a.blocks[n].dof_offset;
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        const double pg = gs * (c1 * w[3 * n + r] + c4 * cc[3 * n + r] + F[3 * r] * t[3 * n] + F[3 * r + 1] * t[3 * n + 1] + F[3 * r + 2] * t[3 * n + 2]);
                        atomicAdd(a.grad + off + 3 * node[n] + r, pg);
                        if (a.g_elem) a.g_elem[(size_t)e * 12 + 3 * n + r] = pg;
                    }
                }
            }
            // Hessian: + hs [ dcFF w_a w_b^T + (g_a . S g_b) I + kappa/2 (F F^T (g_a . g_b) + w_b w_a^T) + cN nn_a nn_b^T + cD d_a d_b^T ]
            const double hk = 0.5 * kappa;
            const double FFm[9] = {FFt[0], FFt[1], FFt[2], FFt[1], FFt[3], FFt[4], FFt[2], FFt[4], FFt[5]};
#pragma unroll 1
            for (int na = 0; na < 4; na++)
#pragma unroll 1
                for (int nb = 0; nb < 4; nb++) {
                    const double gg = G[3 * na] * G[3 * nb] + G[3 * na + 1] * G[3 * nb + 1] + G[3 * na + 2] * G[3 * nb + 2];
                    const double iso = G[3 * na] * t[3 * nb] + G[3 * na + 1] * t[3 * nb + 1] + G[3 * na + 2] * t[3 * nb + 2];
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int s = 0; s < 3; s++) {
                            double v = dcFF * w[3 * na + r] * w[3 * nb + s] + hk * (FFm[3 * r + s] * gg + w[3 * nb + r] * w[3 * na + s])
                                     + cN * nn[3 * na + r] * nn[3 * nb + s] + cD * d[3 * na + r] * d[3 * nb + s];
                            if (r == s) v += iso;
                            sH[(3 * na + r) * 12 + 3 * nb + s] += hs * v;
                        }
                }
        }
        if (live) a.E_elem[e] = vol * Psi;

        // ---- this element's 1,152 B leave through the TMA engine ----
        double* dstH = a.H + (size_t)e * 144;
        if (bulk_ok) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy
            if (live) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dstH), "r"(sH_addr), "r"(144 * 8) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            store_pending = true;
        } else if (live) {
#pragma unroll 1
            for (int k = 0; k < 144; k++) dstH[k] = sH[k];
        }
    }
    if (store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Energy only (line-search evaluations): the same expressions as above up to Psi, one thread per element, nothing staged.
// Reads 16 B of connectivity per tet plus the gathered node data (L2-resident), writes 8 B.
template<bool COMPLETE, bool CANON>
__global__ void __launch_bounds__(128) k_tet_energy(const __grid_constant__ TetParams P)
{
    constexpr int NIN = COMPLETE ? 43 : 40;
    const EvalArgs& a = P.a;
    const int e = blockIdx.x * 128 + threadIdx.x;
    if (e >= a.n_elem) return;
    const int32_t* ce = a.conn + (size_t)e * a.conn_stride;
    auto in = [&](int slot) -> double {
        const FetchSlot& fs = P.slot[slot];
        const int row = (fs.conn_col >= 0) ? ce[fs.conn_col] : 0;
        return fs.base[(size_t)row * fs.stride + fs.off];
    };
    double vv[12], xx[12], XX[12];
    if (CANON) {
        int node[4];
#pragma unroll
        for (int n = 0; n < 4; n++) node[n] = ce[P.slot[3 * n].conn_col];
        const double* bv = P.slot[0].base + P.slot[0].off;
        const double* bx = P.slot[12].base + P.slot[12].off;
        const double* bX = P.slot[24].base + P.slot[24].off;
        const int sv = P.slot[0].stride, sx = P.slot[12].stride, sX = P.slot[24].stride;
#pragma unroll
        for (int n = 0; n < 4; n++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                vv[3 * n + c] = bv[(size_t)node[n] * sv + c];
                xx[3 * n + c] = bx[(size_t)node[n] * sx + c];
                XX[3 * n + c] = bX[(size_t)node[n] * sX + c];
            }
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) { vv[k] = in(k); xx[k] = in(12 + k); XX[k] = in(24 + k); }
    }
    const double dt = in(NIN - 1), scale = in(36), ym = in(37), nu = in(38);
    double G[12];
    double vol;
    {
#pragma unroll
        for (int k = 0; k < 12; k++) XX[k] = scale * XX[k];
        double DX[9];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) DX[3 * r + c] = XX[3 * (c + 1) + r] - XX[r];
        const double c00 = DX[4] * DX[8] - DX[5] * DX[7], c01 = DX[5] * DX[6] - DX[3] * DX[8], c02 = DX[3] * DX[7] - DX[4] * DX[6];
        const double det = DX[0] * c00 + DX[1] * c01 + DX[2] * c02;
        const double rd = 1.0 / det;
        double B[9];
        B[0] = c00 * rd; B[1] = (DX[2] * DX[7] - DX[1] * DX[8]) * rd; B[2] = (DX[1] * DX[5] - DX[2] * DX[4]) * rd;
        B[3] = c01 * rd; B[4] = (DX[0] * DX[8] - DX[2] * DX[6]) * rd; B[5] = (DX[2] * DX[3] - DX[0] * DX[5]) * rd;
        B[6] = c02 * rd; B[7] = (DX[1] * DX[6] - DX[0] * DX[7]) * rd; B[8] = (DX[0] * DX[4] - DX[1] * DX[3]) * rd;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            G[c] = -(B[c] + B[3 + c] + B[6 + c]);
            G[3 + c] = B[c]; G[6 + c] = B[3 + c]; G[9 + c] = B[6 + c];
        }
        vol = det / 6.0;
    }
    double F[9];
#pragma unroll
    for (int k = 0; k < 9; k++) F[k] = 0.0;
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double x = xx[3 * n + r] + dt * vv[3 * n + r];
#pragma unroll
            for (int c = 0; c < 3; c++) F[3 * r + c] += x * G[3 * n + c];
        }
    double Ic = 0.0;
#pragma unroll
    for (int k = 0; k < 9; k++) Ic += F[k] * F[k];
    const double J = F[0] * (F[4] * F[8] - F[5] * F[7]) + F[1] * (F[5] * F[6] - F[3] * F[8]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
    const double mu = ym / (2.0 * (1.0 + nu));
    const double lambda = (ym * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu));
    const double mu_ = 4.0 / 3.0 * mu, lambda_ = lambda + 5.0 / 6.0 * mu;
    const double alpha = 1.0 + mu_ / lambda_ - mu_ / (4.0 * lambda_);
    double Psi = 0.5 * mu_ * (Ic - 3.0) + 0.5 * lambda_ * (J - alpha) * (J - alpha) - 0.5 * mu_ * log(Ic + 1.0);
    if (COMPLETE) {
        const double limit = in(39), k_sl = in(40), damping = in(41);
        double E1[6];
        {
            int q = 0;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = i; j < 3; j++) {
                    E1[q] = 0.5 * (F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j] - (i == j ? 1.0 : 0.0));
                    q++;
                }
        }
        if (damping != 0.0) {
            double F0[9];
#pragma unroll
            for (int k = 0; k < 9; k++) F0[k] = 0.0;
#pragma unroll
            for (int n = 0; n < 4; n++)
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) F0[3 * r + c] += xx[3 * n + r] * G[3 * n + c];
            const double k2 = damping / (dt * dt);
            int q = 0;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = i; j < 3; j++) {
                    const double e0 = 0.5 * (F0[i] * F0[j] + F0[3 + i] * F0[3 + j] + F0[6 + i] * F0[6 + j] - (i == j ? 1.0 : 0.0));
                    const double D = E1[q] - e0;
                    acc += (i == j ? 1.0 : 2.0) * D * D;
                    q++;
                }
            Psi += 0.5 * k2 * acc;
        }
        const double m = (E1[0] + E1[3] + E1[5]) / 3.0;
        const double dv[6] = {E1[0] - m, E1[1], E1[2], E1[3] - m, E1[4], E1[5] - m};
        const double dn = sqrt(dv[0] * dv[0] + dv[3] * dv[3] + dv[5] * dv[5] + 2.0 * (dv[1] * dv[1] + dv[2] * dv[2] + dv[4] * dv[4]));
        const double dl = m + sqrt(2.0 / 3.0) * dn - limit;
        if (dl > 0.0) Psi += k_sl * dl * dl * dl / 3.0;
    }
    a.E_elem[e] = vol * Psi;
}

template<bool COMPLETE> static void launch_tet_analytic_p(const EvalArgs& a, cudaStream_t s)
{
    TetParams P;
    P.a = a;
    const int nin = COMPLETE ? 43 : 40;
    for (int i = 0; i < nin; i++) P.slot[i] = a.slots_host[i];
    const int grid = (a.n_elem + 127) / 128;
    if (tet_layout_is_canonical(P.slot)) k_tet_energy<COMPLETE, true><<<grid, 128, 0, s>>>(P);
    else k_tet_energy<COMPLETE, false><<<grid, 128, 0, s>>>(P);
}

template<bool COMPLETE> static void launch_tet_analytic_pgh(const EvalArgs& a, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_tet_analytic<COMPLETE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TET_SMEM_BYTES);
        cudaFuncSetAttribute(k_tet_analytic<COMPLETE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TET_SMEM_BYTES);
        configured = true;
    }
    TetParams P;
    P.a = a;
    const int nin = COMPLETE ? 43 : 40;
    for (int i = 0; i < nin; i++) P.slot[i] = a.slots_host[i];
    const int n_tiles = (a.n_elem + TET_TILE - 1) / TET_TILE;
    const int ctas_needed = (n_tiles + TET_WARPS - 1) / TET_WARPS;
    static const int env_cap = std::getenv("SB_TET_GRID") ? std::max(1, std::atoi(std::getenv("SB_TET_GRID"))) : 0;   // (experiment hook)
    const int grid_cap = env_cap ? env_cap : g_tet_grid_cap;
    const int grid = ctas_needed < grid_cap ? ctas_needed : grid_cap;   // persistent: 3 CTAs per SM (2 beside the collision detection)
    if (tet_layout_is_canonical(P.slot)) k_tet_analytic<COMPLETE, true><<<grid, TET_THREADS, TET_SMEM_BYTES, s>>>(P);
    else k_tet_analytic<COMPLETE, false><<<grid, TET_THREADS, TET_SMEM_BYTES, s>>>(P);
}

}  // namespace sb
