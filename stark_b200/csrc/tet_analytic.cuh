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
// Mapping: 16 lanes per element, lane = block (a, b); one warp = two consecutive elements, whose 2 x 1152 B of Hessian are
// staged in shared memory and written back as nine fully coalesced 256 B stores.  Inputs are gathered once per CTA into
// shared memory through the generic fetch table (same binding contract as every other potential).
#pragma once

namespace sb {

constexpr int TET_THREADS = 128;
constexpr int TET_ELEMS = TET_THREADS / 16;

template<bool COMPLETE>
__global__ void __launch_bounds__(TET_THREADS) k_tet_analytic(const EvalArgs a)
{
    constexpr int NIN = COMPLETE ? 43 : 40;
    __shared__ double s_in[TET_ELEMS * NIN];
    __shared__ double s_H[TET_ELEMS * 144];
    const int tid = threadIdx.x;
    const int e_base = blockIdx.x * TET_ELEMS;
    for (int idx = tid; idx < TET_ELEMS * NIN; idx += TET_THREADS) {
        const int el = idx / NIN, slot = idx - el * NIN;
        const int e = e_base + el;
        if (e < a.n_elem) {
            const FetchSlot fs = a.slots[slot];
            const int row = (fs.conn_col >= 0) ? a.conn[(size_t)e * a.conn_stride + fs.conn_col] : 0;
            s_in[idx] = fs.base[(size_t)row * fs.stride + fs.off];
        }
    }
    __syncthreads();

    const int el = tid >> 4, l = tid & 15;
    const int ba = l >> 2, bb = l & 3;
    const int e = e_base + el;
    const bool live = e < a.n_elem;
    if (live) {
        const double* in = s_in + el * NIN;
        const double dt = in[NIN - 1], scale = in[36], ym = in[37], nu = in[38];
        // rest shape: B = DX^-1, vol = det(DX) / 6, G = shape-function gradients (4 x 3)
        double DX[9];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) DX[3 * r + c] = scale * in[24 + 3 * (c + 1) + r] - scale * in[24 + r];
        const double c00 = DX[4] * DX[8] - DX[5] * DX[7], c01 = DX[5] * DX[6] - DX[3] * DX[8], c02 = DX[3] * DX[7] - DX[4] * DX[6];
        const double det = DX[0] * c00 + DX[1] * c01 + DX[2] * c02;
        const double rd = 1.0 / det;
        double B[9];
        B[0] = c00 * rd; B[1] = (DX[2] * DX[7] - DX[1] * DX[8]) * rd; B[2] = (DX[1] * DX[5] - DX[2] * DX[4]) * rd;
        B[3] = c01 * rd; B[4] = (DX[0] * DX[8] - DX[2] * DX[6]) * rd; B[5] = (DX[2] * DX[3] - DX[0] * DX[5]) * rd;
        B[6] = c02 * rd; B[7] = (DX[1] * DX[6] - DX[0] * DX[7]) * rd; B[8] = (DX[0] * DX[4] - DX[1] * DX[3]) * rd;
        const double vol = det / 6.0;
        double G[12];
        for (int c = 0; c < 3; c++) {
            G[c] = -(B[c] + B[3 + c] + B[6 + c]);
            G[3 + c] = B[c]; G[6 + c] = B[3 + c]; G[9 + c] = B[6 + c];
        }
        // F = sum_a x_a (x) g_a
        double F[9];
        for (int k = 0; k < 9; k++) F[k] = 0.0;
        for (int n = 0; n < 4; n++)
            for (int r = 0; r < 3; r++) {
                const double x = in[12 + 3 * n + r] + dt * in[3 * n + r];
                for (int c = 0; c < 3; c++) F[3 * r + c] += x * G[3 * n + c];
            }
        double Ic = 0.0;
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
        double cFF = 2.0 * mu_ * ri * ri;     // coefficient of (F (x) F)
        const double c3 = lambda_;
        const double c4 = lambda_ * (J - alpha);
        double Psi = 0.5 * mu_ * (Ic - 3.0) + 0.5 * lambda_ * (J - alpha) * (J - alpha) - 0.5 * mu_ * log(Ic + 1.0);

        // second Piola-Kirchhoff-like stress S = dPsi/dE of the E-based terms (symmetric: 00 01 02 11 12 22) and coefficients
        double S[6] = {0, 0, 0, 0, 0, 0};
        double kappa = 0.0, cN = 0.0, cD = 0.0, dn_inv = 0.0;
        double FD[9];           // F dev(E)
        bool limit_active = false;
        double FFt[6] = {0, 0, 0, 0, 0, 0};
        if (COMPLETE) {
            const double limit = in[39], k_sl = in[40], damping = in[41];
            double E1[6];
            {
                int q = 0;
                for (int i = 0; i < 3; i++)
                    for (int j = i; j < 3; j++) {
                        E1[q] = 0.5 * (F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j] - (i == j ? 1.0 : 0.0));
                        FFt[q] = F[3 * i] * F[3 * j] + F[3 * i + 1] * F[3 * j + 1] + F[3 * i + 2] * F[3 * j + 2];
                        q++;
                    }
            }
            if (damping != 0.0) {
                double F0[9];
                for (int k = 0; k < 9; k++) F0[k] = 0.0;
                for (int n = 0; n < 4; n++)
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++) F0[3 * r + c] += in[12 + 3 * n + r] * G[3 * n + c];
                const double k2 = damping / (dt * dt);
                int q = 0;
                double acc = 0.0;
                for (int i = 0; i < 3; i++)
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
            if (dl > 0.0) {
                limit_active = true;
                dn_inv = 1.0 / dn;
                Psi += k_sl * dl * dl * dl / 3.0;
                const double sN = k_sl * dl * dl;
                // N = I/3 + sqrt(2/3) dev/|dev|
                S[0] += sN * (1.0 / 3.0 + s23 * dv[0] * dn_inv); S[1] += sN * s23 * dv[1] * dn_inv; S[2] += sN * s23 * dv[2] * dn_inv;
                S[3] += sN * (1.0 / 3.0 + s23 * dv[3] * dn_inv); S[4] += sN * s23 * dv[4] * dn_inv; S[5] += sN * (1.0 / 3.0 + s23 * dv[5] * dn_inv);
                const double beta = sN * s23;
                kappa += beta * dn_inv;
                cFF -= beta * dn_inv / 3.0;
                cN = 2.0 * k_sl * dl;
                cD = -beta * dn_inv * dn_inv * dn_inv;
                // F dev
                const double dm[9] = {dv[0], dv[1], dv[2], dv[1], dv[3], dv[4], dv[2], dv[4], dv[5]};
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) FD[3 * r + c] = F[3 * r] * dm[c] + F[3 * r + 1] * dm[3 + c] + F[3 * r + 2] * dm[6 + c];
            }
        }

        // ---- block (ba, bb) ----
        const double* ga = G + 3 * ba;
        const double* gb = G + 3 * bb;
        const double gg = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
        double wa[3], wb[3], ca[3], cb[3];
        for (int r = 0; r < 3; r++) {
            wa[r] = F[3 * r] * ga[0] + F[3 * r + 1] * ga[1] + F[3 * r + 2] * ga[2];
            wb[r] = F[3 * r] * gb[0] + F[3 * r + 1] * gb[1] + F[3 * r + 2] * gb[2];
            ca[r] = Cf[3 * r] * ga[0] + Cf[3 * r + 1] * ga[1] + Cf[3 * r + 2] * ga[2];
            cb[r] = Cf[3 * r] * gb[0] + Cf[3 * r + 1] * gb[1] + Cf[3 * r + 2] * gb[2];
        }
        const double xg[3] = {ga[1] * gb[2] - ga[2] * gb[1], ga[2] * gb[0] - ga[0] * gb[2], ga[0] * gb[1] - ga[1] * gb[0]};
        double q[3];
        for (int r = 0; r < 3; r++) q[r] = F[3 * r] * xg[0] + F[3 * r + 1] * xg[1] + F[3 * r + 2] * xg[2];
        double iso = c1 * gg;
        double blk[9];
        if (COMPLETE) {
            const double Sg[3] = {S[0] * gb[0] + S[1] * gb[1] + S[2] * gb[2], S[1] * gb[0] + S[3] * gb[1] + S[4] * gb[2], S[2] * gb[0] + S[4] * gb[1] + S[5] * gb[2]};
            iso += ga[0] * Sg[0] + ga[1] * Sg[1] + ga[2] * Sg[2];
        }
        for (int r = 0; r < 3; r++)
            for (int s = 0; s < 3; s++) blk[3 * r + s] = cFF * wa[r] * wb[s] + c3 * ca[r] * cb[s];
        blk[0] += iso; blk[4] += iso; blk[8] += iso;
        blk[1] += c4 * q[2]; blk[2] -= c4 * q[1]; blk[3] -= c4 * q[2]; blk[5] += c4 * q[0]; blk[6] += c4 * q[1]; blk[7] -= c4 * q[0];
        if (COMPLETE) {
            if (kappa != 0.0) {
                const double hk = 0.5 * kappa;
                const double FFm[9] = {FFt[0], FFt[1], FFt[2], FFt[1], FFt[3], FFt[4], FFt[2], FFt[4], FFt[5]};
                for (int r = 0; r < 3; r++)
                    for (int s = 0; s < 3; s++) blk[3 * r + s] += hk * (FFm[3 * r + s] * gg + wb[r] * wa[s]);
            }
            if (limit_active) {
                const double s23d = sqrt(2.0 / 3.0) * dn_inv;
                double da[3], db[3], na[3], nb[3];
                for (int r = 0; r < 3; r++) {
                    da[r] = FD[3 * r] * ga[0] + FD[3 * r + 1] * ga[1] + FD[3 * r + 2] * ga[2];
                    db[r] = FD[3 * r] * gb[0] + FD[3 * r + 1] * gb[1] + FD[3 * r + 2] * gb[2];
                    na[r] = wa[r] / 3.0 + s23d * da[r];
                    nb[r] = wb[r] / 3.0 + s23d * db[r];
                }
                for (int r = 0; r < 3; r++)
                    for (int s = 0; s < 3; s++) blk[3 * r + s] += cN * na[r] * nb[s] + cD * da[r] * db[s];
            }
        }
        const double hs = vol * dt * dt;
        double* He = s_H + el * 144;
        for (int r = 0; r < 3; r++)
            for (int s = 0; s < 3; s++) He[(3 * ba + r) * 12 + 3 * bb + s] = hs * blk[3 * r + s];

        // ---- gradient (lanes bb == 0), block rows, energy ----
        const int32_t* ce = a.conn + (size_t)e * a.conn_stride;
        if (bb == 0) {
            // P g_a with P = c1 F + c4 C + F S
            double pg[3];
            for (int r = 0; r < 3; r++) pg[r] = c1 * wa[r] + c4 * ca[r];
            if (COMPLETE) {
                const double Sg[3] = {S[0] * ga[0] + S[1] * ga[1] + S[2] * ga[2], S[1] * ga[0] + S[3] * ga[1] + S[4] * ga[2], S[2] * ga[0] + S[4] * ga[1] + S[5] * ga[2]};
                for (int r = 0; r < 3; r++) pg[r] += F[3 * r] * Sg[0] + F[3 * r + 1] * Sg[1] + F[3 * r + 2] * Sg[2];
            }
            const DofBlock b = a.blocks[ba];
            const int node = ce[b.conn_col];
            const double gs = vol * dt;
            for (int r = 0; r < 3; r++) {
                atomicAdd(a.grad + b.dof_offset + 3 * node + r, gs * pg[r]);
                if (a.g_elem) a.g_elem[(size_t)e * 12 + 3 * ba + r] = gs * pg[r];
            }
            a.rows[(size_t)e * 4 + ba] = b.dof_offset / 3 + node;
        }
        if (l == 0) a.E_elem[e] = vol * Psi;
    }
    __syncwarp();
    // coalesced write-back of the warp's two adjacent element Hessians (2 x 144 doubles)
    {
        const int w = tid >> 5, lane = tid & 31;
        const int e0 = e_base + 2 * w;
        const int n_here = min(2, a.n_elem - e0);
        if (n_here > 0) {
            double* dst = a.H + (size_t)e0 * 144;
            const double* src = s_H + (2 * w) * 144;
            for (int k = lane; k < n_here * 144; k += 32) dst[k] = src[k];
        }
    }
}

template<bool COMPLETE> static void launch_tet_analytic_pgh(const EvalArgs& a, cudaStream_t s)
{
    const int grid = (a.n_elem + TET_ELEMS - 1) / TET_ELEMS;
    k_tet_analytic<COMPLETE><<<grid, TET_THREADS, 0, s>>>(a);
}

}  // namespace sb
