// Element evaluation kernels: gather -> energy / gradient / Hessian -> scatter.
//
// Replaces CompiledInLoop<double>::run (symx/compile/CompiledInLoop_run.h:66-472) and the per-element callbacks of
// SecondOrderCompiledPotential (symx/solver/second_order/SecondOrderCompiledPotential.cpp:94-181):
//   * gather: every element's in[] is staged ONCE in shared memory by the whole CTA (the reference memcpy-gathers
//     8 elements at a time into a thread buffer, CompiledInLoop_run.h:262-282);
//   * evaluate: n(n+1)/2 lanes per element, one Hessian entry each (ad.cuh);
//   * scatter: dense row-major n x n element Hessian + global block rows (the ElementHessians contract,
//     ElementHessians.cpp:20-41), gradient scatter-add into the flat gradient, per-element energy.
#include "internal.h"
#include "potentials.cuh"
#include <cstring>

namespace sb {

template<class Pot> struct Geo {
    static constexpr int N = Pot::N_DOF;
    static constexpr int L = N * (N + 1) / 2;                       // lanes per element
    static constexpr int BLOCK = (L <= EVAL_THREADS) ? EVAL_THREADS : ((L + 31) / 32) * 32;
    static constexpr int G = BLOCK / L;                              // elements per CTA
};

__device__ __forceinline__ void pair_from_index(int p, int& i, int& j)
{
    int r = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= p) r++;
    while (r * (r + 1) / 2 > p) r--;
    i = r;
    j = p - r * (r + 1) / 2;
}

// body of the P+G+H evaluation of one CTA's elements; BLOCK threads take part (the kernel may have more), G = BLOCK / L elements
template<class Pot, int BLOCK>
__device__ __forceinline__ void eval_pgh_body(const EvalArgs& a, int cta_first, int cta_stride, double* s_in)
{
    constexpr int N = Pot::N_DOF, L = Geo<Pot>::L, G = BLOCK / L, NIN = Pot::N_IN, NB = Pot::NB;
    const int tid = threadIdx.x;
    // element count and output locations: launch arguments, or (fused detection + evaluation) the device-side layout
    int n_elem = a.n_elem;
    double* Hb = a.H; int32_t* rows_b = a.rows; double* Eb = a.E_elem;
    if (a.dyn) { n_elem = a.dyn->n_elem; Hb = a.H + a.dyn->H_off; rows_b = a.rows + a.dyn->rows_off; Eb = a.E_elem + a.dyn->E_off; }
    for (int cta = cta_first; cta * G < n_elem; cta += cta_stride) {   // (one trip unless the launch was sized by an estimate)
        const int e_base = cta * G;
        // cooperative gather of G elements' inputs
        for (int idx = tid; idx < G * NIN; idx += blockDim.x) {
            const int el = idx / NIN, slot = idx - el * NIN;
            const int e = e_base + el;
            if (e < n_elem) {
                const FetchSlot fs = a.slots[slot];
                const int row = (fs.conn_col >= 0) ? a.conn[(size_t)e * a.conn_stride + fs.conn_col] : 0;
                s_in[idx] = fs.base[(size_t)row * fs.stride + fs.off];
            }
        }
        __syncthreads();

        const int el = tid / L;
        const int p = tid - el * L;
        const int e = e_base + el;
        if (el < G && e < n_elem) {
            sbad::Seed<sbad::D2> seed;
            pair_from_index(p, seed.i, seed.j);
            const sbad::D2 r = Pot::template energy<sbad::D2>(s_in + el * NIN, seed);

            const int i = seed.i, j = seed.j;
            double* He = Hb + (size_t)e * N * N;
            He[i * N + j] = r.h;
            if (i != j) He[j * N + i] = r.h;
            const int32_t* ce = a.conn + (size_t)e * a.conn_stride;
            if (j == 0) {
                const DofBlock b = a.blocks[i / 3];
                atomicAdd(a.grad + b.dof_offset + 3 * ce[b.conn_col] + (i % 3), r.gi);
                if (a.g_elem) a.g_elem[(size_t)e * N + i] = r.gi;
            }
            if (p < NB) {
                const DofBlock b = a.blocks[p];
                rows_b[(size_t)e * NB + p] = b.dof_offset / 3 + ce[b.conn_col];
            }
            if (p == 0) Eb[e] = r.v;
        }
        __syncthreads();   // (the staged inputs are overwritten by the next trip)
    }
}
template<class Pot>
__global__ void __launch_bounds__(Geo<Pot>::BLOCK) k_eval_pgh(const EvalArgs a)
{
    __shared__ double s_in[Geo<Pot>::G * Pot::N_IN];
    eval_pgh_body<Pot, Geo<Pot>::BLOCK>(a, blockIdx.x, gridDim.x, s_in);
}

// energy only (line-search evaluations): the same cooperative gather (all loads of a CTA's elements in flight together --
// a thread walking its own fetch table serially makes a 100-element contact table cost 25 us of dependent loads), then one
// thread per element
constexpr int EVAL_P_SMEM_DOUBLES = 5120;   // 40 KB of gathered inputs per CTA
template<class Pot> struct GeoP {
    static constexpr int BY_SMEM = EVAL_P_SMEM_DOUBLES / Pot::N_IN;
    static constexpr int EP = BY_SMEM >= 128 ? 128 : (BY_SMEM / 32) * 32;   // elements per CTA
    static_assert(EP >= 32, "fetch table too long for the shared-memory gather");
};
template<class Pot>
__device__ __forceinline__ void eval_p_body(const FetchSlot* __restrict__ slots, const int32_t* __restrict__ conn, int conn_stride, int n_elem,
                                            double* __restrict__ E_elem, int cta, double* s_in)
{
    constexpr int NIN = Pot::N_IN, EP = GeoP<Pot>::EP;
    const int e_base = cta * EP;
    for (int idx = threadIdx.x; idx < EP * NIN; idx += 128) {
        const int el = idx / NIN, slot = idx - el * NIN;
        const int e = e_base + el;
        if (e < n_elem) {
            const FetchSlot fs = slots[slot];
            const int row = (fs.conn_col >= 0) ? conn[(size_t)e * conn_stride + fs.conn_col] : 0;
            s_in[idx] = fs.base[(size_t)row * fs.stride + fs.off];
        }
    }
    __syncthreads();
    const int e = e_base + threadIdx.x;
    if (threadIdx.x >= EP || e >= n_elem) return;
    sbad::Seed<double> seed;
    E_elem[e] = Pot::template energy<double>(s_in + threadIdx.x * NIN, seed);
}
template<class Pot>
__global__ void __launch_bounds__(128) k_eval_p(const EvalArgs a)
{
    __shared__ double s_in[GeoP<Pot>::EP * Pot::N_IN];
    eval_p_body<Pot>(a.slots, a.conn, a.conn_stride, a.n_elem, a.E_elem, blockIdx.x, s_in);
}

template<class Pot> static void launch_pgh(const EvalArgs& a, cudaStream_t s)
{
    const int grid = (a.n_elem + Geo<Pot>::G - 1) / Geo<Pot>::G;
    k_eval_pgh<Pot><<<grid, Geo<Pot>::BLOCK, 0, s>>>(a);
}
template<class Pot> static void launch_p(const EvalArgs& a, cudaStream_t s)
{
    const int grid = (a.n_elem + GeoP<Pot>::EP - 1) / GeoP<Pot>::EP;
    k_eval_p<Pot><<<grid, 128, 0, s>>>(a);
}


// ---- all small potentials of one energy-only evaluation in ONE launch ----
// A line-search evaluation of a contact scene touches a dozen potentials with a few hundred elements each; launched one by
// one (even spread over side streams) the evaluation is bound by the host's launch rate.  Here the CTA looks up which
// potential its block index falls into and runs that potential's body.
#define SB_ALL_POTS(X) \
    X(EnergyLumpedInertia) \
    X(EnergyPrescribedPositions) \
    X(EnergySegmentStrain) \
    X(EnergySegmentStrain_Elasticity_Only) \
    X(EnergyTriangleStrain) \
    X(EnergyTriangleStrain_Elasticity_Only) \
    X(EnergyDiscreteShells) \
    X(EnergyBendingFlat) \
    X(EnergyTetStrain) \
    X(EnergyTetStrain_Elasticity_Only) \
    X(EnergyRigidBodyInertia_Linear) \
    X(EnergyRigidBodyInertia_Angular) \
    X(rb_constraint_global_points) \
    X(rb_constraint_global_directions) \
    X(rb_constraint_points) \
    X(rb_constraint_point_on_axis) \
    X(rb_constraint_distances) \
    X(rb_constraint_distance_limits) \
    X(rb_constraint_directions) \
    X(rb_constraint_angle_limits) \
    X(rb_constraint_damped_spring) \
    X(rb_constraint_linear_velocity) \
    X(rb_constraint_angular_velocity) \
    X(EnergyAttachments_d_d_p_p) \
    X(EnergyAttachments_d_d_p_e) \
    X(EnergyAttachments_d_d_p_t) \
    X(EnergyAttachments_d_d_e_e) \
    X(EnergyAttachments_rb_d) \
    X(contact_d_d_pt_pp) \
    X(contact_d_d_pt_pe) \
    X(contact_d_d_pt_pt) \
    X(contact_d_d_ee_pp) \
    X(contact_d_d_ee_pe) \
    X(contact_d_d_ee_ee) \
    X(contact_rb_rb_pt_pp) \
    X(contact_rb_rb_pt_pe) \
    X(contact_rb_rb_pt_pt) \
    X(contact_rb_rb_ee_pp) \
    X(contact_rb_rb_ee_pe) \
    X(contact_rb_rb_ee_ee) \
    X(contact_rb_d_pt_pp) \
    X(contact_rb_d_pt_pe) \
    X(contact_rb_d_pt_pt) \
    X(contact_rb_d_pt_ep) \
    X(contact_rb_d_pt_tp) \
    X(contact_rb_d_ee_pp) \
    X(contact_rb_d_ee_pe) \
    X(contact_rb_d_ee_ee) \
    X(contact_rb_d_ee_ep) \
    X(friction_d_d_pp) \
    X(friction_d_d_pe) \
    X(friction_d_d_pt) \
    X(friction_d_d_ee) \
    X(friction_rb_rb_pp) \
    X(friction_rb_rb_pe) \
    X(friction_rb_rb_pt) \
    X(friction_rb_rb_ee) \
    X(friction_rb_d_pp) \
    X(friction_rb_d_pe) \
    X(friction_rb_d_pt) \
    X(friction_rb_d_ee) \
    X(friction_rb_d_ep) \
    X(friction_rb_d_tp)

enum PotKind {
#define X(S) PK_##S,
    SB_ALL_POTS(X)
#undef X
    PK_COUNT
};
__global__ void __launch_bounds__(128) k_eval_p_multi(const __grid_constant__ MultiPArgs M)
{
    __shared__ double s_in[EVAL_P_SMEM_DOUBLES];
    int i = 0;
    while (i + 1 < M.n && (int)blockIdx.x >= M.it[i + 1].cta0) i++;
    const MultiPItem& a = M.it[i];
    const int cta = blockIdx.x - a.cta0;
    switch (a.kind) {
#define X(S) case PK_##S: eval_p_body<sbpot::S>(a.slots, a.conn, a.conn_stride, a.n_elem, a.E_elem, cta, s_in); break;
    SB_ALL_POTS(X)
#undef X
    default: break;
    }
}
int multi_p_ctas(int p_kind, int n_elem)
{
    switch (p_kind) {
#define X(S) case PK_##S: return (n_elem + GeoP<sbpot::S>::EP - 1) / GeoP<sbpot::S>::EP;
    SB_ALL_POTS(X)
#undef X
    default: return 0;
    }
}
void launch_p_multi(const MultiPArgs& M, int total_ctas, cudaStream_t s)
{
    k_eval_p_multi<<<total_ctas, 128, 0, s>>>(M);
}

// ---- the contact and friction tables of one P+G+H evaluation in ONE launch ----
// Seven to thirty-five potentials with a few hundred elements each: launched one by one (a launch, a fetch-table refresh and a
// share of the fork / join events each) they cost the host ~5 us apiece in the stretch between the collision detection and the
// evaluation's reductions, where nothing hides it.  Same scheme as k_eval_p_multi; potentials whose lane count exceeds the
// block (24-DoF elements: 300 lanes) keep their own kernel.
// (six kernels, one per family of tables: ONE kernel holding all 35 differentiated energies needs 255 registers with spills and
//  runs 33 us for a thousand elements -- every CTA executes a different stretch of a very large code; per family it is a third)
#define SB_FAM_0(X) X(contact_d_d_pt_pp) X(contact_d_d_pt_pe) X(contact_d_d_pt_pt) X(contact_d_d_ee_pp) X(contact_d_d_ee_pe) X(contact_d_d_ee_ee)
#define SB_FAM_1(X) X(contact_rb_rb_pt_pp) X(contact_rb_rb_pt_pe) X(contact_rb_rb_pt_pt) X(contact_rb_rb_ee_pp) X(contact_rb_rb_ee_pe) X(contact_rb_rb_ee_ee)
#define SB_FAM_2(X) X(contact_rb_d_pt_pp) X(contact_rb_d_pt_pe) X(contact_rb_d_pt_pt) X(contact_rb_d_pt_ep) X(contact_rb_d_pt_tp) \
                    X(contact_rb_d_ee_pp) X(contact_rb_d_ee_pe) X(contact_rb_d_ee_ee) X(contact_rb_d_ee_ep)
#define SB_FAM_3(X) X(friction_d_d_pp) X(friction_d_d_pe) X(friction_d_d_pt) X(friction_d_d_ee)
#define SB_FAM_4(X) X(friction_rb_rb_pp) X(friction_rb_rb_pe) X(friction_rb_rb_pt) X(friction_rb_rb_ee)
#define SB_FAM_5(X) X(friction_rb_d_pp) X(friction_rb_d_pe) X(friction_rb_d_pt) X(friction_rb_d_ee) X(friction_rb_d_ep) X(friction_rb_d_tp)
#define SB_TABLE_POTS(X) SB_FAM_0(X) SB_FAM_1(X) SB_FAM_2(X) SB_FAM_3(X) SB_FAM_4(X) SB_FAM_5(X)
constexpr int MULTI_G_BLOCK = 256;
template<class Pot> struct MultiGOk { static constexpr bool value = Geo<Pot>::L <= MULTI_G_BLOCK; };
template<class Pot> constexpr int multi_g_smem() { return MultiGOk<Pot>::value ? (MULTI_G_BLOCK / Geo<Pot>::L) * Pot::N_IN : 0; }
constexpr int multi_g_smem_max()
{
    int m = 0;
#define X(S) m = multi_g_smem<sbpot::S>() > m ? multi_g_smem<sbpot::S>() : m;
    SB_TABLE_POTS(X)
#undef X
    return m;
}
template<class Pot, bool OK = MultiGOk<Pot>::value> struct MultiGBody {
    static __device__ __forceinline__ void run(const EvalArgs& a, int cta, int n_ctas, double* s_in) { eval_pgh_body<Pot, MULTI_G_BLOCK>(a, cta, n_ctas, s_in); }
};
template<class Pot> struct MultiGBody<Pot, false> { static __device__ __forceinline__ void run(const EvalArgs&, int, int, double*) {} };
#define SB_MULTI_G_KERNEL(FAM) \
__global__ void __launch_bounds__(MULTI_G_BLOCK) k_eval_pgh_multi_##FAM(const __grid_constant__ MultiGArgs M) \
{ \
    __shared__ double s_in[multi_g_smem_max()]; \
    int i = 0; \
    while (i + 1 < M.n && (int)blockIdx.x >= M.cta0[i + 1]) i++; \
    const EvalArgs& a = M.it[i]; \
    const int cta = blockIdx.x - M.cta0[i]; \
    const int n_ctas = ((i + 1 < M.n) ? M.cta0[i + 1] : (int)gridDim.x) - M.cta0[i];   /* this potential's share of the grid */ \
    switch (M.kind[i]) { \
    SB_FAM_##FAM(SB_MULTI_G_CASE) \
    default: break; \
    } \
}
#define SB_MULTI_G_CASE(S) case PK_##S: MultiGBody<sbpot::S>::run(a, cta, n_ctas, s_in); break;
SB_MULTI_G_KERNEL(0)
SB_MULTI_G_KERNEL(1)
SB_MULTI_G_KERNEL(2)
SB_MULTI_G_KERNEL(3)
SB_MULTI_G_KERNEL(4)
SB_MULTI_G_KERNEL(5)
// CTAs an eligible potential needs in the multi-potential launch (0: not eligible)
int multi_g_ctas(int p_kind, int n_elem)
{
    switch (p_kind) {
#define X(S) case PK_##S: return MultiGOk<sbpot::S>::value ? (n_elem + (MULTI_G_BLOCK / Geo<sbpot::S>::L) - 1) / (MULTI_G_BLOCK / Geo<sbpot::S>::L) : 0;
    SB_TABLE_POTS(X)
#undef X
    default: return 0;
    }
}
int multi_g_family(int p_kind)
{
    switch (p_kind) {
#define X(S) case PK_##S: return 0;
    SB_FAM_0(X)
#undef X
#define X(S) case PK_##S: return 1;
    SB_FAM_1(X)
#undef X
#define X(S) case PK_##S: return 2;
    SB_FAM_2(X)
#undef X
#define X(S) case PK_##S: return 3;
    SB_FAM_3(X)
#undef X
#define X(S) case PK_##S: return 4;
    SB_FAM_4(X)
#undef X
#define X(S) case PK_##S: return 5;
    SB_FAM_5(X)
#undef X
    default: return -1;
    }
}
void launch_pgh_multi(int family, const MultiGArgs& M, int total_ctas, cudaStream_t s)
{
    switch (family) {
    case 0: k_eval_pgh_multi_0<<<total_ctas, MULTI_G_BLOCK, 0, s>>>(M); break;
    case 1: k_eval_pgh_multi_1<<<total_ctas, MULTI_G_BLOCK, 0, s>>>(M); break;
    case 2: k_eval_pgh_multi_2<<<total_ctas, MULTI_G_BLOCK, 0, s>>>(M); break;
    case 3: k_eval_pgh_multi_3<<<total_ctas, MULTI_G_BLOCK, 0, s>>>(M); break;
    case 4: k_eval_pgh_multi_4<<<total_ctas, MULTI_G_BLOCK, 0, s>>>(M); break;
    default: k_eval_pgh_multi_5<<<total_ctas, MULTI_G_BLOCK, 0, s>>>(M); break;
    }
}

}  // namespace sb
#include "tet_analytic.cuh"
namespace sb {

#define SB_KERNEL(STRUCT, NAME) \
    KernelInfo { NAME, sbpot::STRUCT::N_IN, sbpot::STRUCT::N_DOF, sbpot::STRUCT::NB, sbpot::STRUCT::DOF_SLOT, &launch_pgh<sbpot::STRUCT>, &launch_p<sbpot::STRUCT>, PK_##STRUCT }

const std::vector<KernelInfo>& all_kernels()
{
    static const std::vector<KernelInfo> k = {
        SB_KERNEL(EnergyLumpedInertia, "EnergyLumpedInertia"),
        SB_KERNEL(EnergyPrescribedPositions, "EnergyPrescribedPositions"),
        SB_KERNEL(EnergySegmentStrain, "EnergySegmentStrain"),
        SB_KERNEL(EnergySegmentStrain_Elasticity_Only, "EnergySegmentStrain_Elasticity_Only"),
        SB_KERNEL(EnergyTriangleStrain, "EnergyTriangleStrain"),
        SB_KERNEL(EnergyTriangleStrain_Elasticity_Only, "EnergyTriangleStrain_Elasticity_Only"),
        SB_KERNEL(EnergyDiscreteShells, "EnergyDiscreteShells"),
        SB_KERNEL(EnergyBendingFlat, "EnergyBendingFlat"),
        // hand-derived analytic tet kernels (tet_analytic.cuh) are the product path; the AD ones stay as cross-checks
        KernelInfo { "EnergyTetStrain", 43, 12, 4, sbpot::EnergyTetStrain::DOF_SLOT, &launch_tet_analytic_pgh<true>, &launch_tet_analytic_p<true>, -1 },
        KernelInfo { "EnergyTetStrain_Elasticity_Only", 40, 12, 4, sbpot::EnergyTetStrain_Elasticity_Only::DOF_SLOT, &launch_tet_analytic_pgh<false>, &launch_tet_analytic_p<false>, -1 },
        SB_KERNEL(EnergyTetStrain, "EnergyTetStrain_AD"),
        SB_KERNEL(EnergyTetStrain_Elasticity_Only, "EnergyTetStrain_Elasticity_Only_AD"),
        SB_KERNEL(EnergyRigidBodyInertia_Linear, "EnergyRigidBodyInertia_Linear"),
        SB_KERNEL(EnergyRigidBodyInertia_Angular, "EnergyRigidBodyInertia_Angular"),
        SB_KERNEL(rb_constraint_global_points, "rb_constraint_global_points"),
        SB_KERNEL(rb_constraint_global_directions, "rb_constraint_global_directions"),
        SB_KERNEL(rb_constraint_points, "rb_constraint_points"),
        SB_KERNEL(rb_constraint_point_on_axis, "rb_constraint_point_on_axis"),
        SB_KERNEL(rb_constraint_distances, "rb_constraint_distances"),
        SB_KERNEL(rb_constraint_distance_limits, "rb_constraint_distance_limits"),
        SB_KERNEL(rb_constraint_directions, "rb_constraint_directions"),
        SB_KERNEL(rb_constraint_angle_limits, "rb_constraint_angle_limits"),
        SB_KERNEL(rb_constraint_damped_spring, "rb_constraint_damped_spring"),
        SB_KERNEL(rb_constraint_linear_velocity, "rb_constraint_linear_velocity"),
        SB_KERNEL(rb_constraint_angular_velocity, "rb_constraint_angular_velocity"),
        SB_KERNEL(EnergyAttachments_d_d_p_p, "EnergyAttachments_d_d_p_p"),
        SB_KERNEL(EnergyAttachments_d_d_p_e, "EnergyAttachments_d_d_p_e"),
        SB_KERNEL(EnergyAttachments_d_d_p_t, "EnergyAttachments_d_d_p_t"),
        SB_KERNEL(EnergyAttachments_d_d_e_e, "EnergyAttachments_d_d_e_e"),
        SB_KERNEL(EnergyAttachments_rb_d, "EnergyAttachments_rb_d"),
        SB_KERNEL(contact_d_d_pt_pp, "contact_d_d_pt_pp_cubic"),
        SB_KERNEL(contact_d_d_pt_pe, "contact_d_d_pt_pe_cubic"),
        SB_KERNEL(contact_d_d_pt_pt, "contact_d_d_pt_pt_cubic"),
        SB_KERNEL(contact_d_d_ee_pp, "contact_d_d_ee_pp_cubic"),
        SB_KERNEL(contact_d_d_ee_pe, "contact_d_d_ee_pe_cubic"),
        SB_KERNEL(contact_d_d_ee_ee, "contact_d_d_ee_ee_cubic"),
        SB_KERNEL(contact_rb_rb_pt_pp, "contact_rb_rb_pt_pp_cubic"),
        SB_KERNEL(contact_rb_rb_pt_pe, "contact_rb_rb_pt_pe_cubic"),
        SB_KERNEL(contact_rb_rb_pt_pt, "contact_rb_rb_pt_pt_cubic"),
        SB_KERNEL(contact_rb_rb_ee_pp, "contact_rb_rb_ee_pp_cubic"),
        SB_KERNEL(contact_rb_rb_ee_pe, "contact_rb_rb_ee_pe_cubic"),
        SB_KERNEL(contact_rb_rb_ee_ee, "contact_rb_rb_ee_ee_cubic"),
        SB_KERNEL(contact_rb_d_pt_pp, "contact_rb_d_pt_pp_cubic"),
        SB_KERNEL(contact_rb_d_pt_pe, "contact_rb_d_pt_pe_cubic"),
        SB_KERNEL(contact_rb_d_pt_pt, "contact_rb_d_pt_pt_cubic"),
        SB_KERNEL(contact_rb_d_pt_ep, "contact_rb_d_pt_ep_cubic"),
        SB_KERNEL(contact_rb_d_pt_tp, "contact_rb_d_pt_tp_cubic"),
        SB_KERNEL(contact_rb_d_ee_pp, "contact_rb_d_ee_pp_cubic"),
        SB_KERNEL(contact_rb_d_ee_pe, "contact_rb_d_ee_pe_cubic"),
        SB_KERNEL(contact_rb_d_ee_ee, "contact_rb_d_ee_ee_cubic"),
        SB_KERNEL(contact_rb_d_ee_ep, "contact_rb_d_ee_ep_cubic"),
        SB_KERNEL(friction_d_d_pp, "friction_d_d_pp_C0"),
        SB_KERNEL(friction_d_d_pe, "friction_d_d_pe_C0"),
        SB_KERNEL(friction_d_d_pt, "friction_d_d_pt_C0"),
        SB_KERNEL(friction_d_d_ee, "friction_d_d_ee_C0"),
        SB_KERNEL(friction_rb_rb_pp, "friction_rb_rb_pp_C0"),
        SB_KERNEL(friction_rb_rb_pe, "friction_rb_rb_pe_C0"),
        SB_KERNEL(friction_rb_rb_pt, "friction_rb_rb_pt_C0"),
        SB_KERNEL(friction_rb_rb_ee, "friction_rb_rb_ee_C0"),
        SB_KERNEL(friction_rb_d_pp, "friction_rb_d_pp_C0"),
        SB_KERNEL(friction_rb_d_pe, "friction_rb_d_pe_C0"),
        SB_KERNEL(friction_rb_d_pt, "friction_rb_d_pt_C0"),
        SB_KERNEL(friction_rb_d_ee, "friction_rb_d_ee_C0"),
        SB_KERNEL(friction_rb_d_ep, "friction_rb_d_ep_C0"),
        SB_KERNEL(friction_rb_d_tp, "friction_rb_d_tp_C0"),
    };
    return k;
}

// CUDA loads a kernel's code at its first launch (lazy module loading): a contact or friction potential first used in the
// middle of a run would stall that Newton iteration for milliseconds.  Touching a kernel's attributes loads it.
template<class Pot> static void preload_pot()
{
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_eval_pgh<Pot>);
    cudaFuncGetAttributes(&fa, k_eval_p<Pot>);
}
void preload_eval_kernels()
{
#define X(S) preload_pot<sbpot::S>();
    SB_ALL_POTS(X)
#undef X
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_eval_p_multi);
    cudaFuncGetAttributes(&fa, k_eval_pgh_multi_0); cudaFuncGetAttributes(&fa, k_eval_pgh_multi_1); cudaFuncGetAttributes(&fa, k_eval_pgh_multi_2);
    cudaFuncGetAttributes(&fa, k_eval_pgh_multi_3); cudaFuncGetAttributes(&fa, k_eval_pgh_multi_4); cudaFuncGetAttributes(&fa, k_eval_pgh_multi_5);
    cudaFuncGetAttributes(&fa, k_tet_analytic<true, true>); cudaFuncGetAttributes(&fa, k_tet_analytic<true, false>);
    cudaFuncGetAttributes(&fa, k_tet_analytic<false, true>); cudaFuncGetAttributes(&fa, k_tet_analytic<false, false>);
    cudaFuncGetAttributes(&fa, k_tet_energy<true, true>); cudaFuncGetAttributes(&fa, k_tet_energy<true, false>);
    cudaFuncGetAttributes(&fa, k_tet_energy<false, true>); cudaFuncGetAttributes(&fa, k_tet_energy<false, false>);
    cudaGetLastError();
}

const KernelInfo* find_kernel(const char* name)
{
    for (const auto& k : all_kernels())
        if (std::strcmp(k.name, name) == 0) return &k;
    return nullptr;
}

}  // namespace sb
