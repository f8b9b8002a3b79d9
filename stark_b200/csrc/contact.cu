// Collision detection and contact / friction list construction on the device.
//
// Replaces, for the Newton hot path:
//   * tmcd::ProximityDetection::run  (tmcd/ProximityDetection.cpp:74-220): float AABBs enlarged by (enl + FLT_EPSILON)
//     (tmcd/AABBs.cpp:22-180), candidate pairs = overlapping AABBs minus shared-vertex ("orphan") and blacklisted pairs
//     (tmcd/BroadPhasePTEEBase.cpp:270-420), narrow phase = IPC-toolkit distance-type classification + squared distance
//     (tmcd/ipc_toolkit_geometry_functions.cpp:42-306), kept if d^2 < enl^2;
//   * tmcd::IntersectionDetection::run (tmcd/IntersectionDetection.cpp:54-108, geometry :565-585);
//   * EnergyFrictionalContact::{_update_vertices, _before_energy_evaluation__update_contacts,
//     _before_time_step__update_friction_contacts} (S/models/interactions/EnergyFrictionalContact.cpp:219-262, 368-773):
//     the serial host loops that turn proximity pairs into the 21 connectivity tables and the friction data
//     (T, bary, fn, mu -- S/models/interactions/friction_geometry.cpp).
//
// The reference's octree only de-duplicates and prunes; the candidate SET is "all overlapping AABB pairs", which is what
// the tiled all-pairs kernels below enumerate directly (one CTA per 256 x 256 tile, AABBs staged in shared memory).
// Candidates are compacted into a list and classified one pair per thread by a second kernel, so the long narrow-phase
// code does not diverge inside the tile loop.  Compile this file with -fmad=false: the classification thresholds of the
// narrow phase are decided by the last bit for grid-aligned scenes, and un-fused IEEE arithmetic in the reference's
// operation order is the closest reproducible restatement.
#include "internal.h"
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <cstdlib>
#include <cub/cub.cuh>

namespace sb {

enum Role { R_SOFT_V1, R_SOFT_X0, R_SOFT_X, R_RB_V1, R_RB_W1, R_RB_T0, R_RB_Q0, R_DT, R_THICKNESS, R_STIFFNESS, R_RB_LOCAL, R_EPSV,
            R_F_T, R_F_BARY, R_F_MU, R_F_FN };
struct LayoutEntry { int role, col, slot, stride; };
struct Layout { const char* name; int conn_stride; int n; LayoutEntry e[40]; };
static const Layout LAYOUTS[] = {
#include "contact_layouts.inc"
};
constexpr int N_TABLES = 35;       // 21 contact + 14 friction, in the order of contact_layouts.inc
constexpr int N_CONTACT_TABLES = 21;
constexpr int N_FRICTION = 14;
enum Table {
    CT_DD_PT_PP, CT_DD_PT_PE, CT_DD_PT_PT, CT_DD_EE_PP, CT_DD_EE_PE, CT_DD_EE_EE,
    CT_RR_PT_PP, CT_RR_PT_PE, CT_RR_PT_PT, CT_RR_EE_PP, CT_RR_EE_PE, CT_RR_EE_EE,
    CT_RD_PT_PP, CT_RD_PT_PE, CT_RD_PT_PT, CT_RD_PT_EP, CT_RD_PT_TP, CT_RD_EE_PP, CT_RD_EE_PE, CT_RD_EE_EE, CT_RD_EE_EP,
    FT_DD_PP, FT_DD_PE, FT_DD_PT, FT_DD_EE, FT_RR_PP, FT_RR_PE, FT_RR_PT, FT_RR_EE, FT_RD_PP, FT_RD_PE, FT_RD_PT, FT_RD_EE, FT_RD_EP, FT_RD_TP
};
constexpr int N_LISTS = 7;         // pt_pp pt_pe pt_pt ee_pp ee_pe ee_ee intersections
static const int LIST_WIDTH[N_LISTS] = {5, 6, 4, 6, 5, 4, 4};
constexpr int MAX_GROUPS = 64;

struct Group {
    int ps, body, n_v, n_t, n_e;
    int v_off, t_off, e_off;   // offsets in the concatenated collision vertex / triangle / edge arrays
    double thickness;
};

// everything the device kernels need, passed by value
struct Dev {
    int n_v, n_t, n_e, n_groups;
    double* x;                   // [n_v][3] current collision vertex positions
    const int32_t* v_group;      // [n_v]
    const int32_t* v_ps_index;   // [n_v] soft: global node; rigid: row in the rigid local-vertex array
    const int32_t* tri;          // [n_t][3] global collision vertex ids
    const int32_t* t_group;
    const int32_t* edge;         // [n_e][2]
    const int32_t* e_group;
    float* bb_p; float* bb_t; float* bb_e;   // [n][6] bottom xyz, top xyz
    const int32_t* perm_p; const int32_t* perm_t; const int32_t* perm_e;   // slot -> primitive (Morton order of the AABB centres)
    float* tb_p; float* tb_t; float* tb_e;   // [n_tiles][6] boxes of TILE consecutive SLOTS
    int2* tile_pairs[3];                     // per broad-phase kind: overlapping (tile A, tile B) pairs
    int tile_pair_cap[3];
    const int32_t* g_ps; const int32_t* g_body; const int32_t* g_voff; const int32_t* g_toff; const int32_t* g_eoff;
    const double* g_thickness;
    const uint8_t* blacklist;    // [MAX_GROUPS * MAX_GROUPS]
    const double* mu;            // [MAX_GROUPS * MAX_GROUPS]
    // candidates
    int2* cand_pt; int2* cand_ee; int2* cand_et;
    int cand_cap;
    int* counters;               // [0] pt cands [1] ee cands [2] et cands [3] overflow flag, [8 + l] list counts, [16 + t] table counts
    // raw results
    int32_t* list_ids[N_LISTS];
    double* list_dist[N_LISTS];
    int list_cap;
    // contact / friction tables
    int32_t* table[N_TABLES];
    int table_stride[N_TABLES];
    int table_cap;
    unsigned long long* hash;    // [2 * N_CONTACT_TABLES] order-independent 128-bit digest of every contact table's rows
    double* fT[N_FRICTION]; double* fmu[N_FRICTION]; double* ffn[N_FRICTION]; double* fbary[N_FRICTION];
};

struct Contact {
    sb_contact_bindings bind;
    std::vector<Group> groups;
    std::vector<int32_t> h_v_group, h_v_ps, h_tri, h_t_group, h_edge, h_e_group;
    std::vector<double> h_rb_local;          // concatenated rigid local vertices
    std::vector<uint8_t> h_blacklist;
    std::vector<double> h_mu;
    bool topology_dirty = true;
    double stiffness = 1e3, epsv = 0.1;
    int enable_pt = 1, enable_ee = 1, enable_friction = 1;
    // ctx arrays owned by the module
    int a_thickness = -1, a_stiffness = -1, a_rb_local = -1, a_epsv = -1;
    int a_fT[N_FRICTION], a_fbary[N_FRICTION], a_fmu[N_FRICTION], a_ffn[N_FRICTION];
    int pot[N_TABLES];
    // device storage
    DevBuf<double> x;
    DevBuf<int32_t> v_group, v_ps, tri, t_group, edge, e_group, g_i32;
    DevBuf<double> g_f64, mu;
    DevBuf<uint8_t> blacklist;
    DevBuf<float> bb_p, bb_t, bb_e, tb_p, tb_t, tb_e;
    DevBuf<float> bb0_p, bb0_t, bb0_e, tb0_p, tb0_t, tb0_e;   // the intersection pass's own (non-enlarged) boxes: it runs beside the proximity pass
    DevBuf<int32_t> perm_p, perm_t, perm_e, perm_tmp;
    DevBuf<uint32_t> morton, morton_tmp;
    DevBuf<float> bounds;            // scene box (6 floats)
    DevBuf<uint8_t> sort_temp;
    int reorder_countdown = 0;       // detections until the spatial order of the primitives is refreshed
    DevBuf<int2> tile_pairs[3];
    DevBuf<int2> cand_pt, cand_ee, cand_et;
    DevBuf<int> counters;
    DevBuf<int32_t> list_ids[N_LISTS];
    DevBuf<double> list_dist[N_LISTS];
    DevBuf<int32_t> table[N_TABLES];
    // Contact tables are written into BACK buffers and swapped in only when the new tables differ from the current ones as sets
    // of rows (count + 128-bit digest per table): in resting / sliding contact consecutive detections find the same pairs, and
    // keeping the current tables (same rows, old order) keeps the potentials' connectivity, the element layout and the symbolic
    // phase of the assembly valid -- nothing downstream changes.
    DevBuf<int32_t> table_back[N_CONTACT_TABLES];
    DevBuf<unsigned long long> hash;
    unsigned long long* h_hash = nullptr;
    unsigned long long cur_hash[2 * N_CONTACT_TABLES] = {0};
    bool cur_valid = false;
    long long n_tables_same = 0, n_tables_changed = 0;
    int cand_cap = 1 << 16, list_cap = 1 << 14, table_cap = 1 << 14;
    int* h_counters = nullptr;
    int h_list_count[N_LISTS] = {0};
    int h_table_count[N_TABLES] = {0};
    bool external_vertices = false;
    // detection results are a pure function of the state: a repeated call at an unchanged state (the line search's last
    // trial and the next iteration's evaluation see the same DoFs) is answered from the cache
    double uploaded_stiffness = -1.0, uploaded_epsv = -1.0;
    uint64_t contacts_state = 0, intersections_state = 0;
    int cached_intersections = 0;
};

// ---------------------------------------------------------------------------------------------------
// vertex update (EnergyFrictionalContact.cpp:219-250)
// ---------------------------------------------------------------------------------------------------
__global__ void k_update_vertices(Dev d, const double* __restrict__ v1, const double* __restrict__ x0, const double* __restrict__ rb_v, const double* __restrict__ rb_w,
                                  const double* __restrict__ rb_t0, const double* __restrict__ rb_q0, const double* __restrict__ rb_local, const double* __restrict__ dt_ptr, int zero_dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_v) return;
    const double dt = zero_dt ? 0.0 : dt_ptr[0];
    const int g = d.v_group[i];
    const int k = d.v_ps_index[i];
    if (d.g_ps[g] == 0) {
        for (int c = 0; c < 3; c++) d.x[3 * i + c] = x0[3 * k + c] + dt * v1[3 * k + c];
    } else {
        const int b = d.g_body[g];
        // quat_time_integration (rigidbody_transformations.cpp:33-40): q1 = normalize(q0 + dt/2 (0,w) * q0); q0 as (w,x,y,z)
        const double qw = rb_q0[4 * b], qx = rb_q0[4 * b + 1], qy = rb_q0[4 * b + 2], qz = rb_q0[4 * b + 3];
        const double wx = rb_w[3 * b], wy = rb_w[3 * b + 1], wz = rb_w[3 * b + 2];
        const double pw = -wx * qx - wy * qy - wz * qz;
        const double px = wx * qw + wy * qz - wz * qy;
        const double py = wy * qw + wz * qx - wx * qz;
        const double pz = wz * qw + wx * qy - wy * qx;
        const double h = 0.5 * dt;
        double ew = qw + h * pw, ex = qx + h * px, ey = qy + h * py, ez = qz + h * pz;
        const double n = sqrt(ex * ex + ey * ey + ez * ez + ew * ew);
        ew /= n; ex /= n; ey /= n; ez /= n;
        const double tx = 2.0 * ex, ty = 2.0 * ey, tz = 2.0 * ez;
        const double twx = tx * ew, twy = ty * ew, twz = tz * ew;
        const double txx = tx * ex, txy = ty * ex, txz = tz * ex;
        const double tyy = ty * ey, tyz = tz * ey, tzz = tz * ez;
        const double R[9] = {1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1.0 - (txx + tyy)};
        const double X0 = rb_local[3 * k], X1 = rb_local[3 * k + 1], X2 = rb_local[3 * k + 2];
        for (int c = 0; c < 3; c++) {
            const double t1 = rb_t0[3 * b + c] + dt * rb_v[3 * b + c];
            d.x[3 * i + c] = t1 + (R[3 * c] * X0 + R[3 * c + 1] * X1 + R[3 * c + 2] * X2);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// AABBs (tmcd/AABBs.cpp:22-140): float min / max of the vertices, enlarged by extra = (float)enl + FLT_EPSILON
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bb_init(float* b) { b[0] = b[1] = b[2] = FLT_MAX; b[3] = b[4] = b[5] = -FLT_MAX; }
__device__ __forceinline__ void bb_expand(float* b, const double* x)
{
    for (int c = 0; c < 3; c++) { const float v = (float)x[c]; b[c] = fminf(b[c], v); b[3 + c] = fmaxf(b[3 + c], v); }
}
constexpr int TILE = 64;   // slots per broad-phase tile

// AABBs are stored per SLOT: slot i holds primitive perm[i] (spatially sorted, so that a tile of consecutive slots is compact)
__global__ void k_aabbs(Dev d, float extra, int do_points)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float b[6];
    if (do_points && i < d.n_v) {
        const int p = d.perm_p[i];
        bb_init(b); bb_expand(b, d.x + 3 * p);
        for (int c = 0; c < 3; c++) { d.bb_p[6 * i + c] = b[c] - extra; d.bb_p[6 * i + 3 + c] = b[3 + c] + extra; }
    }
    if (i < d.n_t) {
        const int t = d.perm_t[i];
        bb_init(b);
        for (int k = 0; k < 3; k++) bb_expand(b, d.x + 3 * d.tri[3 * t + k]);
        for (int c = 0; c < 3; c++) { d.bb_t[6 * i + c] = b[c] - extra; d.bb_t[6 * i + 3 + c] = b[3 + c] + extra; }
    }
    if (i < d.n_e) {
        const int e = d.perm_e[i];
        bb_init(b);
        for (int k = 0; k < 2; k++) bb_expand(b, d.x + 3 * d.edge[2 * e + k]);
        for (int c = 0; c < 3; c++) { d.bb_e[6 * i + c] = b[c] - extra; d.bb_e[6 * i + 3 + c] = b[3 + c] + extra; }
    }
}

// One launch per detection: AABBs of every slot, the box of every tile (CTA = tile) and the counters this detection rewrites.
// CTAs [0, Tv) are point tiles, [Tv, Tv + Tt) triangle tiles, the rest edge tiles.  clear_mask bit k clears counters
// [8 k, 8 k + 8) (k = 0: candidates / tile pairs / overflow, 1: proximity lists, 2-4: contact tables, 5-6: friction tables ...).
__global__ void __launch_bounds__(TILE) k_aabbs_tiles(Dev d, float extra, int do_points, int Tv, int Tt, unsigned long long clear_mask)
{
    __shared__ float s_lo[3][TILE / 32], s_hi[3][TILE / 32];
    if (blockIdx.x == 0 && threadIdx.x < 64 && ((clear_mask >> threadIdx.x) & 1ull)) d.counters[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x < 2 * N_CONTACT_TABLES && ((clear_mask >> 16) & 1ull)) d.hash[threadIdx.x] = 0ull;   // (the contact tables are rewritten)
    int cls, tile;
    if ((int)blockIdx.x < Tv) { cls = 0; tile = blockIdx.x; }
    else if ((int)blockIdx.x < Tv + Tt) { cls = 1; tile = blockIdx.x - Tv; }
    else { cls = 2; tile = blockIdx.x - Tv - Tt; }
    if (cls == 0 && !do_points) return;
    const int n = (cls == 0) ? d.n_v : (cls == 1 ? d.n_t : d.n_e);
    float* bb = (cls == 0) ? d.bb_p : (cls == 1 ? d.bb_t : d.bb_e);
    float* tb = (cls == 0) ? d.tb_p : (cls == 1 ? d.tb_t : d.tb_e);
    const int i = tile * TILE + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        float b[6];
        bb_init(b);
        if (cls == 0) bb_expand(b, d.x + 3 * d.perm_p[i]);
        else if (cls == 1) { const int t = d.perm_t[i]; for (int k = 0; k < 3; k++) bb_expand(b, d.x + 3 * d.tri[3 * t + k]); }
        else { const int e = d.perm_e[i]; for (int k = 0; k < 2; k++) bb_expand(b, d.x + 3 * d.edge[2 * e + k]); }
        for (int c = 0; c < 3; c++) { lo[c] = b[c] - extra; hi[c] = b[3 + c] + extra; bb[6 * i + c] = lo[c]; bb[6 * i + 3 + c] = hi[c]; }
    }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) { lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o)); hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o)); }
        if ((threadIdx.x & 31) == 0) { s_lo[c][threadIdx.x >> 5] = lo[c]; s_hi[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float l = FLT_MAX, h = -FLT_MAX;
        for (int w = 0; w < TILE / 32; w++) { l = fminf(l, s_lo[c][w]); h = fmaxf(h, s_hi[c][w]); }
        tb[6 * tile + c] = l; tb[6 * tile + 3 + c] = h;
    }
}

// ---- spatial order of the primitives: 30-bit Morton code of the AABB centre inside the scene box ----
__global__ void k_iota(int32_t* __restrict__ p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
// scene box = union of the tile boxes of one primitive class (one CTA)
__global__ void __launch_bounds__(256) k_scene_bounds(const float* __restrict__ tb, int n_tiles, float* __restrict__ bounds, int accumulate)
{
    __shared__ float s[6][256];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int t = threadIdx.x; t < n_tiles; t += 256)
        for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], tb[6 * t + c]); hi[c] = fmaxf(hi[c], tb[6 * t + 3 + c]); }
    for (int c = 0; c < 3; c++) { s[c][threadIdx.x] = lo[c]; s[3 + c][threadIdx.x] = hi[c]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int c = threadIdx.x;
        float v = s[c][0];
        for (int k = 1; k < 256; k++) v = (c < 3) ? fminf(v, s[c][k]) : fmaxf(v, s[c][k]);
        if (accumulate) v = (c < 3) ? fminf(v, bounds[c]) : fmaxf(v, bounds[c]);
        bounds[c] = v;
    }
}
__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
    v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_morton(const float* __restrict__ bb, const int32_t* __restrict__ perm, const float* __restrict__ bounds,
                         uint32_t* __restrict__ code, int32_t* __restrict__ ids, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t q[3];
    for (int c = 0; c < 3; c++) {
        const float ext = fmaxf(bounds[3 + c] - bounds[c], 1e-30f);
        const float u = (0.5f * (bb[6 * i + c] + bb[6 * i + 3 + c]) - bounds[c]) / ext;
        q[c] = (uint32_t)fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f);
    }
    // A primitive much larger than its neighbours (a floor triangle among the elements of a fine mesh) would blow up the box of
    // whatever tile its centre falls into, and that tile would then be tested against every other one: such primitives are
    // sorted behind all others (bit 30) and share tiles among themselves.
    float ext = 0.0f, scene = 0.0f;
    for (int c = 0; c < 3; c++) { ext = fmaxf(ext, bb[6 * i + 3 + c] - bb[6 * i + c]); scene = fmaxf(scene, bounds[3 + c] - bounds[c]); }
    const uint32_t big = (ext > 0.0625f * scene) ? 0x40000000u : 0u;
    code[i] = big | (spread3(q[2]) << 2) | (spread3(q[1]) << 1) | spread3(q[0]);
    ids[i] = perm[i];
}

// tmcd/helpers.h:35-44
__device__ __forceinline__ bool bb_overlap(const float* a, const float* b)
{
    bool o = true;
    for (int c = 0; c < 3; c++) if (a[c] > b[3 + c] || b[c] > a[3 + c]) o = false;
    return o;
}

// ---------------------------------------------------------------------------------------------------
// broad phase: tiled all-pairs.  KIND 0: point(A) x triangle(B); 1: edge(A) x edge(B), A < B; 2: edge(A) x triangle(B)
// ---------------------------------------------------------------------------------------------------

// box of every tile of TILE consecutive primitives (one CTA per tile)
__global__ void __launch_bounds__(TILE) k_tile_boxes(const float* __restrict__ bb, float* __restrict__ tb, int n)
{
    __shared__ float s_lo[3][TILE / 32], s_hi[3][TILE / 32];
    const int i = blockIdx.x * TILE + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) for (int c = 0; c < 3; c++) { lo[c] = bb[6 * i + c]; hi[c] = bb[6 * i + 3 + c]; }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) { lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o)); hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o)); }
        if ((threadIdx.x & 31) == 0) { s_lo[c][threadIdx.x >> 5] = lo[c]; s_hi[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float l = FLT_MAX, h = -FLT_MAX;
        for (int w = 0; w < TILE / 32; w++) { l = fminf(l, s_lo[c][w]); h = fmaxf(h, s_hi[c][w]); }
        tb[6 * blockIdx.x + c] = l; tb[6 * blockIdx.x + 3 + c] = h;
    }
}

// overlapping tile pairs of one broad-phase kind (KIND 1: edge tiles, B >= A only)
template<int KIND>
__device__ __forceinline__ void tile_pair_test(const Dev& d, int t, int nTa, int nTb)
{
    const int ta = t / nTb, tb = t - ta * nTb;
    if (KIND == 1 && tb < ta) return;
    const float* A = ((KIND == 0) ? d.tb_p : d.tb_e) + 6 * ta;
    const float* B = ((KIND == 1) ? d.tb_e : d.tb_t) + 6 * tb;
    if (!bb_overlap(A, B)) return;
    const int slot = atomicAdd(d.counters + 4 + KIND, 1);
    if (slot < d.tile_pair_cap[KIND]) d.tile_pairs[KIND][slot] = make_int2(ta, tb); else d.counters[3] = 1;
}
// proximity: point-triangle (n0 = Tv Tt threads) and edge-edge (n1 = Te Te threads) in one launch; intersections: n0 = 0
// and the n1 range tests edge-triangle tiles
__global__ void k_tile_pairs_all(Dev d, int Tv, int Tt, int Te, int n0, int n1, int intersections)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n0) tile_pair_test<0>(d, t, Tv, Tt);
    else if (t < n0 + n1) {
        if (intersections) tile_pair_test<2>(d, t - n0, Te, Tt);
        else tile_pair_test<1>(d, t - n0, Te, Te);
    }
}

// all-pairs test inside the overlapping tile pairs (persistent CTAs, B tile staged in shared memory)
constexpr int QCAP = 384;      // (a tile pair yields a dozen candidates on average; 4 KB of queue lets 32 CTAs share an SM)
struct alignas(16) BroadSmem {
    float bb[TILE][8];       // (min xyz, max xyz, 2 pad: one 128-bit + one 64-bit shared load per box)
    int v[TILE][3];
    int g[TILE];
    int id[TILE];
    int2 q[QCAP];
    int qn, qbase;
    int slot[TILE];          // slot of every staged B primitive (the B tile is compacted, see broad_kind)
    int nb;                  // staged B primitives
};
template<int KIND>
__device__ void broad_kind(const Dev& d, BroadSmem& S, int first_pair)
{
    float (*s_bb)[8] = S.bb;
    int (*s_v)[3] = S.v;
    int* s_g = S.g;
    int* s_id = S.id;
    int2* s_q = S.q;
    int& s_qn = S.qn;
    int& s_qbase = S.qbase;
    int* s_slot = S.slot;
    int& s_nb = S.nb;
    const float* tbA = (KIND == 0) ? d.tb_p : d.tb_e;
    const float* tbB = (KIND == 1) ? d.tb_e : d.tb_t;
    const int nA = (KIND == 0) ? d.n_v : d.n_e;
    const int nB = (KIND == 1) ? d.n_e : d.n_t;
    const float* bbA = (KIND == 0) ? d.bb_p : d.bb_e;
    const float* bbB = (KIND == 1) ? d.bb_e : d.bb_t;
    const int32_t* permA = (KIND == 0) ? d.perm_p : d.perm_e;
    const int32_t* permB = (KIND == 1) ? d.perm_e : d.perm_t;
    // candidates of one tile pair are queued in shared memory and appended to the global list with ONE global atomic
    // (tens of thousands of same-address global atomics would serialise in L2 and dominate the kernel)
    int2* out = (KIND == 0) ? d.cand_pt : (KIND == 1 ? d.cand_ee : d.cand_et);
    const int n_pairs = min(d.counters[4 + KIND], d.tile_pair_cap[KIND]);
    for (int pi = first_pair; pi < n_pairs; pi += gridDim.x) {
        const int2 tp = d.tile_pairs[KIND][pi];
        const int a = tp.x * TILE + threadIdx.x;     // slot of A
        const int b0 = tp.y * TILE;                  // first slot of the B tile
        __syncthreads();   // previous pair done with the shared tile and queue
        if (threadIdx.x == 0) { s_qn = 0; s_nb = 0; }
        __syncthreads();
        {
            // Two tiles that overlap usually share a boundary strip only: B primitives outside the box of tile A cannot overlap
            // any A primitive and are not staged (and A primitives outside the box of tile B sit the pair out, below) -- the
            // all-pairs loop runs over what is left.
            const int b = b0 + threadIdx.x;
            if (b < nB) {
                float bx[6];
                for (int c = 0; c < 6; c++) bx[c] = bbB[6 * b + c];
                if (bb_overlap(bx, tbA + 6 * tp.x)) {
                    const int k = atomicAdd(&s_nb, 1);
                    for (int c = 0; c < 6; c++) s_bb[k][c] = bx[c];
                    const int pb = permB[b];
                    s_id[k] = pb;
                    s_slot[k] = b;
                    if (KIND == 1) { s_v[k][0] = d.edge[2 * pb]; s_v[k][1] = d.edge[2 * pb + 1]; s_v[k][2] = -1; s_g[k] = d.e_group[pb]; }
                    else { s_v[k][0] = d.tri[3 * pb]; s_v[k][1] = d.tri[3 * pb + 1]; s_v[k][2] = d.tri[3 * pb + 2]; s_g[k] = d.t_group[pb]; }
                }
            }
        }
        __syncthreads();
        {
            // (every thread walks the B tile, so that the warp can skip a box none of its lanes overlaps with one vote: with a
            //  hit rate of a few per thousand tests, everything after the box test is off the common path)
            const int aa = (a < nA) ? a : 0;
            float ba[6];
            for (int c = 0; c < 6; c++) ba[c] = bbA[6 * aa + c];
            const bool act = a < nA && bb_overlap(ba, tbB + 6 * tp.y);
            const int pa = permA[aa];
            int va0, va1, ga;
            if (KIND == 0) { va0 = pa; va1 = -2; ga = d.v_group[pa]; }
            else { va0 = d.edge[2 * pa]; va1 = d.edge[2 * pa + 1]; ga = d.e_group[pa]; }
            const int nb = __any_sync(0xffffffffu, act) ? s_nb : 0;   // (a warp with no A primitive inside tile B's box skips the pair)
            for (int j = 0; j < nb; j++) {
                const int b = s_slot[j];
                const float4 blo = *reinterpret_cast<const float4*>(&s_bb[j][0]);   // min x, min y, min z, max x
                const float2 bhi = *reinterpret_cast<const float2*>(&s_bb[j][4]);   // max y, max z
                bool hit = act && !(ba[0] > blo.w || blo.x > ba[3] || ba[1] > bhi.x || blo.y > ba[4] || ba[2] > bhi.y || blo.z > ba[5]);
                if (KIND == 1 && b <= a) hit = false;  // every unordered pair of slots once
                if (!__any_sync(0xffffffffu, hit)) continue;
                if (!hit) continue;
                const int gb = s_g[j];
                // shared-vertex ("orphan") discard: same set and a common vertex (collision vertex ids are global, so equality suffices)
                const bool orphan = (va0 == s_v[j][0]) || (va0 == s_v[j][1]) || (va0 == s_v[j][2]) || (va1 == s_v[j][0]) || (va1 == s_v[j][1]) || (va1 == s_v[j][2]);
                if (orphan) continue;
                if (d.blacklist[ga * MAX_GROUPS + gb]) continue;
                const int pb = s_id[j];
                // edge-edge pairs keep the reference's role convention: first edge = lower primitive id
                const int2 cand = (KIND == 1 && pb < pa) ? make_int2(pb, pa) : make_int2(pa, pb);
                const int k = atomicAdd(&s_qn, 1);
                if (k < QCAP) s_q[k] = cand;
                else {   // queue full (dense contact): straight to the global list
                    const int slot = atomicAdd(d.counters + KIND, 1);
                    if (slot < d.cand_cap) out[slot] = cand; else d.counters[3] = 1;
                }
            }
        }
        __syncthreads();
        const int nq = min(s_qn, QCAP);
        if (threadIdx.x == 0 && nq > 0) s_qbase = atomicAdd(d.counters + KIND, nq);
        __syncthreads();
        if (nq > 0) {
            const int base = s_qbase;
            for (int i = threadIdx.x; i < nq; i += TILE) {
                if (base + i < d.cand_cap) out[base + i] = s_q[i]; else d.counters[3] = 1;
            }
        }
    }
}

// all broad-phase kinds of one detection in one persistent launch (K1 < 0: only K0)
template<int K0, int K1>
__global__ void __launch_bounds__(TILE) k_broad_all(Dev d)
{
    __shared__ BroadSmem S;
    broad_kind<K0>(d, S, blockIdx.x);
    if (K1 >= 0) {
        // the second kind's pairs start where the first kind's left off: the kernel is bound by the latency of ONE tile pair
        // (dependent gathers, a few barriers), so what matters is that no CTA gets more pairs than ceil(total / grid)
        const int n0 = min(d.counters[4 + K0], d.tile_pair_cap[K0]);
        const int first = (int)((blockIdx.x + gridDim.x - (unsigned)(n0 % (int)gridDim.x)) % gridDim.x);
        __syncthreads();
        broad_kind<(K1 >= 0 ? K1 : 0)>(d, S, first);
    }
}

// ---------------------------------------------------------------------------------------------------
// narrow phase geometry (tmcd/ipc_toolkit_geometry_functions.cpp), same operation order
// ---------------------------------------------------------------------------------------------------
struct V { double x, y, z; };
__device__ __forceinline__ V ld(const double* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V operator*(double s, V a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dotv(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V crossv(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double sqn(V a) { return dotv(a, a); }
__device__ __forceinline__ V normalizedv(V a) { const double n = sqrt(sqn(a)); return {a.x / n, a.y / n, a.z / n}; }

__device__ __forceinline__ double pp_sq(V p0, V p1) { return sqn(p1 - p0); }
__device__ __forceinline__ double pl_sq(V p, V e0, V e1) { return sqn(crossv(e0 - p, e1 - p)) / sqn(e1 - e0); }
__device__ __forceinline__ double ll_sq(V ea0, V ea1, V eb0, V eb1)
{
    const V n = crossv(ea1 - ea0, eb1 - eb0);
    const double l = dotv(eb0 - ea0, n);
    return l * l / sqn(n);
}
__device__ __forceinline__ double ppl_sq(V p, V t0, V t1, V t2)
{
    const V n = crossv(t1 - t0, t2 - t0);
    const double l = dotv(p - t0, n);
    return l * l / sqn(n);
}

// sympy-unrolled edge parametrisation (:204-247)
__device__ void edge_param(V p, V e0, V e1, V n, double& p0, double& p1)
{
    const double x0 = e0.x * e0.x, x1 = e0.y * e0.y, x2 = e0.z * e0.z;
    const double x3 = e1.x * e1.x, x4 = e1.y * e1.y, x5 = e1.z * e1.z;
    const double x6 = 2 * e0.x, x7 = e1.x * x6, x8 = 2 * e0.y, x9 = e1.y * x8, x10 = 2 * e1.z, x11 = e0.z * x10;
    const double x12 = -e0.x + e1.x, x13 = -e0.x + p.x, x14 = -e0.y + e1.y, x15 = -e0.y + p.y, x16 = -e0.z + e1.z, x17 = -e0.z + p.z;
    const double x18 = n.x * x6, x19 = n.y * x18, x20 = e0.z * n.z, x21 = e1.z * n.z, x22 = n.y * x8, x23 = e1.x * n.x, x24 = 2 * x23;
    const double x25 = e1.y * n.y, x26 = n.z * x10, x27 = n.y * n.y, x28 = n.z * n.z, x29 = n.x * n.x;
    p0 = (x12 * x13 + x14 * x15 + x16 * x17) / (x0 + x1 - x11 + x2 + x3 + x4 + x5 - x7 - x9);
    p1 = (x13 * (-n.y * x16 + n.z * x14) + x15 * (n.x * x16 - n.z * x12) + x17 * (-n.x * x14 + n.y * x12))
       / (-e0.y * x19 + e1.y * x19 + x0 * x27 + x0 * x28 + x1 * x28 + x1 * x29 - x11 * x27 - x11 * x29 - x18 * x20 + x18 * x21 + x2 * x27 + x2 * x29
          - x20 * x22 + x20 * x24 + 2 * x20 * x25 + x21 * x22 + x22 * x23 - x23 * x26 - x24 * x25 - x25 * x26 + x27 * x3 + x27 * x5 - x27 * x7 + x28 * x3
          + x28 * x4 - x28 * x7 - x28 * x9 + x29 * x4 + x29 * x5 - x29 * x9);
}
enum PT { P_T0, P_T1, P_T2, P_E0, P_E1, P_E2, P_T };
__device__ int pt_type(V p, V t0, V t1, V t2)
{
    const V n = crossv(t1 - t0, t2 - t0);
    double a0, a1, b0, b1, c0, c1;
    edge_param(p, t0, t1, n, a0, a1);
    if (a0 > 0.0 && a0 < 1.0 && a1 >= 0.0) return P_E0;
    edge_param(p, t1, t2, n, b0, b1);
    if (b0 > 0.0 && b0 < 1.0 && b1 >= 0.0) return P_E1;
    edge_param(p, t2, t0, n, c0, c1);
    if (c0 > 0.0 && c0 < 1.0 && c1 >= 0.0) return P_E2;
    if (a0 <= 0.0 && c0 >= 1.0) return P_T0;
    if (b0 <= 0.0 && a0 >= 1.0) return P_T1;
    if (c0 <= 0.0 && b0 >= 1.0) return P_T2;
    return P_T;
}
enum EE { EA0_EB0, EA0_EB1, EA1_EB0, EA1_EB1, EA_EB0, EA_EB1, EA0_EB, EA1_EB, EA_EB };
__device__ int ee_parallel_type(V ea0, V ea1, V eb0, V eb1)
{
    const V ea = ea1 - ea0;
    const double alpha = dotv(eb0 - ea0, ea) / sqn(ea);
    const double beta = dotv(eb1 - ea0, ea) / sqn(ea);
    int eac, ebc;
    if (alpha < 0) { eac = (0 <= beta && beta <= 1) ? 2 : 0; ebc = (beta <= alpha) ? 0 : (beta <= 1 ? 1 : 2); }
    else if (alpha > 1) { eac = (0 <= beta && beta <= 1) ? 2 : 1; ebc = (beta >= alpha) ? 0 : (0 <= beta ? 1 : 2); }
    else { eac = 2; ebc = 0; }
    return ebc < 2 ? (eac << 1 | ebc) : (6 + eac);
}
__device__ int ee_type(V ea0, V ea1, V eb0, V eb1, double tol)
{
    const V u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const double a = sqn(u), b = dotv(u, v), c = sqn(v), dd = dotv(u, w), e = dotv(v, w);
    const double D = a * c - b * b;
    if (a == 0.0 && c == 0.0) return EA0_EB0;
    else if (a == 0.0) return EA0_EB;
    else if (c == 0.0) return EA_EB0;
    if (sqn(crossv(u, v)) < tol) return ee_parallel_type(ea0, ea1, eb0, eb1);
    int def = EA_EB;
    const double sN = (b * e - c * dd);
    double tN, tD;
    if (sN <= 0.0) { tN = e; tD = c; def = EA0_EB; }
    else if (sN >= D) { tN = e + b; tD = c; def = EA1_EB; }
    else {
        tN = (a * e - b * dd); tD = D;
        if (tN > 0.0 && tN < tD && sqn(crossv(u, v)) < tol) {
            if (sN < D / 2) { tN = e; tD = c; def = EA0_EB; }
            else { tN = e + b; tD = c; def = EA1_EB; }
        }
    }
    if (tN <= 0.0) {
        if (-dd <= 0.0) return EA0_EB0;
        else if (-dd >= a) return EA1_EB0;
        else return EA_EB0;
    } else if (tN >= tD) {
        if ((-dd + b) <= 0.0) return EA0_EB1;
        else if ((-dd + b) >= a) return EA1_EB1;
        else return EA_EB1;
    }
    return def;
}
// :565-585
__device__ bool edge_hits_triangle(V Q1, V Q2, V A, V B, V C)
{
    const V E1 = B - A, E2 = C - A;
    const V N = crossv(E1, E2);
    const V Dir = Q2 - Q1;
    const double det = -dotv(Dir, N);
    const double invdet = 1.0 / det;
    const V AO = Q1 - A;
    const V DAO = crossv(AO, Dir);
    const double u = dotv(E2, DAO) * invdet;
    const double v = -dotv(E1, DAO) * invdet;
    const double t = dotv(AO, N) * invdet;
    return (fabs(det) >= 1e-14 && t >= 0.0 && t <= 1.0 && u >= 0.0 && v >= 0.0 && (u + v) <= 1.0);
}

// ---- friction geometry (S/models/interactions/friction_geometry.cpp) ----
__device__ void proj_point_point(V p, V a, double* T)
{
    const V n = normalizedv(p - a);
    const V e = (n.z < 0.99) ? V{0, 0, 1} : V{1, 0, 0};
    const V u = normalizedv(crossv(e, n));
    const V v = normalizedv(crossv(u, n));
    T[0] = u.x; T[1] = u.y; T[2] = u.z; T[3] = v.x; T[4] = v.y; T[5] = v.z;
}
__device__ void proj_point_edge(V p, V a, V b, double* T)
{
    const V u = normalizedv(b - a);
    const V v = normalizedv(crossv(u, p - a));
    T[0] = u.x; T[1] = u.y; T[2] = u.z; T[3] = v.x; T[4] = v.y; T[5] = v.z;
}
__device__ void proj_triangle(V a, V b, V c, double* T)
{
    const V v01 = a - c, v02 = b - c;
    const V u = normalizedv(v01);
    const V v = normalizedv(crossv(crossv(v01, v02), u));
    T[0] = u.x; T[1] = u.y; T[2] = u.z; T[3] = v.x; T[4] = v.y; T[5] = v.z;
}
__device__ void proj_edge_edge(V a, V b, V p, V q, double* T)
{
    const V u = normalizedv(b - a);
    const V n = crossv(u, q - p);
    const V v = normalizedv(crossv(u, n));
    T[0] = u.x; T[1] = u.y; T[2] = u.z; T[3] = v.x; T[4] = v.y; T[5] = v.z;
}
__device__ void bary_point_edge(V p, V a, V b, double* o)
{
    const V ab = b - a;
    const double alpha = dotv(p - a, ab) / sqn(ab);
    o[0] = 1.0 - alpha; o[1] = alpha;
}
__device__ void bary_point_triangle(V p, V a, V b, V c, double* o)
{
    const V v0 = b - a, v1 = c - a, v2 = p - a;
    const double d00 = dotv(v0, v0), d01 = dotv(v0, v1), d11 = dotv(v1, v1), d20 = dotv(v2, v0), d21 = dotv(v2, v1);
    const double inv = 1.0 / (d00 * d11 - d01 * d01);
    const double v = (d11 * d20 - d01 * d21) * inv, w = (d00 * d21 - d01 * d20) * inv;
    o[0] = 1.0 - v - w; o[1] = v; o[2] = w;
}
__device__ void bary_edge_edge(V A, V B, V P, V Q, double* o)
{
    const V da = B - A, db = Q - P, r = A - P;
    const double a = dotv(da, da), e = dotv(db, db), f = dotv(db, r), b = dotv(da, db), c = dotv(da, r);
    const double denom = a * e - b * b;
    if (denom < 1e-16) { o[0] = 0.5; o[1] = 0.5; return; }
    const double s = (b * f - c * e) / denom;
    o[0] = s; o[1] = (b * s + f) / e;
}

// ---------------------------------------------------------------------------------------------------
// list / table writers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int push_row(const Dev& d, int table, const int* row)
{
    const int slot = atomicAdd(d.counters + 16 + table, 1);
    if (slot >= d.table_cap) { d.counters[3] = 1; return -1; }
    int32_t* dst = d.table[table] + (size_t)slot * d.table_stride[table];
    unsigned long long x = 0x9E3779B97F4A7C15ull * (unsigned long long)(table + 1);
    for (int k = 0; k < d.table_stride[table]; k++) {
        dst[k] = row[k];
        x = (x ^ (unsigned long long)(unsigned)row[k]) * 0xBF58476D1CE4E5B9ull;
        x ^= x >> 29;
    }
    // digest of the table as a SET of rows (the append order depends on scheduling): sums of two independent mixes of every row
    unsigned long long h1 = x * 0x94D049BB133111EBull; h1 ^= h1 >> 31;
    unsigned long long h2 = (x ^ 0xD6E8FEB86659FD93ull) * 0xFF51AFD7ED558CCDull; h2 ^= h2 >> 33; h2 *= 0xC4CEB9FE1A85EC53ull; h2 ^= h2 >> 33;
    atomicAdd(d.hash + 2 * table, h1);
    atomicAdd(d.hash + 2 * table + 1, h2);
    return slot;
}
__device__ __forceinline__ void push_list(const Dev& d, int list, int width, const int* ids, double dist)
{
    const int slot = atomicAdd(d.counters + 8 + list, 1);
    if (slot >= d.list_cap) { d.counters[3] = 1; return; }
    for (int k = 0; k < width; k++) d.list_ids[list][(size_t)slot * width + k] = ids[k];
    d.list_dist[list][slot] = dist;
}
// friction rows carry their own index in column 0 (LabelledConnectivity::numbered_push_back)
__device__ __forceinline__ int push_friction(const Dev& d, int table, int* row, const double* T, const double* bary, int nbary, double mu, double fn)
{
    const int slot = atomicAdd(d.counters + 16 + table, 1);
    if (slot >= d.table_cap) { d.counters[3] = 1; return -1; }
    row[0] = slot;
    int32_t* dst = d.table[table] + (size_t)slot * d.table_stride[table];
    for (int k = 0; k < d.table_stride[table]; k++) dst[k] = row[k];
    const int f = table - N_CONTACT_TABLES;
    for (int k = 0; k < 6; k++) d.fT[f][6 * (size_t)slot + k] = T[k];
    for (int k = 0; k < nbary; k++) d.fbary[f][(size_t)nbary * slot + k] = bary[k];
    d.fmu[f][slot] = mu;
    d.ffn[f][slot] = fn;
    return slot;
}

struct Side { int group, ps, body; };
__device__ __forceinline__ Side side_of(const Dev& d, int g) { return {g, d.g_ps[g], d.g_body[g]}; }

// One point (A) - primitive of a triangle (B) proximity pair.  nB = number of B vertices (1 point, 2 edge, 3 triangle).
// mode 0: contact tables (EnergyFrictionalContact.cpp:381-455), mode 1: friction tables (:600-690)
__device__ void emit_pt_one(const Dev& d, int mode, double dist, double stiffness, int nB, int pA, const int* vB, Side A, Side B)
{
    if (mode > 1) return;   // raw lists only
    const double dhat = d.g_thickness[A.group] + d.g_thickness[B.group];
    if (dist > dhat) return;
    const int a = d.v_ps_index[pA];
    int b[3];
    for (int k = 0; k < nB; k++) b[k] = d.v_ps_index[vB[k]];
    const bool A_soft = (A.ps == 0), B_soft = (B.ps == 0);
    if (mode == 0) {
        int row[8];
        if (A_soft && B_soft) {
            row[0] = A.group; row[1] = B.group; row[2] = a;
            for (int k = 0; k < nB; k++) row[3 + k] = b[k];
            push_row(d, nB == 1 ? CT_DD_PT_PP : (nB == 2 ? CT_DD_PT_PE : CT_DD_PT_PT), row);
        } else if (!A_soft && !B_soft) {   // rigid point vs rigid primitive
            row[0] = A.group; row[1] = B.group; row[2] = A.body; row[3] = B.body; row[4] = a;
            for (int k = 0; k < nB; k++) row[5 + k] = b[k];
            push_row(d, nB == 1 ? CT_RR_PT_PP : (nB == 2 ? CT_RR_PT_PE : CT_RR_PT_PT), row);
        } else if (!A_soft) {   // rigid point vs deformable primitive
            row[0] = A.group; row[1] = B.group; row[2] = A.body; row[3] = a;
            for (int k = 0; k < nB; k++) row[4 + k] = b[k];
            push_row(d, nB == 1 ? CT_RD_PT_PP : (nB == 2 ? CT_RD_PT_PE : CT_RD_PT_PT), row);
        } else {                // deformable point vs rigid primitive
            row[0] = B.group; row[1] = A.group; row[2] = B.body;
            for (int k = 0; k < nB; k++) row[3 + k] = b[k];
            row[3 + nB] = a;
            push_row(d, nB == 1 ? CT_RD_PT_PP : (nB == 2 ? CT_RD_PT_EP : CT_RD_PT_TP), row);
        }
    } else {
        const double mu = d.mu[A.group * MAX_GROUPS + B.group];
        if (mu == 0.0) return;
        const double fn = stiffness * (dhat - dist) * (dhat - dist);   // _barrier_force, cubic (:1239-1243)
        const V P = ld(d.x + 3 * pA);
        double T[6], bary[3];
        int row[8];
        if (nB == 1) {
            proj_point_point(P, ld(d.x + 3 * vB[0]), T);
            if (A_soft && B_soft) { row[1] = a; row[2] = b[0]; push_friction(d, FT_DD_PP, row, T, bary, 0, mu, fn); }
            else if (!A_soft && !B_soft) { row[1] = A.body; row[2] = B.body; row[3] = a; row[4] = b[0]; push_friction(d, FT_RR_PP, row, T, bary, 0, mu, fn); }
            else if (!A_soft) { row[1] = A.body; row[2] = a; row[3] = b[0]; push_friction(d, FT_RD_PP, row, T, bary, 0, mu, fn); }
            else { row[1] = B.body; row[2] = b[0]; row[3] = a; push_friction(d, FT_RD_PP, row, T, bary, 0, mu, fn); }
        } else if (nB == 2) {
            const V E0 = ld(d.x + 3 * vB[0]), E1 = ld(d.x + 3 * vB[1]);
            bary_point_edge(P, E0, E1, bary);
            proj_point_edge(P, E0, E1, T);
            if (A_soft && B_soft) { row[1] = a; row[2] = b[0]; row[3] = b[1]; push_friction(d, FT_DD_PE, row, T, bary, 2, mu, fn); }
            else if (!A_soft && !B_soft) { row[1] = A.body; row[2] = B.body; row[3] = a; row[4] = b[0]; row[5] = b[1]; push_friction(d, FT_RR_PE, row, T, bary, 2, mu, fn); }
            else if (!A_soft) { row[1] = A.body; row[2] = a; row[3] = b[0]; row[4] = b[1]; push_friction(d, FT_RD_PE, row, T, bary, 2, mu, fn); }
            else { row[1] = B.body; row[2] = b[0]; row[3] = b[1]; row[4] = a; push_friction(d, FT_RD_EP, row, T, bary, 2, mu, fn); }
        } else {
            const V T0 = ld(d.x + 3 * vB[0]), T1 = ld(d.x + 3 * vB[1]), T2 = ld(d.x + 3 * vB[2]);
            bary_point_triangle(P, T0, T1, T2, bary);
            proj_triangle(T0, T1, T2, T);
            if (A_soft && B_soft) { row[1] = a; row[2] = b[0]; row[3] = b[1]; row[4] = b[2]; push_friction(d, FT_DD_PT, row, T, bary, 3, mu, fn); }
            else if (!A_soft && !B_soft) { row[1] = A.body; row[2] = B.body; row[3] = a; row[4] = b[0]; row[5] = b[1]; row[6] = b[2]; push_friction(d, FT_RR_PT, row, T, bary, 3, mu, fn); }
            else if (!A_soft) { row[1] = A.body; row[2] = a; row[3] = b[0]; row[4] = b[1]; row[5] = b[2]; push_friction(d, FT_RD_PT, row, T, bary, 3, mu, fn); }
            else { row[1] = B.body; row[2] = b[0]; row[3] = b[1]; row[4] = b[2]; row[5] = a; push_friction(d, FT_RD_TP, row, T, bary, 3, mu, fn); }
        }
    }
}

__device__ __forceinline__ void emit_pt(const Dev& d, int mode, double dist, double stiffness, int nB, int pA, const int* vB, Side A, Side B);
__device__ void narrow_pt(const Dev& d, double enl_sq, int mode, double stiffness)
{
    const int total = min(d.counters[0], d.cand_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int p = d.cand_pt[i].x, t = d.cand_pt[i].y;
        const int v0 = d.tri[3 * t], v1 = d.tri[3 * t + 1], v2 = d.tri[3 * t + 2];
        const V P = ld(d.x + 3 * p), A = ld(d.x + 3 * v0), B = ld(d.x + 3 * v1), C = ld(d.x + 3 * v2);
        const int type = pt_type(P, A, B, C);
        double d2;
        switch (type) {
        case P_T0: d2 = pp_sq(P, A); break;
        case P_T1: d2 = pp_sq(P, B); break;
        case P_T2: d2 = pp_sq(P, C); break;
        case P_E0: d2 = pl_sq(P, A, B); break;
        case P_E1: d2 = pl_sq(P, B, C); break;
        case P_E2: d2 = pl_sq(P, C, A); break;
        default: d2 = ppl_sq(P, A, B, C); break;
        }
        if (!(d2 < enl_sq)) continue;
        const double dist = sqrt(d2);
        const int gp = d.v_group[p], gt = d.t_group[t];
        const int lp = p - d.g_voff[gp], lt = t - d.g_toff[gt], tvo = d.g_voff[gt];
        const Side SA = side_of(d, gp), SB = side_of(d, gt);
        int ids[6], vB[3];
        if (type <= P_T2) {
            vB[0] = (type == P_T0) ? v0 : (type == P_T1 ? v1 : v2);
            ids[0] = gp; ids[1] = lp; ids[2] = gt; ids[3] = lt; ids[4] = vB[0] - tvo;
            push_list(d, 0, 5, ids, dist);
            emit_pt(d, mode, dist, stiffness, 1, p, vB, SA, SB);
        } else if (type <= P_E2) {
            vB[0] = (type == P_E0) ? v0 : (type == P_E1 ? v1 : v2);
            vB[1] = (type == P_E0) ? v1 : (type == P_E1 ? v2 : v0);
            ids[0] = gp; ids[1] = lp; ids[2] = gt; ids[3] = lt; ids[4] = vB[0] - tvo; ids[5] = vB[1] - tvo;
            push_list(d, 1, 6, ids, dist);
            emit_pt(d, mode, dist, stiffness, 2, p, vB, SA, SB);
        } else {
            vB[0] = v0; vB[1] = v1; vB[2] = v2;
            ids[0] = gp; ids[1] = lp; ids[2] = gt; ids[3] = lt;
            push_list(d, 2, 4, ids, dist);
            emit_pt(d, mode, dist, stiffness, 3, p, vB, SA, SB);
        }
    }
}

// Edge-edge derived pairs.  kind 0: point(A edge-point) - point(B edge-point); 1: edge-point(A) - edge(B); 2: edge(A) - edge(B)
// (EnergyFrictionalContact.cpp:457-529 contact, :692-772 friction).  eA / eB: the full edges (for the mollifier tables).
// mode 3: contact AND friction tables from one detection (start of a time step: both are built at the same positions)
__device__ __forceinline__ void emit_pt(const Dev& d, int mode, double dist, double stiffness, int nB, int pA, const int* vB, Side A, Side B)
{
    if (mode == 3) { emit_pt_one(d, 0, dist, stiffness, nB, pA, vB, A, B); emit_pt_one(d, 1, dist, stiffness, nB, pA, vB, A, B); }
    else emit_pt_one(d, mode, dist, stiffness, nB, pA, vB, A, B);
}
__device__ void emit_ee_one(const Dev& d, int mode, double dist, double stiffness, int kind, const int* eA, int pA, const int* eB, int pB, Side A, Side B)
{
    if (mode > 1) return;   // raw lists only
    const double dhat = d.g_thickness[A.group] + d.g_thickness[B.group];
    if (dist > dhat) return;
    const bool A_soft = (A.ps == 0), B_soft = (B.ps == 0);
    const int a0 = d.v_ps_index[eA[0]], a1 = d.v_ps_index[eA[1]], b0 = d.v_ps_index[eB[0]], b1 = d.v_ps_index[eB[1]];
    const int ap = (pA >= 0) ? d.v_ps_index[pA] : -1, bp = (pB >= 0) ? d.v_ps_index[pB] : -1;
    if (mode == 0) {
        int row[10];
        if (kind == 0) {
            if (A_soft && B_soft) { const int r[8] = {A.group, B.group, a0, a1, ap, b0, b1, bp}; push_row(d, CT_DD_EE_PP, r); }
            else if (!A_soft && !B_soft) { const int r[10] = {A.group, B.group, A.body, B.body, a0, a1, ap, b0, b1, bp}; push_row(d, CT_RR_EE_PP, r); }
            else if (!A_soft) { const int r[9] = {A.group, B.group, A.body, a0, a1, ap, b0, b1, bp}; push_row(d, CT_RD_EE_PP, r); }
            else { const int r[9] = {B.group, A.group, B.body, b0, b1, bp, a0, a1, ap}; push_row(d, CT_RD_EE_PP, r); }
        } else if (kind == 1) {
            if (A_soft && B_soft) { const int r[7] = {A.group, B.group, a0, a1, ap, b0, b1}; push_row(d, CT_DD_EE_PE, r); }
            else if (!A_soft && !B_soft) { const int r[9] = {A.group, B.group, A.body, B.body, a0, a1, ap, b0, b1}; push_row(d, CT_RR_EE_PE, r); }
            else if (!A_soft) { const int r[8] = {A.group, B.group, A.body, a0, a1, ap, b0, b1}; push_row(d, CT_RD_EE_PE, r); }
            else { const int r[8] = {B.group, A.group, B.body, b0, b1, a0, a1, ap}; push_row(d, CT_RD_EE_EP, r); }
        } else {
            if (A_soft && B_soft) { const int r[6] = {A.group, B.group, a0, a1, b0, b1}; push_row(d, CT_DD_EE_EE, r); }
            else if (!A_soft && !B_soft) { const int r[8] = {A.group, B.group, A.body, B.body, a0, a1, b0, b1}; push_row(d, CT_RR_EE_EE, r); }
            else if (!A_soft) { const int r[7] = {A.group, B.group, A.body, a0, a1, b0, b1}; push_row(d, CT_RD_EE_EE, r); }
            else { const int r[7] = {B.group, A.group, B.body, b0, b1, a0, a1}; push_row(d, CT_RD_EE_EE, r); }
        }
        (void)row;
    } else {
        const double mu = d.mu[A.group * MAX_GROUPS + B.group];
        if (mu == 0.0) return;
        const double fn = stiffness * (dhat - dist) * (dhat - dist);
        double T[6], bary[3];
        int row[8];
        if (kind == 0) {
            proj_point_point(ld(d.x + 3 * pA), ld(d.x + 3 * pB), T);
            if (A_soft && B_soft) { row[1] = ap; row[2] = bp; push_friction(d, FT_DD_PP, row, T, bary, 0, mu, fn); }
            else if (!A_soft && !B_soft) { row[1] = A.body; row[2] = B.body; row[3] = ap; row[4] = bp; push_friction(d, FT_RR_PP, row, T, bary, 0, mu, fn); }
            else if (!A_soft) { row[1] = A.body; row[2] = ap; row[3] = bp; push_friction(d, FT_RD_PP, row, T, bary, 0, mu, fn); }
            else { row[1] = B.body; row[2] = bp; row[3] = ap; push_friction(d, FT_RD_PP, row, T, bary, 0, mu, fn); }
        } else if (kind == 1) {
            const V P = ld(d.x + 3 * pA), E0 = ld(d.x + 3 * eB[0]), E1 = ld(d.x + 3 * eB[1]);
            bary_point_edge(P, E0, E1, bary);
            proj_point_edge(P, E0, E1, T);
            if (A_soft && B_soft) { row[1] = ap; row[2] = b0; row[3] = b1; push_friction(d, FT_DD_PE, row, T, bary, 2, mu, fn); }
            else if (!A_soft && !B_soft) { row[1] = A.body; row[2] = B.body; row[3] = ap; row[4] = b0; row[5] = b1; push_friction(d, FT_RR_PE, row, T, bary, 2, mu, fn); }
            else if (!A_soft) { row[1] = A.body; row[2] = ap; row[3] = b0; row[4] = b1; push_friction(d, FT_RD_PE, row, T, bary, 2, mu, fn); }
            else { row[1] = B.body; row[2] = b0; row[3] = b1; row[4] = ap; push_friction(d, FT_RD_EP, row, T, bary, 2, mu, fn); }
        } else {
            const V EA0 = ld(d.x + 3 * eA[0]), EA1 = ld(d.x + 3 * eA[1]), EB0 = ld(d.x + 3 * eB[0]), EB1 = ld(d.x + 3 * eB[1]);
            bary_edge_edge(EA0, EA1, EB0, EB1, bary);     // in detection order, like the reference (even when the table swaps A and B)
            proj_edge_edge(EA0, EA1, EB0, EB1, T);
            if (A_soft && B_soft) { row[1] = a0; row[2] = a1; row[3] = b0; row[4] = b1; push_friction(d, FT_DD_EE, row, T, bary, 2, mu, fn); }
            else if (!A_soft && !B_soft) { row[1] = A.body; row[2] = B.body; row[3] = a0; row[4] = a1; row[5] = b0; row[6] = b1; push_friction(d, FT_RR_EE, row, T, bary, 2, mu, fn); }
            else if (!A_soft) { row[1] = A.body; row[2] = a0; row[3] = a1; row[4] = b0; row[5] = b1; push_friction(d, FT_RD_EE, row, T, bary, 2, mu, fn); }
            else { row[1] = B.body; row[2] = b0; row[3] = b1; row[4] = a0; row[5] = a1; push_friction(d, FT_RD_EE, row, T, bary, 2, mu, fn); }
        }
    }
}

__device__ __forceinline__ void emit_ee(const Dev& d, int mode, double dist, double stiffness, int kind, const int* eA, int pA, const int* eB, int pB, Side A, Side B)
{
    if (mode == 3) { emit_ee_one(d, 0, dist, stiffness, kind, eA, pA, eB, pB, A, B); emit_ee_one(d, 1, dist, stiffness, kind, eA, pA, eB, pB, A, B); }
    else emit_ee_one(d, mode, dist, stiffness, kind, eA, pA, eB, pB, A, B);
}
__device__ void narrow_ee(const Dev& d, double enl_sq, int mode, double stiffness, double parallel_tol)
{
    const int total = min(d.counters[1], d.cand_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ea = d.cand_ee[i].x, eb = d.cand_ee[i].y;
        const int eA[2] = {d.edge[2 * ea], d.edge[2 * ea + 1]}, eB[2] = {d.edge[2 * eb], d.edge[2 * eb + 1]};
        const V a = ld(d.x + 3 * eA[0]), b = ld(d.x + 3 * eA[1]), p = ld(d.x + 3 * eB[0]), q = ld(d.x + 3 * eB[1]);
        const int type = ee_type(a, b, p, q, parallel_tol);
        double d2;
        switch (type) {
        case EA0_EB0: d2 = pp_sq(a, p); break;
        case EA0_EB1: d2 = pp_sq(a, q); break;
        case EA1_EB0: d2 = pp_sq(b, p); break;
        case EA1_EB1: d2 = pp_sq(b, q); break;
        case EA_EB0: d2 = pl_sq(p, a, b); break;
        case EA_EB1: d2 = pl_sq(q, a, b); break;
        case EA0_EB: d2 = pl_sq(a, p, q); break;
        case EA1_EB: d2 = pl_sq(b, p, q); break;
        default: d2 = ll_sq(a, b, p, q); break;
        }
        if (!(d2 < enl_sq)) continue;
        if (sqn(crossv(b - a, q - p)) <= parallel_tol) continue;
        const double dist = sqrt(d2);
        const int ga = d.e_group[ea], gb = d.e_group[eb];
        const int la = ea - d.g_eoff[ga], lb = eb - d.g_eoff[gb], vao = d.g_voff[ga], vbo = d.g_voff[gb];
        const Side SA = side_of(d, ga), SB = side_of(d, gb);
        int ids[6];
        if (type <= EA1_EB1) {
            const int pa = (type == EA0_EB0 || type == EA0_EB1) ? eA[0] : eA[1];
            const int pb = (type == EA0_EB0 || type == EA1_EB0) ? eB[0] : eB[1];
            ids[0] = ga; ids[1] = la; ids[2] = pa - vao; ids[3] = gb; ids[4] = lb; ids[5] = pb - vbo;
            push_list(d, 3, 6, ids, dist);
            emit_ee(d, mode, dist, stiffness, 0, eA, pa, eB, pb, SA, SB);
        } else if (type == EA_EB0 || type == EA_EB1) {   // point of B vs edge A: the edge-point is on B
            const int pb = (type == EA_EB0) ? eB[0] : eB[1];
            ids[0] = gb; ids[1] = lb; ids[2] = pb - vbo; ids[3] = ga; ids[4] = la;
            push_list(d, 4, 5, ids, dist);
            emit_ee(d, mode, dist, stiffness, 1, eB, pb, eA, -1, SB, SA);
        } else if (type == EA0_EB || type == EA1_EB) {
            const int pa = (type == EA0_EB) ? eA[0] : eA[1];
            ids[0] = ga; ids[1] = la; ids[2] = pa - vao; ids[3] = gb; ids[4] = lb;
            push_list(d, 4, 5, ids, dist);
            emit_ee(d, mode, dist, stiffness, 1, eA, pa, eB, -1, SA, SB);
        } else {
            ids[0] = ga; ids[1] = la; ids[2] = gb; ids[3] = lb;
            push_list(d, 5, 4, ids, dist);
            emit_ee(d, mode, dist, stiffness, 2, eA, -1, eB, -1, SA, SB);
        }
    }
}

__global__ void __launch_bounds__(128) k_narrow_et(Dev d)
{
    const int total = min(d.counters[2], d.cand_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = d.cand_et[i].x, t = d.cand_et[i].y;
        const V e0 = ld(d.x + 3 * d.edge[2 * e]), e1 = ld(d.x + 3 * d.edge[2 * e + 1]);
        const V t0 = ld(d.x + 3 * d.tri[3 * t]), t1 = ld(d.x + 3 * d.tri[3 * t + 1]), t2 = ld(d.x + 3 * d.tri[3 * t + 2]);
        if (edge_hits_triangle(e0, e1, t0, t1, t2)) {
            const int ge = d.e_group[e], gt = d.t_group[t];
            const int ids[4] = {ge, e - d.g_eoff[ge], gt, t - d.g_toff[gt]};
            push_list(d, 6, 4, ids, 0.0);
        }
    }
}

// point-triangle and edge-edge narrow phases of one detection in one launch
__global__ void __launch_bounds__(128) k_narrow_all(Dev d, double enl_sq, int mode, double stiffness, double parallel_tol, int do_pt, int do_ee)
{
    // (the tables' digests go to global memory with L2 reductions: a shared-memory stage was tried and is slower -- 64-bit shared
    //  atomics are compare-and-swap loops, and every point-triangle row of a warp hits the same word)
    if (do_pt) narrow_pt(d, enl_sq, mode, stiffness);
    if (do_ee) narrow_ee(d, enl_sq, mode, stiffness, parallel_tol);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
void contact_destroy(sb_context* ctx)
{
    Contact* C = ctx->contact;
    if (!C) return;
    C->x.release(); C->v_group.release(); C->v_ps.release(); C->tri.release(); C->t_group.release(); C->edge.release(); C->e_group.release();
    C->g_i32.release(); C->g_f64.release(); C->mu.release(); C->blacklist.release(); C->bb_p.release(); C->bb_t.release(); C->bb_e.release(); C->tb_p.release(); C->tb_t.release(); C->tb_e.release();
    C->bb0_p.release(); C->bb0_t.release(); C->bb0_e.release(); C->tb0_p.release(); C->tb0_t.release(); C->tb0_e.release();
    C->perm_p.release(); C->perm_t.release(); C->perm_e.release(); C->perm_tmp.release(); C->morton.release(); C->morton_tmp.release(); C->bounds.release(); C->sort_temp.release();
    for (int k = 0; k < 3; k++) C->tile_pairs[k].release();
    C->cand_pt.release(); C->cand_ee.release(); C->cand_et.release(); C->counters.release();
    for (int l = 0; l < N_LISTS; l++) { C->list_ids[l].release(); C->list_dist[l].release(); }
    for (int t = 0; t < N_TABLES; t++) C->table[t].release();
    for (int t = 0; t < N_CONTACT_TABLES; t++) C->table_back[t].release();
    C->hash.release();
    if (std::getenv("SB_CONTACT_DUMP")) fprintf(stderr, "[stark_b200 contact] detections that rebuilt the contact tables: %lld unchanged as sets (tables kept), %lld changed\n", C->n_tables_same, C->n_tables_changed);
    if (C->h_counters) cudaFreeHost(C->h_counters);
    if (C->h_hash) cudaFreeHost(C->h_hash);
    delete C;
    ctx->contact = nullptr;
}
bool contact_active(sb_context* ctx) { return ctx->contact && !ctx->contact->groups.empty(); }

static int new_array(sb_context* ctx, const char* label, int stride)
{
    Array a;
    a.label = label;
    a.stride = stride;
    ctx->arrays.push_back(a);
    return (int)ctx->arrays.size() - 1;
}

static int upload_topology(sb_context* ctx, Contact* C)
{
    cudaStream_t st = ctx->stream;
    const int nv = (int)C->h_v_group.size(), nt = (int)C->h_t_group.size(), ne = (int)C->h_e_group.size();
    auto up = [&](auto& buf, const auto& h) {
        buf.ensure(h.size() + 1);
        if (!h.empty()) cudaMemcpyAsync(buf.p, h.data(), h.size() * sizeof(h[0]), cudaMemcpyHostToDevice, st);
    };
    up(C->v_group, C->h_v_group); up(C->v_ps, C->h_v_ps); up(C->tri, C->h_tri); up(C->t_group, C->h_t_group); up(C->edge, C->h_edge); up(C->e_group, C->h_e_group);
    up(C->blacklist, C->h_blacklist); up(C->mu, C->h_mu);
    const int G = (int)C->groups.size();
    std::vector<int32_t> gi(5 * MAX_GROUPS, 0);
    std::vector<double> gf(MAX_GROUPS, 0.0);
    for (int g = 0; g < G; g++) {
        gi[g] = C->groups[g].ps; gi[MAX_GROUPS + g] = C->groups[g].body; gi[2 * MAX_GROUPS + g] = C->groups[g].v_off;
        gi[3 * MAX_GROUPS + g] = C->groups[g].t_off; gi[4 * MAX_GROUPS + g] = C->groups[g].e_off;
        gf[g] = C->groups[g].thickness;
    }
    up(C->g_i32, gi); up(C->g_f64, gf);
    C->x.ensure(3 * (size_t)nv + 3);
    C->bb_p.ensure(6 * (size_t)nv + 6); C->bb_t.ensure(6 * (size_t)nt + 6); C->bb_e.ensure(6 * (size_t)ne + 6);
    C->bb0_p.ensure(6 * (size_t)nv + 6); C->bb0_t.ensure(6 * (size_t)nt + 6); C->bb0_e.ensure(6 * (size_t)ne + 6);
    C->perm_p.ensure(nv + 1); C->perm_t.ensure(nt + 1); C->perm_e.ensure(ne + 1);
    if (nv) k_iota<<<(nv + 255) / 256, 256, 0, st>>>(C->perm_p.p, nv);
    if (nt) k_iota<<<(nt + 255) / 256, 256, 0, st>>>(C->perm_t.p, nt);
    if (ne) k_iota<<<(ne + 255) / 256, 256, 0, st>>>(C->perm_e.p, ne);
    C->reorder_countdown = 0;
    SB_CUDA(ctx, cudaStreamSynchronize(st));
    // module-owned arrays seen by the potentials
    Array& th = ctx->arrays[C->a_thickness];
    th.d.ensure(MAX_GROUPS); th.n_rows = G;
    SB_CUDA(ctx, cudaMemcpy(th.d.p, gf.data(), G * sizeof(double), cudaMemcpyHostToDevice));
    Array& rl = ctx->arrays[C->a_rb_local];
    rl.d.ensure(C->h_rb_local.size() + 3); rl.n_rows = (int)C->h_rb_local.size() / 3;
    if (!C->h_rb_local.empty()) SB_CUDA(ctx, cudaMemcpy(rl.d.p, C->h_rb_local.data(), C->h_rb_local.size() * sizeof(double), cudaMemcpyHostToDevice));
    C->topology_dirty = false;
    return 0;
}

static void ensure_capacities(sb_context* ctx, Contact* C)
{
    C->cand_pt.ensure(C->cand_cap); C->cand_ee.ensure(C->cand_cap); C->cand_et.ensure(C->cand_cap);
    C->counters.ensure(64);
    {   // tile boxes and the (never overflowing) lists of overlapping tile pairs
        const size_t nv = C->h_v_group.size(), nt = C->h_t_group.size(), ne = C->h_e_group.size();
        const size_t Tv = (nv + TILE - 1) / TILE, Tt = (nt + TILE - 1) / TILE, Te = (ne + TILE - 1) / TILE;
        C->tb_p.ensure(6 * Tv + 6); C->tb_t.ensure(6 * Tt + 6); C->tb_e.ensure(6 * Te + 6);
        C->tb0_p.ensure(6 * Tv + 6); C->tb0_t.ensure(6 * Tt + 6); C->tb0_e.ensure(6 * Te + 6);
        C->tile_pairs[0].ensure(Tv * Tt + 1); C->tile_pairs[1].ensure(Te * Te + 1); C->tile_pairs[2].ensure(Te * Tt + 1);
    }
    for (int l = 0; l < N_LISTS; l++) { C->list_ids[l].ensure((size_t)C->list_cap * LIST_WIDTH[l]); C->list_dist[l].ensure(C->list_cap); }
    // Contact and friction tables share one capacity, but a detection rewrites only one of the two families (friction tables
    // are built once per time step): a table that is not being rewritten must survive a growth caused by the other family,
    // so every reallocation keeps the rows in use, and the potentials are re-pointed at the new buffers.
    bool moved = false;
    for (int t = 0; t < N_TABLES; t++) {
        const int32_t* before = C->table[t].p;
        const size_t stride = LAYOUTS[t].conn_stride;
        C->table[t].ensure_keep((size_t)C->table_cap * stride, (size_t)C->h_table_count[t] * stride, ctx->stream);
        moved = moved || before != C->table[t].p;
    }
    for (int f = 0; f < N_FRICTION; f++) {
        const size_t n = (size_t)C->h_table_count[N_CONTACT_TABLES + f];
        ctx->arrays[C->a_fT[f]].d.ensure_keep(6 * (size_t)C->table_cap, 6 * n, ctx->stream);
        ctx->arrays[C->a_fmu[f]].d.ensure_keep(C->table_cap, n, ctx->stream);
        ctx->arrays[C->a_ffn[f]].d.ensure_keep(C->table_cap, n, ctx->stream);
        if (C->a_fbary[f] >= 0) { const size_t st = (size_t)ctx->arrays[C->a_fbary[f]].stride; ctx->arrays[C->a_fbary[f]].d.ensure_keep(st * C->table_cap, st * n, ctx->stream); }
    }
    if (moved) for (int t = 0; t < N_TABLES; t++) if (C->pot[t] >= 0) ctx->potentials[C->pot[t]].conn_ext = C->table[t].p;
    for (int t = 0; t < N_CONTACT_TABLES; t++) C->table_back[t].ensure((size_t)C->table_cap * LAYOUTS[t].conn_stride);
    C->hash.ensure(2 * N_CONTACT_TABLES);
}

static Dev make_dev(sb_context* ctx, Contact* C, bool back_tables = false)
{
    Dev d;
    d.n_v = (int)C->h_v_group.size(); d.n_t = (int)C->h_t_group.size(); d.n_e = (int)C->h_e_group.size(); d.n_groups = (int)C->groups.size();
    d.x = C->x.p; d.v_group = C->v_group.p; d.v_ps_index = C->v_ps.p; d.tri = C->tri.p; d.t_group = C->t_group.p; d.edge = C->edge.p; d.e_group = C->e_group.p;
    d.bb_p = C->bb_p.p; d.bb_t = C->bb_t.p; d.bb_e = C->bb_e.p;
    d.tb_p = C->tb_p.p; d.tb_t = C->tb_t.p; d.tb_e = C->tb_e.p;
    d.perm_p = C->perm_p.p; d.perm_t = C->perm_t.p; d.perm_e = C->perm_e.p;
    for (int k = 0; k < 3; k++) { d.tile_pairs[k] = C->tile_pairs[k].p; d.tile_pair_cap[k] = (int)C->tile_pairs[k].cap; }
    d.g_ps = C->g_i32.p; d.g_body = C->g_i32.p + MAX_GROUPS; d.g_voff = C->g_i32.p + 2 * MAX_GROUPS; d.g_toff = C->g_i32.p + 3 * MAX_GROUPS; d.g_eoff = C->g_i32.p + 4 * MAX_GROUPS;
    d.g_thickness = C->g_f64.p; d.blacklist = C->blacklist.p; d.mu = C->mu.p;
    d.cand_pt = C->cand_pt.p; d.cand_ee = C->cand_ee.p; d.cand_et = C->cand_et.p; d.cand_cap = C->cand_cap;
    d.counters = C->counters.p;
    for (int l = 0; l < N_LISTS; l++) { d.list_ids[l] = C->list_ids[l].p; d.list_dist[l] = C->list_dist[l].p; }
    d.list_cap = C->list_cap;
    for (int t = 0; t < N_TABLES; t++) { d.table[t] = (back_tables && t < N_CONTACT_TABLES) ? C->table_back[t].p : C->table[t].p; d.table_stride[t] = LAYOUTS[t].conn_stride; }
    d.table_cap = C->table_cap;
    d.hash = C->hash.p;
    for (int f = 0; f < N_FRICTION; f++) {
        d.fT[f] = ctx->arrays[C->a_fT[f]].d.p; d.fmu[f] = ctx->arrays[C->a_fmu[f]].d.p; d.ffn[f] = ctx->arrays[C->a_ffn[f]].d.p;
        d.fbary[f] = (C->a_fbary[f] >= 0) ? ctx->arrays[C->a_fbary[f]].d.p : nullptr;
    }
    return d;
}

static int update_vertices(sb_context* ctx, Contact* C, bool zero_dt)
{
    if (C->external_vertices) return 0;
    Dev d = make_dev(ctx, C);
    const sb_contact_bindings& b = C->bind;
    auto ptr = [&](int a) -> const double* { return (a >= 0) ? ctx->arrays[a].d.p : nullptr; };
    k_update_vertices<<<(d.n_v + 255) / 256, 256, 0, ctx->stream>>>(d, ptr(b.soft_v1), ptr(b.soft_x0), ptr(b.rb_v1), ptr(b.rb_w1), ptr(b.rb_t0), ptr(b.rb_q0),
                                                                      ctx->arrays[C->a_rb_local].d.p, ptr(b.dt), zero_dt ? 1 : 0);
    ctx->launches++;
    return 0;
}

// Refresh the spatial (Morton) order of points, triangles and edges from the current vertex positions: the broad phase
// culls whole 256-slot tiles, which only works when consecutive slots are neighbours in space.  Run at the first detection
// after a topology change and every REORDER_PERIOD detections (deformation moves primitives slowly); ~10 small launches.
constexpr int REORDER_PERIOD = 64;
static int reorder_primitives(sb_context* ctx, Contact* C)
{
    cudaStream_t st = ctx->stream;
    ensure_capacities(ctx, C);
    Dev d = make_dev(ctx, C);
    const int nmax = std::max(d.n_v, std::max(d.n_t, d.n_e));
    if (nmax == 0) return 0;
    C->bounds.ensure(8); C->morton.ensure(nmax + 1); C->morton_tmp.ensure(nmax + 1); C->perm_tmp.ensure(nmax + 1);
    const int Tv = (d.n_v + TILE - 1) / TILE, Tt = (d.n_t + TILE - 1) / TILE, Te = (d.n_e + TILE - 1) / TILE;
    k_aabbs<<<(nmax + 255) / 256, 256, 0, st>>>(d, 0.0f, 1);                 // boxes in the CURRENT slot order
    if (d.n_v) k_tile_boxes<<<Tv, TILE, 0, st>>>(d.bb_p, d.tb_p, d.n_v);
    if (d.n_t) k_tile_boxes<<<Tt, TILE, 0, st>>>(d.bb_t, d.tb_t, d.n_t);
    if (d.n_e) k_tile_boxes<<<Te, TILE, 0, st>>>(d.bb_e, d.tb_e, d.n_e);
    if (d.n_v) k_scene_bounds<<<1, 256, 0, st>>>(d.tb_p, Tv, C->bounds.p, 0);  // every edge / triangle vertex is a point
    ctx->launches += 5;
    struct Cls { int n; const float* bb; DevBuf<int32_t>* perm; } cls[3] = {{d.n_v, d.bb_p, &C->perm_p}, {d.n_t, d.bb_t, &C->perm_t}, {d.n_e, d.bb_e, &C->perm_e}};
    for (auto& c : cls) {
        if (c.n == 0) continue;
        k_morton<<<(c.n + 255) / 256, 256, 0, st>>>(c.bb, c.perm->p, C->bounds.p, C->morton.p, C->perm_tmp.p, c.n);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, C->morton.p, C->morton_tmp.p, C->perm_tmp.p, c.perm->p, c.n, 0, 31, st);
        C->sort_temp.ensure(tb + 16);
        tb = C->sort_temp.cap;
        SB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(C->sort_temp.p, tb, C->morton.p, C->morton_tmp.p, C->perm_tmp.p, c.perm->p, c.n, 0, 31, st));
        ctx->launches += 5;
    }
    SB_CUDA(ctx, cudaGetLastError());
    C->reorder_countdown = REORDER_PERIOD;
    return 0;
}

// mode 0: proximity + contact tables; 1: proximity + friction tables; 2: intersections only; 3: proximity lists only + intersections;
// 4: proximity + contact tables + intersections (a line-search trial: validity test and the tables of the evaluation that follows)
// One detection = issue (all launches + the read-back of the counters, no synchronisation) -> [host sync] -> overflow check
// (grow + run again) -> publish (table sizes to the potentials).  The three steps are separate so that sb_newton_solve can queue
// the evaluation of the contact potentials BEHIND an issued detection and synchronise once for both (core.cu, fused path).
// mode 0: proximity + contact tables; 1: proximity + friction tables; 2: intersections only; 3: proximity lists only + intersections;
// 4: proximity + contact tables + intersections (a line-search trial: validity test and the tables of the evaluation that follows);
// 5: 4 + friction tables (start of a time step)
static int detect_issue(sb_context* ctx, Contact* C, int mode, double enlargement, bool defer_readback = false)
{
    cudaStream_t st = ctx->stream;
    ensure_capacities(ctx, C);
    Dev d = make_dev(ctx, C, mode == 0 || mode == 4 || mode == 5);   // (contact tables go to the back buffers, see Contact::table_back)
    Dev d0 = d;   // the intersection pass's view: its own boxes
    d0.bb_p = C->bb0_p.p; d0.bb_t = C->bb0_t.p; d0.bb_e = C->bb0_e.p; d0.tb_p = C->tb0_p.p; d0.tb_t = C->tb0_t.p; d0.tb_e = C->tb0_e.p;
    const bool both = (mode == 3 || mode == 4 || mode == 5);
    const int nmax = std::max(d.n_v, std::max(d.n_t, d.n_e));
    // only the counters this mode rewrites are cleared (contact and friction tables live side by side): bit t = counters[t]
    auto bits = [](int lo, int n) { unsigned long long m = 0; for (int t = lo; t < lo + n; t++) m |= 1ull << t; return m; };
    unsigned long long clear = bits(0, 8);
    if (mode == 0 || mode == 3 || mode == 4 || mode == 5) clear |= bits(8, 6) | bits(16, N_CONTACT_TABLES);
    if (mode == 1 || mode == 5) clear |= bits(8, 6) | bits(16 + N_CONTACT_TABLES, N_FRICTION);
    if (mode == 2 || mode == 3 || mode == 4 || mode == 5) clear |= bits(8 + 6, 1);
    const int Tv = (d.n_v + TILE - 1) / TILE, Tt = (d.n_t + TILE - 1) / TILE, Te = (d.n_e + TILE - 1) / TILE;
    auto broad_grid = [](long long n_tile_pairs) { return (int)std::max(1ll, std::min(n_tile_pairs, 148ll * 32)); };
    (void)nmax;
    if (mode != 2) {
        const float extra = (float)enlargement + FLT_EPSILON;
        const bool pt = C->enable_pt && d.n_t > 0 && d.n_v > 0, ee = C->enable_ee && d.n_e > 1;
        const int n0 = pt ? Tv * Tt : 0, n1 = ee ? Te * Te : 0;
        timeline_point(st, "detect: begin");
        k_aabbs_tiles<<<Tv + Tt + Te, TILE, 0, st>>>(d, extra, 1, Tv, Tt, clear);
        timeline_point(st, "detect: aabbs");
        if (both) SB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));   // (the counters are cleared: the intersection pass may start)
        if (n0 + n1 > 0) {
            k_tile_pairs_all<<<(n0 + n1 + 255) / 256, 256, 0, st>>>(d, Tv, Tt, Te, n0, n1, 0);
            if (pt && ee) k_broad_all<0, 1><<<broad_grid((long long)n0 + n1), TILE, 0, st>>>(d);
            else if (pt) k_broad_all<0, -1><<<broad_grid(n0), TILE, 0, st>>>(d);
            else k_broad_all<1, -1><<<broad_grid(n1), TILE, 0, st>>>(d);
            const int emit_mode = (mode == 3) ? 2 : (mode == 4 ? 0 : (mode == 5 ? 3 : mode));   // 2 = lists only (no table matches mode 2 inside emit_*), 3 = contact + friction tables
            timeline_point(st, "detect: broad");
            k_narrow_all<<<148 * 2, 128, 0, st>>>(d, enlargement * enlargement, emit_mode, C->stiffness, 1e-30, pt ? 1 : 0, ee ? 1 : 0);
            timeline_point(st, "detect: narrow");
            ctx->launches += 3;
        }
        ctx->launches += 1;
        clear = 0;   // (mode 3: the intersection pass below must not wipe what the proximity pass just counted)
    }
    if (mode == 2 || mode == 3 || mode == 4 || mode == 5) {
        // The intersection pass has its own boxes, tile-pair list, candidate list and counters: next to a proximity pass it
        // runs on a side stream at the same time (both are chains of latency-bound kernels).
        const float extra = 0.0f + FLT_EPSILON;   // IntersectionDetection uses non-enlarged AABBs (tmcd/BroadPhaseET.cpp:38)
        cudaStream_t s2 = both ? ctx->side[0] : st;
        if (both) SB_CUDA(ctx, cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
        k_aabbs_tiles<<<Tv + Tt + Te, TILE, 0, s2>>>(d0, extra, 0, Tv, Tt, clear);
        ctx->launches++;
        if (d.n_e > 0 && d.n_t > 0) {
            k_tile_pairs_all<<<(Te * Tt + 255) / 256, 256, 0, s2>>>(d0, Tv, Tt, Te, 0, Te * Tt, 1);
            k_broad_all<2, -1><<<broad_grid((long long)Te * Tt), TILE, 0, s2>>>(d0);
            ctx->launches += 2;
        }
        timeline_point(s2, "detect: et broad");
        k_narrow_et<<<148, 128, 0, s2>>>(d0);
        timeline_point(s2, "detect: et narrow");
        ctx->launches++;
        if (both) {
            SB_CUDA(ctx, cudaEventRecord(ctx->ev_join[0], s2));
            SB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[0], 0));
        }
    }
    if (defer_readback) return 0;   // (fused path: the counters travel to the host in one copy with the evaluation's results)
    SB_CUDA(ctx, cudaMemcpyAsync(C->h_counters, C->counters.p, 64 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (mode == 0 || mode == 4 || mode == 5) SB_CUDA(ctx, cudaMemcpyAsync(C->h_hash, C->hash.p, 2 * N_CONTACT_TABLES * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    return 0;
}
// after the synchronisation: true = something overflowed, capacities have been grown, the detection must run again
static bool detect_overflowed(sb_context* ctx, Contact* C)
{
    if (ctx->profile) for (int k = 0; k < 3; k++) { ctx->stage_calls[ST_TILE_PAIRS_PT + k] += C->h_counters[4 + k]; ctx->stage_calls[ST_CAND_PT + k] += C->h_counters[k]; }
    if (!C->h_counters[3]) return false;
    const int mc = std::max(C->h_counters[0], std::max(C->h_counters[1], C->h_counters[2]));
    if (mc > C->cand_cap) C->cand_cap = mc + mc / 2;
    int ml = 0, mt = 0;
    for (int l = 0; l < N_LISTS; l++) ml = std::max(ml, C->h_counters[8 + l]);
    for (int t = 0; t < N_TABLES; t++) mt = std::max(mt, C->h_counters[16 + t]);
    if (ml > C->list_cap) C->list_cap = ml + ml / 2;
    if (mt > C->table_cap) C->table_cap = mt + mt / 2;
    return true;
}
// force_swap: the new contact tables become the current ones even if they hold the same rows (something has already read them)
static int detect_publish(sb_context* ctx, Contact* C, int mode, bool force_swap)
{
    for (int l = 0; l < N_LISTS; l++) C->h_list_count[l] = C->h_counters[8 + l];
    // publish the table sizes to the potentials / friction arrays
    auto publish = [&](int t0, int t1) {
        bool any = false;   // empty before and empty now: neither the connectivity nor the pattern changed
        for (int t = t0; t < t1; t++) any = any || C->h_counters[16 + t] != 0 || C->h_table_count[t] != 0 || ctx->potentials[C->pot[t]].n_elem != 0;
        for (int t = t0; t < t1; t++) {
            const int n = C->h_counters[16 + t];
            C->h_table_count[t] = n;
            Potential& p = ctx->potentials[C->pot[t]];
            p.conn_ext = C->table[t].p;
            p.n_elem = n;
            if (t >= N_CONTACT_TABLES) {
                const int f = t - N_CONTACT_TABLES;
                ctx->arrays[C->a_fT[f]].n_rows = n; ctx->arrays[C->a_fmu[f]].n_rows = n; ctx->arrays[C->a_ffn[f]].n_rows = n;
                if (C->a_fbary[f] >= 0) ctx->arrays[C->a_fbary[f]].n_rows = n;
            }
        }
        if (any) {
            ctx->dynamic_version++;
            ctx->have_pgh = false;
        }
    };
    if (mode == 0 || mode == 4 || mode == 5) {
        static const bool no_reuse = std::getenv("SB_NO_TABLE_REUSE") != nullptr;   // diagnostic hook
        bool same = C->cur_valid && !no_reuse && !force_swap;
        for (int t = 0; t < N_CONTACT_TABLES && same; t++)
            same = C->h_counters[16 + t] == C->h_table_count[t] && C->h_hash[2 * t] == C->cur_hash[2 * t] && C->h_hash[2 * t + 1] == C->cur_hash[2 * t + 1];
        if (same) C->n_tables_same++;   // same rows as the current tables: keep them (and everything built on them)
        else {
            C->n_tables_changed++;
            for (int t = 0; t < N_CONTACT_TABLES; t++) { std::swap(C->table[t].p, C->table_back[t].p); std::swap(C->table[t].cap, C->table_back[t].cap); }
            for (int k = 0; k < 2 * N_CONTACT_TABLES; k++) C->cur_hash[k] = C->h_hash[k];
            C->cur_valid = true;
            publish(0, N_CONTACT_TABLES);
        }
    }
    if (mode == 1 || mode == 5) publish(N_CONTACT_TABLES, N_TABLES);
    return 0;
}


static int detect(sb_context* ctx, Contact* C, int mode, double enlargement)
{
    if (C->reorder_countdown-- <= 0) { int r = reorder_primitives(ctx, C); if (r) return r; }
    for (int attempt = 0; attempt < 8; attempt++) {
        int r = detect_issue(ctx, C, mode, enlargement);
        if (r) return r;
        SB_CUDA(ctx, hot_sync(ctx));
        SB_CUDA(ctx, cudaGetLastError());
        if (!detect_overflowed(ctx, C)) break;
        if (attempt == 7) return fail(ctx, SB_ERR_STATE, "contact detection: buffers keep overflowing");
    }
    return detect_publish(ctx, C, mode, false);
}

static double max_thickness(Contact* C)
{
    double m = 0.0;
    for (auto& g : C->groups) m = std::max(m, g.thickness);
    return m;
}
static int refresh_params(sb_context* ctx, Contact* C)
{
    // contact stiffness / stick-slide threshold are tiny and may change between retries: re-upload every time
    Array& ks = ctx->arrays[C->a_stiffness];
    Array& ev = ctx->arrays[C->a_epsv];
    ks.d.ensure(1); ks.n_rows = 1; ev.d.ensure(1); ev.n_rows = 1;
    if (C->uploaded_stiffness == C->stiffness && C->uploaded_epsv == C->epsv) return 0;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the pinned staging words may still be in flight
    C->uploaded_stiffness = C->stiffness; C->uploaded_epsv = C->epsv;
    ctx->h_scalars[8] = C->stiffness; ctx->h_scalars[9] = C->epsv;
    SB_CUDA(ctx, cudaMemcpyAsync(ks.d.p, ctx->h_scalars + 8, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(ev.d.p, ctx->h_scalars + 9, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int contact_update_internal(sb_context* ctx)
{
    Contact* C = ctx->contact;
    if (!C || C->groups.empty()) return 0;
    if (!C->external_vertices && !C->topology_dirty && C->contacts_state == ctx->state_version) return 0;
    StageTimer timer(ctx, ST_CONTACT_UPDATE);
    int r;
    if (C->topology_dirty && (r = upload_topology(ctx, C))) return r;
    if ((r = refresh_params(ctx, C))) return r;
    if ((r = update_vertices(ctx, C, false))) return r;
    if ((r = detect(ctx, C, 0, 2.0 * max_thickness(C)))) return r;
    C->contacts_state = ctx->state_version;
    return 0;
}
int contact_intersections_internal(sb_context* ctx, int* out_count)
{
    Contact* C = ctx->contact;
    *out_count = 0;
    if (!C || C->groups.empty()) return 0;
    if (!C->external_vertices && !C->topology_dirty && C->intersections_state == ctx->state_version) { *out_count = C->cached_intersections; return 0; }
    StageTimer timer(ctx, ST_INTERSECTIONS);
    int r;
    if (C->topology_dirty && (r = upload_topology(ctx, C))) return r;
    // A state that passes this test is evaluated next (line-search trial -> energy; accepted step -> next iteration), and
    // that evaluation starts with a contact update at the very same state: both detections share one vertex update, one
    // launch sequence and one host synchronisation.  (If the state is rejected the tables are simply rebuilt at the next one.)
    const bool with_tables = !C->external_vertices && C->contacts_state != ctx->state_version;
    if (with_tables && (r = refresh_params(ctx, C))) return r;
    if ((r = update_vertices(ctx, C, false))) return r;
    if ((r = detect(ctx, C, with_tables ? 4 : 2, with_tables ? 2.0 * max_thickness(C) : 0.0))) return r;
    *out_count = C->h_list_count[6];
    C->cached_intersections = *out_count;
    C->intersections_state = ctx->state_version;
    if (with_tables) C->contacts_state = ctx->state_version;
    return 0;
}

// ---- fused detection + evaluation (core.cu: eval_fused) ----
// Can the detection of a trial state be issued without synchronising?  (internal vertex update, clean topology)
bool contact_fusable(sb_context* ctx)
{
    Contact* C = ctx->contact;
    // (friction switched off: the host zeroes the friction potentials while the device counters keep their last values)
    return C && !C->groups.empty() && !C->external_vertices && !C->topology_dirty && C->reorder_countdown > 0 && C->enable_friction;
}
// vertex update + every launch of a mode-4 detection + the read-back of the counters, NO synchronisation
int contact_fused_issue(sb_context* ctx)
{
    Contact* C = ctx->contact;
    int r;
    if ((r = refresh_params(ctx, C))) return r;
    if ((r = update_vertices(ctx, C, false))) return r;
    C->reorder_countdown--;
    return detect_issue(ctx, C, 4, 2.0 * max_thickness(C), true);
}
// the detection's counters / table digests on the device, and their delivery to the host copies detect_* reads
void contact_readback_sources(sb_context* ctx, const int** counters, const unsigned long long** hash)
{
    *counters = ctx->contact->counters.p;
    *hash = ctx->contact->hash.p;
}
void contact_readback_deliver(sb_context* ctx, const int* counters, const unsigned long long* hash)
{
    std::memcpy(ctx->contact->h_counters, counters, 64 * sizeof(int));
    std::memcpy(ctx->contact->h_hash, hash, 2 * N_CONTACT_TABLES * sizeof(unsigned long long));
}
int contact_n_digest_words() { return 2 * N_CONTACT_TABLES; }
// where the issued detection writes the table of potential `pot` (contact tables: the back buffer), the table's counter on the
// device, the common capacity; -1 if `pot` is not a contact / friction table
int contact_issue_table(sb_context* ctx, int pot, const int32_t** conn, const int** count_dev, int* cap)
{
    Contact* C = ctx->contact;
    for (int t = 0; t < N_TABLES; t++)
        if (C->pot[t] == pot) {
            *conn = (t < N_CONTACT_TABLES) ? C->table_back[t].p : C->table[t].p;
            *count_dev = C->counters.p + 16 + t;
            *cap = C->table_cap;
            return t;
        }
    return -1;
}
// after the synchronisation: overflow -> *retry (capacities grown, nothing published); otherwise the tables are published (always
// swapped in: the evaluation queued behind the detection has read the new buffers) and the caches are set
int contact_fused_finish(sb_context* ctx, int* out_intersections, bool* retry)
{
    Contact* C = ctx->contact;
    *retry = detect_overflowed(ctx, C);
    if (*retry) return 0;
    int r = detect_publish(ctx, C, 4, true);
    if (r) return r;
    *out_intersections = C->h_list_count[6];
    C->cached_intersections = *out_intersections;
    C->intersections_state = ctx->state_version;
    C->contacts_state = ctx->state_version;
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_contact_init(sb_context* ctx, const sb_contact_bindings* b)
{
    if (!ctx || !b) return fail(ctx, SB_ERR_ARG, "sb_contact_init: bad argument");
    if (ctx->contact) return fail(ctx, SB_ERR_STATE, "sb_contact_init: already initialised");
    const int ids[8] = {b->soft_v1, b->soft_x0, b->soft_X, b->rb_v1, b->rb_w1, b->rb_t0, b->rb_q0, b->dt};
    for (int i = 0; i < 8; i++)
        if (ids[i] < 0 || ids[i] >= (int)ctx->arrays.size()) return fail(ctx, SB_ERR_ARG, "sb_contact_init: unknown array handle in bindings");
    Contact* C = new Contact();
    C->bind = *b;
    for (int t = 0; t < N_TABLES; t++) C->pot[t] = -1;
    // test hook: a small initial table capacity makes tables overflow (and grow) in the middle of a step
    if (const char* cap = std::getenv("SB_CONTACT_TABLE_CAP")) C->table_cap = std::max(1, std::atoi(cap));
    C->h_blacklist.assign(MAX_GROUPS * MAX_GROUPS, 0);
    C->h_mu.assign(MAX_GROUPS * MAX_GROUPS, 0.0);
    cudaMallocHost(&C->h_counters, 64 * sizeof(int));
    cudaMallocHost(&C->h_hash, 2 * N_CONTACT_TABLES * sizeof(unsigned long long));
    C->a_thickness = new_array(ctx, "contact_thicknesses", 1);
    C->a_stiffness = new_array(ctx, "contact_stiffness", 1);
    C->a_rb_local = new_array(ctx, "rigidbody_local_vertices", 3);
    C->a_epsv = new_array(ctx, "friction_stick_slide_threshold", 1);
    ctx->contact = C;
    // one potential per contact / friction table, bound exactly like the reference's lambdas bind their symbols
    for (int t = 0; t < N_TABLES; t++) {
        const Layout& L = LAYOUTS[t];
        const int f = t - N_CONTACT_TABLES;
        if (f >= 0) { C->a_fT[f] = C->a_fmu[f] = C->a_ffn[f] = C->a_fbary[f] = -1; }
        std::vector<sb_fetch> fetch;
        for (int k = 0; k < L.n; k++) {
            const LayoutEntry& e = L.e[k];
            int arr = -1;
            switch (e.role) {
            case R_SOFT_V1: arr = b->soft_v1; break;
            case R_SOFT_X0: arr = b->soft_x0; break;
            case R_SOFT_X: arr = b->soft_X; break;
            case R_RB_V1: arr = b->rb_v1; break;
            case R_RB_W1: arr = b->rb_w1; break;
            case R_RB_T0: arr = b->rb_t0; break;
            case R_RB_Q0: arr = b->rb_q0; break;
            case R_DT: arr = b->dt; break;
            case R_THICKNESS: arr = C->a_thickness; break;
            case R_STIFFNESS: arr = C->a_stiffness; break;
            case R_RB_LOCAL: arr = C->a_rb_local; break;
            case R_EPSV: arr = C->a_epsv; break;
            case R_F_T: arr = C->a_fT[f] = new_array(ctx, "friction_T", 6); break;
            case R_F_BARY: arr = C->a_fbary[f] = new_array(ctx, "friction_bary", e.stride); break;
            case R_F_MU: arr = C->a_fmu[f] = new_array(ctx, "friction_mu", 1); break;
            case R_F_FN: arr = C->a_ffn[f] = new_array(ctx, "friction_fn", 1); break;
            }
            fetch.push_back({arr, e.col, e.slot, e.stride});
        }
        int pot = -1;
        int r = sb_potential_create(ctx, L.name, L.conn_stride, fetch.data(), (int)fetch.size(), &pot);
        if (r) return r;
        C->pot[t] = pot;
        ctx->potentials[pot].conn_ext = nullptr;
        ctx->potentials[pot].dynamic = true;
    }
    return SB_OK;
}

int sb_contact_add_mesh(sb_context* ctx, const sb_contact_mesh* m, int* out_group)
{
    if (!ctx || !m) return fail(ctx, SB_ERR_ARG, "sb_contact_add_mesh: bad argument");
    Contact* C = ctx->contact;
    if (!C) return fail(ctx, SB_ERR_STATE, "sb_contact_add_mesh: call sb_contact_init first");
    if ((int)C->groups.size() >= MAX_GROUPS) return fail(ctx, SB_ERR_STATE, "sb_contact_add_mesh: too many contact groups");
    if (m->n_vertices <= 0 || m->contact_thickness <= 0.0) return fail(ctx, SB_ERR_ARG, "sb_contact_add_mesh: contact thickness must be positive and the mesh non-empty");
    Group g;
    g.ps = m->physical_system; g.body = m->rigid_body; g.n_v = m->n_vertices; g.n_t = m->n_triangles; g.n_e = m->n_edges;
    g.v_off = (int)C->h_v_group.size(); g.t_off = (int)C->h_t_group.size(); g.e_off = (int)C->h_e_group.size();
    g.thickness = m->contact_thickness;
    const int gi = (int)C->groups.size();
    const int rb_off = (int)C->h_rb_local.size() / 3;
    for (int i = 0; i < g.n_v; i++) {
        C->h_v_group.push_back(gi);
        if (g.ps == 0) {
            if (!m->vertex_global) return fail(ctx, SB_ERR_ARG, "sb_contact_add_mesh: deformable mesh without vertex_global");
            C->h_v_ps.push_back(m->vertex_global[i]);
        } else {
            if (!m->vertices_local) return fail(ctx, SB_ERR_ARG, "sb_contact_add_mesh: rigid mesh without vertices_local");
            C->h_v_ps.push_back(rb_off + i);
            for (int c = 0; c < 3; c++) C->h_rb_local.push_back(m->vertices_local[3 * i + c]);
        }
    }
    for (int i = 0; i < g.n_t; i++) { for (int k = 0; k < 3; k++) C->h_tri.push_back(g.v_off + m->triangles[3 * i + k]); C->h_t_group.push_back(gi); }
    for (int i = 0; i < g.n_e; i++) { for (int k = 0; k < 2; k++) C->h_edge.push_back(g.v_off + m->edges[2 * i + k]); C->h_e_group.push_back(gi); }
    C->groups.push_back(g);
    if (g.ps == 1) { C->h_blacklist[gi * MAX_GROUPS + gi] = 1; }   // rigid meshes never self-collide (EnergyFrictionalContact.cpp:209)
    C->topology_dirty = true;
    ctx->state_version++;
    if (out_group) *out_group = gi;
    return SB_OK;
}

int sb_contact_blacklist(sb_context* ctx, int a, int b)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C || a < 0 || b < 0 || a >= (int)C->groups.size() || b >= (int)C->groups.size()) return fail(ctx, SB_ERR_ARG, "sb_contact_blacklist: bad group");
    C->h_blacklist[a * MAX_GROUPS + b] = C->h_blacklist[b * MAX_GROUPS + a] = 1;
    C->topology_dirty = true;
    ctx->state_version++;
    return SB_OK;
}
int sb_contact_set_friction(sb_context* ctx, int a, int b, double mu)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C || a < 0 || b < 0 || a >= (int)C->groups.size() || b >= (int)C->groups.size()) return fail(ctx, SB_ERR_ARG, "sb_contact_set_friction: bad group");
    C->h_mu[a * MAX_GROUPS + b] = C->h_mu[b * MAX_GROUPS + a] = mu;
    C->topology_dirty = true;
    ctx->state_version++;
    return SB_OK;
}
int sb_contact_set_params(sb_context* ctx, double contact_stiffness, double epsv, int enable_pt, int enable_ee, int enable_friction)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C) return fail(ctx, SB_ERR_STATE, "sb_contact_set_params: call sb_contact_init first");
    C->stiffness = contact_stiffness; C->epsv = epsv; C->enable_pt = enable_pt; C->enable_ee = enable_ee; C->enable_friction = enable_friction;
    ctx->state_version++;
    return SB_OK;
}
int sb_contact_update(sb_context* ctx)
{
    if (!ctx) return SB_ERR_ARG;
    return contact_update_internal(ctx);
}
int sb_contact_update_friction(sb_context* ctx)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C) return fail(ctx, SB_ERR_STATE, "sb_contact_update_friction: call sb_contact_init first");
    if (C->groups.empty()) return SB_OK;
    int r;
    if (C->topology_dirty && (r = upload_topology(ctx, C))) return r;
    if ((r = refresh_params(ctx, C))) return r;
    if (!C->enable_friction) {
        for (int t = N_CONTACT_TABLES; t < N_TABLES; t++) ctx->potentials[C->pot[t]].n_elem = 0;
        return SB_OK;
    }
    if ((r = update_vertices(ctx, C, true))) return r;   // dt = 0 (EnergyFrictionalContact.cpp:543)
    return detect(ctx, C, 1, 2.0 * max_thickness(C));
}
// before_time_step of the contact model.  Friction tables are built from the positions at dt = 0
// (EnergyFrictionalContact.cpp:531-773).  When the caller guarantees that every DoF array is zero at this point -- the
// reference's own order: PointDynamics / RigidBodyDynamics zero v1 / w1 in their before_time_step callbacks, registered before
// the contact model's -- the positions at dt = 0 ARE the positions of the initial Newton state, so the same detection also
// yields the contact tables and the intersection count that sb_newton_solve asks for first (it then finds them cached).
int sb_contact_begin_time_step(sb_context* ctx, int dofs_are_zero)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C) return fail(ctx, SB_ERR_STATE, "sb_contact_begin_time_step: call sb_contact_init first");
    if (C->groups.empty()) return SB_OK;
    if (!dofs_are_zero || C->external_vertices) return sb_contact_update_friction(ctx);
    int r;
    if (C->topology_dirty && (r = upload_topology(ctx, C))) return r;
    if ((r = refresh_params(ctx, C))) return r;
    if (!C->enable_friction)
        for (int t = N_CONTACT_TABLES; t < N_TABLES; t++) ctx->potentials[C->pot[t]].n_elem = 0;
    StageTimer timer(ctx, ST_INTERSECTIONS);
    if ((r = update_vertices(ctx, C, true))) return r;
    if ((r = detect(ctx, C, C->enable_friction ? 5 : 4, 2.0 * max_thickness(C)))) return r;
    C->cached_intersections = C->h_list_count[6];
    C->intersections_state = ctx->state_version;
    C->contacts_state = ctx->state_version;
    return SB_OK;
}
int sb_contact_count_intersections(sb_context* ctx, int* out_count)
{
    if (!ctx || !out_count) return SB_ERR_ARG;
    return contact_intersections_internal(ctx, out_count);
}
int sb_contact_detect(sb_context* ctx, double enlargement, int with_intersections)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C) return fail(ctx, SB_ERR_STATE, "sb_contact_detect: call sb_contact_init first");
    int r;
    if (C->topology_dirty && (r = upload_topology(ctx, C))) return r;
    if ((r = refresh_params(ctx, C))) return r;
    if ((r = update_vertices(ctx, C, false))) return r;
    C->contacts_state = 0; C->intersections_state = 0;   // tables / lists are rewritten with a caller-chosen enlargement
    return detect(ctx, C, with_intersections ? 3 : 0, enlargement);
}
int sb_contact_get_proximity(sb_context* ctx, int kind, int32_t* host_ids, double* host_dist, int capacity, int* out_count, int* out_width)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C || kind < 0 || kind >= N_LISTS) return fail(ctx, SB_ERR_ARG, "sb_contact_get_proximity: bad argument");
    const int n = C->h_list_count[kind];
    if (out_count) *out_count = n;
    if (out_width) *out_width = LIST_WIDTH[kind];
    if (host_ids && n > 0) {
        if (capacity < n) return fail(ctx, SB_ERR_ARG, "sb_contact_get_proximity: buffer too small");
        SB_CUDA(ctx, cudaMemcpyAsync(host_ids, C->list_ids[kind].p, sizeof(int32_t) * (size_t)n * LIST_WIDTH[kind], cudaMemcpyDeviceToHost, ctx->stream));
        if (host_dist) SB_CUDA(ctx, cudaMemcpyAsync(host_dist, C->list_dist[kind].p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SB_OK;
}
int sb_contact_get_vertices(sb_context* ctx, int group, double* host_xyz)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C || group < 0 || group >= (int)C->groups.size() || !host_xyz) return fail(ctx, SB_ERR_ARG, "sb_contact_get_vertices: bad argument");
    const Group& g = C->groups[group];
    SB_CUDA(ctx, cudaMemcpyAsync(host_xyz, C->x.p + 3 * (size_t)g.v_off, sizeof(double) * 3 * g.n_v, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int sb_contact_set_vertices(sb_context* ctx, int group, const double* host_xyz)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C || group < 0 || group >= (int)C->groups.size() || !host_xyz) return fail(ctx, SB_ERR_ARG, "sb_contact_set_vertices: bad argument");
    int r;
    if (C->topology_dirty && (r = upload_topology(ctx, C))) return r;
    const Group& g = C->groups[group];
    SB_CUDA(ctx, cudaMemcpyAsync(C->x.p + 3 * (size_t)g.v_off, host_xyz, sizeof(double) * 3 * g.n_v, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->state_version++;
    C->external_vertices = true;
    return SB_OK;
}
int sb_contact_potential(sb_context* ctx, const char* name, int* out_potential)
{
    Contact* C = ctx ? ctx->contact : nullptr;
    if (!C || !name || !out_potential) return fail(ctx, SB_ERR_ARG, "sb_contact_potential: bad argument");
    *out_potential = -1;
    for (int t = 0; t < N_TABLES; t++)
        if (std::strcmp(LAYOUTS[t].name, name) == 0) *out_potential = C->pot[t];
    return SB_OK;
}

}  // extern "C"
