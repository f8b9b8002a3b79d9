// Potentials without a hand-written kernel: symx operation sequence -> CUDA source -> NVRTC -> cubin -> element kernel.
//
// Replaces, for user-defined potentials (GlobalPotential::add_potential with an arbitrary energy lambda, e.g. the magnetic
// attraction of examples/main.cpp:666-690), the reference's code generator + host-compiler JIT:
// symx/src/compile/Compilation.cpp:381-469 (`_add_instructions_scalar`: one C statement per symx::core::Op,
// compile/FixedBranchSequence.h:44-83) and Compilation.cpp's g++ / dlopen round trip.  The caller (the reference-side shim)
// differentiates the energy with symx's own symbolic engine exactly as SecondOrderCompiledPotential.cpp:62-80 does and hands over
// the two operation sequences [E] and [E | grad | hess] as plain arrays of sb_op; nothing symbolic lives here.
//
// Generated kernel: one thread per element -- gather of the in[] values through the potential's fetch table, the straight-line
// code of the sequence (in[] / out[] indices are literals, so both arrays are registers or, for big elements, local memory), then
// the element-output contract of every other kernel (E per element, FP64 atomic gradient scatter, dense row-major n x n Hessian,
// block rows).  Compiled once per (source, architecture): cubins are cached on disk under $SB_CACHE_DIR (default
// ~/.cache/stark_b200), keyed by a 128-bit hash of the generated source.  NVRTC is loaded with dlopen at the first use, so the
// library itself does not depend on it.
#include "internal.h"
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace sb {

// symx::ExprType (symx/src/symbol/Expr.h:12-44)
enum : int { OP_BRANCH = 2, OP_CONST = 4, OP_OUT = 5, OP_ADD = 6, OP_SUB = 7, OP_MUL = 8, OP_RECIP = 9, OP_POWN = 10, OP_POWF = 11, OP_SQRT = 12, OP_LN = 13,
             OP_LOG10 = 14, OP_EXP = 15, OP_SIN = 16, OP_COS = 17, OP_TAN = 18, OP_ASIN = 19, OP_ACOS = 20, OP_ATAN = 21, OP_PRINT = 22 };

struct UserKernel {
    std::string name;
    std::vector<int> dof_slot;
    KernelInfo info;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_pgh = nullptr, k_p = nullptr;
};

// argument block of the generated kernels (restated verbatim in the generated source)
struct UArgs {
    const FetchSlot* slots;
    const int32_t* conn;
    int32_t conn_stride, n_elem;
    int32_t dof_offset[MAX_BLOCKS], conn_col[MAX_BLOCKS];
    double* grad;
    double* H;
    int32_t* rows;
    double* E_elem;
    double* g_elem;   // optional per-element gradient (parity dumps), may be null
};

static std::string fmt_double(double v)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEni") == std::string::npos) s += ".0";   // (inf / nan never come out of symx constants)
    return s;
}

// one C statement per operation: the scalar back-end of the reference's generator, for CUDA
static int emit_body(const sb_op* ops, int n_ops, int n_in, int n_out, std::string& code, std::string& err)
{
    auto idx = [&](int i) { return i < n_in ? "in[" + std::to_string(i) + "]" : "v" + std::to_string(i); };
    int depth = 1;
    auto tab = [&]() { return std::string((size_t)depth, '\t'); };
    for (int k = 0; k < n_ops; k++) {
        const sb_op& op = ops[k];
        const std::string d = "const double " + idx(op.dst) + " = ";
        switch (op.type) {
        case OP_OUT:
            if (op.dst < 0 || op.dst >= n_out) { err = "output index out of range"; return SB_ERR_ARG; }
            code += tab() + "out[" + std::to_string(op.dst) + "] = " + idx(op.a) + ";\n"; break;
        case OP_CONST: code += tab() + d + fmt_double(op.constant) + ";\n"; break;
        case OP_ADD: code += tab() + d + idx(op.a) + " + " + idx(op.b) + ";\n"; break;
        case OP_SUB: code += tab() + d + idx(op.a) + " - " + idx(op.b) + ";\n"; break;
        case OP_MUL: code += tab() + d + idx(op.a) + " * " + idx(op.b) + ";\n"; break;
        case OP_RECIP: code += tab() + d + "1.0 / " + idx(op.a) + ";\n"; break;
        case OP_POWN: code += tab() + d + "pow(" + idx(op.a) + ", (double)" + std::to_string(op.b) + ");\n"; break;   // (std::pow(double, int) promotes the exponent)
        case OP_POWF: code += tab() + d + "pow(" + idx(op.a) + ", " + idx(op.b) + ");\n"; break;
        case OP_SQRT: code += tab() + d + "sqrt(" + idx(op.a) + ");\n"; break;
        case OP_LN: code += tab() + d + "(" + idx(op.a) + " <= 0.0) ? -SB_INF : log(" + idx(op.a) + ");\n"; break;
        case OP_LOG10: code += tab() + d + "(" + idx(op.a) + " <= 0.0) ? -SB_INF : log10(" + idx(op.a) + ");\n"; break;
        case OP_EXP: code += tab() + d + "exp(" + idx(op.a) + ");\n"; break;
        case OP_SIN: code += tab() + d + "sin(" + idx(op.a) + ");\n"; break;
        case OP_COS: code += tab() + d + "cos(" + idx(op.a) + ");\n"; break;
        case OP_TAN: code += tab() + d + "tan(" + idx(op.a) + ");\n"; break;
        case OP_ASIN: code += tab() + d + "asin(" + idx(op.a) + ");\n"; break;
        case OP_ACOS: code += tab() + d + "acos(" + idx(op.a) + ");\n"; break;
        case OP_ATAN: code += tab() + d + "atan(" + idx(op.a) + ");\n"; break;
        case OP_PRINT: code += tab() + d + "0.0;\n"; break;   // (the reference prints the operand from every thread; dropped here)
        case OP_BRANCH:
            // core::Op encoding (FixedBranchSequence.h:58-82): cond == -2 end-if; a == 0 positive branch; a == 1 negative branch
            if (op.cond == -2) {
                if (--depth < 1) { err = "unbalanced branch"; return SB_ERR_ARG; }
                code += tab() + "}\n";
            }
            if (op.a == 0) { code += tab() + "if (" + idx(op.cond) + " > 0.0)\n" + tab() + "{\n"; depth++; }
            else if (op.a == 1) {
                if (--depth < 1) { err = "unbalanced branch"; return SB_ERR_ARG; }
                code += tab() + "}\n" + tab() + "else\n" + tab() + "{\n";
                depth++;
            }
            break;
        default:
            err = "operation type " + std::to_string(op.type) + " has no CUDA translation";
            return SB_ERR_ARG;
        }
    }
    if (depth != 1) { err = "unbalanced branch"; return SB_ERR_ARG; }
    return 0;
}

static int generate_source(const char* name, int n_in, int nb, const sb_op* ops_p, int n_ops_p, const sb_op* ops_pgh, int n_ops_pgh, std::string& src, std::string& err)
{
    const int n = 3 * nb, n_out = 1 + n + n * n;
    std::string body_p, body_pgh;
    int r = emit_body(ops_p, n_ops_p, n_in, 1, body_p, err);
    if (r) return r;
    r = emit_body(ops_pgh, n_ops_pgh, n_in, n_out, body_pgh, err);
    if (r) return r;
    std::ostringstream o;
    o << "// stark_b200 generated element kernel of potential '" << name << "'\n"
      << "#define N_IN " << n_in << "\n#define NB " << nb << "\n#define N_DOF " << n << "\n#define N_OUT " << n_out << "\n"
      << "#define SB_INF __longlong_as_double(0x7ff0000000000000LL)\n"
      << "struct FetchSlot { const double* base; int conn_col; int stride; int off; int pad; };\n"
      << "struct UArgs { const FetchSlot* slots; const int* conn; int conn_stride, n_elem; int dof_offset[" << MAX_BLOCKS << "], conn_col[" << MAX_BLOCKS << "];\n"
      << "               double* grad; double* H; int* rows; double* E_elem; double* g_elem; };\n"
      << "__device__ __forceinline__ void f_p(const double* __restrict__ in, double* __restrict__ out)\n{\n" << body_p << "}\n"
      << "__device__ __forceinline__ void f_pgh(const double* __restrict__ in, double* __restrict__ out)\n{\n" << body_pgh << "}\n"
      << "__device__ __forceinline__ void gather(const UArgs& a, int e, double* in)\n{\n"
      << "#pragma unroll\n\tfor (int s = 0; s < N_IN; s++) {\n\t\tconst FetchSlot fs = a.slots[s];\n"
      << "\t\tconst int row = (fs.conn_col >= 0) ? a.conn[(size_t)e * a.conn_stride + fs.conn_col] : 0;\n"
      << "\t\tin[s] = fs.base[(size_t)row * fs.stride + fs.off];\n\t}\n}\n"
      << "extern \"C\" __global__ void __launch_bounds__(128) user_p(const UArgs a)\n{\n"
      << "\tconst int e = blockIdx.x * 128 + threadIdx.x;\n\tif (e >= a.n_elem) return;\n"
      << "\tdouble in[N_IN], out[1];\n\tgather(a, e, in);\n\tf_p(in, out);\n\ta.E_elem[e] = out[0];\n}\n"
      << "extern \"C\" __global__ void __launch_bounds__(128) user_pgh(const UArgs a)\n{\n"
      << "\tconst int e = blockIdx.x * 128 + threadIdx.x;\n\tif (e >= a.n_elem) return;\n"
      << "\tdouble in[N_IN], out[N_OUT];\n\tgather(a, e, in);\n\tf_pgh(in, out);\n\ta.E_elem[e] = out[0];\n"
      << "\tconst int* ce = a.conn + (size_t)e * a.conn_stride;\n"
      << "#pragma unroll\n\tfor (int b = 0; b < NB; b++) {\n\t\tconst int base = a.dof_offset[b] + 3 * ce[a.conn_col[b]];\n"
      << "\t\ta.rows[(size_t)e * NB + b] = base / 3;\n"
      << "#pragma unroll\n\t\tfor (int k = 0; k < 3; k++) atomicAdd(a.grad + base + k, out[1 + 3 * b + k]);\n\t}\n"
      << "\tif (a.g_elem) {\n#pragma unroll\n\t\tfor (int i = 0; i < N_DOF; i++) a.g_elem[(size_t)e * N_DOF + i] = out[1 + i];\n\t}\n"
      << "\tdouble* He = a.H + (size_t)e * N_DOF * N_DOF;\n"
      << "#pragma unroll\n\tfor (int i = 0; i < N_DOF * N_DOF; i++) He[i] = out[1 + N_DOF + i];\n}\n";
    src = o.str();
    return 0;
}

// ---- NVRTC through dlopen ----
struct Nvrtc {
    void* h = nullptr;
    int (*createProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*compileProgram)(void*, int, const char* const*) = nullptr;
    int (*getCUBINSize)(void*, size_t*) = nullptr;
    int (*getCUBIN)(void*, char*) = nullptr;
    int (*getProgramLogSize)(void*, size_t*) = nullptr;
    int (*getProgramLog)(void*, char*) = nullptr;
    int (*destroyProgram)(void**) = nullptr;
    int (*version)(int*, int*) = nullptr;
    bool ok = false;
};
static Nvrtc& nvrtc()
{
    static Nvrtc N;
    static bool tried = false;
    if (tried) return N;
    tried = true;
    for (const char* p : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) {
        N.h = dlopen(p, RTLD_NOW | RTLD_LOCAL);
        if (N.h) break;
    }
    if (!N.h) return N;
    auto sym = [&](const char* s) { return dlsym(N.h, s); };
    N.createProgram = (decltype(N.createProgram))sym("nvrtcCreateProgram");
    N.compileProgram = (decltype(N.compileProgram))sym("nvrtcCompileProgram");
    N.getCUBINSize = (decltype(N.getCUBINSize))sym("nvrtcGetCUBINSize");
    N.getCUBIN = (decltype(N.getCUBIN))sym("nvrtcGetCUBIN");
    N.getProgramLogSize = (decltype(N.getProgramLogSize))sym("nvrtcGetProgramLogSize");
    N.getProgramLog = (decltype(N.getProgramLog))sym("nvrtcGetProgramLog");
    N.destroyProgram = (decltype(N.destroyProgram))sym("nvrtcDestroyProgram");
    N.version = (decltype(N.version))sym("nvrtcVersion");
    N.ok = N.createProgram && N.compileProgram && N.getCUBINSize && N.getCUBIN && N.getProgramLogSize && N.getProgramLog && N.destroyProgram;
    return N;
}

static const char* ARCH_OPT = "--gpu-architecture=sm_100a";

// 2 x 64-bit FNV-1a over the source and the compile options: the cache key
static std::string source_key(const std::string& src)
{
    unsigned long long h1 = 1469598103934665603ull, h2 = 0x9ae16a3b2f90404full;
    auto mix = [&](unsigned char c) { h1 = (h1 ^ c) * 1099511628211ull; h2 = (h2 ^ (c + 0x9eu)) * 0x100000001b3ull; h2 ^= h2 >> 29; };
    for (unsigned char c : src) mix(c);
    for (const char* p = ARCH_OPT; *p; p++) mix((unsigned char)*p);
    char buf[40];
    snprintf(buf, sizeof(buf), "%016llx%016llx", h1, h2);
    return buf;
}
static std::string cache_dir()
{
    if (const char* d = getenv("SB_CACHE_DIR")) return d;
    const char* home = getenv("HOME");
    return std::string(home && *home ? home : "/tmp") + "/.cache/stark_b200";
}

static int compile_to_cubin(const std::string& src, const char* name, std::vector<char>& cubin, bool& cached, std::string& err)
{
    const std::string dir = cache_dir(), path = dir + "/" + source_key(src) + ".cubin";
    cached = false;
    {
        std::ifstream f(path, std::ios::binary);
        if (f) {
            cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            if (!cubin.empty()) { cached = true; return 0; }
        }
    }
    Nvrtc& N = nvrtc();
    if (!N.ok) { err = "NVRTC (libnvrtc.so.12) could not be loaded: user potentials need it"; return SB_ERR_NO_KERNEL; }
    void* prog = nullptr;
    if (N.createProgram(&prog, src.c_str(), (std::string(name) + ".cu").c_str(), 0, nullptr, nullptr) != 0) { err = "nvrtcCreateProgram failed"; return SB_ERR_CUDA; }
    const char* opts[] = {ARCH_OPT, "-std=c++17", "-lineinfo"};
    const int rc = N.compileProgram(prog, 3, opts);
    if (rc != 0) {
        size_t n = 0;
        N.getProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) N.getProgramLog(prog, &log[0]);
        N.destroyProgram(&prog);
        err = "NVRTC compilation of the generated kernel failed:\n" + log.substr(0, 2000);
        return SB_ERR_CUDA;
    }
    size_t n = 0;
    N.getCUBINSize(prog, &n);
    cubin.resize(n);
    N.getCUBIN(prog, cubin.data());
    N.destroyProgram(&prog);
    // best effort: a cache that cannot be written only costs the next start-up a compilation
    mkdir((dir.substr(0, dir.find_last_of('/'))).c_str(), 0755);
    mkdir(dir.c_str(), 0755);
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    {
        std::ofstream f(tmp, std::ios::binary);
        if (f) { f.write(cubin.data(), (std::streamsize)cubin.size()); f.close(); rename(tmp.c_str(), path.c_str()); }
    }
    return 0;
}

static void user_launch(const EvalArgs& a, cudaStream_t st, bool pgh)
{
    const UserKernel* U = static_cast<const UserKernel*>(a.user);
    if (!U || a.n_elem <= 0) return;
    UArgs ua;
    ua.slots = a.slots; ua.conn = a.conn; ua.conn_stride = a.conn_stride; ua.n_elem = a.n_elem;
    for (int b = 0; b < MAX_BLOCKS; b++) { ua.dof_offset[b] = a.blocks[b].dof_offset; ua.conn_col[b] = a.blocks[b].conn_col; }
    ua.grad = a.grad; ua.H = a.H; ua.rows = a.rows; ua.E_elem = a.E_elem; ua.g_elem = a.g_elem;
    void* args[] = {&ua};
    cudaLaunchKernel((const void*)(pgh ? U->k_pgh : U->k_p), dim3((a.n_elem + 127) / 128), dim3(128), args, 0, st);
}
static void user_launch_pgh(const EvalArgs& a, cudaStream_t st) { user_launch(a, st, true); }
static void user_launch_p(const EvalArgs& a, cudaStream_t st) { user_launch(a, st, false); }

int potential_create_with_kernel(sb_context* ctx, const KernelInfo* k, const char* kernel_name, int conn_stride, const sb_fetch* fetch, int n_fetch, int* out_potential);

void user_kernels_destroy(sb_context* ctx)
{
    for (void* p : ctx->user_kernels) {
        UserKernel* U = static_cast<UserKernel*>(p);
        if (U->lib) cudaLibraryUnload(U->lib);
        delete U;
    }
    ctx->user_kernels.clear();
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_user_codegen(const char* name, int n_in, int n_blocks, const sb_op* ops_p, int n_ops_p, const sb_op* ops_pgh, int n_ops_pgh, char* out_source, long long capacity, long long* out_length)
{
    if (!name || n_in <= 0 || n_blocks <= 0 || n_blocks > MAX_BLOCKS || !ops_p || !ops_pgh || n_ops_p <= 0 || n_ops_pgh <= 0) return SB_ERR_ARG;
    std::string src, err;
    const int r = generate_source(name, n_in, n_blocks, ops_p, n_ops_p, ops_pgh, n_ops_pgh, src, err);
    if (r) return r;
    if (out_length) *out_length = (long long)src.size();
    if (out_source && capacity > 0) {
        const size_t n = std::min((size_t)capacity - 1, src.size());
        memcpy(out_source, src.data(), n);
        out_source[n] = '\0';
    }
    return 0;
}

int sb_user_compile(const char* source, long long* out_cubin_bytes, int* out_was_cached, char* out_log, int log_capacity)
{
    if (!source) return SB_ERR_ARG;
    std::vector<char> cubin;
    bool cached = false;
    std::string err;
    const int r = compile_to_cubin(source, "sb_user", cubin, cached, err);
    if (out_log && log_capacity > 0) { strncpy(out_log, err.c_str(), (size_t)log_capacity - 1); out_log[log_capacity - 1] = '\0'; }
    if (out_cubin_bytes) *out_cubin_bytes = (long long)cubin.size();
    if (out_was_cached) *out_was_cached = cached ? 1 : 0;
    return r;
}

int sb_potential_create_user(sb_context* ctx, const char* name, int conn_stride, const sb_fetch* fetch, int n_fetch, int n_in, int n_blocks,
                             const int32_t* dof_block_slots, const sb_op* ops_p, int n_ops_p, const sb_op* ops_pgh, int n_ops_pgh, int* out_potential)
{
    if (!ctx) return SB_ERR_ARG;
    if (!name || !fetch || n_fetch <= 0 || n_in <= 0 || n_blocks <= 0 || n_blocks > MAX_BLOCKS || !dof_block_slots || !ops_p || !ops_pgh || n_ops_p <= 0 || n_ops_pgh <= 0)
        return fail(ctx, SB_ERR_ARG, "sb_potential_create_user: bad argument");
    std::string src, err;
    int r = generate_source(name, n_in, n_blocks, ops_p, n_ops_p, ops_pgh, n_ops_pgh, src, err);
    if (r) return fail(ctx, r, std::string("sb_potential_create_user(") + name + "): " + err);
    std::vector<char> cubin;
    bool cached = false;
    r = compile_to_cubin(src, name, cubin, cached, err);
    if (r) return fail(ctx, r, std::string("sb_potential_create_user(") + name + "): " + err);
    std::unique_ptr<UserKernel> U(new UserKernel());
    U->name = name;
    U->dof_slot.assign(dof_block_slots, dof_block_slots + n_blocks);
    if (cudaLibraryLoadData(&U->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&U->k_pgh, U->lib, "user_pgh") != cudaSuccess || cudaLibraryGetKernel(&U->k_p, U->lib, "user_p") != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        if (U->lib) cudaLibraryUnload(U->lib);
        return fail(ctx, SB_ERR_CUDA, std::string("sb_potential_create_user(") + name + "): loading the compiled kernel failed: " + cudaGetErrorString(e));
    }
    U->info.name = U->name.c_str();
    U->info.n_in = n_in; U->info.n_dof = 3 * n_blocks; U->info.nb = n_blocks;
    U->info.dof_slot = U->dof_slot.data();
    U->info.launch_pgh = user_launch_pgh; U->info.launch_p = user_launch_p;
    U->info.p_kind = -1;
    U->info.user = U.get();
    r = potential_create_with_kernel(ctx, &U->info, name, conn_stride, fetch, n_fetch, out_potential);
    if (r) { cudaLibraryUnload(U->lib); return r; }
    ctx->user_kernels.push_back(U.release());
    return SB_OK;
}

}  // extern "C"
