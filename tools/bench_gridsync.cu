// Micro-benchmark: cost of one "grid-wide all-reduce of two doubles that is also a barrier" on a persistent cooperative grid
// (one 1024-thread CTA per SM), the synchronisation step of the PCG kernel.  Variants are timed in isolation, with and
// without global stores in flight before the synchronisation (the vector phase of PCG writes x and u just before it).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/bench_gridsync tools/bench_gridsync.cu && gpurun_out/bench_gridsync
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int T = 1024;

__device__ __forceinline__ double block_sum(double v, double* s)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < T / 32) t = s[threadIdx.x];
    if (w == 0)
        for (int o = T / 64; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    return t;
}

__device__ __forceinline__ void ll_store(uint4* slot, double v, unsigned epoch)
{
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(slot), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch) : "memory");
}
template<int ACQ> __device__ __forceinline__ bool ll_try(const uint4* slot, unsigned epoch, double& v)
{
    unsigned lo, f0, hi, f1;
    if (ACQ == 1) asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot) : "memory");
    v = __hiloint2double((int)hi, (int)lo);
    return f0 == epoch && f1 == epoch;
}

struct Args {
    unsigned* counter; double* part; uint4* inbox; double* scratch; double* out; long long* cycles;
    int iters, variant, stores;
};

// V0: atomic counter barrier + re-read of partials
__device__ __forceinline__ double v0(const Args& a, double v, unsigned& epoch, double* s, double* bc)
{
    double t = block_sum(v, s);
    if (threadIdx.x == 0) __stcg(a.part + blockIdx.x, t);
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.counter, 1u);
        const unsigned target = epoch * gridDim.x;
        unsigned c;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(a.counter) : "memory"); } while (c < target);
    }
    __syncthreads();
    double acc = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += T) acc += __ldcg(a.part + i);
    t = block_sum(acc, s);
    if (threadIdx.x == 0) *bc = t;
    __syncthreads();
    return *bc;
}

// push-model flagged slots.  FENCE: 0 none, 1 warp-0 fence.acq_rel, 2 thread-0 fence + barrier then 148 threads store
// ACQ: 1 ld.acquire polls, 0 relaxed polls + one fence by pollers, 2 relaxed polls, no fence
template<int FENCE, int ACQ, int PUSHERS>
__device__ __forceinline__ double vll(const Args& a, double v, unsigned epoch, double* s, double* bc)
{
    const int G = gridDim.x, me = blockIdx.x;
    double t = block_sum(v, s);
    if (threadIdx.x == 0) {
        if (FENCE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        *bc = t;
    }
    __syncthreads();
    t = *bc;
    if (PUSHERS == 32) {
        if (threadIdx.x < 32) {
            if (FENCE == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (int i = threadIdx.x; i < G; i += 32) ll_store(a.inbox + (size_t)i * G + me, t, epoch);
        }
    } else {
        for (int i = threadIdx.x; i < G; i += T) {
            if (FENCE == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            ll_store(a.inbox + (size_t)i * G + me, t, epoch);
        }
    }
    double acc = 0.0;
    for (int i = threadIdx.x; i < G; i += T) {
        double vi = 0.0;
        while (!ll_try<ACQ>(a.inbox + (size_t)me * G + i, epoch, vi)) { }
        acc += vi;
    }
    if (ACQ == 0 && threadIdx.x < G) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    t = block_sum(acc, s);
    if (threadIdx.x == 0) *bc = t;
    __syncthreads();
    return *bc;
}

// pull model: one slot per CTA, everybody polls everybody's slot
template<int FENCE>
__device__ __forceinline__ double vpull(const Args& a, double v, unsigned epoch, double* s, double* bc, int stride)
{
    const int G = gridDim.x, me = blockIdx.x;
    double t = block_sum(v, s);
    if (threadIdx.x == 0) {
        if (FENCE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        ll_store(a.inbox + (size_t)me * stride, t, epoch);
    }
    double acc = 0.0;
    for (int i = threadIdx.x; i < G; i += T) {
        double vi = 0.0;
        while (!ll_try<1>(a.inbox + (size_t)i * stride, epoch, vi)) { }
        acc += vi;
    }
    t = block_sum(acc, s);
    if (threadIdx.x == 0) *bc = t;
    __syncthreads();
    return *bc;
}

__global__ void __launch_bounds__(T) k_bench(const Args a)
{
    __shared__ double s[64];
    __shared__ double bc;
    unsigned epoch = 0;
    double v = 1.0 + threadIdx.x * 1e-6;
    double sum = 0.0;
    const long long t0 = clock64();
    for (int it = 1; it <= a.iters; it++) {
        if (a.stores) {
            double* q = a.scratch + ((size_t)blockIdx.x * T + threadIdx.x) * 6;
            for (int k = 0; k < a.stores; k++) __stcg(q + k, v + it);
        }
        double r;
        Args b = a;
        b.inbox = a.inbox + (size_t)(it & 1) * gridDim.x * gridDim.x * 8;   // two inboxes in turn: a slot is rewritten only after every CTA has read it
        {
            const Args& a = b;
        switch (a.variant) {
        case 0: r = v0(a, v, epoch, s, &bc); break;
        case 1: r = vll<1, 1, 32>(a, v, (unsigned)it, s, &bc); break;     // warp-0 fence, acquire polls
        case 2: r = vll<0, 1, 148>(a, v, (unsigned)it, s, &bc); break;    // no fence, acquire polls
        case 3: r = vll<0, 2, 148>(a, v, (unsigned)it, s, &bc); break;    // no fence, relaxed polls
        case 4: r = vll<1, 0, 32>(a, v, (unsigned)it, s, &bc); break;     // warp-0 fence, relaxed polls + fence
        case 5: r = vll<2, 1, 148>(a, v, (unsigned)it, s, &bc); break;    // thread-0 fence, 148 pushers, acquire polls
        case 6: r = vll<1, 1, 148>(a, v, (unsigned)it, s, &bc); break;    // every pusher fences
        case 7: r = vpull<1>(a, v, (unsigned)it, s, &bc, 1); break;       // pull, packed slots
        case 8: r = vpull<1>(a, v, (unsigned)it, s, &bc, 8); break;       // pull, one slot per 128 B line
        case 9: r = vll<2, 2, 148>(a, v, (unsigned)it, s, &bc); break;    // thread-0 fence, relaxed polls, no acquire
        default: r = 0.0;
        }
        }
        sum += r;
        v = v * 1.0000001;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { a.out[blockIdx.x] = sum; a.cycles[blockIdx.x] = t1 - t0; }
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int G = sms;
    Args a;
    cudaMalloc(&a.counter, 4); cudaMalloc(&a.part, 8 * G); cudaMalloc(&a.inbox, sizeof(uint4) * (size_t)G * G * 16);
    cudaMalloc(&a.scratch, 8 * (size_t)G * T * 6); cudaMalloc(&a.out, 8 * G); cudaMalloc(&a.cycles, 8 * G);
    a.iters = 2000;
    const char* names[] = {"atomic barrier + re-read", "push: warp0 fence, acq polls", "push: no fence, acq polls", "push: no fence, relaxed polls",
                           "push: warp0 fence, relaxed polls + fence", "push: thread0 fence, 148 pushers, acq polls", "push: every pusher fences",
                           "pull packed", "pull 128B slots", "push: thread0 fence, relaxed polls"};
    for (int stores = 0; stores <= 6; stores += 6)
        for (int variant = 0; variant < 10; variant++) {
            a.variant = variant; a.stores = stores;
            cudaMemset(a.counter, 0, 4); cudaMemset(a.inbox, 0, sizeof(uint4) * (size_t)G * G * 16);
            void* args[] = {&a};
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cudaError_t err = cudaLaunchCooperativeKernel((const void*)k_bench, dim3(G), dim3(T), args, 0, 0);
            cudaEventRecord(e1);
            cudaError_t err2 = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double out = 0;
            cudaMemcpy(&out, a.out, 8, cudaMemcpyDeviceToHost);
            printf("stores=%d  V%d %-46s %7.3f us/sync   (%s %s) check=%.6g\n", stores, variant, names[variant], 1e3 * ms / a.iters,
                   cudaGetErrorString(err), cudaGetErrorString(err2), out);
        }
    return 0;
}
