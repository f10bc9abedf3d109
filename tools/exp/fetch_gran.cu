// Experiment (not product): how many DRAM bytes does a sparse pick of 2 floats per lattice cell cost on B200, as a
// function of the L2 fetch granularity (cudaLimitMaxL2FetchGranularity) and the load flavour?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fetch_gran tools/exp/fetch_gran.cu
//   gpurun_out/fetch_gran            (prints time per variant; run under ncu --metrics dram__bytes_read.sum for bytes)
//
// Pattern = k_gather's: cell c (V floats) -> (lp[c*V + blank], lp[c*V + label[c % U]]), pairs written as float2.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int FLAVOUR>
__device__ __forceinline__ float ld(const float *p) {
    float v;
    if (FLAVOUR == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (FLAVOUR == 1) asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (FLAVOUR == 2) asm volatile("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (FLAVOUR == 3) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if (FLAVOUR == 4) asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    }
    return v;
}

template <int FLAVOUR>
__global__ void __launch_bounds__(256) k_pick(const float *__restrict__ lp, const int *__restrict__ lab, int U, int V,
                                               float2 *__restrict__ out, int64_t cells) {
    constexpr int ILP = 4;
    for (int64_t c0 = ((int64_t)blockIdx.x * 256) * ILP; c0 < cells; c0 += (int64_t)gridDim.x * 256 * ILP) {
        float vb[ILP], vl[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const int64_t c = c0 + k * 256 + threadIdx.x;
            if (c < cells) {
                const float *row = lp + c * V;
                vb[k] = ld<FLAVOUR>(row);
                vl[k] = ld<FLAVOUR>(row + lab[c % U]);
            }
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const int64_t c = c0 + k * 256 + threadIdx.x;
            if (c < cells) out[c] = make_float2(vb[k], vl[k]);
        }
    }
}

__global__ void __launch_bounds__(256) k_stream(const float4 *__restrict__ lp, int64_t n4, float *__restrict__ out) {
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 v;
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(lp + i));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void k_flush(float4 *p, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) p[i] = make_float4(1, 2, 3, 4);
}

template <int FLAVOUR>
static float run_pick(const float *lp, const int *lab, int U, int V, float2 *out, int64_t cells, float4 *flush, int64_t fl4) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        k_flush<<<1184, 256>>>(flush, fl4);
        cudaEventRecord(e0);
        k_pick<FLAVOUR><<<148 * 16, 256>>>(lp, lab, U, V, out, cells);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    const int grans[3] = {32, 64, 128};
    struct Cfg { const char *name; int U, V; int64_t cells; } cfgs[2] = {{"c2-like V=28", 40, 28, (int64_t)128 * 150 * 40 * 4},
                                                                         {"c4-like V=50", 300, 50, (int64_t)16 * 1500 * 300}};
    size_t lim = 0;
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("default cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    const int64_t fl4 = (int64_t)256 << 20 >> 4;       // 256 MB flush buffer
    float4 *flush; CK(cudaMalloc(&flush, fl4 * 16));
    for (auto &c : cfgs) {
        float *lp; int *lab; float2 *out; float *dummy;
        const int64_t nfl = c.cells * c.V;
        CK(cudaMalloc(&lp, nfl * 4)); CK(cudaMalloc(&lab, c.U * 4)); CK(cudaMalloc(&out, c.cells * 8)); CK(cudaMalloc(&dummy, 4));
        CK(cudaMemset(lp, 0, nfl * 4));
        int *h = (int *)malloc(c.U * 4);
        srand(1);
        for (int u = 0; u < c.U; ++u) h[u] = 1 + rand() % (c.V - 1);
        CK(cudaMemcpy(lab, h, c.U * 4, cudaMemcpyHostToDevice));
        printf("== %s: %lld cells, tensor %.1f MB, useful %.1f MB, pairs out %.1f MB\n", c.name, (long long)c.cells, nfl * 4 / 1e6,
               c.cells * 8 / 1e6, c.cells * 8 / 1e6);
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_flush<<<1184, 256>>>(flush, fl4);
            cudaEventRecord(e0);
            k_stream<<<148 * 16, 256>>>((const float4 *)lp, nfl / 4, dummy);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("   stream whole tensor: %.3f ms = %.0f GB/s\n", ms, nfl * 4 / ms / 1e6);
        }
        for (int g : grans) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
            size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            printf("   granularity %d (set: %s, now %zu)\n", g, cudaGetErrorString(e), got);
            const float t0 = run_pick<0>(lp, lab, c.U, c.V, out, c.cells, flush, fl4);
            const float t1 = run_pick<1>(lp, lab, c.U, c.V, out, c.cells, flush, fl4);
            const float t2 = run_pick<2>(lp, lab, c.U, c.V, out, c.cells, flush, fl4);
            const float t3 = run_pick<3>(lp, lab, c.U, c.V, out, c.cells, flush, fl4);
            const float t4 = run_pick<4>(lp, lab, c.U, c.V, out, c.cells, flush, fl4);
            const float t5 = run_pick<5>(lp, lab, c.U, c.V, out, c.cells, flush, fl4);
            printf("      nc %.3f ms | nc.L2::64B %.3f | nc.L2::128B %.3f | cg %.3f | L1::no_allocate %.3f | nc evict_first %.3f   (tensor-equivalent %.0f GB/s for nc)\n",
                   t0, t1, t2, t3, t4, t5, nfl * 4 / t0 / 1e6);
        }
        cudaFree(lp); cudaFree(lab); cudaFree(out); cudaFree(dummy); free(h);
    }
    return 0;
}
