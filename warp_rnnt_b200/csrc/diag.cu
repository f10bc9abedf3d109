// diag.cu -- general path, second generation: the same "C columns per lane, diagonal-major
// operands" wavefront as fused.cu, but with the staged arrays in HBM, for lattices that do not fit
// shared memory (BASELINE configs 4 and 5, the compact layout).
//
//   k_diag_border   sentinel slots of the staged operand arrays (blank 0, label kBig)
//   k_diag_gather   one pass over log_probs: each cell's blank / label log-prob is written to the
//                   slots that consume it (alpha: the target cell's slot, beta: the own cell's slot), + loc
//   k_diag_wavefront  grid (2,N) in 2-CTA clusters (alpha CTA, beta CTA), ONE warp each: a diagonal's
//                   operands are a contiguous row -> cp.async (LDGSTS) ring, 16 bytes per lane; results
//                   leave as coalesced vector stores.  Every warp-level memory access touches a handful
//                   of lines instead of 32 (the row-major k_wavefront is LSU-wavefront bound: 96 lines
//                   per warp-step).  Cost / mismatch guard after barrier.cluster like k_wavefront.
//   k_diag_grads    (cells,2) gradients, threads in beta-diagonal order so all five operand reads coalesce
//
// Replaces, like wavefront.cu: core.cu:41-370, core_gather.cu, core_compact.cu:29-436.
#include "common.cuh"
#include "kernels.cuh"

namespace rnnt {

static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

constexpr float kBigD = -1.0e30f;
constexpr int kRingSteps = 8;          // diagonals of operands in flight (cp.async ring)

// staged layout of lattice n: six planes [nd][Wd]; idxA(t,u) = (t+u)*Wd + u (alpha side),
// idxB(t,u) = ((Tn-1-t) + jp)*Wd + jp with jp = Wd-1-u (beta side, "primed" coordinates)
struct DiagLayout {
    float *WBa, *WLa, *WBb, *WLb, *AL, *BE;   // each N * plane floats
    int64_t plane;                            // nd * Wd
    int Wd, nd;
};

__device__ __forceinline__ int64_t cell_index(const Problem &p, const Lattice &L, int t, int u) {
    return L.base + (int64_t)t * L.stride + u;
}

// ---- sentinels -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diag_border(Problem p, DiagLayout D) {
    const int n = blockIdx.y;
    const Lattice L = get_lattice(p, n);
    const int Tn = L.ok ? L.Tn : 0, Un = L.ok ? L.Un : 0;
    const int T1 = Tn - 1, U1 = Un - 1, Wd = D.Wd;
    const int64_t base = (int64_t)n * D.plane;
    for (int k = blockIdx.x * 256 + threadIdx.x; k < D.plane; k += gridDim.x * 256) {
        const int d = k / Wd, j = k - d * Wd;
        const int i = d - j;                                   // alpha row of this slot
        const bool col = j <= U1;
        if (!(col && i >= 1 && i <= T1)) D.WBa[base + k] = 0.0f;
        if (!(col && j >= 1 && i >= 0 && i <= T1)) D.WLa[base + k] = kBigD;
        const int u = Wd - 1 - j, t = T1 - i;                  // beta: own cell of this slot
        const bool cell = (u >= 0 && u <= U1 && t >= 0 && t <= T1);
        if (!cell) D.WBb[base + k] = 0.0f;
        if (!(cell && u < U1)) D.WLb[base + k] = kBigD;
    }
}

// ---- gather ----------------------------------------------------------------------------------
constexpr int kDGThreads = 256;
constexpr int kDGIlp = 4;

__global__ void __launch_bounds__(kDGThreads)
k_diag_gather(Problem p, DiagLayout D, const float *__restrict__ lp, const int *__restrict__ labels, int V,
              int blank, int pairs_in, int64_t *__restrict__ loc) {
    const int n = blockIdx.y;
    const Lattice L = get_lattice(p, n);
    if (!L.ok) return;
    const int Tn = L.Tn, Un = L.Un, T1 = Tn - 1, U1 = Un - 1, Wd = D.Wd;
    const int cells = Tn * Un;
    const int64_t base = (int64_t)n * D.plane;
    const float inv = 1.0f / (float)Un;
    for (int c0 = blockIdx.x * kDGThreads * kDGIlp; c0 < cells; c0 += gridDim.x * kDGThreads * kDGIlp) {
        float vb[kDGIlp], vl[kDGIlp];
        int tt[kDGIlp], uu[kDGIlp], lab[kDGIlp];
#pragma unroll
        for (int g = 0; g < kDGIlp; ++g) {
            const int c = c0 + g * kDGThreads + threadIdx.x;
            const bool in = c < cells;
            int t = (int)(((float)c + 0.5f) * inv);
            int u = c - t * Un;
            if (u < 0) { --t; u += Un; } else if (u >= Un) { ++t; u -= Un; }
            if (!in) { t = 0; u = 0; }
            tt[g] = in ? t : -1;
            uu[g] = u;
            lab[g] = (in && u < U1 && !pairs_in) ? labels[L.lab_base + u] : blank;
        }
#pragma unroll
        for (int g = 0; g < kDGIlp; ++g) {
            const int64_t cell = cell_index(p, L, tt[g] < 0 ? 0 : tt[g], uu[g]);
            if (pairs_in) {
                const float2 w2 = __ldg(reinterpret_cast<const float2 *>(lp) + cell);
                vb[g] = w2.x;
                vl[g] = w2.y;
            } else {
                const float *row = lp + cell * (int64_t)V;
                vb[g] = __ldg(row + blank);
                vl[g] = __ldg(row + lab[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < kDGIlp; ++g) {
            const int t = tt[g], u = uu[g];
            if (t < 0) continue;
            const int jp = Wd - 1 - u;
            const int64_t ib = base + (int64_t)(T1 - t + jp) * Wd + jp;
            D.WBb[ib] = vb[g];
            if (u < U1) D.WLb[ib] = vl[g];
            if (t < T1) D.WBa[base + (int64_t)(t + 1 + u) * Wd + u] = vb[g];
            if (u < U1) D.WLa[base + (int64_t)(t + u + 1) * Wd + u + 1] = vl[g];
            if (loc) loc[cell_index(p, L, t, u)] = lab[g];     // blank on the last column (core_compact.cu:424-431)
        }
    }
}

// ---- wavefront -------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int C>
__device__ __forceinline__ void lds_c(uint32_t a, float (&v)[C]) {
    if constexpr (C == 1) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(a) : "memory");
    } else if constexpr (C == 2) {
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a) : "memory");
    } else {
#pragma unroll
        for (int q = 0; q < C / 4; ++q)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3])
                         : "r"(a + 16u * q)
                         : "memory");
    }
}
template <int C>
__device__ __forceinline__ void stg_c(float *p, const float (&v)[C]) {
    if constexpr (C == 1) {
        *p = v[0];
    } else if constexpr (C == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < C / 4; ++q)
            reinterpret_cast<float4 *>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

// One direction, one warp.  ring: kRingSteps slots of [2][Wd] floats (blank-edge row, label-edge row).
template <int KIND, int C>
__device__ __forceinline__ void sweep_diag_gmem(const float *__restrict__ wb, const float *__restrict__ wl,
                                                float *__restrict__ out, int Wd, int ndiag, int lane, int first_col,
                                                const float *pre, int pre_rows, float *ring) {
    const bool act = C * lane < Wd;                          // lanes past the staged width do nothing
    float val[C];
#pragma unroll
    for (int c = 0; c < C; ++c) val[c] = (C * lane + c == first_col) ? 0.0f : kBigD;
    const uint32_t ring_a = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t slot_bytes = 8u * (uint32_t)Wd;
    const int nvec = Wd / 4;                                 // 16-byte chunks per row (Wd % 4 == 0 when C >= 4)
    auto issue = [&](int d) {                                // operands of diagonal d -> ring slot d % kRingSteps
        const uint32_t dst = ring_a + (uint32_t)(d % kRingSteps) * slot_bytes;
        const float *sb = wb + (int64_t)d * Wd, *sl = wl + (int64_t)d * Wd;
        if constexpr (C >= 4) {
            for (int q = lane; q < nvec; q += 32) {
                cp_async16(dst + 16u * q, sb + 4 * q);
                cp_async16(dst + 4u * Wd + 16u * q, sl + 4 * q);
            }
        } else {
            for (int q = lane; q < Wd; q += 32) {
                cp_async4(dst + 4u * q, sb + q);
                cp_async4(dst + 4u * Wd + 4u * q, sl + q);
            }
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int d = 0; d < kRingSteps; ++d) issue(d);
    const int l0 = first_col / C, c0 = first_col - l0 * C;
    float *op = out + C * lane;
#pragma unroll 1
    for (int d = 0; d < ndiag; ++d) {
        cp_async_wait<kRingSteps - 1>();                     // diagonal d has landed (groups complete in order)
        __syncwarp();
        float b[C], l[C];
        const uint32_t src = ring_a + (uint32_t)(d % kRingSteps) * slot_bytes + 4u * (uint32_t)(C * lane);
        if (act) {
            lds_c<C>(src, b);
            lds_c<C>(src + 4u * Wd, l);
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) { b[c] = 0.0f; l[c] = kBigD; }
        }
        const float left = __shfl_up_sync(0xffffffffu, val[C - 1], 1);
        float nv[C];
        nv[0] = lse<KIND>(val[0] + b[0], left + l[0]);
#pragma unroll
        for (int c = 1; c < C; ++c) nv[c] = lse<KIND>(val[c] + b[c], val[c - 1] + l[c]);
        if (KIND != kFast) {
            const int i = d - first_col;
            if (lane == l0 && i >= 1 && i < pre_rows) {
                const float pv = pre[i];
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c == c0) nv[c] = pv;
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) val[c] = nv[c];
        if (act) stg_c<C>(op, val);
        op += Wd;
        __syncwarp();                                        // every lane is done with this ring slot
        issue(d + kRingSteps);                               // rows past ndiag are allocated and initialised
    }
    cp_async_wait<0>();
}

template <int KIND, int C>
__global__ void __launch_bounds__(32, 1)
k_diag_wavefront(Problem p, DiagLayout D, float *__restrict__ ws_ll, int *__restrict__ bad, float *__restrict__ costs,
                 int beta_only, int guard, int t_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);                    // [kRingSteps][2][Wd]
    float *pre = ring + (size_t)kRingSteps * 2 * D.Wd;                    // [t_cap] exact-mode column scan
    const int lane = threadIdx.x;
    const int n = blockIdx.y;
    const bool beta = beta_only || (blockIdx.x == 1);
    const Lattice L = get_lattice(p, n);
    const int Tn = L.Tn, Un = L.Un, T1 = Tn - 1, Wd = D.Wd;
    const int64_t base = (int64_t)n * D.plane;
    const bool ok = L.ok && Tn <= t_cap && Un <= Wd;
    if (ok) {
        const int first_col = beta ? (Wd - Un) : 0;
        const int ndiag = beta ? (Tn + Wd - 1) : (Tn + Un - 1);
        const float *wb = (beta ? D.WBb : D.WBa) + base;
        const float *wl = (beta ? D.WLb : D.WLa) + base;
        float *out = (beta ? D.BE : D.AL) + base;
        if (KIND != kFast) {
            // reference-order column scan (core.cu:92-110 / :197-215), see fused.cu
            float bs = beta ? wb[(int64_t)first_col * Wd + first_col] : 0.0f;
            if (lane == 0) pre[0] = bs;
            for (int p0 = 0; p0 < T1; p0 += 32) {
                const int i = p0 + lane + 1;
                float bsum = 0.0f;
                if (i <= T1) bsum = wb[(int64_t)(i + first_col) * Wd + first_col];
#pragma unroll
                for (int k = 1; k < 32; k <<= 1) {
                    const float a = __shfl_up_sync(0xffffffffu, bsum, k);
                    if (k <= lane) bsum += a;
                }
                const float v = bs + bsum;
                if (i <= T1) pre[i] = v;
                bs = __shfl_sync(0xffffffffu, v, 31);
            }
            __syncwarp();
        }
        sweep_diag_gmem<KIND, C>(wb, wl, out, Wd, ndiag, lane, first_col, pre, Tn, ring);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            // alpha side: alpha[T-1,U-1] + blank[T-1,U-1] (core.cu:346) ; beta side: beta[0,0]
            const int U1 = Un - 1;
            const int jp0 = Wd - Un, jpL = Wd - 1;
            float v;
            if (beta) v = __ldcg(D.BE + base + (int64_t)(T1 + jpL) * Wd + jpL);
            else v = __ldcg(D.AL + base + (int64_t)(T1 + U1) * Wd + U1) + __ldcg(D.WBb + base + (int64_t)jp0 * Wd + jp0);
            ws_ll[(beta ? p.N : 0) + n] = v;
        }
    }
    if (!beta_only) {
        __threadfence();
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    if (beta && lane == 0) {
        float cost = NAN;
        int isbad = 0;
        if (ok) {
            float b = __ldcg(ws_ll + p.N + n);
            if (!beta_only && guard) {
                const float a = __ldcg(ws_ll + n);
                const float ratio = fabsf(a - b) / fabsf(fmaxf(a, b));
                if (ratio > 0.001f) {
                    printf("\nWARNING: sample %d [%d, %d] has a forward/backward mismatch %f / %f\n", n, Tn, Un - 1, a, b);
                    b = (a + b) / 2.0f;
                    isbad = 1;
                }
            }
            cost = -b;
        } else {
            isbad = 1;
        }
        costs[n] = cost;
        if (bad) bad[n] = isbad;
    }
}

// ---- (cells,2) gradients ---------------------------------------------------------------------
// blocks [0, gslots): one thread per beta-diagonal slot (d', jp) -> coalesced reads of BE (own, one
// diagonal earlier same / previous column), WBb, WLb and of AL along its own diagonal; the pair is
// written at the cell's row-major position.  blocks [gslots, ...): dense layout only, zero the padded
// cells of the (N,T,U,2) output (disjoint from the first part, no ordering needed).
__global__ void __launch_bounds__(256)
k_diag_grads(Problem p, DiagLayout D, const int *__restrict__ bad, float fastemit_lambda, float2 *__restrict__ out,
             int gslots) {
    const int n = blockIdx.y;
    const Lattice L = get_lattice(p, n);
    const int Wd = D.Wd;
    const int64_t base = (int64_t)n * D.plane;
    const bool live = L.ok && !(bad && bad[n]);
    if ((int)blockIdx.x < gslots) {
        if (!L.ok) return;
        const int Tn = L.Tn, Un = L.Un, T1 = Tn - 1, U1 = Un - 1;
        const int jpL = Wd - 1;
        const float b00 = live ? D.BE[base + (int64_t)(T1 + jpL) * Wd + jpL] : 0.0f;
        const int nslots = (Tn + Wd) * Wd;
        const bool has_lam = fastemit_lambda != 0.0f;
        for (int k = blockIdx.x * 256 + threadIdx.x; k < nslots; k += gslots * 256) {
            const int d = k / Wd, jp = k - d * Wd;
            const int u = Wd - 1 - jp, t = T1 - (d - jp);
            if (u < 0 || u > U1 || t < 0 || t > T1) continue;
            float gb = 0.0f, gl = 0.0f;
            if (live) {
                const int64_t ib = base + k;
                const float al = D.AL[base + (int64_t)(t + u) * Wd + u];
                const bool last_t = (t == T1), last_u = (u == U1);
                if (!(last_t && !last_u)) {                    // core.cu:284-294
                    float a = al;
                    if (!last_t) a += D.BE[ib - Wd];
                    a = expf(a + D.WBb[ib] - b00);
                    gb = -a;
                }
                if (!last_u) {                                 // core.cu:319-331
                    float a = al + D.BE[ib - Wd - 1];
                    a = expf(a + D.WLb[ib] - b00);
                    if (has_lam) a = (float)((1.0 + (double)fastemit_lambda) * (double)a);
                    gl = -a;
                }
            }
            out[cell_index(p, L, t, u)] = make_float2(gb, gl);
        }
    } else if (!p.compact) {
        const int total = p.T * p.U;
        const int Tn = L.ok ? L.Tn : 0, Un = L.ok ? L.Un : 0;
        const int64_t slab = (int64_t)n * total;
        const int nb = gridDim.x - gslots;
        for (int r = (blockIdx.x - gslots) * 256 + threadIdx.x; r < total; r += nb * 256) {
            const int t = r / p.U, u = r - t * p.U;
            if (t >= Tn || u >= Un) out[slab + r] = make_float2(0.0f, 0.0f);
        }
    }
}

// ---- host side -------------------------------------------------------------------------------
static int pick_c(int U) {
    if (U <= 32) return 1;
    if (U <= 64) return 2;
    if (U <= 128) return 4;
    if (U <= 256) return 8;
    if (U <= 384) return 12;
    if (U <= 512) return 16;
    return 0;
}

bool diag_plan(int N, int t_max, int u_max, DiagPlan *plan) {
    if (N < 1 || t_max < 1 || u_max < 1) return false;
    const int C = pick_c(u_max);
    if (C == 0) return false;
    int Wd = (u_max + C - 1) / C * C;
    if (C < 4) Wd = (Wd + 3) / 4 * 4;                        // keep rows 16-byte aligned
    const int nd = t_max + Wd + kRingSteps + 1;
    const size_t smem = sizeof(float) * ((size_t)kRingSteps * 2 * Wd + (size_t)t_max) + 64;
    if (smem > 200 * 1024) return false;
    plan->C = C; plan->Wd = Wd; plan->nd = nd; plan->t_cap = t_max; plan->smem = smem;
    plan->plane = (int64_t)nd * Wd;
    plan->scratch_bytes = sizeof(float) * 6 * (size_t)N * (size_t)plan->plane;
    return true;
}

template <int KIND, int C>
static cudaError_t launch_wf(cudaStream_t s, const Problem &p, const DiagLayout &D, const DiagPlan &plan, float *ws_ll,
                             int *bad, float *costs, int beta_only, int guard) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_diag_wavefront<KIND, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(beta_only ? 1 : 2, p.N, 1);
    cfg.blockDim = dim3(32, 1, 1);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = beta_only ? 1 : 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    return cudaLaunchKernelEx(&cfg, k_diag_wavefront<KIND, C>, p, D, ws_ll, bad, costs, beta_only, guard, plan.t_cap);
}

template <int KIND>
static cudaError_t launch_wf_kind(cudaStream_t s, const Problem &p, const DiagLayout &D, const DiagPlan &plan,
                                  float *ws_ll, int *bad, float *costs, int beta_only, int guard) {
    switch (plan.C) {
        case 1: return launch_wf<KIND, 1>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard);
        case 2: return launch_wf<KIND, 2>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard);
        case 4: return launch_wf<KIND, 4>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard);
        case 8: return launch_wf<KIND, 8>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard);
        case 12: return launch_wf<KIND, 12>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard);
        default: return launch_wf<KIND, 16>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard);
    }
}

// gather + wavefront (+ gradients in (cells,2) form when pg != nullptr) on the diagonal-major path.
// scratch: plan.scratch_bytes of device memory.
cudaError_t launch_diag_forward(cudaStream_t s, int kind, const Problem &p, const DiagPlan &plan, void *scratch,
                                const float *lp, const int *labels, int V, int blank, int pairs_in, int64_t *loc,
                                float *ws_ll, int *bad, float *costs, float2 *pg, float fastemit_lambda, int guard) {
    DiagLayout D;
    float *f = reinterpret_cast<float *>(scratch);
    const int64_t arr = (int64_t)p.N * plan.plane;
    D.WBa = f; D.WLa = f + arr; D.WBb = f + 2 * arr; D.WLb = f + 3 * arr; D.AL = f + 4 * arr; D.BE = f + 5 * arr;
    D.plane = plan.plane; D.Wd = plan.Wd; D.nd = plan.nd;
    cudaError_t e;
    {
        const int gx = (int)imin64((plan.plane + 255) / 256, 2048);
        k_diag_border<<<dim3(gx, p.N), 256, 0, s>>>(p, D);
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    {
        const int64_t per = (int64_t)plan.t_cap * plan.Wd;
        const int gx = (int)imin64((per + kDGThreads * kDGIlp - 1) / (kDGThreads * kDGIlp), 4096);
        k_diag_gather<<<dim3(max(gx, 1), p.N), kDGThreads, 0, s>>>(p, D, lp, labels, V, blank, pairs_in, loc);
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    const int beta_only = pg == nullptr;
    switch (kind) {
        case kExactDense: e = launch_wf_kind<kExactDense>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard); break;
        case kExactCompact: e = launch_wf_kind<kExactCompact>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard); break;
        default: e = launch_wf_kind<kFast>(s, p, D, plan, ws_ll, bad, costs, beta_only, guard); break;
    }
    if (e != cudaSuccess) return e;
    if (pg) {
        const int64_t slots = (int64_t)(plan.t_cap + plan.Wd) * plan.Wd;
        const int gslots = (int)imin64((slots + 255) / 256, 4096);
        const int gpad = p.compact ? 0 : (int)imin64(((int64_t)p.T * p.U + 255) / 256, 1024);
        k_diag_grads<<<dim3(gslots + gpad, p.N), 256, 0, s>>>(p, D, bad, fastemit_lambda, pg, gslots);
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace rnnt
