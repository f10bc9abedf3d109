// wavefront.cu -- general (any-size) path: prefix sums, log-prob gather, the alpha/beta wavefront
// and the (cells,2) gradient kernel.
//
// What it replaces in the reference (/root/reference):
//   k_prefix      : host-side cumsum + 4 .item() syncs        pytorch_binding/binding.cpp:132-158
//   k_gather      : kernel_fill_gather                        core_compact.cu:403-436
//                   python-level gather (index tensor + torch.gather)   warp_rnnt/__init__.py:118-128
//   k_wavefront   : kernel_warp_alphas / kernel_warp_betas + the counts scheduler
//                   core.cu:41-258, core_gather.cu:37-246, core_compact.cu:29-269
//                   kernel_fill_costs (+ mismatch guard)      core.cu:334-370, core_compact.cu:347-358
//   k_grads_pairs : kernel_grads_blank / kernel_grads_label   core.cu:260-332, core_compact.cu:271-345
//
// Design (not a port): one CTA per (lattice, direction); alpha and beta CTAs of a lattice form a
// 2-CTA cluster and meet on a cluster barrier for the cost / mismatch guard.  Inside a CTA lane l
// of warp w owns lattice column 32w+l and walks the anti-diagonals: one __shfl_up per step moves
// the cell value to the right-hand neighbour, warps hand their boundary column to the next warp
// through a shared-memory ring (tagged 64-bit slots, no barrier on the dependent chain).  The
// reference instead tiles the lattice into 32x1 blocks ordered by global atomics + __threadfence.
#include "common.cuh"
#include "kernels.cuh"

namespace rnnt {

// ------------------------------------------------------------------------------------------
// k_prefix: exclusive prefix sums of xn*(yn+1) and yn (int64), totals for host validation.
// One block; N is small (batch size).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_prefix(const int *__restrict__ xn, const int *__restrict__ yn, int N,
                                                  int64_t *__restrict__ mem_pref, int64_t *__restrict__ lab_pref,
                                                  int *__restrict__ totals) {
    __shared__ int64_t s_mem[32], s_lab[32];
    __shared__ int s_tm[32], s_um[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (N + blockDim.x - 1) / blockDim.x;
    const int lo = min(tid * per, N), hi = min(lo + per, N);
    int64_t sm = 0, sl = 0;
    int tm = 0, um = 0;
    for (int i = lo; i < hi; ++i) {
        const int x = xn[i], y = yn[i];
        sm += (int64_t)x * (y + 1);
        sl += y;
        tm = max(tm, x);
        um = max(um, y + 1);
    }
    // inclusive warp scan of (sm, sl), warp max of (tm, um)
    int64_t im = sm, il = sl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t a = __shfl_up_sync(0xffffffffu, im, o);
        int64_t b = __shfl_up_sync(0xffffffffu, il, o);
        if (lane >= o) { im += a; il += b; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tm = max(tm, __shfl_xor_sync(0xffffffffu, tm, o));
        um = max(um, __shfl_xor_sync(0xffffffffu, um, o));
    }
    if (lane == 31) { s_mem[warp] = im; s_lab[warp] = il; }
    if (lane == 0) { s_tm[warp] = tm; s_um[warp] = um; }
    __syncthreads();
    if (warp == 0) {
        int64_t a = s_mem[lane], b = s_lab[lane];
        const int nw = blockDim.x >> 5;
        if (lane >= nw) { a = 0; b = 0; }
        int64_t ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t x = __shfl_up_sync(0xffffffffu, ia, o);
            int64_t y = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += x; ib += y; }
        }
        s_mem[lane] = ia - a;  // exclusive over warps
        s_lab[lane] = ib - b;
        int t2 = lane < nw ? s_tm[lane] : 0, u2 = lane < nw ? s_um[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t2 = max(t2, __shfl_xor_sync(0xffffffffu, t2, o));
            u2 = max(u2, __shfl_xor_sync(0xffffffffu, u2, o));
        }
        if (lane == 31 && totals) {
            totals[0] = (int)min(ia, (int64_t)INT32_MAX);
            totals[1] = (int)min(ib, (int64_t)INT32_MAX);
            totals[2] = t2;
            totals[3] = u2;
        }
    }
    __syncthreads();
    int64_t rm = s_mem[warp] + (im - sm), rl = s_lab[warp] + (il - sl);  // exclusive prefix of this thread
    for (int i = lo; i < hi; ++i) {
        mem_pref[i] = rm;
        lab_pref[i] = rl;
        rm += (int64_t)xn[i] * (yn[i] + 1);
        rl += yn[i];
    }
}

// ------------------------------------------------------------------------------------------
// k_gather: pairs[cell] = (lp[cell, blank], lp[cell, label(u)]) for valid cells; optional loc.
// grid (gx, N); every thread handles kIlp cells with all loads issued before the stores.
// ------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherIlp = 4;

__global__ void __launch_bounds__(kGatherThreads)
k_gather(Problem p, const float *__restrict__ lp, const int *__restrict__ labels, int V, int blank,
         float2 *__restrict__ pairs, int64_t *__restrict__ loc) {
    const int n = blockIdx.y;
    const Lattice L = get_lattice(p, n);
    if (!L.ok) return;
    const int cells = p.compact ? L.Tn * L.Un : L.Tn * L.stride;  // dense: rows [0,Tn) incl. padded columns
    for (int c0 = (blockIdx.x * kGatherThreads) * kGatherIlp; c0 < cells; c0 += gridDim.x * kGatherThreads * kGatherIlp) {
        float vb[kGatherIlp], vl[kGatherIlp];
        int lab[kGatherIlp];
        int64_t cell[kGatherIlp];
        bool ok[kGatherIlp];
#pragma unroll
        for (int k = 0; k < kGatherIlp; ++k) {
            const int r = c0 + k * kGatherThreads + threadIdx.x;
            ok[k] = r < cells;
            const int t = ok[k] ? r / L.stride : 0;
            const int u = ok[k] ? r - t * L.stride : 0;
            ok[k] = ok[k] && (u < L.Un);
            cell[k] = L.base + r;
            // the last column has no label transition; the reference's compact gather stores the
            // blank there (core_compact.cu:424-431)
            lab[k] = (ok[k] && u < L.Un - 1) ? labels[L.lab_base + u] : blank;
        }
#pragma unroll
        for (int k = 0; k < kGatherIlp; ++k) {
            if (ok[k]) {
                const float *row = lp + cell[k] * (int64_t)V;
                vb[k] = __ldg(row + blank);
                vl[k] = __ldg(row + lab[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < kGatherIlp; ++k) {
            if (ok[k]) {
                pairs[cell[k]] = make_float2(vb[k], vl[k]);
                if (loc) loc[cell[k]] = lab[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_wavefront
// ------------------------------------------------------------------------------------------
constexpr int kRing = 128;     // boundary ring slots per warp boundary
constexpr int kPrefetch = 8;   // steps of log-prob prefetch (registers)

struct __align__(8) Slot { float val; int row; };

__device__ __forceinline__ Slot ld_slot(const Slot *p) {
    Slot s;
    asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];"
                 : "=f"(s.val), "=r"(s.row)
                 : "r"((uint32_t)__cvta_generic_to_shared(p))
                 : "memory");
    return s;
}
__device__ __forceinline__ void st_slot(Slot *p, float v, int row) {
    asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v),
                 "r"(row)
                 : "memory");
}
__device__ __forceinline__ int ld_vol_s32(const int *p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol_s32(int *p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// One direction of one lattice.  "Primed" coordinates (i,j): alpha uses (t,u); beta uses
// (Tn-1-t, Un-1-u), so both are the same recurrence
//   val[i,j] = LSE(val[i-1,j] + wB(i,j), val[i,j-1] + wL(i,j)),  val[0,0] = init
// alpha: wB = blank[i-1,j], wL = label[i,j-1], init 0          (core.cu:80-134)
// beta : wB = blank[t,u],   wL = label[t,u] at the own cell, init = blank[T-1,U-1]  (core.cu:171-239)
template <int KIND, bool BETA>
__device__ void wavefront_dir(const Lattice &L, const float2 *__restrict__ pairs, float *__restrict__ out,
                              Slot *ring, int *cons, float *ll_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int Tn = L.Tn, Un = L.Un, st = L.stride;
    const int T1 = Tn - 1, U1 = Un - 1;
    const float2 *pr = pairs + L.base;
    float *o = out + L.base;
    const int cols_per_pass = 32 * nwarps;

    for (int col0 = 0; col0 < Un; col0 += cols_per_pass) {
        if (col0 > 0) __syncthreads();  // previous pass complete (its last column is our left boundary)
        // reset ring tags / consumer counters for this pass
        for (int k = threadIdx.x; k < nwarps * kRing; k += blockDim.x) ring[k].row = -1;
        if (threadIdx.x < nwarps) cons[threadIdx.x] = 0;
        __syncthreads();

        const int wcol = col0 + 32 * warp;
        if (wcol >= Un) continue;  // warp idle in this pass (still takes part in the barriers above)
        const int j = wcol + lane;
        const bool col_ok = j < Un;
        const int active = min(32, Un - wcol);
        const int nsteps = Tn + active - 1;
        const bool has_next = (wcol + 32 < Un) && (warp + 1 < nwarps);  // we publish our lane-31 column
        const bool from_ring = warp > 0;                                // lane 0's left neighbour is in the ring
        const bool from_gmem = (warp == 0 && col0 > 0);                 // ... or in global memory (previous pass)
        Slot *ring_in = ring + (size_t)(warp - 1) * kRing;
        Slot *ring_out = ring + (size_t)warp * kRing;

        // Exact mode, column 0: the reference builds it with a 32-wide Kogge-Stone scan per tile
        // plus the tile's base value (core.cu:92-110 / :197-215); reproduce that summation order
        // so alpha/beta are bit-identical.  Done up front by warp 0; lane 0 then reads it back.
        const bool col0_scan = (KIND != kFast) && (wcol == 0);
        if (col0_scan) {
            float base = BETA ? pr[(int64_t)T1 * st + U1].x : 0.0f;
            if (lane == 0) o[BETA ? ((int64_t)T1 * st + U1) : 0] = base;
            for (int p0 = 0; p0 < T1; p0 += 32) {
                const int i = p0 + lane + 1;  // row computed by this lane
                float b = 0.0f;
                if (i <= T1) b = BETA ? pr[(int64_t)(T1 - i) * st + U1].x : pr[(int64_t)(i - 1) * st].x;
#pragma unroll
                for (int k = 1; k < 32; k <<= 1) {
                    const float a = __shfl_up_sync(0xffffffffu, b, k);
                    if (k <= lane) b += a;
                }
                const float v = base + b;
                if (i <= T1) o[BETA ? ((int64_t)(T1 - i) * st + U1) : ((int64_t)i * st)] = v;
                base = __shfl_sync(0xffffffffu, v, 31);
            }
            __syncwarp();
        }

        float val = kNegInf;      // val[i-1, j] (own column, previous row); -inf drops the skip term on row 0
        int cons_seen = 0;        // producer-side cache of the next warp's consumed-row counter
        float wbA[kPrefetch], wlA[kPrefetch], bndA[kPrefetch];
        float wbB[kPrefetch], wlB[kPrefetch], bndB[kPrefetch];

        // log-probs (and memory-resident boundary values) for steps [s0, s0+kPrefetch)
        auto load_chunk = [&](float (&wb)[kPrefetch], float (&wl)[kPrefetch], float (&bnd)[kPrefetch], int s0) {
#pragma unroll
            for (int k = 0; k < kPrefetch; ++k) {
                const int i = s0 + k - lane;
                wb[k] = 0.0f;
                wl[k] = 0.0f;
                bnd[k] = kNegInf;
                if (col_ok && i >= 0 && i < Tn) {
                    if (BETA) {
                        const float2 w = pr[(int64_t)(T1 - i) * st + (U1 - j)];
                        wb[k] = w.x;
                        wl[k] = w.y;
                    } else {
                        if (i >= 1) wb[k] = pr[(int64_t)(i - 1) * st + j].x;
                        if (j >= 1) wl[k] = pr[(int64_t)i * st + (j - 1)].y;
                    }
                    if (lane == 0) {
                        if (from_gmem) bnd[k] = __ldcg(&o[BETA ? ((int64_t)(T1 - i) * st + (U1 - (j - 1))) : ((int64_t)i * st + (j - 1))]);
                        if (col0_scan && i >= 1) bnd[k] = __ldcg(&o[BETA ? ((int64_t)(T1 - i) * st + U1) : ((int64_t)i * st)]);
                    }
                }
            }
        };

        auto run_chunk = [&](const float (&wb)[kPrefetch], const float (&wl)[kPrefetch], const float (&bnd)[kPrefetch], int s0) {
#pragma unroll
            for (int k = 0; k < kPrefetch; ++k) {
                const int s = s0 + k;
                if (s >= nsteps) break;           // warp-uniform
                const int i = s - lane;
                float left = __shfl_up_sync(0xffffffffu, val, 1);
                if (lane == 0) {
                    left = kNegInf;
                    if (from_gmem) left = bnd[k];
                }
                if (from_ring && s < Tn) {        // warp-uniform: lane 0's row s needs the boundary value
                    Slot sl = ld_slot(&ring_in[s & (kRing - 1)]);
                    while (sl.row != s) {
                        __nanosleep(20);
                        sl = ld_slot(&ring_in[s & (kRing - 1)]);
                    }
                    if (lane == 0) {
                        left = sl.val;
                        st_vol_s32(&cons[warp], s + 1);
                    }
                }
                const bool act = col_ok && i >= 0 && i < Tn;
                if (act) {
                    float v;
                    if (i == 0 && j == 0) {
                        v = BETA ? wb[k] : 0.0f;                      // beta[T-1,U-1] = blank there; alpha[0,0] = 0
                    } else if (col0_scan && lane == 0) {
                        v = bnd[k];                                   // column 0 from the scan pre-pass
                    } else {
                        const float skip = val + wb[k];
                        const float emit = left + wl[k];
                        if (i == 0) v = emit;                         // first row: label transitions only (core.cu:80-90)
                        else if (j == 0) v = skip;                    // first column: blank transitions only
                        else v = lse<KIND>(skip, emit);
                    }
                    val = v;
                    o[BETA ? ((int64_t)(T1 - i) * st + (U1 - j)) : ((int64_t)i * st + j)] = v;
                }
                if (has_next && s >= 31 && s - 31 < Tn) {             // warp-uniform: lane 31 finished row s-31
                    const int row = s - 31;
                    if (row - kRing >= cons_seen) {                   // ring slot still unread? wait for the consumer
                        int c = ld_vol_s32(&cons[warp + 1]);
                        while (row - kRing >= c) {
                            __nanosleep(20);
                            c = ld_vol_s32(&cons[warp + 1]);
                        }
                        cons_seen = c;
                    }
                    if (lane == 31) st_slot(&ring_out[row & (kRing - 1)], val, row);
                }
            }
        };

        // double-buffered: the loads of the next chunk are in flight while this chunk's chain runs
        load_chunk(wbA, wlA, bndA, 0);
        for (int s0 = 0; s0 < nsteps; s0 += 2 * kPrefetch) {
            load_chunk(wbB, wlB, bndB, s0 + kPrefetch);
            run_chunk(wbA, wlA, bndA, s0);
            load_chunk(wbA, wlA, bndA, s0 + 2 * kPrefetch);
            run_chunk(wbB, wlB, bndB, s0 + kPrefetch);
        }
        // the thread that owns the last cell reports the log-likelihood seen from this direction
        if (j == U1 && ll_out != nullptr) {
            // alpha side: alpha[T-1,U-1] + blank[T-1,U-1] (core.cu:346) ; beta side: beta[0,0]
            *ll_out = BETA ? val : val + pr[(int64_t)T1 * st + U1].x;
        }
    }
}

// grid (2, N) with cluster (2,1,1): rank 0 = alpha, rank 1 = beta (beta_only: grid (1,N), no cluster).
// ws_ll: (2,N) floats {alpha-side ll, beta-side ll}; bad: (N) ints (1 = mismatch guard fired).
template <int KIND>
__global__ void __launch_bounds__(512) k_wavefront(Problem p, const float2 *__restrict__ pairs,
                                                    float *__restrict__ alphas, float *__restrict__ betas,
                                                    float *__restrict__ ws_ll, int *__restrict__ bad,
                                                    float *__restrict__ costs, int beta_only, int guard) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5;
    Slot *ring = reinterpret_cast<Slot *>(smem_raw);
    int *cons = reinterpret_cast<int *>(smem_raw + sizeof(Slot) * (size_t)nwarps * kRing);
    const int n = blockIdx.y;
    const bool is_beta = beta_only || (blockIdx.x == 1);
    const Lattice L = get_lattice(p, n);
    float *ll = ws_ll + (is_beta ? p.N : 0) + n;
    if (L.ok) {
        if (is_beta) wavefront_dir<KIND, true>(L, pairs, betas, ring, cons, ll);
        else wavefront_dir<KIND, false>(L, pairs, alphas, ring, cons, ll);
    }
    if (!beta_only) {
        // alpha and beta CTAs of a lattice meet here; release/acquire orders the ll writes.
        __threadfence();
        cluster_arrive_release();
        cluster_wait_acquire();
    } else {
        __syncthreads();
    }
    if (is_beta && threadIdx.x == 0) {
        float cost = NAN;
        int isbad = 0;
        if (L.ok) {
            float b = __ldcg(ws_ll + p.N + n);
            if (!beta_only && guard) {
                // forward/backward mismatch guard, core.cu:346-367
                const float a = __ldcg(ws_ll + n);
                const float ratio = fabsf(a - b) / fabsf(fmaxf(a, b));
                if (ratio > 0.001f) {
                    printf("\nWARNING: sample %d [%d, %d] has a forward/backward mismatch %f / %f\n", n, L.Tn,
                           L.Un - 1, a, b);
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

// ------------------------------------------------------------------------------------------
// k_grads_pairs: out[cell] = (blank grad, label grad) for every cell of the layout (zeros on
// padding and on the structural zeros).  Elementwise, all SMs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cell_grads(const Lattice &L, const float2 *__restrict__ pairs,
                                             const float *__restrict__ alphas, const float *__restrict__ betas,
                                             int t, int u, float b00, float lam_scale_dummy, float fastemit_lambda) {
    (void)lam_scale_dummy;
    const int64_t c = L.base + (int64_t)t * L.stride + u;
    const float2 w = pairs[c];
    const float al = alphas[c];
    float gb = 0.0f, gl = 0.0f;
    const bool last_t = (t == L.Tn - 1), last_u = (u == L.Un - 1);
    if (!(last_t && !last_u)) {
        // core.cu:284-294
        float a = al;
        if (!last_t) a += betas[c + L.stride];
        a = expf(a + w.x - b00);
        gb = -a;
    }
    if (!last_u) {
        // core.cu:319-331 ; (1. + lambda) * a is a double multiply in the reference
        float a = al + betas[c + 1];
        a = expf(a + w.y - b00);
        a = (float)((1.0 + (double)fastemit_lambda) * (double)a);
        gl = -a;
    }
    return make_float2(gb, gl);
}

constexpr int kGradThreads = 256;

__global__ void __launch_bounds__(kGradThreads)
k_grads_pairs(Problem p, const float2 *__restrict__ pairs, const float *__restrict__ alphas,
              const float *__restrict__ betas, const int *__restrict__ bad, float fastemit_lambda,
              float2 *__restrict__ out) {
    const int n = blockIdx.y;
    const Lattice L = get_lattice(p, n);
    const int total = p.compact ? (L.ok ? L.Tn * L.Un : 0) : p.T * p.U;  // dense: the whole padded slab
    const int64_t slab = p.compact ? L.base : (int64_t)n * p.T * p.U;
    const int stride = p.compact ? L.stride : p.U;
    const bool live = L.ok && !(bad && bad[n]);
    const float b00 = live ? betas[L.base] : 0.0f;
    for (int r = blockIdx.x * kGradThreads + threadIdx.x; r < total; r += gridDim.x * kGradThreads) {
        const int t = r / stride, u = r - t * stride;
        float2 g = make_float2(0.0f, 0.0f);
        if (live && t < L.Tn && u < L.Un) g = cell_grads(L, pairs, alphas, betas, t, u, b00, 0.0f, fastemit_lambda);
        out[slab + r] = g;
    }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_prefix(cudaStream_t s, const int *xn, const int *yn, int N, int64_t *mem_pref,
                          int64_t *lab_pref, int *totals) {
    k_prefix<<<1, 1024, 0, s>>>(xn, yn, N, mem_pref, lab_pref, totals);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_gather(cudaStream_t s, const Problem &p, const float *lp, const int *labels, int V, int blank,
                          float2 *pairs, int64_t *loc, int64_t cells_hint) {
    // cells per lattice: dense T*U; compact unknown on the host -> average * 2, grid-stride covers the rest
    int64_t per = p.compact ? (cells_hint / (p.N > 0 ? p.N : 1)) * 2 + 1 : (int64_t)p.T * p.U;
    int gx = (int)((per + kGatherThreads * kGatherIlp - 1) / (kGatherThreads * kGatherIlp));
    gx = max(1, min(gx, 4096));
    dim3 grid(gx, p.N);
    k_gather<<<grid, kGatherThreads, 0, s>>>(p, lp, labels, V, blank, pairs, loc);
    count_launch();
    return cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_wavefront_kind(cudaStream_t s, const Problem &p, const float2 *pairs, float *alphas,
                                         float *betas, float *ws_ll, int *bad, float *costs, int beta_only,
                                         int guard, int u_hint) {
    int nwarps = (u_hint + 31) / 32;
    nwarps = max(1, min(nwarps, 16));
    const size_t smem = sizeof(Slot) * (size_t)nwarps * kRing + sizeof(int) * nwarps;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(beta_only ? 1 : 2, p.N, 1);
    cfg.blockDim = dim3(32 * nwarps, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = beta_only ? 1 : 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    return cudaLaunchKernelEx(&cfg, k_wavefront<KIND>, p, pairs, alphas, betas, ws_ll, bad, costs, beta_only, guard);
}

cudaError_t launch_wavefront(cudaStream_t s, int kind, const Problem &p, const float2 *pairs, float *alphas,
                             float *betas, float *ws_ll, int *bad, float *costs, int beta_only, int guard,
                             int u_hint) {
    switch (kind) {
        case kExactDense:
            return launch_wavefront_kind<kExactDense>(s, p, pairs, alphas, betas, ws_ll, bad, costs, beta_only, guard, u_hint);
        case kExactCompact:
            return launch_wavefront_kind<kExactCompact>(s, p, pairs, alphas, betas, ws_ll, bad, costs, beta_only, guard, u_hint);
        default:
            return launch_wavefront_kind<kFast>(s, p, pairs, alphas, betas, ws_ll, bad, costs, beta_only, guard, u_hint);
    }
}

cudaError_t launch_grads_pairs(cudaStream_t s, const Problem &p, const float2 *pairs, const float *alphas,
                               const float *betas, const int *bad, float fastemit_lambda, float2 *out,
                               int64_t cells_hint) {
    int64_t per = p.compact ? (cells_hint / (p.N > 0 ? p.N : 1)) * 2 + 1 : (int64_t)p.T * p.U;
    int gx = (int)((per + kGradThreads - 1) / kGradThreads);
    gx = max(1, min(gx, 8192));
    dim3 grid(gx, p.N);
    k_grads_pairs<<<grid, kGradThreads, 0, s>>>(p, pairs, alphas, betas, bad, fastemit_lambda, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rnnt
