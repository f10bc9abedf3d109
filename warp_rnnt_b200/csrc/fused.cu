// fused.cu -- single-kernel loss + gradient for lattices that fit one SM's shared memory
// ((T+U)*U <~ 9k cells: BASELINE configs 1-3).  One launch does everything the reference spreads
// over zeros_like + 4 kernels (+ python gather / mul_):
//
//   phase 0  gather      all warps stage the lattice's blank/label log-probs into shared memory in
//                        DIAGONAL-MAJOR, target-indexed form (replaces the python-level gather,
//                        __init__.py:118-128, and the strided loads inside kernel_warp, core.cu:115-120)
//   phase 1  wavefront   ONE warp runs alpha and one runs beta: lane l owns C adjacent lattice columns,
//                        every anti-diagonal is one step = one __shfl_up + C independent LSE chains,
//                        operands and results move as C-wide vector LDS/STS (conflict-free, because a
//                        diagonal's cells are contiguous in the staged layout).  No inter-warp hand-off,
//                        no barrier, no atomics on the recurrence (replaces kernel_warp + the
//                        global-memory counts scheduler, core.cu:41-258)
//            zero-fill   meanwhile the other 14 warps stream zeros over this CTA's slice of the dense
//                        gradient with 256-bit evict_last stores (replaces at::zeros_like, binding.cpp:58);
//                        the wavefront warps join through a shared work counter when they finish
//   phase 2  cost/guard  kernel_fill_costs (core.cu:334-370)
//            patch       the <= 2 non-zeros per row are written into the freshly zeroed, still L2-resident
//                        lines (kernel_grads_blank/label, core.cu:260-332), or -- MODE 1 -- the gradients
//                        are emitted in (N,T,U,2) form for the deferred dense backward.
//
// grid (S, N): S CTAs per lattice recompute the (cheap) wavefront redundantly and split the
// (bandwidth-bound) fill/patch of the lattice's rows, so small batches still use every SM.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace rnnt {

constexpr int kFusedThreads = 512;
constexpr int kGatherWarps = 12;    // MODE 0: warps that stage log-probs; the other 4 start the zero-fill at once
constexpr float kBigF = -1.0e30f;   // finite stand-in for -inf (see wavefront.cu)

// L2 residency control (B200: 126 MB L2).  The dense gradient slab is zero-filled while the
// wavefront runs and patched afterwards; the patch is a partial-sector write, so it must still HIT
// in L2 or ECC forces a DRAM read-modify-write per touched sector.  Filled lines are therefore
// stored evict_last (256-bit STG.E.ELL2.256), and the one-pass log-prob gather is loaded
// evict_first so that it cannot displace them.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float ldg_hint(const float *ptr, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ldg_hint2(const float2 *ptr, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_zero256_evict_last(float *ptr) {
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(ptr), "r"(0) : "memory");
}

// C-wide shared-memory vector load / store (C in {1,2,4,8}); asm volatile keeps program order so
// the operand prefetch stays where it is written.
template <int C>
__device__ __forceinline__ void lds_vec(uint32_t a, float (&v)[C]) {
    if constexpr (C == 1) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(a) : "memory");
    } else if constexpr (C == 2) {
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a) : "memory");
    } else if constexpr (C == 4) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a) : "memory");
    } else {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a + 16u) : "memory");
    }
}
template <int C>
__device__ __forceinline__ void sts_vec(uint32_t a, const float (&v)[C]) {
    if constexpr (C == 1) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v[0]) : "memory");
    } else if constexpr (C == 2) {
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v[0]), "f"(v[1]) : "memory");
    } else if constexpr (C == 4) {
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else {
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 16u), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    }
}

struct FusedArgs {
    const float *lp;        // dense (N,T,U,V) or, pairs_in, (N,T,U,2)
    const int *labels;      // (N,U-1)
    const int *xn, *yn;
    float *costs;           // (N)
    float *grads;           // MODE 0: dense (N,T,U,V)
    float2 *pair_grads;     // MODE 1: (N,T,U) float2
    const float *scale;     // (N) or null
    int N, T, U, V, blank;
    float lam;
    int pairs_in, guard, slices;
    int Wd;                 // staged row stride (floats) = C * ceil(U / C), even
    int nd;                 // staged rows (diagonals) allocated = T + Wd + 16
    int gw;                 // MODE 0: warps that gather; the rest start the zero-fill immediately
};

// One direction of one lattice, one warp.  Diagonal-major, target-indexed operands:
//   wb[d][j] = weight of the row edge    (i-1,j) -> (i,j)       with i = d - j
//   wl[d][j] = weight of the column edge (i,j-1) -> (i,j)
// and out[d][j] = val[i,j].  Slots outside the lattice hold wb = 0, wl = kBig, which makes the one
// uniform step below reproduce every edge rule of core.cu:64-134 / :171-239 without a branch:
//   row 0 (val starts at kBig -> skip vanishes), first column (wl = kBig -> emit vanishes),
//   cell (0,0) (val of the first real column starts at 0 and wb(0,0) carries the initial value),
//   cells past the last row / column only ever feed other out-of-lattice slots.
// Column j at diagonal d reads its own previous value (row edge) and column j-1's previous value
// (column edge): inside a lane that is a register, across lanes one __shfl_up of the lane's last column.
template <int KIND, int C>
__device__ __forceinline__ void sweep_diag(uint32_t wb, uint32_t wl, uint32_t out, int Wd, int ndiag, int lane,
                                           int first_col, const float *pre, int pre_rows) {
    constexpr int P = (C <= 2) ? 4 : 2;                       // diagonals of operand prefetch
    float val[C];
#pragma unroll
    for (int c = 0; c < C; ++c) val[c] = (C * lane + c == first_col) ? 0.0f : kBigF;
    const uint32_t stride = 4u * (uint32_t)Wd;
    const uint32_t off = 4u * (uint32_t)(C * lane);
    uint32_t a_wb = wb + off, a_wl = wl + off, a_out = out + off;
    float b[P][C], l[P][C];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        lds_vec<C>(a_wb, b[k]);
        lds_vec<C>(a_wl, l[k]);
        a_wb += stride;
        a_wl += stride;
    }
    // exact mode: the first real column is taken from the reference-order prefix scan (core.cu:92-110)
    const int l0 = first_col / C, c0 = first_col - l0 * C;
    for (int d0 = 0; d0 < ndiag; d0 += P) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const float left = __shfl_up_sync(0xffffffffu, val[C - 1], 1);   // lane 0 gets its own value: wl = kBig there
            float nv[C];
            nv[0] = lse<KIND>(val[0] + b[k][0], left + l[k][0]);
#pragma unroll
            for (int c = 1; c < C; ++c) nv[c] = lse<KIND>(val[c] + b[k][c], val[c - 1] + l[k][c]);
            if (KIND != kFast) {
                const int i = d0 + k - first_col;              // row of the first real column on this diagonal
                if (lane == l0 && i >= 1 && i < pre_rows) {
                    const float p = pre[i];
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (c == c0) nv[c] = p;
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) val[c] = nv[c];
            sts_vec<C>(a_out, val);
            a_out += stride;
            lds_vec<C>(a_wb, b[k]);                            // operands P diagonals ahead (rows past ndiag are allocated)
            lds_vec<C>(a_wl, l[k]);
            a_wb += stride;
            a_wl += stride;
        }
    }
}

// MODE 0: dense gradients (zero-fill + patch).  MODE 1: (N,T,U,2) gradients.  Both write costs.
template <int KIND, int MODE, int C>
__global__ void __launch_bounds__(kFusedThreads, 1) k_fused(FusedArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y, slice = blockIdx.x;
    const int T = A.T, U = A.U, V = A.V, Wd = A.Wd;
    const int Tn = A.xn[n], Un = A.yn[n] + 1;
    const bool ok = (Tn >= 1 && Tn <= T && Un >= 1 && Un <= U);
    const int T1 = Tn - 1, U1 = Un - 1;

    // ---- shared memory carve-up: six diagonal-major arrays [nd][Wd], prefix-scan columns, labels
    const size_t plane = (size_t)A.nd * Wd;
    float *WBa = reinterpret_cast<float *>(smem_raw);   // alpha: blank edge into (t,u)   at [t+u][u]
    float *WLa = WBa + plane;                           // alpha: label edge into (t,u)   at [t+u][u]
    float *WBb = WLa + plane;                           // beta : blank edge out of (t,u) at [d'][j'], j' = Wd-1-u, d' = (T1-t)+j'
    float *WLb = WBb + plane;                           // beta : label edge out of (t,u) at [d'][j']
    float *AL = WLb + plane;                            // alpha[t,u] at [t+u][u]
    float *BE = AL + plane;                             // beta[t,u]  at [d'][j']
    float *preA = BE + plane;                           // [T] exact-mode column scans
    float *preB = preA + T;
    int *s_lab = reinterpret_cast<int *>(preB + T);     // [U]
    __shared__ int s_next;                              // fill work counter
    __shared__ int s_bad;
    auto idxA = [&](int t, int u) { return (t + u) * Wd + u; };
    auto idxB = [&](int t, int u) { const int jp = Wd - 1 - u; return (T1 - t + jp) * Wd + jp; };

    // rows of the padded slab this CTA fills / patches
    const int t0 = (int)((int64_t)T * slice / A.slices), t1 = (int)((int64_t)T * (slice + 1) / A.slices);
    const int64_t slab = (int64_t)n * T * U;            // first cell of this lattice

    // ---- phase 0: sentinels, labels, gather
    if (tid == 0) { s_next = 0; s_bad = ok ? 0 : 1; }
    if (ok) {
        const int used = (Tn + Wd + 8) * Wd;            // diagonals any sweep or its prefetch can touch
        const int lim = min(used, (int)plane);
        for (int k = tid; k < lim; k += kFusedThreads) {
            WBa[k] = 0.0f; WBb[k] = 0.0f;
            WLa[k] = kBigF; WLb[k] = kBigF;
        }
        if (!A.pairs_in)
            for (int u = tid; u < U1; u += kFusedThreads) s_lab[u] = A.labels[(int64_t)n * (U - 1) + u];
    }
    __syncthreads();
    // MODE 0: warps [0,GW) gather, then warps 0/1 sweep and the rest of them join the zero-fill that
    // warps [GW,32) started right after the barrier above: HBM reads (gather) and writes (fill) overlap.
    // The fill is HBM-write bound (~20 us for cfg 2's 86 MB) and the gather HBM-read bound (~14 us);
    // serialising them costs their sum, overlapping them costs ~the larger.
    const int GW = (MODE == 0) ? A.gw : kFusedThreads / 32;
    const int gthreads = GW * 32;
    // Roles go by a rotated warp index: the two wavefront warps are the HIGHEST-numbered hardware
    // warps (14 and 15, on different sub-partitions).  The SM's issue arbiter favours higher warp
    // ids, so the latency-critical recurrence is never starved by the fill / gather warps sharing its
    // scheduler (with the sweep on warps 0/1 the exact flavour measured anywhere from 65 to 250 us).
    const int lw = (warp + 2) & (kFusedThreads / 32 - 1);
    const int ltid = lw * 32 + lane;
    if (ok && lw < GW) {
        const int cells = Tn * Un;
        const float inv = 1.0f / (float)Un;
        const uint64_t pol_first = policy_evict_first();
        constexpr int G = 4;                            // cells per thread per pass: all loads first
        for (int cb = ltid; cb < cells; cb += gthreads * G) {
            float vb[G], vl[G];
            int tt[G], uu[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c = cb + g * gthreads;
                const bool in = c < cells;
                int t = (int)(((float)c + 0.5f) * inv);
                int u = c - t * Un;
                if (u < 0) { --t; u += Un; } else if (u >= Un) { ++t; u -= Un; }
                if (!in) { t = 0; u = 0; }
                tt[g] = in ? t : -1;
                uu[g] = u;
                const int64_t cell = slab + (int64_t)t * U + u;
                vl[g] = kBigF;
                if (A.pairs_in) {
                    const float2 w2 = ldg_hint2(reinterpret_cast<const float2 *>(A.lp) + cell, pol_first);
                    vb[g] = w2.x;
                    if (u < U1) vl[g] = w2.y;
                } else {
                    const float *row = A.lp + cell * V;
                    vb[g] = ldg_hint(row + A.blank, pol_first);
                    if (u < U1) vl[g] = ldg_hint(row + s_lab[u], pol_first);
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int t = tt[g], u = uu[g];
                if (t >= 0) {
                    const int ib = idxB(t, u);
                    WBb[ib] = vb[g];
                    WLb[ib] = vl[g];                    // kBig on the last column: beta's first column has no column edge
                    if (t < T1) WBa[idxA(t + 1, u)] = vb[g];
                    if (u < U1) WLa[idxA(t, u + 1)] = vl[g];
                }
            }
        }
    }
    if (GW == kFusedThreads / 32) __syncthreads();
    else if (lw < GW) asm volatile("bar.sync 1, %0;" ::"r"(gthreads) : "memory");   // gather warps only

    // ---- phase 1: logical warp 0 = alpha, 1 = beta (then both help filling); other warps zero-fill
    if (ok && lw < 2) {
        const bool beta = (lw == 1);
        const int first_col = beta ? (Wd - Un) : 0;
        const int ndiag = beta ? (Tn + Wd - 1) : (Tn + Un - 1);
        const float *wbp = beta ? WBb : WBa;
        float *pre = beta ? preB : preA;
        if (KIND != kFast) {
            // column 0 in the reference's summation order: 32-wide Kogge-Stone scan per tile + the tile's
            // base (core.cu:92-110 / :197-215), so that exact mode is bit-identical
            float base = beta ? wbp[first_col * Wd + first_col] : 0.0f;
            if (lane == 0) pre[0] = base;
            for (int p0 = 0; p0 < T1; p0 += 32) {
                const int i = p0 + lane + 1;
                float bsum = 0.0f;
                if (i <= T1) bsum = wbp[(i + first_col) * Wd + first_col];   // blank edge into row i of the first column
#pragma unroll
                for (int k = 1; k < 32; k <<= 1) {
                    const float a = __shfl_up_sync(0xffffffffu, bsum, k);
                    if (k <= lane) bsum += a;
                }
                const float v = base + bsum;
                if (i <= T1) pre[i] = v;
                base = __shfl_sync(0xffffffffu, v, 31);
            }
            __syncwarp();
        }
        const uint32_t wb_a = (uint32_t)__cvta_generic_to_shared(wbp);
        const uint32_t wl_a = (uint32_t)__cvta_generic_to_shared(beta ? WLb : WLa);
        const uint32_t out_a = (uint32_t)__cvta_generic_to_shared(beta ? BE : AL);
        sweep_diag<KIND, C>(wb_a, wl_a, out_a, Wd, ndiag, lane, first_col, pre, Tn);
    }
    if (MODE == 0) {
        // zero-fill rows [t0,t1) of this lattice's slab: floats [f0,f1); 32 KB chunks handed out by
        // s_next; 256-bit evict_last stores on the 32-byte aligned interior
        const int64_t f0 = (slab + (int64_t)t0 * U) * V, f1 = (slab + (int64_t)t1 * U) * V;
        float *g = A.grads;
        const bool vec = ((reinterpret_cast<uintptr_t>(g) & 31u) == 0);
        const int64_t a0 = vec ? min(f1, (f0 + 7) & ~(int64_t)7) : f1;
        const int64_t a1 = vec ? max(a0, f1 & ~(int64_t)7) : f1;
        constexpr int kChunk = 8192;                    // floats per chunk
        const int64_t nchunks = (a1 - a0 + kChunk - 1) / kChunk;
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_next, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= nchunks) break;
            const int64_t b = a0 + (int64_t)c * kChunk;
            const int64_t e = min(a1, b + kChunk);
#pragma unroll 4
            for (int64_t f = b + 8 * lane; f < e; f += 256) stg_zero256_evict_last(g + f);
        }
        if (lw == kFusedThreads / 32 - 1) {             // unaligned head / tail (<= 7 floats each, or all if !vec)
            for (int64_t f = f0 + lane; f < a0; f += 32) g[f] = 0.0f;
            for (int64_t f = a1 + lane; f < f1; f += 32) g[f] = 0.0f;
        }
    }
    __syncthreads();

    // ---- phase 2: cost (+ mismatch guard, core.cu:346-367), gradients
    if (tid == 0) {
        float cost = NAN;
        if (ok) {
            float b = BE[idxB(0, 0)];                   // beta[0,0]
            if (A.guard) {
                const float a = AL[idxA(T1, U1)] + WBb[idxB(T1, U1)];   // alpha-side ll (core.cu:346)
                const float ratio = fabsf(a - b) / fabsf(fmaxf(a, b));
                if (ratio > 0.001f) {
                    if (slice == 0)
                        printf("\nWARNING: sample %d [%d, %d] has a forward/backward mismatch %f / %f\n", n, Tn, Un - 1, a, b);
                    b = (a + b) / 2.0f;
                    s_bad = 1;
                }
            }
            cost = -b;
        }
        if (slice == 0) A.costs[n] = cost;
    }
    __syncthreads();
    const bool live = ok && !s_bad;
    const float b00 = live ? BE[idxB(0, 0)] : 0.0f;
    const float sc = A.scale ? A.scale[n] : 1.0f;
    const bool has_lam = (A.lam != 0.0f);
    auto cell_grad = [&](int t, int u) -> float2 {
        // same operation order as core.cu:284-294 and :319-331
        const int ib = idxB(t, u);
        const float a0v = AL[idxA(t, u)];
        float gb = 0.0f, gl = 0.0f;
        const bool last_t = (t == T1), last_u = (u == U1);
        if (!(last_t && !last_u)) {
            float a = a0v;
            if (!last_t) a += BE[ib - Wd];              // beta[t+1,u]: one diagonal earlier, same column
            a = expf(a + WBb[ib] - b00);
            gb = -a;
        }
        if (!last_u) {
            float a = a0v + BE[ib - Wd - 1];            // beta[t,u+1]: previous diagonal, previous primed column
            a = expf(a + WLb[ib] - b00);
            // (1. + lambda) * a is a double multiply in the reference (core.cu:327-329); with lambda == 0
            // it returns a unchanged, so the fp64 round trip is skipped without changing a bit
            if (has_lam) a = (float)((1.0 + (double)A.lam) * (double)a);
            gl = -a;
        }
        if (A.scale) { gb *= sc; gl *= sc; }
        return make_float2(gb, gl);
    };
    if (MODE == 0) {
        if (live) {
            const int r0 = t0, r1 = min(t1, Tn);
            const int cells = (r1 - r0) * Un;
            const float inv = 1.0f / (float)Un;
            // four cells per thread and pass: all shared-memory reads and both expf chains of the four
            // cells first (independent, so they overlap), then the eight scattered stores
            constexpr int G = 4;
            for (int cb = tid; cb < cells; cb += kFusedThreads * G) {
                float2 gq[G];
                int tq[G], uq[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int c = cb + g * kFusedThreads;
                    int tt = (int)(((float)c + 0.5f) * inv);
                    int u = c - tt * Un;
                    if (u < 0) { --tt; u += Un; } else if (u >= Un) { ++tt; u -= Un; }
                    const bool in = c < cells;
                    tq[g] = in ? r0 + tt : -1;
                    uq[g] = in ? u : 0;
                    gq[g] = in ? cell_grad(r0 + tt, u) : make_float2(0.0f, 0.0f);
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int t = tq[g], u = uq[g];
                    if (t < 0) continue;
                    float *row = A.grads + (slab + (int64_t)t * U + u) * V;
                    if (!(t == T1 && u < U1)) row[A.blank] = gq[g].x;
                    if (u < U1) row[s_lab[u]] = gq[g].y;   // after the blank: a label equal to blank wins (core.cu:383-390)
                }
            }
        }
    } else if (A.pair_grads) {
        const int cells = (t1 - t0) * U;
#pragma unroll 4
        for (int c = tid; c < cells; c += kFusedThreads) {
            const int tt = c / U, u = c - tt * U;
            const int t = t0 + tt;
            float2 gq = make_float2(0.0f, 0.0f);
            if (live && t < Tn && u < Un) gq = cell_grad(t, u);
            A.pair_grads[slab + (int64_t)t * U + u] = gq;
        }
    }
}

// ---- host side -------------------------------------------------------------------------------
static size_t fused_smem_bytes(int T, int U, int Wd, int nd) {
    return sizeof(float) * ((size_t)6 * nd * Wd + (size_t)2 * T) + sizeof(int) * (size_t)U + 64;
}

// Can the fused kernel take this shape?  Fills `plan` with the derived launch parameters.
bool fused_plan(int N, int T, int U, FusedPlan *plan) {
    if (N < 1 || T < 1 || U < 1 || U > 256) return false;
    const int C = U <= 32 ? 1 : (U <= 64 ? 2 : (U <= 128 ? 4 : 8));
    int Wd = (U + C - 1) / C * C;
    if (Wd & 1) ++Wd;                                   // C == 1: keep the diagonal stride even
    const int nd = T + Wd + 16;                         // diagonals + prefetch overshoot
    const size_t smem = fused_smem_bytes(T, U, Wd, nd);
    if (smem > 220 * 1024) return false;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // CTAs per lattice: fill every SM (two CTAs per SM when two staged lattices fit in its shared memory)
    const int per_sm = (smem <= 110 * 1024) ? 2 : 1;
    int slices = (per_sm * sms) / N;
    slices = max(1, min(slices, T));
    plan->W = Wd; plan->ring = nd; plan->nw = C; plan->slices = slices; plan->smem = smem;
    return true;
}

template <int KIND, int MODE, int C>
static cudaError_t launch_fused_kmc(cudaStream_t s, const FusedArgs &a, size_t smem) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_fused<KIND, MODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(a.slices, a.N);
    k_fused<KIND, MODE, C><<<grid, kFusedThreads, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

template <int KIND, int MODE>
static cudaError_t launch_fused_km(cudaStream_t s, const FusedArgs &a, size_t smem, int C) {
    switch (C) {
        case 1: return launch_fused_kmc<KIND, MODE, 1>(s, a, smem);
        case 2: return launch_fused_kmc<KIND, MODE, 2>(s, a, smem);
        case 4: return launch_fused_kmc<KIND, MODE, 4>(s, a, smem);
        default: return launch_fused_kmc<KIND, MODE, 8>(s, a, smem);
    }
}

cudaError_t launch_fused(cudaStream_t s, int kind, const FusedPlan &plan, const float *lp, const int *labels,
                         const int *xn, const int *yn, float *costs, float *grads, float2 *pair_grads,
                         const float *scale, int N, int T, int U, int V, int blank, float lam, int pairs_in,
                         int guard) {
    FusedArgs a;
    a.lp = lp; a.labels = labels; a.xn = xn; a.yn = yn; a.costs = costs; a.grads = grads; a.pair_grads = pair_grads;
    a.scale = scale; a.N = N; a.T = T; a.U = U; a.V = V; a.blank = blank; a.lam = lam; a.pairs_in = pairs_in;
    a.guard = guard; a.slices = (grads || pair_grads) ? plan.slices : 1; a.Wd = plan.W; a.nd = plan.ring;
    {
        static int gw_env = -1;                         // tuning knob: RNNT_B200_GATHER_WARPS in [2,16]
        if (gw_env < 0) {
            const char *e = getenv("RNNT_B200_GATHER_WARPS");
            gw_env = e ? atoi(e) : 0;
            if (gw_env < 2 || gw_env > kFusedThreads / 32) gw_env = 0;
        }
        // fast LSE: the fill (HBM-write bound) outlasts the wavefront, so start it during the gather;
        // exact LSE: the wavefront outlasts the fill, so let every warp gather and start it sooner
        a.gw = gw_env ? gw_env : (kind == kFast ? kGatherWarps : kFusedThreads / 32);
    }
    const bool dense = grads != nullptr;
    const int C = plan.nw;
    if (kind == kFast)
        return dense ? launch_fused_km<kFast, 0>(s, a, plan.smem, C) : launch_fused_km<kFast, 1>(s, a, plan.smem, C);
    // dense-layout exact flavour (the compact layout never takes the fused path)
    return dense ? launch_fused_km<kExactDense, 0>(s, a, plan.smem, C) : launch_fused_km<kExactDense, 1>(s, a, plan.smem, C);
}

}  // namespace rnnt
