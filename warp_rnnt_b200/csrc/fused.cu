// fused.cu -- single-kernel loss + gradient for lattices that fit one SM's shared memory
// (T*U <~ 12k cells: BASELINE configs 1-3).  One launch does everything the reference spreads over
// zeros_like + 4 kernels (+ python gather / mul_):
//
//   phase 0  gather      all warps stage the lattice's blank/label log-probs into shared memory
//                        (replaces the python-level gather, __init__.py:118-128, and the strided
//                        loads inside kernel_warp, core.cu:115-120)
//   phase 1  wavefront   warps [0,nw) run alpha, [nw,2nw) run beta, entirely on shared memory: one
//                        __shfl_up per anti-diagonal inside a warp, tagged shared-memory slots between
//                        warps (replaces kernel_warp + the global-memory counts scheduler,
//                        core.cu:41-258)
//            zero-fill   meanwhile the remaining warps stream zeros over this CTA's slice of the
//                        dense gradient with 128-bit stores (replaces at::zeros_like, binding.cpp:58);
//                        wavefront warps join through a shared work counter when they finish
//   phase 2  cost/guard  kernel_fill_costs (core.cu:334-370)
//            patch       the <= 2 non-zeros per row are written into the freshly zeroed (L2-resident)
//                        lines (kernel_grads_blank/label, core.cu:260-332), or -- MODE 1 -- the
//                        gradients are emitted in (N,T,U,2) form for the deferred dense backward.
//
// grid (S, N): S CTAs per lattice recompute the (cheap) wavefront redundantly and split the
// (bandwidth-bound) fill/patch of the lattice's rows, so small batches still use every SM.
#include "common.cuh"
#include "kernels.cuh"

namespace rnnt {

constexpr int kFusedThreads = 512;
constexpr int kPad = 40;            // sentinel rows above and below the staged log-probs (32 + unroll/prefetch overshoot)
constexpr float kBigF = -1.0e30f;   // finite stand-in for -inf (see wavefront.cu)
constexpr int kFusedPrefetch = 4;

struct __align__(8) FSlot { float val; int row; };

__device__ __forceinline__ float lds_pred(uint32_t addr, bool pred, float dflt) {
    float v;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.f32 %0, %3;\n\t@p ld.shared.f32 %0, [%1];\n\t}"
                 : "=f"(v)
                 : "r"(addr), "r"((int)pred), "f"(dflt)
                 : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ FSlot ld_fslot(uint32_t addr) {
    FSlot s;
    asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=f"(s.val), "=r"(s.row) : "r"(addr) : "memory");
    return s;
}
__device__ __forceinline__ void st_fslot(uint32_t addr, float v, int row) {
    asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "f"(v), "r"(row) : "memory");
}

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
    int W;                  // shared-memory row stride (floats), even, >= U+1
    int ring;               // boundary ring slots per warp boundary (power of two >= T)
    int nw;                 // warps per direction = ceil(U/32)
};

enum { kSrcNone = 0, kSrcRing = 1, kSrcCol0 = 2 };

// One warp's sweep over its 32 columns, on shared memory.  Same recurrence and operand encoding
// as sweep_warp in wavefront.cu; staged arrays carry the sentinels (rows outside [0,Tn): blank 0,
// label kBig; label column U-1 and column -1: kBig) so the loop has no edge cases.
// pb/pl index of cell (r,c): (r + kPad) * W + c + 1.   out index: r * W + c.
// The step loop runs in three phases: ramp-up (s < 32) and ramp-down (s >= Tn) use per-step
// store / publish predicates; the steady phase in between (every lane inside the lattice) has none:
// per step it is one shuffle, the LSE chain, one STS, two LDS and three pointer bumps.
template <int KIND, bool BETA, int SRC>
__device__ __forceinline__ void sweep_smem(uint32_t pb, uint32_t pl, uint32_t out, int W, int Tn, int Un, int wcol,
                                           int lane, uint32_t ring_in, uint32_t ring_out, bool publish) {
    const int j = wcol + lane;
    const bool col_ok = j < Un;
    const int jc = col_ok ? j : Un - 1;                      // lanes beyond the lattice shadow the last column
    const unsigned rows = col_ok ? (unsigned)Tn : 0u;
    const int T1 = Tn - 1, U1 = Un - 1;
    const int nsteps = Tn + min(32, Un - wcol) - 1;
    float val = (j == 0) ? 0.0f : kBigF;
    // running byte addresses of the operands of step s; each step advances by +-W floats
    const int dW = (BETA ? -W : W) * 4;
    uint32_t a_wb, a_wl, a_out;
    if (BETA) {
        const int own = (T1 + lane + kPad) * W + (U1 - jc) + 1;
        a_wb = pb + 4u * own;
        a_wl = pl + 4u * own;
        a_out = out + 4u * ((T1 + lane) * W + (U1 - jc));
    } else {
        a_wb = pb + 4u * ((-lane - 1 + kPad) * W + jc + 1);  // blank[i-1, j]
        a_wl = pl + 4u * ((-lane + kPad) * W + jc);          // label[i, j-1]
        a_out = out + 4u * (-lane * W + jc);
    }
    uint32_t a_c0 = a_out;                                   // exact mode: pre-scanned column 0 (lane 0)
    uint32_t a_put = ring_out - 8u * 31u;                    // slot of row s-31 (ring >= T: no wrap)

    float wb[kFusedPrefetch], wl[kFusedPrefetch], c0v[kFusedPrefetch];
    int fs = 0;                                              // step index of the next fetch (exact mode only)
    auto fetch = [&](int k) {                                // operands kFusedPrefetch steps ahead
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wb[k]) : "r"(a_wb) : "memory");
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wl[k]) : "r"(a_wl) : "memory");
        a_wb += dW;
        a_wl += dW;
        if (SRC == kSrcCol0) {
            c0v[k] = lds_pred(a_c0, lane == 0 && fs < Tn, 0.0f);   // column 0 has Tn rows only
            a_c0 += dW;
            ++fs;
        }
    };
    // consumer side of the inter-warp hand-off: once per group of kFusedPrefetch steps wait for the
    // LAST row of the group (the producer's lane 31 publishes rows in order), then read the values
    // with plain loads -- no tag check on the per-step dependent chain.
    float bv[kFusedPrefetch];
    auto ring_group = [&](int s0) {
        if (s0 < Tn) {                                       // warp-uniform
            const int need = min(s0 + kFusedPrefetch - 1, Tn - 1);
            while (ld_fslot(ring_in + 8u * (uint32_t)need).row != need) {}
#pragma unroll
            for (int k = 0; k < kFusedPrefetch; ++k) {
                const int r = min(s0 + k, Tn - 1);
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv[k]) : "r"(ring_in + 8u * (uint32_t)r) : "memory");
            }
        }
    };
    auto step = [&](int k, int s, bool steady) {
        float left = __shfl_up_sync(0xffffffffu, val, 1);
        if (SRC == kSrcRing) {
            if (lane == 0) left = bv[k];
        }
        const float skip = val + wb[k];
        const float emit = left + wl[k];
        float v = lse<KIND>(skip, emit);
        if (SRC == kSrcCol0) {
            if (lane == 0 && s >= 1) v = c0v[k];              // exact mode: column 0 from the scan pre-pass
        }
        val = v;
        if (steady ? col_ok : ((unsigned)(s - lane) < rows)) sts_f32(a_out, v);
        a_out += dW;
        if (publish && (steady || (unsigned)(s - 31) < (unsigned)Tn)) {
            if (lane == 31) st_fslot(a_put, v, s - 31);
        }
        a_put += 8u;
        fetch(k);
    };
#pragma unroll
    for (int k = 0; k < kFusedPrefetch; ++k) fetch(k);

    int s = 0;
    const int ramp = min(32, nsteps);                          // multiples of kFusedPrefetch below
    for (; s < ramp; s += kFusedPrefetch) {
        if (SRC == kSrcRing) ring_group(s);
#pragma unroll
        for (int k = 0; k < kFusedPrefetch; ++k) step(k, s + k, false);
    }
    for (; s + kFusedPrefetch <= Tn; s += kFusedPrefetch) {
        if (SRC == kSrcRing) ring_group(s);
#pragma unroll
        for (int k = 0; k < kFusedPrefetch; ++k) step(k, s + k, true);
    }
    for (; s < nsteps; s += kFusedPrefetch) {
        if (SRC == kSrcRing) ring_group(s);
#pragma unroll
        for (int k = 0; k < kFusedPrefetch; ++k) step(k, s + k, false);
    }
}

// Column 0 in the reference's summation order (32-wide Kogge-Stone scan per tile + tile base,
// core.cu:92-110 / :197-215) so that exact mode is bit-identical.  One warp.
template <bool BETA>
__device__ __forceinline__ void col0_scan_smem(const float *pb, float *out, int W, int Tn, int Un, int lane) {
    const int T1 = Tn - 1, U1 = Un - 1;
    auto PB = [&](int r, int c) { return pb[(r + kPad) * W + c + 1]; };
    float base = BETA ? PB(T1, U1) : 0.0f;
    if (lane == 0) out[BETA ? (T1 * W + U1) : 0] = base;
    for (int p0 = 0; p0 < T1; p0 += 32) {
        const int i = p0 + lane + 1;
        float b = 0.0f;
        if (i <= T1) b = BETA ? PB(T1 - i, U1) : PB(i - 1, 0);
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const float a = __shfl_up_sync(0xffffffffu, b, k);
            if (k <= lane) b += a;
        }
        const float v = base + b;
        if (i <= T1) out[BETA ? ((T1 - i) * W + U1) : (i * W)] = v;
        base = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
}

// MODE 0: dense gradients (zero-fill + patch).  MODE 1: (N,T,U,2) gradients.  Both write costs.
template <int KIND, int MODE>
__global__ void __launch_bounds__(kFusedThreads, 1) k_fused(FusedArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y, slice = blockIdx.x;
    const int T = A.T, U = A.U, V = A.V, W = A.W;
    const int Tn = A.xn[n], Un = A.yn[n] + 1;
    const bool ok = (Tn >= 1 && Tn <= T && Un >= 1 && Un <= U);
    const int T1 = Tn - 1, U1 = Un - 1;

    // ---- shared memory carve-up
    const int prow = T + 2 * kPad;
    float *pb = reinterpret_cast<float *>(smem_raw);            // [prow][W] blank log-probs (+sentinels)
    float *pl = pb + (size_t)prow * W;                          // [prow][W] label log-probs (+sentinels)
    float *al = pl + (size_t)prow * W;                          // [T][W] alpha
    float *be = al + (size_t)T * W;                             // [T][W] beta
    FSlot *ring = reinterpret_cast<FSlot *>(be + (size_t)T * W);   // [2*nw][ring]
    int *s_lab = reinterpret_cast<int *>(ring + (size_t)2 * A.nw * A.ring);   // [U]
    __shared__ int s_next;                                      // fill work counter
    __shared__ int s_bad;

    // rows of the padded slab this CTA fills / patches
    const int t0 = (int)((int64_t)T * slice / A.slices), t1 = (int)((int64_t)T * (slice + 1) / A.slices);
    const int64_t slab = (int64_t)n * T * U;                    // first cell of this lattice

    // ---- phase 0: sentinels, labels, gather
    if (tid == 0) { s_next = 0; s_bad = ok ? 0 : 1; }
    if (ok) {
        for (int k = tid; k < prow * W; k += kFusedThreads) { pb[k] = 0.0f; pl[k] = kBigF; }
        for (int k = tid; k < 2 * A.nw * A.ring; k += kFusedThreads) ring[k].row = -1;
        if (!A.pairs_in)
            for (int u = tid; u < U1; u += kFusedThreads) s_lab[u] = A.labels[(int64_t)n * (U - 1) + u];
    }
    __syncthreads();
    // Warps [0,GW) gather (then sweep / help filling); warps [GW,16), if any, would start the
    // zero-fill right away.  Measured on B200 (cfg 2): the gather is HBM-bound (it drags in every
    // 64-byte DRAM atom of log_probs) and sits on the critical path, so fill traffic competing with
    // it only delays the wavefront -- GW = all warps is faster (60 us vs 69 us with GW = 8).
    const int nw = A.nw;
    const int GW = kFusedThreads / 32;
    const int gthreads = GW * 32;
    if (ok && warp < GW) {
        const int cells = Tn * Un;
        const float inv = 1.0f / (float)Un;
        const uint64_t pol_first = policy_evict_first();
        constexpr int G = 4;                                    // cells per thread per pass: all loads first
        for (int cb = tid; cb < cells; cb += gthreads * G) {
            float vb[G], vl[G];
            int idx[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c = cb + g * gthreads;
                const bool in = c < cells;
                int t = (int)(((float)c + 0.5f) * inv);
                int u = c - t * Un;
                if (u < 0) { --t; u += Un; } else if (u >= Un) { ++t; u -= Un; }
                if (!in) { t = 0; u = 0; }
                const int64_t cell = slab + (int64_t)t * U + u;
                idx[g] = in ? (t + kPad) * W + u + 1 : -1;
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
                if (idx[g] >= 0) {
                    pb[idx[g]] = vb[g];
                    pl[idx[g]] = vl[g];                          // label column U-1 stays kBig (beta's first column)
                }
            }
        }
    }
    if (warp < GW) asm volatile("bar.sync 1, %0;" ::"r"(gthreads) : "memory");   // gather complete (gather warps only)

    // ---- phase 1: wavefront warps + zero-fill warps
    const uint32_t pb_a = (uint32_t)__cvta_generic_to_shared(pb), pl_a = (uint32_t)__cvta_generic_to_shared(pl);
    if (ok && warp < 2 * nw) {
        const bool beta = warp >= nw;
        const int w = beta ? warp - nw : warp;
        const int wcol = 32 * w;
        if (wcol < Un) {
            float *outp = beta ? be : al;
            const uint32_t out_a = (uint32_t)__cvta_generic_to_shared(outp);
            const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring) + 8u * (uint32_t)((beta ? nw : 0) * A.ring);
            const uint32_t ring_in = ring_base + 8u * (uint32_t)((w > 0 ? w - 1 : 0) * A.ring);
            const uint32_t ring_out = ring_base + 8u * (uint32_t)(w * A.ring);
            const bool publish = (wcol + 32 < Un);
            if (w > 0) {
                beta ? sweep_smem<KIND, true, kSrcRing>(pb_a, pl_a, out_a, W, Tn, Un, wcol, lane, ring_in, ring_out, publish)
                           : sweep_smem<KIND, false, kSrcRing>(pb_a, pl_a, out_a, W, Tn, Un, wcol, lane, ring_in, ring_out, publish);
            } else if (KIND != kFast) {
                if (beta) col0_scan_smem<true>(pb, outp, W, Tn, Un, lane);
                else col0_scan_smem<false>(pb, outp, W, Tn, Un, lane);
                beta ? sweep_smem<KIND, true, kSrcCol0>(pb_a, pl_a, out_a, W, Tn, Un, wcol, lane, ring_in, ring_out, publish)
                           : sweep_smem<KIND, false, kSrcCol0>(pb_a, pl_a, out_a, W, Tn, Un, wcol, lane, ring_in, ring_out, publish);
            } else {
                beta ? sweep_smem<KIND, true, kSrcNone>(pb_a, pl_a, out_a, W, Tn, Un, wcol, lane, ring_in, ring_out, publish)
                           : sweep_smem<KIND, false, kSrcNone>(pb_a, pl_a, out_a, W, Tn, Un, wcol, lane, ring_in, ring_out, publish);
            }
        }
    }
    if (MODE == 0) {
        // zero-fill rows [t0,t1) of this lattice's slab: floats [f0,f1); 32 KB chunks handed out by
        // s_next; 256-bit evict_last stores on the 32-byte aligned interior
        const int64_t f0 = (slab + (int64_t)t0 * U) * V, f1 = (slab + (int64_t)t1 * U) * V;
        float *g = A.grads;
        const bool vec = ((reinterpret_cast<uintptr_t>(g) & 31u) == 0);
        const int64_t a0 = vec ? min(f1, (f0 + 7) & ~(int64_t)7) : f1;
        const int64_t a1 = vec ? max(a0, f1 & ~(int64_t)7) : f1;
        constexpr int kChunk = 8192;                            // floats per chunk
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
        if (warp == kFusedThreads / 32 - 1) {                   // unaligned head / tail (<= 7 floats each, or all if !vec)
            for (int64_t f = f0 + lane; f < a0; f += 32) g[f] = 0.0f;
            for (int64_t f = a1 + lane; f < f1; f += 32) g[f] = 0.0f;
        }
    }
    __syncthreads();

    // ---- phase 2: cost (+ mismatch guard, core.cu:346-367), gradients
    if (tid == 0) {
        float cost = NAN;
        if (ok) {
            float b = be[0];                                         // beta[0,0]
            if (A.guard) {
                const float a = al[T1 * W + U1] + pb[(T1 + kPad) * W + U1 + 1];   // alpha-side ll (core.cu:346)
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
    const float b00 = live ? be[0] : 0.0f;
    const float sc = A.scale ? A.scale[n] : 1.0f;
    auto cell_grad = [&](int t, int u) -> float2 {
        // same operation order as core.cu:284-294 and :319-331
        const float a0v = al[t * W + u];
        float gb = 0.0f, gl = 0.0f;
        const bool last_t = (t == T1), last_u = (u == U1);
        if (!(last_t && !last_u)) {
            float a = a0v;
            if (!last_t) a += be[(t + 1) * W + u];
            a = expf(a + pb[(t + kPad) * W + u + 1] - b00);
            gb = -a;
        }
        if (!last_u) {
            float a = a0v + be[t * W + u + 1];
            a = expf(a + pl[(t + kPad) * W + u + 1] - b00);
            a = (float)((1.0 + (double)A.lam) * (double)a);
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
#pragma unroll 4
            for (int c = tid; c < cells; c += kFusedThreads) {
                int tt = (int)(((float)c + 0.5f) * inv);
                int u = c - tt * Un;
                if (u < 0) { --tt; u += Un; } else if (u >= Un) { ++tt; u -= Un; }
                const int t = r0 + tt;
                const float2 gq = cell_grad(t, u);
                float *row = A.grads + (slab + (int64_t)t * U + u) * V;
                if (!(t == T1 && u < U1)) row[A.blank] = gq.x;
                if (u < U1) row[s_lab[u]] = gq.y;               // after the blank: a label equal to blank wins (core.cu:383-390)
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
static size_t fused_smem_bytes(int T, int W, int nw, int ring, int U) {
    return sizeof(float) * ((size_t)2 * (T + 2 * kPad) * W + (size_t)2 * T * W) + sizeof(FSlot) * (size_t)2 * nw * ring +
           sizeof(int) * (size_t)U + 64;
}

// Can the fused kernel take this shape?  Fills `a` with the derived launch parameters.
bool fused_plan(int N, int T, int U, FusedPlan *plan) {
    if (N < 1 || T < 1 || U < 1 || U > 256) return false;
    const int nw = (U + 31) / 32;
    int W = U + 1;
    if (W & 1) ++W;
    int ring = 32;
    while (ring < T) ring *= 2;
    if (nw == 1) ring = 1;
    const size_t smem = fused_smem_bytes(T, W, nw, ring, U);
    if (smem > 200 * 1024) return false;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int slices = sms / N;
    slices = max(1, min(slices, T));
    plan->W = W; plan->ring = ring; plan->nw = nw; plan->slices = slices; plan->smem = smem;
    return true;
}

template <int KIND, int MODE>
static cudaError_t launch_fused_km(cudaStream_t s, const FusedArgs &a, size_t smem) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_fused<KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(a.slices, a.N);
    k_fused<KIND, MODE><<<grid, kFusedThreads, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_fused(cudaStream_t s, int kind, const FusedPlan &plan, const float *lp, const int *labels,
                         const int *xn, const int *yn, float *costs, float *grads, float2 *pair_grads,
                         const float *scale, int N, int T, int U, int V, int blank, float lam, int pairs_in,
                         int guard) {
    FusedArgs a;
    a.lp = lp; a.labels = labels; a.xn = xn; a.yn = yn; a.costs = costs; a.grads = grads; a.pair_grads = pair_grads;
    a.scale = scale; a.N = N; a.T = T; a.U = U; a.V = V; a.blank = blank; a.lam = lam; a.pairs_in = pairs_in;
    a.guard = guard; a.slices = (grads || pair_grads) ? plan.slices : 1; a.W = plan.W; a.ring = plan.ring; a.nw = plan.nw;
    const bool dense = grads != nullptr;
    switch (kind) {
        case kExactDense:
            return dense ? launch_fused_km<kExactDense, 0>(s, a, plan.smem) : launch_fused_km<kExactDense, 1>(s, a, plan.smem);
        case kExactCompact:
            return dense ? launch_fused_km<kExactCompact, 0>(s, a, plan.smem) : launch_fused_km<kExactCompact, 1>(s, a, plan.smem);
        default:
            return dense ? launch_fused_km<kFast, 0>(s, a, plan.smem) : launch_fused_km<kFast, 1>(s, a, plan.smem);
    }
}

}  // namespace rnnt
