// Hand-written sm_100a kernels for the batched NLP callbacks. Included by every generated
// model translation unit AFTER the generated `struct DtoModel` (element device functions +
// kind dispatchers), so that the element code inlines into the kernels.
//
// Work decomposition (replaces the reference's serial `for t` loops,
// /root/reference/src/dynamics.jl:103-127, src/costs.jl:49-73, src/constraints.jl:80-104):
//   * one LANE per (problem, knot) item, items numbered flat g = b*T + t so warps stay full
//     even when T is not a multiple of 32 (a warp may straddle two problems);
//   * one WARP is an independent tile of 32 consecutive items (31 + 1 halo item when a
//     dynamics Hessian reaches next-state rows): no __syncthreads anywhere;
//   * every lane writes its element values into the warp's shared-memory segments at the
//     item's flat offset, then the warp streams each segment to HBM with coalesced stores:
//     problem-major outputs J[b][:], H[b][:] are contiguous per (segment, problem) range;
//   * Hessian slots are gathered owner-computes: the warp owns the slots whose ROW belongs to
//     its knots and sums the contributing terms in the reference's += order
//     (cost, dynamics t-1 (as y), dynamics t (as x), stage; /root/reference/src/moi.jl:88-111),
//     so results are deterministic and no atomics / zero-fill are needed.
//   * the GeneralConstraint block (/root/reference/src/general_constraint.jl:73-91) runs as a
//     second tiny kernel on the same stream: J slots assigned, H slots += (it is last in the
//     reference's order too).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dto_model_abi.h"

#ifndef DTO_WARPS
#define DTO_WARPS 4
#endif
#ifndef DTO_MIN_CTAS
#define DTO_MIN_CTAS 1
#endif
#ifndef DTO_GATHER_UNROLL
#define DTO_GATHER_UNROLL 8
#endif
#ifndef DTO_L2_PREFETCH
#define DTO_L2_PREFETCH 1
#endif

#define DTO_MODE_G 1
#define DTO_MODE_C 2
#define DTO_MODE_J 4
#define DTO_MODE_H 8

namespace dto {

__device__ __forceinline__ dto_knot_entry load_knot(const dto_knot_entry* __restrict__ tab, int t)
{
    // 64-byte entry = 4 x 16-byte read-only loads
    const int4* p = reinterpret_cast<const int4*>(tab + t);
    int4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    dto_knot_entry e;
    e.zofs = a.x; e.nx = a.y; e.wofs = a.z; e.kdyn = a.w;
    e.kcost = b.x; e.kstage = b.y; e.rdyn = b.z; e.rstage = b.w;
    e.jdyn = c.x; e.jstage = c.y; e.hterm = c.z; e.hslot = c.w;
    e.hclass = d.x; e.hprev = d.y; e.pad0 = d.z; e.pad1 = d.w;
    return e;
}

// flat item index -> (problem, knot) without an integer divide (magic number from the runtime)
__device__ __forceinline__ void split_item(const dto_launch_args& a, int g, int& b, int& t)
{
    b = (int)(((unsigned long long)(unsigned)g * a.div_mul) >> a.div_shift);
    t = g - b * a.T;
}

__device__ __forceinline__ void warp_stream_out(double* __restrict__ dst, const double* __restrict__ src, int n, int lane)
{
    // dst is 8-byte aligned only (slot ranges start anywhere); one 256-byte row per warp instruction,
    // four rows in flight per trip
    int i = lane;
    for (; i + 96 < n; i += 128) {
        const double v0 = src[i], v1 = src[i + 32], v2 = src[i + 64], v3 = src[i + 96];
        dst[i] = v0;
        dst[i + 32] = v1;
        dst[i + 64] = v2;
        dst[i + 96] = v3;
    }
    for (; i < n; i += 32) dst[i] = src[i];
}

template <int MODE>
__host__ __device__ constexpr bool seg_active(int s)
{
    return (s == DTO_SEG_G && (MODE & DTO_MODE_G)) || ((s == DTO_SEG_CDYN || s == DTO_SEG_CSTAGE) && (MODE & DTO_MODE_C)) ||
           ((s == DTO_SEG_JDYN || s == DTO_SEG_JSTAGE) && (MODE & DTO_MODE_J)) || (s == DTO_SEG_HTERM && (MODE & DTO_MODE_H));
}

template <int MODE>
__host__ __device__ inline int smem_doubles_per_warp(const dto_launch_args& a, int* base)
{
    int per_warp = 0;
    for (int s = 0; s < 6; ++s) {
        if (seg_active<MODE>(s)) {
            if (base) base[s] = per_warp + a.seg_pad[s];
            per_warp += a.seg_pad[s] + a.seg_cap[s];
        } else if (base) {
            base[s] = 0;
        }
    }
    return (per_warp + 1) & ~1;  // keep warps 16-byte aligned
}

// ---------------------------------------------------------------------------------------
// per-knot kernel: gradient / residuals / Jacobian / Hessian / fused Jacobian+Hessian
// ---------------------------------------------------------------------------------------
template <class M, int MODE>
__global__ void __launch_bounds__(DTO_WARPS * 32, DTO_MIN_CTAS) knot_kernel(const __grid_constant__ dto_launch_args a)
{
    extern __shared__ __align__(16) double dto_smem[];
    constexpr bool DO_G = (MODE & DTO_MODE_G) != 0, DO_C = (MODE & DTO_MODE_C) != 0;
    constexpr bool DO_J = (MODE & DTO_MODE_J) != 0, DO_H = (MODE & DTO_MODE_H) != 0;
    constexpr bool HALO = DO_H && (M::HESS_HALO != 0);
    constexpr int OWN = HALO ? 31 : 32;

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int T = a.T;
    const int total = (int)(a.B * T);  // the runtime guarantees B*T < 2^31 per shard
    const int g0 = (int)((blockIdx.x * DTO_WARPS + wib) * OWN);  // first OWN item of this warp
    if (g0 >= total) return;
    const int g1 = (g0 + OWN < total) ? g0 + OWN : total;
    int b0, t0;
    split_item(a, g0, b0, t0);

    // L2 prefetch of the inputs of the tile that will start when this one retires (the grid is
    // consumed in order, `tiles_in_flight` warps at a time): its loads then hit L2, not HBM.
    if (a.tiles_in_flight > 0) {
        const long long ga = (long long)g0 + (long long)a.tiles_in_flight * OWN;
        if (ga < total) {
            int ba, ta_;
            split_item(a, (int)ga, ba, ta_);
            const size_t zf = (size_t)ba * a.N_z + (size_t)ta_ * a.z_per_knot;   // approximate start is fine
            const size_t lf = (size_t)ba * a.N_c + (size_t)ta_ * a.c_per_knot;
            const int zl = ((OWN + 1) * a.z_per_knot) / 16 + 2;                   // 128-byte lines of the z range
            const int ll = (OWN * a.c_per_knot) / 16 + 2;
            if (lane < zl) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.z + zf + lane * 16));
            if ((MODE & DTO_MODE_H) && lane < ll) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.lam + lf + lane * 16));
        }
    }

    int base[6];
    const int per_warp = smem_doubles_per_warp<MODE>(a, base);
    double* __restrict__ sm = dto_smem + (size_t)wib * per_warp;

    const dto_knot_entry k0 = load_knot(a.knot, t0);
    const dto_knot_entry kb = load_knot(a.knot, 0);
    const dto_knot_entry kT = load_knot(a.knot, T);
    const int L_g = kT.zofs, L_cd = kT.rdyn - kb.rdyn, L_cs = kT.rstage - kb.rstage;
    const int L_jd = kT.jdyn - kb.jdyn, L_js = kT.jstage - kb.jstage, L_h = kT.hterm;

    // ---- compute phase: one item per lane ----
    {
        const int g = g0 + lane - (HALO ? 1 : 0);
        const bool in = (g < g1) && (g >= g0 || (HALO && t0 > 0));
        if (in) {
            const bool own = g >= g0;
            int b, t;
            split_item(a, g, b, t);
            const int db = b - b0;
            const dto_knot_entry ke = load_knot(a.knot, t);
            const dto_knot_entry kn = load_knot(a.knot, t + 1);
            const double* __restrict__ zb = a.z + (size_t)b * a.N_z;
            const double* __restrict__ x = zb + ke.zofs;
            const double* __restrict__ u = x + ke.nx;
            const double* __restrict__ y = zb + kn.zofs;
            const double* __restrict__ w = a.w + (size_t)b * a.N_w + ke.wofs;
            const double* __restrict__ lam = a.lam + (size_t)b * a.N_c;

            double* hterm = sm + base[DTO_SEG_HTERM] + db * L_h + (ke.hterm - k0.hterm);
            // cost
            if (own) {
                if (DO_G) M::cost_grad(ke.kcost, x, u, w, sm + base[DTO_SEG_G] + db * L_g + (ke.zofs - k0.zofs));
                if (DO_H) M::cost_hess(ke.kcost, x, u, w, __ldg(a.sigma + b), hterm);
            }
            if (DO_H) hterm += M::cost_nh(ke.kcost);
            // dynamics (also evaluated by the halo lane: its y-row Hessian terms feed our first knot)
            if (ke.kdyn >= 0) {
                if (DO_C) M::dyn_res(ke.kdyn, y, x, u, w, sm + base[DTO_SEG_CDYN] + db * L_cd + (ke.rdyn - k0.rdyn));
                double* jd = sm + base[DTO_SEG_JDYN] + db * L_jd + (ke.jdyn - k0.jdyn);
                if (DO_J && DO_H) M::dyn_jac_hess(ke.kdyn, y, x, u, w, lam + ke.rdyn, jd, hterm);
                else if (DO_J) M::dyn_jac(ke.kdyn, y, x, u, w, jd);
                else if (DO_H) M::dyn_hess(ke.kdyn, y, x, u, w, lam + ke.rdyn, hterm);
                if (DO_H) hterm += M::dyn_nh(ke.kdyn);
            }
            // stage constraint
            if (own && ke.kstage >= 0) {
                if (DO_C) M::stage_res(ke.kstage, x, u, w, sm + base[DTO_SEG_CSTAGE] + db * L_cs + (ke.rstage - k0.rstage));
                double* js = sm + base[DTO_SEG_JSTAGE] + db * L_js + (ke.jstage - k0.jstage);
                if (DO_J && DO_H) M::stage_jac_hess(ke.kstage, x, u, w, lam + ke.rstage, js, hterm);
                else if (DO_J) M::stage_jac(ke.kstage, x, u, w, js);
                else if (DO_H) M::stage_hess(ke.kstage, x, u, w, lam + ke.rstage, hterm);
            }
        }
    }
    __syncwarp();

    // ---- compiled Hessian gather (recipes.py): each lane sums the terms of ITS knot's slots with
    // straight-line code (no table loads), then the slot values overwrite the (now dead) term
    // buffer so that the stream-out below is a plain coalesced copy.
    constexpr bool HG = DO_H && (M::HG_NCLASS > 0);
    if (HG && a.use_hclass) {
        double v[M::HG_VMAX > 0 ? M::HG_VMAX : 1];
        const int g = g0 + lane - (HALO ? 1 : 0);
        const bool own = (g < g1) && (g >= g0);
        int cls = -1;
        double* dst = nullptr;
        int b = 0, t = 0;
        if (own) {
            split_item(a, g, b, t);
            const dto_knot_entry ke = load_knot(a.knot, t);
            const double* ownp = sm + base[DTO_SEG_HTERM] + (b - b0) * L_h + (ke.hterm - k0.hterm);
            cls = ke.hclass;
            dst = sm + base[DTO_SEG_HTERM] + (b - b0) * a.nnz_H + (ke.hslot - k0.hslot);
            M::hg_compute(cls, ownp, ownp - ke.hprev, v);
        }
        __syncwarp();
        if (own) M::hg_store(cls, v, dst);
        __syncwarp();
        // general-constraint Hessian entries owned by this lane's knot: last in the reference's +=
        // order (src/moi.jl:112-118), added in shared memory so H is written to HBM exactly once
        if (a.gen_nhess > 0) {
            if (own) {
                const int p0 = __ldg(a.gh_ptr + t), p1 = __ldg(a.gh_ptr + t + 1);
                const int khs = load_knot(a.knot, t).hslot;
                for (int p = p0; p < p1; ++p) {
                    const int2 e = __ldg(reinterpret_cast<const int2*>(a.gh_ent) + p);  // slot, instance
                    const int4 inst = __ldg(reinterpret_cast<const int4*>(a.gen_inst[2]) + e.y);
                    const double val = M::gen_eval(2, inst.x, a.z + (size_t)b * a.N_z + inst.y, a.w + (size_t)b * a.N_w + inst.z,
                                                   a.lam + (size_t)b * a.N_c + a.gen_row0 + inst.w);
                    dst[e.x - khs] += val;
                }
            }
            __syncwarp();
        }
    }

    // ---- stream-out phase: per (problem) sub-tile, coalesced ----
    {
        int b = b0, ta = t0;
        int rem = g1 - g0;
        while (rem > 0) {
            const int cnt = (rem < T - ta) ? rem : (T - ta);
            const int tb = ta + cnt;
            const int db = b - b0;
            const dto_knot_entry ea = load_knot(a.knot, ta);
            const dto_knot_entry eb = load_knot(a.knot, tb);
            if (DO_G)
                warp_stream_out(a.g + (size_t)b * a.N_z + ea.zofs, sm + base[DTO_SEG_G] + db * L_g + (ea.zofs - k0.zofs),
                                eb.zofs - ea.zofs, lane);
            if (DO_C) {
                warp_stream_out(a.c + (size_t)b * a.N_c + ea.rdyn, sm + base[DTO_SEG_CDYN] + db * L_cd + (ea.rdyn - k0.rdyn),
                                eb.rdyn - ea.rdyn, lane);
                warp_stream_out(a.c + (size_t)b * a.N_c + ea.rstage,
                                sm + base[DTO_SEG_CSTAGE] + db * L_cs + (ea.rstage - k0.rstage), eb.rstage - ea.rstage, lane);
            }
            if (DO_J) {
                warp_stream_out(a.J + (size_t)b * a.nnz_J + ea.jdyn, sm + base[DTO_SEG_JDYN] + db * L_jd + (ea.jdyn - k0.jdyn),
                                eb.jdyn - ea.jdyn, lane);
                warp_stream_out(a.J + (size_t)b * a.nnz_J + ea.jstage,
                                sm + base[DTO_SEG_JSTAGE] + db * L_js + (ea.jstage - k0.jstage), eb.jstage - ea.jstage, lane);
            }
            if (HG && a.use_hclass) {
                warp_stream_out(a.H + (size_t)b * a.nnz_H + ea.hslot,
                                sm + base[DTO_SEG_HTERM] + db * a.nnz_H + (ea.hslot - k0.hslot), eb.hslot - ea.hslot, lane);
            } else if (DO_H) {
                const double* __restrict__ smh = sm + base[DTO_SEG_HTERM] + db * L_h - k0.hterm;
                double* __restrict__ Hb = a.H + (size_t)b * a.nnz_H;
                const int4* __restrict__ src4 = reinterpret_cast<const int4*>(a.hsrc4);
                // DTO_GATHER_UNROLL table records per lane are in flight before the first use: the
                // records live in L2 (one 16-byte coalesced load each), so the exposed latency is
                // one L2 round trip per 32*UNROLL slots instead of one per 32 slots.
                int s = ea.hslot + lane;
                for (; s + 32 * (DTO_GATHER_UNROLL - 1) < eb.hslot; s += 32 * DTO_GATHER_UNROLL) {
                    int4 q[DTO_GATHER_UNROLL];
#pragma unroll
                    for (int k = 0; k < DTO_GATHER_UNROLL; ++k) q[k] = __ldg(src4 + s + 32 * k);
#pragma unroll
                    for (int k = 0; k < DTO_GATHER_UNROLL; ++k) {
                        double acc = q[k].x >= 0 ? smh[q[k].x] : 0.0;
                        if (q[k].y >= 0) acc += smh[q[k].y];
                        if (q[k].z >= 0) acc += smh[q[k].z];
                        if (q[k].w >= 0) acc += smh[q[k].w];
                        Hb[s + 32 * k] = acc;
                    }
                }
                for (; s < eb.hslot; s += 32) {
                    const int4 q0 = __ldg(src4 + s);
                    double acc = q0.x >= 0 ? smh[q0.x] : 0.0;
                    if (q0.y >= 0) acc += smh[q0.y];
                    if (q0.z >= 0) acc += smh[q0.z];
                    if (q0.w >= 0) acc += smh[q0.w];
                    Hb[s] = acc;
                }
            }
            rem -= cnt;
            ++b;
            ta = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// objective value: one warp per problem, lanes stride over knots, xor-shuffle tree
// (/root/reference/src/costs.jl:49-56)
// ---------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(DTO_WARPS * 32) objective_kernel(const __grid_constant__ dto_launch_args a)
{
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * DTO_WARPS + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const double* __restrict__ zb = a.z + (size_t)b * a.N_z;
    double acc = 0.0;
    for (int t = lane; t < a.T; t += 32) {
        const dto_knot_entry ke = load_knot(a.knot, t);
        const double* x = zb + ke.zofs;
        acc += M::cost_val(ke.kcost, x, x + ke.nx, a.w + (size_t)b * a.N_w + ke.wofs);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) a.f[b] = acc;
}

// ---------------------------------------------------------------------------------------
// general constraint: one thread per (problem, output instance)
// ---------------------------------------------------------------------------------------
template <class M, int CLS>
__global__ void __launch_bounds__(128) general_kernel(const __grid_constant__ dto_launch_args a)
{
    const int n = (CLS == 0) ? a.gen_nrow : (CLS == 1) ? a.gen_njac : a.gen_nhess;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * (long long)n) return;
    const long long b = idx / n;
    const int i = (int)(idx - b * n);
    const int4 inst = __ldg(reinterpret_cast<const int4*>(a.gen_inst[CLS]) + i);  // tmpl, zbase, wbase, lbase
    const double* __restrict__ z = a.z + (size_t)b * a.N_z + inst.y;
    const double* __restrict__ w = a.w + (size_t)b * a.N_w + inst.z;
    if (CLS == 0) {
        a.c[(size_t)b * a.N_c + a.gen_row0 + i] = M::gen_eval(0, inst.x, z, w, nullptr);
    } else if (CLS == 1) {
        a.J[(size_t)b * a.nnz_J + a.gen_jac0 + i] = M::gen_eval(1, inst.x, z, w, nullptr);
    } else {
        const double* __restrict__ lam = a.lam + (size_t)b * a.N_c + a.gen_row0 + inst.w;
        double* h = a.H + (size_t)b * a.nnz_H + __ldg(a.gen_hslot + i);
        *h += M::gen_eval(2, inst.x, z, w, lam);  // general is last in the reference's += order
    }
}

// ---------------------------------------------------------------------------------------
// host side of the model library
// ---------------------------------------------------------------------------------------
template <int MODE>
inline int64_t knot_smem_bytes(const dto_launch_args& a)
{
    return (int64_t)smem_doubles_per_warp<MODE>(a, nullptr) * DTO_WARPS * (int64_t)sizeof(double);
}

template <class M, int MODE>
inline int launch_knot(const dto_launch_args& a, cudaStream_t st)
{
    constexpr bool HALO = ((MODE & DTO_MODE_H) != 0) && (M::HESS_HALO != 0);
    const long long own = HALO ? 31 : 32;
    const long long total = a.B * (long long)a.T;
    if (total == 0) return 0;
    const long long warps = (total + own - 1) / own;
    const long long ctas = (warps + DTO_WARPS - 1) / DTO_WARPS;
    const int64_t smem = knot_smem_bytes<MODE>(a);
    if (smem > 48 * 1024) {  // opt in to large dynamic shared memory (per device, cheap)
        cudaError_t e = cudaFuncSetAttribute(knot_kernel<M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            fprintf(stderr, "[dto] cudaFuncSetAttribute(mode %d, smem %lld) failed: %s\n", MODE, (long long)smem, cudaGetErrorString(e));
            return (int)e;
        }
    }
    dto_launch_args b = a;
    b.tiles_in_flight = 0;
#if DTO_L2_PREFETCH
    {
        static int resident_warps[16] = {0};  // per device, for the shared-memory size it was computed with
        static int64_t resident_smem[16] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 16) {
            if (resident_warps[dev] == 0 || resident_smem[dev] != smem) {
                resident_smem[dev] = smem;
                int sms = 0, per_sm = 0;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knot_kernel<M, MODE>, DTO_WARPS * 32, (size_t)smem) != cudaSuccess)
                    per_sm = 0;
                resident_warps[dev] = sms * per_sm * DTO_WARPS;
                cudaGetLastError();
            }
            b.tiles_in_flight = resident_warps[dev];
        }
    }
#endif
    knot_kernel<M, MODE><<<(unsigned)ctas, DTO_WARPS * 32, (size_t)smem, st>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        fprintf(stderr, "[dto] knot_kernel<mode %d> launch failed: %s (grid %lld, block %d, smem %lld)\n", MODE, cudaGetErrorString(e),
                ctas, DTO_WARPS * 32, (long long)smem);
    return (int)e;
}

template <class M, int CLS>
inline int launch_general(const dto_launch_args& a, cudaStream_t st)
{
    const int n = (CLS == 0) ? a.gen_nrow : (CLS == 1) ? a.gen_njac : a.gen_nhess;
    const long long total = a.B * (long long)n;
    if (total == 0) return 0;
    general_kernel<M, CLS><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}

template <class M>
inline int launch(int kernel_id, const dto_launch_args* pa, void* stream)
{
    const dto_launch_args& a = *pa;
    cudaStream_t st = (cudaStream_t)stream;
    int e = 0;
    switch (kernel_id) {
    case DTO_K_OBJECTIVE:
        if (a.B == 0) return 0;
        objective_kernel<M><<<(unsigned)((a.B + DTO_WARPS - 1) / DTO_WARPS), DTO_WARPS * 32, 0, st>>>(a);
        return (int)cudaGetLastError();
    case DTO_K_GRADIENT:
        return launch_knot<M, DTO_MODE_G>(a, st);
    case DTO_K_CONSTRAINT:
        if ((e = launch_knot<M, DTO_MODE_C>(a, st))) return e;
        return launch_general<M, 0>(a, st);
    case DTO_K_JACOBIAN:
        if ((e = launch_knot<M, DTO_MODE_J>(a, st))) return e;
        return launch_general<M, 1>(a, st);
    case DTO_K_HESSIAN:
        if ((e = launch_knot<M, DTO_MODE_H>(a, st))) return e;
        if (M::HG_NCLASS > 0 && a.use_hclass) return 0;  // general Hessian fused into the knot kernel
        return launch_general<M, 2>(a, st);
    case DTO_K_JAC_HESS:
        if ((e = launch_knot<M, DTO_MODE_J | DTO_MODE_H>(a, st))) return e;
        if ((e = launch_general<M, 1>(a, st))) return e;
        if (M::HG_NCLASS > 0 && a.use_hclass) return 0;
        return launch_general<M, 2>(a, st);
    default:
        return (int)cudaErrorInvalidValue;
    }
}

inline int64_t smem_bytes(int kernel_id, const dto_launch_args* pa)
{
    switch (kernel_id) {
    case DTO_K_GRADIENT: return knot_smem_bytes<DTO_MODE_G>(*pa);
    case DTO_K_CONSTRAINT: return knot_smem_bytes<DTO_MODE_C>(*pa);
    case DTO_K_JACOBIAN: return knot_smem_bytes<DTO_MODE_J>(*pa);
    case DTO_K_HESSIAN: return knot_smem_bytes<DTO_MODE_H>(*pa);
    case DTO_K_JAC_HESS: return knot_smem_bytes<DTO_MODE_J | DTO_MODE_H>(*pa);
    default: return 0;
    }
}

}  // namespace dto
