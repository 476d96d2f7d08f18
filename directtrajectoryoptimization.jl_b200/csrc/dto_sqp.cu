// dto_sqp.cu -- per-problem bookkeeping kernels of the native lock-step Newton-KKT solver (dto_sqp_solve).
//
// Reference anchor: solve!(solver) (/root/reference/src/solver.jl:45-47) hands the MOI callbacks to Ipopt; Ipopt is
// absent here and a CPU solver per problem would put PCIe back between the callbacks and their consumer, so the
// caller of the callback path is this batched solver. The algorithm is stated once, in
// directtrajectoryoptimization.jl_b200/sqp.py (`solve`); this file and dto_sqp_host.inc are its native form: the same
// statements in the same order, one warp per problem for everything that reduces over a problem's row, so that the
// python twin driven by the CPU oracle (tests/sqp_oracle.py) checks it iterate by iterate.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "dto_sqp_dev.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// max |x| that propagates NaN like a comparison-free maximum would not: a row with a NaN is "not finite" separately
__device__ __forceinline__ double absmax_row(const double* __restrict__ r, int n, int lane)
{
    double m = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double v = fabs(r[i]);
        m = (v > m || v != v) ? v : m;   // keep a NaN once seen (torch's amax propagates NaN)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(FULL, m, o);
        m = (w > m || w != w) ? w : m;
    }
    return m;
}
__device__ __forceinline__ bool finite_row(const double* __restrict__ r, int n, int lane)
{
    bool ok = true;
    for (int i = lane; i < n; i += 32) ok = ok && isfinite(r[i]);
    return __all_sync(FULL, ok);
}

#define WARP_PROBLEM()                                                        \
    const int lane = threadIdx.x & 31;                                        \
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); \
    if (b >= a.B) return;

// ---- delta = lm; the callbacks see z and lam * exact (Hessian multipliers), the factor kernel sees delta
__global__ void k_begin(const dto_sqp_args a)
{
    WARP_PROBLEM();
    const double ex = a.exact[b];
    for (int i = lane; i < a.N_z; i += 32) a.bz[b * a.N_z + i] = a.z[b * a.N_z + i];
    for (int i = lane; i < a.N_c; i += 32) a.blam[b * a.N_c + i] = a.lam[b * a.N_c + i] * ex;
    if (lane == 0) {
        const double d = a.lm[b];
        a.delta[b] = d;
        a.preg[b] = d;
    }
}

__global__ void k_set_lam(const dto_sqp_args a)
{
    WARP_PROBLEM();
    for (int i = lane; i < a.N_c; i += 32) a.blam[b * a.N_c + i] = a.lam[b * a.N_c + i];
}

// ---- after the callbacks and the first factorisation of the iteration
__global__ void k_after_first(const dto_sqp_args a)
{
    WARP_PROBLEM();
    const double* c = a.bc + b * a.N_c;
    const double* rz = a.rhs + b * a.dim;
    double dr = 0.0;
    for (int i = lane; i < a.N_z; i += 32) {
        const double v = fabs(rz[i] * a.free[i]);
        dr = (v > dr || v != v) ? v : dr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(FULL, dr, o);
        dr = (w > dr || w != w) ? w : dr;
    }
    const double cv = absmax_row(c, a.N_c, lane);
    for (int i = lane; i < a.N_c; i += 32) a.ckeep[b * a.N_c + i] = c[i];
    const bool fin = finite_row(a.sol + b * a.dim, a.dim, lane);
    if (lane == 0) {
        a.cv[b] = cv;
        a.dr[b] = dr;
        a.fcur[b] = a.bf[b];
        a.exact[b] = cv <= a.p.exact_below ? 1.0 : 0.0;
        bool done = a.done[b] != 0;
        const bool newly = !done && cv <= a.p.tol_constraint && dr <= a.p.tol_dual;
        if (newly) a.iters[b] = a.it;
        done = done || newly;
        a.done[b] = done ? 1 : 0;
        const bool bad = (a.nneg[b] != a.N_c || !fin) && !done;
        a.bad[b] = bad ? 1 : 0;
        a.first[b] = 1;
        if (done) atomicAdd(a.counters + DTO_SQP_N_DONE, 1);
        if (bad) atomicAdd(a.counters + DTO_SQP_N_BAD, 1);
    }
}

// ---- inertia control (Ipopt's IC rule per problem): next regularisation of the problems whose pivot count was wrong
__global__ void k_reg_next(const dto_sqp_args a)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B || !a.bad[b]) return;
    const double dl = a.delta_last[b], d = a.delta[b];
    const double start = dl == 0.0 ? a.p.reg_first : fmax(a.p.reg_min, a.p.reg_dec * dl);
    const double grow = dl == 0.0 ? a.p.reg_inc_first : a.p.reg_inc;
    const double nxt = a.first[b] ? fmax(start, 2.0 * d) : fmin(a.p.reg_max, grow * d);
    a.delta[b] = nxt;
    a.preg[b] = nxt;
    a.first[b] = 0;
    a.idx[atomicAdd(a.counters + DTO_SQP_N_IDX, 1)] = (int32_t)b;
}

__global__ void k_recheck(const dto_sqp_args a, int32_t count)
{
    const int lane = threadIdx.x & 31;
    const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= count) return;
    const int64_t b = a.idx[k];
    const bool fin = finite_row(a.sol + b * a.dim, a.dim, lane);
    if (lane == 0) {
        const bool bad = a.nneg[b] != a.N_c || !fin;   // (these problems are not done)
        a.bad[b] = bad ? 1 : 0;
        if (bad) atomicAdd(a.counters + DTO_SQP_N_BAD, 1);
    }
}

// ---- step, l1-merit quantities (curvature rule for the penalty), first trial point
__global__ void k_direction(const dto_sqp_args a)
{
    WARP_PROBLEM();
    const double* sol = a.sol + b * a.dim;
    const double* g = a.bg + b * a.N_z;
    const double* ck = a.ckeep + b * a.N_c;
    double gd = 0.0, c1 = 0.0, cl = 0.0;
    for (int i = lane; i < a.N_z; i += 32) {
        const double d = -sol[i] * a.free[i];
        a.dz[b * a.N_z + i] = d;
        gd += g[i] * d;
        a.bz[b * a.N_z + i] = a.z[b * a.N_z + i] + 1.0 * d;   // first trial: alpha = 1
    }
    for (int i = lane; i < a.N_c; i += 32) {
        const double dl = -sol[a.N_z + i];
        a.dlam[b * a.N_c + i] = dl;
        c1 += fabs(ck[i]);
        cl += ck[i] * (a.lam[b * a.N_c + i] + dl);
    }
    gd = wsum(gd);
    c1 = wsum(c1);
    cl = wsum(cl);
    if (lane == 0) {
        const double delta = a.delta[b], lm = a.lm[b];
        if (delta > lm) a.delta_last[b] = delta;
        const double curv = fmax(-gd + cl, 0.0);
        const double nu_need = (gd + 0.5 * curv) / ((1.0 - a.p.merit_rho) * fmax(c1, 1.0e-300));
        const double nu = fmax(a.p.merit_min, a.p.merit_margin * nu_need);
        a.nu[b] = nu;
        a.c1[b] = c1;
        a.slope[b] = gd - nu * c1;
        a.phi0[b] = a.fcur[b] + nu * c1;
        a.alpha[b] = 1.0;
        a.accepted[b] = (a.done[b] || a.bad[b]) ? 1 : 0;   // converged problems and failed factorisations do not move
    }
}

// ---- one backtracking round: bz holds the trial point, bf / bc its objective and constraint values
__global__ void k_ls_round(const dto_sqp_args a, int32_t round)
{
    WARP_PROBLEM();
    const double* ct = a.bc + b * a.N_c;
    double ct1 = 0.0;
    for (int i = lane; i < a.N_c; i += 32) ct1 += fabs(ct[i]);
    ct1 = wsum(ct1);
    bool accepted = a.accepted[b] != 0;
    double alpha = a.alpha[b];
    const double nu = a.nu[b];
    const double phit = a.bf[b] + nu * ct1;
    const bool ok = (phit <= a.phi0[b] + a.p.armijo * alpha * a.slope[b]) && !accepted;
    if (ok) {
        for (int i = lane; i < a.N_z; i += 32) a.z[b * a.N_z + i] = a.bz[b * a.N_z + i];
        for (int i = lane; i < a.N_c; i += 32) a.lam[b * a.N_c + i] += alpha * a.dlam[b * a.N_c + i];
    }
    accepted = accepted || ok;
    if (round == 0 && a.p.soc) {
        // second-order correction candidates: rejected full steps that did not even reduce the constraint violation
        const bool need = !accepted && ct1 >= a.c1[b];
        if (need)
            for (int i = lane; i < a.N_c; i += 32) a.bc[b * a.N_c + i] = a.ckeep[b * a.N_c + i] + ct[i];   // c(z) + c(z + dz)
        if (lane == 0) {
            a.need[b] = need ? 1 : 0;
            if (need) a.idx[atomicAdd(a.counters + DTO_SQP_N_NEED, 1)] = (int32_t)b;
        }
    }
    if (!accepted) alpha = 0.5 * alpha;
    __syncwarp();
    for (int i = lane; i < a.N_z; i += 32) a.bz[b * a.N_z + i] = a.z[b * a.N_z + i] + alpha * a.dz[b * a.N_z + i];   // next trial
    if (lane == 0) {
        a.accepted[b] = accepted ? 1 : 0;
        a.alpha[b] = alpha;
        if (!accepted) atomicAdd(a.counters + DTO_SQP_N_OPEN, 1);   // the host stops the search when no problem is open
    }
}

__global__ void k_soc_trial(const dto_sqp_args a, int32_t count)
{
    const int lane = threadIdx.x & 31;
    const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= count) return;
    const int64_t b = a.idx[k];
    const double* sol = a.sol + b * a.dim;
    for (int i = lane; i < a.N_z; i += 32) a.bz[b * a.N_z + i] = a.z[b * a.N_z + i] + (-sol[i] * a.free[i]);
}

__global__ void k_soc_accept(const dto_sqp_args a, int32_t count)
{
    const int lane = threadIdx.x & 31;
    const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= count) return;
    const int64_t b = a.idx[k];
    const double* cs = a.bc + b * a.N_c;
    const double* sol = a.sol + b * a.dim;
    double cs1 = 0.0;
    for (int i = lane; i < a.N_c; i += 32) cs1 += fabs(cs[i]);
    cs1 = wsum(cs1);
    const bool fin = finite_row(sol, a.dim, lane);
    const bool oks = (a.bf[b] + a.nu[b] * cs1 <= a.phi0[b] + a.p.armijo * a.slope[b]) && fin;
    double alpha = a.alpha[b];
    if (oks) {
        for (int i = lane; i < a.N_z; i += 32) a.z[b * a.N_z + i] = a.bz[b * a.N_z + i];
        for (int i = lane; i < a.N_c; i += 32) a.lam[b * a.N_c + i] -= sol[a.N_z + i];
        alpha = 1.0;   // the full (corrected) step was taken: round 0 had already halved alpha for this problem
    }
    __syncwarp();
    for (int i = lane; i < a.N_z; i += 32) a.bz[b * a.N_z + i] = a.z[b * a.N_z + i] + alpha * a.dz[b * a.N_z + i];   // trial of round 1
    if (lane == 0) {
        if (oks) {
            a.accepted[b] = 1;
            atomicAdd(a.counters + DTO_SQP_N_SOC_OK, 1);
        }
        a.alpha[b] = alpha;
    }
}

// ---- end of the iteration: multiplier safeguard, Levenberg-Marquardt damping, regularisation memory
__global__ void k_end(const dto_sqp_args a)
{
    WARP_PROBLEM();
    if (a.p.lam_max > 0.0) {
        const double m = absmax_row(a.lam + b * a.N_c, a.N_c, lane);
        if (m > a.p.lam_max)
            for (int i = lane; i < a.N_c; i += 32) a.lam[b * a.N_c + i] = 0.0;
    }
    if (lane == 0) {
        const bool accepted = a.accepted[b] != 0, done = a.done[b] != 0, bad = a.bad[b] != 0;
        const bool moved = !done && !bad;
        const double alpha = a.alpha[b], delta = a.delta[b];
        double lm = a.lm[b];
        if (moved && alpha < a.p.lm_grow_below) lm = fmax(a.p.lm_min, a.p.lm_grow * fmax(lm, delta));
        if (moved && accepted && alpha >= 1.0) lm = a.p.lm_shrink * lm;
        if (lm < a.p.lm_zero) lm = 0.0;
        a.lm[b] = lm;
        if (!accepted || bad) a.delta_last[b] = fmax(a.p.reg_first, a.p.reg_inc * fmax(a.delta_last[b], delta));
    }
}

template <class K, class... Args>
int launch_warp_per(K kernel, int64_t n, void* stream, Args... args)
{
    if (n <= 0) return 0;
    kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(args...);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace

extern "C" int dto_sqp_k_begin(const dto_sqp_args* a, void* s) { return launch_warp_per(k_begin, a->B, s, *a); }
extern "C" int dto_sqp_k_set_lam(const dto_sqp_args* a, void* s) { return launch_warp_per(k_set_lam, a->B, s, *a); }
extern "C" int dto_sqp_k_after_first(const dto_sqp_args* a, void* s) { return launch_warp_per(k_after_first, a->B, s, *a); }
extern "C" int dto_sqp_k_reg_next(const dto_sqp_args* a, void* s)
{
    if (a->B <= 0) return 0;
    k_reg_next<<<(unsigned)((a->B + 127) / 128), 128, 0, (cudaStream_t)s>>>(*a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}
extern "C" int dto_sqp_k_recheck(const dto_sqp_args* a, int32_t count, void* s) { return launch_warp_per(k_recheck, count, s, *a, count); }
extern "C" int dto_sqp_k_direction(const dto_sqp_args* a, void* s) { return launch_warp_per(k_direction, a->B, s, *a); }
extern "C" int dto_sqp_k_ls_round(const dto_sqp_args* a, int32_t round, void* s) { return launch_warp_per(k_ls_round, a->B, s, *a, round); }
extern "C" int dto_sqp_k_soc_trial(const dto_sqp_args* a, int32_t count, void* s) { return launch_warp_per(k_soc_trial, count, s, *a, count); }
extern "C" int dto_sqp_k_soc_accept(const dto_sqp_args* a, int32_t count, void* s) { return launch_warp_per(k_soc_accept, count, s, *a, count); }
extern "C" int dto_sqp_k_end(const dto_sqp_args* a, void* s) { return launch_warp_per(k_end, a->B, s, *a); }
