// Warp-specialised per-knot kernel (the default for FP64-heavy models). Included by
// dto_kernels.cuh inside namespace dto, after the shared helpers (tile_geom, piece_off, bulk_*).
//
// One CTA of 12 warps per SM for the whole launch:
//   * warps 0-3 ("helpers", 40 registers after setmaxnreg.dec) do every piece of integer work: tile
//     geometry, the bulk input copies (cp.async.bulk, completion on the stage's mbarrier), and a
//     per-item DESCRIPTOR (shared-memory offsets of the item's inputs and output slots, element
//     kinds, gather class) plus the list of output pieces;
//   * warps 4-11 ("compute", 232 registers after setmaxnreg.inc: no spills, room for ILP) wait on the
//     stage's mbarrier, load their descriptor (3 x LDS.128), run the generated FP64 code of
//     /root/reference/src/dynamics.jl:103-127, src/costs.jl:58-73, src/constraints.jl:80-104, gather
//     the Hessian slots (reference += order, src/moi.jl:88-118) and issue the bulk stores listed by
//     the helper.
// Helper h serves compute warps h and h+4 (all three live on SM sub-partition h, whose 512
// registers per lane are split 40 + 232 + 232). Two input stages per compute warp, full/empty
// mbarriers per stage; the output staging is single-buffered and guarded by
// cp.async.bulk.wait_group.read.
#pragma once

#define DTO_WS_COMPUTE 8
#define DTO_WS_HELPERS 4
#define DTO_WS_DESC_DOUBLES 192  /* 3 x int4 per item, 32 items */
#define DTO_WS_PIECE_DOUBLES 64  /* 1 x int4 per piece, 32 pieces */

// yterms = doubles of the per-lane dynamics-term exchange buffer (32 * MAXD when a dynamics Hessian
// reaches next-state rows, else 0); *yt_off receives its offset
template <int MODE>
__host__ __device__ inline int ws_layout(const dto_launch_args& a, int yterms, int* base, int* ioff, int* in_sz, int* stage_sz, int* yt_off)
{
    constexpr bool DO_H = (MODE & DTO_MODE_H) != 0;
    int n = 0;
    for (int k = 0; k < 5; ++k) {
        const bool need = (k == DTO_IN_Z) || (k == DTO_IN_W && a.w_flat) || (DO_H && k != DTO_IN_W);
        if (ioff) ioff[k] = n;
        if (need) n += a.in_cap[k];
    }
    if (in_sz) *in_sz = n;
    const int st = n + DTO_WS_DESC_DOUBLES + DTO_WS_PIECE_DOUBLES;
    if (stage_sz) *stage_sz = st;
    int off = 4 + 2 * st;  // four mbarriers, two stages
    if (yt_off) *yt_off = off;
    off += (yterms + 1) & ~1;
    for (int s = 0; s < 6; ++s) {
        if (seg_active<MODE>(s)) {
            // Hessian terms stay in registers here: the HTERM segment stages slot values only
            const int pad = s == DTO_SEG_HTERM ? 0 : (a.seg_pad[s] + 1) & ~1;
            const int cap = s == DTO_SEG_HTERM ? a.hslot_cap : a.seg_cap[s];
            if (base) base[s] = off + pad;
            off += pad + ((cap + 1) & ~1) + 2 * a.nsub_max + 2;
        } else if (base) {
            base[s] = 0;
        }
    }
    return (off + 1) & ~1;
}

template <class M, int MODE>
__global__ void __launch_bounds__((DTO_WS_COMPUTE + DTO_WS_HELPERS) * 32, 1) knot_kernel_ws(const __grid_constant__ dto_launch_args a)
{
    extern __shared__ __align__(16) double dto_smem[];
    constexpr bool DO_G = (MODE & DTO_MODE_G) != 0, DO_C = (MODE & DTO_MODE_C) != 0;
    constexpr bool DO_J = (MODE & DTO_MODE_J) != 0, DO_H = (MODE & DTO_MODE_H) != 0;
    constexpr bool HALO = DO_H && (M::HESS_HALO != 0);
    constexpr int OWN = HALO ? 31 : 32;
    constexpr bool HG = DO_H && (M::HG_NCLASS > 0);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
    const int T = a.T;
    const int total = (int)(a.B * T);
    const int tiles = (total + OWN - 1) / OWN;

    const int kt_doubles = a.kt_smem ? (T + 1) * 8 : 0;
    if (a.kt_smem) {
        const int4* src = reinterpret_cast<const int4*>(a.knot);
        int4* dst = reinterpret_cast<int4*>(dto_smem);
        for (int i = threadIdx.x; i < (T + 1) * 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const dto_knot_entry* tab = a.kt_smem ? reinterpret_cast<const dto_knot_entry*>(dto_smem) : a.knot;

    constexpr int YTERMS = HALO ? 32 * M::MAXD : 0;
    int base[6], ioff[5], in_sz, stage_sz, yt_off;
    const int per_warp = ws_layout<MODE>(a, YTERMS, base, ioff, &in_sz, &stage_sz, &yt_off);
    if (warp >= DTO_WS_HELPERS && lane == 0) {
        const uint32_t bar = smem_u32(dto_smem + kt_doubles + (size_t)(warp - DTO_WS_HELPERS) * per_warp);
        mbar_init(bar, 1);
        mbar_init(bar + 8, 1);
        mbar_init(bar + 16, 1);
        mbar_init(bar + 24, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    __syncthreads();
    const int stride = gridDim.x * DTO_WS_COMPUTE;

    if (warp < DTO_WS_HELPERS) {
        // =============================== helper warp ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DTO_WS_HREG));
        for (int k = 0;; ++k) {
            bool any = false;
#pragma unroll 1
            for (int ci = 0; ci < DTO_WS_COMPUTE / DTO_WS_HELPERS; ++ci) {
                const int c = warp + ci * DTO_WS_HELPERS;
                const int tile = blockIdx.x * DTO_WS_COMPUTE + c + k * stride;
                if (tile >= tiles) continue;
                any = true;
                double* smc = dto_smem + kt_doubles + (size_t)c * per_warp;
                const uint32_t bar0 = smem_u32(smc);
                const int st = k & 1;
                const uint32_t full = bar0 + st * 8, empty = bar0 + 16 + st * 8;
                const int sbase = 4 + st * stage_sz;   // stage offset inside the region (doubles)
                mbar_wait(empty, ((k >> 1) & 1) ^ 1);  // the compute warp released this stage (passes on first use)

                const tile_t q = tile_geom<HALO>(a, tile, total);
                const dto_knot_entry kb = ld_knot(tab, 0), kT = ld_knot(tab, T);
                const dto_knot_entry k0 = ld_knot(tab, q.t0), ktf = ld_knot(tab, q.tf);
                // ---- (1) bulk input copies, one lane per range ----
                {
                    const double* src = nullptr;
                    int len = 0, slot = 0;
                    if (lane == 0) {
                        const dto_knot_entry kl1 = ld_knot(tab, q.tl + 1);
                        src = a.z + (size_t)q.b0 * a.N_z + ktf.zofs;
                        len = (q.bl - q.b0) * a.N_z + kl1.zofs + (q.tl + 1 < T ? kl1.nx : 0) - ktf.zofs;
                        slot = ioff[DTO_IN_Z];
                    } else if (lane == 1) {
                        if (DO_H) {
                            src = a.sigma + q.b0;
                            len = q.nsub;
                            slot = ioff[DTO_IN_SIGMA];
                        }
                    } else if (lane == 2) {
                        if (a.w_flat) {
                            const dto_knot_entry kl = ld_knot(tab, q.tl);
                            src = a.w + (size_t)q.b0 * a.N_w + ktf.wofs;
                            len = (q.bl - q.b0) * a.N_w + kl.wofs + kl.pad0 - ktf.wofs;
                            slot = ioff[DTO_IN_W];
                        }
                    } else if (DO_H) {
                        const int j = (lane - 3) >> 1, sg = (lane - 3) & 1;
                        if (j < q.nsub) {
                            const dto_knot_entry ea = ld_knot(tab, j == 0 ? q.tf : 0);
                            const dto_knot_entry eb = ld_knot(tab, j == q.nsub - 1 ? q.tl + 1 : T);
                            const int r0 = sg ? ea.rstage : ea.rdyn, r1 = sg ? eb.rstage : eb.rdyn;
                            const int rT = sg ? kT.rstage : kT.rdyn;
                            const int Ls = sg ? kT.rstage - kT.rdyn : kT.rdyn;  // rows per problem (stage rows follow the dynamics rows)
                            const int flat = j == 0 ? 0 : (rT - (sg ? ktf.rstage : ktf.rdyn)) + (j - 1) * Ls;
                            src = a.lam + (size_t)(q.b0 + j) * a.N_c + r0;
                            len = r1 - r0;
                            slot = ioff[sg ? DTO_IN_LSTAGE : DTO_IN_LDYN] + ((flat + 1) & ~1) + 2 * j;
                        }
                    }
                    if (len > 0) {
                        const int mis = ptr_parity(src);  // the range lands at slot + mis: same 16-byte phase as in HBM
                        const uint32_t bytes = (uint32_t)((len + mis + 1) & ~1) * 8u;
                        mbar_expect_tx(full, bytes);
                        bulk_load(smem_u32(smc + sbase + slot), src - mis, bytes, full);
                    }
                }
                // ---- (2) item descriptors: everything the compute lane needs, as region offsets ----
                {
                    const item_t m = tile_item<HALO>(a, q, lane);
                    const dto_knot_entry ke = ld_knot(tab, m.t);
                    const int kn_zofs = ld_knot(tab, m.t + 1).zofs;
                    const int b = m.b, db = m.db;
                    const int x_off =
                        sbase + ioff[DTO_IN_Z] + (((q.b0 & a.N_z) ^ ktf.zofs ^ ptr_parity(a.z)) & 1) + db * a.N_z + (ke.zofs - ktf.zofs);
                    const int y_off = x_off + (kn_zofs - ke.zofs);
                    const int w_off =
                        a.w_flat ? sbase + ioff[DTO_IN_W] + (((q.b0 & a.N_w) ^ ktf.wofs ^ ptr_parity(a.w)) & 1) + db * a.N_w + (ke.wofs - ktf.wofs)
                                 : ke.wofs;
                    int ld_off = 0, ls_off = 0, sg_off = 0;
                    if (DO_H) {
                        const int pl = ptr_parity(a.lam);
                        ld_off = sbase + ioff[DTO_IN_LDYN] + piece_off(db, b, a.N_c, pl, ktf.rdyn, kb.rdyn, kT.rdyn, ke.rdyn);
                        ls_off = sbase + ioff[DTO_IN_LSTAGE] + piece_off(db, b, a.N_c, pl, ktf.rstage, kb.rstage, kT.rstage, ke.rstage);
                        sg_off = sbase + ioff[DTO_IN_SIGMA] + ((q.b0 ^ ptr_parity(a.sigma)) & 1) + db;
                    }
                    const int g_off = DO_G ? base[DTO_SEG_G] + piece_off(db, b, a.N_z, ptr_parity(a.g), k0.zofs, kb.zofs, kT.zofs, ke.zofs) : 0;
                    const int cd_off = DO_C ? base[DTO_SEG_CDYN] + piece_off(db, b, a.N_c, ptr_parity(a.c), k0.rdyn, kb.rdyn, kT.rdyn, ke.rdyn) : 0;
                    const int cs_off =
                        DO_C ? base[DTO_SEG_CSTAGE] + piece_off(db, b, a.N_c, ptr_parity(a.c), k0.rstage, kb.rstage, kT.rstage, ke.rstage) : 0;
                    const int jd_off = DO_J ? base[DTO_SEG_JDYN] + piece_off(db, b, a.nnz_J, ptr_parity(a.J), k0.jdyn, kb.jdyn, kT.jdyn, ke.jdyn) : 0;
                    const int js_off =
                        DO_J ? base[DTO_SEG_JSTAGE] + piece_off(db, b, a.nnz_J, ptr_parity(a.J), k0.jstage, kb.jstage, kT.jstage, ke.jstage) : 0;
                    const int hd_off =
                        DO_H ? base[DTO_SEG_HTERM] + piece_off(db, b, a.nnz_H, ptr_parity(a.H), k0.hslot, kb.hslot, kT.hslot, ke.hslot) : 0;
                    const int flags = (m.in ? 1 : 0) | (m.own ? 2 : 0);
                    int4 d0, d1, d2;
                    d0.x = x_off | (ke.nx << 16);
                    d0.y = y_off | (w_off << 16);
                    d0.z = ld_off | (ls_off << 16);
                    d0.w = sg_off | (flags << 16);
                    d1.x = (ke.kcost & 255) | ((ke.kdyn & 255) << 8) | ((ke.kstage & 255) << 16) | ((ke.hclass & 255) << 24);
                    d1.y = g_off | (cd_off << 16);
                    d1.z = cs_off | (jd_off << 16);
                    d1.w = js_off;
                    d2.x = hd_off;
                    d2.y = b;
                    d2.z = m.t;
                    d2.w = ke.hslot;
                    int4* dsc = reinterpret_cast<int4*>(smc + sbase + in_sz);
                    dsc[lane] = d0;
                    dsc[32 + lane] = d1;
                    dsc[64 + lane] = d2;
                }
                // ---- (3) output piece list: lane = 8*segment + problem ----
                {
                    const int si = lane >> 3, j = lane & 7;
                    int sg = -1;
                    {
                        int cnt = 0;
#pragma unroll
                        for (int r = 0; r < 6; ++r)
                            if (seg_active<MODE>(r) && (r != DTO_SEG_HTERM || HG)) {
                                if (cnt == si) sg = r;
                                ++cnt;
                            }
                    }
                    int4 pd = make_int4(0, 0, 0, 0);
                    if (sg >= 0 && j < q.nsub) {
                        const dto_knot_entry ea = ld_knot(tab, j == 0 ? q.t0 : 0);
                        const dto_knot_entry eb = ld_knot(tab, j == q.nsub - 1 ? q.tl + 1 : T);
                        const int x0 = seg_field(ea, sg);
                        const int len = seg_field(eb, sg) - x0;
                        const bool isc = sg == DTO_SEG_CDYN || sg == DTO_SEG_CSTAGE, isj = sg == DTO_SEG_JDYN || sg == DTO_SEG_JSTAGE;
                        const int N_s = sg == DTO_SEG_G ? a.N_z : isc ? a.N_c : isj ? a.nnz_J : a.nnz_H;
                        double* arr = sg == DTO_SEG_G ? a.g : isc ? a.c : isj ? a.J : a.H;
                        double* dst = arr + (size_t)(q.b0 + j) * N_s + x0;
                        const int bs = sg == DTO_SEG_G ? base[0] : sg == DTO_SEG_CDYN ? base[1] : sg == DTO_SEG_CSTAGE ? base[2]
                                       : sg == DTO_SEG_JDYN ? base[3] : sg == DTO_SEG_JSTAGE ? base[4] : base[5];
                        const int so =
                            bs + piece_off(j, q.b0 + j, N_s, ptr_parity(arr), seg_field(k0, sg), seg_field(kb, sg), seg_field(kT, sg), x0);
                        const unsigned long long dp = reinterpret_cast<unsigned long long>(dst);
                        pd.x = (int)(unsigned)(dp & 0xffffffffull);
                        pd.y = (int)(unsigned)(dp >> 32);
                        pd.z = so;
                        pd.w = len;
                    }
                    reinterpret_cast<int4*>(smc + sbase + in_sz + DTO_WS_DESC_DOUBLES)[lane] = pd;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
            }
            if (!any) break;
        }
    } else {
        // =============================== compute warp ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DTO_WS_CREG));
        const int c = warp - DTO_WS_HELPERS;
        double* __restrict__ smc = dto_smem + kt_doubles + (size_t)c * per_warp;
        const uint32_t bar0 = smem_u32(smc);
        const bool hg_on = HG && a.use_hclass;
        int k = 0;
        for (int tile = blockIdx.x * DTO_WS_COMPUTE + c; tile < tiles; tile += stride, ++k) {
            const int st = k & 1;
            const int4* dsc = reinterpret_cast<const int4*>(smc + 4 + st * stage_sz + in_sz);
            mbar_wait(bar0 + st * 8, (k >> 1) & 1);  // inputs + descriptors of this tile are in the stage
            bulk_wait_read();                        // the previous tile's output staging has been read
            __syncwarp();
            {
                const int4 d0 = dsc[lane], d1 = dsc[32 + lane], d2 = dsc[64 + lane];
                const int flags = d0.w >> 16;
                const bool own = (flags & 2) != 0;
                // Hessian terms of this lane's knot, by role: registers (all indices are compile-time)
                double tc[M::MAXC], td[M::MAXD], ts[M::MAXS];
                double* yt = smc + yt_off + lane * M::MAXD;  // dynamics terms for the next lane's gather
                if (flags & 1) {
                    const double* __restrict__ x = smc + (d0.x & 0xffff);
                    const double* __restrict__ u = x + (d0.x >> 16);
                    const double* __restrict__ y = smc + (d0.y & 0xffff);
                    const int kcost = d1.x & 255, kdyn = (d1.x >> 8) & 255, kstage = (d1.x >> 16) & 255;
                    const double* __restrict__ w = smc + ((unsigned)d0.y >> 16);
                    if (!a.w_flat) w = a.w + (size_t)d2.y * a.N_w + ((unsigned)d0.y >> 16);
                    const double* __restrict__ lam_d = smc + (d0.z & 0xffff);
                    const double* __restrict__ lam_s = smc + ((unsigned)d0.z >> 16);
                    if (own) {
                        if (DO_G) M::cost_grad(kcost, x, u, w, smc + (d1.y & 0xffff));
                        if (DO_H) M::cost_hess(kcost, x, u, w, smc[d0.w & 0xffff], tc);
                    }
                    if (kdyn != 255) {
                        if (DO_C) M::dyn_res(kdyn, y, x, u, w, smc + ((unsigned)d1.y >> 16));
                        double* jd = smc + ((unsigned)d1.z >> 16);
                        if (DO_J && DO_H) M::dyn_jac_hess(kdyn, y, x, u, w, lam_d, jd, td);
                        else if (DO_J) M::dyn_jac(kdyn, y, x, u, w, jd);
                        else if (DO_H) M::dyn_hess(kdyn, y, x, u, w, lam_d, td);
                        if (HALO) {
#pragma unroll
                            for (int i = 0; i < M::MAXD; ++i) yt[i] = td[i];
                        }
                    }
                    if (own && kstage != 255) {
                        if (DO_C) M::stage_res(kstage, x, u, w, smc + (d1.z & 0xffff));
                        double* js = smc + (d1.w & 0xffff);
                        if (DO_J && DO_H) M::stage_jac_hess(kstage, x, u, w, lam_s, js, ts);
                        else if (DO_J) M::stage_jac(kstage, x, u, w, js);
                        else if (DO_H) M::stage_hess(kstage, x, u, w, lam_s, ts);
                    }
                }
                if (HALO) __syncwarp();
                // Hessian slots of this lane's knot: own terms from registers + the previous knot's
                // dynamics terms, summed in the reference's += order (src/moi.jl:88-118)
                if (hg_on && own) {
                    const int hclass = (int)((unsigned)d1.x >> 24);
                    double v[M::HG_VMAX > 0 ? M::HG_VMAX : 1];
                    double* dst = smc + (d2.x & 0xffff);
                    M::hg_compute_r(hclass, tc, td, ts, yt - M::MAXD, v);
                    M::hg_store(hclass, v, dst);
                    if (a.gen_nhess > 0) {
                        const int b = d2.y, t = d2.z;
                        const int p0 = __ldg(a.gh_ptr + t), p1 = __ldg(a.gh_ptr + t + 1);
                        for (int p = p0; p < p1; ++p) {
                            const int2 e = __ldg(reinterpret_cast<const int2*>(a.gh_ent) + p);  // slot, instance
                            const int4 inst = __ldg(reinterpret_cast<const int4*>(a.gen_inst[2]) + e.y);
                            const double val = M::gen_eval(2, inst.x, a.z + (size_t)b * a.N_z + inst.y, a.w + (size_t)b * a.N_w + inst.z,
                                                           a.lam + (size_t)b * a.N_c + a.gen_row0 + inst.w);
                            dst[e.x - d2.w] += val;
                        }
                    }
                }
            }
            fence_async_smem();  // this lane's generic-proxy writes -> visible to the bulk-store engine
            __syncwarp();
            {   // stream-out from the helper's piece list
                const int4 pd = reinterpret_cast<const int4*>(smc + 4 + st * stage_sz + in_sz + DTO_WS_DESC_DOUBLES)[lane];
                int len = pd.w;
                if (len > 0) {
                    double* dst = reinterpret_cast<double*>(((unsigned long long)(unsigned)pd.y << 32) | (unsigned long long)(unsigned)pd.x);
                    const double* sp = smc + pd.z;
                    if (ptr_parity(dst)) {  // odd position: single head store
                        *dst = *sp;
                        ++dst; ++sp; --len;
                    }
                    if (len & 1) {
                        dst[len - 1] = sp[len - 1];
                        --len;
                    }
                    if (len > 0) bulk_store(dst, smem_u32(sp), (uint32_t)len * 8u);
                }
                bulk_commit();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 16 + st * 8);  // stage free for the helper
        }
        bulk_wait_all();
    }
}

// launch plan: shared memory, or 0 if the shape / mode is not covered by the specialised kernel
template <class M, int MODE>
inline int64_t plan_ws(dto_launch_args& b)
{
    constexpr bool DO_H = (MODE & DTO_MODE_H) != 0;
    if (!DTO_WS || !b.persist_ok) return 0;
    if (DO_H && !(M::HG_NCLASS > 0 && b.use_hclass)) return 0;     // table gather: other kernels
    if (M::N_KINDS_MAX >= 255 || M::HG_NCLASS >= 255) return 0;    // descriptor packs kinds in 8 bits
    if (!b.w_flat && b.N_w > 65535) return 0;
    constexpr bool HALO = DO_H && (M::HESS_HALO != 0);
    const int64_t per_warp = (int64_t)ws_layout<MODE>(b, HALO ? 32 * M::MAXD : 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (per_warp > 65535) return 0;                                // 16-bit region offsets
    for (int kt = 1; kt >= 0; --kt) {
        if (kt && b.T + 1 > DTO_KT_SMEM_MAX) continue;
        const int64_t smem = (kt ? (int64_t)(b.T + 1) * 64 : 0) + per_warp * 8 * DTO_WS_COMPUTE;
        if (smem <= DTO_SMEM_LIMIT - 1024) {
            b.kt_smem = kt;
            return smem;
        }
    }
    return 0;
}
