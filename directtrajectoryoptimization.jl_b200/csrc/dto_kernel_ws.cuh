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
// registers per lane are split DTO_WS_HREG + 2 x DTO_WS_CREG). Per compute warp: two input stages
// (mbarrier full[2]) and one or two output staging buffers (mbarriers out_full / out_empty), two when
// shared memory allows. Round kk of a helper, per served compute warp: wait out_full(kk-2) -> issue
// that tile's bulk stores -> produce tile kk into the input stage tile kk-2 just vacated -> wait for
// the stores' shared-memory reads -> arrive out_empty. The compute warp only waits on full(k) and
// out_empty(k - NOUT), both normally long since complete.
#pragma once

#define DTO_WS_COMPUTE 8
#ifndef DTO_WS_HELPERS
#define DTO_WS_HELPERS 4         /* 4: helper h serves compute warps h, h+4; 8: one helper per compute warp */
#endif
#define DTO_WS_DESC_DOUBLES 192  /* 3 x int4 per item, 32 items */
#define DTO_WS_PIECE_DOUBLES 64  /* 1 x int4 per piece, 32 pieces */

template <int MODE>
__host__ __device__ inline int ws_layout(const dto_launch_args& a, int* base, int* ioff, int* in_sz, int* stage_sz, int* out0, int* out_sz)
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
    int off = 8 + 2 * st;  // six mbarriers (8 doubles reserved), two input stages
    const int out_begin = off;
    if (out0) *out0 = off;
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
    off = (off + 1) & ~1;
    if (out_sz) *out_sz = off - out_begin;
    return off;  // size with ONE output buffer; a second one adds *out_sz
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

    int base[6], ioff[5], in_sz, stage_sz, out0, out_sz;
    const int NOUT = a.ws_nout;  // output staging buffers per compute warp (1 or 2)
    const int per_warp1 = ws_layout<MODE>(a, base, ioff, &in_sz, &stage_sz, &out0, &out_sz);
    const int per_warp = per_warp1 + (NOUT - 1) * out_sz;
    if (warp >= DTO_WS_HELPERS && lane == 0) {
        const uint32_t bar = smem_u32(dto_smem + kt_doubles + (size_t)(warp - DTO_WS_HELPERS) * per_warp);
        for (int i = 0; i < 6; ++i) mbar_init(bar + 8 * i, 1);  // full[2], out_full[2], out_empty[2]
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    __syncthreads();
    const int stride = gridDim.x * DTO_WS_COMPUTE;

    if (warp < DTO_WS_HELPERS) {
        // =============================== helper warp ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DTO_WS_HREG));
        for (int kk = 0;; ++kk) {
            bool any = false;
#pragma unroll 1
            for (int ci = 0; ci < DTO_WS_COMPUTE / DTO_WS_HELPERS; ++ci) {
                const int c = warp + ci * DTO_WS_HELPERS;
                const int tile0 = blockIdx.x * DTO_WS_COMPUTE + c;
                const int tile = tile0 + kk * stride;                                // tile to produce
                const bool have_drain = kk >= 2 && tile0 + (kk - 2) * stride < tiles;  // tile kk-2 to drain
                const bool have_prod = tile < tiles;
                if (!have_drain && !have_prod) continue;
                any = true;
                double* smc = dto_smem + kt_doubles + (size_t)c * per_warp;
                const uint32_t bar0 = smem_u32(smc);
                const int st = kk & 1;  // input stage of tile kk and of tile kk-2
                const uint32_t full = bar0 + st * 8;
                const int sbase = 8 + st * stage_sz;  // stage offset inside the region (doubles)
                const int od = (kk - 2) & (NOUT - 1), nd = (kk - 2) >> (NOUT - 1);  // output buffer / use count of tile kk-2 (NOUT is 1 or 2)
                if (have_drain) {
                    // ---- (0) tile kk-2 is complete in output buffer od: issue its stores (piece list of stage st)
                    mbar_wait(bar0 + 16 + od * 8, nd & 1);
                    const int4 pd = reinterpret_cast<const int4*>(smc + sbase + in_sz + DTO_WS_DESC_DOUBLES)[lane];
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
                    __syncwarp();  // every lane has read its piece before the stage is refilled
                }
                if (have_prod) {
                    const int ob = (kk & (NOUT - 1)) * out_sz;  // output buffer of tile kk: offset added to the segment bases
                    const tile_t q = tile_geom<HALO>(a, tile, total);
                    const int* tabw = reinterpret_cast<const int*>(tab);  // knot entry = 16 words
                    // ---- (1) flat input ranges z / sigma / w: lanes 0..2 ----
                    {
                        const double* src = nullptr;
                        int len = 0, slot = 0;
                        const int tfz = tabw[q.tf * 16 + 0], tfw = tabw[q.tf * 16 + 2];
                        if (lane == 0) {
                            const int e1z = tabw[(q.tl + 1) * 16 + 0], e1n = tabw[(q.tl + 1) * 16 + 1];
                            src = a.z + (size_t)q.b0 * a.N_z + tfz;
                            len = (q.bl - q.b0) * a.N_z + e1z + (q.tl + 1 < T ? e1n : 0) - tfz;
                            slot = ioff[DTO_IN_Z];
                        } else if (lane == 1) {
                            if (DO_H) {
                                src = a.sigma + q.b0;
                                len = q.nsub;
                                slot = ioff[DTO_IN_SIGMA];
                            }
                        } else if (lane == 2) {
                            if (a.w_flat) {
                                src = a.w + (size_t)q.b0 * a.N_w + tfw;
                                len = (q.bl - q.b0) * a.N_w + tabw[q.tl * 16 + 2] + tabw[q.tl * 16 + 14] - tfw;
                                slot = ioff[DTO_IN_W];
                            }
                        }
                        if (len > 0) {
                            const int mis = ptr_parity(src);  // the range lands at slot + mis: same 16-byte phase as in HBM
                            const uint32_t bytes = (uint32_t)((len + mis + 1) & ~1) * 8u;
                            mbar_expect_tx(full, bytes);
                            bulk_load(smem_u32(smc + sbase + slot), src - mis, bytes, full);
                        }
                    }
                    // ---- (2) per (segment row, problem) ranges: lane = 4*row + j. Rows 0..5 are the output
                    // segments (G, CDYN, CSTAGE, JDYN, JSTAGE, Hessian slots), rows 6,7 the dynamics / stage
                    // multipliers. Each lane derives where its range starts in HBM and in the region; cst =
                    // start - (prefix field of the range's first knot), so an item of that range sits at
                    // cst + (its own prefix field).
                    int cst = 0;
                    {
                        const int row = lane >> 2, j = lane & 3;
                        const bool is_in = row >= 6;
                        const bool active = is_in ? DO_H : (seg_active<MODE>(row) && (row != DTO_SEG_HTERM || HG));
                        int4 pd = make_int4(0, 0, 0, 0);
                        if (active && j < q.nsub) {
                            const int wi = (0x76B98760u >> (4 * row)) & 15;  // word of the row's prefix field in a knot entry
                            const int tfirst = is_in ? q.tf : q.t0;
                            const int kfx = tabw[tfirst * 16 + wi], kTx = tabw[T * 16 + wi], kbx = tabw[wi];
                            const int ea = j == 0 ? kfx : kbx;
                            const int eb = tabw[(j == q.nsub - 1 ? q.tl + 1 : T) * 16 + wi];
                            const int len = eb - ea;
                            const bool isc = row == 1 || row == 2 || is_in, isj = row == 3 || row == 4;
                            const int N_s = row == 0 ? a.N_z : isc ? a.N_c : isj ? a.nnz_J : a.nnz_H;
                            const double* arr = row == 0 ? a.g : is_in ? a.lam : isc ? a.c : isj ? a.J : a.H;
                            const double* gp = arr + (size_t)(q.b0 + j) * N_s + ea;
                            const int par = ptr_parity(gp);
                            const int flat = j == 0 ? 0 : (kTx - kfx) + (j - 1) * (kTx - kbx);
                            const int sb = row == 0 ? base[0] : row == 1 ? base[1] : row == 2 ? base[2] : row == 3 ? base[3]
                                           : row == 4 ? base[4] : row == 5 ? base[5] : row == 6 ? ioff[DTO_IN_LDYN] : ioff[DTO_IN_LSTAGE];
                            const int start = (is_in ? sbase : ob) + sb + ((flat + 1) & ~1) + 2 * j + par;
                            cst = start - ea;
                            if (is_in) {
                                if (len > 0) {
                                    const uint32_t bytes = (uint32_t)((len + par + 1) & ~1) * 8u;
                                    mbar_expect_tx(full, bytes);
                                    bulk_load(smem_u32(smc + start - par), gp - par, bytes, full);
                                }
                            } else {
                                const unsigned long long dp = reinterpret_cast<unsigned long long>(gp);
                                pd.x = (int)(unsigned)(dp & 0xffffffffull);
                                pd.y = (int)(unsigned)(dp >> 32);
                                pd.z = start;
                                pd.w = len;
                            }
                        }
                        reinterpret_cast<int4*>(smc + sbase + in_sz + DTO_WS_DESC_DOUBLES)[lane] = pd;
                    }
                    // ---- (3) item descriptors: region offsets of everything the compute lane touches ----
                    {
                        const item_t m = tile_item<HALO>(a, q, lane);
                        const dto_knot_entry ke = ld_knot(tab, m.t);
                        const int kn_zofs = tabw[(m.t + 1) * 16];
                        const int db = m.db;
                        const int tfz = tabw[q.tf * 16 + 0], tfw = tabw[q.tf * 16 + 2];
                        const int x_off = sbase + ioff[DTO_IN_Z] + (((q.b0 & a.N_z) ^ tfz ^ ptr_parity(a.z)) & 1) + db * a.N_z + (ke.zofs - tfz);
                        const int y_off = x_off + (kn_zofs - ke.zofs);
                        const int w_off =
                            a.w_flat ? sbase + ioff[DTO_IN_W] + (((q.b0 & a.N_w) ^ tfw ^ ptr_parity(a.w)) & 1) + db * a.N_w + (ke.wofs - tfw) : ke.wofs;
                        const int g_off = __shfl_sync(0xffffffffu, cst, 0 * 4 + db) + ke.zofs;
                        const int cd_off = __shfl_sync(0xffffffffu, cst, 1 * 4 + db) + ke.rdyn;
                        const int cs_off = __shfl_sync(0xffffffffu, cst, 2 * 4 + db) + ke.rstage;
                        const int jd_off = __shfl_sync(0xffffffffu, cst, 3 * 4 + db) + ke.jdyn;
                        const int js_off = __shfl_sync(0xffffffffu, cst, 4 * 4 + db) + ke.jstage;
                        const int hd_off = __shfl_sync(0xffffffffu, cst, 5 * 4 + db) + ke.hslot;
                        const int ld_off = __shfl_sync(0xffffffffu, cst, 6 * 4 + db) + ke.rdyn;
                        const int ls_off = __shfl_sync(0xffffffffu, cst, 7 * 4 + db) + ke.rstage;
                        const int sg_off = sbase + ioff[DTO_IN_SIGMA] + ((q.b0 ^ ptr_parity(a.sigma)) & 1) + db;
                        const int flags = (m.in ? 1 : 0) | (m.own ? 2 : 0);
                        int4 d0, d1, d2;
                        d0.x = x_off | (ke.nx << 16);
                        d0.y = y_off | (w_off << 16);
                        d0.z = DO_H ? (ld_off | (ls_off << 16)) : 0;
                        d0.w = (DO_H ? sg_off : 0) | (flags << 16);
                        d1.x = (ke.kcost & 255) | ((ke.kdyn & 255) << 8) | ((ke.kstage & 255) << 16) | ((ke.hclass & 255) << 24);
                        d1.y = (DO_G ? g_off : 0) | ((DO_C ? cd_off : 0) << 16);
                        d1.z = (DO_C ? cs_off : 0) | ((DO_J ? jd_off : 0) << 16);
                        d1.w = DO_J ? js_off : 0;
                        d2.x = DO_H ? hd_off : 0;
                        d2.y = m.b;
                        d2.z = m.t;
                        d2.w = ke.hslot;
                        int4* dsc = reinterpret_cast<int4*>(smc + sbase + in_sz);
                        dsc[lane] = d0;
                        dsc[32 + lane] = d1;
                        dsc[64 + lane] = d2;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full);
                }  // have_prod
                if (have_drain) {
                    bulk_wait_read();  // the stores of tile kk-2 have read their shared-memory source
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar0 + 32 + od * 8);  // out_empty: buffer od may be overwritten
                }
            }
            if (!any && kk >= 2) break;  // (a lone tile is produced in round 0 and drained in round 2)
        }
        bulk_wait_all();
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
            const int o = k & (NOUT - 1), no = k >> (NOUT - 1);
            const int4* dsc = reinterpret_cast<const int4*>(smc + 8 + st * stage_sz + in_sz);
            mbar_wait(bar0 + st * 8, (k >> 1) & 1);          // inputs + descriptors of this tile are in the stage
            mbar_wait(bar0 + 32 + o * 8, (no & 1) ^ 1);      // output buffer o has been drained (passes on first use)
            {
                const int4 d0 = dsc[lane], d1 = dsc[32 + lane], d2 = dsc[64 + lane];
                const int flags = d0.w >> 16;
                const bool own = (flags & 2) != 0;
                // Hessian terms of this lane's knot, by role: registers (all indices are compile-time)
                double tc[M::MAXC], td[M::MAXD], ts[M::MAXS];
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
                    }
                    if (own && kstage != 255) {
                        if (DO_C) M::stage_res(kstage, x, u, w, smc + (d1.z & 0xffff));
                        double* js = smc + (d1.w & 0xffff);
                        if (DO_J && DO_H) M::stage_jac_hess(kstage, x, u, w, lam_s, js, ts);
                        else if (DO_J) M::stage_jac(kstage, x, u, w, js);
                        else if (DO_H) M::stage_hess(kstage, x, u, w, lam_s, ts);
                    }
                }
                // the previous knot's dynamics terms that feed this knot's rows come from the lane below
                // (lane = consecutive items; with a halo, lane 0 is the item before the tile)
                double pd[M::MAXD];
                if (HALO) M::hg_prev_exchange(td, pd);
                // Hessian slots of this lane's knot: own terms from registers + the previous knot's
                // dynamics terms, summed in the reference's += order (src/moi.jl:88-118)
                if (hg_on && own) {
                    const int hclass = (int)((unsigned)d1.x >> 24);
                    double v[M::HG_VMAX > 0 ? M::HG_VMAX : 1];
                    double* dst = smc + (d2.x & 0xffff);
                    M::hg_compute_r(hclass, tc, td, ts, pd, v);
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
            if (lane == 0) mbar_arrive(bar0 + 16 + o * 8);  // out_full: the helper streams the tile out and refills the stage
        }
    }
}

// launch plan: shared memory, or 0 if the shape / mode is not covered by the specialised kernel
template <class M, int MODE>
inline int64_t plan_ws(dto_launch_args& b)
{
    constexpr bool DO_H = (MODE & DTO_MODE_H) != 0;
    if (!DTO_WS || !b.persist_ok || b.nsub_max > 4) return 0;
    // light, HBM-bound work (models with few FP64 ops per knot; gradient / residual / Jacobian-only
    // passes of any model): the plain kernel's occupancy wins (profiles/sweep_r01_s_per_kernel.jsonl)
    if (!DO_H || M::OPS_FUSED < DTO_WS_MIN_OPS) return 0;   // (segment, problem) lane map: 8 rows x 4 problems
    if (DO_H && !(M::HG_NCLASS > 0 && b.use_hclass)) return 0;     // table gather: other kernels
    if (M::N_KINDS_MAX >= 255 || M::HG_NCLASS >= 255) return 0;    // descriptor packs kinds in 8 bits
    if (!b.w_flat && b.N_w > 65535) return 0;
    int out_sz = 0;
    const int64_t one = (int64_t)ws_layout<MODE>(b, nullptr, nullptr, nullptr, nullptr, nullptr, &out_sz);
    for (int nout = 2; nout >= 1; --nout) {
        const int64_t per_warp = one + (nout - 1) * (int64_t)out_sz;
        if (per_warp > 65535) continue;  // 16-bit region offsets
        for (int kt = 1; kt >= 0; --kt) {
            if (kt && b.T + 1 > DTO_KT_SMEM_MAX) continue;
            const int64_t smem = (kt ? (int64_t)(b.T + 1) * 64 : 0) + per_warp * 8 * DTO_WS_COMPUTE;
            if (smem <= DTO_SMEM_LIMIT - 1024) {
                b.kt_smem = kt;
                b.ws_nout = nout;
                return smem;
            }
        }
    }
    return 0;
}
