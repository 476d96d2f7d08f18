// libdto.so -- host runtime behind include/dto.h.
//
//  * loads generated model libraries (dlopen) and checks their ABI;
//  * assembles a problem shape: z-layout, constraint rows, Jacobian COO structure, sorted-unique
//    Hessian structure and every per-knot slot table, in O(nnz log nnz) -- the restatement of
//    /root/reference/src/data.jl:61-104,150-220 and the *_indices / sparsity_* helpers of
//    src/dynamics.jl:129-204, src/costs.jl:75-104, src/constraints.jl:106-183,
//    src/general_constraint.jl:93-139 (which are O(T^2)/O(nnz^2) there);
//  * owns device memory, streams and the contiguous batch shards (one per listed device, no
//    collective); keeps z / lambda / sigma / w resident between callbacks
//    (replaces trajectory!/duals!, src/data.jl:258-278);
//  * enqueues the model library's kernels for the five MOI callbacks (src/moi.jl:1-120).
//
// No CPU evaluation path exists in this file by design.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <utility>
#include <chrono>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a tool (nsys / ncu --nvtx) is attached

#include "../../include/dto.h"
#include "dto_model_abi.h"

// NVTX range over one host-API phase (tracing): shows up as "dto:<what>" in a timeline
struct DtoRange {
    explicit DtoRange(const char* name) { nvtxRangePushA(name); }
    ~DtoRange() { nvtxRangePop(); }
    DtoRange(const DtoRange&) = delete;
    DtoRange& operator=(const DtoRange&) = delete;
};

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define DTO_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return fail(e__ == cudaErrorMemoryAllocation ? DTO_ERR_OOM : DTO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                          \
    } while (0)

#define DTO_REQUIRE(cond, ...)                                \
    do {                                                      \
        if (!(cond)) return fail(DTO_ERR_BAD_ARG, __VA_ARGS__); \
    } while (0)

// ------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------
struct dto_model {
    void* handle = nullptr;
    const dto_model_vtable* vt = nullptr;
    std::string path;
};

struct dto_shape {
    dto_model* model = nullptr;
    int32_t T = 0;
    bool use_general = false;
    int64_t N_z = 0, N_c = 0, N_w = 0, nnz_J = 0, nnz_H = 0, nnz_H_nonunique = 0;
    int64_t n_dyn_rows = 0, n_stage_rows = 0, n_gen_rows = 0;
    int64_t n_dyn_jac = 0, n_stage_jac = 0, n_gen_jac = 0;
    bool hessian_available = true;
    std::vector<int32_t> nx, nu;                 // [T]
    std::vector<dto_knot_entry> knot;            // [T+1]
    std::vector<int64_t> jac_row, jac_col;       // 1-based
    std::vector<int64_t> hess_row, hess_col;     // 1-based, sorted unique
    std::vector<int32_t> hptr, hsrc;             // CSR slot -> term ids (knot elements only)
    std::vector<int32_t> hsrc4;                  // [nnz_H][4] packed, -1 padded; slot with no knot term: zero term
    std::vector<int32_t> gen_inst[3];            // [n][4]
    std::vector<int32_t> gen_hslot;
    std::vector<int32_t> gh_ptr, gh_ent;         // general Hessian entries grouped by owning knot
    std::vector<int32_t> gh_rec;                 // the same entries flattened (8 ints each) for the ws kernel's tile plans
    std::vector<double> c_lower, c_upper;
    int32_t seg_cap[6] = {0, 0, 0, 0, 0, 0};
    int32_t seg_pad[6] = {0, 0, 0, 0, 0, 0};
    bool use_hclass = false;
    int32_t in_cap[5] = {0, 0, 0, 0, 0};         // persistent kernel: input staging capacities per warp tile
    int32_t nsub_max = 1;
    int32_t hslot_cap = 0;
    int64_t shape_id = 0;
    bool persist_ok = false, w_flat = false;
};

struct dto_shard {
    int device = 0;
    int64_t begin = 0, size = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    double* arr[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    dto_knot_entry* d_knot = nullptr;
    int32_t* d_hsrc4 = nullptr;
    int32_t* d_gen_inst[3] = {nullptr, nullptr, nullptr};
    int32_t* d_gen_hslot = nullptr;
    int32_t* d_gh_ptr = nullptr;
    int32_t* d_gh_ent = nullptr;
    int32_t* d_gh_rec = nullptr;
    cudaStream_t pipe[4] = {nullptr, nullptr, nullptr, nullptr};  // chunk pipeline of the one-call host path
};

struct dto_batch {
    dto_shape* shape = nullptr;
    int64_t B = 0;
    std::vector<dto_shard> shards;
    bool have_x = false, have_duals = false;
    int64_t launches = 0;
};

static int64_t array_width(const dto_shape* s, int array)
{
    switch (array) {
    case DTO_ARRAY_Z: return s->N_z;
    case DTO_ARRAY_LAMBDA: return s->N_c;
    case DTO_ARRAY_SIGMA: return 1;
    case DTO_ARRAY_W: return s->N_w;
    case DTO_ARRAY_F: return 1;
    case DTO_ARRAY_G: return s->N_z;
    case DTO_ARRAY_C: return s->N_c;
    case DTO_ARRAY_J: return s->nnz_J;
    case DTO_ARRAY_H: return s->nnz_H;
    default: return -1;
    }
}

// ------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------
extern "C" int dto_abi_version(void) { return DTO_ABI_VERSION; }
extern "C" const char* dto_last_error(void) { return g_err; }
extern "C" const char* dto_status_string(int status)
{
    switch (status) {
    case DTO_OK: return "ok";
    case DTO_ERR_BAD_ARG: return "bad argument";
    case DTO_ERR_CUDA: return "CUDA error";
    case DTO_ERR_OOM: return "out of memory";
    case DTO_ERR_MODEL: return "model library error";
    case DTO_ERR_NO_HESSIAN: return "Hessian not available";
    case DTO_ERR_STATE: return "invalid state";
    case DTO_ERR_UNSUPPORTED: return "unsupported shape";
    default: return "unknown status";
    }
}
extern "C" int dto_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
extern "C" int dto_model_load(const char* path, dto_model** out)
{
    if (!path || !out) return fail(DTO_ERR_BAD_ARG, "dto_model_load: null argument");
    *out = nullptr;
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail(DTO_ERR_MODEL, "dto_model_load: dlopen(%s) failed: %s", path, dlerror());
    typedef const dto_model_vtable* (*entry_fn)(void);
    entry_fn entry = (entry_fn)dlsym(h, "dto_model_entry");
    if (!entry) {
        dlclose(h);
        return fail(DTO_ERR_MODEL, "dto_model_load: %s does not export dto_model_entry", path);
    }
    const dto_model_vtable* vt = entry();
    if (!vt || vt->abi_version != DTO_MODEL_ABI_VERSION) {
        int got = vt ? vt->abi_version : -1;
        dlclose(h);
        return fail(DTO_ERR_MODEL, "dto_model_load: %s has model ABI %d, runtime expects %d (rebuild the model)", path, got,
                    DTO_MODEL_ABI_VERSION);
    }
    dto_model* m = new (std::nothrow) dto_model();
    if (!m) {
        dlclose(h);
        return fail(DTO_ERR_OOM, "dto_model_load: out of host memory");
    }
    m->handle = h;
    m->vt = vt;
    m->path = path;
    *out = m;
    return DTO_OK;
}

extern "C" void dto_model_destroy(dto_model* m)
{
    if (!m) return;
    // the library stays mapped: CUDA module teardown at dlclose time is not worth the risk
    delete m;
}
extern "C" const char* dto_model_name(const dto_model* m) { return m ? m->vt->name : ""; }
extern "C" const char* dto_model_hash(const dto_model* m) { return m ? m->vt->source_hash : ""; }

static const dto_element_desc* kind_desc(const dto_model* m, int role, int kind)
{
    const dto_model_vtable* vt = m->vt;
    int n = role == 0 ? vt->n_dyn : role == 1 ? vt->n_cost : role == 2 ? vt->n_stage : 0;
    if (kind < 0 || kind >= n) return nullptr;
    return (role == 0 ? vt->dyn : role == 1 ? vt->cost : vt->stage) + kind;
}

extern "C" int dto_model_num_kinds(const dto_model* m, int role)
{
    if (!m) return fail(DTO_ERR_BAD_ARG, "dto_model_num_kinds: null model");
    switch (role) {
    case 0: return m->vt->n_dyn;
    case 1: return m->vt->n_cost;
    case 2: return m->vt->n_stage;
    default: return fail(DTO_ERR_BAD_ARG, "dto_model_num_kinds: role %d", role);
    }
}

extern "C" int dto_model_kind_dims(const dto_model* m, int role, int kind, int32_t dims[6])
{
    if (!m || !dims) return fail(DTO_ERR_BAD_ARG, "dto_model_kind_dims: null argument");
    const dto_element_desc* d = kind_desc(m, role, kind);
    if (!d) return fail(DTO_ERR_BAD_ARG, "dto_model_kind_dims: no kind %d for role %d", kind, role);
    dims[0] = d->n_out;
    dims[1] = d->nx;
    dims[2] = d->nu;
    dims[3] = d->nw;
    dims[4] = d->nnz_jac;
    dims[5] = d->has_hess ? d->nnz_hess : 0;
    return DTO_OK;
}
extern "C" int dto_model_has_general(const dto_model* m) { return (m && m->vt->general) ? 1 : 0; }

// ------------------------------------------------------------------------------------------
// shape assembly
// ------------------------------------------------------------------------------------------
namespace {
struct Term {
    int64_t row, col;  // 1-based global
    int32_t id;        // term id (knot elements) or general index
};

int32_t cyclic_window_max(const std::vector<int32_t>& size, int W)
{
    const int T = (int)size.size();
    int32_t best = 0;
    for (int t0 = 0; t0 < T; ++t0) {
        int64_t s = 0;
        for (int i = 0; i < W; ++i) s += size[(t0 + i) % T];
        best = std::max<int64_t>(best, s);
    }
    return best;
}
}  // namespace

extern "C" int dto_shape_create(dto_model* m, const dto_shape_desc* d, dto_shape** out)
{
    if (!m || !d || !out) return fail(DTO_ERR_BAD_ARG, "dto_shape_create: null argument");
    *out = nullptr;
    const int T = d->T;
    DTO_REQUIRE(T >= 2, "dto_shape_create: T=%d, need at least 2 knots", T);
    DTO_REQUIRE(d->dynamics_kind && d->cost_kind && d->stage_kind, "dto_shape_create: null kind array");
    const dto_general_desc* gen = nullptr;
    if (d->use_general) {
        gen = m->vt->general;
        if (!gen) return fail(DTO_ERR_MODEL, "dto_shape_create: use_general=1 but model '%s' has no general constraint", m->vt->name);
    }
    dto_shape* s = new (std::nothrow) dto_shape();
    if (!s) return fail(DTO_ERR_OOM, "dto_shape_create: out of host memory");
    struct Guard {
        dto_shape* s;
        ~Guard() { delete s; }
    } guard{s};
    s->model = m;
    s->T = T;
    s->use_general = gen != nullptr;

    // ---- dims (src/dynamics.jl:206-211) and kinds
    std::vector<const dto_element_desc*> dyn(T, nullptr), cost(T, nullptr), stage(T, nullptr);
    for (int t = 0; t < T; ++t) {
        if (t < T - 1) {
            dyn[t] = kind_desc(m, 0, d->dynamics_kind[t]);
            DTO_REQUIRE(dyn[t], "dto_shape_create: dynamics_kind[%d]=%d out of range", t, d->dynamics_kind[t]);
        }
        cost[t] = kind_desc(m, 1, d->cost_kind[t]);
        DTO_REQUIRE(cost[t], "dto_shape_create: cost_kind[%d]=%d out of range", t, d->cost_kind[t]);
        if (d->stage_kind[t] >= 0) {
            stage[t] = kind_desc(m, 2, d->stage_kind[t]);
            DTO_REQUIRE(stage[t], "dto_shape_create: stage_kind[%d]=%d out of range", t, d->stage_kind[t]);
        }
    }
    s->nx.resize(T);
    s->nu.resize(T);
    for (int t = 0; t < T - 1; ++t) {
        s->nx[t] = dyn[t]->nx;
        s->nu[t] = dyn[t]->nu;
        if (t > 0)
            DTO_REQUIRE(dyn[t - 1]->n_out == dyn[t]->nx, "dto_shape_create: dynamics[%d].num_next_state=%d != dynamics[%d].num_state=%d",
                        t - 1, dyn[t - 1]->n_out, t, dyn[t]->nx);
    }
    s->nx[T - 1] = dyn[T - 2]->n_out;
    s->nu[T - 1] = 0;
    for (int t = 0; t < T; ++t) {
        DTO_REQUIRE(cost[t]->nx + cost[t]->nu == s->nx[t] + s->nu[t],
                    "dto_shape_create: objective[%d] has %d+%d variables, knot has %d+%d (gradient slice would not fit, "
                    "src/costs.jl:61)", t, cost[t]->nx, cost[t]->nu, s->nx[t], s->nu[t]);
        if (stage[t]) {
            DTO_REQUIRE(stage[t]->nx == s->nx[t], "dto_shape_create: constraints[%d].num_state=%d != %d", t, stage[t]->nx, s->nx[t]);
            // A Constraint may be declared with more actions than its knot has (examples/pendulum builds the terminal
            // one with num_action = m) as long as it does not USE them: a Jacobian / Hessian entry beyond the knot's
            // [x; u] would read the next problem's z (the reference raises a BoundsError there)
            const int32_t nv = s->nx[t] + s->nu[t];
            for (int32_t e = 0; e < stage[t]->nnz_jac; ++e)
                DTO_REQUIRE(stage[t]->jac_col[e] >= 1 && stage[t]->jac_col[e] <= nv,
                            "dto_shape_create: constraints[%d] uses variable %d of [x; u] but knot %d has only %d+%d (an action "
                            "the knot does not have?)", t, stage[t]->jac_col[e], t, s->nx[t], s->nu[t]);
            for (int32_t e = 0; stage[t]->has_hess && e < stage[t]->nnz_hess; ++e)
                DTO_REQUIRE(stage[t]->hess_row[e] >= 1 && stage[t]->hess_row[e] <= nv && stage[t]->hess_col[e] >= 1 && stage[t]->hess_col[e] <= nv,
                            "dto_shape_create: constraints[%d] has a Hessian entry (%d,%d) outside the knot's %d variables", t,
                            stage[t]->hess_row[e], stage[t]->hess_col[e], nv);
        }
        if (!cost[t]->has_hess) s->hessian_available = false;
    }

    // ---- prefix tables
    s->knot.resize(T + 1);
    std::vector<int32_t> hterm_cost(T, 0), hterm_dyn(T, 0), hterm_stage(T, 0);
    {
        int64_t zofs = 0, rdyn = 0, jdyn = 0, hterm = 0, wsum = 0;
        for (int t = 0; t < T; ++t) {
            dto_knot_entry& k = s->knot[t];
            k.zofs = (int32_t)zofs;
            k.nx = s->nx[t];
            const int32_t pdim = d->parameter_dim ? d->parameter_dim[t] : 0;
            DTO_REQUIRE(pdim >= 0, "dto_shape_create: parameter_dim[%d] negative", t);
            k.wofs = d->parameter_offset ? d->parameter_offset[t] : (int32_t)wsum;
            wsum += pdim;
            if (d->parameter_offset)
                DTO_REQUIRE(k.wofs >= 0 && k.wofs + pdim <= d->num_parameter, "dto_shape_create: parameter slice of knot %d outside [0,%d)", t,
                            d->num_parameter);
            const dto_element_desc* els[3] = {dyn[t], cost[t], stage[t]};
            for (const dto_element_desc* e : els)
                if (e) DTO_REQUIRE(e->nw <= pdim, "dto_shape_create: an element at knot %d reads %d parameters, w_t has %d", t, e->nw, pdim);
            k.kdyn = t < T - 1 ? d->dynamics_kind[t] : -1;
            k.kcost = d->cost_kind[t];
            k.kstage = stage[t] ? d->stage_kind[t] : -1;
            k.rdyn = (int32_t)rdyn;
            k.jdyn = (int32_t)jdyn;
            k.hterm = (int32_t)hterm;
            hterm_cost[t] = cost[t]->has_hess ? cost[t]->nnz_hess : 0;
            hterm_dyn[t] = (dyn[t] && dyn[t]->has_hess) ? dyn[t]->nnz_hess : 0;
            hterm_stage[t] = (stage[t] && stage[t]->has_hess) ? stage[t]->nnz_hess : 0;
            hterm += hterm_cost[t] + hterm_dyn[t] + hterm_stage[t];
            zofs += s->nx[t] + s->nu[t];
            if (dyn[t]) {
                rdyn += dyn[t]->n_out;
                jdyn += dyn[t]->nnz_jac;
            }
        }
        s->N_z = zofs;
        s->N_w = d->parameter_offset ? d->num_parameter : wsum;
        s->n_dyn_rows = rdyn;
        s->n_dyn_jac = jdyn;
        int64_t rstage = rdyn, jstage = jdyn;
        for (int t = 0; t < T; ++t) {
            s->knot[t].rstage = (int32_t)rstage;
            s->knot[t].jstage = (int32_t)jstage;
            if (stage[t]) {
                rstage += stage[t]->n_out;
                jstage += stage[t]->nnz_jac;
            }
        }
        s->n_stage_rows = rstage - rdyn;
        s->n_stage_jac = jstage - jdyn;
        dto_knot_entry& e = s->knot[T];
        e.zofs = (int32_t)zofs;
        e.nx = 0;
        e.wofs = 0;
        e.kdyn = e.kcost = e.kstage = -1;
        e.rdyn = (int32_t)rdyn;
        e.rstage = (int32_t)rstage;
        e.jdyn = (int32_t)jdyn;
        e.jstage = (int32_t)jstage;
        e.hterm = (int32_t)hterm;
        e.hslot = 0;  // filled below
        for (int t = 0; t <= T; ++t) {
            s->knot[t].hclass = -1;
            s->knot[t].hprev = t > 0 ? s->knot[t].hterm - s->knot[t - 1].hterm : 0;
            s->knot[t].pad0 = (t < T && d->parameter_dim) ? d->parameter_dim[t] : 0;  // length of w_t
            s->knot[t].pad1 = 0;
        }
    }
    if (gen) {
        DTO_REQUIRE(gen->num_variables == s->N_z, "dto_shape_create: general constraint built for %d variables, shape has %lld",
                    gen->num_variables, (long long)s->N_z);
        DTO_REQUIRE(gen->num_parameter <= s->N_w, "dto_shape_create: general constraint reads %d parameters, problems carry %lld",
                    gen->num_parameter, (long long)s->N_w);
        s->n_gen_rows = gen->num_constraint;
        s->n_gen_jac = gen->nnz_jac;
    }
    s->N_c = s->n_dyn_rows + s->n_stage_rows + s->n_gen_rows;
    s->nnz_J = s->n_dyn_jac + s->n_stage_jac + s->n_gen_jac;
    DTO_REQUIRE(s->N_z < (1ll << 30) && s->nnz_J < (1ll << 30), "dto_shape_create: shape too large for 32-bit tables");

    // ---- constraint bounds (src/data.jl:135-148)
    s->c_lower.assign(s->N_c, 0.0);
    s->c_upper.assign(s->N_c, 0.0);
    for (int t = 0; t < T; ++t)
        if (stage[t])
            for (int i = 0; i < stage[t]->n_ineq; ++i) {
                const int r = stage[t]->ineq[i];
                DTO_REQUIRE(r >= 1 && r <= stage[t]->n_out, "dto_shape_create: constraints[%d] inequality index %d out of range", t, r);
                s->c_lower[s->knot[t].rstage + r - 1] = -std::numeric_limits<double>::infinity();
            }
    if (gen)
        for (int i = 0; i < gen->n_ineq; ++i) {
            const int r = gen->ineq[i];
            DTO_REQUIRE(r >= 1 && r <= gen->num_constraint, "dto_shape_create: general inequality index %d out of range", r);
            s->c_lower[s->n_dyn_rows + s->n_stage_rows + r - 1] = -std::numeric_limits<double>::infinity();
        }

    // ---- Jacobian COO structure (src/data.jl:170-175, Q3)
    s->jac_row.reserve(s->nnz_J);
    s->jac_col.reserve(s->nnz_J);
    for (int t = 0; t < T - 1; ++t)
        for (int k = 0; k < dyn[t]->nnz_jac; ++k) {
            s->jac_row.push_back(s->knot[t].rdyn + dyn[t]->jac_row[k]);
            s->jac_col.push_back(s->knot[t].zofs + dyn[t]->jac_col[k]);
        }
    for (int t = 0; t < T; ++t)
        if (stage[t])
            for (int k = 0; k < stage[t]->nnz_jac; ++k) {
                s->jac_row.push_back(s->knot[t].rstage + stage[t]->jac_row[k]);
                s->jac_col.push_back(s->knot[t].zofs + stage[t]->jac_col[k]);
            }
    if (gen)
        for (int k = 0; k < gen->nnz_jac; ++k) {
            s->jac_row.push_back(s->n_dyn_rows + s->n_stage_rows + gen->jac_row[k]);
            s->jac_col.push_back(gen->jac_col[k]);
        }

    // ---- Hessian terms in the reference's concatenation order (src/data.jl:178-182, Q4)
    std::vector<Term> terms;       // knot-element terms: id = smem term id
    std::vector<Term> gen_terms;   // general terms: id = index into the general nzval
    for (int t = 0; t < T; ++t)    // objective
        for (int k = 0; k < hterm_cost[t]; ++k)
            terms.push_back({s->knot[t].zofs + cost[t]->hess_row[k], s->knot[t].zofs + cost[t]->hess_col[k], s->knot[t].hterm + k});
    for (int t = 0; t < T - 1; ++t)  // dynamics
        for (int k = 0; k < hterm_dyn[t]; ++k)
            terms.push_back({s->knot[t].zofs + dyn[t]->hess_row[k], s->knot[t].zofs + dyn[t]->hess_col[k],
                             s->knot[t].hterm + hterm_cost[t] + k});
    for (int t = 0; t < T; ++t)  // stage
        for (int k = 0; k < hterm_stage[t]; ++k)
            terms.push_back({s->knot[t].zofs + stage[t]->hess_row[k], s->knot[t].zofs + stage[t]->hess_col[k],
                             s->knot[t].hterm + hterm_cost[t] + hterm_dyn[t] + k});
    if (gen && gen->has_hess)
        for (int k = 0; k < gen->nnz_hess; ++k) gen_terms.push_back({gen->hess_row[k], gen->hess_col[k], k});
    s->nnz_H_nonunique = (int64_t)terms.size() + (int64_t)gen_terms.size();

    // key = sort(unique(list)): lexicographic on (row, col) (src/data.jl:184)
    std::vector<std::pair<int64_t, int64_t>> key;
    key.reserve(terms.size() + gen_terms.size());
    for (const Term& t : terms) key.emplace_back(t.row, t.col);
    for (const Term& t : gen_terms) key.emplace_back(t.row, t.col);
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    s->nnz_H = (int64_t)key.size();
    DTO_REQUIRE(s->nnz_H < (1ll << 30), "dto_shape_create: Hessian too large for 32-bit tables");
    s->hess_row.resize(key.size());
    s->hess_col.resize(key.size());
    for (size_t i = 0; i < key.size(); ++i) {
        s->hess_row[i] = key[i].first;
        s->hess_col[i] = key[i].second;
        DTO_REQUIRE(key[i].first >= 1 && key[i].first <= s->N_z && key[i].second >= 1 && key[i].second <= s->N_z,
                    "dto_shape_create: Hessian entry (%lld,%lld) outside the %lld variables", (long long)key[i].first,
                    (long long)key[i].second, (long long)s->N_z);
    }
    auto slot_of = [&](int64_t r, int64_t c) -> int32_t {
        return (int32_t)(std::lower_bound(key.begin(), key.end(), std::make_pair(r, c)) - key.begin());
    };
    // first slot owned by each knot (rows are variables of knot t)
    for (int t = 0; t <= T; ++t)
        s->knot[t].hslot = (int32_t)(std::lower_bound(key.begin(), key.end(), std::make_pair((int64_t)s->knot[t].zofs + 1, (int64_t)0)) - key.begin());
    // CSR slot -> contributing term ids, in reference += order (stable counting sort)
    s->hptr.assign(s->nnz_H + 1, 0);
    std::vector<int32_t> tslot(terms.size());
    for (size_t i = 0; i < terms.size(); ++i) {
        tslot[i] = slot_of(terms[i].row, terms[i].col);
        s->hptr[tslot[i] + 1]++;
    }
    for (int64_t i = 0; i < s->nnz_H; ++i) s->hptr[i + 1] += s->hptr[i];
    s->hsrc.resize(terms.size());
    {
        std::vector<int32_t> fill(s->hptr.begin(), s->hptr.end() - 1);
        for (size_t i = 0; i < terms.size(); ++i) s->hsrc[fill[tslot[i]]++] = terms[i].id;
    }
    // locality check: a slot owned by knot t may only draw from knot t and dynamics[t-1]
    for (int t = 0; t < T; ++t) {
        const int32_t lo = t > 0 ? s->knot[t - 1].hterm + hterm_cost[t - 1] : s->knot[t].hterm;
        const int32_t lo_end = t > 0 ? lo + hterm_dyn[t - 1] : lo;
        const int32_t own0 = s->knot[t].hterm, own1 = s->knot[t + 1].hterm;
        for (int32_t sl = s->knot[t].hslot; sl < s->knot[t + 1].hslot; ++sl)
            for (int32_t p = s->hptr[sl]; p < s->hptr[sl + 1]; ++p) {
                const int32_t id = s->hsrc[p];
                const bool ok = (id >= own0 && id < own1) || (id >= lo && id < lo_end);
                DTO_REQUIRE(ok, "dto_shape_create: Hessian slot %d (row %lld) draws from an element outside knots %d..%d; "
                                "element reaches beyond its [x;u;y] window", sl, (long long)s->hess_row[sl], t - 1, t);
            }
    }
    // packed gather table: <= 4 contributors per slot (cost_t, dynamics_{t-1}, dynamics_t, stage_t)
    s->hsrc4.assign((size_t)s->nnz_H * 4, -1);
    for (int64_t sl = 0; sl < s->nnz_H; ++sl) {
        const int32_t n = s->hptr[sl + 1] - s->hptr[sl];
        DTO_REQUIRE(n <= 4, "dto_shape_create: Hessian slot %lld has %d knot contributors (max 4)", (long long)sl, n);
        for (int32_t k = 0; k < n; ++k) s->hsrc4[4 * sl + k] = s->hsrc[s->hptr[sl] + k];
    }
    // match every knot's gather recipe against the compiled classes of the model library
    {
        const dto_model_vtable* vt = m->vt;
        bool all = vt->n_hg_classes > 0;
        for (int t = 0; t < T && all; ++t) {
            const int32_t s0 = s->knot[t].hslot, s1 = s->knot[t + 1].hslot;
            int found = -1;
            for (int c = 0; c < vt->n_hg_classes && found < 0; ++c) {
                if (vt->hg_nslots[c] != s1 - s0) continue;
                if (vt->hg_meta[3 * c] != hterm_cost[t] || vt->hg_meta[3 * c + 1] != hterm_dyn[t] ||
                    vt->hg_meta[3 * c + 2] != (t > 0 ? hterm_cost[t - 1] : 0))
                    continue;
                const int32_t* src = vt->hg_src + 4 * (size_t)vt->hg_ofs[c];
                bool same = true;
                for (int32_t sl = s0; sl < s1 && same; ++sl)
                    for (int k = 0; k < 4 && same; ++k) {
                        const int32_t id = s->hsrc4[4 * (size_t)sl + k];
                        int32_t enc = -1;
                        if (id >= 0) enc = id >= s->knot[t].hterm ? id - s->knot[t].hterm : -2 - (id - s->knot[t - 1].hterm);
                        same = enc == src[4 * (size_t)(sl - s0) + k];
                    }
                if (same) found = c;
            }
            if (found < 0) all = false;
            else s->knot[t].hclass = found;
        }
        if (!all)
            for (int t = 0; t < T; ++t) s->knot[t].hclass = -1;
        s->use_hclass = all;
    }
    s->gen_hslot.resize(gen_terms.size());
    for (size_t i = 0; i < gen_terms.size(); ++i) s->gen_hslot[i] = slot_of(gen_terms[i].row, gen_terms[i].col);
    {   // group by the knot owning the slot (hslot[] is monotone in t)
        s->gh_ptr.assign(T + 1, 0);
        std::vector<int> owner(gen_terms.size());
        for (size_t i = 0; i < gen_terms.size(); ++i) {
            int lo = 0, hi = T;  // last t with hslot[t] <= slot
            while (hi - lo > 1) {
                const int mid = (lo + hi) / 2;
                if (s->knot[mid].hslot <= s->gen_hslot[i]) lo = mid; else hi = mid;
            }
            owner[i] = lo;
            s->gh_ptr[lo + 1]++;
        }
        for (int t = 0; t < T; ++t) s->gh_ptr[t + 1] += s->gh_ptr[t];
        s->gh_ent.assign(gen_terms.size() * 2, 0);
        std::vector<int32_t> fill(s->gh_ptr.begin(), s->gh_ptr.end() - 1);
        for (size_t i = 0; i < gen_terms.size(); ++i) {
            const int32_t p = fill[owner[i]]++;
            s->gh_ent[2 * p] = s->gen_hslot[i];
            s->gh_ent[2 * p + 1] = (int32_t)i;
        }
        s->gh_rec.assign(gen_terms.size() * 8, 0);
        for (size_t p = 0; p < gen_terms.size(); ++p) {
            const int32_t i = s->gh_ent[2 * p + 1];
            const int32_t tmpl = gen->inst_tmpl[2][i];
            DTO_REQUIRE(tmpl >= 0 && tmpl < gen->n_hess_templates, "dto_shape_create: general Hessian template %d out of range", tmpl);
            int32_t* r = &s->gh_rec[8 * p];
            r[0] = s->gh_ent[2 * p];
            r[1] = tmpl;
            r[2] = gen->inst_zbase[2][i];
            r[3] = gen->inst_lbase[2][i];
            r[4] = gen->inst_wbase[2][i];
            r[5] = gen->hess_span[3 * tmpl + 0];
            r[6] = gen->hess_span[3 * tmpl + 2];
            r[7] = 0;
        }
    }
    if (gen) {
        const int n[3] = {gen->num_constraint, gen->nnz_jac, gen->has_hess ? gen->nnz_hess : 0};
        for (int cls = 0; cls < 3; ++cls) {
            s->gen_inst[cls].resize((size_t)n[cls] * 4);
            for (int i = 0; i < n[cls]; ++i) {
                s->gen_inst[cls][4 * i + 0] = gen->inst_tmpl[cls][i];
                s->gen_inst[cls][4 * i + 1] = gen->inst_zbase[cls][i];
                s->gen_inst[cls][4 * i + 2] = gen->inst_wbase[cls][i];
                s->gen_inst[cls][4 * i + 3] = gen->inst_lbase[cls][i];
            }
        }
    }

    // ---- shared-memory segment capacities: max flat extent of 32 consecutive (b,t) items
    {
        std::vector<int32_t> sz[6];
        for (int k = 0; k < 6; ++k) sz[k].resize(T);
        for (int t = 0; t < T; ++t) {
            sz[DTO_SEG_G][t] = s->nx[t] + s->nu[t];
            sz[DTO_SEG_CDYN][t] = dyn[t] ? dyn[t]->n_out : 0;
            sz[DTO_SEG_CSTAGE][t] = stage[t] ? stage[t]->n_out : 0;
            sz[DTO_SEG_JDYN][t] = dyn[t] ? dyn[t]->nnz_jac : 0;
            sz[DTO_SEG_JSTAGE][t] = stage[t] ? stage[t]->nnz_jac : 0;
            sz[DTO_SEG_HTERM][t] = hterm_cost[t] + hterm_dyn[t] + hterm_stage[t];
        }
        for (int k = 0; k < 6; ++k) {
            s->seg_cap[k] = cyclic_window_max(sz[k], 32);
            s->seg_pad[k] = 0;
        }
        {   // the compiled gather writes slot values over the term buffer: size it for both
            std::vector<int32_t> nslot(T);
            for (int t = 0; t < T; ++t) nslot[t] = s->knot[t + 1].hslot - s->knot[t].hslot;
            s->hslot_cap = cyclic_window_max(nslot, 32);
            s->seg_cap[DTO_SEG_HTERM] = std::max(s->seg_cap[DTO_SEG_HTERM], s->hslot_cap);
        }
        if (m->vt->hess_halo) {
            s->seg_pad[DTO_SEG_JDYN] = *std::max_element(sz[DTO_SEG_JDYN].begin(), sz[DTO_SEG_JDYN].end());
            s->seg_pad[DTO_SEG_HTERM] = *std::max_element(sz[DTO_SEG_HTERM].begin(), sz[DTO_SEG_HTERM].end());
        }
    }
    // ---- persistent pipeline kernel: input staging capacities, maximised over every tile start.
    // A tile is at most 32 consecutive (b,t) items (31 own + halo, or 32 own); its z / sigma / w inputs
    // are one flat range each, its dynamics / stage multipliers one range per problem touched.
    {
        std::vector<int32_t> nw(T);
        bool mono = true;
        for (int t = 0; t < T; ++t) {
            nw[t] = s->knot[t].pad0;
            if (t > 0 && (s->knot[t].wofs < s->knot[t - 1].wofs || s->knot[t].wofs + nw[t] < s->knot[t - 1].wofs + nw[t - 1])) mono = false;
        }
        s->w_flat = mono && s->N_w > 0;
        int64_t cap[5] = {0, 0, 0, 0, 0};
        int nsub_max = 1;
        for (int tf = 0; tf < T; ++tf) {
            int64_t zl = 0, ld = 0, ls = 0;
            int nprob = 1, t = tf, tl = tf;
            for (int i = 0; i < 32; ++i) {
                zl += s->nx[t] + s->nu[t];
                ld += s->knot[t + 1].rdyn - s->knot[t].rdyn;
                ls += s->knot[t + 1].rstage - s->knot[t].rstage;
                tl = t;
                if (++t == T) { t = 0; if (i < 31) ++nprob; }
            }
            if (tl + 1 < T) zl += s->nx[tl + 1];
            const int64_t wl = (int64_t)(nprob - 1) * s->N_w + s->knot[tl].wofs + nw[tl] - s->knot[tf].wofs;
            cap[DTO_IN_Z] = std::max(cap[DTO_IN_Z], zl + 2);
            cap[DTO_IN_SIGMA] = std::max<int64_t>(cap[DTO_IN_SIGMA], nprob + 2);
            cap[DTO_IN_W] = std::max(cap[DTO_IN_W], s->w_flat ? wl + 2 : 0);
            cap[DTO_IN_LDYN] = std::max(cap[DTO_IN_LDYN], ld + 3 * nprob + 2);
            cap[DTO_IN_LSTAGE] = std::max(cap[DTO_IN_LSTAGE], ls + 3 * nprob + 2);
            nsub_max = std::max(nsub_max, nprob);
        }
        for (int k = 0; k < 5; ++k) s->in_cap[k] = (int32_t)((cap[k] + 1) & ~(int64_t)1);
        s->nsub_max = nsub_max;
        s->persist_ok = nsub_max <= 8;
    }
    {
        static std::atomic<int64_t> next_shape_id{1};
        s->shape_id = next_shape_id.fetch_add(1);
    }
    guard.s = nullptr;
    *out = s;
    return DTO_OK;
}

extern "C" void dto_shape_destroy(dto_shape* s) { delete s; }
extern "C" int64_t dto_num_variables(const dto_shape* s) { return s ? s->N_z : -1; }
extern "C" int64_t dto_num_constraint(const dto_shape* s) { return s ? s->N_c : -1; }
extern "C" int64_t dto_num_jacobian(const dto_shape* s) { return s ? s->nnz_J : -1; }
extern "C" int64_t dto_num_hessian(const dto_shape* s) { return s ? s->nnz_H : -1; }
extern "C" int64_t dto_num_hessian_nonunique(const dto_shape* s) { return s ? s->nnz_H_nonunique : -1; }
extern "C" int64_t dto_num_parameter(const dto_shape* s) { return s ? s->N_w : -1; }
extern "C" int dto_hessian_available(const dto_shape* s) { return (s && s->hessian_available) ? 1 : 0; }

extern "C" int dto_jacobian_structure(const dto_shape* s, int64_t* rows, int64_t* cols)
{
    if (!s || !rows || !cols) return fail(DTO_ERR_BAD_ARG, "dto_jacobian_structure: null argument");
    std::copy(s->jac_row.begin(), s->jac_row.end(), rows);
    std::copy(s->jac_col.begin(), s->jac_col.end(), cols);
    return DTO_OK;
}
extern "C" int dto_hessian_lagrangian_structure(const dto_shape* s, int64_t* rows, int64_t* cols)
{
    if (!s || !rows || !cols) return fail(DTO_ERR_BAD_ARG, "dto_hessian_lagrangian_structure: null argument");
    std::copy(s->hess_row.begin(), s->hess_row.end(), rows);
    std::copy(s->hess_col.begin(), s->hess_col.end(), cols);
    return DTO_OK;
}
extern "C" int dto_constraint_bounds(const dto_shape* s, double* lower, double* upper)
{
    if (!s || !lower || !upper) return fail(DTO_ERR_BAD_ARG, "dto_constraint_bounds: null argument");
    std::copy(s->c_lower.begin(), s->c_lower.end(), lower);
    std::copy(s->c_upper.begin(), s->c_upper.end(), upper);
    return DTO_OK;
}
extern "C" int dto_knot_layout(const dto_shape* s, int64_t* state_start, int32_t* num_state, int64_t* action_start, int32_t* num_action)
{
    if (!s) return fail(DTO_ERR_BAD_ARG, "dto_knot_layout: null shape");
    for (int t = 0; t < s->T; ++t) {
        if (state_start) state_start[t] = s->knot[t].zofs + 1;
        if (num_state) num_state[t] = s->nx[t];
        if (action_start) action_start[t] = s->knot[t].zofs + s->nx[t] + 1;
        if (num_action) num_action[t] = s->nu[t];
    }
    return DTO_OK;
}
extern "C" int64_t dto_algorithmic_bytes_per_problem(const dto_shape* s)
{
    return s ? 8 * (s->N_z + s->N_c + s->N_w + s->nnz_J + s->nnz_H) + 8 : -1;
}

static void fill_args(const dto_shape* s, const dto_shard* sh, dto_launch_args* a)
{
    memset(a, 0, sizeof(*a));
    a->B = sh ? sh->size : 0;
    a->T = s->T;
    a->N_z = (int32_t)s->N_z;
    a->N_c = (int32_t)s->N_c;
    a->N_w = (int32_t)s->N_w;
    a->nnz_J = (int32_t)s->nnz_J;
    a->nnz_H = (int32_t)s->nnz_H;
    if (sh) {
        a->z = sh->arr[DTO_ARRAY_Z];
        a->lam = sh->arr[DTO_ARRAY_LAMBDA];
        a->sigma = sh->arr[DTO_ARRAY_SIGMA];
        a->w = sh->arr[DTO_ARRAY_W];
        a->f = sh->arr[DTO_ARRAY_F];
        a->g = sh->arr[DTO_ARRAY_G];
        a->c = sh->arr[DTO_ARRAY_C];
        a->J = sh->arr[DTO_ARRAY_J];
        a->H = sh->arr[DTO_ARRAY_H];
        a->knot = sh->d_knot;
        a->hsrc4 = sh->d_hsrc4;
        for (int k = 0; k < 3; ++k) a->gen_inst[k] = sh->d_gen_inst[k];
        a->gen_hslot = sh->d_gen_hslot;
        a->gh_ptr = sh->d_gh_ptr;
        a->gh_ent = sh->d_gh_ent;
        a->gh_rec = sh->d_gh_rec;
    }
    a->gen_nrow = (int32_t)s->n_gen_rows;
    a->gen_njac = (int32_t)s->n_gen_jac;
    a->gen_nhess = (int32_t)s->gen_hslot.size();
    a->gen_row0 = (int32_t)(s->n_dyn_rows + s->n_stage_rows);
    a->gen_jac0 = (int32_t)(s->n_dyn_jac + s->n_stage_jac);
    for (int k = 0; k < 6; ++k) {
        a->seg_cap[k] = s->seg_cap[k];
        a->seg_pad[k] = s->seg_pad[k];
    }
    a->use_hclass = s->use_hclass ? 1 : 0;
    for (int k = 0; k < 5; ++k) a->in_cap[k] = s->in_cap[k];
    a->nsub_max = s->nsub_max;
    a->persist_ok = s->persist_ok ? 1 : 0;
    a->w_flat = s->w_flat ? 1 : 0;
    a->hslot_cap = s->hslot_cap;
    a->shape_id = s->shape_id;
    {   // magic number for g / T, exact for g < 2^31 (k = 31 + ceil(log2 T), M = ceil(2^k / T) < 2^32)
        int lg = 0;
        while ((1ll << lg) < s->T) ++lg;
        const int k = 31 + lg;
        a->div_shift = k;
        a->div_mul = (uint32_t)((((unsigned __int128)1 << k) + (unsigned)s->T - 1) / (unsigned)s->T);
        a->z_per_knot = (int32_t)(s->N_z / s->T);
        a->c_per_knot = (int32_t)(s->N_c / s->T);
    }
}

extern "C" int dto_shape_compiled_gather(const dto_shape* s) { return (s && s->use_hclass) ? 1 : 0; }

extern "C" int64_t dto_kernel_smem_bytes(const dto_shape* s, int kernel_id)
{
    if (!s) return -1;
    dto_launch_args a;
    fill_args(s, nullptr, &a);
    return s->model->vt->smem_bytes(kernel_id, &a);
}

// ------------------------------------------------------------------------------------------
// batch
// ------------------------------------------------------------------------------------------
template <class V>
static int upload(V** dst, const std::vector<V>& src, cudaStream_t st)
{
    *dst = nullptr;
    if (src.empty()) return DTO_OK;
    DTO_CUDA(cudaMalloc((void**)dst, src.size() * sizeof(V)));
    DTO_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(V), cudaMemcpyHostToDevice, st));
    return DTO_OK;
}

static int ensure_array(dto_batch* b, dto_shard& sh, int array)
{
    if (sh.arr[array]) return DTO_OK;
    const int64_t n = std::max<int64_t>(1, array_width(b->shape, array) * sh.size);
    DTO_CUDA(cudaSetDevice(sh.device));
    // +2 doubles: the persistent kernel's 16-byte bulk copies may touch one element past an odd-length array
    DTO_CUDA(cudaMalloc((void**)&sh.arr[array], (size_t)(n + 2) * sizeof(double)));
    if (array == DTO_ARRAY_SIGMA) {
        std::vector<double> ones((size_t)n, 1.0);
        DTO_CUDA(cudaMemcpy(sh.arr[array], ones.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        DTO_CUDA(cudaMemsetAsync(sh.arr[array], 0, (size_t)n * sizeof(double), sh.stream));
    }
    return DTO_OK;
}

extern "C" void dto_batch_destroy(dto_batch* b)
{
    if (!b) return;
    for (dto_shard& sh : b->shards) {
        cudaSetDevice(sh.device);
        if (sh.stream) cudaStreamSynchronize(sh.stream);
        for (double*& p : sh.arr)
            if (p) cudaFree(p);
        if (sh.d_knot) cudaFree(sh.d_knot);
        if (sh.d_hsrc4) cudaFree(sh.d_hsrc4);
        for (int32_t*& p : sh.d_gen_inst)
            if (p) cudaFree(p);
        if (sh.d_gen_hslot) cudaFree(sh.d_gen_hslot);
        if (sh.d_gh_ptr) cudaFree(sh.d_gh_ptr);
        if (sh.d_gh_ent) cudaFree(sh.d_gh_ent);
        if (sh.d_gh_rec) cudaFree(sh.d_gh_rec);
        if (sh.stream && sh.own_stream) cudaStreamDestroy(sh.stream);
        for (cudaStream_t& ps : sh.pipe)
            if (ps) cudaStreamDestroy(ps);
    }
    cudaGetLastError();
    delete b;
}

extern "C" int dto_batch_create(dto_shape* s, int64_t B, const int* devices, int ndev, dto_batch** out)
{
    if (!s || !out) return fail(DTO_ERR_BAD_ARG, "dto_batch_create: null argument");
    *out = nullptr;
    DTO_REQUIRE(B >= 1, "dto_batch_create: B=%lld", (long long)B);
    DTO_REQUIRE(ndev >= 0 && (ndev == 0 || devices), "dto_batch_create: bad device list");
    int count = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) {
            cudaGetLastError();
            return fail(DTO_ERR_CUDA, "dto_batch_create: no CUDA device available (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        }
    }
    std::vector<int> devs;
    if (ndev == 0) {
        int cur = 0;
        DTO_CUDA(cudaGetDevice(&cur));
        devs.push_back(cur);
    } else {
        devs.assign(devices, devices + ndev);
    }
    for (int dv : devs) DTO_REQUIRE(dv >= 0 && dv < count, "dto_batch_create: device %d not in [0,%d)", dv, count);
    DTO_REQUIRE(B * std::max<int64_t>(s->nnz_J, s->nnz_H) < (1ll << 40), "dto_batch_create: batch too large");
    {
        const int64_t nsh = ndev == 0 ? 1 : ndev;
        DTO_REQUIRE(((B + nsh - 1) / nsh) * (int64_t)s->T < (1ll << 31) - 64,
                    "dto_batch_create: %lld problems x %d knots per shard exceed 2^31 items; use more shards", (long long)((B + nsh - 1) / nsh), s->T);
    }

    dto_batch* b = new (std::nothrow) dto_batch();
    if (!b) return fail(DTO_ERR_OOM, "dto_batch_create: out of host memory");
    b->shape = s;
    b->B = B;
    const int64_t chunk = (B + (int64_t)devs.size() - 1) / (int64_t)devs.size();
    for (size_t i = 0; i < devs.size(); ++i) {
        dto_shard sh;
        sh.device = devs[i];
        sh.begin = std::min<int64_t>(B, (int64_t)i * chunk);
        sh.size = std::min<int64_t>(chunk, B - sh.begin);
        b->shards.push_back(sh);
    }
    int rc = DTO_OK;
    for (dto_shard& sh : b->shards) {
        auto setup = [&]() -> int {
            DTO_CUDA(cudaSetDevice(sh.device));
            DTO_CUDA(cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking));
            int r;
            if ((r = upload(&sh.d_knot, s->knot, sh.stream))) return r;
            if ((r = upload(&sh.d_hsrc4, s->hsrc4, sh.stream))) return r;
            for (int k = 0; k < 3; ++k)
                if ((r = upload(&sh.d_gen_inst[k], s->gen_inst[k], sh.stream))) return r;
            if ((r = upload(&sh.d_gen_hslot, s->gen_hslot, sh.stream))) return r;
            if ((r = upload(&sh.d_gh_ptr, s->gh_ptr, sh.stream))) return r;
            if ((r = upload(&sh.d_gh_ent, s->gh_ent, sh.stream))) return r;
            if ((r = upload(&sh.d_gh_rec, s->gh_rec, sh.stream))) return r;
            // inputs are allocated eagerly, outputs on first use
            for (int arr : {DTO_ARRAY_Z, DTO_ARRAY_LAMBDA, DTO_ARRAY_SIGMA, DTO_ARRAY_W})
                if ((r = ensure_array(b, sh, arr))) return r;
            DTO_CUDA(cudaStreamSynchronize(sh.stream));
            return DTO_OK;
        };
        if ((rc = setup())) break;
    }
    if (rc) {
        std::string msg = g_err;
        dto_batch_destroy(b);
        snprintf(g_err, sizeof(g_err), "%s", msg.c_str());
        return rc;
    }
    *out = b;
    return DTO_OK;
}

extern "C" int64_t dto_batch_size(const dto_batch* b) { return b ? b->B : -1; }
extern "C" int dto_batch_num_shards(const dto_batch* b) { return b ? (int)b->shards.size() : -1; }
extern "C" int64_t dto_shard_begin(const dto_batch* b, int shard)
{
    return (b && shard >= 0 && shard < (int)b->shards.size()) ? b->shards[shard].begin : -1;
}
extern "C" int64_t dto_shard_size(const dto_batch* b, int shard)
{
    return (b && shard >= 0 && shard < (int)b->shards.size()) ? b->shards[shard].size : -1;
}
extern "C" int dto_shard_device(const dto_batch* b, int shard)
{
    return (b && shard >= 0 && shard < (int)b->shards.size()) ? b->shards[shard].device : -1;
}

static int sync_all(dto_batch* b)
{
    for (dto_shard& sh : b->shards) {
        DTO_CUDA(cudaSetDevice(sh.device));
        DTO_CUDA(cudaStreamSynchronize(sh.stream));
    }
    return DTO_OK;
}

static int h2d(dto_batch* b, int array, const double* host)
{
    const int64_t wdt = array_width(b->shape, array);
    if (wdt == 0) return DTO_OK;
    if (!host) return fail(DTO_ERR_BAD_ARG, "null host pointer for input array %d", array);
    for (dto_shard& sh : b->shards) {
        if (sh.size == 0) continue;
        int r;
        if ((r = ensure_array(b, sh, array))) return r;
        DTO_CUDA(cudaSetDevice(sh.device));
        DTO_CUDA(cudaMemcpyAsync(sh.arr[array], host + sh.begin * wdt, (size_t)(sh.size * wdt) * sizeof(double), cudaMemcpyHostToDevice,
                                 sh.stream));
    }
    return DTO_OK;
}

static int d2h(dto_batch* b, int array, double* host)
{
    const int64_t wdt = array_width(b->shape, array);
    if (wdt == 0 || !host) return DTO_OK;
    for (dto_shard& sh : b->shards) {
        if (sh.size == 0) continue;
        DTO_CUDA(cudaSetDevice(sh.device));
        DTO_CUDA(cudaMemcpyAsync(host + sh.begin * wdt, sh.arr[array], (size_t)(sh.size * wdt) * sizeof(double), cudaMemcpyDeviceToHost,
                                 sh.stream));
    }
    return DTO_OK;
}

extern "C" int dto_set_parameters(dto_batch* b, const double* w)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_set_parameters: null batch");
    int r;
    if ((r = h2d(b, DTO_ARRAY_W, w))) return r;
    return sync_all(b);
}
extern "C" int dto_set_x(dto_batch* b, const double* z)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_set_x: null batch");
    int r;
    if ((r = h2d(b, DTO_ARRAY_Z, z))) return r;
    if ((r = sync_all(b))) return r;
    b->have_x = true;
    return DTO_OK;
}
extern "C" int dto_set_duals(dto_batch* b, const double* sigma, const double* lambda)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_set_duals: null batch");
    int r;
    if ((r = h2d(b, DTO_ARRAY_SIGMA, sigma))) return r;
    if ((r = h2d(b, DTO_ARRAY_LAMBDA, lambda))) return r;
    if ((r = sync_all(b))) return r;
    b->have_duals = true;
    return DTO_OK;
}

static int launch_all(dto_batch* b, int kernel_id)
{
    static const char* const names[] = {"dto:objective", "dto:gradient", "dto:constraint", "dto:jacobian", "dto:hessian", "dto:jacobian+hessian"};
    DtoRange range(kernel_id >= 0 && kernel_id < 6 ? names[kernel_id] : "dto:launch");
    const dto_shape* s = b->shape;
    if (kernel_id < 0 || kernel_id >= DTO_K_COUNT) return fail(DTO_ERR_BAD_ARG, "dto_launch: kernel id %d", kernel_id);
    if (!b->have_x) return fail(DTO_ERR_STATE, "evaluation requested before dto_set_x");
    const bool hess = kernel_id == DTO_K_HESSIAN || kernel_id == DTO_K_JAC_HESS;
    if (hess && !s->hessian_available)
        return fail(DTO_ERR_NO_HESSIAN, "Hessian requested but a Cost was built without evaluate_hessian (reference throws at src/costs.jl:68)");
    if (hess && !b->have_duals) return fail(DTO_ERR_STATE, "Hessian requested before dto_set_duals");
    for (dto_shard& sh : b->shards) {
        if (sh.size == 0) continue;
        int r;
        std::vector<int> outs;
        switch (kernel_id) {
        case DTO_K_OBJECTIVE: outs = {DTO_ARRAY_F}; break;
        case DTO_K_GRADIENT: outs = {DTO_ARRAY_G}; break;
        case DTO_K_CONSTRAINT: outs = {DTO_ARRAY_C}; break;
        case DTO_K_JACOBIAN: outs = {DTO_ARRAY_J}; break;
        case DTO_K_HESSIAN: outs = {DTO_ARRAY_H}; break;
        default: outs = {DTO_ARRAY_J, DTO_ARRAY_H}; break;
        }
        for (int o : outs)
            if ((r = ensure_array(b, sh, o))) return r;
        DTO_CUDA(cudaSetDevice(sh.device));
        dto_launch_args a;
        fill_args(s, &sh, &a);
        const int e = s->model->vt->launch(kernel_id, &a, (void*)sh.stream);
        if (e < 0) return fail(DTO_ERR_CUDA, "kernel launch (id %d) failed: %s", kernel_id, cudaGetErrorString((cudaError_t)(-e)));
        const int64_t n = e;
        b->launches += n;
    }
    return DTO_OK;
}

static int eval_to_host(dto_batch* b, int kernel_id, int array, double* host)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "null batch");
    if (!host && array_width(b->shape, array) > 0) return fail(DTO_ERR_BAD_ARG, "null output pointer");
    DtoRange range("dto:eval_to_host");
    int r;
    if ((r = launch_all(b, kernel_id))) return r;
    if ((r = d2h(b, array, host))) return r;
    return sync_all(b);
}

extern "C" int dto_eval_objective(dto_batch* b, double* f) { return eval_to_host(b, DTO_K_OBJECTIVE, DTO_ARRAY_F, f); }
extern "C" int dto_eval_objective_gradient(dto_batch* b, double* g) { return eval_to_host(b, DTO_K_GRADIENT, DTO_ARRAY_G, g); }
extern "C" int dto_eval_constraint(dto_batch* b, double* c) { return eval_to_host(b, DTO_K_CONSTRAINT, DTO_ARRAY_C, c); }
extern "C" int dto_eval_constraint_jacobian(dto_batch* b, double* J) { return eval_to_host(b, DTO_K_JACOBIAN, DTO_ARRAY_J, J); }
extern "C" int dto_eval_hessian_lagrangian(dto_batch* b, double* H) { return eval_to_host(b, DTO_K_HESSIAN, DTO_ARRAY_H, H); }
extern "C" int dto_eval_jacobian_hessian(dto_batch* b, double* J, double* H)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_eval_jacobian_hessian: null batch");
    int r;
    if ((r = launch_all(b, DTO_K_JAC_HESS))) return r;
    if ((r = d2h(b, DTO_ARRAY_J, J))) return r;
    if ((r = d2h(b, DTO_ARRAY_H, H))) return r;
    return sync_all(b);
}

extern "C" int dto_eval_jacobian_hessian_host(dto_batch* b, const double* z, const double* sigma, const double* lambda, double* J,
                                              double* H, int nchunks)
{
    if (!b || !z || !sigma || !lambda) return fail(DTO_ERR_BAD_ARG, "dto_eval_jacobian_hessian_host: null argument");
    DtoRange range("dto:eval_jacobian_hessian_host (chunked H2D | kernel | D2H pipeline)");
    const dto_shape* s = b->shape;
    if (!s->hessian_available)
        return fail(DTO_ERR_NO_HESSIAN, "Hessian requested but a Cost was built without evaluate_hessian (reference throws at src/costs.jl:68)");
    if (nchunks <= 0) nchunks = 8;
    const int64_t wz = s->N_z, wl = s->N_c, wj = s->nnz_J, wh = s->nnz_H;
    for (dto_shard& sh : b->shards) {
        if (sh.size == 0) continue;
        int r;
        for (int arr : {DTO_ARRAY_Z, DTO_ARRAY_LAMBDA, DTO_ARRAY_SIGMA, DTO_ARRAY_W, DTO_ARRAY_J, DTO_ARRAY_H})
            if ((r = ensure_array(b, sh, arr))) return r;
        DTO_CUDA(cudaSetDevice(sh.device));
        DTO_CUDA(cudaStreamSynchronize(sh.stream));  // parameters / earlier work on the shard's main stream
        for (cudaStream_t& ps : sh.pipe)
            if (!ps) DTO_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
        const int64_t nc = std::min<int64_t>(nchunks, sh.size);
        const int64_t per = (sh.size + nc - 1) / nc;
        for (int64_t k = 0; k < nc; ++k) {
            const int64_t first = k * per, cnt = std::min<int64_t>(per, sh.size - first);
            if (cnt <= 0) break;
            cudaStream_t st = sh.pipe[k % 4];
            const int64_t g = sh.begin + first;  // global problem index
            DTO_CUDA(cudaMemcpyAsync(sh.arr[DTO_ARRAY_Z] + first * wz, z + g * wz, (size_t)(cnt * wz) * sizeof(double), cudaMemcpyHostToDevice, st));
            DTO_CUDA(cudaMemcpyAsync(sh.arr[DTO_ARRAY_LAMBDA] + first * wl, lambda + g * wl, (size_t)(cnt * wl) * sizeof(double), cudaMemcpyHostToDevice, st));
            DTO_CUDA(cudaMemcpyAsync(sh.arr[DTO_ARRAY_SIGMA] + first, sigma + g, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, st));
            dto_launch_args a;
            fill_args(s, &sh, &a);
            a.B = cnt;
            a.z += first * wz;
            a.lam += first * wl;
            a.sigma += first;
            a.w += first * s->N_w;
            a.J += first * wj;
            a.H += first * wh;
            const int e = s->model->vt->launch(DTO_K_JAC_HESS, &a, (void*)st);
            if (e < 0) return fail(DTO_ERR_CUDA, "kernel launch (fused, chunk %lld) failed: %s", (long long)k, cudaGetErrorString((cudaError_t)(-e)));
            b->launches += e;
            if (J) DTO_CUDA(cudaMemcpyAsync(J + g * wj, sh.arr[DTO_ARRAY_J] + first * wj, (size_t)(cnt * wj) * sizeof(double), cudaMemcpyDeviceToHost, st));
            if (H) DTO_CUDA(cudaMemcpyAsync(H + g * wh, sh.arr[DTO_ARRAY_H] + first * wh, (size_t)(cnt * wh) * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
    }
    for (dto_shard& sh : b->shards) {
        if (sh.size == 0) continue;
        DTO_CUDA(cudaSetDevice(sh.device));
        for (cudaStream_t ps : sh.pipe)
            if (ps) DTO_CUDA(cudaStreamSynchronize(ps));
    }
    b->have_x = b->have_duals = true;
    return DTO_OK;
}

extern "C" int dto_get_problem(dto_batch* b, int array, int64_t problem, double* out)
{
    if (!b || !out) return fail(DTO_ERR_BAD_ARG, "dto_get_problem: null argument");
    const int64_t wdt = array_width(b->shape, array);
    DTO_REQUIRE(wdt >= 0, "dto_get_problem: array id %d", array);
    DTO_REQUIRE(problem >= 0 && problem < b->B, "dto_get_problem: problem %lld not in [0,%lld)", (long long)problem, (long long)b->B);
    if (wdt == 0) return DTO_OK;
    for (dto_shard& sh : b->shards) {
        if (problem < sh.begin || problem >= sh.begin + sh.size) continue;
        if (!sh.arr[array]) return fail(DTO_ERR_STATE, "dto_get_problem: array %d has not been computed yet", array);
        DTO_CUDA(cudaSetDevice(sh.device));
        DTO_CUDA(cudaMemcpyAsync(out, sh.arr[array] + (problem - sh.begin) * wdt, (size_t)wdt * sizeof(double), cudaMemcpyDeviceToHost,
                                 sh.stream));
        DTO_CUDA(cudaStreamSynchronize(sh.stream));
        return DTO_OK;
    }
    return fail(DTO_ERR_STATE, "dto_get_problem: problem not found in any shard");
}

extern "C" int dto_get_last_x(const dto_batch* b, int64_t problem, double* z)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_get_last_x: null batch");
    if (!b->have_x) return fail(DTO_ERR_STATE, "dto_get_last_x: no z has been set");
    return dto_get_problem(const_cast<dto_batch*>(b), DTO_ARRAY_Z, problem, z);
}

extern "C" void* dto_device_pointer(dto_batch* b, int array, int shard)
{
    if (!b || shard < 0 || shard >= (int)b->shards.size() || array < 0 || array > DTO_ARRAY_H) {
        fail(DTO_ERR_BAD_ARG, "dto_device_pointer: bad argument");
        return nullptr;
    }
    if (ensure_array(b, b->shards[shard], array)) return nullptr;
    if (array == DTO_ARRAY_Z) b->have_x = true;  // caller writes the inputs on the device
    if (array == DTO_ARRAY_LAMBDA || array == DTO_ARRAY_SIGMA) b->have_duals = true;
    return b->shards[shard].arr[array];
}
extern "C" void* dto_get_stream(dto_batch* b, int shard)
{
    return (b && shard >= 0 && shard < (int)b->shards.size()) ? (void*)b->shards[shard].stream : nullptr;
}
extern "C" int dto_set_stream(dto_batch* b, int shard, void* cuda_stream)
{
    if (!b || shard < 0 || shard >= (int)b->shards.size()) return fail(DTO_ERR_BAD_ARG, "dto_set_stream: bad argument");
    dto_shard& sh = b->shards[shard];
    DTO_CUDA(cudaSetDevice(sh.device));
    DTO_CUDA(cudaStreamSynchronize(sh.stream));
    if (sh.own_stream && sh.stream) cudaStreamDestroy(sh.stream);
    sh.stream = (cudaStream_t)cuda_stream;
    sh.own_stream = false;
    return DTO_OK;
}
extern "C" int dto_launch(dto_batch* b, int kernel_id)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_launch: null batch");
    return launch_all(b, kernel_id);
}
extern "C" int dto_sync(dto_batch* b)
{
    if (!b) return fail(DTO_ERR_BAD_ARG, "dto_sync: null batch");
    return sync_all(b);
}
extern "C" int64_t dto_launch_count(const dto_batch* b) { return b ? b->launches : -1; }

// ------------------------------------------------------------------------------------------
// device-side KKT consumer
// ------------------------------------------------------------------------------------------
#include "dto_kkt_host.inc"

// ------------------------------------------------------------------------------------------
// batched Newton-KKT solver: the caller of the callback path
// ------------------------------------------------------------------------------------------
#include "dto_sqp_host.inc"
