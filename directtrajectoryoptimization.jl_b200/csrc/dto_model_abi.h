/* Internal ABI between libdto.so (host runtime, C ABI of include/dto.h) and a generated
 * model library (element device functions + the hand-written kernels of dto_kernels.cuh
 * instantiated on them, compiled by nvcc for sm_100a).
 *
 * A model library is self-describing: it carries the dims and LOCAL sparsity patterns of
 * every element kind (what the reference keeps in the element structs,
 * /root/reference/src/dynamics.jl:1-16, src/costs.jl:1-11, src/constraints.jl:1-17,
 * src/general_constraint.jl:1-16), so the runtime can assemble the global structures
 * (/root/reference/src/data.jl:150-220) without any symbolic machinery.
 */
#ifndef DTO_MODEL_ABI_H
#define DTO_MODEL_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTO_MODEL_ABI_VERSION 14

/* One element kind (a distinct Dynamics / Cost / Constraint object of the reference). All
 * patterns are 1-based local indices in the element's own variable order
 * ([x;u;y] dynamics, [x;u] cost/stage), stored in CSC order = value-slot order. */
typedef struct dto_element_desc {
    int32_t n_out;        /* dynamics: num_next_state; stage: num_constraint; cost: 1 */
    int32_t nx, nu, nw;   /* num_state, num_action, num_parameter */
    int32_t nnz_jac;      /* dynamics/stage: num_jacobian; cost: nx+nu (dense gradient) */
    const int32_t* jac_row;
    const int32_t* jac_col;
    int32_t has_hess;     /* element built with evaluate_hessian=true */
    int32_t nnz_hess;
    const int32_t* hess_row;
    const int32_t* hess_col;
    int32_t n_ineq;       /* stage: indices_inequality (1-based local rows) */
    const int32_t* ineq;
} dto_element_desc;

/* The (single) GeneralConstraint block over the whole z. Outputs are evaluated as
 * INSTANCES of de-duplicated expression templates: instance i of output class k
 * (k = 0 residual rows, 1 Jacobian nonzeros, 2 Hessian nonzeros) evaluates template
 * tmpl[i] with its z / w / lambda indices shifted by zbase/wbase/lbase. */
typedef struct dto_general_desc {
    int32_t num_variables, num_parameter, num_constraint;
    int32_t nnz_jac;
    const int32_t* jac_row; /* 1-based local row */
    const int32_t* jac_col; /* 1-based global column */
    int32_t has_hess;
    int32_t nnz_hess;
    const int32_t* hess_row;
    const int32_t* hess_col;
    int32_t n_ineq;
    const int32_t* ineq;
    /* instance tables, one per output class, length = num_constraint / nnz_jac / nnz_hess */
    const int32_t* inst_tmpl[3];
    const int32_t* inst_zbase[3];
    const int32_t* inst_wbase[3];
    const int32_t* inst_lbase[3];
    /* per Hessian template (class 2): how many consecutive z / w / lambda entries from its base it may read */
    int32_t n_hess_templates;
    const int32_t* hess_span; /* [n_hess_templates][3] = zspan, wspan, lspan */
} dto_general_desc;

/* Per-knot static table entry (device). Entry t describes knot t (0-based); the table
 * has T+1 entries, entry T closing every prefix sum. */
typedef struct dto_knot_entry {
    int32_t zofs;    /* offset of x_t in z (entry T: num_variables)                       */
    int32_t nx;      /* num_state at knot t: u_t lives at zofs+nx                          */
    int32_t wofs;    /* offset of w_t in the problem's parameter vector                    */
    int32_t kdyn;    /* dynamics kind of knot t (-1: none, i.e. the last knot)             */
    int32_t kcost;   /* cost kind                                                          */
    int32_t kstage;  /* stage-constraint kind (-1: empty Constraint())                     */
    int32_t rdyn;    /* first dynamics constraint row (0-based, absolute in c)             */
    int32_t rstage;  /* first stage constraint row (absolute in c)                         */
    int32_t jdyn;    /* first dynamics Jacobian slot (absolute in J)                       */
    int32_t jstage;  /* first stage Jacobian slot (absolute in J)                          */
    int32_t hterm;   /* first Hessian TERM of knot t; per knot: [cost][dynamics][stage]    */
    int32_t hslot;   /* first Hessian slot whose row is a variable of knot t               */
    int32_t hclass;  /* compiled gather-recipe class of knot t (-1: use the table gather)  */
    int32_t hprev;   /* hterm[t] - hterm[t-1] (0 for t = 0)                                */
    int32_t pad0;    /* length of w_t (parameter_dim[t])                                   */
    int32_t pad1;
} dto_knot_entry;    /* 16 x int32 = 64 bytes */

/* Kernel ids for dto_model_vtable.launch */
enum {
    DTO_K_OBJECTIVE = 0,  /* f[B]                               (src/moi.jl:1-13)   */
    DTO_K_GRADIENT = 1,   /* g[B][N_z]                          (src/moi.jl:15-30)  */
    DTO_K_CONSTRAINT = 2, /* c[B][N_c]                          (src/moi.jl:32-50)  */
    DTO_K_JACOBIAN = 3,   /* J[B][nnz_J]                        (src/moi.jl:52-70)  */
    DTO_K_HESSIAN = 4,    /* H[B][nnz_H]                        (src/moi.jl:72-120) */
    DTO_K_JAC_HESS = 5,   /* J and H in one pass (the benchmark unit)               */
    DTO_K_COUNT = 6
};

/* Everything a launch needs; all pointers are device pointers of ONE shard. */
typedef struct dto_launch_args {
    int64_t B;       /* problems in this shard */
    int32_t T;
    int32_t N_z, N_c, N_w, nnz_J, nnz_H;
    const double* z;      /* [B][N_z]   */
    const double* lam;    /* [B][N_c]   */
    const double* sigma;  /* [B]        */
    const double* w;      /* [B][N_w]   */
    double* f;            /* [B]        */
    double* g;            /* [B][N_z]   */
    double* c;            /* [B][N_c]   */
    double* J;            /* [B][nnz_J] */
    double* H;            /* [B][nnz_H] */
    const dto_knot_entry* knot; /* [T+1] */
    /* [nnz_H][4] per Hessian slot: the (at most four: cost_t, dynamics_{t-1}, dynamics_t, stage_t)
     * contributing term ids in the reference's += order (Q5), -1 = none. One 16-byte record per
     * slot: the gather issues a single coalesced, independent load per output value. */
    const int32_t* hsrc4;
    /* general constraint (device copies of the instance tables + output placement) */
    int32_t gen_nrow, gen_njac, gen_nhess;
    int32_t gen_row0;     /* first general row in c / lambda        */
    int32_t gen_jac0;     /* first general slot in J                */
    const int32_t* gen_inst[3]; /* per class: [n][4] = tmpl, zbase, wbase, lbase */
    const int32_t* gen_hslot;   /* [gen_nhess] Hessian slot of each general nonzero */
    /* shared-memory capacities (doubles per warp) computed by the runtime for the shape:
     * index = segment id below */
    int32_t seg_cap[6];
    int32_t seg_pad[6];   /* halo pad (doubles) in front of a segment, 0 unless halo */
    int32_t use_hclass;   /* 1: every knot matched a compiled gather class of the model library */
    /* general-constraint Hessian entries grouped by the knot that owns their slot (fused into the
     * knot kernel when use_hclass): entries gh_ptr[t]..gh_ptr[t+1], each = (slot, instance) */
    const int32_t* gh_ptr;  /* [T+1] */
    const int32_t* gh_ent;  /* [gen_nhess][2] */
    /* the same entries flattened for the ws kernel's tile plans, 2 x int4 per entry in gh_ent order:
     * {slot, template, zbase, lbase}, {wbase, zspan, lspan, 0} */
    const int32_t* gh_rec;  /* [gen_nhess][8] */
    /* filled by the model library at launch: how many warp tiles run concurrently on the device; a
     * starting warp prefetches (L2) the inputs of the tile that many positions ahead */
    int32_t tiles_in_flight;
    /* division of a flat item index g < 2^31 by T without a divide: g / T == (g * div_mul) >> div_shift */
    uint32_t div_mul;
    int32_t div_shift;
    int32_t z_per_knot, c_per_knot;  /* N_z / T, N_c / T (prefetch address estimate) */
    /* persistent pipeline kernel (knot_kernel_p): per-warp input staging capacities in doubles (even),
     * index = DTO_IN_*; computed by the runtime over every possible tile start */
    int32_t in_cap[5];
    int32_t nsub_max;     /* most problems one 32-item tile can touch                              */
    int32_t persist_ok;   /* shape is eligible for the persistent kernel (piece count fits a warp)  */
    int32_t w_flat;       /* per-knot parameter slices are monotone: a tile's w is one flat range   */
    int32_t kt_smem;      /* the model library sets it: knot table is staged in shared memory       */
    int64_t shape_id;     /* unique per dto_shape in this process (key of the model library's tile-plan cache) */
    const void* ws_plan;  /* the model library sets it: tile-plan table of the ws kernel (NULL: compute in-kernel) */
    int32_t ws_nout;      /* the model library sets it: output staging buffers per compute warp (ws kernel) */
    int32_t hslot_cap;    /* most Hessian slots 32 consecutive items own (slot staging of the ws kernel) */
} dto_launch_args;

enum { DTO_IN_Z = 0, DTO_IN_SIGMA = 1, DTO_IN_W = 2, DTO_IN_LDYN = 3, DTO_IN_LSTAGE = 4 };

enum { DTO_SEG_G = 0, DTO_SEG_CDYN = 1, DTO_SEG_CSTAGE = 2, DTO_SEG_JDYN = 3, DTO_SEG_JSTAGE = 4, DTO_SEG_HTERM = 5 };

typedef struct dto_model_vtable {
    int32_t abi_version;
    const char* name;
    const char* source_hash;   /* content hash of the generating spec */
    int32_t n_dyn, n_cost, n_stage;
    const dto_element_desc* dyn;
    const dto_element_desc* cost;
    const dto_element_desc* stage;
    const dto_general_desc* general; /* NULL if the model has no general constraint */
    int32_t hess_halo;  /* 1 if a dynamics Hessian has entries in next-state (y) rows */
    int32_t warps_per_cta;
    /* fp64 instruction estimate per knot for the fused Jac+Hess pass (codegen op counts) */
    int32_t ops_fused_per_knot;
    /* compiled Hessian gather recipes (recipes.py): class c has hg_nslots[c] slots whose 4 encoded
     * sources start at hg_src[4*hg_ofs[c]] (k>=0 own term, k<=-2 previous knot's term -k-2, -1 none) */
    int32_t n_hg_classes;
    const int32_t* hg_nslots;
    const int32_t* hg_ofs;
    const int32_t* hg_src;
    /* per class: {own cost terms, own dynamics terms, cost terms of the previous knot}: a knot matches a
     * class only if these agree too (the register-resident gather splits flat term ids by role) */
    const int32_t* hg_meta;
    /* Enqueue kernel `kernel_id` (+ its general-constraint companions) on `stream`.
     * Returns the number of kernels launched (>= 0) or minus a cudaError_t value. */
    int (*launch)(int kernel_id, const dto_launch_args* args, void* stream);
    /* dynamic shared memory (bytes per CTA) the knot kernel `kernel_id` needs */
    int64_t (*smem_bytes)(int kernel_id, const dto_launch_args* args);
} dto_model_vtable;

/* The one symbol a model library exports. */
const dto_model_vtable* dto_model_entry(void);

#ifdef __cplusplus
}
#endif
#endif /* DTO_MODEL_ABI_H */
