# DTOB200.jl -- Julia glue for the B200 batched NLP-callback engine (libdto.so).
#
# Written against DirectTrajectoryOptimization.jl as vendored at /root/reference
# (Julia 1.6, Symbolics 0.1.29-0.1.32, MathOptInterface 1.3, Ipopt.jl 1.0.2; Project.toml:16-23).
# NOT EXECUTED in this repository's CI: the build image has no Julia. It is kept small and literal:
# every ccall below binds one entry point of include/dto.h, and the evaluator methods mirror
# /root/reference/src/moi.jl one for one.
#
# What it does
#   1. `export_spec(solver)`  re-traces nothing: it re-builds the Symbolics expressions exactly like
#      the reference constructors do (src/dynamics.jl:23-35, src/costs.jl:18-27,
#      src/constraints.jl:27-40, src/general_constraint.jl:23-36), prints them as infix text and
#      writes the JSON model spec that `python -m dto_b200.spec_io spec.json` compiles to a CUDA model
#      library (content-addressed: the reference's "#TODO: option to load/save methods").
#   2. `BatchedNLPData` <: MOI.AbstractNLPEvaluator wraps a dto_batch handle. With batch = 1 it is a
#      drop-in for the reference's NLPData inside MOI.NLPBlockData (src/data.jl:233-234); with
#      batch = B it serves B lock-step Ipopt instances (see INTEGRATION.md, "Batched Ipopt driver").
module DTOB200

using MathOptInterface
const MOI = MathOptInterface
using Symbolics
using LinearAlgebra: dot

const libdto = get(ENV, "DTO_LIB", "libdto.so")

# ---------------------------------------------------------------- thin ccall layer (include/dto.h)
check(status::Cint) = status == 0 || error("dto: " * unsafe_string(ccall((:dto_last_error, libdto), Cstring, ())))

struct ShapeDesc                       # dto_shape_desc
    T::Int32
    dynamics_kind::Ptr{Int32}
    cost_kind::Ptr{Int32}
    stage_kind::Ptr{Int32}
    use_general::Int32
    parameter_dim::Ptr{Int32}
    parameter_offset::Ptr{Int32}
    num_parameter::Int32
end

function model_load(path::String)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dto_model_load, libdto), Cint, (Cstring, Ref{Ptr{Cvoid}}), path, h))
    h[]
end

function shape_create(model, T, kd::Vector{Int32}, kc::Vector{Int32}, ks::Vector{Int32}, use_general::Bool,
                      pdim::Vector{Int32})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve kd kc ks pdim begin
        d = Ref(ShapeDesc(T, pointer(kd), pointer(kc), pointer(ks), use_general, pointer(pdim), C_NULL, 0))
        check(ccall((:dto_shape_create, libdto), Cint, (Ptr{Cvoid}, Ref{ShapeDesc}, Ref{Ptr{Cvoid}}), model, d, h))
    end
    h[]
end

function batch_create(shape, B::Integer, devices::Vector{Cint})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dto_batch_create, libdto), Cint, (Ptr{Cvoid}, Int64, Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}),
                shape, B, devices, length(devices), h))
    h[]
end

num_variables(s)  = ccall((:dto_num_variables, libdto), Int64, (Ptr{Cvoid},), s)
num_constraint(s) = ccall((:dto_num_constraint, libdto), Int64, (Ptr{Cvoid},), s)
num_jacobian(s)   = ccall((:dto_num_jacobian, libdto), Int64, (Ptr{Cvoid},), s)
num_hessian(s)    = ccall((:dto_num_hessian, libdto), Int64, (Ptr{Cvoid},), s)

function structure(f::Symbol, s, n)
    r = Vector{Int64}(undef, n); c = Vector{Int64}(undef, n)
    check(ccall((f, libdto), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), s, r, c))
    collect(zip(r, c))                  # Vector{Tuple{Int,Int}}, 1-based, reference order
end

# ---------------------------------------------------------------- evaluator
mutable struct BatchedNLPData <: MOI.AbstractNLPEvaluator
    model::Ptr{Cvoid}
    shape::Ptr{Cvoid}
    batch::Ptr{Cvoid}
    B::Int
    hessian_lagrangian::Bool
    jacobian_sparsity::Vector{Tuple{Int,Int}}
    hessian_lagrangian_sparsity::Vector{Tuple{Int,Int}}
    sigma::Vector{Float64}
end

function BatchedNLPData(model_path, T, kd, kc, ks; use_general=false, pdim=zeros(Int32, T), batch=1,
                        devices=Cint[0], evaluate_hessian=false)
    m = model_load(model_path)
    s = shape_create(m, T, Int32.(kd), Int32.(kc), Int32.(ks), use_general, Int32.(pdim))
    b = batch_create(s, batch, Cint.(devices))
    BatchedNLPData(m, s, b, batch, evaluate_hessian,
                   structure(:dto_jacobian_structure, s, num_jacobian(s)),
                   structure(:dto_hessian_lagrangian_structure, s, num_hessian(s)), ones(batch))
end

set_x!(nlp, z) = check(ccall((:dto_set_x, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, z))

# the five callbacks, method for method as /root/reference/src/moi.jl:1-120 (batch = 1: z is a Vector;
# batch = B: z, outputs are B x n row-major, i.e. Julia matrices of size (n, B))
function MOI.eval_objective(nlp::BatchedNLPData, z)
    set_x!(nlp, z)
    f = Vector{Float64}(undef, nlp.B)
    check(ccall((:dto_eval_objective, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, f))
    nlp.B == 1 ? f[1] : f
end
function MOI.eval_objective_gradient(nlp::BatchedNLPData, g, z)
    set_x!(nlp, z)
    check(ccall((:dto_eval_objective_gradient, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, g)); return
end
function MOI.eval_constraint(nlp::BatchedNLPData, c, z)
    set_x!(nlp, z)
    check(ccall((:dto_eval_constraint, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, c)); return
end
function MOI.eval_constraint_jacobian(nlp::BatchedNLPData, J, z)
    set_x!(nlp, z)
    check(ccall((:dto_eval_constraint_jacobian, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, J)); return
end
function MOI.eval_hessian_lagrangian(nlp::BatchedNLPData, H, z, σ, λ)
    set_x!(nlp, z)
    fill!(nlp.sigma, σ)
    check(ccall((:dto_set_duals, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.batch, nlp.sigma, λ))
    check(ccall((:dto_eval_hessian_lagrangian, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, H)); return
end
MOI.features_available(nlp::BatchedNLPData) = nlp.hessian_lagrangian ? [:Grad, :Jac, :Hess] : [:Grad, :Jac]
MOI.initialize(nlp::BatchedNLPData, features) = nothing
MOI.jacobian_structure(nlp::BatchedNLPData) = nlp.jacobian_sparsity
MOI.hessian_lagrangian_structure(nlp::BatchedNLPData) = nlp.hessian_lagrangian_sparsity

# ---------------------------------------------------------------- spec export (Symbolics -> JSON text)
ascii_name(prefix, i) = string(prefix, i)
function plain_variables(prefix, n)      # @variables x[1:n] but with ASCII names x1..xn
    [Symbolics.variable(Symbol(ascii_name(prefix, i))) for i in 1:n]
end
totext(e) = replace(string(e), "π" => string(Float64(pi)))   # Julia infix: + - * / ^, sin cos tan ...

"Re-trace `f` like Dynamics(f, ny, nx, nu; ...) does (src/dynamics.jl:23-35) and describe it for the code generator."
function dynamics_spec(f, ny, nx, nu; num_parameter=0, evaluate_hessian=false)
    y, x, u, w = plain_variables("y", ny), plain_variables("x", nx), plain_variables("u", nu), plain_variables("w", num_parameter)
    ev = f(y, x, u, w)
    jac = Symbolics.sparsejacobian(ev, [x; u; y])
    I, J, _ = findnz(jac)
    d = Dict("num_next_state" => ny, "num_state" => nx, "num_action" => nu, "num_parameter" => num_parameter,
             "evaluate" => totext.(ev), "jacobian_sparsity" => [I, J], "evaluate_hessian" => evaluate_hessian,
             "hessian_sparsity" => [Int[], Int[]])
    if evaluate_hessian
        λ = plain_variables("lam", ny)
        Hs = Symbolics.hessian_sparsity(dot(λ, ev), [x; u; y])
        HI, HJ, _ = findnz(Hs)
        d["hessian_sparsity"] = [HI, HJ]
    end
    d
end
# cost_spec / constraint_spec / general_spec follow the same three lines with the variable lists of
# src/costs.jl:18-27, src/constraints.jl:27-40, src/general_constraint.jl:23-36.

# ---------------------------------------------------------------- device-resident KKT consumer (dto_kkt_*)
# What examples/pendulum/pendulum.jl:138-211 does with the callback outputs (assemble
# [[H + primal_reg I, C'], [C, -dual_reg I]] and h = [grad + C'y; c], qdldl, solve!), for the whole
# batch on the GPU; only the solution comes back.
mutable struct BatchedKKT
    handle::Ptr{Cvoid}
    nlp::BatchedNLPData
    dim::Int
end
function BatchedKKT(nlp::BatchedNLPData; primal_reg=1.0e-5, dual_reg=1.0e-5)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dto_kkt_create, libdto), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ref{Ptr{Cvoid}}), nlp.batch, primal_reg, dual_reg, h))
    BatchedKKT(h[], nlp, Int(ccall((:dto_kkt_dim, libdto), Int64, (Ptr{Cvoid},), h[])))
end
"sol[dim, B] = K \\ h at (z, y) for every problem (column b = problem b; Julia is column-major, the ABI problem-major)"
function solve!(sol::Matrix{Float64}, kkt::BatchedKKT, z::Matrix{Float64}, y::Matrix{Float64})
    nlp = kkt.nlp
    check(ccall((:dto_set_x, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), nlp.batch, z))
    fill!(nlp.sigma, 1.0)                                               # pendulum.jl:136 evaluates with sigma = 1.0
    check(ccall((:dto_set_duals, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), nlp.batch, nlp.sigma, y))
    check(ccall((:dto_kkt_solve, libdto), Cint, (Ptr{Cvoid}, Ptr{Float64}), kkt.handle, sol))
    sol
end

end # module
