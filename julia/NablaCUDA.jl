# NablaCUDA.jl — the Julia-side glue a Nabla.jl maintainer adds to route the array-valued reverse
# sweep to libnabla_cuda.so (include/nabla_cuda.h).  Nabla's own source stays unchanged: this file only
# ADDS a device array type and methods of Nabla's extension surface for it
#   * forward primal dispatch on the unboxed argument types (src/core.jl:85-89, chainrules.jl:183)
#   * preprocess(f, y, ȳ, xs...) -> p  run once per Branch (src/sensitivity.jl:206-221)
#   * ∇(f, Arg{N}, p, y, ȳ, xs...) = p[N]  (src/core.jl:189-193), accumulate ∇(x̄, f, Arg{N}, ...) (:195-205)
#   * ChainRulesCore.rrule for `*`, sum, mean, dot, adjoint on DevArray (picked up automatically by
#     on_new_rule(generate_overload, rrule), src/Nabla.jl:77)
#
# NOT EXECUTED in the build container (no Julia toolchain there): the tested host is the Python mirror
# nabla.jl_b200/, which makes the identical sequence of C-ABI calls.
module NablaCUDA

using Nabla
import Nabla: ∇, preprocess, Arg, Node, unbox
using ChainRulesCore
using LinearAlgebra

const LIB = get(ENV, "NABLA_CUDA_LIB", joinpath(@__DIR__, "..", "nabla.jl_b200", "libnabla_cuda.so"))
const NB_F32, NB_F64 = Int32(0), Int32(1)
nbdtype(::Type{Float32}) = NB_F32
nbdtype(::Type{Float64}) = NB_F64

struct NablaCudaError <: Exception
    status::Int32
    msg::String
end

mutable struct Context
    h::Ptr{Cvoid}
end
function Context(device::Integer=0; pool_bytes::Integer=0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:nb_ctx_create, LIB), Int32, (Int32, UInt64, Ptr{Ptr{Cvoid}}), device, pool_bytes, out)
    st == 0 || throw(NablaCudaError(st, unsafe_string(ccall((:nb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    ctx = Context(out[])
    finalizer(c -> ccall((:nb_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), ctx)
    return ctx
end
const CTX = Ref{Context}()
context() = (isassigned(CTX) || (CTX[] = Context()); CTX[])

function check(st::Int32, ctx::Context=context())
    st == 0 && return nothing
    msg = unsafe_string(ccall((:nb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
    st == 6 && throw(DimensionMismatch(msg))
    st == 4 && throw(MethodError(nothing, (msg,)))      # no device rule: same failure mode as Nabla
    throw(NablaCudaError(st, msg))
end

# nb_tensor (include/nabla_cuda.h): {void* data; int32 dtype; int32 ndim; int64 shape[4]}
struct NbTensor
    data::Ptr{Cvoid}
    dtype::Int32
    ndim::Int32
    shape::NTuple{4,Int64}
end

"Dense column-major array in HBM (ndim ≤ 4); storage is a refcounted block of the ctx pool."
mutable struct DevArray{T<:Union{Float32,Float64},N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    ctx::Context
    function DevArray{T,N}(::UndefInitializer, dims::NTuple{N,Int}; ctx=context()) where {T,N}
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:nb_alloc, LIB), Int32, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Cvoid}}), ctx.h, max(prod(dims) * sizeof(T), 1), out), ctx)
        a = new{T,N}(out[], dims, ctx)
        # nb_release is safe from the GC finalizer thread (pool is mutex-guarded, stream-ordered reuse)
        finalizer(x -> ccall((:nb_release, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, x.ptr), a)
        return a
    end
end
Base.size(a::DevArray) = a.dims
Base.similar(a::DevArray{T}, ::Type{T}, dims::Dims{N}) where {T,N} = DevArray{T,N}(undef, dims; ctx=a.ctx)
tensor(a::DevArray{T,N}) where {T,N} = NbTensor(a.ptr, nbdtype(T), N, ntuple(i -> i <= N ? a.dims[i] : 1, 4))

function DevArray(x::Array{T,N}; ctx=context()) where {T,N}
    a = DevArray{T,N}(undef, size(x); ctx=ctx)
    check(ccall((:nb_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, UInt64), ctx.h, a.ptr, x, sizeof(x)), ctx)
    check(ccall((:nb_sync, LIB), Int32, (Ptr{Cvoid},), ctx.h), ctx)
    return a
end
function Base.Array(a::DevArray{T,N}) where {T,N}
    x = Array{T,N}(undef, a.dims)
    check(ccall((:nb_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{T}, Ptr{Cvoid}, UInt64), a.ctx.h, x, a.ptr, sizeof(x)), a.ctx)
    return x
end

# ---- family (3): `*` through ChainRules (the bridge turns this rrule into Node overloads) ----------
function gemm!(tA::Char, tB::Char, α, A::DevArray{T}, B::DevArray{T}, β, C::DevArray{T}, m, n, k) where {T}
    lda, ldb = size(A, 1), size(B, 1)
    check(ccall((:nb_gemm, LIB), Int32,
                (Ptr{Cvoid}, Cchar, Cchar, Int64, Int64, Int64, Float64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64,
                 Float64, Ptr{Cvoid}, Int64, Int32, Int32),
                C.ctx.h, tA, tB, m, n, k, α, A.ptr, lda, B.ptr, ldb, β, C.ptr, m, nbdtype(T), 0), C.ctx)
    return C
end
function Base.:*(A::DevArray{T,2}, B::DevArray{T,2}) where {T}
    m, k = size(A); n = size(B, 2)
    size(B, 1) == k || throw(DimensionMismatch("A has dimensions $(size(A)) but B has dimensions $(size(B))"))
    return gemm!('N', 'N', 1.0, A, B, 0.0, DevArray{T,2}(undef, (m, n); ctx=A.ctx), m, n, k)
end
function ChainRulesCore.rrule(::typeof(*), A::DevArray{T,2}, B::DevArray{T,2}) where {T}
    Y = A * B
    m, k = size(A); n = size(B, 2)
    function times_pullback(Ȳ)
        Ȳd = unthunk(Ȳ)
        # lazy: Nabla only forces the thunks of Node operands (src/core.jl:151)
        Ā = InplaceableThunk(X̄ -> gemm!('N', 'T', 1.0, Ȳd, B, 1.0, X̄, m, k, n),               # β = 1: add!!
                             @thunk(gemm!('N', 'T', 1.0, Ȳd, B, 0.0, similar(A), m, k, n)))
        B̄ = InplaceableThunk(X̄ -> gemm!('T', 'N', 1.0, A, Ȳd, 1.0, X̄, k, n, m),
                             @thunk(gemm!('T', 'N', 1.0, A, Ȳd, 0.0, similar(B), k, n, m)))
        return NoTangent(), Ā, B̄
    end
    return Y, times_pullback
end

# ---- tape-level fusion around a product (SURVEY §8(f) rank 1; the tested mirror is nabla.jl_b200/fusion.py):
#      the launch of a Float32 product can be deferred so that a following `.+ bias` / `f.` (forward) or
#      `.* f'(y)` / bias row sum (reverse) joins its epilogue.  This is the C entry point those hooks call.
struct NbGemmEpilogue
    kind::Int32          # 1 forward  C = f.(αAB .+ bias);  2 pullback  C = (αAB) .* f'(y)
    op::Int32            # unary nb_op code
    bias::Ptr{Cvoid}
    y::Ptr{Cvoid}
    ldy::Int64
    rowsum::Ptr{Cvoid}   # nullable: rowsum[i] (+)= Σ_j C[i,j]  (the bias sensitivity)
    rowsum_acc::Int32
    emit_split::Int32
end
function gemm_fused!(tA::Char, tB::Char, α, A::DevArray{Float32}, B::DevArray{Float32}, C::DevArray{Float32}, m, n, k,
                     epi::NbGemmEpilogue)
    st = ccall((:nb_gemm_fused, LIB), Int32,
               (Ptr{Cvoid}, Cchar, Cchar, Int64, Int64, Int64, Float64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64,
                Ptr{Cvoid}, Int64, Int32, Int32, Ref{NbGemmEpilogue}),
               C.ctx.h, tA, tB, m, n, k, α, A.ptr, size(A, 1), B.ptr, size(B, 1), C.ptr, m, nbdtype(Float32), 0, epi)
    st == 4 && return nothing      # NB_ERR_UNSUPPORTED: the caller runs the unfused launches
    check(st, C.ctx)
    return C
end

# ---- family (1): broadcast.  Scalar functions are lowered to device programs by `program(f, nargs)`
#      (catalogue opcode for Base functions; a tracer over a symbolic scalar type for user lambdas,
#      mirroring nabla.jl_b200/scalar.py); unknown primitives throw MethodError — no CPU fallback. ----
function bcast_fwd(prog::Ptr{Cvoid}, args::Vector{<:DevArray{T}}, out::DevArray{T}) where {T}
    ts = [tensor(a) for a in args]
    check(ccall((:nb_bcast_fwd, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{NbTensor}, Ref{NbTensor}),
                out.ctx.h, prog, length(ts), ts, tensor(out)), out.ctx)
    return out
end
Base.broadcast(f, A::DevArray{T}, rest::DevArray{T}...) where {T} =
    bcast_fwd(program(f, 1 + length(rest)), [A, rest...],
              DevArray{T,length(Broadcast.broadcast_shape(size(A), size.(rest)...))}(undef,
                       Broadcast.broadcast_shape(size(A), size.(rest)...); ctx=A.ctx))

# One fused launch for every Node operand, run ONCE per Branch through `preprocess`; `∇` then just
# picks its slot — exactly how the ChainRules bridge uses p (chainrules.jl:243-251, 286).
function Nabla.preprocess(::typeof(broadcast), y::Node{<:DevArray}, ȳ, f, args...)
    xs = map(unbox, args)
    want = UInt32(0)
    outs = NbTensor[]
    bufs = Any[]
    for (k, a) in enumerate(args)
        if a isa Node
            want |= UInt32(1) << (k - 1)
            b = similar(unbox(a)); push!(bufs, b); push!(outs, tensor(b))
        else
            push!(bufs, nothing); push!(outs, NbTensor(C_NULL, 0, 0, (1, 1, 1, 1)))
        end
    end
    ts = [tensor(x) for x in xs]
    ctx = unbox(y).ctx
    check(ccall((:nb_bcast_pullback, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{NbTensor}, Ref{NbTensor}, Int32, Ptr{NbTensor}, UInt32, Ptr{NbTensor}, UInt32),
                ctx.h, program(f, length(xs)), tensor(ȳ), tensor(unbox(y)), length(ts), ts, want, outs, UInt32(0)), ctx)
    return (nothing, bufs...)          # p[1] is the (non-differentiable) function slot
end
Nabla.∇(::typeof(broadcast), ::Type{Arg{N}}, p, y::DevArray, ȳ, f, args...) where {N} = p[N]
# accumulate: add!! in place on the slot buffer (src/core.jl:203-205) — one fused `+=` launch
function Nabla.∇(x̄::DevArray{T}, ::typeof(broadcast), ::Type{Arg{N}}, p, y::DevArray, ȳ, f, args...) where {T,N}
    t = tensor(p[N])
    check(ccall((:nb_bcast_reduce, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{NbTensor}, Float64, Ref{NbTensor}, Int32),
                x̄.ctx.h, C_NULL, 1, t, 1.0, tensor(x̄), 1), x̄.ctx)
    return x̄
end

# ---- family (2): reductions (forward shown; pullbacks go through nb_bcast_pullback with the
#      identity / f program and ȳ broadcast from the reduced shape, reducedim.jl:11-20) --------------
function Base.sum(A::DevArray{T,N}; dims=:) where {T,N}
    rd = dims === Colon() ? () : ntuple(i -> i in dims ? 1 : size(A, i), N)
    out = DevArray{T,length(rd)}(undef, rd; ctx=A.ctx)
    check(ccall((:nb_bcast_reduce, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{NbTensor}, Float64, Ref{NbTensor}, Int32),
                A.ctx.h, C_NULL, 1, tensor(A), 1.0, tensor(out), 0), A.ctx)
    return out
end

# ---- batch-sharded gradients: after ∇(y) on every rank's shard ------------------------------------
function allreduce_sum!(bufs::Vector{<:DevArray{T}}) where {T}
    ctx = first(bufs).ctx
    ptrs = [b.ptr for b in bufs]; counts = Int64[length(b) for b in bufs]
    check(ccall((:nb_allreduce_sum, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int32),
                ctx.h, length(bufs), ptrs, counts, nbdtype(T)), ctx)
end

# ---- scalar-function lowering: `program(f, nargs)` -> nb_prog* (nb_prog_create) ---------------------
#      The tested twin is nabla.jl_b200/scalar.py (compile_scalar).  A catalogue function maps to its opcode; any
#      other scalar function is called ONCE on symbolic scalars that record a register program (with structural
#      common-subexpression elimination); functions that branch on values or call primitives outside the
#      catalogue throw MethodError, as an un-ruled function does in Nabla (test/core.jl:159).
import SpecialFunctions
const SF = SpecialFunctions
const NB_MAX_ARGS, NB_MAX_INSTR, NB_MAX_CONSTS = 6, 32, 16

# function => nb_op (include/nabla_cuda.h; the values are ABI).  The list is test/runtests.jl:28-52.
const UNARY_OPS = Dict{Any,Int32}(
    identity => 0, abs => 2, abs2 => 3, sqrt => 4, cbrt => 5, inv => 6, exp => 7, exp2 => 8, exp10 => 9,
    expm1 => 10, log => 11, log2 => 12, log10 => 13, log1p => 14, sin => 15, cos => 16, tan => 17, sec => 18,
    csc => 19, cot => 20, sinh => 21, cosh => 22, tanh => 23, sech => 24, csch => 25, coth => 26, asin => 27,
    acos => 28, atan => 29, asec => 30, acsc => 31, acot => 32, asinh => 33, acosh => 34, atanh => 35,
    asech => 36, acsch => 37, acoth => 38, sind => 39, cosd => 40, tand => 41, secd => 42, cscd => 43,
    cotd => 44, asind => 45, acosd => 46, atand => 47, asecd => 48, acscd => 49, acotd => 50, sinpi => 51,
    cospi => 52, deg2rad => 53, rad2deg => 54,
    SF.erf => 56, SF.erfc => 57, SF.gamma => 96, SF.digamma => 97, SF.trigamma => 98, SF.besselj0 => 99,
    SF.besselj1 => 100, SF.bessely0 => 101, SF.bessely1 => 102, SF.erfinv => 103, SF.erfcinv => 104,
    SF.erfcx => 105, SF.airyai => 106, SF.airyaiprime => 107, SF.airybi => 108, SF.airybiprime => 109,
    SF.dawson => 110, SF.erfi => 111, SF.invdigamma => 112)
const BINARY_OPS = Dict{Any,Int32}(
    (+) => 64, (-) => 65, (*) => 66, (/) => 67, (\) => 68, (^) => 69, hypot => 70, max => 71, min => 72,
    SF.beta => 74,
    # differentiable in the second argument only; the first is the integer order (test/runtests.jl:43-47)
    SF.besselj => 75, SF.bessely => 76, SF.besseli => 77, SF.besselk => 78, SF.polygamma => 79)
const OP_NEG, OP_LOGISTIC, OP_ATAN2 = Int32(1), Int32(55), Int32(73)

struct Sym <: Real
    kind::Symbol                 # :arg, :const, :op
    op::Int32
    a::Union{Sym,Nothing}
    b::Union{Sym,Nothing}
    value::Float64
    index::Int32
end
symarg(i) = Sym(:arg, 0, nothing, nothing, 0.0, i)
symconst(v::Real) = Sym(:const, 0, nothing, nothing, Float64(v), 0)
symop(op, a, b=nothing) = Sym(:op, op, a, b, 0.0, 0)
Base.promote_rule(::Type{Sym}, ::Type{<:Real}) = Sym
Base.convert(::Type{Sym}, v::Real) = symconst(v)
Base.convert(::Type{Sym}, s::Sym) = s
for (f, op) in UNARY_OPS
    f === identity && continue
    m = parentmodule(f); n = nameof(f)
    @eval $m.$n(x::Sym) = symop($op, x)
end
Base.:-(x::Sym) = symop(OP_NEG, x)
for (f, op) in BINARY_OPS
    m = parentmodule(f); n = nameof(f)
    @eval $m.$n(x::Sym, y::Sym) = symop($op, x, y)
    @eval $m.$n(x::Sym, y::Real) = symop($op, x, symconst(y))
    @eval $m.$n(x::Real, y::Sym) = symop($op, symconst(x), y)
end
Base.atan(y::Sym, x::Sym) = symop(OP_ATAN2, y, x)
Base.literal_pow(::typeof(^), x::Sym, ::Val{p}) where {p} = symop(BINARY_OPS[^], x, symconst(p))
# comparisons would make the trace depend on values: refuse
Base.isless(::Sym, ::Sym) = throw(MethodError(isless, (Sym, Sym)))
Base.:<(::Sym, ::Real) = throw(MethodError(<, (Sym, Real)))

const PROGRAMS = Dict{Tuple{UInt,Int},Ptr{Cvoid}}()

function program(f, nargs::Int)
    get!(PROGRAMS, (objectid(f), nargs)) do
        nargs <= NB_MAX_ARGS || throw(MethodError(f, ntuple(_ -> Sym, nargs)))
        args = [symarg(Int32(i - 1)) for i in 1:nargs]
        root = if nargs == 1 && haskey(UNARY_OPS, f)
            f === identity ? args[1] : symop(UNARY_OPS[f], args[1])
        elseif nargs == 2 && haskey(BINARY_OPS, f)
            symop(BINARY_OPS[f], args[1], args[2])
        else
            convert(Sym, f(args...))          # trace
        end
        ops, ia, ib, consts = Int32[], Int32[], Int32[], Float64[]
        memo = Dict{Tuple{Int32,Int32,Int32},Int32}()
        function emit(s::Sym)::Int32
            s.kind === :arg && return s.index
            if s.kind === :const
                k = findfirst(c -> c === s.value, consts)
                k === nothing && (push!(consts, s.value); k = length(consts))
                length(consts) <= NB_MAX_CONSTS || throw(MethodError(f, ntuple(_ -> Sym, nargs)))
                return Int32(-k)                # consts[c] is named -(c+1), c zero-based
            end
            a = emit(s.a)
            b = s.b === nothing ? Int32(0) : emit(s.b)
            get!(memo, (s.op, a, b)) do
                push!(ops, s.op); push!(ia, a); push!(ib, b)
                length(ops) <= NB_MAX_INSTR || throw(MethodError(f, ntuple(_ -> Sym, nargs)))
                Int32(nargs + length(ops) - 1)
            end
        end
        r = emit(root)
        if root.kind === :const                 # constant function: c + 0 keeps operand 0 for the shape
            push!(ops, BINARY_OPS[+]); push!(ia, r); push!(ib, emit(symconst(0.0)))
        elseif root.kind === :arg && r != 0     # f(x, y) = y
            push!(ops, Int32(0)); push!(ia, r); push!(ib, Int32(0))
        end
        out = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:nb_prog_create, LIB), Int32,
                   (Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                   nargs, length(ops), ops, ia, ib, length(consts), consts, out)
        st == 0 || throw(MethodError(f, ntuple(_ -> Sym, nargs)))   # NB_ERR_UNSUPPORTED: outside the catalogue
        out[]
    end
end
# logistic(x) = 1 / (1 + exp(-x)) (examples/mlp.jl:42) traces to four instructions; the Python twin's peephole
# (scalar.py:_simplify) rewrites that pattern to the single NB_OP_LOGISTIC the fused epilogues know — do the same
# here by registering the user's function:  NablaCUDA.UNARY_OPS[logistic] = NablaCUDA.OP_LOGISTIC.

end # module
