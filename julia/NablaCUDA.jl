# NablaCUDA.jl — the Julia-side glue a Nabla.jl maintainer adds to route the array-valued reverse
# sweep to libnabla_cuda.so (include/nabla_cuda.h).  Nabla's own source stays unchanged: this file only
# ADDS a device array type and methods of Nabla's extension surface for it
#   * forward primal dispatch on the unboxed argument types (src/core.jl:85-89, chainrules.jl:183)
#   * preprocess(f, y, ȳ, xs...) -> p  run once per Branch (src/sensitivity.jl:206-221)
#   * ∇(f, Arg{N}, p, y, ȳ, xs...) = p[N]  (src/core.jl:189-193), accumulate ∇(x̄, f, Arg{N}, ...) (:195-205)
#   * ChainRulesCore.rrule for `*` (matrix*matrix, matrix*vector), sum, sum(abs2, .), mean, dot, adjoint on DevArray
#     (picked up automatically by on_new_rule(generate_overload, rrule), src/Nabla.jl:77)
#   * a Broadcast style, so that `f.(args...)` on DevArrays (the bodies of Nabla's own rules, e.g. `2ȳ .* A`,
#     reducedim.jl:50) lowers to ONE traced device program instead of scalar indexing
#
# NOT EXECUTED in the build container (no Julia toolchain there): the tested host is the Python mirror
# nabla.jl_b200/, which makes the identical sequence of C-ABI calls.  What CAN be checked without Julia is checked by
# tests/test_host_logic.py: every ccall against the header's prototypes, the op tables against the header's enum.
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
    # finalizer order is unspecified: null the handle so that array finalizers running later skip nb_release
    # (nb_ctx_destroy frees every block of the pool itself)
    finalizer(ctx) do c
        h = c.h
        c.h = C_NULL
        h == C_NULL || ccall((:nb_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), h)
    end
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
        finalizer(x -> x.ctx.h == C_NULL || ccall((:nb_release, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, x.ptr), a)
        return a
    end
end
Base.size(a::DevArray) = a.dims
# the data lives in HBM: no scalar indexing (generic AbstractArray fallbacks must fail loudly, not crawl or crash)
Base.getindex(::DevArray, ::Int...) = error("scalar indexing of a NablaCUDA.DevArray is not supported; copy it with Array(x)")
Base.setindex!(::DevArray, _, ::Int...) = error("scalar indexing of a NablaCUDA.DevArray is not supported")
Base.show(io::IO, a::DevArray{T,N}) where {T,N} = print(io, "DevArray{", T, ",", N, "}", a.dims)
Base.show(io::IO, ::MIME"text/plain", a::DevArray) = show(io, a)
function Base.fill!(a::DevArray{T}, v) where {T}
    check(ccall((:nb_fill, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, UInt64, Float64),
                a.ctx.h, a.ptr, nbdtype(T), length(a), Float64(v)), a.ctx)
    return a
end
Base.zero(a::DevArray) = fill!(similar(a), 0)      # unused arguments get zero(arg) (src/core.jl:220-224)
Base.one(a::DevArray{T,0}) where {T} = fill!(similar(a), 1)
Base.copy(a::DevArray{T,N}) where {T,N} = (b = similar(a);
    check(ccall((:nb_d2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64), a.ctx.h, b.ptr, a.ptr, length(a) * sizeof(T)), a.ctx); b)
# A device objective is a 0-dim DevArray; Nabla's ∇(y::Node{<:∇Scalar}) only seeds Numbers (src/core.jl:201)
Nabla.∇(y::Node{<:DevArray{T,0}}) where {T} = ∇(y, one(unbox(y)))
Base.getindex(a::DevArray{T,0}) where {T} = Array(a)[]      # reading the objective synchronises (nb_d2h)
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

# matrix * vector and its adjoints (ChainRules rrule(*) with a vector operand; examples/symmetric_quadratic.jl:15)
function gemv!(tA::Char, α, A::DevArray{T,2}, x::DevArray{T,1}, β, y::DevArray{T,1}) where {T}
    m, n = size(A)
    check(ccall((:nb_gemv, LIB), Int32,
                (Ptr{Cvoid}, Cchar, Int64, Int64, Float64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Int32),
                y.ctx.h, tA, m, n, α, A.ptr, m, x.ptr, β, y.ptr, nbdtype(T)), y.ctx)
    return y
end
function ger!(α, x::DevArray{T,1}, y::DevArray{T,1}, β, A::DevArray{T,2}) where {T}
    m, n = size(A)
    check(ccall((:nb_ger, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Int64, Int32),
                A.ctx.h, m, n, α, x.ptr, y.ptr, β, A.ptr, m, nbdtype(T)), A.ctx)
    return A
end
function Base.:*(A::DevArray{T,2}, x::DevArray{T,1}) where {T}
    size(A, 2) == length(x) || throw(DimensionMismatch("A has dimensions $(size(A)) but x has length $(length(x))"))
    return gemv!('N', 1.0, A, x, 0.0, DevArray{T,1}(undef, (size(A, 1),); ctx=A.ctx))
end
function ChainRulesCore.rrule(::typeof(*), A::DevArray{T,2}, x::DevArray{T,1}) where {T}
    y = A * x
    function times_vec_pullback(ȳ)
        ȳd = unthunk(ȳ)
        Ā = InplaceableThunk(X̄ -> ger!(1.0, ȳd, x, 1.0, X̄), @thunk(ger!(1.0, ȳd, x, 0.0, similar(A))))
        x̄ = InplaceableThunk(X̄ -> gemv!('T', 1.0, A, ȳd, 1.0, X̄), @thunk(gemv!('T', 1.0, A, ȳd, 0.0, similar(x))))
        return NoTangent(), Ā, x̄
    end
    return y, times_vec_pullback
end
function Base.adjoint(A::DevArray{T,2}) where {T}      # materialised (ChainRules rrule(adjoint): x̄ = ȳ')
    m, n = size(A)
    out = DevArray{T,2}(undef, (n, m); ctx=A.ctx)
    check(ccall((:nb_transpose, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32),
                A.ctx.h, m, n, A.ptr, m, out.ptr, n, nbdtype(T)), A.ctx)
    return out
end
Base.transpose(A::DevArray{T,2}) where {T} = adjoint(A)
ChainRulesCore.rrule(::typeof(adjoint), A::DevArray{T,2}) where {T} =
    adjoint(A), Ȳ -> (NoTangent(), adjoint(unthunk(Ȳ)))

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

# out (+)= scale * Σ_{dims where out has extent 1} f.(args...) — the one reduction entry point (nb_bcast_reduce)
function reduce!(prog::Ptr{Cvoid}, args::Vector{<:DevArray{T}}, scale, out::DevArray{T}; accumulate::Bool=false) where {T}
    ts = [tensor(a) for a in args]
    check(ccall((:nb_bcast_reduce, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{NbTensor}, Float64, Ref{NbTensor}, Int32),
                out.ctx.h, prog, length(ts), ts, Float64(scale), tensor(out), accumulate ? 1 : 0), out.ctx)
    return out
end
reduced_dims(A::DevArray{T,N}, dims) where {T,N} = dims === Colon() ? () : ntuple(i -> i in dims ? 1 : size(A, i), N)
denominator(A::DevArray, dims) = dims === Colon() ? length(A) : prod(size(A, d) for d in dims)      # _denom, reducedim.jl:60-62
# ȳ (reduced shape) broadcast back to size(x), scaled: the device form of ChainRules' sum / mean pullback
function expand(ȳ::DevArray{T}, like::DevArray{T,N}, scale) where {T,N}
    return reduce!(C_NULL, [ȳ], scale, DevArray{T,N}(undef, size(like); ctx=like.ctx))
end
function ChainRulesCore.rrule(::typeof(sum), A::DevArray{T}; dims=:) where {T}
    return sum(A; dims=dims), ȳ -> (NoTangent(), @thunk(expand(unthunk(ȳ), A, 1.0)))
end
function Statistics_mean(A::DevArray{T,N}; dims=:) where {T,N}
    rd = reduced_dims(A, dims)
    return reduce!(C_NULL, [A], 1.0 / denominator(A, dims), DevArray{T,length(rd)}(undef, rd; ctx=A.ctx))
end
# sum(abs2, A): forward one fused reduction; pullback 2ȳ .* A (functional/reducedim.jl:42-51) as ONE fused launch
function Base.sum(::typeof(abs2), A::DevArray{T,N}; dims=:) where {T,N}
    rd = reduced_dims(A, dims)
    return reduce!(program(abs2, 1), [A], 1.0, DevArray{T,length(rd)}(undef, rd; ctx=A.ctx))
end
function pullback!(prog::Ptr{Cvoid}, ȳ::DevArray{T}, ysaved, args::Vector{<:DevArray{T}}, outs::Vector, acc::UInt32) where {T}
    want = UInt32(0)
    ots = NbTensor[]
    for (k, o) in enumerate(outs)
        if o === nothing
            push!(ots, NbTensor(C_NULL, 0, 0, (1, 1, 1, 1)))
        else
            want |= UInt32(1) << (k - 1); push!(ots, tensor(o))
        end
    end
    ts = [tensor(x) for x in args]
    ctx = ȳ.ctx
    ys = ysaved === nothing ? Ptr{NbTensor}(C_NULL) : Ref(tensor(ysaved))
    check(ccall((:nb_bcast_pullback, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{NbTensor}, Ptr{NbTensor}, Int32, Ptr{NbTensor}, UInt32, Ptr{NbTensor}, UInt32),
                ctx.h, prog, tensor(ȳ), ys, length(ts), ts, want, ots, acc), ctx)
    return outs
end
function ChainRulesCore.rrule(::typeof(sum), ::typeof(abs2), A::DevArray{T}; dims=:) where {T}
    y = sum(abs2, A; dims=dims)
    return y, ȳ -> (NoTangent(), NoTangent(), @thunk(pullback!(program(abs2, 1), unthunk(ȳ), nothing, [A], Any[similar(A)], UInt32(0))[1]))
end
# dot(x, y): x̄ = ȳ·y, ȳ' = ȳ·x — one fused reduction forward, one fused pullback launch for both operands
function LinearAlgebra.dot(x::DevArray{T,N}, y::DevArray{T,N}) where {T,N}
    return reduce!(program(*, 2), [x, y], 1.0, DevArray{T,0}(undef, (); ctx=x.ctx))
end
function ChainRulesCore.rrule(::typeof(dot), x::DevArray{T}, y::DevArray{T}) where {T}
    function dot_pullback(s̄)
        g = pullback!(program(*, 2), unthunk(s̄), nothing, [x, y], Any[similar(x), similar(y)], UInt32(0))
        return NoTangent(), g[1], g[2]
    end
    return dot(x, y), dot_pullback
end

# ---- broadcasting: `f.(args...)` with DevArray operands (and scalars, baked in as constants) is ONE device
#      program — this is what makes the bodies of Nabla's own rules (`2ȳ .* A`, `ȳ .* f'(A)`) run on the device
struct DevStyle <: Broadcast.AbstractArrayStyle{Any} end
DevStyle(::Val) = DevStyle()
Base.BroadcastStyle(::Type{<:DevArray}) = DevStyle()
Base.BroadcastStyle(s::DevStyle, ::Broadcast.DefaultArrayStyle{0}) = s
function Base.copy(bc::Broadcast.Broadcasted{DevStyle})
    fb = Broadcast.flatten(bc)
    isdev = map(a -> a isa DevArray, fb.args)
    devs = DevArray[a for a in fb.args if a isa DevArray]
    all(a -> a isa DevArray || a isa Number || a isa Base.RefValue, fb.args) ||
        throw(MethodError(broadcast, (fb.f, fb.args...)))       # a host Array among device operands: no silent copy
    T = eltype(first(devs))
    consts = map(a -> a isa DevArray ? nothing : (a isa Base.RefValue ? a[] : a), fb.args)
    g = function (xs...)
        k = 0
        fb.f(map((c, d) -> d ? xs[k += 1] : c, consts, isdev)...)
    end
    shp = Broadcast.broadcast_shape(map(size, devs)...)
    out = DevArray{T,length(shp)}(undef, shp; ctx=first(devs).ctx)
    return bcast_fwd(program(g, length(devs); cache=false), convert(Vector{DevArray{T}}, devs), out)
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
"`step(x) = x > 0 ? 1 : 0` (NB_OP_STEP)"
step(x::Real) = x > 0 ? one(x) : zero(x)
const NB_MAX_ARGS, NB_MAX_INSTR, NB_MAX_CONSTS = 10, 32, 16

# function => nb_op (include/nabla_cuda.h; the values are ABI).  The list is test/runtests.jl:28-52.
const UNARY_OPS = Dict{Any,Int32}(
    identity => 0, abs => 2, abs2 => 3, sqrt => 4, cbrt => 5, inv => 6, exp => 7, exp2 => 8, exp10 => 9,
    expm1 => 10, log => 11, log2 => 12, log10 => 13, log1p => 14, sin => 15, cos => 16, tan => 17, sec => 18,
    csc => 19, cot => 20, sinh => 21, cosh => 22, tanh => 23, sech => 24, csch => 25, coth => 26, asin => 27,
    acos => 28, atan => 29, asec => 30, acsc => 31, acot => 32, asinh => 33, acosh => 34, atanh => 35,
    asech => 36, acsch => 37, acoth => 38, sind => 39, cosd => 40, tand => 41, secd => 42, cscd => 43,
    cotd => 44, asind => 45, acosd => 46, atand => 47, asecd => 48, acscd => 49, acotd => 50, sinpi => 51,
    cospi => 52, deg2rad => 53, rad2deg => 54,
    sign => 58, step => 59, floor => 60, ceil => 61, round => 62, trunc => 63,   # derivative zero (ForwardDiff)
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
# comparisons (1.0 / 0.0) and selection: ifelse(c, a, b) = MASK(c, a) + MASKN(c, b)
const OP_GT, OP_LT, OP_GE, OP_LE, OP_MASK, OP_MASKN = Int32(80), Int32(81), Int32(82), Int32(83), Int32(84), Int32(85)

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
# Value-dependent scalar functions.  `x > 0 ? a : b` needs a Bool, which a symbolic trace cannot give (the trace
# would depend on values): comparisons of Syms return a Sym (1.0 / 0.0, derivative zero) and the selection is written
# `ifelse(x > 0, a, b)` — both branches traced, value and derivative of the branch taken, which is what ForwardDiff's
# duals give under fmad (src/core.jl:285-292).  Using a comparison as a Bool (`if`, `?:`, `&&`) raises a TypeError.
for (f, op) in ((:>, OP_GT), (:<, OP_LT), (:>=, OP_GE), (:<=, OP_LE))
    @eval Base.$f(x::Sym, y::Sym) = symop($op, x, y)
    @eval Base.$f(x::Sym, y::Real) = symop($op, x, symconst(y))
    @eval Base.$f(x::Real, y::Sym) = symop($op, symconst(x), y)
end
Base.ifelse(c::Sym, a, b) = symop(BINARY_OPS[+], symop(OP_MASK, c, convert(Sym, a)), symop(OP_MASKN, c, convert(Sym, b)))
Base.isless(::Sym, ::Sym) = throw(MethodError(isless, (Sym, Sym)))   # sorting / max of traces: use max, min

# Cache: only functions that ARE their type (no captured state: named functions, non-capturing lambdas) are keyed —
# by the function object itself, which the Dict keeps alive, so an id can never be recycled.  A closure is re-traced on
# every call (`cache=false` path): keying it by objectid would serve a stale program once a captured value changes.
const PROGRAMS = Dict{Tuple{Any,Int},Ptr{Cvoid}}()

function program(f, nargs::Int; cache::Bool=true)
    (cache && Base.issingletontype(typeof(f))) || return trace_program(f, nargs)
    return get!(() -> trace_program(f, nargs), PROGRAMS, (f, nargs))
end

function trace_program(f, nargs::Int)
    begin
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
