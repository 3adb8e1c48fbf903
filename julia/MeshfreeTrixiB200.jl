# MeshfreeTrixiB200.jl -- methods that give MeshfreeTrixi.jl's reserved `RBFFDEngineCUDA` engine a body by calling
# libmft_b200.so (C ABI: include/mft_b200.h).  `include` this file after `using MeshfreeTrixi`.
#
# NOT executed in the build container (no Julia toolchain there); the identical call sequence is exercised through
# ctypes by meshfreetrixi.jl_b200/api.py and the test-suite.  References are file:line in the MeshfreeTrixi.jl repository.

using MeshfreeTrixi
using MeshfreeTrixi: RBFSolver, RBFFDEngineCUDA, PointCloudDomain, PointData, RefPointData, SourceIGR, compute_flux_operator,
                     PositivityPreservingLimiterZhangShu,
                     SourceResidualViscosityTominec, SourceUpwindViscosityTominec,
                     SourceHyperviscosityTominec, SourceHyperviscosityFlyer,
                     BoundaryConditionDoNothing, boundary_condition_slip_wall, time_deriv_weights!
using Trixi
using Trixi: nvariables, BoundaryConditionDirichlet, CompressibleEulerEquations2D, LinearScalarAdvectionEquation2D
using StructArrays, SparseArrays, StaticArrays, SimpleUnPack

const libmft = get(ENV, "MFT_B200_LIB", joinpath(@__DIR__, "..", "meshfreetrixi.jl_b200", "libmft_b200.so"))

const MFT_MEM_HOST, MFT_MEM_DEVICE = Cint(0), Cint(1)

mft_error() = unsafe_string(ccall((:mft_last_error, libmft), Cstring, ()))
mft_check(rc) = rc == 0 ? nothing : error("libmft_b200 error $rc: ", mft_error())

mutable struct MftContext
    ptr::Ptr{Cvoid}
end

function MftContext(device::Integer, n_local::Integer, n_halo::Integer, nvars::Integer, k::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    mft_check(ccall((:mft_ctx_create, libmft), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64, Int64, Cint, Cint, Cint),
                    out, device, n_local, n_halo, nvars, 2, k))
    ctx = MftContext(out[])
    finalizer(c -> (c.ptr == C_NULL || ccall((:mft_ctx_destroy, libmft), Cint, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL), ctx)
    return ctx
end

function set_equation!(ctx::MftContext, eq::CompressibleEulerEquations2D)
    p = Float64[eq.gamma]
    mft_check(ccall((:mft_set_equation, libmft), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), ctx.ptr, 0, p, 1))
end
function set_equation!(ctx::MftContext, eq::LinearScalarAdvectionEquation2D)
    p = Float64[eq.advection_velocity[1], eq.advection_velocity[2]]
    mft_check(ccall((:mft_set_equation, libmft), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), ctx.ptr, 1, p, 2))
end

function add_boundary!(ctx::MftContext, kind::Integer, idx::Vector{Int}, normals::Vector{Float64}, values)
    vptr = values === nothing ? Ptr{Float64}(C_NULL) : pointer(values)
    GC.@preserve values mft_check(ccall((:mft_add_boundary, libmft), Cint,
                                        (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                                        ctx.ptr, kind, length(idx), idx, normals, vptr))
end

# value table of a Dirichlet group at time t, SoA: values[v*nb + j]  (PointCloudBCs.jl:49-63)
function dirichlet_table(bc::BoundaryConditionDirichlet, domain, tag, t, equations)
    nb, V = length(tag.idx), nvariables(equations)
    vals = Vector{Float64}(undef, nb * V)
    for (j, i) in enumerate(tag.idx)
        ub = bc.boundary_value_function(domain.pd.points[i], t, equations)
        for v in 1:V
            vals[(v - 1) * nb + j] = ub[v]
        end
    end
    return vals
end

function add_source!(ctx::MftContext, kind::Integer, params::Vector{Float64}, A::Union{Nothing, SparseMatrixCSC} = nothing)
    if A === nothing
        mft_check(ccall((:mft_add_source, libmft), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                        ctx.ptr, kind, params, length(params), C_NULL, C_NULL, C_NULL))
    else
        mft_check(ccall((:mft_add_source, libmft), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                        ctx.ptr, kind, params, length(params), A.colptr, A.rowval, A.nzval))
    end
end

# ---- Trixi.create_cache for the CUDA engine (CPU engine: src/solvers/pointcloudsolver/rbfsolver.jl:130-165) ----------
function Trixi.create_cache(domain::PointCloudDomain{2}, equations,
                            solver::RBFSolver{<:Any, RBFFDEngineCUDA}, RealT, uEltype)
    pd = domain.pd
    rbf_differentiation_matrices = compute_flux_operator(solver, domain)    # compute_operators.jl:409-453, unchanged
    ctx = MftContext(0, pd.num_points, 0, nvariables(equations), pd.num_neighbors)
    set_equation!(ctx, equations)
    x = Float64[p[1] for p in pd.points]
    y = Float64[p[2] for p in pd.points]
    perm = Vector{Int64}(undef, pd.num_points)
    mft_check(ccall((:mft_sfc_order, libmft), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                    pd.num_points, x, y, perm))
    mft_check(ccall((:mft_set_permutation, libmft), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx.ptr, perm))
    for (slot, A) in enumerate(rbf_differentiation_matrices)
        mft_check(ccall((:mft_set_operator_csc, libmft), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                        ctx.ptr, slot - 1, A.colptr, A.rowval, A.nzval))
    end
    return (; pd, rbf_differentiation_matrices, ctx, registered = Ref(false))
end

# Dirichlet closures get the stage time on every call in the reference (calc_single_boundary_flux!, rbfsolver.jl:311-316).
# Host-mode rhs! therefore re-evaluates every Dirichlet table at the actual `t` of the call (O(sqrt(N)) closure calls, small
# next to the state transfer).  Set STATIC_DIRICHLET[] = true to skip that when the data do not depend on time.
const STATIC_DIRICHLET = Ref(false)

# boundary conditions / sources are only known to rhs! (and to the history callback, whichever runs first); register them on
# first use, in NamedTuple order, with the Dirichlet data of the time of that first call
function register!(cache, domain, equations, boundary_conditions, source_terms, t0 = 0.0)
    ctx = cache.ctx
    for (key, bc) in zip(keys(boundary_conditions), boundary_conditions)     # rbfsolver.jl:280-285
        tag = domain.boundary_tags[key]
        nrm = collect(reinterpret(Float64, tag.normals))
        if bc isa BoundaryConditionDirichlet
            add_boundary!(ctx, 0, tag.idx, nrm, dirichlet_table(bc, domain, tag, t0, equations))
        elseif bc === boundary_condition_slip_wall
            add_boundary!(ctx, 1, tag.idx, nrm, nothing)
        else
            add_boundary!(ctx, 2, tag.idx, nrm, nothing)
        end
    end
    if source_terms !== nothing
        for source in values(source_terms)                                     # rbfsolver.jl:390-394
            c = source.cache
            if source isa SourceResidualViscosityTominec
                add_source!(ctx, 3, Float64[c.c_rv, c.c_uw, domain.pd.dx_avg, length(c.time_history) - 1])
            elseif source isa SourceUpwindViscosityTominec
                add_source!(ctx, 2, Float64[c.c_uw, domain.pd.dx_avg])
            elseif source isa SourceHyperviscosityTominec
                add_source!(ctx, 1, Float64[c.gamma], c.hv_differentiation_matrix)
            elseif source isa SourceHyperviscosityFlyer
                add_source!(ctx, 0, Float64[c.gamma], c.hv_differentiation_matrix)
            elseif source isa MeshfreeTrixi.SourceIGR                           # IGR.jl: maxiter = 20 is hard-wired (:190)
                add_source!(ctx, 4, Float64[c.alpha, 20.0])
            else
                error("source $(typeof(source)) has no CUDA implementation")
            end
        end
    end
    mft_check(ccall((:mft_finalize, libmft), Cint, (Ptr{Cvoid},), ctx.ptr))
    cache.registered[] = true
end

# time-dependent Dirichlet data: refresh the tables for stage time t (closures stay on the Julia side)
function refresh_dirichlet!(cache, domain, equations, boundary_conditions, t)
    for (g, (key, bc)) in enumerate(zip(keys(boundary_conditions), boundary_conditions))
        bc isa BoundaryConditionDirichlet || continue
        vals = dirichlet_table(bc, domain, domain.boundary_tags[key], t, equations)
        mft_check(ccall((:mft_update_boundary_values, libmft), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}),
                        cache.ctx.ptr, g - 1, vals))
    end
end

# ---- Trixi.rhs! (CPU engine: rbfsolver.jl:397-428) ---------------------------------------------------------------------
# Trixi's rhs!(du_ode, u_ode, semi, t) calls this 10-argument method positionally, so nothing here may depend on a keyword.
function Trixi.rhs!(du, u, t, domain, equations, initial_condition, boundary_conditions::BC,
                    source_terms::Source, solver::RBFSolver{<:Any, RBFFDEngineCUDA}, cache) where {BC, Source}
    if !cache.registered[]
        register!(cache, domain, equations, boundary_conditions, source_terms, t)   # tables of the first call's t
    elseif !STATIC_DIRICHLET[]
        refresh_dirichlet!(cache, domain, equations, boundary_conditions, t)
    end
    up = collect(Ptr{Float64}, pointer.(StructArrays.components(u)))
    dup = collect(Ptr{Float64}, pointer.(StructArrays.components(du)))
    GC.@preserve u du up dup begin
        mft_check(ccall((:mft_rhs, libmft), Cint, (Ptr{Cvoid}, Float64, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Cint),
                        cache.ctx.ptr, t, up, dup, MFT_MEM_HOST))
    end
    return nothing
end

# calc_fluxes! for the CUDA engine (CPU engine: rbfsolver.jl:247-265) -- used by test/divergence_test.jl
function MeshfreeTrixi.calc_fluxes!(du, u, domain::PointCloudDomain, have_nonconservative_terms::Trixi.False, equations,
                                    engine::RBFFDEngineCUDA, solver, cache)
    up = collect(Ptr{Float64}, pointer.(StructArrays.components(u)))
    dup = collect(Ptr{Float64}, pointer.(StructArrays.components(du)))
    GC.@preserve u du up dup mft_check(ccall((:mft_calc_fluxes, libmft), Cint,
                                             (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}), cache.ctx.ptr, up, dup))
end

# ---- HistoryCallback (history.jl:53-103) for the CUDA engine -----------------------------------------------------------------
# The stock callback runs update_history!(semi, u, t, approx_order, integrator) (history.jl:80-88), which calls the
# FIVE-argument modify_cache!(source, u, t, approx_order, integrator) of every source.  The stock method for
# SourceResidualViscosityTominec would update the HOST cache only, the device would never see a history push, success_iter
# would stay 0 there and every point would take the first-order branch of update_visc!.  So for a semidiscretization whose
# solver carries the CUDA engine, update_history! itself is routed here: the callback and the OrdinaryDiffEq driver stay as
# they are, the (polydeg+1)^2 weight solve stays in Julia (time_deriv_weights!, history.jl:131-152), the shift of the
# solution history and update_approx_du! happen on the device.
function MeshfreeTrixi.update_history!(semi::Trixi.SemidiscretizationHyperbolic, u, t, approx_order, integrator)
    if !(semi.solver isa RBFSolver{<:Any, RBFFDEngineCUDA})
        return invoke(MeshfreeTrixi.update_history!, Tuple{Any, Any, Any, Any, Any}, semi, u, t, approx_order, integrator)
    end
    cache = semi.cache
    # initialize! of the callback runs before the first rhs!: boundary conditions and sources must be registered (and the
    # layouts built) before the first push
    cache.registered[] || register!(cache, semi.mesh, semi.equations, semi.boundary_conditions, semi.source_terms, t)
    semi.source_terms === nothing && return nothing
    for source in values(semi.source_terms)
        source isa SourceResidualViscosityTominec && push_history!(source, u, t, approx_order, integrator, cache.ctx)
    end
    return nothing
end

function push_history!(source::SourceResidualViscosityTominec, u, t, approx_order, integrator, ctx::MftContext)
    @unpack time_history, time_weights = source.cache
    source.cache.success_iter .= integrator.success_iter
    time_history[2:end] .= time_history[1:(end - 1)]
    time_history[1] = t
    n = min(integrator.success_iter + 1, approx_order + 1)
    integrator.success_iter > 0 && time_deriv_weights!(@view(time_weights[1:n]), @view(time_history[1:n]))
    # the snapshot must be the u the callback received (integrator.u), not whatever the last rhs! left on the device: in
    # host mode the resident state is the last STAGE value, post-BC
    up = collect(Ptr{Float64}, pointer.(StructArrays.components(u)))
    GC.@preserve u up mft_check(ccall((:mft_upload_state, libmft), Cint, (Ptr{Cvoid}, Ptr{Ptr{Float64}}), ctx.ptr, up))
    mft_check(ccall((:mft_history_push_weights, libmft), Cint, (Ptr{Cvoid}, Float64, Int64, Cint, Ptr{Float64}),
                    ctx.ptr, t, integrator.success_iter, n, time_weights))
end

# Dirichlet tables for the two stage times of the next device-resident step (slot 0: t + dt, slot 1: t + dt/2); the
# reference's closures receive the stage time (rbfsolver.jl:311-316)
function stage_dirichlet!(cache, domain, equations, boundary_conditions, t, dt)
    for (g, (key, bc)) in enumerate(zip(keys(boundary_conditions), boundary_conditions))
        bc isa BoundaryConditionDirichlet || continue
        for (slot, ts) in ((0, t + dt), (1, t + dt / 2))
            vals = dirichlet_table(bc, domain, domain.boundary_tags[key], ts, equations)
            mft_check(ccall((:mft_set_stage_boundary_values, libmft), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}),
                            cache.ctx.ptr, g - 1, slot, vals))
        end
    end
end

# ---- device-resident SSPRK33 loop (same shape as Trixi's SimpleSSPRK33: init / step! / solve!) ------------------------------
function solve_ssprk33_resident!(u, semi, tspan, dt; approx_order = nothing, time_dependent_bcs = false)
    cache = semi.cache
    ctx = cache.ctx
    cache.registered[] || register!(cache, semi.mesh, semi.equations, semi.boundary_conditions, semi.source_terms, first(tspan))
    time_dependent_bcs && refresh_dirichlet!(cache, semi.mesh, semi.equations, semi.boundary_conditions, first(tspan))
    up = collect(Ptr{Float64}, pointer.(StructArrays.components(u)))
    GC.@preserve u up mft_check(ccall((:mft_upload_state, libmft), Cint, (Ptr{Cvoid}, Ptr{Ptr{Float64}}), ctx.ptr, up))
    t, iter = first(tspan), 0
    approx_order === nothing || mft_check(ccall((:mft_history_push, libmft), Cint, (Ptr{Cvoid}, Float64, Int64, Cint),
                                                ctx.ptr, t, 0, approx_order))
    while t < last(tspan) - 0.5dt
        time_dependent_bcs && stage_dirichlet!(cache, semi.mesh, semi.equations, semi.boundary_conditions, t, dt)
        mft_check(ccall((:mft_ssprk_step, libmft), Cint, (Ptr{Cvoid}, Cint, Float64, Float64), ctx.ptr, 0, t, dt))
        t += dt
        iter += 1
        approx_order === nothing || mft_check(ccall((:mft_history_push, libmft), Cint, (Ptr{Cvoid}, Float64, Int64, Cint),
                                                    ctx.ptr, t, iter, approx_order))
    end
    GC.@preserve u up mft_check(ccall((:mft_download_state, libmft), Cint, (Ptr{Cvoid}, Ptr{Ptr{Float64}}), ctx.ptr, up))
    return t, iter
end

# ---- multi-GPU bootstrap: one MPI rank per GPU; NCCL id broadcast over MPI (replaces MPICache, ParallelPointCloud.jl:6-71) ----
function init_comm!(ctx::MftContext, comm)   # comm::MPI.Comm
    id = zeros(UInt8, 128)
    MPI.Comm_rank(comm) == 0 && mft_check(ccall((:mft_nccl_unique_id, libmft), Cint, (Ptr{UInt8},), id))
    MPI.Bcast!(id, 0, comm)
    mft_check(ccall((:mft_comm_init, libmft), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}),
                    ctx.ptr, MPI.Comm_size(comm), MPI.Comm_rank(comm), id))
end

function set_halo!(ctx::MftContext, mpi_cache)   # fields of MPICache: send ids (1-based ranks), halo_send_idx, halo_recv_length
    peers = Cint.(mpi_cache.mpi_send_id .- 1)
    send_off = Int64[0; cumsum(length.(mpi_cache.halo_send_idx))]
    send_idx = Int64.(reduce(vcat, mpi_cache.halo_send_idx; init = Int64[]))
    recv = Int64.(mpi_cache.halo_recv_length)
    mft_check(ccall((:mft_set_halo, libmft), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                    ctx.ptr, length(peers), peers, send_off, send_idx, recv))
end


# ---- setup on the device (SURVEY.md section 8 row f1) ------------------------------------------------------------------------
# PointData constructor with the neighbour search on the GPU (CPU: geometry_primatives.jl:322-339)
function MeshfreeTrixi.PointData(medusa_data::Vector{SVector{2, Float64}}, basis::RefPointData, ::RBFFDEngineCUDA; device = 0)
    n, nv = length(medusa_data), basis.nv
    x = [p[1] for p in medusa_data]
    y = [p[2] for p in medusa_data]
    nbr = Matrix{Int64}(undef, nv, n)      # column-major nv x n is the library's n x nv row-major
    dist = Matrix{Float64}(undef, nv, n)
    mft_check(ccall((:mft_setup_knn, libmft), Cint,
                    (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Int64}, Ptr{Float64}),
                    device, n, x, y, nv, nbr, dist))
    return PointData{2, SVector{2, Float64}, Int}(medusa_data, [nbr[:, i] for i in 1:n], n, nv,
                                                   minimum(@view dist[2, :]), sum(@view dist[2, :]) / n)
end

phs_power(basis::RefPointData) = basis.approximation_type.Nrbf   # RBF{PolyharmonicSpline}.Nrbf: r^Nrbf (geometry_primatives.jl:97-114)

# compute_flux_operator with the per-point solves on the GPU (CPU: compute_operators.jl:409-453, k-th derivative :549-594)
function MeshfreeTrixi.compute_flux_operator(solver::RBFSolver{<:Any, RBFFDEngineCUDA}, domain::PointCloudDomain{2},
                                             k::Int = 1; device = 0)
    pd = domain.pd
    n, nv = pd.num_points, pd.num_neighbors
    x = [p[1] for p in pd.points]
    y = [p[2] for p in pd.points]
    nbr = reduce(hcat, pd.neighbors)
    wx = Matrix{Float64}(undef, nv, n)
    wy = similar(wx)
    mft_check(ccall((:mft_setup_rbf_weights, libmft), Cint,
                    (Cint, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Int64}, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
                    device, n, x, y, nv, nbr, phs_power(solver.basis), solver.basis.N, k, wx, wy))
    rows = repeat(1:n, inner = nv)
    return [sparse(rows, vec(nbr), vec(wx), n, n), sparse(rows, vec(nbr), vec(wy), n, n)]
end

# ---- Zhang-Shu positivity limiter (positivity_zhang_shu.jl:29-72, positivity_zhang_shu_point2d.jl:22-82) ------------------------
limiter_variable_code(::typeof(Trixi.density)) = Cint(0)    # MFT_VAR_DENSITY
limiter_variable_code(::typeof(Trixi.pressure)) = Cint(1)   # MFT_VAR_PRESSURE

function set_neighbors!(ctx::MftContext, pd::PointData)
    nbr = reduce(hcat, pd.neighbors)       # nv x n column-major == n x nv row-major, 1-based
    mft_check(ccall((:mft_set_neighbors, libmft), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx.ptr, nbr))
end

# limiter!(u_ode, integrator, semi, t) on host arrays (one pass per (threshold, variable) pair, in order)
function limiter_zhang_shu_device!(u, limiter::PositivityPreservingLimiterZhangShu, ctx::MftContext)
    thr = collect(Float64, limiter.thresholds)
    var = Cint[limiter_variable_code(v) for v in limiter.variables]
    up = collect(Ptr{Float64}, pointer.(StructArrays.components(u)))
    GC.@preserve u begin
        mft_check(ccall((:mft_limiter_zhang_shu, libmft), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Ptr{Float64}}, Cint),
                        ctx.ptr, length(thr), thr, var, up, 0 #= MFT_MEM_HOST =#))
    end
end

# SSPRK33(stage_limiter!) for the device-resident loop
function set_stage_limiter!(ctx::MftContext, limiter::PositivityPreservingLimiterZhangShu)
    thr = collect(Float64, limiter.thresholds)
    var = Cint[limiter_variable_code(v) for v in limiter.variables]
    mft_check(ccall((:mft_set_stage_limiter, libmft), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Cint}),
                    ctx.ptr, length(thr), thr, var))
end

# ---- SourceIGR (IGR.jl): registered like the other sources; sigma comes back through mft_get_field -----------------------------
register_igr!(ctx::MftContext, source::SourceIGR; maxiter = 20) = add_source!(ctx, 4 #= MFT_SRC_IGR =#, [source.cache.alpha, Float64(maxiter)])

function fetch_sigma!(source::SourceIGR, ctx::MftContext)
    mft_check(ccall((:mft_get_field, libmft), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.ptr, 7 #= MFT_FIELD_SIGMA =#, source.cache.sigma))
    return source.cache.sigma
end
