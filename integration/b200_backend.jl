# b200_backend.jl -- the binding a maintainer of carstenbauer/dqmc adds to run the local-update sweep, the Green's-function
# wrap and the UDT stabilization on a B200 through libdqmc_b200.so (C ABI: include/dqmc_b200.h).
#
#   include("b200_backend.jl")      after src/dqmc_framework.jl has been included (it needs AbstractDQMC, CBAssaad, Stack)
#
# The reference has no FFI; its seam is dispatch on the checkerboard tag of AbstractDQMC{C<:Checkerboard}
# (src/dqmc_framework.jl:4-13).  This file adds one tag, CBAssaadB200, and methods for the functions the driver calls
# (src/dqmc_framework.jl:168-172, 398-402, 500-517): initialize_stack, build_stack, propagate, local_updates, global_update,
# plus wrap_greens!, multiply_B_*, measure_tdgfs! for the measurement code.  Everything else (XML, lattice, checkpoints,
# driver loop) is untouched.  `julia` is not part of the build image of this repository, so the file is checked there only
# structurally (tests/test_cabi_cpu.py: every ccall symbol exists in the header with the same argument count).
#
# One new field is needed in Stack{G} (src/stack.jl:45-118):     b200::Ptr{Cvoid}
# and one line in DQMC(p) (src/dqmc_framework.jl:105-115):       CB = p.b200 ? CBAssaadB200 : CBAssaad

using Random

abstract type CBAssaadB200 <: CBAssaad end

const libdqmc = "libdqmc_b200"                      # dqmc_b200/libdqmc_b200.so on LD_LIBRARY_PATH

struct DqmcParamsC                                  # include/dqmc_b200.h: dqmc_params
  L::Int32; flv::Int32; opdim::Int32; slices::Int32; safe_mult::Int32; edrun::Int32
  all_checks::Int32; device::Int32; delay::Int32; reserved::Int32
  delta_tau::Float64; lambda::Float64; r::Float64; c::Float64; u::Float64
end

b200_error(ctx) = unsafe_string(ccall((:dqmc_last_error, libdqmc), Cstring, (Ptr{Cvoid},), ctx))
b200_check(mc, rc) = rc == 0 || error(b200_error(mc.s.b200))

# Parametric types are invariant: methods typed AbstractDQMC{CBAssaad} (slice_matrices.jl:101-203,
# hoppings_checkerboard.jl:20,65,141,165,200) do not apply to the new tag; the setup functions get forwarding methods.
init_checkerboard_matrices(mc::AbstractDQMC{CBAssaadB200}) =
  invoke(init_checkerboard_matrices, Tuple{AbstractDQMC{CBAssaad}}, mc)
init_checkerboard_matrices_Bfield(mc::AbstractDQMC{CBAssaadB200}) =
  invoke(init_checkerboard_matrices_Bfield, Tuple{AbstractDQMC{CBAssaad}}, mc)

# ---------------------------------------------------------------------------------------------- initialize_stack (stack.jl:224-242)
function b200_set_operator(mc, which::Integer, m)
  GC.@preserve m b200_check(mc, ccall((:dqmc_set_operator, libdqmc), Cint,
      (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint),
      mc.s.b200, which, m.m, m.n, m.colptr, m.rowval, m.nzval, eltype(m) <: Complex))
end

function initialize_stack(mc::AbstractDQMC{CBAssaadB200})
  p, l = mc.p, mc.l
  cp = Ref(DqmcParamsC(l.L, p.flv, p.opdim, p.slices, p.safe_mult, p.edrun, p.all_checks, 0, 0, 0,
                       p.delta_tau, p.lambda, p.r, p.c, p.u))
  ctx = Ref{Ptr{Cvoid}}(C_NULL)
  ccall((:dqmc_create, libdqmc), Cint, (Ptr{Ptr{Cvoid}}, Ptr{DqmcParamsC}), ctx, cp) == 0 || error(b200_error(C_NULL))
  mc.s.b200 = ctx[]
  # the sparse factors of the path (lattice.jl:24-47) go over as Julia's own CSC arrays (Int64, 1-based); 6 and 7 are the
  # half-step factors of group A that only effective_greens2greens! (fermion_measurements.jl:1125-1142) needs
  for (which, m) in ((0, l.chkr_hop_half[2]), (1, l.chkr_hop[1]), (2, l.chkr_hop_half_inv[2]),
                     (3, l.chkr_hop_inv[1]), (4, l.chkr_mu), (5, l.chkr_mu_inv),
                     (6, l.chkr_hop_half[1]), (7, l.chkr_hop_half_inv[1]))
    b200_set_operator(mc, which, m)
  end
  nb = l.neighbors
  GC.@preserve nb b200_check(mc, ccall((:dqmc_set_neighbors, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Int64}), mc.s.b200, nb))
  mc.s.greens = zeros(geltype(mc), p.flv * l.sites, p.flv * l.sites)   # host mirror, filled by sync_to_host!
  nothing
end

b200_destroy(mc) = ccall((:dqmc_destroy, libdqmc), Cint, (Ptr{Cvoid},), mc.s.b200)

# ---------------------------------------------------------------------------------------------- stack.jl:251-499
function build_stack(mc::AbstractDQMC{CBAssaadB200})
  h = mc.p.hsfield
  GC.@preserve h b200_check(mc, ccall((:dqmc_set_hsfield, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Float64}), mc.s.b200, h))
  b200_check(mc, ccall((:dqmc_build_stack, libdqmc), Cint, (Ptr{Cvoid},), mc.s.b200))
  mc.s.current_slice = mc.p.slices + 1
  mc.s.direction = -1
  nothing
end

function propagate(mc::AbstractDQMC{CBAssaadB200})
  s, d = Ref{Int32}(0), Ref{Int32}(0)
  b200_check(mc, ccall((:dqmc_propagate, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), mc.s.b200, s, d))
  mc.s.current_slice, mc.s.direction = s[], d[]
  nothing
end

function wrap_greens!(mc::AbstractDQMC{CBAssaadB200}, gf::Matrix{ComplexF64}, slice::Int, direction::Int)
  GC.@preserve gf b200_check(mc, ccall((:dqmc_wrap_greens, libdqmc), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Int32),
                                       mc.s.b200, gf, slice, direction))
  gf
end

function b200_multiply_B!(mc, op::Integer, slice::Int, M::Matrix{ComplexF64})
  GC.@preserve M b200_check(mc, ccall((:dqmc_multiply_B, libdqmc), Cint, (Ptr{Cvoid}, Cint, Int32, Ptr{ComplexF64}),
                                      mc.s.b200, op, slice, M))
  M
end
multiply_B_left!(mc::AbstractDQMC{CBAssaadB200}, slice::Int, M::Matrix{ComplexF64}) = b200_multiply_B!(mc, 0, slice, M)
multiply_B_right!(mc::AbstractDQMC{CBAssaadB200}, slice::Int, M::Matrix{ComplexF64}) = b200_multiply_B!(mc, 1, slice, M)
multiply_B_inv_left!(mc::AbstractDQMC{CBAssaadB200}, slice::Int, M::Matrix{ComplexF64}) = b200_multiply_B!(mc, 2, slice, M)
multiply_B_inv_right!(mc::AbstractDQMC{CBAssaadB200}, slice::Int, M::Matrix{ComplexF64}) = b200_multiply_B!(mc, 3, slice, M)
multiply_daggered_B_left!(mc::AbstractDQMC{CBAssaadB200}, slice::Int, M::Matrix{ComplexF64}) = b200_multiply_B!(mc, 4, slice, M)

function calculate_logdet(mc::AbstractDQMC{CBAssaadB200})
  v = Ref{Float64}(0.0)
  b200_check(mc, ccall((:dqmc_logdet, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Float64}), mc.s.b200, v))
  mc.s.log_det = v[]
end

# ---------------------------------------------------------------------------------------------- local_updates.jl:1-39
function local_updates(mc::AbstractDQMC{CBAssaadB200})
  N = mc.l.sites
  # the reference draws from the global MersenneTwister: opdim uniforms per proposal, one more only if p_acc <= 1
  # (local_updates.jl:9,31).  Generate the worst case from a *copy* of the RNG, let the library consume in that
  # order, then advance the real RNG by exactly `consumed` draws -> identical stream position as the CPU code.
  u = rand(copy(Random.GLOBAL_RNG), 4N)
  consumed, accepted, dS = Ref{Int64}(0), Ref{Int64}(0), Ref{Float64}(0.0)
  GC.@preserve u b200_check(mc, ccall((:dqmc_local_updates, libdqmc), Cint,
      (Ptr{Cvoid}, Float64, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
      mc.s.b200, mc.p.box, u, length(u), consumed, accepted, dS))
  for _ in 1:consumed[]
    rand()
  end
  mc.p.boson_action += dS[]
  return accepted[] / N
end

# fused fast path: `nupdates` x {propagate; local_updates} in one call (the body of dqmc_framework.jl:258-261)
function b200_sweep!(mc::AbstractDQMC{CBAssaadB200}, nupdates::Int)
  N = mc.l.sites
  u = rand(copy(Random.GLOBAL_RNG), 4N * nupdates)
  consumed, accepted, dS = Ref{Int64}(0), Ref{Int64}(0), Ref{Float64}(0.0)
  GC.@preserve u b200_check(mc, ccall((:dqmc_sweep, libdqmc), Cint,
      (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
      mc.s.b200, nupdates, mc.p.box, u, length(u), consumed, accepted, dS))
  for _ in 1:consumed[]
    rand()
  end
  mc.p.boson_action += dS[]
  s, d = Ref{Int32}(0), Ref{Int32}(0)
  ccall((:dqmc_get_state, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), mc.s.b200, s, d)
  mc.s.current_slice, mc.s.direction = s[], d[]
  return accepted[] / (N * nupdates)
end

# mc.s.greens / mc.p.hsfield are read by the driver only at measurement points (dqmc_framework.jl:432-439)
function sync_to_host!(mc::AbstractDQMC{CBAssaadB200})
  g, h = mc.s.greens, mc.p.hsfield
  GC.@preserve g b200_check(mc, ccall((:dqmc_get_greens, libdqmc), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), mc.s.b200, g))
  GC.@preserve h b200_check(mc, ccall((:dqmc_get_hsfield, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Float64}), mc.s.b200, h))
  nothing
end

# ---------------------------------------------------------------------------------------------- global_updates.jl:18-59
function global_update(mc::AbstractDQMC{CBAssaadB200})      # called at (slices, -1)
  u = rand(copy(Random.GLOBAL_RNG), 4)                        # 3 shift draws + the accept draw (used only if p_acc <= 1)
  Snew, acc, consumed = Ref{Float64}(0.0), Ref{Int32}(0), Ref{Int32}(0)
  GC.@preserve u b200_check(mc, ccall((:dqmc_global_update, libdqmc), Cint,
      (Ptr{Cvoid}, Float64, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
      mc.s.b200, mc.p.box_global, u, mc.p.boson_action, Snew, acc, consumed))
  for _ in 1:consumed[]
    rand()
  end
  acc[] == 1 && (mc.p.boson_action = Snew[])
  return Int(acc[])
end

# ---------------------------------------------------------------------------------------------- measurements
function measure_chi_dynamic(mc::AbstractDQMC{CBAssaadB200})    # boson_measurements.jl:6-56, device-resident field
  L, M = mc.l.L, mc.p.slices
  chi = zeros(Float64, div(L, 2) + 1, div(L, 2) + 1, div(M, 2) + 1)
  GC.@preserve chi b200_check(mc, ccall((:dqmc_measure_chi_dynamic, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Float64}), mc.s.b200, chi))
  chi
end

# time-displaced Green's functions stay on the device (2 M n^2 ComplexF64); the measurement code pulls the slices it needs
function measure_tdgfs!(mc::AbstractDQMC{CBAssaadB200})         # fermion_measurements.jl:1343-1407
  b200_check(mc, ccall((:dqmc_measure_tdgfs, libdqmc), Cint, (Ptr{Cvoid},), mc.s.b200))
end
function b200_tdgf(mc, which::Integer, slice::Int)               # which = 0: mc.s.meas.Gt0[slice], 1: G0t[slice]
  g = similar(mc.s.greens)
  GC.@preserve g b200_check(mc, ccall((:dqmc_get_tdgf, libdqmc), Cint, (Ptr{Cvoid}, Cint, Int32, Ptr{ComplexF64}),
                                      mc.s.b200, which, slice, g))
  g
end
deallocate_tdgfs_stacks!(mc::AbstractDQMC{CBAssaadB200}) = ccall((:dqmc_free_tdgfs, libdqmc), Cint, (Ptr{Cvoid},), mc.s.b200)

# the reference's "Propagation instability" telemetry (stack.jl:426,477): max |G_wrapped - G_fresh| since the last call
function b200_checks(mc)
  e, k = Ref{Float64}(0.0), Ref{Int64}(0)
  b200_check(mc, ccall((:dqmc_checks, libdqmc), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), mc.s.b200, e, k))
  e[] > 1e-7 && @printf("->%d \t+1 Propagation instability\t %.1e\n", mc.s.current_slice, e[])
  e[], k[]
end
