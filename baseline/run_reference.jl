# run_reference.jl -- emits iteration logs of the REAL Enlsip.jl for the problems our oracle is checked on.
#
# NOT RUN in this repository's build image or on the GPU box: neither has a `julia` binary (SURVEY.md 8c), which is
# why the oracle is "parity unpinned".  On a machine with Julia >= 1.8:
#
#     julia --project=/path/to/Enlsip.jl baseline/run_reference.jl > baseline/reference_traces.json
#     python baseline/compare_reference.py baseline/reference_traces.json
#
# pins oracle/enlsip_oracle.py against the reference: exit status, iteration count, solution, objective, evaluation
# counts and the per-iteration record the reference keeps (DisplayedInfo, src/structures.jl:117-123: objective,
# ||active constraints||^2, ||p||, steplength, reduction).
using Enlsip, Printf

function dump(io, name, model)
    info = model.model_info
    its = info.iterations_detail
    print(io, "{\"name\": \"", name, "\", \"status\": \"", status(model), "\", \"objective\": ", @sprintf("%.17g", sum_sq_residuals(model)),
          ", \"solution\": [", join((@sprintf("%.17g", v) for v in solution(model)), ", "), "]",
          ", \"nb_function_evaluations\": ", info.nb_function_evaluations,
          ", \"nb_jacobian_evaluations\": ", info.nb_jacobian_evaluations, ", \"iterations\": [")
    for (k, d) in enumerate(its)
        k > 1 && print(io, ", ")
        print(io, "[", @sprintf("%.17g", d.objective), ", ", @sprintf("%.17g", d.sqr_nrm_act_cons), ", ", @sprintf("%.17g", d.nrm_p),
              ", ", @sprintf("%.17g", d.α), ", ", @sprintf("%.17g", d.reduction), "]")
    end
    print(io, "]}")
end

io = stdout
print(io, "[")

# --- Hock-Schittkowski 65 (README / test/problems/HS65.jl) ---
r65(x) = [x[1] - x[2]; (x[1] + x[2] - 10.0) / 3.0; x[3] - 5.0]
j65(x) = [1.0 -1.0 0.0; 1/3 1/3 0.0; 0.0 0.0 1.0]
c65(x) = [48.0 - x[1]^2 - x[2]^2 - x[3]^2]
a65(x) = [-2x[1] -2x[2] -2x[3]]
m = CnlsModel(r65, 3, 3; jacobian_residuals=j65, starting_point=[-5.0, 5.0, 0.0], ineq_constraints=c65, jacobian_ineqcons=a65,
              nb_ineqcons=1, x_low=[-4.5, -4.5, -5.0], x_upp=[4.5, 4.5, 5.0])
solve!(m)
dump(io, "hs65", m)

# --- chained Rosenbrock, n = 10 (test/problems/chained_rosenbrock.jl with n = 10) ---
n = 10
rcr(x) = vcat([10.0 * (x[i]^2 - x[i+1]) for i in 1:n-1], [x[k-n+1] - 1.0 for k in n:2(n-1)])
ccr(x) = [3x[k+1]^3 + 2x[k+2] - 5 + sin(x[k+1] - x[k+2]) * sin(x[k+1] + x[k+2]) + 4x[k+1] - x[k] * exp(x[k] - x[k+1]) - 3 for k in 1:n-2]
x0 = [(mod(i, 2) == 1 ? -1.2 : 1.0) for i in 1:n]
m = CnlsModel(rcr, n, 2(n - 1); starting_point=x0, eq_constraints=ccr, nb_eqcons=n - 2)
solve!(m)
print(io, ", ")
dump(io, "chained_rosenbrock10_forwarddiff", m)

println(io, "]")
