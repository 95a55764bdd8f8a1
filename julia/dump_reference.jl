# dump_reference.jl — run the REAL NumericalEarth.jl on the raw inputs of this repository's golden cases and write its
# outputs under the same .npz keys, so that the first machine with Julia turns "oracle-pinned" into "reference-pinned":
#
#     python tests/golden/export_inputs.py tests/golden/inputs            # raw .npy inputs + manifest.json (seeded, reproducible)
#     julia --project=/path/to/NumericalEarth.jl julia/dump_reference.jl tests/golden/inputs tests/golden/reference
#     NE_GOLDEN_DIR=tests/golden/reference python -m pytest tests/test_golden.py        # oracle and CUDA path vs the reference
#
# Needs NPZ.jl and JSON.jl next to NumericalEarth's own dependencies.  Julia is NOT installed in the container this
# repository was built in: the script has never been executed.  It only uses the reference's public constructors
# (PrescribedAtmosphere / PrescribedRadiation / PrescribedOcean-like ocean_simulation / OceanOnlyModel / OceanSeaIceModel,
# cf. test/test_reactant.jl:20-50 and test/test_surface_fluxes.jl:34-110) and reads the fields update_state! fills.
#
# Arrays arrive as this repository stores them: C-order (ny + 2hy, nx + 2hx) parents with halos, which NPZ.jl returns as
# column-major arrays of the same shape → `permutedims` gives Oceananigans' (nx + 2hx, ny + 2hy) parent layout.

using NPZ, JSON
using Oceananigans
using Oceananigans.Units
using Oceananigans.OutputReaders: Cyclical
using NumericalEarth
using NumericalEarth.EarthSystemModels.InterfaceComputations: interface_kernel_parameters

const FTs = Dict("f64" => Float64, "f32" => Float32)

parent2d(path) = permutedims(npzread(path), (2, 1))                 # (nx + 2hx, ny + 2hy)
parent4d(path) = begin                                              # (nt, ny + 2hy, nx + 2hx) → (nx + 2hx, ny + 2hy, 1, nt)
    a = permutedims(npzread(path), (3, 2, 1))
    reshape(a, size(a, 1), size(a, 2), 1, size(a, 3))
end

"copy a parent-with-halos array into a Field / FieldTimeSeries parent of the same shape"
function set_parent!(dst, src)
    p = parent(dst)
    size(p)[1:2] == size(src)[1:2] || error("halo mismatch: $(size(p)) vs $(size(src)) — build the grids with the manifest's halos")
    copyto!(p, reshape(convert(Array{eltype(p)}, src), size(p)))
    return nothing
end

function build_case(dir, case)
    ex, at = case["exchange"], case["atmosphere"]
    FT, FTa = FTs[ex["FT"]], FTs[at["FT"]]
    arch = CPU()
    Nz = isnothing(case["column"]) ? 1 : case["column"]["nz"]
    zfaces = isnothing(case["column"]) ? (-1, 0) : vcat(0.0, cumsum(npzread(joinpath(dir, "column_dz.npy")))) .- sum(npzread(joinpath(dir, "column_dz.npy")))
    grid = LatitudeLongitudeGrid(arch, FT; size = (ex["nx"], ex["ny"], Nz), halo = (ex["hx"], ex["hy"], max(Nz > 1 ? 3 : 1, 1)),
                                 longitude = Tuple(ex["longitude"]), latitude = Tuple(ex["latitude"]), z = zfaces)
    # land mask → immersed boundary whose inactive_node(i, j, Nz) is exactly the manifest's mask
    inactive = parent2d(joinpath(dir, "ocean_inactive.npy"))[ex["hx"]+1:ex["hx"]+ex["nx"], ex["hy"]+1:ex["hy"]+ex["ny"]] .!= 0
    bottom = ifelse.(inactive, 1.0, -sum(abs, zfaces isa Tuple ? zfaces : (first(zfaces),)))
    grid = ImmersedBoundaryGrid(grid, GridFittedBottom(bottom))

    ocean = ocean_simulation(grid; momentum_advection = nothing, tracer_advection = nothing, closure = nothing, bottom_drag_coefficient = 0)
    for (name, field) in ((:T, ocean.model.tracers.T), (:S, ocean.model.tracers.S), (:u, ocean.model.velocities.u), (:v, ocean.model.velocities.v))
        src = parent2d(joinpath(dir, "ocean_$name.npy"))
        if Nz == 1
            set_parent!(field, src)            # (1-level ocean: the parent is (nx + 2hx, ny + 2hy, 1 + 2hz); fill every level)
        end
        p = parent(field)
        for k in axes(p, 3)
            p[:, :, k] .= convert.(eltype(p), src[axes(p, 1), axes(p, 2)])
        end
    end
    if !isnothing(case["column"])               # 3-D T, S of the frazil kernels, k = 1 (bottom) … Nz (top)
        Hz = Oceananigans.Grids.halo_size(grid)[3]
        for (name, field) in ((:T, ocean.model.tracers.T), (:S, ocean.model.tracers.S))
            a = npzread(joinpath(dir, "column_$name.npy"))                  # (nz, ny + 2hy, nx + 2hx) C order
            a = permutedims(a, (3, 2, 1))
            parent(field)[:, :, Hz+1:Hz+Nz] .= convert.(eltype(field), a)
        end
    end

    atmos_grid = LatitudeLongitudeGrid(arch, FTa; size = (at["nx"], at["ny"]), halo = (at["hx"], at["hy"]), latitude = (-90, 90),
                                       longitude = (0, 360), topology = (Periodic, Bounded, Flat))
    times = convert.(FTa, at["times"])
    atmosphere = PrescribedAtmosphere(atmos_grid, times; surface_layer_height = at["surface_layer_height"],
                                      boundary_layer_height = at["boundary_layer_height"])
    series = Dict(:u => atmosphere.velocities.u, :v => atmosphere.velocities.v, :T => atmosphere.temperature,
                  :q => atmosphere.specific_humidity, :p => atmosphere.pressure,
                  :rain => atmosphere.precipitation_flux.rain, :snow => atmosphere.precipitation_flux.snow)
    for (name, fts) in series
        set_parent!(fts.data, parent4d(joinpath(dir, "atm_$name.npy")))
    end
    sw = FieldTimeSeries{Center, Center, Nothing}(atmos_grid, times; time_indexing = Cyclical())
    lw = FieldTimeSeries{Center, Center, Nothing}(atmos_grid, times; time_indexing = Cyclical())
    set_parent!(sw.data, parent4d(joinpath(dir, "atm_sw.npy")))
    set_parent!(lw.data, parent4d(joinpath(dir, "atm_lw.npy")))
    radiation = PrescribedRadiation(sw, lw)

    if case["sea_ice"]
        sea_ice = sea_ice_simulation(grid, ocean; advection = nothing, dynamics = nothing)
        for (name, field) in ((:concentration, sea_ice.model.ice_concentration), (:hi, sea_ice.model.ice_thickness),
                              (:top_temperature, sea_ice.model.ice_thermodynamics.top_surface_temperature))
            set_parent!(field, parent2d(joinpath(dir, "ocean_$name.npy")))
        end
        model = OceanSeaIceModel(ocean, sea_ice; atmosphere, radiation)
    else
        model = OceanOnlyModel(ocean; atmosphere, radiation)
    end
    return model
end

"exchange-layout (ny + 2hy, nx + 2hx) C-order array, as this repository's .npz files hold them"
out2d(field) = permutedims(Array(parent(field))[:, :, 1], (2, 1))

function dump_case(dir, case, t, outpath)
    model = build_case(dir, case)
    model.clock.time = t
    for c in (model.atmosphere, model.radiation)
        c.clock.time = t
    end
    Oceananigans.TimeSteppers.update_state!(model)
    I = model.interfaces
    out = Dict{String, Any}()
    ex = I.exchanger
    out["frac.i"], out["frac.j"] = out2d(ex.atmosphere.regridder.i), out2d(ex.atmosphere.regridder.j)
    for (k, f) in pairs(ex.atmosphere.state)
        out["atmos." * Dict(:u => "u", :v => "v", :T => "T", :p => "p", :q => "q", :Jʳⁿ => "Jrn", :Jˢⁿ => "Jsn")[k]] = out2d(f)
    end
    out["rad.sw"], out["rad.lw"] = out2d(ex.radiation.state.ℐꜜˢʷ), out2d(ex.radiation.state.ℐꜜˡʷ)
    ao = I.atmosphere_ocean_interface
    for n in (:latent_heat, :sensible_heat, :water_vapor, :x_momentum, :y_momentum, :friction_velocity, :temperature_scale, :water_vapor_scale)
        out["ao.$n"] = out2d(getproperty(ao.fluxes, n))
    end
    out["ao.interface_temperature"] = out2d(ao.temperature)
    net = I.net_fluxes.ocean
    for (k, n) in ((:u, "u"), (:v, "v"), (:T, "T"), (:S, "S"), (:η, "eta"), (:freshwater_heat_content, "freshwater_heat_content"))
        out["net_ocean.$n"] = out2d(getproperty(net, k))
    end
    rf = model.radiation.interface_fluxes.ocean
    for n in (:upwelling_longwave, :downwelling_longwave, :downwelling_shortwave)
        out["rad_ocean.$n"] = out2d(getproperty(rf, n))
    end
    if case["sea_ice"]
        ai, io = I.atmosphere_sea_ice_interface, I.sea_ice_ocean_interface
        for n in (:latent_heat, :sensible_heat, :water_vapor, :x_momentum, :y_momentum)
            out["asi.$n"] = out2d(getproperty(ai.fluxes, n))
        end
        out["asi.top_temperature"] = out2d(ai.temperature)
        for n in (:frazil_heat, :interface_heat, :salt, :freshwater, :x_momentum, :y_momentum)
            out["sio.$n"] = out2d(getproperty(io.fluxes, n))
        end
        top, bottom = I.net_fluxes.sea_ice.top, I.net_fluxes.sea_ice.bottom
        out["net_sea_ice.top_heat"], out["net_sea_ice.top_snowfall"] = out2d(top.heat), out2d(top.snowfall)
        out["net_sea_ice.top_u"], out["net_sea_ice.top_v"] = out2d(top.u), out2d(top.v)
        out["net_sea_ice.bottom_heat"] = out2d(bottom.heat)
    end
    # (iteration counts are not observable in the reference: `ao.iterations` / `asi.iterations` are left out and
    #  tests/test_golden.py skips keys the reference file does not hold)
    npzwrite(outpath, out)
    @info "wrote $outpath" keys = length(out)
end

function main(indir, outdir)
    manifest = JSON.parsefile(joinpath(indir, "manifest.json"))
    mkpath(outdir)
    for (name, case) in manifest["cases"]
        dump_case(joinpath(indir, name), case, manifest["T_STEP"], joinpath(outdir, case["outputs"]))
    end
end

length(ARGS) == 2 || error("usage: julia dump_reference.jl <inputs dir> <outputs dir>")
main(ARGS[1], ARGS[2])
