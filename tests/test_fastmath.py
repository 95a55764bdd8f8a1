"""Host-side accuracy check of the branch-free FP64 elementary functions and the ψ polynomial tables of the
table-driven solve (numericalearth.jl_b200/csrc/ne_fastmath.cuh, ne_flux_tab.cuh).  The header compiles for the
host (MUFU seeds emulated in Float32); tools/fastmath_check.cu compares every function with long double on
400 000 samples and the ψ tables with the long-double closed forms of the Edson functions
(similarity_theory_turbulent_fluxes.jl:501-532, 586-618).  No GPU needed: nvcc builds a host executable."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def report(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("fm") / "fastmath_check")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                        os.path.join(ROOT, "tools", "fastmath_check.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return json.loads(out)


def test_elementary_functions_within_a_few_ulp(report):
    # units: 2^-53 relative (half an ulp of a number in [1, 2))
    assert report["rcp_ulp"] <= 2.0 and report["div_ulp"] <= 2.0 and report["sqrt_ulp"] <= 2.0
    assert report["cbrt_ulp"] <= 3.0 and report["cbrt_wide_ulp"] <= 3.0
    assert report["log_ulp_small"] <= 4.0 and report["log_ulp_large"] <= 4.0 and report["log_abs_near1_ulp1"] <= 3.0
    assert report["exp_ulp"] <= 3.0 and report["exp_ulp_mid"] <= 3.0


def test_short_exponential_and_saturated_constants(report):
    """exp_lo (32-entry 2^(j/32) table, degree 4) serves the roughness length ℓs that only scales ψ(ℓs/L★): 2e-12 relative is
    ample and is what it must deliver; the saturated exp(−ζmax) constants of the stable closed forms equal exp_mid's value."""
    assert report["exp_lo_rel"] <= 3e-12
    assert report["exp_sat_rel"] <= 1e-15 and report["z_sat_ok"] == 1
    # w = a|ζ| + b from the mantissa bits of |ζ| and the per-lane-class copy of the log table change no bit
    assert report["bit_w_and_replicated_log_same_bits"] == 1
    # one third-order step from the seed instead of two Newton steps (units 2^-53)
    assert report["rcp3_ulp"] <= 2.0 and report["sqrt3_ulp"] <= 2.5 and report["cbrt3_ulp"] <= 5.0


def test_psi_tables_reproduce_the_closed_forms(report):
    # error relative to max(1, |ψ|); the library refuses tables worse than 2e-15 and falls back to the closed forms
    assert report["psi_fit_err"] <= 1e-15
    assert report["psi_dense_err"] <= 1e-15
    assert report["psi_tiny_abs"] <= 1e-15
    assert report["psi_micro_abs"] <= 1e-15
    assert report["interval_logic_ok"] == 1


def test_far_unstable_closed_forms_with_branch_free_functions(report):
    """ζ ≤ −2^7 (the second trip of an unstable point): log_pos / sqrt_pos / cbrt_pos / atan_large against long double."""
    assert report["psi_far_err"] <= 2e-15
    assert report["atan_large_abs"] <= 4e-16


def test_general_psi_tables_sea_ice_and_large_yeager_pairs(report):
    """The same piecewise polynomials fitted to Split(SHEBA, Paulson) (the atmosphere-sea-ice default,
    similarity_theory_turbulent_fluxes.jl:779-789) and Split(LinearStable, Paulson) (:766-771)."""
    assert report["psi_seaice_general"] == 1
    assert report["psi_seaice_fit_err"] <= 1e-15 and report["psi_seaice_dense_err"] <= 1e-15
    assert report["psi_ly_fit_err"] <= 1e-15 and report["psi_ly_dense_err"] <= 1e-15


@pytest.mark.gpu
def test_device_functions_against_long_double(tmp_path):
    """The same functions as the kernels run them, with the MUFU seeds of the real hardware (tools/fastmath_gpu_check.cu):
    units 2^-53 relative.  The one-step third-order forms (rcp3 / sqrt3 / cbrt3) need seeds of ≥ 20 bits — this is where
    that assumption is checked."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fastmath_gpu_check")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tools", "fastmath_gpu_check.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    rep = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    print("DEVICE_FASTMATH", json.dumps(rep))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "fastmath_gpu_check.json"), "w") as f:
            json.dump(rep, f)
    assert rep["rcp_ulp"] <= 2.0 and rep["sqrt_ulp"] <= 2.0 and rep["cbrt_ulp"] <= 3.0
    assert rep["rcp3_ulp"] <= 2.0 and rep["sqrt3_ulp"] <= 2.5 and rep["cbrt3_ulp"] <= 5.0
    assert rep["log_ulp"] <= 4.0 and rep["log_rep_ulp"] <= 4.0 and rep["exp_ulp"] <= 3.0
    assert rep["exp_lo_rel"] <= 3e-12
