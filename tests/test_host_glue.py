"""Host-side round glue of the sumcheck engine (csrc/sumcheck_host.hpp, fr_host.hpp) checked on the CPU: the optimised forms
(small-integer Toom interpolation, Horner evaluation, batch inversion, inversion-free Gruen polynomials with the running
normalised claim) must give the same field elements as the plain forms they replace (unipoly.rs:104-134,219-245,
split_eq_poly.rs:379-471, mles_product_sum.rs:330-376)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_glue_matches_plain_forms(tmp_path):
    exe = str(tmp_path / "host_glue_check")
    src = os.path.join(ROOT, "tests", "host", "host_glue_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-mbmi2", "-madx", "-I", os.path.join(ROOT, "jolt_atlas_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "bad=0" in r.stdout
