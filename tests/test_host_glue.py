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


def test_transcript_blake2b_matches_hashlib(tmp_path):
    """Blake2b-256 of the product transcript (runtime-selected AVX2 rows and the unrolled scalar rounds) against hashlib
    on messages around the block boundaries, and the big-endian scalar serialisation (blake2b.rs:138-146)."""
    import hashlib
    exe = str(tmp_path / "blake2b_check")
    src = os.path.join(ROOT, "tests", "host", "blake2b_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-mbmi2", "-madx", "-I", os.path.join(ROOT, "jolt_atlas_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    for args in ([], ["scalar"]):
        out = subprocess.run([exe] + args, stdout=subprocess.PIPE, text=True, check=True).stdout
        n = 0
        for line in out.splitlines():
            p = line.split()
            if p[0] == "be32":
                assert p[1] == bytes(range(0x20, 0, -1)).hex()
                continue
            msg = bytes.fromhex(p[1]) if len(p) == 3 else b""
            assert len(msg) == int(p[0])
            assert hashlib.blake2b(msg, digest_size=32).hexdigest() == p[-1], (args, p[0])
            n += 1
        assert n == 11
