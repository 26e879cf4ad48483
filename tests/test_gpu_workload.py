"""BASELINE.json configs[0] (microgpt) and configs[1] (nanoGPT — the shape bench.py's headline is measured on): the
prove-shaped pipeline on the GPU against the CPU oracle twin, stage by stage —
every one-hot commitment, every sumcheck's final claims, the transcript state after each node (which pins every round
polynomial and challenge, since they are all absorbed), and the final HyperKZG opening.  Bit-exact."""
import numpy as np
import pytest

from oracle import cpu as ORC
from oracle import workload_cpu as WC
from tests.util import to_mont_array

pytestmark = pytest.mark.gpu
TAU = 0x1234567890abcdef1122334455667788


@pytest.mark.parametrize("config", ["microgpt", "nanoGPT"])
def test_pipeline_bit_exact(ctx, config):
    from jolt_atlas_b200 import SRS, MultilinearPolynomial, workload as W
    inputs = W.build_inputs(config)
    n = 1 << inputs["ell"]
    srs_host = ORC.srs_powers(to_mont_array([TAU])[0], n)
    srs = SRS(ctx, srs_host).precompute()
    got = W.run_device(ctx, srs, inputs)
    want = WC.run_cpu(srs_host, inputs)
    assert len(got["states"]) == len(want["states"]) == len(inputs["nodes"]) + 1
    for i, ((gc, gi), (wc, wi)) in enumerate(zip(got["commitments"], want["commitments"])):
        assert np.array_equal(np.asarray(gi, dtype=bool), np.asarray(wi, dtype=bool)), i
        assert np.array_equal(gc, wc), i
    for i, (a, b) in enumerate(zip(got["finals"], want["finals"])):
        assert np.array_equal(a, b), i
    assert got["states"] == want["states"]
    for k in ("com", "v", "w"):
        assert np.array_equal(got["open"][k], want["open"][k]), k
    # device-resident inputs give the same proof
    res = W.make_resident(ctx, inputs)
    again = W.run_device(ctx, srs, inputs, resident=res)
    assert again["states"] == want["states"]
    W.free_resident(res)
    srs.free()
