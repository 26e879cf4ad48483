"""jolt_atlas_b200 — B200-native (sm_100a) proving hot path of jolt-atlas behind a C ABI.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI + the C++ host driver) and
the thin Python mirror of the reference's polynomial / sumcheck / PCS interface (`api.py`).
Importing this package does not load CUDA; `api.Context()` does and fails loudly without a GPU.
"""
from .api import (  # noqa: F401
    BindingOrder, Blake2bTranscriptState, Context, HyperKZGOpening, hyperkzg_open, EqPolynomial, EvalKernel, GruenSplitEqPolynomial, JoltAtlasError,
    MsmWidth, MultilinearPolynomial, OneHotBatch, SRS, bind_many, g1_sum_indexed, g1_sum_indexed_batch, msm_fr, msm_fr_batch,
    msm_host, round_eval, sumcheck_prove, tensor_fold_i32, OneHotAddresses, InstanceKind, commit_one_hot_batches, eval_reduction_h, batched_sumcheck_prove, TensorI32, FusedWitness, PrefixSuffixShout, SuffixKind,
)
