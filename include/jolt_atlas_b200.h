/*
 * jolt_atlas_b200.h — C ABI of the B200-native proving hot path of jolt-atlas.
 *
 * The reference (ICME-Lab/jolt-atlas) has no FFI; its seams are Rust traits.
 * Each entry point below states the reference interface it replaces (file:line under the
 * reference root).  `rust-shim/` and INTEGRATION.md show the binding a maintainer adds.
 *
 * Conventions
 *   Fr       : uint64_t[4] little-endian limbs, MONTGOMERY form (R = 2^256) — the in-memory
 *              layout of ark_bn254::Fr (BigInt<4>); the reference transmutes it the same way
 *              (joltworks/src/field/ark.rs:21-29).
 *   challenge: uint64_t[4] = {0, 0, lo, hi}, the 125-bit MontU128Challenge
 *              (joltworks/src/field/challenge/mont_ark_u128.rs:51-63).
 *   G1 affine: uint64_t[8] = Fq x[4], Fq y[4], Montgomery form; infinity is carried in a flag.
 *   order    : JA_LOW_TO_HIGH / JA_HIGH_TO_LOW == BindingOrder (multilinear_polynomial.rs).
 *   status   : every call returns 0 on success, <0 on error (ja_last_error has the text).
 *              Nothing unwinds across the ABI.  Prover-side callers panic on error (the
 *              reference prover panics on invariant violations); a verifier-side caller maps
 *              errors to ProofVerifyError::InternalError (joltworks/src/utils/errors.rs:16-50).
 *   memory   : host pointers are borrowed for the duration of the call; device objects are
 *              opaque handles owned by the library until the matching *_free.
 *   threads  : a ja_ctx serialises its calls internally (PCS::commit is invoked from rayon
 *              workers, jolt-atlas-core/src/onnx_proof/prover.rs:243-248).
 *   There is NO CPU fallback: without a CUDA device ja_init fails.
 */
#ifndef JOLT_ATLAS_B200_H
#define JOLT_ATLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ja_ctx ja_ctx;
typedef struct ja_poly ja_poly;       /* device-resident MultilinearPolynomial (dense Fr, or compact i32 before its first bind) */
typedef struct ja_spliteq ja_spliteq; /* device-resident GruenSplitEqPolynomial */
typedef struct ja_srs ja_srs;         /* device-resident KZG SRS (g1_powers) */
typedef struct ja_onehot ja_onehot;   /* device-resident batch of one-hot index lists (OneHotPolynomial::nonzero_indices as k*T+t) */
typedef struct ja_addr ja_addr;       /* device-resident one-hot ADDRESSES of a node: d lists x T entries in [0, K) (u32, 0xFFFFFFFF = None) */
typedef struct ja_hkzg ja_hkzg;       /* in-flight HyperKZG::open (folded polynomials resident on the device) */

enum { JA_LOW_TO_HIGH = 0, JA_HIGH_TO_LOW = 1 };

enum {
  JA_OK = 0,
  JA_ERR_CUDA = -1,          /* CUDA runtime failure (text in ja_last_error) */
  JA_ERR_INVALID = -2,       /* bad argument (null, non power of two, length mismatch) */
  JA_ERR_KEY_LENGTH = -3,    /* == ProofVerifyError::KeyLengthError: SRS shorter than the polynomial */
  JA_ERR_NO_DEVICE = -4,
  JA_ERR_UNSUPPORTED = -5
};

/* Round-evaluation bodies (SURVEY §8a addendum).  Family S = split-eq weighted, LowToHigh
 * (GruenSplitEqPolynomial::par_fold_out_in_unreduced, split_eq_poly.rs:569-597); family D = plain
 * products via sumcheck_evals at X in {0,2,3} (multilinear_polynomial.rs:873-905), HighToLow. */
enum {
  JA_EVAL_ADD = 0,      /* [l0+r0]                      ops/add.rs:283-296        n_out=1 */
  JA_EVAL_SUB = 1,      /* [l0-r0]                      ops/sub.rs:267            n_out=1 */
  JA_EVAL_MUL = 2,      /* [l0*r0, dl*dr]               ops/mul.rs:160-176        n_out=2 */
  JA_EVAL_SQUARE = 3,   /* [o0^2, do^2]                 ops/square.rs:163         n_out=2 */
  JA_EVAL_PROD = 4,     /* prod_i (p_i0 + X dp_i) on {1..d-1,inf} * current_scalar is applied by the caller;
                           mles_product_sum.rs:61-129   n_out=d (d = n_polys, 2..32) */
  JA_EVAL_POW = 5,      /* same-MLE power (p0 + X dp)^d on {1..d-1,inf}; d passed in aux_u32; cube.rs:159 */
  JA_EVAL_IDENT = 6,    /* [p0]  ps_shout / identity-RC cycle rounds, opening reduction; n_out=1 */
  JA_EVAL_IFF = 8,      /* [m0*a0+(1-m0)*b0, dm*da-dm*db]  polys (mask, a, b)          ops/iff.rs:189-216   n_out=2 */
  JA_EVAL_DIV = 9,      /* [r0*q0+R0-l0, dr*dq]            polys (l, r, q, R)          ops/div.rs:329-347   n_out=2 */
  JA_EVAL_RSQRT = 10,   /* [x0*quot0+dr0-S^3+gamma*(out0^2+sr0-quot0), dx*dquot+gamma*dout^2]  polys (x, quotient, output,
                           div_remainder, sqrt_remainder), aux_fr = {gamma, S^3}       ops/rsqrt.rs:390-418 n_out=2 */
  JA_EVAL_LIN3 = 11,    /* [tau*q0+r0-in0]  polys (input, quotient, remainder), aux_fr = {tau}
                           neural_teleport/division.rs:231-246; ScalarConstDiv's [l0-R0] (ops/scalar_const_div.rs:227) is JA_EVAL_SUB  n_out=1 */
  JA_EVAL_WIDENT = 12,  /* [p0 * T[g >> s]] split-eq weighted; polys (p, T), T a smaller table, aux_u32 = s
                           softmax_last_axis/recip_mult.rs:196-216 (phase 1)   n_out=1   (ja_round_eval only) */
  JA_EVAL_DOT2 = 16,    /* [sum l(0)r(0), sum l(2)r(2)]             einsum/dot.rs:292-303   n_out=2 */
  JA_EVAL_DOT3 = 17,    /* [sum l r e at 0,2,3], three MLEs of equal length  dot.rs:330-350   n_out=3 */
  JA_EVAL_SUM1 = 18,    /* [sum_i gamma_i * sum_j p_i[2j]]  LowToHigh one eval  hamming_weight.rs:118-139 (aux = gammas) */
  JA_EVAL_SUMHI = 19,   /* [sum_{j<n/2} p[j]]  HighToLow one eval  ops/sum/axis.rs:220-233 */
  /* Two-phase / table-weighted bodies: through ja_round_eval only (the caller owns the phase logic; the table or eq polynomial is the
   * LAST entry of polys and may be shorter than the operands; aux_u32 = s). */
  JA_EVAL_WSUM = 21,        /* [sum_g p[2g] T[g >> s]]  LowToHigh, degree 1   softmax_last_axis/exp_sum.rs:146-158   polys (p, T)   n_out=1 */
  JA_EVAL_WDOT2 = 22,       /* [sum_g T[g >> s] X(k) e(k), k=0,2,3]  LowToHigh  softmax_last_axis/max.rs:185-206   polys (X, e, T)  n_out=3 */
  JA_EVAL_DOT2_L2H = 23,    /* [sum a(0)b(0), sum a(2)b(2)]  LowToHigh   ops/slice.rs:254, reshape.rs:286, concat.rs:290, gather/mod.rs:232 (b = table + gamma*identity)  n_out=2 */
  JA_EVAL_SQ_EQHI = 24,     /* [sum_i l(k)^2 e(k), k=0,2,3]  HighToLow, e = eq polynomial at i >> s (length 1: its final claim)
                               ops/mean_of_squares.rs:363-386   polys (l, eq)   n_out=3 */
  JA_EVAL_DOT2_EQHI = 25,   /* [sum_i l(k) r(k) e(k)]  the same with two operands   ops/einsum/dot.rs:306-326 (EqSchedule::High)  polys (l, r, eq)  n_out=3 */
  JA_EVAL_DOT2_EQLOW = 26,  /* [sum_i l(k) r(k) T[i & (2^s-1)]]   ops/einsum/dot.rs:328-347 (EqSchedule::Low, rounds < log_k)  polys (l, r, T)  n_out=3 */
  JA_EVAL_OPEN = 20     /* dense opening reduction: [sum_{j<n/2} eq(r, j) P[j]] HighToLow with a HighToLow split-eq
                           (DensePolynomialProverOpening, subprotocols/opening_reduction.rs:355-419); n_out=1.
                           Only through ja_sumcheck_prove / ja_batched_sumcheck_prove. */
};

/* ---- context ---------------------------------------------------------------------------- */
int32_t ja_init(int32_t device, ja_ctx** out);
void ja_shutdown(ja_ctx*);
/* copies the calling thread's last error text (NUL terminated) */
void ja_last_error(char* buf, size_t cap);
int32_t ja_sync(ja_ctx*);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
uint64_t ja_launch_count(const ja_ctx*);

/* ---- polynomials: MultilinearPolynomial<Fr> (joltworks/src/poly/multilinear_polynomial.rs:22-35) ---- */
/* LargeScalars(DensePolynomial) from a host Fr array; n must be a power of two. */
int32_t ja_poly_from_fr(ja_ctx*, const uint64_t* z, size_t n, ja_poly** out);
/* I32Scalars(CompactPolynomial<i32>) — `MultilinearPolynomial::from(tensor.padded_next_power_of_two())`
 * (ops/mul.rs:146-147).  Stays 4 B/coeff on device until the first bind (compact_polynomial.rs:272-353). */
int32_t ja_poly_from_i32(ja_ctx*, const int32_t* z, size_t n, ja_poly** out);
/* `count` polynomials of n coefficients each from one row-major i32 matrix: one copy, one synchronisation (the operands of a node) */
int32_t ja_poly_from_i32_many(ja_ctx*, const int32_t* z, size_t count, size_t n, ja_poly** out);
/* RaPolynomial materialisation (joltworks/src/poly/ra_poly.rs:31-81; shout.rs:549-598): out[t] = table[idx[t]] with
 * table = K eq evaluations; idx[t] == 0xFFFFFFFF (None) -> 0.  n must be a power of two. */
int32_t ja_poly_from_lookup(ja_ctx*, const uint64_t* table, size_t K, const uint32_t* idx, size_t n, ja_poly** out);
/* uninitialised dense poly of length n (output of device-side producers) */
int32_t ja_poly_alloc(ja_ctx*, size_t n, ja_poly** out);
int32_t ja_poly_clone(ja_ctx*, const ja_poly*, ja_poly** out);
size_t ja_poly_len(const ja_poly*);
/* copy current (bound) coefficients to the host as Fr; cap >= len */
int32_t ja_poly_to_host(ja_ctx*, const ja_poly*, uint64_t* out, size_t cap);
void ja_poly_free(ja_ctx*, ja_poly*);
void ja_poly_free_many(ja_ctx*, ja_poly* const* polys, size_t n);   /* the d RA polynomials of a node in one call */

/* PolynomialBinding::bind_parallel (multilinear_polynomial.rs:728-742) ->
 * DensePolynomial::bound_poly_var_top_zero_optimized (dense_mlpoly.rs:126-141, HighToLow) /
 * bound_poly_var_bot_01_optimized (:219-239, LowToHigh) / CompactPolynomial::bind_parallel. */
int32_t ja_bind(ja_ctx*, ja_poly*, const uint64_t r[4], int32_t order);
/* several polys, one challenge, one launch (every `ingest_challenge`, e.g. ops/mul.rs:179-185) */
int32_t ja_bind_many(ja_ctx*, ja_poly* const* polys, size_t n_polys, const uint64_t r[4], int32_t order);
/* PolynomialBinding::final_claim (multilinear_polynomial.rs:744-762): len must be 1 */
int32_t ja_final_claim(ja_ctx*, const ja_poly*, uint64_t out[4]);
/* MultilinearPolynomial::evaluate (multilinear_polynomial.rs:766-862 -> dense_mlpoly.rs:265-305):
 * point = m challenges-or-field elements (Fr Montgomery limbs), r[0] binds the MSB. */
int32_t ja_poly_evaluate(ja_ctx*, const ja_poly*, const uint64_t* point, size_t m, uint64_t out[4]);

/* EqPolynomial::evals_with_scaling (joltworks/src/poly/eq_poly.rs:77-101,149-167,225-252):
 * table of 2^m Fr, big-endian index (r[0] = MSB); scale may be NULL (= 1). */
int32_t ja_eq_evals(ja_ctx*, const uint64_t* r, size_t m, const uint64_t* scale_or_null, ja_poly** out);

/* ---- GruenSplitEqPolynomial (joltworks/src/poly/split_eq_poly.rs:67-598) -------------------- */
/* new_with_scaling (:86-145): w = m Fr (Montgomery limbs; challenges are valid Fr limbs) */
int32_t ja_spliteq_new(ja_ctx*, const uint64_t* w, size_t m, int32_t order, const uint64_t* scale_or_null,
                       ja_spliteq** out);
/* bind (:331-372) */
int32_t ja_spliteq_bind(ja_ctx*, ja_spliteq*, const uint64_t r[4]);
/* get_current_scalar (:495) / get_current_w (:499-504) */
int32_t ja_spliteq_current_scalar(const ja_spliteq*, uint64_t out[4]);
int32_t ja_spliteq_current_w(const ja_spliteq*, uint64_t out[4]);
/* merge (:473-493): dense eq table of the unbound variables times current_scalar */
int32_t ja_spliteq_merge(ja_ctx*, const ja_spliteq*, ja_poly** out);
void ja_spliteq_free(ja_ctx*, ja_spliteq*);

/* One round's reduced sums (the host does the O(1) interpolation: gruen_poly_deg_2/3,
 * UniPoly::from_evals_and_hint, finish_mles_product_sum_from_evals).
 * Replaces the body of every `compute_message` listed in the enum above.
 * eq must be non-NULL for family S and NULL for family D.  out_evals = n_out Fr. */
int32_t ja_round_eval(ja_ctx*, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys,
                      const ja_spliteq* eq_or_null, const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32,
                      uint64_t* out_evals, size_t n_out);

/* Sumcheck::prove for ONE instance (joltworks/src/subprotocols/sumcheck.rs:565-599) with the library's Blake2b
 * transcript between the kernels: append_scalar(claim); per round compute_message (ja_round_eval + interpolation),
 * append the compressed round polynomial (unipoly.rs:550-558), r_j = challenge_scalar_optimized, claim = uni(r_j),
 * ingest_challenge (ja_bind_many + split-eq bind).  kind = JA_EVAL_*; family S/PROD/POW take eq_w = one Fr per round
 * (w of GruenSplitEqPolynomial::new, LowToHigh), the others NULL.  The polynomials are consumed (bound to length 1).
 * Outputs: out_coeffs[round][max_coeffs][4] (compressed: every coefficient but the linear one), out_ncoeffs[round],
 * out_challenges[round][4] ({0,0,lo,hi}), out_final_claims[n_polys][4]; transcript_state / n_rounds are read and
 * written back (Blake2bTranscript state + round counter, blake2b.rs:11-26). */
int32_t ja_sumcheck_prove(ja_ctx*, int32_t kind, ja_poly* const* polys, size_t n_polys, const uint64_t* eq_w, size_t eq_m,
                          const uint64_t* aux_fr, size_t n_aux, uint32_t aux_u32, const uint64_t claim[4],
                          uint8_t transcript_state[32], uint32_t* n_rounds, size_t max_coeffs, uint64_t* out_coeffs,
                          uint32_t* out_ncoeffs, uint64_t* out_challenges, uint64_t* out_final_claims);

/* BatchedSumcheck::prove (joltworks/src/subprotocols/sumcheck.rs:30-184, "front-loaded" batching): input claims are
 * appended, one batching coefficient per instance is drawn (challenge_vector), instances with fewer rounds start late
 * and contribute 2^k-scaled constants (:58-65, :91-100); per round the instances' univariates are combined, compressed,
 * appended, and r_j is drawn; every active instance then ingests r_j.  ONE host<->device exchange per round serves all
 * instances.  An instance is described by: */
enum { JA_INST_BOOLEANITY = 32,      /* BooleanitySumcheckProver (subprotocols/booleanity.rs:153-372), input claim 0 */
       JA_INST_HAMMING_TABLES = 33,  /* HammingWeightSumcheckProver over the K-entry G tables (hamming_weight.rs:60-160) */
       JA_INST_OPENING_ONEHOT = 34   /* OneHotPolynomialProverOpening (opening_reduction.rs:503-723) for the d polynomials of one
                                        address batch opened at one point: expands to d instances (d claims, d batching
                                        coefficients); addr, eq_w = r_cycle, aux_fr = r_address (log K), host_tables = the d
                                        input claims (table_len = 1), out_final_claims = d x 4 */ };
typedef struct ja_sc_instance {
  int32_t kind;                  /* JA_EVAL_* (device polynomials) or JA_INST_* */
  uint32_t aux_u32;              /* JA_EVAL_POW: degree d; JA_INST_BOOLEANITY: log_k */
  size_t n_polys;                /* polynomials / tables of the instance (booleanity: d) */
  ja_poly* const* polys;         /* JA_EVAL_*: device polynomials, consumed (bound to length 1) */
  const uint64_t* host_tables;   /* JA_INST_*: n_polys x table_len Fr — the G tables of compute_ra_evals */
  size_t table_len;              /* K */
  const ja_addr* addr;           /* JA_INST_BOOLEANITY: the addresses H_i = RaPolynomial(indices, F) is built from */
  const uint64_t* eq_w;          /* family S / PROD / POW: eq point (one Fr per round); booleanity: r_cycle */
  size_t eq_m;
  const uint64_t* aux_fr;        /* SUM1 / HAMMING_TABLES: gammas; booleanity: gammas (d) followed by r_address (log_k) */
  size_t n_aux;
  uint64_t claim[4];             /* input claim (ignored for booleanity) */
  uint64_t* out_final_claims;    /* n_polys x 4 Fr: final_claim of every polynomial / table (may be NULL) */
} ja_sc_instance;
/* Outputs as ja_sumcheck_prove; rounds = max over instances of num_rounds. */
int32_t ja_batched_sumcheck_prove(ja_ctx*, const ja_sc_instance* instances, size_t n_instances, uint8_t transcript_state[32],
                                  uint32_t* n_rounds, size_t max_coeffs, uint64_t* out_coeffs, uint32_t* out_ncoeffs,
                                  uint64_t* out_challenges);

/* NodeEvalReduction::prove -> compute_h (joltworks/src/subprotocols/evaluation_reduction.rs:91-148, :223-249): the
 * univariate h = mle o l for the degree n-1 curve l through the n opening points (points = n x m Fr, point-major, l(i) =
 * point i).  out_coeffs holds up to m (n-1) + 1 coefficients, trailing zeros trimmed (UniPoly::from_coeff); the caller
 * appends h ("UncompressedUniPoly_begin" ... "_end", unipoly.rs:540-548), draws x' and opens at l(x'). */
int32_t ja_eval_reduction_h(ja_ctx*, const ja_poly* mle, const uint64_t* points, size_t n, size_t m, uint64_t* out_coeffs,
                            size_t* out_ncoeffs);

/* Einsum operand fold / i32 tensor x eq-vector (ops/einsum/mk_kn_mn.rs:47-79):
 *   transpose==0: out[j] = sum_i from_i32(A[i*cols + j]) * eq[i]   (eq has `rows` entries, out has `cols`)
 *   transpose==1: out[i] = sum_j from_i32(A[i*cols + j]) * eq[j]   (eq has `cols` entries, out has `rows`) */
int32_t ja_tensor_fold_i32(ja_ctx*, const int32_t* A, size_t rows, size_t cols, const ja_poly* eq,
                           int32_t transpose, ja_poly** out);
/* The same fold on a tensor kept on the device: model weights (and the node inputs the tracer hands to several provers,
 * jolt-atlas-core/src/onnx_proof/ops/einsum/dot.rs:259-283) are uploaded once per proof / per preprocessing. */
typedef struct ja_tensor_i32 ja_tensor_i32;
int32_t ja_tensor_i32_upload(ja_ctx*, const int32_t* A, size_t rows, size_t cols, ja_tensor_i32** out);
void ja_tensor_i32_free(ja_ctx*, ja_tensor_i32*);
int32_t ja_tensor_fold_resident(ja_ctx*, const ja_tensor_i32* t, const ja_poly* eq, int32_t transpose, ja_poly** out);

/* ---- SRS residency + MSM (commitment half) ---------------------------------------------------------
 * Scalar width tags of ja_msm_host == the variants VariableBaseMSM::msm dispatches on (joltworks/src/msm/mod.rs:27-181). */
enum { JA_MSM_FR = 0, JA_MSM_U8 = 1, JA_MSM_U16 = 2, JA_MSM_U32 = 3, JA_MSM_U64 = 4, JA_MSM_I32 = 5, JA_MSM_I64 = 6 };
/* Upload g1_powers once per ProverSetup (hyperkzg/commitment_scheme.rs:36-44 setup_prover -> kzg.rs:108-143
 * KZGProverKey); the Rust shim repacks ark's G1Affine {x, y, infinity} into x||y Montgomery limbs. */
int32_t ja_srs_upload(ja_ctx*, const uint64_t* g1_affine_xy, size_t n_points, ja_srs** out);
/* SRS::setup's fixed-base loop on device (hyperkzg/kzg.rs:45-66): g1_powers[i] = beta^i * g1.  The caller samples
 * beta and g1 (arkworks UniformRand over ChaCha20 stays on the host, kzg.rs:34-36). */
int32_t ja_srs_generate(ja_ctx*, const uint64_t g1_xy[8], const uint64_t beta[4], size_t n_points, ja_srs** out);
/* Fixed-base window table for full-width MSMs against this SRS: 2^(16 w) * g1_powers[i] for w = 0..15 resident in HBM
 * (16x the SRS: 256 MB at nanoGPT scale, 16 GiB at GPT-2 scale).  Every later MSM of Fr scalars then fills ONE bucket set
 * (no per-window bucket reduction, no doubling tail).  Part of setup_prover; results are unchanged (same group element). */
int32_t ja_srs_precompute(ja_ctx*, ja_srs*);
int32_t ja_srs_to_host(ja_ctx*, const ja_srs*, size_t first, size_t count, uint64_t* out_xy);
size_t ja_srs_len(const ja_srs*);
void ja_srs_free(ja_ctx*, ja_srs*);
/* UnivariateKZG::commit_as_univariate on a device-resident dense polynomial (kzg.rs:285-298 ->
 * VariableBaseMSM::msm LargeScalars, msm/mod.rs:32-37): sum_i Z[i] * g1_powers[i].
 * JA_ERR_KEY_LENGTH when the SRS is shorter than the polynomial (ProofVerifyError::KeyLengthError). */
int32_t ja_msm_fr(ja_ctx*, const ja_srs*, const ja_poly* scalars, uint64_t out_xy[8], int32_t* is_inf);
/* UnivariateKZG::commit_variable_batch / batch_msm (kzg.rs:227-243, msm/mod.rs:309-318): `count` polynomials of
 * any lengths against the same SRS prefix, ONE bucket pipeline for the whole batch.  out_xy = count x 8 limbs. */
int32_t ja_msm_fr_batch(ja_ctx*, const ja_srs*, const ja_poly* const* polys, size_t count, uint64_t* out_xy,
                        int32_t* is_inf);
/* MSM over host scalars of any width (HyperKZG::commit of compact polynomials, commitment_scheme.rs:54-73):
 * scalars = n elements of the type named by width_tag (JA_MSM_FR: n x 4 Montgomery limbs); bases are
 * g1_powers[base_offset .. base_offset + n).  Signed widths reproduce msm(pos) - msm(neg) (msm/mod.rs:93-176). */
int32_t ja_msm_host(ja_ctx*, const ja_srs*, size_t base_offset, const void* scalars, int32_t width_tag, size_t n,
                    uint64_t out_xy[8], int32_t* is_inf);
/* HyperKZG::commit_one_hot (hyperkzg/mod.rs:520-554 -> jolt_optimizations::batch_g1_additions_multi):
 * sum of g1_powers[indices[i]]; the caller forms indices k*T + t exactly as the reference does. */
int32_t ja_g1_sum_indexed(ja_ctx*, const ja_srs*, const uint64_t* indices, size_t n, uint64_t out_xy[8],
                          int32_t* is_inf);
/* batch_commit_one_hot (hyperkzg/mod.rs:558-596): `count` index lists, list i = indices[offsets[i] .. offsets[i+1]). */
int32_t ja_g1_sum_indexed_batch(ja_ctx*, const ja_srs*, const uint64_t* indices, const uint64_t* offsets, size_t count,
                                uint64_t* out_xy, int32_t* is_inf);

/* The same with the index lists resident on the device (uploaded once per proof; the opening reduction reuses them). */
int32_t ja_onehot_upload(ja_ctx*, const uint64_t* indices, const uint64_t* offsets, size_t count, ja_onehot** out);
int32_t ja_onehot_commit(ja_ctx*, const ja_srs*, const ja_onehot*, uint64_t* out_xy, int32_t* is_inf);
void ja_onehot_free(ja_ctx*, ja_onehot*);

/* One-hot address batches: the d chunk-address lists of a node (witness.rs:84-99 build_one_hot_rad_witness ->
 * OneHotPolynomial::from_indices, one_hot_polynomial.rs:62) uploaded once as d x T u32 (0xFFFFFFFF = None, K <= 65536),
 * then consumed on the device by the commitment, the RA materialisations and the G-table scatter. */
int32_t ja_addr_upload(ja_ctx*, const uint32_t* k, size_t d, size_t T, size_t K, ja_addr** out);
/* n batches in one call (all index arrays of a proof: witness.rs:142-214 produces them together): copies and device-side
 * validation enqueued back to back, one synchronisation; on error no handle is returned. */
int32_t ja_addr_upload_many(ja_ctx*, const uint32_t* const* ks, const size_t* ds, const size_t* Ts, const size_t* Ks, size_t n,
                            ja_addr** outs);
void ja_addr_free(ja_ctx*, ja_addr*);
size_t ja_addr_len(const ja_addr*);
size_t ja_addr_count(const ja_addr*);
/* HyperKZG::batch_commit_one_hot (hyperkzg/mod.rs:558-596): C_i = sum_t g1_powers[k_i[t] * T + t]; out_xy = d x 8 limbs */
int32_t ja_addr_commit(ja_ctx*, const ja_srs*, const ja_addr*, uint64_t* out_xy, int32_t* is_inf);
/* commit_to_polynomials over every one-hot polynomial of a proof (jolt-atlas-core/src/onnx_proof/prover.rs:236-249): the
 * lists of all `n_batches` address batches in ONE pair of launches; out_xy = (sum of the batches' d) x 8 limbs, batch-major. */
int32_t ja_addr_commit_many(ja_ctx*, const ja_srs*, const ja_addr* const* batches, size_t n_batches, uint64_t* out_xy,
                            int32_t* is_inf);
/* RaPolynomial::new(indices, eq_evals) for all d lists in one launch (ra_poly.rs:31-81, ra_virtual.rs:113-134):
 * tables = d x K Fr, out_polys[i][t] = tables[i][k_i[t]] (None -> 0); T must be a power of two. */
int32_t ja_addr_gather(ja_ctx*, const ja_addr*, const uint64_t* tables, ja_poly** out_polys);
/* compute_ra_evals (subprotocols/shout.rs:549-598): out_G[i][k] = sum_{t: k_i[t] == k} eq(r_cycle, t); out_G = d x K Fr */
int32_t ja_addr_ra_evals(ja_ctx*, const ja_addr*, const uint64_t* r_cycle, size_t log_t, uint64_t* out_G);
/* The same for many address batches whose points are all known at once (OneHotPolynomialProverOpening::initialize for every
 * committed polynomial of the opening reduction, joltworks/src/subprotocols/opening_reduction.rs:532-571): one
 * synchronisation for all of them.  out_G[j] receives d_j x K_j Fr. */
int32_t ja_addr_ra_evals_many(ja_ctx*, const ja_addr* const* addrs, const uint64_t* const* r_cycles, const size_t* log_ts, size_t n,
                              uint64_t* const* out_G);

/* build_materialized_rlc (joltworks/src/poly/rlc_polynomial.rs:13-78): the joint polynomial sum_i gamma^i P_i of the single
 * HyperKZG opening, built in HBM.  ja_poly_zeros(2^max_num_vars); one ja_rlc_add_onehot per address batch
 * (joint[k_i[t] * T + t] += coeffs[i], :59-74) and one ja_rlc_add_dense per dense polynomial (joint[i] += coeff * P[i], :42-57). */
int32_t ja_poly_zeros(ja_ctx*, size_t n, ja_poly** out);
int32_t ja_rlc_add_onehot(ja_ctx*, ja_poly* joint, const ja_addr*, const uint64_t* coeffs /* d Fr */);
int32_t ja_rlc_add_dense(ja_ctx*, ja_poly* joint, const ja_poly* poly, const uint64_t coeff[4]);

/* ---- prefix-suffix Shout: the T-sized passes of the read-raf sumcheck over a 2^LOG_K-entry table (clamp lookups: LOG_K = 64) ----
 * joltworks/src/subprotocols/ps_shout/mod.rs.  The LOG_K address rounds run in NUM_PHASES phases over m = 2^(LOG_K / NUM_PHASES)-entry
 * suffix polynomials: O(m) work per round, done by ja_psshout_prove_address on the HOST next to the transcript (a device round would
 * pay a PCIe round trip for ~1 us of arithmetic on 256-entry tables).  The device does what scales with T at every phase boundary:
 *   ja_psshout_new          lookup_indices (T x u64) resident; u_evals = EqPolynomial::evals(r_node_output)        mod.rs:226-267
 *   ja_psshout_init_phase   init_phase (:269-303): u_evals[j] *= v[phase-1][k_bound(j)] (v_prev = the m-entry expanding table of the
 *                           finished phase, NULL for phase 0), then init_suffix_polys (:305-335) / RafProverState::init_Q
 *                           (poly/prefix_suffix.rs:294-351): out_Q[s][y] = sum_j u_evals[j] * suffix_s(suffix_bits(k_j)) over the entries
 *                           with prefix_bits(k_j) & (m-1) == y, for the n_suffixes suffix kinds listed (read-checking and raf suffixes in
 *                           one pass).  Suffix kinds = the clamp-table family of lookup_tables/suffixes/ with `bound` = BOUND of the
 *                           table (31 for SaturationTable, 9 for the ONNX Clamp) and the identity suffix of the raf decomposition.
 *   ja_psshout_materialize_ra   init_log_t_rounds (:420-446): ra[j] = prod_phase v[phase][k_bound(j, phase)], v = NUM_PHASES x m Fr;
 *                           the result is the polynomial of the log T cycle rounds (JA_EVAL_IDENT with eq = r_node_output).  v = NULL
 *                           uses the tables of the last ja_psshout_prove_address; `scale` (NULL = 1) multiplies ra by the constant
 *                           val + raf_val of the cycle rounds (mod.rs:484-487), so that JA_EVAL_IDENT over the result emits the
 *                           reference's gruen_poly_deg_2(eval_at_0 * (val + raf_val), claim); its final claim is scale * ra(r).
 *   ja_psshout_prove_address    the LOG_K address rounds of Sumcheck::prove over ReadRafSumcheckProver<SaturationTable-style clamp,
 *                           UnaryRafPS> (subprotocols/sumcheck.rs:565-599; mod.rs:337-418 compute_prefix_suffix_prover_message,
 *                           :491-560 ingest_challenge; lookup_tables/clamp.rs:94-118 combine with prefixes/{higher_all_zero,
 *                           higher_all_one,lower_word,msb}.rs; ps_shout/unary.rs:45-87 + poly/signed_identity_poly.rs:183-217 for the
 *                           raf part): runs init_phase on the device at the 8 phase boundaries and the rounds in between on the
 *                           host, appends every compressed round polynomial [c0, c2] to the transcript, draws the challenges.
 *                           claim_in = rv_claim + gamma * operand_claim (params.input_claim, mod.rs:108-110); NULL = use the sum the
 *                           prover derives from its own phase-0 tables, also returned in out_input_claim.  out_val / out_raf_val =
 *                           val / raf_val of mod.rs:527-556; out_claim = the running claim handed to the cycle rounds. */
typedef struct ja_psshout ja_psshout;
enum { JA_SUF_ONE = 0,               /* suffixes/one.rs */
       JA_SUF_HIGHER_ALL_ZERO = 1,   /* suffixes/higher_all_zero.rs:9-29 */
       JA_SUF_HZERO_MUL_LWORD = 2,   /* suffixes/hzero_mul_lword.rs:9-36 */
       JA_SUF_HONE_MUL_LWORD = 3,    /* suffixes/hone_mul_lword.rs:10-37 */
       JA_SUF_IDENTITY = 4,          /* poly/identity_poly.rs:153-158 (IdentityPolynomial as a SuffixPolynomial: the suffix value) */
       JA_SUF_SHIFT = 5 };           /* poly/identity_poly.rs:160-166 (ShiftSuffixPolynomial: 2^suffix_len) */
int32_t ja_psshout_new(ja_ctx*, const uint64_t* lookup_indices, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                       uint32_t phases, ja_psshout** out);
int32_t ja_psshout_init_phase(ja_ctx*, ja_psshout*, uint32_t phase, const uint64_t* v_prev, const uint32_t* suffix_kinds, size_t n_suffixes,
                              uint32_t bound, uint64_t* out_Q /* n_suffixes x m Fr */);
int32_t ja_psshout_materialize_ra(ja_ctx*, ja_psshout*, const uint64_t* v /* phases x m Fr, or NULL */, const uint64_t* scale /* Fr or NULL */,
                                  ja_poly** out_ra);
int32_t ja_psshout_prove_address(ja_ctx*, ja_psshout*, uint32_t bound, const uint64_t* gamma, const uint64_t* claim_in /* or NULL */,
                                 uint8_t transcript_state[32], uint32_t* transcript_n_rounds, uint64_t* out_coeffs /* LOG_K x 2 Fr */,
                                 uint32_t* out_ncoeffs /* LOG_K */, uint64_t* out_challenges /* LOG_K x 4 limbs */,
                                 uint64_t* out_input_claim, uint64_t* out_val, uint64_t* out_raf_val, uint64_t* out_claim);
/* IdentityRCProver (subprotocols/identity_range_check.rs:140-325): the LOG_K address rounds of a remainder range check over the
 * unsigned identity decomposition (suffixes [Shift, Identity], poly/identity_poly.rs:113-166); phases per IdentityRCProvider::phases
 * (:416-431: LOG_K / 4 if LOG_K % 4 == 0, else LOG_K / 2).  out_raf_val = the identity checkpoint (the constant of the cycle
 * rounds: ja_psshout_materialize_ra(scale = raf_val) then JA_EVAL_IDENT). */
int32_t ja_psshout_prove_identity_rc(ja_ctx*, ja_psshout*, const uint64_t* claim_in /* or NULL */, uint8_t transcript_state[32],
                                     uint32_t* transcript_n_rounds, uint64_t* out_coeffs /* LOG_K x 2 Fr */, uint32_t* out_ncoeffs,
                                     uint64_t* out_challenges, uint64_t* out_input_claim, uint64_t* out_raf_val, uint64_t* out_claim);
int32_t ja_psshout_tables(ja_ctx*, ja_psshout*, uint64_t* out_v /* phases x m Fr: the expanding tables of the address rounds */);
void ja_psshout_free(ja_ctx*, ja_psshout*);

/* ---- witness generation of a fused node on the device (jolt-atlas-core/src/onnx_proof/witness.rs:142-214 generate_node_witnesses) ----
 * From the node's resident i32 operands: acc = einsum_acc_i64 (op EINSUM_MK_KN: A m x k, B k x n; atlas-onnx-tracer/src/ops/einsum.rs:248-258),
 * a * b (MUL, ops/mul.rs:51-60) or a +- b (ADD / SUB, sat_binop_intermediate) in i64; quotient = acc.div_euclid(2^S), remainder =
 * acc.rem_euclid(2^S) (ops/mod.rs:224-249); then, zero-padded to T entries, the 16 ClampRaD chunk lists (4-bit chunks of the
 * quotient as u64, clamp_lookups/mod.rs:245-252 + joltworks/src/config.rs:75-77), the ceil(S/4) RescaleRemainderRaD chunk lists
 * (witness.rs:601-617), the lookup indices of the clamp read-raf and the clamped i32 output - all born in HBM, ready for
 * ja_addr_commit_many / ja_addr_ra_evals / ja_addr_gather / ja_psshout_from_witness.  The address batches are owned by the witness. */
typedef struct ja_witness ja_witness;
enum { JA_WIT_EINSUM_MK_KN = 0, JA_WIT_MUL = 1, JA_WIT_ADD = 2, JA_WIT_SUB = 3 };
int32_t ja_witness_fused(ja_ctx*, int32_t op, const ja_tensor_i32* A, const ja_tensor_i32* B, uint32_t scale_bits, size_t T, ja_witness** out);
const ja_addr* ja_witness_clamp_addr(const ja_witness*);      /* d = 16, K = 16 */
const ja_addr* ja_witness_rem_addr(const ja_witness*);        /* d = ceil(S / 4), K = 16; NULL when scale_bits == 0 */
int32_t ja_witness_to_host(ja_ctx*, const ja_witness*, uint64_t* out_idx /* T */, int32_t* out_i32 /* T */, uint32_t* out_clamp_k /* 16 x T */,
                           uint32_t* out_rem_k /* ceil(S/4) x T */);    /* any pointer may be NULL */
int32_t ja_psshout_from_witness(ja_ctx*, const ja_witness*, const uint64_t* r_cycle, size_t log_t, uint32_t log_k, uint32_t phases, ja_psshout** out);
/* ... and over the rescale remainders (LOG_K = scale_bits): the state of the remainder range check */
int32_t ja_psshout_from_witness_rem(ja_ctx*, const ja_witness*, const uint64_t* r_cycle, size_t log_t, uint32_t phases, ja_psshout** out);
int32_t ja_psshout_new_dev(ja_ctx*, const unsigned long long* d_indices, size_t T, const uint64_t* r_cycle, size_t log_t, uint32_t log_k,
                           uint32_t phases, ja_psshout** out);          /* lookup indices already in device memory */
void ja_witness_free(ja_ctx*, ja_witness*);

/* ---- HyperKZG::open (joltworks/src/poly/commitment/hyperkzg/mod.rs:400-447) --------------------------------------
 * Split at the two transcript interaction points so that a Rust caller keeps its own Blake2bTranscript:
 *   begin    Phase 1: l-1 folds Pi[j] = point[l-i-1]*(prev[2j+1]-prev[2j]) + prev[2j] (:413-428) and
 *            commit_variable_batch(polys[1..]) (kzg.rs:227-243) -> com (l-1 points).  point = l challenges {0,0,lo,hi}.
 *            caller: transcript.append_points(com); r = transcript.challenge_scalar()            (:439-440)
 *   evals    v[i][j] = eval_as_univariate(polys[j], u[i]), u = [r, -r, r^2] (:441, :245-257); v_out = 3 x l Fr, point-major
 *            caller: transcript.append_scalars(v); q_powers = transcript.challenge_scalar_powers(l)   (:258-260)
 *   witness  B = sum_k q^k polys[k] (:262-270), h_i = (B - B(u_i))/(x - u_i) (:213-229), w = commit_batch(h) (3 points)
 *            caller: transcript.append_points(w); transcript.challenge_scalar()                   (:276-277)
 * Errors: JA_ERR_KEY_LENGTH (SRS too short), JA_ERR_INVALID (length != 2^l, as the reference's assert_eq!). */
int32_t ja_hyperkzg_open_begin(ja_ctx*, const ja_srs*, const ja_poly* poly, const uint64_t* point, size_t ell,
                               ja_hkzg** out, uint64_t* com_xy, int32_t* com_inf);
int32_t ja_hyperkzg_open_evals(ja_ctx*, ja_hkzg*, const uint64_t r[4], uint64_t* v_out);
int32_t ja_hyperkzg_open_witness(ja_ctx*, ja_hkzg*, const uint64_t r[4], const uint64_t* q_powers, uint64_t* w_xy,
                                 int32_t* w_inf);
void ja_hyperkzg_open_free(ja_ctx*, ja_hkzg*);
/* The same three steps with the library's own Blake2b transcript (joltworks/src/transcripts/blake2b.rs) in between:
 * transcript_state/n_rounds are the running state and round counter, read on entry and written back. */
int32_t ja_hyperkzg_open(ja_ctx*, const ja_srs*, const ja_poly* poly, const uint64_t* point, size_t ell,
                         uint8_t transcript_state[32], uint32_t* n_rounds, uint64_t* com_xy, int32_t* com_inf,
                         uint64_t* w_xy, int32_t* w_inf, uint64_t* v_out);

/* ---- multi-GPU: one process per GPU; the exchange (an all-gather of <= 17 Fr per round or of one point per
 * MSM and GPU) belongs to the caller; these are the slice entry points and the host-side combine functions. ------------- */
/* Restrict every MSM this context runs (ja_hyperkzg_open_*, ja_onehot_commit, ...) to the index range of shard `index` of
 * `count` (joltworks/src/msm/mod.rs:27-181 pairs split by index; the SRS is resident on every GPU).  Results are PARTIAL
 * points: all-gather them and add with ja_g1_sum_affine.  count == 1 restores the unsharded behaviour. */
int32_t ja_set_msm_shard(ja_ctx*, uint32_t index, uint32_t count);
/* sum_{i in [lo, hi)} Z[i] * g1_powers[i] */
int32_t ja_msm_fr_range(ja_ctx*, const ja_srs*, const ja_poly* scalars, size_t lo, size_t hi, uint64_t out_xy[8], int32_t* is_inf);
/* One GPU's share of a round evaluation over contiguous hypercube slices: polys = the slice, eq = the replicated split-eq
 * of the whole instance, g_offset = slice_start / 2.  Family S / PROD / POW (LowToHigh).  Outputs are partial sums. */
int32_t ja_round_eval_slice(ja_ctx*, int32_t kernel_id, const ja_poly* const* polys, size_t n_polys, const ja_spliteq* eq,
                            uint32_t aux_u32, size_t g_offset, uint64_t* out_evals, size_t n_out);
/* Sharded Sumcheck::prove / BatchedSumcheck::prove: after this call the device polynomials handed to ja_sumcheck_prove /
 * ja_batched_sumcheck_prove are this GPU's contiguous slice (len / world coefficients, rank-major) of each MLE; eq points,
 * claims and the transcript are replicated.  Every round the partial sums (<= 17 Fr) go through `allgather` (send `bytes`,
 * receive world x bytes, rank-major; return 0) and are added; binds stay local; once a slice is down to one coefficient
 * the world remaining coefficients are gathered and the last log2(world) rounds run replicated.  Every rank returns the same
 * proof.  Split-eq (LowToHigh) bodies and PROD / POW only; world == 1 (or a NULL callback) restores the normal mode. */
typedef int32_t (*ja_allgather_fn)(void* user, const void* send, size_t bytes, void* recv);
int32_t ja_set_sumcheck_shard(ja_ctx*, uint32_t rank, uint32_t world, ja_allgather_fn allgather, void* user);
/* In-library exchange (comm.cu): an NCCL communicator owned by the context, collectives on the context's stream.  Rank 0 calls
 * ja_comm_unique_id and hands the 128 bytes to the other ranks by any means; every rank calls ja_comm_init.  From then on, with
 * ja_set_msm_shard(rank, world), every MSM (ja_hyperkzg_open*, ja_msm_fr*, ...) multiplies this rank's index range and the partial
 * POINTS are all-gathered and added inside the call (every rank returns the full result); ja_addr_commit_many deals the lists of a
 * proof round-robin to the ranks and all-gathers the commitments; ja_set_sumcheck_shard(rank, world, NULL, NULL) all-gathers the
 * partial round sums through the same communicator.  JA_ERR_UNSUPPORTED when libnccl.so.2 cannot be loaded.
 * Reference: none (single-process rayon); the split points are joltworks/src/msm/mod.rs:27-181 (pairs by index) and
 * jolt-atlas-core/src/onnx_proof/prover.rs:236-249 (one commitment per polynomial). */
int32_t ja_comm_unique_id(uint8_t out[128]);
int32_t ja_comm_init(ja_ctx*, uint32_t rank, uint32_t world, const uint8_t id[128]);
void ja_comm_free(ja_ctx*);
int32_t ja_comm_allgather(ja_ctx*, const void* send, size_t bytes, void* recv);   /* host buffers, recv = world x bytes, rank-major */
/* Host-only (no ja_ctx, no GPU): add n affine points (complete addition) / add n_parts vectors of n_vals Fr. */
int32_t ja_g1_sum_affine(const uint64_t* xy, const int32_t* is_inf, size_t n, uint64_t out_xy[8], int32_t* out_inf);
int32_t ja_fr_sum(const uint64_t* vals, size_t n_parts, size_t n_vals, uint64_t* out);
/* The library's Blake2b transcript (joltworks/src/transcripts/blake2b.rs) for callers that own state + round counter. */
/* cache_openings (SumcheckInstanceProver::cache_openings -> ProverOpeningAccumulator::append_{dense,sparse,virtual},
 * joltworks/src/poly/opening_proof.rs:281, :338, :398): the reference appends every opening claim an instance caches to the transcript at
 * the end of Sumcheck::prove / BatchedSumcheck::prove.  OFF by default: ja_sumcheck_prove / ja_batched_sumcheck_prove stop after the last
 * round and return the final claims; the caller owns the accumulator and appends what its instances cache.  ON: the drivers append the
 * final claims of every instance (instance order, polynomial order within an instance) with Transcript::append_scalar before they return -
 * right for every instance whose cached openings ARE its polynomials' final claims (all JA_EVAL_* bodies, RaVirtual, Booleanity, the
 * opening reduction).  An instance that caches something else (the read-raf cycle rounds cache ra(r), not the scaled polynomial this
 * library binds) runs with the flag off and the caller appends through ja_transcript_append_scalar_each. */
int32_t ja_set_cache_openings(ja_ctx*, int32_t on);
void ja_transcript_append_scalar_each(uint8_t state[32], uint32_t* n_rounds, const uint64_t* fr, size_t n);
void ja_transcript_new(const char* label, uint8_t state[32], uint32_t* n_rounds);
void ja_transcript_append_points(uint8_t state[32], uint32_t* n_rounds, const uint64_t* xy, const int32_t* is_inf, size_t n);
void ja_transcript_append_scalars(uint8_t state[32], uint32_t* n_rounds, const uint64_t* fr, size_t n);
void ja_transcript_challenge_scalar(uint8_t state[32], uint32_t* n_rounds, uint64_t out[4]);
void ja_transcript_challenge_scalar_powers(uint8_t state[32], uint32_t* n_rounds, size_t n, uint64_t* out);
void ja_transcript_challenge_optimized(uint8_t state[32], uint32_t* n_rounds, size_t n, uint64_t* out);   /* n x challenge_scalar_optimized: {0,0,lo,hi} each */
/* ExpandingTable (joltworks/src/utils/expanding_table.rs:62-89) after n updates from [1]: out = 2^n Fr.  Host-only O(2^n) glue
 * (the expanding tables v[phase] of ps_shout and the F tables of the one-hot address rounds). */
int32_t ja_expanding_table(const uint64_t* challenges, size_t n, int32_t order, uint64_t* out);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------ */
/* CUDA-event timer on the context's own stream (torch.cuda.Event cannot see this stream). */
int32_t ja_timer_begin(ja_ctx*);
int32_t ja_timer_end(ja_ctx*, float* out_ms);
/* Per-kernel-class launch profile: between begin and end every kernel launch of this context is bracketed by CUDA
 * events on the context's stream; end returns, per class k < ja_profile_class_count(), the number of launches and
 * the summed device milliseconds.  Adds event overhead: never wrap a headline timing in it. */
int32_t ja_profile_begin(ja_ctx*);
int32_t ja_profile_end(ja_ctx*, uint64_t* out_launches, double* out_ms, size_t n_classes);
int32_t ja_profile_class_count(void);
const char* ja_profile_class_name(int32_t k);
/* Re-run ONE kernel `iters` times back to back on resident synthetic operands of 2^log_n Fr (inputs are
 * never consumed, so every iteration does identical work; 2^log_n * 32 B should exceed the 126 MB L2).
 *   which: 0 = bind LowToHigh, 1 = bind HighToLow (n_polys polys per launch),
 *          2 = round eval MUL (split-eq, 2 polys), 3 = round eval DOT2, 4 = round eval ADD
 * Returns the average device time per launch in *out_ms. */
int32_t ja_bench_kernel(ja_ctx*, int32_t which, int32_t log_n, int32_t n_polys, int32_t iters, float* out_ms);
/* The same for ONE fused round kernel (bind previous challenge + evaluate; fused_kernels.cuh): which = 0 ADD (2 polys),
 * 1 MUL (2), 2 IDENT (1), 3 product of 4, 4 product of 16, 5 booleanity over 16, 6 opening reduction HighToLow (1).
 * Algorithmic bytes per launch = 48 * 2^log_n per polynomial. */
int32_t ja_bench_fused(ja_ctx*, int32_t which, int32_t log_n, int32_t iters, float* out_ms);
/* device-side pseudo-random canonical Fr (xorshift of the index; synthetic bench operands) */
int32_t ja_poly_random(ja_ctx*, size_t n, uint32_t seed, ja_poly** out);
/* register-resident Montgomery-product loop; returns achieved Fr-mul/s in *out_mul_per_s */
int32_t ja_calibrate_fr_mul(ja_ctx*, int32_t iters, double* out_mul_per_s);

/* ---- test hooks (tests/test_gpu_field.py) ------------------------------------------------------- */
/* The device field arithmetic applied element-wise to host arrays of n Fr (Montgomery limbs), so that golden edge vectors reach
 * fp.cuh directly: joltworks/src/field/ark.rs:76-297 (mul / add / sub / neg / square / from_i64), field/challenge/macros.rs:274-286
 * (F * MontU128Challenge; b = {0,0,lo,hi}), field/mod.rs:286-310 (delayed reduction: MUL_WIDE returns 16 * a * b through
 * fpw_mul_acc x 16 + one fpw_reduce).  FROM_I64 reads the i64 from limb 0 of a. */
enum { JA_TEST_FR_MUL = 0, JA_TEST_FR_ADD = 1, JA_TEST_FR_SUB = 2, JA_TEST_FR_MUL_CHALLENGE = 3, JA_TEST_FR_FROM_I64 = 4,
       JA_TEST_FR_MUL_WIDE = 5, JA_TEST_FR_NEG = 6, JA_TEST_FR_SQR = 7 };
int32_t ja_test_field_ops(ja_ctx*, int32_t op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* JOLT_ATLAS_B200_H */
