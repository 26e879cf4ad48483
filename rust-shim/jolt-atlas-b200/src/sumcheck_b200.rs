//! Seam B2: GPU-backed `SumcheckInstanceProver`s (joltworks/src/subprotocols/sumcheck_prover.rs:10-68).  Two ways in:
//! * round by round (`ja_round_eval` + `ja_bind_many`): the instance keeps the reference's driver and transcript;
//! * whole loops (`ja_sumcheck_prove` / `ja_batched_sumcheck_prove`): the library runs `Sumcheck::prove` / `BatchedSumcheck::prove`
//!   with its own Blake2b transcript seeded from and written back to the caller's (`transcript.state`, `n_rounds`) - the fast path
//!   (round-resident kernels, one host<->device exchange per round for all instances of a batch).
use crate::{check, fr_from_limbs, limbs, Ctx};
use ark_bn254::Fr;
use jolt_atlas_b200_sys as sys;
use joltworks::field::JoltField;
use joltworks::poly::split_eq_poly::GruenSplitEqPolynomial;
use joltworks::poly::unipoly::UniPoly;
use joltworks::subprotocols::sumcheck_prover::SumcheckInstanceProver;
use joltworks::transcripts::Transcript;
use std::os::raw::c_void;
use std::sync::Arc;

/// `MulProver` (jolt-atlas-core/src/onnx_proof/ops/mul.rs:134-201) with the operands and the split-eq tables resident in HBM.
/// The O(1) assembly of the round cubic stays in Rust (`gruen_poly_deg_3`), exactly as in the reference.
pub struct MulProverB200 {
    pub ctx: Arc<Ctx>,
    pub left: *mut c_void,
    pub right: *mut c_void,
    pub eq_dev: *mut c_void,
    pub eq_host: GruenSplitEqPolynomial<Fr>,
}

impl<T: Transcript> SumcheckInstanceProver<Fr, T> for MulProverB200 {
    fn compute_message(&mut self, _round: usize, previous_claim: Fr) -> UniPoly<Fr> {
        let mut q = [0u64; 8];
        let polys = [self.left, self.right];
        check(unsafe {
            sys::ja_round_eval(self.ctx.0, sys::JA_EVAL_MUL, polys.as_ptr() as *mut *mut c_void, 2, self.eq_dev, std::ptr::null_mut(), 0, 0,
                               q.as_mut_ptr() as *mut c_void, 2)
        });
        self.eq_host.gruen_poly_deg_3(fr_from_limbs(&q[0..4]), fr_from_limbs(&q[4..8]), previous_claim)      // mul.rs:177
    }

    fn ingest_challenge(&mut self, r_j: <Fr as JoltField>::Challenge, _round: usize) {
        let r: Fr = r_j.into();
        let l = limbs(&r);
        check(unsafe { sys::ja_spliteq_bind(self.ctx.0, self.eq_dev, l.as_ptr() as *mut c_void) });
        self.eq_host.bind(r_j);
        let polys = [self.left, self.right];
        check(unsafe { sys::ja_bind_many(self.ctx.0, polys.as_ptr() as *mut *mut c_void, 2, l.as_ptr() as *mut c_void, sys::JA_LOW_TO_HIGH) });
    }
    // get_params / cache_openings: unchanged from MulProver; final claims through ja_final_claim(left), ja_final_claim(right).
}

/// The RA one-hot checks of a node (`ra_onehot_provers`, joltworks/src/subprotocols/shout.rs:399-466) as ONE library call:
/// BatchedSumcheck[RaVirtual (product of d), HammingWeight over the G tables, Booleanity].  Returns the compressed round
/// polynomials and challenges; `state` / `n_rounds` are the caller's Blake2bTranscript fields, read and written back.
#[allow(clippy::too_many_arguments)]
pub fn ra_onehot_checks(ctx: &Ctx, addr: *mut c_void, ra: &[*mut c_void], g_tables: &[u64], k: usize, r_cycle: &[u64], r_address: &[u64], gammas: &[u64],
                        hw_gammas: &[u64], claims: [[u64; 4]; 2], state: &mut [u8; 32], n_rounds: &mut u32, max_coeffs: usize)
                        -> (Vec<u64>, Vec<u32>, Vec<u64>) {
    let d = ra.len();
    let log_t = r_cycle.len() / 4;
    let log_k = r_address.len() / 4;
    let aux: Vec<u64> = gammas.iter().chain(r_address.iter()).cloned().collect();
    let z = std::ptr::null();
    let inst = [
        sys::ja_sc_instance { kind: sys::JA_EVAL_PROD, aux_u32: 0, n_polys: d, polys: ra.as_ptr(), host_tables: z, table_len: 0, addr: std::ptr::null(),
                              eq_w: r_cycle.as_ptr(), eq_m: log_t, aux_fr: z, n_aux: 0, claim: claims[0], out_final_claims: std::ptr::null_mut() },
        sys::ja_sc_instance { kind: sys::JA_INST_HAMMING_TABLES, aux_u32: 0, n_polys: d, polys: std::ptr::null(), host_tables: g_tables.as_ptr(), table_len: k,
                              addr: std::ptr::null(), eq_w: z, eq_m: 0, aux_fr: hw_gammas.as_ptr(), n_aux: d, claim: claims[1], out_final_claims: std::ptr::null_mut() },
        sys::ja_sc_instance { kind: sys::JA_INST_BOOLEANITY, aux_u32: log_k as u32, n_polys: d, polys: std::ptr::null(), host_tables: g_tables.as_ptr(), table_len: k,
                              addr, eq_w: r_cycle.as_ptr(), eq_m: log_t, aux_fr: aux.as_ptr(), n_aux: d + log_k, claim: [0; 4], out_final_claims: std::ptr::null_mut() },
    ];
    let rounds = log_k + log_t;
    let (mut coeffs, mut ncoeffs, mut chal) = (vec![0u64; rounds * max_coeffs * 4], vec![0u32; rounds], vec![0u64; rounds * 4]);
    check(unsafe {
        sys::ja_batched_sumcheck_prove(ctx.0, inst.as_ptr() as *mut c_void, 3, state.as_mut_ptr() as *mut std::os::raw::c_char, n_rounds, max_coeffs,
                                       coeffs.as_mut_ptr() as *mut c_void, ncoeffs.as_mut_ptr(), chal.as_mut_ptr() as *mut c_void)
    });
    (coeffs, ncoeffs, chal)
}
