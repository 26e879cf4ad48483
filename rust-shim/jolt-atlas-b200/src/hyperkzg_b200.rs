//! Seam B1: `impl CommitmentScheme for HyperKzgB200` (joltworks/src/poly/commitment/commitment_scheme.rs:11-131).
//! Associated types are the stock HyperKZG ones, so proofs verify with the unchanged CPU verifier.
use crate::{check, fr_from_limbs, limbs, Ctx};
use ark_bn254::{Bn254, Fr, G1Affine};
use jolt_atlas_b200_sys as sys;
use joltworks::field::JoltField;
use joltworks::poly::commitment::commitment_scheme::CommitmentScheme;
use joltworks::poly::commitment::hyperkzg::{HyperKZG, HyperKZGCommitment, HyperKZGProof, HyperKZGProverKey, HyperKZGVerifierKey};
use joltworks::poly::multilinear_polynomial::MultilinearPolynomial;
use joltworks::transcripts::Transcript;
use joltworks::utils::errors::ProofVerifyError;
use std::os::raw::c_void;
use std::sync::Arc;

#[derive(Clone)]
pub struct HyperKzgB200;

/// CPU key (for `setup_verifier` / `combine_commitments`) + the SRS resident in HBM with its fixed-base window tables.
#[derive(Clone)]
pub struct B200ProverSetup {
    pub cpu: HyperKZGProverKey<Bn254>,
    pub ctx: Arc<Ctx>,
    pub srs: *mut c_void,
}
unsafe impl Send for B200ProverSetup {}
unsafe impl Sync for B200ProverSetup {}

fn affine_limbs(p: &G1Affine) -> [u64; 8] {
    let (x, y) = (p.x.0 .0, p.y.0 .0);
    [x[0], x[1], x[2], x[3], y[0], y[1], y[2], y[3]]
}
fn affine_from(xy: &[u64], inf: i32) -> G1Affine {
    if inf != 0 {
        return G1Affine::identity();
    }
    G1Affine::new_unchecked(
        ark_ff::Fp::new_unchecked(ark_ff::BigInt::new([xy[0], xy[1], xy[2], xy[3]])),
        ark_ff::Fp::new_unchecked(ark_ff::BigInt::new([xy[4], xy[5], xy[6], xy[7]])),
    )
}

impl CommitmentScheme for HyperKzgB200 {
    type Field = Fr;
    type ProverSetup = B200ProverSetup;
    type VerifierSetup = HyperKZGVerifierKey<Bn254>;
    type Commitment = HyperKZGCommitment<Bn254>;
    type Proof = HyperKZGProof<Bn254>;
    type BatchedProof = HyperKZGProof<Bn254>;
    type OpeningProofHint = ();

    /// hyperkzg/commitment_scheme.rs:36-44: the SRS is sampled exactly as today (ChaCha20 + arkworks UniformRand stay on the host),
    /// uploaded once (`ja_srs_upload`) and its window tables built when HBM allows (`ja_srs_precompute`: 29 x the SRS at >= 2^21 points).
    fn setup_prover(max_num_vars: usize) -> Self::ProverSetup {
        let cpu = <HyperKZG<Bn254> as CommitmentScheme>::setup_prover(max_num_vars);
        let xy: Vec<u64> = cpu.kzg_pk.g1_powers().iter().flat_map(|p| affine_limbs(p)).collect();
        let ctx = Arc::new(Ctx::new(0));
        let mut srs = std::ptr::null_mut();
        check(unsafe { sys::ja_srs_upload(ctx.0, xy.as_ptr() as *mut c_void, xy.len() / 8, &mut srs) });
        check(unsafe { sys::ja_srs_precompute(ctx.0, srs) });
        B200ProverSetup { cpu, ctx, srs }
    }

    fn setup_verifier(s: &Self::ProverSetup) -> Self::VerifierSetup {
        <HyperKZG<Bn254> as CommitmentScheme>::setup_verifier(&s.cpu)
    }

    /// hyperkzg/commitment_scheme.rs:54-73.  One-hot polynomials commit as a sum of selected SRS points (hyperkzg/mod.rs:520-554),
    /// compact polynomials by their scalar width (msm/mod.rs:27-181), dense ones as a full-width MSM.
    fn commit(poly: &MultilinearPolynomial<Fr>, s: &Self::ProverSetup) -> (Self::Commitment, Self::OpeningProofHint) {
        let (mut xy, mut inf) = ([0u64; 8], 0i32);
        let c = s.ctx.0;
        match poly {
            MultilinearPolynomial::OneHot(p) => {
                let idx: Vec<u64> = p.nonzero_indices.iter().enumerate().filter_map(|(t, k)| k.map(|k| (k as usize * p.T + t) as u64)).collect();
                check(unsafe { sys::ja_g1_sum_indexed(c, s.srs, idx.as_ptr() as *mut c_void, idx.len(), xy.as_mut_ptr() as *mut c_void, &mut inf) });
            }
            MultilinearPolynomial::I32Scalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.coeffs.as_ptr() as *mut c_void, sys::JA_MSM_I32, p.coeffs.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            MultilinearPolynomial::U8Scalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.coeffs.as_ptr() as *mut c_void, sys::JA_MSM_U8, p.coeffs.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            MultilinearPolynomial::U16Scalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.coeffs.as_ptr() as *mut c_void, sys::JA_MSM_U16, p.coeffs.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            MultilinearPolynomial::U32Scalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.coeffs.as_ptr() as *mut c_void, sys::JA_MSM_U32, p.coeffs.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            MultilinearPolynomial::U64Scalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.coeffs.as_ptr() as *mut c_void, sys::JA_MSM_U64, p.coeffs.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            MultilinearPolynomial::I64Scalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.coeffs.as_ptr() as *mut c_void, sys::JA_MSM_I64, p.coeffs.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            MultilinearPolynomial::LargeScalars(p) => check(unsafe {
                sys::ja_msm_host(c, s.srs, 0, p.Z.as_ptr() as *mut c_void, sys::JA_MSM_FR, p.Z.len(), xy.as_mut_ptr() as *mut c_void, &mut inf)
            }),
            _ => return <HyperKZG<Bn254> as CommitmentScheme>::commit(poly, &s.cpu),
        }
        (HyperKZGCommitment(affine_from(&xy, inf)), ())
    }

    fn batch_commit<U: std::borrow::Borrow<MultilinearPolynomial<Fr>> + Sync>(polys: &[U], s: &Self::ProverSetup) -> Vec<(Self::Commitment, ())> {
        polys.iter().map(|p| Self::commit(p.borrow(), s)).collect()
    }

    /// HyperKZG::open (hyperkzg/mod.rs:400-447) split at its two transcript interaction points: the caller's own Blake2bTranscript
    /// absorbs the commitments / evaluations and draws r and the q-powers between the three device calls.
    fn prove<T: Transcript>(s: &Self::ProverSetup, poly: &MultilinearPolynomial<Fr>, point: &[<Fr as JoltField>::Challenge], _h: Option<()>,
                            transcript: &mut T) -> Self::Proof {
        let c = s.ctx.0;
        let ell = point.len();
        let z: Vec<Fr> = (0..poly.len()).map(|i| poly.get_coeff(i)).collect();          // the materialised RLC polynomial is dense
        let mut dev = std::ptr::null_mut();
        check(unsafe { sys::ja_poly_from_fr(c, z.as_ptr() as *mut c_void, z.len(), &mut dev) });
        let pt: Vec<u64> = point.iter().flat_map(|r| limbs(&(*r).into())).collect();
        let (mut com, mut cinf) = (vec![0u64; 8 * (ell - 1)], vec![0i32; ell - 1]);
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::ja_hyperkzg_open_begin(c, s.srs, dev, pt.as_ptr() as *mut c_void, ell, &mut h, com.as_mut_ptr() as *mut c_void, cinf.as_mut_ptr()) });
        let com: Vec<G1Affine> = (0..ell - 1).map(|i| affine_from(&com[8 * i..8 * i + 8], cinf[i])).collect();
        transcript.append_points(&com.iter().map(|p| (*p).into()).collect::<Vec<_>>());   // mod.rs:439
        let r: Fr = transcript.challenge_scalar();                                        // mod.rs:440
        let mut v = vec![0u64; 3 * ell * 4];
        check(unsafe { sys::ja_hyperkzg_open_evals(c, h, limbs(&r).as_ptr() as *mut c_void, v.as_mut_ptr() as *mut c_void) });
        let v: Vec<Vec<Fr>> = (0..3).map(|i| (0..ell).map(|j| fr_from_limbs(&v[4 * (i * ell + j)..])).collect()).collect();
        transcript.append_scalars(&v.iter().flatten().cloned().collect::<Vec<_>>());      // mod.rs:258
        let q: Vec<Fr> = transcript.challenge_scalar_powers(ell);                         // mod.rs:260
        let ql: Vec<u64> = q.iter().flat_map(limbs).collect();
        let (mut w, mut winf) = ([0u64; 24], [0i32; 3]);
        check(unsafe { sys::ja_hyperkzg_open_witness(c, h, limbs(&r).as_ptr() as *mut c_void, ql.as_ptr() as *mut c_void, w.as_mut_ptr() as *mut c_void, winf.as_mut_ptr()) });
        let w: Vec<G1Affine> = (0..3).map(|i| affine_from(&w[8 * i..8 * i + 8], winf[i])).collect();
        transcript.append_points(&w.iter().map(|p| (*p).into()).collect::<Vec<_>>());     // mod.rs:276
        let _d0: Fr = transcript.challenge_scalar();                                      // mod.rs:277
        unsafe {
            sys::ja_hyperkzg_open_free(c, h);
            sys::ja_poly_free(c, dev);
        }
        HyperKZGProof { com, w: [w[0], w[1], w[2]], v: [v[0].clone(), v[1].clone(), v[2].clone()] }
    }

    fn verify<T: Transcript>(proof: &Self::Proof, vk: &Self::VerifierSetup, transcript: &mut T, point: &[<Fr as JoltField>::Challenge], eval: &Fr,
                             commitment: &Self::Commitment) -> Result<(), ProofVerifyError> {
        <HyperKZG<Bn254> as CommitmentScheme>::verify(proof, vk, transcript, point, eval, commitment)     // unchanged CPU verifier
    }

    fn combine_commitments<C: std::borrow::Borrow<Self::Commitment>>(commitments: &[C], coeffs: &[Fr]) -> Self::Commitment {
        <HyperKZG<Bn254> as CommitmentScheme>::combine_commitments(commitments, coeffs)
    }

    fn protocol_name() -> &'static [u8] {
        b"hyperkzg"
    }
}
