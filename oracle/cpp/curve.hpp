// TEST INFRASTRUCTURE ONLY — CPU oracle (C++ restatement) of BN254 G1 and the MSM family.
// The reference delegates to ark-ec / jolt-optimizations (external, un-vendored a16z/arkworks-algebra@76bb3a4):
//   joltworks/src/msm/mod.rs:27-181 (dispatch by scalar width; i32/i64 -> msm(pos) - msm(neg)), :184-190, :309-318
//   joltworks/src/poly/commitment/hyperkzg/mod.rs:520-554 (commit_one_hot -> batch_g1_additions_multi)
// Restated here as the textbook bucket method ark-ec's VariableBaseMSM::msm implements (window-parallel, OpenMP
// in place of rayon).  The result is a group element, so any correct method yields the same affine point.
// Parity unpinned at the byte level.
#pragma once
#include <omp.h>
#include <vector>
#include "field.hpp"

namespace orc {

struct G1Affine { Fq x, y; bool inf; };
struct G1Jac { Fq X, Y, Z; bool is_inf() const { return Z.is_zero(); } };

inline G1Jac jac_inf() { return G1Jac{Fq::one(), Fq::one(), Fq::zero()}; }
inline G1Affine g1_generator() { return G1Affine{Fq::from_u64(1), Fq::from_u64(2), false}; }
inline G1Jac to_jac(const G1Affine& p) { return p.inf ? jac_inf() : G1Jac{p.x, p.y, Fq::one()}; }

inline G1Jac jac_double(const G1Jac& p) {
  if (p.is_inf() || p.Y.is_zero()) return jac_inf();
  Fq A = p.X.sqr(), B = p.Y.sqr(), C = B.sqr();
  Fq t = (p.X + B).sqr() - A - C;
  Fq D = t + t;
  Fq E = A + A + A;
  Fq F = E.sqr();
  Fq X3 = F - D - D;
  Fq C8 = C.dbl().dbl().dbl();
  Fq Y3 = E * (D - X3) - C8;
  Fq Z3 = (p.Y * p.Z).dbl();
  return G1Jac{X3, Y3, Z3};
}
inline G1Jac jac_add(const G1Jac& a, const G1Jac& b) {
  if (a.is_inf()) return b;
  if (b.is_inf()) return a;
  Fq Z1Z1 = a.Z.sqr(), Z2Z2 = b.Z.sqr();
  Fq U1 = a.X * Z2Z2, U2 = b.X * Z1Z1;
  Fq S1 = a.Y * b.Z * Z2Z2, S2 = b.Y * a.Z * Z1Z1;
  if (U1 == U2) return S1 == S2 ? jac_double(a) : jac_inf();
  Fq H = U2 - U1, Rr = S2 - S1;
  Fq HH = H.sqr(), HHH = H * HH, V = U1 * HH;
  Fq X3 = Rr.sqr() - HHH - V - V;
  Fq Y3 = Rr * (V - X3) - S1 * HHH;
  Fq Z3 = a.Z * b.Z * H;
  return G1Jac{X3, Y3, Z3};
}
inline G1Jac jac_add_mixed(const G1Jac& a, const G1Affine& b) {
  if (b.inf) return a;
  if (a.is_inf()) return to_jac(b);
  Fq Z1Z1 = a.Z.sqr();
  Fq U2 = b.x * Z1Z1, S2 = b.y * a.Z * Z1Z1;
  if (a.X == U2) return a.Y == S2 ? jac_double(a) : jac_inf();
  Fq H = U2 - a.X, Rr = S2 - a.Y;
  Fq HH = H.sqr(), HHH = H * HH, V = a.X * HH;
  Fq X3 = Rr.sqr() - HHH - V - V;
  Fq Y3 = Rr * (V - X3) - a.Y * HHH;
  Fq Z3 = a.Z * H;
  return G1Jac{X3, Y3, Z3};
}
inline G1Jac jac_neg(const G1Jac& a) { return G1Jac{a.X, -a.Y, a.Z}; }
inline G1Affine to_affine(const G1Jac& p) {
  if (p.is_inf()) return G1Affine{Fq::zero(), Fq::zero(), true};
  Fq zi = p.Z.inv(), zi2 = zi.sqr();
  return G1Affine{p.X * zi2, p.Y * zi2 * zi, false};
}
inline G1Jac scalar_mul(const G1Affine& p, const uint64_t k[4]) {   // k canonical integer limbs
  G1Jac acc = jac_inf();
  for (int i = 255; i >= 0; i--) {
    acc = jac_double(acc);
    if ((k[i / 64] >> (i % 64)) & 1) acc = jac_add_mixed(acc, p);
  }
  return acc;
}
inline bool on_curve(const G1Affine& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + Fq::from_u64(3); }

// Bucket-method MSM over canonical integer scalars (n x 4 limbs), `nbits` significant bits.
inline G1Jac msm_canonical(const G1Affine* bases, const uint64_t* scalars, size_t n, int nbits = 254) {
  if (n == 0) return jac_inf();
  int c = 3;
  { size_t t = n; int lg = 0; while (t >>= 1) lg++; c = lg < 8 ? 3 : (lg * 69 / 100 + 2); if (c > 16) c = 16; }   // ark-ec's ln-based window heuristic
  if (c > nbits) c = nbits;
  const int nwin = (nbits + c - 1) / c;
  std::vector<G1Jac> wsum(nwin);
#pragma omp parallel for schedule(dynamic, 1)
  for (int w = 0; w < nwin; w++) {
    std::vector<G1Jac> buckets((size_t(1) << c) - 1, jac_inf());
    const int bit0 = w * c;
    for (size_t i = 0; i < n; i++) {
      const uint64_t* s = scalars + 4 * i;
      const int limb = bit0 / 64, off = bit0 % 64;
      uint64_t d = s[limb] >> off;
      if (off + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - off);
      d &= (uint64_t(1) << c) - 1;
      if (d) buckets[d - 1] = jac_add_mixed(buckets[d - 1], bases[i]);
    }
    G1Jac run = jac_inf(), tot = jac_inf();
    for (size_t b = buckets.size(); b-- > 0;) { run = jac_add(run, buckets[b]); tot = jac_add(tot, run); }
    wsum[w] = tot;
  }
  G1Jac acc = jac_inf();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) acc = jac_double(acc);
    acc = jac_add(acc, wsum[w]);
  }
  return acc;
}

// VariableBaseMSM::msm for LargeScalars (msm/mod.rs:32-37): Montgomery Fr scalars
inline G1Jac msm_fr(const G1Affine* bases, const Fr* scalars, size_t n) {
  std::vector<uint64_t> canon(4 * n);
#pragma omp parallel for if (n >= 1024)
  for (size_t i = 0; i < n; i++) scalars[i].to_canonical(&canon[4 * i]);
  return msm_canonical(bases, canon.data(), n, 254);
}
// I32Scalars / I64Scalars (msm/mod.rs:93-176): msm_u64(pos) - msm_u64(neg)
inline G1Jac msm_i64(const G1Affine* bases, const int64_t* scalars, size_t n) {
  std::vector<G1Affine> pb, nb; std::vector<uint64_t> ps, ns;
  for (size_t i = 0; i < n; i++) {
    if (scalars[i] > 0) { pb.push_back(bases[i]); ps.insert(ps.end(), {(uint64_t)scalars[i], 0, 0, 0}); }
    else if (scalars[i] < 0) { nb.push_back(bases[i]); ns.insert(ns.end(), {(uint64_t)(-(scalars[i] + 1)) + 1, 0, 0, 0}); }
  }
  G1Jac p = msm_canonical(pb.data(), ps.data(), pb.size(), 64), q = msm_canonical(nb.data(), ns.data(), nb.size(), 64);
  return jac_add(p, jac_neg(q));
}
// commit_one_hot (hyperkzg/mod.rs:520-554): sum of the selected bases
inline G1Jac sum_indexed(const G1Affine* bases, const uint64_t* idx, size_t n) {
  const int nt = omp_get_max_threads();
  std::vector<G1Jac> part(nt, jac_inf());
#pragma omp parallel if (n >= 256)
  {
    G1Jac acc = jac_inf();
#pragma omp for schedule(static)
    for (size_t i = 0; i < n; i++) acc = jac_add_mixed(acc, bases[idx[i]]);
    part[omp_get_thread_num()] = acc;
  }
  G1Jac tot = jac_inf();
  for (auto& p : part) tot = jac_add(tot, p);
  return tot;
}

}  // namespace orc
