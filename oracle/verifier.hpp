// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Restated plonky2 verifier: the acceptance check behind
// `WormholeVerifier::verify` (/root/reference/wormhole/verifier/src/lib.rs:155-159) and
// `circuit_data.verify` (/root/reference/wormhole/aggregator/src/circuits/tree.rs:210).
// Transcript order per SURVEY.md App. A.6, vanishing identity App. A.7, FRI App. A.8. It is the
// arbiter that pins every convention the commit / quotient / FRI kernels must follow, by
// accepting the reference's own shipped proof (/root/reference/wormhole/bench-data/proof.bin).
#pragma once
#include <string>

#include "challenger.hpp"
#include "merkle.hpp"
#include "plonk.hpp"

namespace orc {

struct Challenges {
  std::vector<u64> betas, gammas, alphas;
  E2 zeta;
  E2 fri_alpha;
  std::vector<E2> fri_betas;
  u64 pow_response;
  std::vector<size_t> query_indices;
};

static inline Challenges get_challenges(const CommonData& c, const VerifierOnly& vo,
                                        const Proof& pf) {
  Challenges ch;
  Challenger t;
  Hash pih = hash_no_pad(pf.public_inputs.data(), pf.public_inputs.size());
  t.observe_hash(vo.circuit_digest);
  t.observe_hash(pih);
  t.observe_cap(pf.wires_cap);
  for (u64 i = 0; i < c.num_challenges; i++) ch.betas.push_back(t.get_challenge());
  for (u64 i = 0; i < c.num_challenges; i++) ch.gammas.push_back(t.get_challenge());
  t.observe_cap(pf.zs_cap);
  for (u64 i = 0; i < c.num_challenges; i++) ch.alphas.push_back(t.get_challenge());
  t.observe_cap(pf.quotient_cap);
  ch.zeta = t.get_ext_challenge();
  for (auto* v : {&pf.constants, &pf.sigmas, &pf.wires, &pf.zs, &pf.partial_products, &pf.quotient})
    for (E2 e : *v) t.observe_ext(e);
  for (E2 e : pf.zs_next) t.observe_ext(e);
  ch.fri_alpha = t.get_ext_challenge();
  for (auto& cap : pf.fri_caps) {
    t.observe_cap(cap);
    ch.fri_betas.push_back(t.get_ext_challenge());
  }
  for (E2 e : pf.final_poly) t.observe_ext(e);
  t.observe(pf.pow_witness);
  ch.pow_response = t.get_challenge();
  size_t lde_size = (size_t)1 << c.lde_bits();
  for (u64 i = 0; i < c.num_query_rounds; i++) ch.query_indices.push_back(t.get_challenge() % lde_size);
  return ch;
}

// eval_vanishing_poly at an extension point: returns one value per challenge.
static inline std::vector<E2> eval_vanishing_poly_ext(const CommonData& c, E2 x, const Proof& pf,
                                                      const u64* pi_hash, const Challenges& ch) {
  size_t nch = c.num_challenges, npp = c.num_partial_products, qdf = c.quotient_degree_factor;
  std::vector<E2> terms;
  u64 n = (u64)1 << c.degree_bits;
  E2 xn = e2pow(x, n);
  E2 l0 = (x == e2(1)) ? e2(1) : (xn - e2(1)) * e2inv(scale(x - e2(1), n % P));
  for (size_t i = 0; i < nch; i++) terms.push_back(l0 * (pf.zs[i] - e2(1)));
  for (size_t i = 0; i < nch; i++) {
    std::vector<E2> accs;
    accs.push_back(pf.zs[i]);
    for (size_t j = 0; j < npp; j++) accs.push_back(pf.partial_products[i * npp + j]);
    accs.push_back(pf.zs_next[i]);
    E2 beta = e2(ch.betas[i]), gamma = e2(ch.gammas[i]);
    size_t nchunks = (c.num_routed_wires + qdf - 1) / qdf;
    for (size_t k = 0; k < nchunks; k++) {
      E2 num = e2(1), den = e2(1);
      for (size_t j = k * qdf; j < (k + 1) * qdf && j < c.num_routed_wires; j++) {
        E2 w = pf.wires[j];
        num = num * (w + scale(x, c.k_is[j]) * beta + gamma);
        den = den * (w + pf.sigmas[j] * beta + gamma);
      }
      terms.push_back(accs[k] * num - accs[k + 1] * den);
    }
  }
  std::vector<E2> gate(c.num_gate_constraints, e2(0));
  eval_gate_constraints<E2>(c, pf.constants.data(), pf.wires.data(), pi_hash, gate.data());
  for (E2 g : gate) terms.push_back(g);
  std::vector<E2> out;
  for (size_t i = 0; i < nch; i++) {
    E2 a = e2(ch.alphas[i]), acc = e2(0);
    for (size_t t = terms.size(); t-- > 0;) acc = acc * a + terms[t];
    out.push_back(acc);
  }
  return out;
}

// Lagrange interpolation of (xs[i], ys[i]) evaluated at z.
static inline E2 interpolate_at(const std::vector<E2>& xs, const std::vector<E2>& ys, E2 z) {
  E2 acc = e2(0);
  for (size_t i = 0; i < xs.size(); i++) {
    E2 num = e2(1), den = e2(1);
    for (size_t j = 0; j < xs.size(); j++)
      if (j != i) {
        num = num * (z - xs[j]);
        den = den * (xs[i] - xs[j]);
      }
    acc = acc + ys[i] * num * e2inv(den);
  }
  return acc;
}

// Verification result: 0 = accepted; otherwise a code naming the first failed check.
enum VerifyCode {
  V_OK = 0,
  V_BAD_SHAPE = 1,
  V_VANISHING = 2,
  V_POW = 3,
  V_INITIAL_MERKLE = 4,
  V_FOLD_CONSISTENCY = 5,
  V_STEP_MERKLE = 6,
  V_FINAL_POLY = 7,
};

static inline int verify_proof(const CommonData& c, const VerifierOnly& vo, const Proof& pf,
                               Challenges* out_ch = nullptr) {
  if (pf.public_inputs.size() != c.num_public_inputs) return V_BAD_SHAPE;
  Hash pih = hash_no_pad(pf.public_inputs.data(), pf.public_inputs.size());
  Challenges ch = get_challenges(c, vo, pf);
  if (out_ch) *out_ch = ch;

  // --- vanishing identity at zeta ---
  std::vector<E2> van = eval_vanishing_poly_ext(c, ch.zeta, pf, pih.e, ch);
  u64 n = (u64)1 << c.degree_bits;
  E2 zeta_n = e2pow(ch.zeta, n);
  E2 zh = zeta_n - e2(1);
  size_t qdf = c.quotient_degree_factor;
  for (size_t i = 0; i < c.num_challenges; i++) {
    E2 acc = e2(0);
    for (size_t j = qdf; j-- > 0;) acc = acc * zeta_n + pf.quotient[i * qdf + j];
    if (van[i] != zh * acc) return V_VANISHING;
  }

  // --- FRI ---
  unsigned lz = ch.pow_response ? (unsigned)__builtin_clzll(ch.pow_response) : 64;
  if (lz < c.pow_bits) return V_POW;
  unsigned lde_bits = c.lde_bits();
  E2 alpha = ch.fri_alpha;
  // batches: zeta (all polys), g*zeta (zs)
  std::vector<E2> open0;
  for (auto* v : {&pf.constants, &pf.sigmas, &pf.wires, &pf.zs, &pf.partial_products, &pf.quotient})
    for (E2 e : *v) open0.push_back(e);
  std::vector<E2> open1 = pf.zs_next;
  auto reduce = [&](const std::vector<E2>& v) {
    E2 acc = e2(0);
    for (size_t i = v.size(); i-- > 0;) acc = acc * alpha + v[i];
    return acc;
  };
  E2 red0 = reduce(open0), red1 = reduce(open1);
  E2 zeta_next = scale(ch.zeta, root_of_unity(c.degree_bits));
  std::vector<size_t> widths = oracle_widths(c);
  size_t salt = c.salt_size();
  const std::vector<Hash>* init_caps[4] = {&vo.constants_sigmas_cap, &pf.wires_cap, &pf.zs_cap,
                                           &pf.quotient_cap};
  for (size_t q = 0; q < pf.queries.size(); q++) {
    const FriQueryRound& qr = pf.queries[q];
    size_t x_index = ch.query_indices[q];
    for (size_t o = 0; o < 4; o++) {
      const FriInitialOpen& io = qr.init[o];
      if (io.evals.size() != widths[o]) return V_BAD_SHAPE;
      if (io.path.size() != lde_bits - c.cap_height) return V_BAD_SHAPE;
      if (!merkle_verify(io.evals.data(), io.evals.size(), x_index, init_caps[o]->data(),
                         io.path.data(), io.path.size()))
        return V_INITIAL_MERKLE;
    }
    u64 subgroup_x = mul(GEN, pow(root_of_unity(lde_bits), bitrev(x_index, lde_bits)));
    // fri_combine_initial
    std::vector<E2> ev0, ev1;
    for (size_t o = 0; o < 4; o++) {
      size_t unsalted = widths[o] - (o == 0 ? 0 : salt);
      for (size_t i = 0; i < unsalted; i++) ev0.push_back(e2(qr.init[o].evals[i]));
    }
    for (size_t i = 0; i < c.num_challenges; i++) ev1.push_back(e2(qr.init[2].evals[i]));
    E2 sx = e2(subgroup_x);
    E2 sum = e2(0);
    sum = sum * e2pow(alpha, ev0.size()) + (reduce(ev0) - red0) * e2inv(sx - ch.zeta);
    sum = sum * e2pow(alpha, ev1.size()) + (reduce(ev1) - red1) * e2inv(sx - zeta_next);
    E2 old_eval = sum;
    size_t cur_bits = lde_bits;
    for (size_t s = 0; s < c.reduction_arity_bits.size(); s++) {
      unsigned ab = c.reduction_arity_bits[s];
      size_t arity = (size_t)1 << ab;
      const FriStep& st = qr.steps[s];
      if (st.evals.size() != arity) return V_BAD_SHAPE;
      size_t coset_index = x_index >> ab, within = x_index & (arity - 1);
      if (st.evals[within] != old_eval) return V_FOLD_CONSISTENCY;
      // compute_evaluation
      u64 g = root_of_unity(ab);
      size_t rev_within = bitrev(within, ab);
      u64 coset_start = mul(subgroup_x, pow(g, arity - rev_within));
      std::vector<E2> xs(arity), ys(arity);
      u64 y = 1;
      for (size_t i = 0; i < arity; i++) {
        xs[i] = e2(mul(coset_start, y));
        ys[i] = st.evals[bitrev(i, ab)];
        y = mul(y, g);
      }
      old_eval = interpolate_at(xs, ys, ch.fri_betas[s]);
      std::vector<u64> flat;
      for (E2 e : st.evals) {
        flat.push_back(e.a);
        flat.push_back(e.b);
      }
      cur_bits -= ab;
      if (st.path.size() != cur_bits - c.cap_height && !(cur_bits < c.cap_height && st.path.empty()))
        return V_BAD_SHAPE;
      if (!merkle_verify(flat.data(), flat.size(), coset_index, pf.fri_caps[s].data(),
                         st.path.data(), st.path.size()))
        return V_STEP_MERKLE;
      for (unsigned k = 0; k < ab; k++) subgroup_x = sqr(subgroup_x);
      x_index = coset_index;
    }
    if (eval_poly_e2(pf.final_poly, e2(subgroup_x)) != old_eval) return V_FINAL_POLY;
  }
  return V_OK;
}

}  // namespace orc
