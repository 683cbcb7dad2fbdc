// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// CPU restatement of qp-plonky2 1.1.1's `prove()` (plonk/prover.rs) and the pieces it drives:
// Z / partial products (H8), `compute_quotient_polys` (H9), `OpeningSet::new` (H10),
// `PolynomialBatch::prove_openings` (H11), `fri_committed_trees` (H12), `fri_proof_of_work` (H13)
// and `fri_prover_query_rounds` (H14) - SURVEY.md §3(A) steps 2-10, §8(a). Entered in the
// reference from /root/reference/wormhole/prover/src/lib.rs:233-237,
// /root/reference/wormhole/aggregator/src/circuits/tree.rs:136 and /root/reference/voting/src/lib.rs:356.
// Witness generation and circuit building stay outside (inputs here: the full wire matrix and the
// constants/sigma columns). Two reference non-determinisms are pinned for comparisons:
// salts are injected by the caller, and the proof-of-work witness is the SMALLEST valid one.
// Everything this prover emits is accepted or rejected by verifier.hpp, which is itself pinned by
// the reference's shipped proof.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "challenger.hpp"
#include "merkle.hpp"
#include "plonk.hpp"
#include "verifier.hpp"

namespace orc {

struct CircuitProverData {
  CommonData common;
  Hash circuit_digest;
  std::vector<std::vector<u64>> constants_sigmas;  // [num_constants + num_routed][n] values on H
  PolyBatch cs_batch;                               // committed once per circuit (build())
};

static inline CircuitProverData make_circuit(const CommonData& c, const Hash& digest,
                                             std::vector<std::vector<u64>> constants_sigmas,
                                             unsigned threads) {
  CircuitProverData d;
  d.common = c;
  d.circuit_digest = digest;
  d.constants_sigmas = std::move(constants_sigmas);
  d.cs_batch = batch_from_values(d.constants_sigmas, c.rate_bits, c.cap_height, nullptr, 0, threads);
  return d;
}

static inline std::vector<u64> batch_inverse(const std::vector<u64>& x) {
  size_t n = x.size();
  std::vector<u64> pre(n), out(n);
  u64 acc = 1;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    acc = mul(acc, x[i]);
  }
  u64 ia = inv(acc);
  for (size_t i = n; i-- > 0;) {
    out[i] = mul(ia, pre[i]);
    ia = mul(ia, x[i]);
  }
  return out;
}

// H8: [Z_0.., pp_0[0..npp), pp_1[..) ...] as value vectors on H (the zs_partial_products batch).
static inline std::vector<std::vector<u64>> compute_zs_partial_products(
    const CircuitProverData& d, const std::vector<std::vector<u64>>& wires,
    const std::vector<u64>& betas, const std::vector<u64>& gammas, unsigned threads = 1) {
  const CommonData& c = d.common;
  size_t n = (size_t)1 << c.degree_bits, nr = c.num_routed_wires, npp = c.num_partial_products;
  size_t nch = c.num_challenges, chunk = c.quotient_degree_factor;
  size_t nchunks = (nr + chunk - 1) / chunk;
  if (nchunks != npp + 1) throw std::runtime_error("partial product count mismatch");
  std::vector<std::vector<u64>> zs(nch, std::vector<u64>(n)), pps(nch * npp, std::vector<u64>(n));
  u64 w = root_of_unity(c.degree_bits);
  // pass 1 (rows in parallel, as plonky2's par_iter): the chunk quotients of every row; pass 2: the running
  // product over the rows, which is sequential
  std::vector<std::vector<u64>> quot(nch * nchunks, std::vector<u64>(n));
  const size_t rows_per_task = 256, ntasks = (n + rows_per_task - 1) / rows_per_task;
  for (size_t ch = 0; ch < nch; ch++) {
    parallel_for(ntasks, threads, [&](size_t task) {
      size_t i0 = task * rows_per_task, i1 = i0 + rows_per_task < n ? i0 + rows_per_task : n;
      u64 x = pow(w, i0);
      std::vector<u64> num(nr), den(nr);
      for (size_t i = i0; i < i1; i++) {
        for (size_t j = 0; j < nr; j++) {
          u64 wv = wires[j][i];
          num[j] = add(add(wv, mul(betas[ch], mul(c.k_is[j], x))), gammas[ch]);
          den[j] = add(add(wv, mul(betas[ch], d.constants_sigmas[c.num_constants + j][i])), gammas[ch]);
        }
        std::vector<u64> di = batch_inverse(den);
        for (size_t k = 0; k < nchunks; k++) {
          u64 prod = 1;
          for (size_t j = k * chunk; j < (k + 1) * chunk && j < nr; j++) prod = mul(prod, mul(num[j], di[j]));
          quot[ch * nchunks + k][i] = prod;
        }
        x = mul(x, w);
      }
    });
    u64 z = 1;
    for (size_t i = 0; i < n; i++) {
      zs[ch][i] = z;
      u64 acc = z;
      for (size_t k = 0; k < nchunks; k++) {
        acc = mul(acc, quot[ch * nchunks + k][i]);
        if (k < npp) pps[ch * npp + k][i] = acc;
      }
      z = acc;  // Z(g x)
    }
  }
  std::vector<std::vector<u64>> out;
  for (auto& v : zs) out.push_back(std::move(v));
  for (auto& v : pps) out.push_back(std::move(v));
  return out;
}

// eval_vanishing_poly_base at one LDE point x (index i in natural order on g*<w_{n*2^qdb}>).
static inline void eval_vanishing_base(const CommonData& c, u64 x, u64 zh_x, const u64* consts_sigmas,
                                       const u64* wires, const u64* zs_pp, const u64* zs_pp_next,
                                       const u64* pi_hash, const std::vector<u64>& betas,
                                       const std::vector<u64>& gammas, const std::vector<u64>& alphas,
                                       u64* out) {
  size_t nch = c.num_challenges, npp = c.num_partial_products, qdf = c.quotient_degree_factor;
  size_t nr = c.num_routed_wires;
  u64 n = (u64)1 << c.degree_bits;
  std::vector<u64> terms;
  // L_0(x) = Z_H(x) / (n (x - 1))
  u64 l0 = mul(zh_x, inv(mul(n % P, sub(x, 1))));
  for (size_t i = 0; i < nch; i++) terms.push_back(mul(l0, sub(zs_pp[i], 1)));
  const u64* sig = consts_sigmas + c.num_constants;
  for (size_t i = 0; i < nch; i++) {
    std::vector<u64> accs;
    accs.push_back(zs_pp[i]);
    for (size_t j = 0; j < npp; j++) accs.push_back(zs_pp[nch + i * npp + j]);
    accs.push_back(zs_pp_next[i]);
    size_t nchunks = (nr + qdf - 1) / qdf;
    for (size_t k = 0; k < nchunks; k++) {
      u64 num = 1, den = 1;
      for (size_t j = k * qdf; j < (k + 1) * qdf && j < nr; j++) {
        num = mul(num, add(add(wires[j], mul(betas[i], mul(c.k_is[j], x))), gammas[i]));
        den = mul(den, add(add(wires[j], mul(betas[i], sig[j])), gammas[i]));
      }
      terms.push_back(sub(mul(accs[k], num), mul(accs[k + 1], den)));
    }
  }
  std::vector<Fp> gate(c.num_gate_constraints), lc(c.num_constants), lw(c.num_wires);
  for (size_t j = 0; j < c.num_constants; j++) lc[j] = Fp{consts_sigmas[j]};
  for (size_t j = 0; j < c.num_wires; j++) lw[j] = Fp{wires[j]};
  eval_gate_constraints<Fp>(c, lc.data(), lw.data(), pi_hash, gate.data());
  for (Fp g : gate) terms.push_back(g.v);
  for (size_t i = 0; i < nch; i++) {
    u64 acc = 0;
    for (size_t t = terms.size(); t-- > 0;) acc = add(mul(acc, alphas[i]), terms[t]);
    out[i] = acc;
  }
}

// H9: quotient chunk coefficient vectors, [nch * qdf][n].  Throws if not divisible by Z_H.
static inline std::vector<std::vector<u64>> compute_quotient_chunks(
    const CircuitProverData& d, const PolyBatch& wires_b, const PolyBatch& zs_b, const u64* pi_hash,
    const std::vector<u64>& betas, const std::vector<u64>& gammas, const std::vector<u64>& alphas,
    unsigned threads) {
  const CommonData& c = d.common;
  unsigned qdb = log2_strict(c.quotient_degree_factor);
  if (((size_t)1 << qdb) != c.quotient_degree_factor || qdb > c.rate_bits)
    throw std::runtime_error("unsupported quotient degree factor");
  size_t step = (size_t)1 << (c.rate_bits - qdb), next_step = (size_t)1 << qdb;
  unsigned lb = c.degree_bits + qdb;
  size_t lde = (size_t)1 << lb, n = (size_t)1 << c.degree_bits, nch = c.num_challenges;
  u64 wl = root_of_unity(lb);
  // Z_H on the coset: x^n - 1 takes 2^qdb values
  std::vector<u64> zh(next_step), zh_inv(next_step);
  u64 gn = pow(GEN, n), wq = root_of_unity(qdb);
  for (size_t i = 0; i < next_step; i++) {
    zh[i] = sub(mul(gn, pow(wq, i)), 1);
    zh_inv[i] = inv(zh[i]);
  }
  std::vector<std::vector<u64>> qv(nch, std::vector<u64>(lde));
  size_t nblk = (lde + 1023) / 1024;
  parallel_for(nblk, threads, [&](size_t blk) {
    for (size_t i = blk * 1024; i < (blk + 1) * 1024 && i < lde; i++) {
      u64 x = mul(GEN, pow(wl, i));
      size_t inext = (i + next_step) % lde;
      u64 out[8];
      eval_vanishing_base(c, x, zh[i % next_step], batch_lde_row(d.cs_batch, i, step),
                          batch_lde_row(wires_b, i, step), batch_lde_row(zs_b, i, step),
                          batch_lde_row(zs_b, inext, step), pi_hash, betas, gammas, alphas, out);
      for (size_t ch = 0; ch < nch; ch++) qv[ch][i] = mul(out[ch], zh_inv[i % next_step]);
    }
  });
  std::vector<std::vector<u64>> chunks;
  for (size_t ch = 0; ch < nch; ch++) {
    std::vector<u64> coeffs = coset_ifft(qv[ch], GEN);
    // trim_to_len(qdf * n): everything above must be zero (here lde == qdf*n, so nothing to trim)
    for (size_t j = 0; j < c.quotient_degree_factor; j++)
      chunks.emplace_back(coeffs.begin() + j * n, coeffs.begin() + (j + 1) * n);
  }
  return chunks;
}

// Synthetic division: (P(X) - P(z)) / (X - z), one coefficient shorter; padded back with a zero.
static inline std::vector<E2> divide_by_linear(const std::vector<E2>& p, E2 z) {
  size_t n = p.size();
  std::vector<E2> q(n, e2(0));
  E2 acc = e2(0);
  for (size_t m = n; m-- > 1;) {
    acc = p[m] + z * acc;
    q[m - 1] = acc;
  }
  return q;
}

struct ProverSalts {
  const u64* wires = nullptr;     // [4][N] or NULL
  const u64* zs_pp = nullptr;
  const u64* quotient = nullptr;
};

struct ProveTrace {  // intermediate values exposed for stage-by-stage parity tests
  std::vector<u64> betas, gammas, alphas;
  E2 zeta, fri_alpha;
  std::vector<E2> fri_betas;
  std::vector<std::vector<u64>> zs_pp_values, quotient_chunks;
  std::vector<E2> final_poly_coeffs_initial;  // the polynomial that enters FRI (n coefficients)
};

// ORC_TIMING=1: per-stage wall time of prove() on stderr (where the CPU baseline spends its time)
struct StageTimer {
  bool on = getenv("ORC_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "  [orc] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

static inline Proof prove(const CircuitProverData& d, const std::vector<std::vector<u64>>& wires,
                          const std::vector<u64>& public_inputs, const ProverSalts& salts,
                          unsigned threads, ProveTrace* trace = nullptr) {
  const CommonData& c = d.common;
  size_t n = (size_t)1 << c.degree_bits, nch = c.num_challenges;
  unsigned salt_cols = c.salt_size();
  if (wires.size() != c.num_wires) throw std::runtime_error("wire count mismatch");
  if (c.hiding && !(salts.wires && salts.zs_pp && salts.quotient))
    throw std::runtime_error("hiding circuit needs salts");
  Proof pf;
  pf.public_inputs = public_inputs;
  Hash pih = hash_no_pad(public_inputs.data(), public_inputs.size());
  StageTimer tm;

  PolyBatch wires_b = batch_from_values(wires, c.rate_bits, c.cap_height, c.hiding ? salts.wires : nullptr,
                                        salt_cols, threads);
  tm.lap("commit wires");
  Challenger ch;
  ch.observe_hash(d.circuit_digest);
  ch.observe_hash(pih);
  ch.observe_cap(wires_b.tree.cap);
  std::vector<u64> betas, gammas, alphas;
  for (size_t i = 0; i < nch; i++) betas.push_back(ch.get_challenge());
  for (size_t i = 0; i < nch; i++) gammas.push_back(ch.get_challenge());

  std::vector<std::vector<u64>> zs_pp = compute_zs_partial_products(d, wires, betas, gammas, threads);
  tm.lap("zs / partial products");
  PolyBatch zs_b = batch_from_values(zs_pp, c.rate_bits, c.cap_height, c.hiding ? salts.zs_pp : nullptr,
                                     salt_cols, threads);
  tm.lap("commit zs");
  ch.observe_cap(zs_b.tree.cap);
  for (size_t i = 0; i < nch; i++) alphas.push_back(ch.get_challenge());

  std::vector<std::vector<u64>> qchunks =
      compute_quotient_chunks(d, wires_b, zs_b, pih.e, betas, gammas, alphas, threads);
  tm.lap("quotient");
  PolyBatch q_b = batch_from_coeffs(qchunks, c.rate_bits, c.cap_height, c.hiding ? salts.quotient : nullptr,
                                    salt_cols, threads);
  tm.lap("commit quotient");
  ch.observe_cap(q_b.tree.cap);
  E2 zeta = ch.get_ext_challenge();
  if (e2pow(zeta, n) == e2(1)) throw std::runtime_error("Opening point is in the subgroup.");

  // H10 openings
  const PolyBatch* oracles[4] = {&d.cs_batch, &wires_b, &zs_b, &q_b};
  auto eval_all = [&](const PolyBatch& b, size_t from, size_t to, E2 x, std::vector<E2>& out) {
    size_t base = out.size();
    out.resize(base + (to - from));
    parallel_for(to - from, threads, [&](size_t j) { out[base + j] = eval_base_poly_at_e2(b.coeffs[from + j].data(), n, x); });
  };
  E2 zeta_next = scale(zeta, root_of_unity(c.degree_bits));
  eval_all(d.cs_batch, 0, c.num_constants, zeta, pf.constants);
  eval_all(d.cs_batch, c.num_constants, c.num_constants + c.num_routed_wires, zeta, pf.sigmas);
  eval_all(wires_b, 0, c.num_wires, zeta, pf.wires);
  eval_all(zs_b, 0, nch, zeta, pf.zs);
  eval_all(zs_b, 0, nch, zeta_next, pf.zs_next);
  eval_all(zs_b, nch, nch * (1 + c.num_partial_products), zeta, pf.partial_products);
  eval_all(q_b, 0, nch * c.quotient_degree_factor, zeta, pf.quotient);
  for (auto* v : {&pf.constants, &pf.sigmas, &pf.wires, &pf.zs, &pf.partial_products, &pf.quotient})
    for (E2 e : *v) ch.observe_ext(e);
  for (E2 e : pf.zs_next) ch.observe_ext(e);

  tm.lap("openings");
  // H11 prove_openings: final_poly = sum_batches alpha^(..) (F_b(X) - F_b(z_b)) / (X - z_b)
  E2 alpha = ch.get_ext_challenge();
  std::vector<E2> final_poly(n, e2(0));
  {
    // batch 0: every polynomial of the four oracles, opened at zeta
    std::vector<E2> comp(n, e2(0));
    E2 ap = e2(1);
    size_t count = 0;
    for (const PolyBatch* b : oracles)
      for (size_t j = 0; j < b->ncols; j++) {
        for (size_t m = 0; m < n; m++) comp[m] = comp[m] + scale(ap, b->coeffs[j][m]);
        ap = ap * alpha;
        count++;
      }
    std::vector<E2> q0 = divide_by_linear(comp, zeta);
    // final_poly (currently zero) * alpha^count + q0
    final_poly = q0;
    // batch 1: the Z polynomials, opened at g*zeta
    std::vector<E2> comp1(n, e2(0));
    ap = e2(1);
    size_t count1 = 0;
    for (size_t j = 0; j < nch; j++) {
      for (size_t m = 0; m < n; m++) comp1[m] = comp1[m] + scale(ap, zs_b.coeffs[j][m]);
      ap = ap * alpha;
      count1++;
    }
    std::vector<E2> q1 = divide_by_linear(comp1, zeta_next);
    E2 shift = e2pow(alpha, count1);
    for (size_t m = 0; m < n; m++) final_poly[m] = final_poly[m] * shift + q1[m];
    (void)count;
  }
  if (trace) trace->final_poly_coeffs_initial = final_poly;

  tm.lap("fri combine");
  // H12 commit phase
  unsigned lde_bits = c.lde_bits();
  std::vector<E2> coeffs = final_poly;  // logical length n; LDE padding is implicit
  size_t cur_n = n;                     // number of (possibly non-zero) coefficients
  u64 shift = GEN;
  std::vector<MerkleTree> fri_trees;
  auto ext_coset_lde = [&](const std::vector<E2>& cf, size_t len, u64 sh) {
    // values of the polynomial on sh*<w_(len*2^r)>, natural order
    std::vector<u64> a(len), b(len);
    for (size_t i = 0; i < len; i++) { a[i] = cf[i].a; b[i] = cf[i].b; }
    std::vector<u64> va = coset_fft(lde(a, c.rate_bits), sh), vb = coset_fft(lde(b, c.rate_bits), sh);
    std::vector<E2> v(va.size());
    for (size_t i = 0; i < va.size(); i++) v[i] = E2{va[i], vb[i]};
    return v;
  };
  std::vector<E2> values = ext_coset_lde(coeffs, cur_n, shift);
  std::vector<E2> fri_betas;
  for (u64 ab : c.reduction_arity_bits) {
    size_t arity = (size_t)1 << ab;
    size_t nv = values.size();
    unsigned vb = log2_strict(nv);
    std::vector<u64> leaves(nv * 2);
    for (size_t i = 0; i < nv; i++) {  // reverse_index_bits, then chunks of `arity`, flattened
      size_t p = bitrev(i, vb);
      leaves[2 * p] = values[i].a;
      leaves[2 * p + 1] = values[i].b;
    }
    MerkleTree t = merkle_new(std::move(leaves), nv / arity, 2 * arity, c.cap_height, threads);
    ch.observe_cap(t.cap);
    pf.fri_caps.push_back(t.cap);
    fri_trees.push_back(std::move(t));
    E2 beta = ch.get_ext_challenge();
    fri_betas.push_back(beta);
    std::vector<E2> folded(cur_n / arity);
    for (size_t m = 0; m < folded.size(); m++) {
      E2 acc = e2(0);
      for (size_t j = arity; j-- > 0;) acc = acc * beta + coeffs[m * arity + j];
      folded[m] = acc;
    }
    coeffs = std::move(folded);
    cur_n /= arity;
    for (u64 k = 0; k < ab; k++) shift = sqr(shift);
    values = ext_coset_lde(coeffs, cur_n, shift);
  }
  pf.final_poly = coeffs;  // already truncated to len >> rate_bits (the padding was implicit)
  for (E2 e : pf.final_poly) ch.observe_ext(e);

  tm.lap("fri commit phase");
  // H13 proof of work: smallest witness whose response has >= pow_bits leading zeros
  {
    State st = ch.state;
    size_t pos = ch.in.size();
    for (size_t i = 0; i < pos; i++) st[i] = ch.in[i];
    // plonky2 searches (0..p) with rayon find_any; here windows of candidates are tried in parallel (eight per
    // permutation call where AVX-512 is available) and the SMALLEST witness of the first window with a hit is
    // kept, so that the result stays deterministic
    u64 wv = 0;
    const u64 per_task = 512, tasks = 4 * (u64)(threads ? threads : 1);
    for (u64 start = 0;; start += per_task * tasks) {
      std::vector<u64> found(tasks, ~0ull);
      parallel_for(tasks, threads, [&](size_t t) {
        u64 lo = start + t * per_task;
        u64 w8[8];
        for (u64 cand = lo; cand < lo + per_task; cand += 8) {
          for (int j = 0; j < 8; j++) w8[j] = cand + j;
          u64 resp[8];
          pow_responses_x8(st, pos, w8, resp);
          for (int j = 0; j < 8; j++) {
            unsigned lz = resp[j] ? (unsigned)__builtin_clzll(resp[j]) : 64;
            if (lz >= c.pow_bits) {
              found[t] = w8[j];
              return;
            }
          }
        }
      });
      u64 best = ~0ull;
      for (u64 f : found) best = f < best ? f : best;
      if (best != ~0ull) {
        wv = best;
        break;
      }
    }
    pf.pow_witness = wv;
    ch.observe(wv);
    u64 resp = ch.get_challenge();
    unsigned lz = resp ? (unsigned)__builtin_clzll(resp) : 64;
    if (lz < c.pow_bits) throw std::runtime_error("pow response mismatch");
  }

  tm.lap("proof of work");
  // H14 query rounds
  size_t lde_size = (size_t)1 << lde_bits;
  for (u64 q = 0; q < c.num_query_rounds; q++) {
    size_t x_index = ch.get_challenge() % lde_size;
    FriQueryRound qr;
    for (const PolyBatch* b : oracles) {
      FriInitialOpen io;
      io.evals.assign(b->tree.leaf(x_index), b->tree.leaf(x_index) + b->tree.leaf_len);
      io.path = merkle_prove(b->tree, x_index);
      qr.init.push_back(std::move(io));
    }
    for (size_t s = 0; s < fri_trees.size(); s++) {
      x_index >>= c.reduction_arity_bits[s];
      FriStep st;
      const u64* lf = fri_trees[s].leaf(x_index);
      for (size_t j = 0; j < fri_trees[s].leaf_len / 2; j++) st.evals.push_back(E2{lf[2 * j], lf[2 * j + 1]});
      st.path = merkle_prove(fri_trees[s], x_index);
      qr.steps.push_back(std::move(st));
    }
    pf.queries.push_back(std::move(qr));
  }
  tm.lap("queries");
  pf.wires_cap = wires_b.tree.cap;
  pf.zs_cap = zs_b.tree.cap;
  pf.quotient_cap = q_b.tree.cap;
  if (trace) {
    trace->betas = betas; trace->gammas = gammas; trace->alphas = alphas;
    trace->zeta = zeta; trace->fri_alpha = alpha; trace->fri_betas = fri_betas;
    trace->zs_pp_values = zs_pp; trace->quotient_chunks = qchunks;
  }
  return pf;
}

}  // namespace orc
