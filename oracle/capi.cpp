// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
// Flat C entry points over the CPU restatement, for tests/ (ctypes), __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs. Never linked into the product library.
#include <cstdio>
#include <cstring>
#include <string>

#include "challenger.hpp"
#include "merkle.hpp"
#include "ntt.hpp"
#include "plonk.hpp"
#include "poseidon.hpp"
#include "prover.hpp"
#include "verifier.hpp"

using namespace orc;

static thread_local std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_selfcheck() { return poseidon_selfcheck(); }

void orc_poseidon(u64* state12) {
  State s;
  for (int i = 0; i < 12; i++) s[i] = canon(state12[i]);
  poseidon(s);
  for (int i = 0; i < 12; i++) state12[i] = s[i];
}
void orc_poseidon_naive(u64* state12) {
  State s;
  for (int i = 0; i < 12; i++) s[i] = canon(state12[i]);
  poseidon_naive(s);
  for (int i = 0; i < 12; i++) state12[i] = s[i];
}
// Poseidon tables so tests can pin the product's generated constant header against the oracle.
void orc_poseidon_tables(u64* rc360, u64* fast_first12, u64* fast_rc22, u64* fast_init121,
                         u64* fast_w_hat242, u64* fast_v242) {
  const PoseidonTables& T = tables();
  memcpy(rc360, T.rc, sizeof T.rc);
  memcpy(fast_first12, T.fast_first, sizeof T.fast_first);
  memcpy(fast_rc22, T.fast_rc, sizeof T.fast_rc);
  memcpy(fast_init121, T.fast_init, sizeof T.fast_init);
  memcpy(fast_w_hat242, T.fast_w_hat, sizeof T.fast_w_hat);
  memcpy(fast_v242, T.fast_v, sizeof T.fast_v);
}
void orc_hash_no_pad(const u64* x, u64 n, u64* out4) {
  Hash h = hash_no_pad(x, n);
  memcpy(out4, h.e, 32);
}
void orc_hash_or_noop(const u64* x, u64 n, u64* out4) {
  Hash h = hash_or_noop(x, n);
  memcpy(out4, h.e, 32);
}
void orc_two_to_one(const u64* l, const u64* r, u64* out4) {
  Hash a, b;
  memcpy(a.e, l, 32);
  memcpy(b.e, r, 32);
  Hash h = two_to_one(a, b);
  memcpy(out4, h.e, 32);
}

u64 orc_mul(u64 a, u64 b) { return mul(canon(a), canon(b)); }
u64 orc_inv(u64 a) { return inv(canon(a)); }
u64 orc_pow(u64 a, u64 e) { return pow(canon(a), e); }
u64 orc_root_of_unity(unsigned bits) { return root_of_unity(bits); }

void orc_fft(u64* a, u64 n) {
  std::vector<u64> v(a, a + n);
  fft_inplace(v);
  memcpy(a, v.data(), n * 8);
}
void orc_ifft(u64* a, u64 n) {
  std::vector<u64> v = ifft(std::vector<u64>(a, a + n));
  memcpy(a, v.data(), n * 8);
}
void orc_coset_fft(u64* a, u64 n, u64 shift) {
  std::vector<u64> v = coset_fft(std::vector<u64>(a, a + n), shift);
  memcpy(a, v.data(), n * 8);
}
void orc_naive_coset_eval(const u64* coeffs, u64 ncoeffs, u64 npoints, u64 shift, u64* out) {
  std::vector<u64> v = naive_coset_eval(std::vector<u64>(coeffs, coeffs + ncoeffs), npoints, shift);
  memcpy(out, v.data(), npoints * 8);
}

// MerkleTree::new over row-major leaves. digests_out: 2*(nleaves - 2^cap_height) hashes in
// plonky2's layout; cap_out: 2^cap_height hashes.
int orc_merkle_new(const u64* leaves, u64 nleaves, u64 leaf_len, unsigned cap_height,
                   unsigned threads, u64* digests_out, u64* cap_out) {
  MerkleTree t = merkle_new(std::vector<u64>(leaves, leaves + nleaves * leaf_len), nleaves,
                            leaf_len, cap_height, threads);
  if (digests_out) memcpy(digests_out, t.digests.data(), t.digests.size() * 32);
  memcpy(cap_out, t.cap.data(), t.cap.size() * 32);
  return 0;
}
// MerkleTree::prove out of a digest buffer in plonky2's layout.
int orc_merkle_prove(const u64* digests, u64 nleaves, unsigned cap_height, u64 leaf_index,
                     u64* siblings_out) {
  MerkleTree t;
  t.nleaves = nleaves;
  t.cap_height = cap_height;
  size_t nd = 2 * (nleaves - ((u64)1 << cap_height));
  t.digests.resize(nd);
  memcpy(t.digests.data(), digests, nd * 32);
  std::vector<Hash> s = merkle_prove(t, leaf_index);
  memcpy(siblings_out, s.data(), s.size() * 32);
  return (int)s.size();
}
int orc_merkle_verify(const u64* leaf, u64 leaf_len, u64 leaf_index, const u64* cap,
                      const u64* siblings, u64 nsib) {
  return merkle_verify(leaf, leaf_len, leaf_index, (const Hash*)cap, (const Hash*)siblings, nsib)
             ? 1 : 0;
}

// PolynomialBatch::from_values / from_coeffs. `in` column-major [ncols][n]; salts NULL or
// [salt_cols][n<<rate_bits] natural order. Any output pointer may be NULL.
int orc_batch_commit(const u64* in, int is_coeffs, u64 ncols, unsigned degree_bits,
                     unsigned rate_bits, unsigned cap_height, const u64* salts, unsigned salt_cols,
                     unsigned threads, u64* coeffs_out, u64* leaves_out, u64* digests_out,
                     u64* cap_out) {
  try {
    size_t n = (size_t)1 << degree_bits;
    std::vector<std::vector<u64>> cols(ncols);
    for (size_t c = 0; c < ncols; c++) {
      cols[c].assign(in + c * n, in + (c + 1) * n);
      for (auto& v : cols[c]) v = canon(v);
    }
    PolyBatch b = is_coeffs
                      ? batch_from_coeffs(std::move(cols), rate_bits, cap_height, salts, salt_cols, threads)
                      : batch_from_values(cols, rate_bits, cap_height, salts, salt_cols, threads);
    if (coeffs_out)
      for (size_t c = 0; c < ncols; c++) memcpy(coeffs_out + c * n, b.coeffs[c].data(), n * 8);
    if (leaves_out) memcpy(leaves_out, b.tree.leaves.data(), b.tree.leaves.size() * 8);
    if (digests_out) memcpy(digests_out, b.tree.digests.data(), b.tree.digests.size() * 32);
    if (cap_out) memcpy(cap_out, b.tree.cap.data(), b.tree.cap.size() * 32);
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- Challenger handle ----
void* orc_challenger_new() { return new Challenger(); }
void orc_challenger_free(void* c) { delete (Challenger*)c; }
void orc_challenger_observe(void* c, const u64* x, u64 n) {
  for (u64 i = 0; i < n; i++) ((Challenger*)c)->observe(canon(x[i]));
}
u64 orc_challenger_get(void* c) { return ((Challenger*)c)->get_challenge(); }

// ---- verifier ----
// common: CommonCircuitData bytes; vonly: VerifierOnlyCircuitData bytes (may be followed by
// trailing bytes, e.g. verifier.bin = vonly || common). Returns VerifyCode, or -1 on parse error.
// challenges_out (optional, >= 16 + num_query_rounds u64): betas[2] gammas[2] alphas[2] zeta[2]
// fri_alpha[2] pow_response, then query indices.
int orc_verify(const uint8_t* common, u64 common_len, const uint8_t* vonly, u64 vonly_len,
               const uint8_t* proof, u64 proof_len, u64* challenges_out) {
  try {
    CommonData c = parse_common(common, common_len);
    size_t used = 0;
    VerifierOnly vo = parse_verifier_only(vonly, vonly_len, &used);
    Proof pf = parse_proof(c, proof, proof_len);
    Challenges ch;
    int rc = verify_proof(c, vo, pf, &ch);
    if (challenges_out) {
      u64* o = challenges_out;
      for (u64 v : ch.betas) *o++ = v;
      for (u64 v : ch.gammas) *o++ = v;
      for (u64 v : ch.alphas) *o++ = v;
      *o++ = ch.zeta.a; *o++ = ch.zeta.b;
      *o++ = ch.fri_alpha.a; *o++ = ch.fri_alpha.b;
      *o++ = ch.pow_response;
      for (size_t q : ch.query_indices) *o++ = q;
    }
    return rc;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// parse -> serialise round trip; returns 1 when the bytes are reproduced exactly.
int orc_proof_roundtrip(const uint8_t* common, u64 common_len, const uint8_t* proof, u64 proof_len) {
  try {
    CommonData c = parse_common(common, common_len);
    Proof pf = parse_proof(c, proof, proof_len);
    std::vector<uint8_t> b = proof_to_bytes(pf);
    return b.size() == proof_len && memcmp(b.data(), proof, proof_len) == 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Only the Merkle openings of a proof against the caps INSIDE the proof (oracles 1..3 and the FRI
// round trees); used for the non-ZK dummy proofs whose verifier data is not shipped.
// x_indices: one per query round, supplied by the caller. Returns number of verified paths or -1.
int orc_check_proof_paths(const uint8_t* common, u64 common_len, const uint8_t* proof,
                          u64 proof_len, const u64* x_indices) {
  try {
    CommonData c = parse_common(common, common_len);
    Proof pf = parse_proof(c, proof, proof_len);
    const std::vector<Hash>* caps[4] = {nullptr, &pf.wires_cap, &pf.zs_cap, &pf.quotient_cap};
    int ok = 0;
    for (size_t q = 0; q < pf.queries.size(); q++) {
      size_t x = x_indices[q];
      for (int o = 1; o < 4; o++) {
        auto& io = pf.queries[q].init[o];
        if (!merkle_verify(io.evals.data(), io.evals.size(), x, caps[o]->data(), io.path.data(), io.path.size()))
          return -2;
        ok++;
      }
      for (size_t s = 0; s < pf.queries[q].steps.size(); s++) {
        x >>= c.reduction_arity_bits[s];
        auto& st = pf.queries[q].steps[s];
        std::vector<u64> flat;
        for (E2 e : st.evals) { flat.push_back(e.a); flat.push_back(e.b); }
        if (!merkle_verify(flat.data(), flat.size(), x, pf.fri_caps[s].data(), st.path.data(), st.path.size()))
          return -3;
        ok++;
      }
    }
    return ok;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- prover (restated prove()) ----
struct OrcCircuit {
  CircuitProverData d;
  ProveTrace trace;
};

// constants_sigmas: column-major [num_constants + num_routed][n] values on the subgroup.
void* orc_circuit_new(const uint8_t* common, u64 common_len, const u64* digest4,
                      const u64* constants_sigmas, unsigned threads) {
  try {
    CommonData c = parse_common(common, common_len);
    size_t n = (size_t)1 << c.degree_bits, cols = c.num_constants + c.num_routed_wires;
    std::vector<std::vector<u64>> cs(cols);
    for (size_t j = 0; j < cols; j++) cs[j].assign(constants_sigmas + j * n, constants_sigmas + (j + 1) * n);
    Hash dg;
    memcpy(dg.e, digest4, 32);
    OrcCircuit* oc = new OrcCircuit();
    oc->d = make_circuit(c, dg, std::move(cs), threads);
    return oc;
  } catch (std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_circuit_free(void* h) { delete (OrcCircuit*)h; }
// VerifierOnlyCircuitData bytes: cap_height, cap, circuit_digest. Returns length.
u64 orc_circuit_verifier_only(void* h, uint8_t* out, u64 cap) {
  OrcCircuit* oc = (OrcCircuit*)h;
  Writer w;
  w.usize(oc->d.common.cap_height);
  for (auto& hh : oc->d.cs_batch.tree.cap) w.hash(hh);
  w.hash(oc->d.circuit_digest);
  if (out && w.b.size() <= cap) memcpy(out, w.b.data(), w.b.size());
  return w.b.size();
}
void orc_circuit_cs_coeffs(void* h, u64* out) {
  OrcCircuit* oc = (OrcCircuit*)h;
  size_t n = (size_t)1 << oc->d.common.degree_bits;
  for (size_t j = 0; j < oc->d.cs_batch.ncols; j++) memcpy(out + j * n, oc->d.cs_batch.coeffs[j].data(), n * 8);
}

// wires: column-major [num_wires][n]. salts_*: NULL or [4][N] natural order. Writes the serialised
// ProofWithPublicInputs; returns its length, or -1 (error text in orc_last_error).
int64_t orc_prove(void* h, const u64* wires, const u64* public_inputs, u64 npi, const u64* salts_w,
                  const u64* salts_z, const u64* salts_q, unsigned threads, uint8_t* out, u64 out_cap) {
  try {
    OrcCircuit* oc = (OrcCircuit*)h;
    const CommonData& c = oc->d.common;
    size_t n = (size_t)1 << c.degree_bits;
    std::vector<std::vector<u64>> w(c.num_wires);
    for (size_t j = 0; j < c.num_wires; j++) {
      w[j].assign(wires + j * n, wires + (j + 1) * n);
      for (auto& v : w[j]) v = canon(v);
    }
    std::vector<u64> pi(public_inputs, public_inputs + npi);
    ProverSalts sl;
    sl.wires = salts_w; sl.zs_pp = salts_z; sl.quotient = salts_q;
    Proof pf = prove(oc->d, w, pi, sl, threads, &oc->trace);
    std::vector<uint8_t> b = proof_to_bytes(pf);
    if (b.size() > out_cap) throw std::runtime_error("output buffer too small");
    memcpy(out, b.data(), b.size());
    return (int64_t)b.size();
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// Intermediate values of the last orc_prove on this circuit, for stage-by-stage parity tests.
// challenges: betas[nch] gammas[nch] alphas[nch] zeta[2] fri_alpha[2] fri_betas[2*rounds]
void orc_trace_challenges(void* h, u64* out) {
  OrcCircuit* oc = (OrcCircuit*)h;
  u64* o = out;
  for (u64 v : oc->trace.betas) *o++ = v;
  for (u64 v : oc->trace.gammas) *o++ = v;
  for (u64 v : oc->trace.alphas) *o++ = v;
  *o++ = oc->trace.zeta.a; *o++ = oc->trace.zeta.b;
  *o++ = oc->trace.fri_alpha.a; *o++ = oc->trace.fri_alpha.b;
  for (E2 e : oc->trace.fri_betas) { *o++ = e.a; *o++ = e.b; }
}
void orc_trace_zs_pp(void* h, u64* out) {  // [nch*(1+npp)][n]
  OrcCircuit* oc = (OrcCircuit*)h;
  size_t n = (size_t)1 << oc->d.common.degree_bits;
  for (size_t j = 0; j < oc->trace.zs_pp_values.size(); j++) memcpy(out + j * n, oc->trace.zs_pp_values[j].data(), n * 8);
}
void orc_trace_quotient_chunks(void* h, u64* out) {  // [nch*qdf][n]
  OrcCircuit* oc = (OrcCircuit*)h;
  size_t n = (size_t)1 << oc->d.common.degree_bits;
  for (size_t j = 0; j < oc->trace.quotient_chunks.size(); j++) memcpy(out + j * n, oc->trace.quotient_chunks[j].data(), n * 8);
}
void orc_trace_final_poly(void* h, u64* out) {  // [n][2]
  OrcCircuit* oc = (OrcCircuit*)h;
  for (size_t m = 0; m < oc->trace.final_poly_coeffs_initial.size(); m++) {
    out[2 * m] = oc->trace.final_poly_coeffs_initial[m].a;
    out[2 * m + 1] = oc->trace.final_poly_coeffs_initial[m].b;
  }
}

int orc_have_avx512(void) { return orc::have_avx512() ? 1 : 0; }
}  // extern "C"
