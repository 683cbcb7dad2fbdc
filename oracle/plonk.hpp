// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Plonk-side data model, binary (de)serialisation and gate constraint evaluation of
// qp-plonky2 1.1.1, restated per SURVEY.md App. A.7 (vanishing identity) and App. B (byte layouts
// of `CommonCircuitData`, `VerifierOnlyCircuitData`, `ProofWithPublicInputs`). These formats are
// what /root/reference/wormhole/prover/src/lib.rs:114-121,
// /root/reference/wormhole/verifier/src/lib.rs:102-106 and
// /root/reference/wormhole/verifier/benches/verifier.rs:22-25 read and write.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "gl.hpp"
#include "poseidon.hpp"

namespace orc {

// ---- field wrappers so gate evaluation is written once for F_p and F_p^2 ----
struct Fp {
  u64 v;
};
static inline Fp operator+(Fp a, Fp b) { return Fp{add(a.v, b.v)}; }
static inline Fp operator-(Fp a, Fp b) { return Fp{sub(a.v, b.v)}; }
static inline Fp operator*(Fp a, Fp b) { return Fp{mul(a.v, b.v)}; }
template <class T> struct FieldOps;
template <> struct FieldOps<Fp> {
  static Fp from(u64 x) { return Fp{canon(x)}; }
};
template <> struct FieldOps<E2> {
  static E2 from(u64 x) { return E2{canon(x), 0}; }
};

enum GateId : uint32_t {
  GATE_ARITHMETIC = 0,
  GATE_BASE_SUM_2 = 2,
  GATE_CONSTANT = 3,
  GATE_NOOP = 9,
  GATE_POSEIDON = 11,
  GATE_PUBLIC_INPUT = 12,
};
struct GateInfo {
  uint32_t id;
  u64 param;  // num_consts / num_limbs / num_ops; 0 when the gate has none
};

struct CommonData {
  u64 num_wires, num_routed_wires, cfg_num_constants, security_bits, num_challenges, max_qdf;
  bool use_base_arithmetic_gate, zero_knowledge;
  u64 rate_bits, cap_height, num_query_rounds;
  uint32_t pow_bits;
  u64 strategy_arity_bits, strategy_final_poly_bits;
  std::vector<u64> reduction_arity_bits;
  u64 degree_bits;
  bool hiding;
  std::vector<u64> selector_indices;
  std::vector<std::pair<u64, u64>> groups;
  u64 quotient_degree_factor, num_gate_constraints, num_constants, num_public_inputs;
  std::vector<u64> k_is;
  u64 num_partial_products;
  std::vector<GateInfo> gates;
  u64 salt_size() const { return hiding ? 4 : 0; }
  u64 lde_bits() const { return degree_bits + rate_bits; }
};

struct Reader {
  const uint8_t* p;
  size_t n, off = 0;
  Reader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  void need(size_t k) {
    if (off + k > n) throw std::runtime_error("short read at " + std::to_string(off));
  }
  uint8_t u8() { need(1); return p[off++]; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p + off, 4); off += 4; return v; }
  u64 usize() { need(8); u64 v; memcpy(&v, p + off, 8); off += 8; return v; }
  u64 felt() {
    u64 v = usize();
    if (v >= P) throw std::runtime_error("non-canonical field element");
    return v;
  }
  E2 ext() { u64 a = felt(); u64 b = felt(); return E2{a, b}; }
  Hash hash() { Hash h; for (int i = 0; i < 4; i++) h.e[i] = felt(); return h; }
  bool boolean() { return u8() != 0; }
};
struct Writer {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void usize(u64 v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void felt(u64 v) { usize(canon(v)); }
  void ext(E2 x) { felt(x.a); felt(x.b); }
  void hash(const Hash& h) { for (int i = 0; i < 4; i++) felt(h.e[i]); }
};

static inline void read_fri_config(Reader& r, CommonData& c) {
  c.rate_bits = r.usize();
  c.cap_height = r.usize();
  c.num_query_rounds = r.usize();
  c.pow_bits = r.u32();
  uint8_t tag = r.u8();
  if (tag != 1) throw std::runtime_error("unsupported FRI reduction strategy tag");
  c.strategy_arity_bits = r.usize();
  c.strategy_final_poly_bits = r.usize();
}

static inline CommonData parse_common(const uint8_t* p, size_t n, size_t* consumed = nullptr) {
  Reader r(p, n);
  CommonData c;
  c.num_wires = r.usize();
  c.num_routed_wires = r.usize();
  c.cfg_num_constants = r.usize();
  c.security_bits = r.usize();
  c.num_challenges = r.usize();
  c.max_qdf = r.usize();
  c.use_base_arithmetic_gate = r.boolean();
  c.zero_knowledge = r.boolean();
  read_fri_config(r, c);
  read_fri_config(r, c);  // FriParams repeats the config
  u64 nar = r.usize();
  for (u64 i = 0; i < nar; i++) c.reduction_arity_bits.push_back(r.usize());
  c.degree_bits = r.usize();
  c.hiding = r.boolean();
  u64 nsel = r.usize();
  for (u64 i = 0; i < nsel; i++) c.selector_indices.push_back(r.usize());
  u64 ngroups = r.usize();
  for (u64 i = 0; i < ngroups; i++) {
    u64 a = r.usize(), b = r.usize();
    c.groups.push_back({a, b});
  }
  c.quotient_degree_factor = r.usize();
  c.num_gate_constraints = r.usize();
  c.num_constants = r.usize();
  c.num_public_inputs = r.usize();
  u64 nk = r.usize();
  for (u64 i = 0; i < nk; i++) c.k_is.push_back(r.felt());
  c.num_partial_products = r.usize();
  u64 num_lookup_polys = r.usize(), num_lookup_selectors = r.usize(), nluts = r.usize();
  if (num_lookup_polys || num_lookup_selectors || nluts)
    throw std::runtime_error("lookup tables are not supported");
  u64 ngates = r.usize();
  for (u64 i = 0; i < ngates; i++) {
    GateInfo g{r.u32(), 0};
    switch (g.id) {
      case GATE_NOOP: case GATE_PUBLIC_INPUT: case GATE_POSEIDON: break;
      case GATE_CONSTANT: case GATE_BASE_SUM_2: case GATE_ARITHMETIC: g.param = r.usize(); break;
      default: throw std::runtime_error("unsupported gate id " + std::to_string(g.id));
    }
    c.gates.push_back(g);
  }
  if (c.gates.size() != c.selector_indices.size()) throw std::runtime_error("selector/gate mismatch");
  if (consumed) *consumed = r.off;
  return c;
}

struct VerifierOnly {
  std::vector<Hash> constants_sigmas_cap;
  Hash circuit_digest;
};
static inline VerifierOnly parse_verifier_only(const uint8_t* p, size_t n, size_t* consumed) {
  Reader r(p, n);
  VerifierOnly v;
  u64 h = r.usize();
  if (h > 32) throw std::runtime_error("bad cap height");
  for (u64 i = 0; i < ((u64)1 << h); i++) v.constants_sigmas_cap.push_back(r.hash());
  v.circuit_digest = r.hash();
  if (consumed) *consumed = r.off;
  return v;
}

struct FriInitialOpen {
  std::vector<u64> evals;
  std::vector<Hash> path;
};
struct FriStep {
  std::vector<E2> evals;
  std::vector<Hash> path;
};
struct FriQueryRound {
  std::vector<FriInitialOpen> init;
  std::vector<FriStep> steps;
};
struct Proof {
  std::vector<Hash> wires_cap, zs_cap, quotient_cap;
  std::vector<E2> constants, sigmas, wires, zs, zs_next, partial_products, quotient;
  std::vector<std::vector<Hash>> fri_caps;
  std::vector<FriQueryRound> queries;
  std::vector<E2> final_poly;
  u64 pow_witness = 0;
  std::vector<u64> public_inputs;
};

static inline std::vector<size_t> oracle_widths(const CommonData& c) {
  size_t s = c.salt_size();
  return {(size_t)(c.num_constants + c.num_routed_wires), (size_t)c.num_wires + s,
          (size_t)(c.num_challenges * (1 + c.num_partial_products)) + s,
          (size_t)(c.num_challenges * c.quotient_degree_factor) + s};
}

static inline Proof parse_proof(const CommonData& c, const uint8_t* p, size_t n) {
  Reader r(p, n);
  Proof pf;
  size_t ncap = (size_t)1 << c.cap_height;
  auto cap = [&](std::vector<Hash>& v) { for (size_t i = 0; i < ncap; i++) v.push_back(r.hash()); };
  auto exts = [&](std::vector<E2>& v, size_t k) { for (size_t i = 0; i < k; i++) v.push_back(r.ext()); };
  auto path = [&](std::vector<Hash>& v) { size_t k = r.u8(); for (size_t i = 0; i < k; i++) v.push_back(r.hash()); };
  cap(pf.wires_cap);
  cap(pf.zs_cap);
  cap(pf.quotient_cap);
  exts(pf.constants, c.num_constants);
  exts(pf.sigmas, c.num_routed_wires);
  exts(pf.wires, c.num_wires);
  exts(pf.zs, c.num_challenges);
  exts(pf.zs_next, c.num_challenges);
  exts(pf.partial_products, c.num_challenges * c.num_partial_products);
  exts(pf.quotient, c.num_challenges * c.quotient_degree_factor);
  pf.fri_caps.resize(c.reduction_arity_bits.size());
  for (auto& v : pf.fri_caps) cap(v);
  std::vector<size_t> widths = oracle_widths(c);
  pf.queries.resize(c.num_query_rounds);
  for (auto& q : pf.queries) {
    q.init.resize(widths.size());
    for (size_t o = 0; o < widths.size(); o++) {
      for (size_t i = 0; i < widths[o]; i++) q.init[o].evals.push_back(r.felt());
      path(q.init[o].path);
    }
    q.steps.resize(c.reduction_arity_bits.size());
    for (size_t s = 0; s < q.steps.size(); s++) {
      exts(q.steps[s].evals, (size_t)1 << c.reduction_arity_bits[s]);
      path(q.steps[s].path);
    }
  }
  u64 total_arity = 0;
  for (u64 a : c.reduction_arity_bits) total_arity += a;
  exts(pf.final_poly, (size_t)1 << (c.degree_bits - total_arity));
  pf.pow_witness = r.felt();
  u64 npi = r.usize();
  for (u64 i = 0; i < npi; i++) pf.public_inputs.push_back(r.felt());
  if (r.off != n) throw std::runtime_error("trailing bytes after proof");
  return pf;
}

static inline std::vector<uint8_t> proof_to_bytes(const Proof& pf) {
  Writer w;
  auto cap = [&](const std::vector<Hash>& v) { for (auto& h : v) w.hash(h); };
  auto exts = [&](const std::vector<E2>& v) { for (auto& e : v) w.ext(e); };
  auto path = [&](const std::vector<Hash>& v) { w.u8((uint8_t)v.size()); for (auto& h : v) w.hash(h); };
  cap(pf.wires_cap); cap(pf.zs_cap); cap(pf.quotient_cap);
  exts(pf.constants); exts(pf.sigmas); exts(pf.wires); exts(pf.zs); exts(pf.zs_next);
  exts(pf.partial_products); exts(pf.quotient);
  for (auto& c : pf.fri_caps) cap(c);
  for (auto& q : pf.queries) {
    for (auto& o : q.init) { for (u64 v : o.evals) w.felt(v); path(o.path); }
    for (auto& s : q.steps) { exts(s.evals); path(s.path); }
  }
  exts(pf.final_poly);
  w.felt(pf.pow_witness);
  w.usize(pf.public_inputs.size());
  for (u64 v : pf.public_inputs) w.felt(v);
  return w.b;
}

// ---- gate constraints ----
template <class T> static inline T sbox_t(T x) {
  T x2 = x * x, x4 = x2 * x2, x3 = x * x2;
  return x3 * x4;
}
template <class T> static inline void mds_layer_t(T* s) {
  T o[12];
  for (int r = 0; r < 12; r++) {
    T acc = FieldOps<T>::from(0);
    for (int i = 0; i < 12; i++) acc = acc + s[(i + r) % 12] * FieldOps<T>::from(MDS_CIRC[i]);
    acc = acc + s[r] * FieldOps<T>::from(MDS_DIAG[r]);
    o[r] = acc;
  }
  for (int i = 0; i < 12; i++) s[i] = o[i];
}

// Appends the PoseidonGate's 123 constraints to out[0..123).
template <class T> static inline void poseidon_gate_constraints(const T* w, T* out) {
  typedef FieldOps<T> FO;
  const PoseidonTables& TB = tables();
  int k = 0;
  T swap = w[24];
  out[k++] = swap * (swap - FO::from(1));
  for (int i = 0; i < 4; i++) out[k++] = swap * (w[i + 4] - w[i]) - w[25 + i];
  T s[12];
  for (int i = 0; i < 4; i++) {
    s[i] = w[i] + w[25 + i];
    s[i + 4] = w[i + 4] - w[25 + i];
  }
  for (int i = 8; i < 12; i++) s[i] = w[i];
  int round = 0;
  for (int r = 0; r < HALF_N_FULL; r++, round++) {
    for (int i = 0; i < 12; i++) s[i] = s[i] + FO::from(TB.rc[12 * round + i]);
    if (r != 0)
      for (int i = 0; i < 12; i++) {
        T in = w[29 + 12 * (r - 1) + i];
        out[k++] = s[i] - in;
        s[i] = in;
      }
    for (int i = 0; i < 12; i++) s[i] = sbox_t(s[i]);
    mds_layer_t(s);
  }
  for (int i = 0; i < 12; i++) s[i] = s[i] + FO::from(TB.fast_first[i]);
  {
    T o[12];
    o[0] = s[0];
    for (int c = 1; c < 12; c++) o[c] = FO::from(0);
    for (int r = 1; r < 12; r++)
      for (int c = 1; c < 12; c++) o[c] = o[c] + s[r] * FO::from(TB.fast_init[r - 1][c - 1]);
    for (int i = 0; i < 12; i++) s[i] = o[i];
  }
  for (int r = 0; r < N_PARTIAL; r++) {
    T in = w[65 + r];
    out[k++] = s[0] - in;
    s[0] = sbox_t(in);
    if (r < N_PARTIAL - 1) s[0] = s[0] + FO::from(TB.fast_rc[r]);
    T d = s[0] * FO::from(MDS_CIRC[0] + MDS_DIAG[0]);
    for (int i = 1; i < 12; i++) d = d + s[i] * FO::from(TB.fast_w_hat[r][i - 1]);
    T s0 = s[0];
    for (int i = 1; i < 12; i++) s[i] = s[i] + s0 * FO::from(TB.fast_v[r][i - 1]);
    s[0] = d;
  }
  round += N_PARTIAL;
  for (int r = 0; r < HALF_N_FULL; r++, round++) {
    for (int i = 0; i < 12; i++) s[i] = s[i] + FO::from(TB.rc[12 * round + i]);
    for (int i = 0; i < 12; i++) {
      T in = w[87 + 12 * r + i];
      out[k++] = s[i] - in;
      s[i] = in;
    }
    for (int i = 0; i < 12; i++) s[i] = sbox_t(s[i]);
    mds_layer_t(s);
  }
  for (int i = 0; i < 12; i++) out[k++] = s[i] - w[12 + i];
}

// evaluate_gate_constraints: out[num_gate_constraints] = sum_g filter_g * constraint_{g,slot}.
// local_constants has num_constants entries (selectors first), local_wires num_wires.
template <class T>
static inline void eval_gate_constraints(const CommonData& c, const T* local_constants,
                                         const T* local_wires, const u64* pi_hash, T* out) {
  typedef FieldOps<T> FO;
  size_t num_selectors = c.groups.size();
  for (size_t j = 0; j < c.num_gate_constraints; j++) out[j] = FO::from(0);
  const T* gc = local_constants + num_selectors;  // gate-local constants
  std::vector<T> tmp(c.num_gate_constraints, FO::from(0));
  for (size_t g = 0; g < c.gates.size(); g++) {
    size_t si = c.selector_indices[g];
    T s = local_constants[si];
    T filter = FO::from(1);
    for (u64 j = c.groups[si].first; j < c.groups[si].second; j++)
      if (j != g) filter = filter * (FO::from(j) - s);
    if (num_selectors > 1) filter = filter * (FO::from(0xFFFFFFFFULL) - s);
    size_t nc = 0;
    const GateInfo& gi = c.gates[g];
    switch (gi.id) {
      case GATE_NOOP: break;
      case GATE_CONSTANT:
        for (u64 i = 0; i < gi.param; i++) tmp[nc++] = gc[i] - local_wires[i];
        break;
      case GATE_PUBLIC_INPUT:
        for (int i = 0; i < 4; i++) tmp[nc++] = local_wires[i] - FO::from(pi_hash[i]);
        break;
      case GATE_BASE_SUM_2: {
        T sum = FO::from(0);
        for (u64 j = gi.param; j-- > 0;) sum = sum * FO::from(2) + local_wires[1 + j];
        tmp[nc++] = sum - local_wires[0];
        for (u64 j = 0; j < gi.param; j++) {
          T l = local_wires[1 + j];
          tmp[nc++] = l * (l - FO::from(1));
        }
        break;
      }
      case GATE_ARITHMETIC:
        for (u64 i = 0; i < gi.param; i++) {
          const T* w = local_wires + 4 * i;
          tmp[nc++] = w[3] - (w[0] * w[1] * gc[0] + w[2] * gc[1]);
        }
        break;
      case GATE_POSEIDON:
        poseidon_gate_constraints(local_wires, tmp.data());
        nc = 123;
        break;
      default: throw std::runtime_error("unsupported gate");
    }
    for (size_t j = 0; j < nc; j++) out[j] = out[j] + filter * tmp[j];
  }
}

}  // namespace orc
