// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Plonk-side data model, binary (de)serialisation and gate constraint evaluation of
// qp-plonky2 1.1.1, restated per SURVEY.md App. A.7 (vanishing identity) and App. B (byte layouts
// of `CommonCircuitData`, `VerifierOnlyCircuitData`, `ProofWithPublicInputs`). These formats are
// what /root/reference/wormhole/prover/src/lib.rs:114-121,
// /root/reference/wormhole/verifier/src/lib.rs:102-106 and
// /root/reference/wormhole/verifier/benches/verifier.rs:22-25 read and write.
#pragma once
#include <algorithm>
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

// Gate ids = position in plonky2's default gate serializer list (the six wormhole ids are read off
// wormhole/bench-data/common.bin; the recursion gates follow the same alphabetical list).
enum GateId : uint32_t {
  GATE_ARITHMETIC = 0,
  GATE_ARITHMETIC_EXT = 1,
  GATE_BASE_SUM_2 = 2,
  GATE_CONSTANT = 3,
  GATE_COSET_INTERPOLATION = 4,
  GATE_EXPONENTIATION = 5,
  GATE_MUL_EXT = 8,
  GATE_NOOP = 9,
  GATE_POSEIDON_MDS = 10,
  GATE_POSEIDON = 11,
  GATE_PUBLIC_INPUT = 12,
  GATE_RANDOM_ACCESS = 13,
  GATE_REDUCING_EXT = 14,
  GATE_REDUCING = 15,
};
struct GateInfo {
  uint32_t id = 0;
  u64 param = 0;   // num_consts / num_limbs / num_ops / num_coeffs / num_power_bits / bits / subgroup_bits
  u64 param2 = 0;  // RandomAccess: num_copies; CosetInterpolation: degree
  u64 param3 = 0;  // RandomAccess: num_extra_constants
  std::vector<u64> weights;  // CosetInterpolation: barycentric weights
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
    GateInfo g;
    g.id = r.u32();
    switch (g.id) {
      case GATE_NOOP: case GATE_PUBLIC_INPUT: case GATE_POSEIDON: case GATE_POSEIDON_MDS: break;
      case GATE_CONSTANT: case GATE_BASE_SUM_2: case GATE_ARITHMETIC: case GATE_ARITHMETIC_EXT:
      case GATE_MUL_EXT: case GATE_REDUCING: case GATE_REDUCING_EXT: case GATE_EXPONENTIATION:
        g.param = r.usize();
        break;
      case GATE_RANDOM_ACCESS:
        g.param = r.usize(); g.param2 = r.usize(); g.param3 = r.usize();
        break;
      case GATE_COSET_INTERPOLATION: {
        g.param = r.usize(); g.param2 = r.usize();
        u64 nw = r.usize();
        if (nw != ((u64)1 << g.param)) throw std::runtime_error("coset interpolation: weight count");
        for (u64 j = 0; j < nw; j++) g.weights.push_back(r.felt());
        break;
      }
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

// ---- the recursion gate set (restated from upstream plonky2's gates/*.rs - qp-plonky2 is
// un-vendored and the reference ships no aggregator circuit data, so unlike the six wormhole gates
// these are NOT pinned by a fixture: "parity unpinned" until a cargo-side comparison can run) ----
// ExtensionAlgebra over T: pairs (a, b) = a + b*X with X^2 = 7, components in T (T = F_p on the prover's
// coset, T = F_p^2 at zeta in the verifier).
template <class T> struct Alg {
  T a, b;
};
template <class T> static inline Alg<T> operator+(Alg<T> x, Alg<T> y) { return Alg<T>{x.a + y.a, x.b + y.b}; }
template <class T> static inline Alg<T> operator-(Alg<T> x, Alg<T> y) { return Alg<T>{x.a - y.a, x.b - y.b}; }
template <class T> static inline Alg<T> operator*(Alg<T> x, Alg<T> y) {
  T seven = FieldOps<T>::from(7);
  return Alg<T>{x.a * y.a + seven * (x.b * y.b), x.a * y.b + x.b * y.a};
}
template <class T> static inline Alg<T> alg_scale(Alg<T> x, T s) { return Alg<T>{x.a * s, x.b * s}; }
template <class T> static inline Alg<T> alg_at(const T* w, size_t i) { return Alg<T>{w[i], w[i + 1]}; }
template <class T> static inline Alg<T> alg_const(u64 v) { return Alg<T>{FieldOps<T>::from(v), FieldOps<T>::from(0)}; }
template <class T> static inline void alg_push(T* out, size_t& nc, Alg<T> x) { out[nc++] = x.a; out[nc++] = x.b; }

// CosetInterpolationGate layout helpers (D = 2)
struct CosetLayout {
  u64 npoints, degree, nint, start_eval_point, start_eval_value, start_int, start_shifted;
  CosetLayout(u64 subgroup_bits, u64 deg) {
    npoints = (u64)1 << subgroup_bits;
    degree = deg;
    nint = (npoints - 2) / (degree - 1);
    start_eval_point = 1 + 2 * npoints;
    start_eval_value = start_eval_point + 2;
    start_int = start_eval_value + 2;
    start_shifted = start_int + 4 * nint;
  }
  u64 num_wires() const { return start_shifted + 2; }
  u64 num_constraints() const { return 2 + 4 * nint + 2; }
};

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
      case GATE_ARITHMETIC_EXT:  // output - (m0*m1*c0 + addend*c1), 4*D wires per op
        for (u64 i = 0; i < gi.param; i++) {
          const T* w = local_wires + 8 * i;
          Alg<T> m0 = alg_at(w, 0), m1 = alg_at(w, 2), ad = alg_at(w, 4), o = alg_at(w, 6);
          alg_push(tmp.data(), nc, o - (alg_scale(m0 * m1, gc[0]) + alg_scale(ad, gc[1])));
        }
        break;
      case GATE_MUL_EXT:  // output - m0*m1*c0, 3*D wires per op
        for (u64 i = 0; i < gi.param; i++) {
          const T* w = local_wires + 6 * i;
          alg_push(tmp.data(), nc, alg_at(w, 4) - alg_scale(alg_at(w, 0) * alg_at(w, 2), gc[0]));
        }
        break;
      case GATE_POSEIDON_MDS:  // out_r - sum_i circ[i]*in[(i+r)%12] - diag[r]*in[r], on algebra elements
        for (int r = 0; r < 12; r++) {
          Alg<T> acc = alg_const<T>(0);
          for (int i = 0; i < 12; i++)
            acc = acc + alg_scale(alg_at(local_wires, 2 * ((i + r) % 12)), FO::from(MDS_CIRC[i]));
          acc = acc + alg_scale(alg_at(local_wires, 2 * r), FO::from(MDS_DIAG[r]));
          alg_push(tmp.data(), nc, alg_at(local_wires, 2 * (12 + r)) - acc);
        }
        break;
      case GATE_RANDOM_ACCESS: {
        const u64 bits = gi.param, copies = gi.param2, extra = gi.param3, vec = (u64)1 << bits;
        const u64 routed = (2 + vec) * copies + extra;
        for (u64 cp = 0; cp < copies; cp++) {
          const T* w = local_wires + (2 + vec) * cp;
          const T* b = local_wires + routed + cp * bits;
          for (u64 i = 0; i < bits; i++) tmp[nc++] = b[i] * (b[i] - FO::from(1));
          T rec = FO::from(0);
          for (u64 i = bits; i-- > 0;) rec = rec + rec + b[i];
          tmp[nc++] = rec - w[0];
          std::vector<T> items(w + 2, w + 2 + vec);
          for (u64 i = 0; i < bits; i++) {
            std::vector<T> nxt;
            for (size_t j = 0; j + 1 < items.size(); j += 2) nxt.push_back(items[j] + b[i] * (items[j + 1] - items[j]));
            items.swap(nxt);
          }
          tmp[nc++] = items[0] - w[1];
        }
        for (u64 i = 0; i < extra; i++) tmp[nc++] = gc[i] - local_wires[(2 + vec) * copies + i];
        break;
      }
      case GATE_REDUCING: case GATE_REDUCING_EXT: {  // acc*alpha + coeff_i - acc_i
        const bool ext = gi.id == GATE_REDUCING_EXT;
        const u64 ncf = gi.param, cw = ext ? 2 : 1;
        const u64 start_accs = 6 + ncf * cw;
        Alg<T> alpha = alg_at(local_wires, 2), acc = alg_at(local_wires, 4);
        for (u64 i = 0; i < ncf; i++) {
          Alg<T> cf = ext ? alg_at(local_wires, 6 + 2 * i) : Alg<T>{local_wires[6 + i], FO::from(0)};
          Alg<T> ai = i == ncf - 1 ? alg_at(local_wires, 0) : alg_at(local_wires, start_accs + 2 * i);
          alg_push(tmp.data(), nc, acc * alpha + cf - ai);
          acc = ai;
        }
        break;
      }
      case GATE_EXPONENTIATION: {
        const u64 nb = gi.param;
        T base = local_wires[0];
        const T* iv = local_wires + 2 + nb;
        for (u64 i = 0; i < nb; i++) {
          T prev = i == 0 ? FO::from(1) : iv[i - 1] * iv[i - 1];
          T bit = local_wires[1 + (nb - 1 - i)];   // bits are little-endian, accumulated big-endian
          tmp[nc++] = prev * (bit * base + (FO::from(1) - bit)) - iv[i];
        }
        tmp[nc++] = local_wires[1 + nb] - iv[nb - 1];
        break;
      }
      case GATE_COSET_INTERPOLATION: {
        CosetLayout L(gi.param, gi.param2);
        T shift = local_wires[0];
        Alg<T> x = alg_at(local_wires, L.start_eval_point), xs = alg_at(local_wires, L.start_shifted);
        alg_push(tmp.data(), nc, x - alg_scale(xs, shift));
        u64 g16 = root_of_unity((unsigned)gi.param);
        std::vector<u64> dom(L.npoints);
        dom[0] = 1;
        for (u64 i = 1; i < L.npoints; i++) dom[i] = mul(dom[i - 1], g16);
        // partial barycentric interpolation over points [lo, hi), continuing from (eval, prod)
        auto partial = [&](u64 lo, u64 hi, Alg<T> eval, Alg<T> prod, Alg<T>& oe, Alg<T>& op) {
          for (u64 i = lo; i < hi; i++) {
            Alg<T> term = xs - alg_const<T>(dom[i]);
            Alg<T> val = alg_at(local_wires, 1 + 2 * i);
            eval = eval * term + alg_scale(val * prod, FO::from(gi.weights[i]));
            prod = prod * term;
          }
          oe = eval;
          op = prod;
        };
        Alg<T> ce, cp;
        partial(0, L.degree, alg_const<T>(0), alg_const<T>(1), ce, cp);
        for (u64 i = 0; i < L.nint; i++) {
          Alg<T> ie = alg_at(local_wires, L.start_int + 2 * i), ip = alg_at(local_wires, L.start_int + 2 * (L.nint + i));
          alg_push(tmp.data(), nc, ie - ce);
          alg_push(tmp.data(), nc, ip - cp);
          u64 lo = 1 + (L.degree - 1) * (i + 1), hi = std::min(lo + L.degree - 1, L.npoints);
          partial(lo, hi, ie, ip, ce, cp);
        }
        alg_push(tmp.data(), nc, alg_at(local_wires, L.start_eval_value) - ce);
        break;
      }
      default: throw std::runtime_error("unsupported gate");
    }
    for (size_t j = 0; j < nc; j++) out[j] = out[j] + filter * tmp[j];
  }
}

}  // namespace orc
