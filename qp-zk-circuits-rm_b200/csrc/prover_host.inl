// Host-side driver of one proof: the order of operations of `ProverCircuitData::prove` in
// qp-plonky2 1.1.1 (plonk/prover.rs), with every data-parallel step on
// the device and only the Fiat-Shamir `Challenger` (a few dozen Poseidon permutations over caps and
// openings - SURVEY H15 "host") and the byte serialisation (H16) on the host.
// Included by qpzk.cu after the PolynomialBatch implementation.
//
// Reference entry points this replaces: /root/reference/wormhole/prover/src/lib.rs:233-237
// (`prove`), /root/reference/wormhole/circuit/src/circuit.rs:98-108 (`build`: the constants|sigmas
// commit done once in qpzk_circuit_create), /root/reference/wormhole/aggregator/src/circuits/tree.rs:127,136,
// /root/reference/voting/src/lib.rs:355-356.

namespace qpzk {

// ---- host Poseidon for the transcript (same tables the device uses) ----
// ~100 permutations per proof sit on the latency path between device stages, so this is written
// branch-free on any-u64 representatives (canonical only at the end), like the device code.
struct HostPoseidon {
  typedef unsigned __int128 u128;
  const PoseidonTablesHost* T;
  static inline u64 red(u128 x) {  // 2^64 == EPS, 2^96 == -1; neither correction can wrap twice
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    t0 -= ((u64)0 - (u64)(lo < hh)) & GL_EPS;
    u64 t1 = hl * GL_EPS;
    u64 t2 = t0 + t1;
    t2 += ((u64)0 - (u64)(t2 < t1)) & GL_EPS;
    return t2;
  }
  static inline u64 mulr(u64 a, u64 b) { return red((u128)a * b); }
  static inline u64 addr(u64 a, u64 b) {  // both operands any u64: the first correction may wrap once more
    u64 s = a + b;
    u64 s2 = s + (((u64)0 - (u64)(s < a)) & GL_EPS);
    return s2 + (((u64)0 - (u64)(s2 < s)) & GL_EPS);
  }
  static inline u64 canon(u64 a) { return a >= GL_P ? a - GL_P : a; }
  static inline u64 sbox(u64 x) {
    u64 x2 = mulr(x, x), x4 = mulr(x2, x2), x3 = mulr(x, x2);
    return mulr(x3, x4);
  }
  static void mds(u64* s) {
    u64 o[12];
    for (int r = 0; r < 12; r++) {
      u128 acc = 0;  // 12 terms of < 2^70: no overflow
      for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12] * kMdsCirc[i];
      if (r == 0) acc += (u128)s[0] * kMdsDiag0;
      o[r] = red(acc);
    }
    memcpy(s, o, sizeof o);
  }
  void permute(u64* s) const {
    for (int r = 0; r < 4; r++) {
      for (int i = 0; i < 12; i++) s[i] = sbox(addr(s[i], T->rc[12 * r + i]));
      mds(s);
    }
    for (int i = 0; i < 12; i++) s[i] = addr(s[i], T->fast_first[i]);
    u64 o[12] = {s[0]};
    for (int c = 1; c < 12; c++) {
      u64 acc = 0;
      for (int r = 1; r < 12; r++) acc = addr(acc, mulr(s[r], T->fast_init[(r - 1) * 11 + (c - 1)]));
      o[c] = acc;
    }
    memcpy(s, o, sizeof o);
    for (int r = 0; r < 22; r++) {
      u64 s0 = addr(sbox(s[0]), T->fast_rc[r]);
      u64 d = mulr(s0, 25);
      for (int i = 1; i < 12; i++) d = addr(d, mulr(s[i], T->fast_w_hat[r * 11 + i - 1]));
      for (int i = 1; i < 12; i++) s[i] = addr(s[i], mulr(s0, T->fast_v[r * 11 + i - 1]));
      s[0] = d;
    }
    for (int r = 0; r < 4; r++) {
      for (int i = 0; i < 12; i++) s[i] = sbox(addr(s[i], T->rc[12 * (26 + r) + i]));
      mds(s);
    }
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
  }
};

static const PoseidonTablesHost* host_tables() {
  static PoseidonTablesHost* T = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!T) {
    T = new PoseidonTablesHost();
    build_poseidon_tables(T);
  }
  return T;
}

// Duplex-sponge Challenger (iop/challenger.rs): buffer up to 8 inputs, OVERWRITE state[0..len),
// permute, outputs are state[0..8) popped from the end.
struct HostChallenger {
  HostPoseidon H{host_tables()};
  u64 state[12] = {0};
  std::vector<u64> in, out;
  void duplex() {
    for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
    in.clear();
    H.permute(state);
    out.assign(state, state + 8);
  }
  void observe(u64 x) {
    out.clear();
    in.push_back(x >= GL_P ? x - GL_P : x);
    if (in.size() == 8) duplex();
  }
  void observe_n(const u64* x, size_t n) {
    for (size_t i = 0; i < n; i++) observe(x[i]);
  }
  u64 get() {
    if (!in.empty() || out.empty()) duplex();
    u64 v = out.back();
    out.pop_back();
    return v;
  }
};

static void host_hash_no_pad(const u64* x, size_t n, u64* out4) {
  HostPoseidon H{host_tables()};
  u64 s[12] = {0};
  for (size_t off = 0; off < n; off += 8) {
    size_t len = n - off < 8 ? n - off : 8;
    for (size_t i = 0; i < len; i++) s[i] = x[off + i] >= GL_P ? x[off + i] - GL_P : x[off + i];
    H.permute(s);
  }
  memcpy(out4, s, 32);
}

// ---- CommonCircuitData (layout: SURVEY.md App. B) ----
struct CommonHost {
  u64 num_wires, num_routed, cfg_constants, security_bits, num_challenges, max_qdf;
  bool base_arith, zk;
  u64 rate_bits, cap_height, num_queries;
  u32 pow_bits;
  std::vector<u64> arities;
  u64 degree_bits;
  bool hiding;
  std::vector<u64> selector_indices;
  std::vector<std::pair<u64, u64>> groups;
  u64 qdf, num_gate_constraints, num_constants, num_public_inputs;
  std::vector<u64> k_is;
  u64 num_partial_products;
  std::vector<std::pair<u32, u64>> gates;
  std::vector<u64> gate_p2, gate_p3;        // RandomAccess copies / extra constants; CosetInterpolation degree
  std::vector<u64> coset_weights;           // CosetInterpolation barycentric weights
  bool recursion = false;                   // any gate outside the wormhole / voting set
};

struct ByteReader {
  const uint8_t* p;
  size_t n, off;
  bool ok;
  ByteReader(const uint8_t* p_, size_t n_) : p(p_), n(n_), off(0), ok(true) {}
  u64 u(size_t bytes) {
    if (off + bytes > n) {
      ok = false;
      return 0;
    }
    u64 v = 0;
    memcpy(&v, p + off, bytes);
    off += bytes;
    return v;
  }
};

static bool parse_common_host(const uint8_t* p, size_t n, CommonHost* c, std::string* err) {
  ByteReader r(p, n);
  c->num_wires = r.u(8); c->num_routed = r.u(8); c->cfg_constants = r.u(8); c->security_bits = r.u(8);
  c->num_challenges = r.u(8); c->max_qdf = r.u(8);
  c->base_arith = r.u(1) != 0; c->zk = r.u(1) != 0;
  for (int rep = 0; rep < 2; rep++) {  // FriConfig, then again inside FriParams
    c->rate_bits = r.u(8); c->cap_height = r.u(8); c->num_queries = r.u(8); c->pow_bits = (u32)r.u(4);
    u64 tag = r.u(1);
    if (tag != 1) { *err = "unsupported FRI reduction strategy"; return false; }
    r.u(8); r.u(8);
  }
  u64 na = r.u(8);
  if (na > 64) { *err = "bad arity list"; return false; }
  for (u64 i = 0; i < na; i++) c->arities.push_back(r.u(8));
  c->degree_bits = r.u(8);
  c->hiding = r.u(1) != 0;
  u64 ns = r.u(8);
  if (ns > QPZK_MAX_GATES) { *err = "too many gates"; return false; }
  for (u64 i = 0; i < ns; i++) c->selector_indices.push_back(r.u(8));
  u64 ng = r.u(8);
  if (ng > QPZK_MAX_GATES) { *err = "too many selector groups"; return false; }
  for (u64 i = 0; i < ng; i++) { u64 a = r.u(8), b = r.u(8); c->groups.push_back({a, b}); }
  c->qdf = r.u(8); c->num_gate_constraints = r.u(8); c->num_constants = r.u(8); c->num_public_inputs = r.u(8);
  u64 nk = r.u(8);
  if (nk > 4096) { *err = "bad k_is"; return false; }
  for (u64 i = 0; i < nk; i++) c->k_is.push_back(r.u(8));
  c->num_partial_products = r.u(8);
  u64 l0 = r.u(8), l1 = r.u(8), l2 = r.u(8);
  if (l0 || l1 || l2) { *err = "lookup tables are not supported"; return false; }
  u64 ngates = r.u(8);
  if (ngates != ns) { *err = "gate/selector count mismatch"; return false; }
  for (u64 i = 0; i < ngates; i++) {
    u32 id = (u32)r.u(4);
    u64 param = 0, p2 = 0, p3 = 0;
    switch (id) {
      case G_NOOP: case G_PUBLIC_INPUT: case G_POSEIDON: case G_POSEIDON_MDS: break;
      case G_CONSTANT: case G_BASE_SUM_2: case G_ARITHMETIC: case G_ARITHMETIC_EXT: case G_MUL_EXT:
      case G_REDUCING: case G_REDUCING_EXT: case G_EXPONENTIATION:
        param = r.u(8);
        break;
      case G_RANDOM_ACCESS:
        param = r.u(8); p2 = r.u(8); p3 = r.u(8);
        if (param > 6 || p3 > 2) { *err = "RandomAccessGate: bits > 6 or more than 2 extra constants"; return false; }
        break;
      case G_COSET_INTERPOLATION: {
        param = r.u(8); p2 = r.u(8);
        u64 nw = r.u(8);
        if (param > 6 || p2 < 2 || nw != (1ull << param) || !c->coset_weights.empty()) {
          *err = "CosetInterpolationGate: unsupported parameters";
          return false;
        }
        for (u64 j = 0; j < nw; j++) c->coset_weights.push_back(r.u(8));
        break;
      }
      default: *err = "unsupported gate id " + std::to_string(id) + " (lookup gates are not built)"; return false;
    }
    if (gate_is_recursion_only(id)) c->recursion = true;
    c->gates.push_back({id, param});
    c->gate_p2.push_back(p2);
    c->gate_p3.push_back(p3);
  }
  if (!r.ok) { *err = "truncated common data"; return false; }
  return true;
}

struct ByteWriter {
  std::vector<uint8_t> b;
  void u(u64 v, size_t bytes) { for (size_t i = 0; i < bytes; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void felts(const u64* x, size_t n) { for (size_t i = 0; i < n; i++) u(x[i], 8); }
};

}  // namespace qpzk

struct qpzk_circuit {
  qpzk_ctx* ctx;
  CommonHost common;
  CircuitDesc desc;
  u64 digest[4];
  u64* k_is_dev = nullptr;
  u64* coset_aux_dev = nullptr;
  u64* l0_den_inv_dev = nullptr;   // [2^(degree_bits + qdb)], see k_build_l0_den_inv
  u64* cs_values = nullptr;  // [num_constants + num_routed][n] values on the subgroup (for Z)
  qpzk_batch* cs_batch = nullptr;
  std::vector<u64> cs_cap;
  // debug trace of the last proof
  std::vector<u64> tr_challenges, tr_zs_pp, tr_quotient, tr_final_poly;
  float stage_ms[16] = {0};
};

namespace qpzk {

struct DevBuf {  // scoped stream-ordered allocation
  qpzk_ctx* c;
  u64* p = nullptr;
  explicit DevBuf(qpzk_ctx* c_) : c(c_) {}
  ~DevBuf() { dev_free(c, p); }
  int alloc(size_t bytes) { return dev_alloc(c, bytes, &p); }
};

}  // namespace qpzk

namespace qpzk {
// fri_proof_of_work: the smallest w such that the permutation of the sponge state with w written at
// `pos` has >= min_lz leading zero bits in output word 7. Candidate windows grow from the expected
// witness size (2^min_lz) upwards: a window much larger than that only burns permutations behind
// the witness before the early exit can see it.
static int grind_pow(qpzk_ctx* c, const PowState& ps, u32 pos, u32 min_lz, u64* witness) {
  DevBuf best(c);
  QP(best.alloc(8));
  unsigned long long init = ~0ull;
  CU(cudaMemcpyAsync(best.p, &init, 8, cudaMemcpyHostToDevice, c->stream));
  u64 batch = 1ull << (min_lz < 12 ? 12 : (min_lz > 20 ? 20 : min_lz));
  u64 start = 0;
  unsigned long long found = ~0ull;
  for (int round = 0; found == ~0ull; round++) {
    if (round >= 2 && batch < (1ull << 22)) batch <<= 1;
    k_pow_grind<<<(unsigned)(batch / 128), 128, 0, c->stream>>>(ps, pos, min_lz, start, batch,
                                                               (unsigned long long*)best.p);
    c->launches++;
    CU(cudaMemcpyAsync(&found, best.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    start += batch;
    if (start >= (1ull << 40)) return fail(QPZK_ERR_CUDA, "proof of work failed");
  }
  *witness = found;
  return QPZK_OK;
}
}  // namespace qpzk

extern "C" {

int qpzk_fri_pow(qpzk_ctx* c, const uint64_t* sponge_state, uint32_t input_pos, uint32_t min_leading_zeros,
                 uint64_t* witness_out) {
  if (!c || !sponge_state || !witness_out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (input_pos >= 12 || min_leading_zeros > 40) return fail(QPZK_ERR_BAD_ARG, "bad position or difficulty");
  CU(cudaSetDevice(c->device));
  PowState ps;
  memcpy(ps.s, sponge_state, sizeof ps.s);
  return grind_pow(c, ps, input_pos, min_leading_zeros, witness_out);
}

int qpzk_batch_eval_ext(const qpzk_batch* b, const uint64_t* point, uint64_t* out) {
  if (!b || !point || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_ctx* c = b->ctx;
  CU(cudaSetDevice(c->device));
  const u64 n = 1ull << b->degree_bits;
  DevBuf zpow(c), res(c);
  QP(zpow.alloc(n * 16));
  QP(res.alloc((size_t)b->ncols * 16));
  k_ext_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(gl2{point[0], point[1]}, n, zpow.p);
  k_eval_at_ext<<<b->ncols, 256, 0, c->stream>>>(b->coeffs, n, zpow.p, res.p);
  c->launches += 2;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, res.p, (size_t)b->ncols * 16, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_circuit_create(qpzk_ctx* c, const uint8_t* common_bytes, size_t common_len, const uint64_t* digest4,
                        const uint64_t* constants_sigmas, qpzk_circuit** out) {
  if (!c || !common_bytes || !digest4 || !constants_sigmas || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(c->device));
  qpzk_circuit* q = new qpzk_circuit();
  q->ctx = c;
  std::string err;
  if (!parse_common_host(common_bytes, common_len, &q->common, &err)) {
    delete q;
    return fail(QPZK_ERR_UNSUPPORTED, err);
  }
  const CommonHost& cm = q->common;
  u32 qdb = 0;
  while ((1ull << qdb) < cm.qdf) qdb++;
  if ((1ull << qdb) != cm.qdf || qdb > cm.rate_bits || cm.num_challenges > 2 || cm.num_partial_products + 1 > 12 ||
      cm.num_constants < cm.groups.size() || cm.k_is.size() != cm.num_routed) {
    delete q;
    return fail(QPZK_ERR_UNSUPPORTED, "unsupported circuit configuration");
  }
  CircuitDesc& d = q->desc;
  memset(&d, 0, sizeof d);
  d.degree_bits = (u32)cm.degree_bits; d.rate_bits = (u32)cm.rate_bits; d.quotient_degree_bits = qdb;
  d.num_wires = (u32)cm.num_wires; d.num_routed = (u32)cm.num_routed; d.num_constants = (u32)cm.num_constants;
  d.num_challenges = (u32)cm.num_challenges; d.num_partial_products = (u32)cm.num_partial_products;
  d.qdf = (u32)cm.qdf; d.num_selectors = (u32)cm.groups.size(); d.num_gates = (u32)cm.gates.size();
  d.num_gate_constraints = (u32)cm.num_gate_constraints;
  for (size_t g = 0; g < cm.gates.size(); g++) {
    d.gate_id[g] = cm.gates[g].first;
    d.gate_param[g] = (u32)cm.gates[g].second;
    d.gate_param2[g] = (u32)cm.gate_p2[g];
    d.gate_param3[g] = (u32)cm.gate_p3[g];
    d.gate_selector[g] = (u32)cm.selector_indices[g];
  }
  for (size_t s = 0; s < cm.groups.size(); s++) {
    d.group_lo[s] = (u32)cm.groups[s].first;
    d.group_hi[s] = (u32)cm.groups[s].second;
  }
  memcpy(q->digest, digest4, 32);
  const u64 n = 1ull << cm.degree_bits;
  const u32 ncs = (u32)(cm.num_constants + cm.num_routed);
  QP(dev_alloc(c, cm.k_is.size() * 8, &q->k_is_dev));
  CU(cudaMemcpyAsync(q->k_is_dev, cm.k_is.data(), cm.k_is.size() * 8, cudaMemcpyHostToDevice, c->stream));
  if (!cm.coset_weights.empty()) {  // subgroup points, then the barycentric weights from the common data
    const size_t np = cm.coset_weights.size();
    u32 bits = 0;
    while ((1ull << bits) < np) bits++;
    std::vector<u64> aux(2 * np);
    u64 g = glh::root_of_unity(bits);
    aux[0] = 1;
    for (size_t i = 1; i < np; i++) aux[i] = glh::mul(aux[i - 1], g);
    for (size_t i = 0; i < np; i++) aux[np + i] = cm.coset_weights[i];
    QP(dev_alloc(c, aux.size() * 8, &q->coset_aux_dev));
    CU(cudaMemcpyAsync(q->coset_aux_dev, aux.data(), aux.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(ctx_wait(c));
    d.coset_aux = q->coset_aux_dev;
  }
  QP(dev_alloc(c, (size_t)ncs * n * 8, &q->cs_values));
  CU(cudaMemcpyAsync(q->cs_values, constants_sigmas, (size_t)ncs * n * 8, cudaMemcpyHostToDevice, c->stream));
  // build(): PolynomialBatch::from_values(constants | sigmas), never blinded
  int rc = commit_impl(c, q->cs_values, false, false, ncs, (u32)cm.degree_bits, (u32)cm.rate_bits, (u32)cm.cap_height,
                       nullptr, false, 0, &q->cs_batch);
  if (rc != QPZK_OK) {
    dev_free(c, q->k_is_dev);
    dev_free(c, q->cs_values);
    delete q;
    return rc;
  }
  q->cs_cap.resize(4ull << cm.cap_height);
  QP(qpzk_batch_cap(q->cs_batch, q->cs_cap.data()));
  {
    const u32 qlb = (u32)cm.degree_bits + qdb;
    if (cm.num_challenges * (1 + cm.num_partial_products + 1) + cm.num_gate_constraints > QPZK_APW_STRIDE) {
      qpzk_circuit_free(q);
      return fail(QPZK_ERR_UNSUPPORTED, "too many constraint terms for the alpha-power table");
    }
    RootTab tab_q;
    QP(get_root_tab(c, (int)qlb, false, &tab_q));
    QP(dev_alloc(c, sizeof(u64) << qlb, &q->l0_den_inv_dev));
    k_build_l0_den_inv<<<(unsigned)(((1ull << qlb) + 127) / 128), 128, 0, c->stream>>>(q->l0_den_inv_dev, tab_q,
                                                                                    (u32)cm.degree_bits, qlb);
    c->launches++;
    CU(cudaGetLastError());
    CU(ctx_wait(c));
  }
  *out = q;
  return QPZK_OK;
}

int qpzk_circuit_cap(const qpzk_circuit* q, uint64_t* out) {
  if (!q || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  memcpy(out, q->cs_cap.data(), q->cs_cap.size() * 8);
  return QPZK_OK;
}
// VerifierOnlyCircuitData bytes (cap_height, constants_sigmas_cap, circuit_digest); returns length.
size_t qpzk_circuit_verifier_only(const qpzk_circuit* q, uint8_t* out, size_t cap) {
  if (!q) return 0;
  ByteWriter w;
  w.u(q->common.cap_height, 8);
  w.felts(q->cs_cap.data(), q->cs_cap.size());
  w.felts(q->digest, 4);
  if (out && w.b.size() <= cap) memcpy(out, w.b.data(), w.b.size());
  return w.b.size();
}
void qpzk_circuit_free(qpzk_circuit* q) {
  if (!q) return;
  cudaSetDevice(q->ctx->device);
  dev_free(q->ctx, q->k_is_dev);
  dev_free(q->ctx, q->coset_aux_dev);
  dev_free(q->ctx, q->l0_den_inv_dev);
  dev_free(q->ctx, q->cs_values);
  qpzk_batch_free(q->cs_batch);
  delete q;
}

// Debug/parity hook: intermediate values of the last qpzk_prove on this circuit (flags & 1).
// which: 0 challenges [betas|gammas|alphas|zeta(2)|fri_alpha(2)|fri_betas(2 each)],
//        1 zs_partial_products values [nch*(1+npp)][n], 2 quotient chunk coefficients [nch*qdf][n],
//        3 FRI input polynomial [n][2]. Returns the number of u64 written (or needed if out is NULL).
size_t qpzk_prove_trace(const qpzk_circuit* q, int which, uint64_t* out) {
  if (!q) return 0;
  const std::vector<u64>* v = which == 0 ? &q->tr_challenges : which == 1 ? &q->tr_zs_pp : which == 2 ? &q->tr_quotient
                                                                                                      : &q->tr_final_poly;
  if (out) memcpy(out, v->data(), v->size() * 8);
  return v->size();
}
int qpzk_prove_stage_ms(const qpzk_circuit* q, float* out16) {
  if (!q || !out16) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  memcpy(out16, q->stage_ms, sizeof q->stage_ms);
  return QPZK_OK;
}

}  // extern "C"

namespace qpzk {

// Proof of one witness. wires_host: [num_wires][n]. Salts: NULL or host [4][N] per blinded oracle.
// ---- H11-H14: the FRI prover as a resumable object (prove_openings, fri_committed_trees,
// fri_prover_query_rounds of qp-plonky2 fri/oracle.rs, fri/prover.rs). The transcript stays with the
// caller: begin -> [commit_round -> (observe cap, squeeze beta) -> fold]* -> final_poly -> queries. ----
struct FriTree {
  u64* leaves = nullptr;  // AoS [nleaves][2*arity]
  u64* levels = nullptr;
  u32 log_n = 0, arity_bits = 0;
};
}  // namespace qpzk

struct qpzk_fri {
  qpzk_circuit* q = nullptr;
  qpzk_ctx* c = nullptr;
  const qpzk_batch* oracles[4] = {nullptr, nullptr, nullptr, nullptr};
  u64 *fpoly = nullptr, *fold_a = nullptr, *fold_b = nullptr, *vals = nullptr;  // device
  u64* coeffs_cur = nullptr;  // SoA [2][cur_n]
  u64 cur_n = 0, shift = GL_GEN;
  u32 cur_k = 0;
  bool flip = false;
  size_t round = 0;
  std::vector<qpzk::FriTree> trees;
  ~qpzk_fri() {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& t : trees) {
      dev_free(c, t.leaves);
      dev_free(c, t.levels);
    }
    dev_free(c, fpoly);
    dev_free(c, fold_a);
    dev_free(c, fold_b);
    dev_free(c, vals);
  }
};

namespace qpzk {

// prove_openings up to the polynomial that enters FRI: batch 0 = every polynomial of the four oracles
// at zeta, batch 1 = the Z polynomials at g*zeta; final = q0 * alpha^(len batch 1) + q1 (no multiply-by-X).
static int fri_begin(qpzk_circuit* q, qpzk_batch* const* oracles, const u64* zeta, const u64* alpha, qpzk_fri** out) {
  qpzk_ctx* c = q->ctx;
  const CircuitDesc& d = q->desc;
  const u32 k = d.degree_bits, r = d.rate_bits, nch = d.num_challenges;
  const u64 n = 1ull << k, N = n << r;
  std::unique_ptr<qpzk_fri> F(new qpzk_fri());
  F->q = q;
  F->c = c;
  u32 total_polys = 0;
  for (int o = 0; o < 4; o++) {
    F->oracles[o] = oracles[o];
    total_polys += oracles[o]->ncols;
  }
  u64 wn = glh::root_of_unity(k);
  u64 zeta_next[2] = {glh::mul(zeta[0], wn), glh::mul(zeta[1], wn)};
  std::vector<u64> apow((size_t)total_polys * 2);
  {
    u64 a = 1, b = 0;
    for (u32 j = 0; j < total_polys; j++) {
      apow[2 * j] = a;
      apow[2 * j + 1] = b;
      u64 na = glh::add(glh::mul(a, alpha[0]), glh::mul(7, glh::mul(b, alpha[1])));
      u64 nb = glh::add(glh::mul(a, alpha[1]), glh::mul(b, alpha[0]));
      a = na;
      b = nb;
    }
  }
  DevBuf apow_dev(c), comp0(c), comp1(c), q0(c), q1(c);
  QP(apow_dev.alloc(apow.size() * 8));
  CU(cudaMemcpyAsync(apow_dev.p, apow.data(), apow.size() * 8, cudaMemcpyHostToDevice, c->stream));
  QP(comp0.alloc(n * 16)); QP(comp1.alloc(n * 16)); QP(q0.alloc(n * 16)); QP(q1.alloc(n * 16));
  QP(dev_alloc(c, n * 16, &F->fpoly));
  QP(dev_alloc(c, n * 16 / 2 + 64, &F->fold_a));
  QP(dev_alloc(c, n * 16 / 2 + 64, &F->fold_b));
  QP(dev_alloc(c, (size_t)2 * N * 8, &F->vals));
  PolyList pl0;
  memset(&pl0, 0, sizeof pl0);
  pl0.noracles = 4;
  for (int o = 0; o < 4; o++) { pl0.base[o] = oracles[o]->coeffs; pl0.count[o] = oracles[o]->ncols; }
  PolyList pl1;
  memset(&pl1, 0, sizeof pl1);
  pl1.noracles = 1; pl1.base[0] = oracles[2]->coeffs; pl1.count[0] = nch;
  k_fri_compose<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(pl0, n, apow_dev.p, comp0.p);
  k_fri_compose<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(pl1, n, apow_dev.p, comp1.p);
  k_divide_by_linear<<<1, 1024, 0, c->stream>>>(comp0.p, q0.p, n, gl2{zeta[0], zeta[1]});
  k_divide_by_linear<<<1, 1024, 0, c->stream>>>(comp1.p, q1.p, n, gl2{zeta_next[0], zeta_next[1]});
  gl2 shift_s = gl2{apow[2 * nch], apow[2 * nch + 1]};
  k_ext_axpy<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(q0.p, q1.p, shift_s, n, F->fpoly);
  c->launches += 5;
  CU(cudaGetLastError());
  F->coeffs_cur = F->fpoly;
  F->cur_n = n;
  F->cur_k = k;
  *out = F.release();
  return QPZK_OK;
}

// One commit-phase round: LDE of the current coefficients on the coset shift*<w>, leaves of 2^arity_bits
// extension evaluations, MerkleTree::new; the cap goes back to the caller's transcript.
static int fri_commit_round(qpzk_fri* F, u64* cap_out) {
  qpzk_ctx* c = F->c;
  const CommonHost& cm = F->q->common;
  const u32 r = F->q->desc.rate_bits, h = (u32)cm.cap_height;
  if (F->round >= cm.arities.size() || F->trees.size() != F->round) return fail(QPZK_ERR_BAD_ARG, "FRI round out of order");
  const u64 ab = cm.arities[F->round];
  const u64 NV = F->cur_n << r;  // values in this round
  QP(launch_lde_shift(c, F->coeffs_cur, F->cur_n, F->vals, NV, 2, (int)F->cur_k, (int)r, F->shift));
  FriTree t;
  t.arity_bits = (u32)ab;
  t.log_n = F->cur_k + r - (u32)ab;
  if (h > t.log_n) return fail(QPZK_ERR_UNSUPPORTED, "FRI tree smaller than the cap");
  QP(dev_alloc(c, NV * 16, &t.leaves));
  F->trees.push_back(t);  // owned by F from here on
  QP(dev_alloc(c, (2ull << t.log_n) * 32, &F->trees.back().levels));
  t = F->trees.back();
  k_ext_interleave<<<(unsigned)((NV + 255) / 256), 256, 0, c->stream>>>(F->vals, NV, t.leaves);
  c->launches++;
  QP(build_tree(c, t.leaves, 2ull << ab, 1, 2u << ab, t.log_n, h, t.levels, nullptr));
  CU(cudaMemcpyAsync(cap_out, cap_ptr(t.levels, t.log_n, h), (size_t)32 << h, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

// Fold in coefficient space: chunks of 2^arity_bits coefficients combined with powers of beta.
static int fri_fold(qpzk_fri* F, u64 b0, u64 b1) {
  qpzk_ctx* c = F->c;
  const CommonHost& cm = F->q->common;
  if (F->round >= cm.arities.size() || F->trees.size() != F->round + 1) return fail(QPZK_ERR_BAD_ARG, "FRI fold out of order");
  const u64 ab = cm.arities[F->round];
  u64* dst = F->flip ? F->fold_b : F->fold_a;
  F->flip = !F->flip;
  u64 n_out = F->cur_n >> ab;
  k_fri_fold<<<(unsigned)((n_out + 127) / 128), 128, 0, c->stream>>>(F->coeffs_cur, F->cur_n, (u32)ab, gl2{b0, b1}, dst);
  c->launches++;
  CU(cudaGetLastError());
  F->coeffs_cur = dst;
  F->cur_n = n_out;
  F->cur_k -= (u32)ab;
  for (u64 e = 0; e < ab; e++) F->shift = glh::mul(F->shift, F->shift);
  F->round++;
  return QPZK_OK;
}

// The polynomial left after the last fold, as interleaved extension coefficients [len][2].
static int fri_final_poly(qpzk_fri* F, std::vector<u64>* out) {
  qpzk_ctx* c = F->c;
  if (F->round != F->q->common.arities.size()) return fail(QPZK_ERR_BAD_ARG, "FRI rounds not finished");
  const u64 m = F->cur_n;
  std::vector<u64> soa(2 * m);
  out->resize(2 * m);
  CU(cudaMemcpyAsync(soa.data(), F->coeffs_cur, m * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(soa.data() + m, F->coeffs_cur + m, m * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  for (u64 i = 0; i < m; i++) {
    (*out)[2 * i] = soa[i];
    (*out)[2 * i + 1] = soa[m + i];
  }
  return QPZK_OK;
}

// fri_prover_query_rounds for nq indices at once. init_open[o] = nq x (salted row | path), step_open[s] =
// nq x (2^arity ext evaluations | path).
static int fri_queries(qpzk_fri* F, const u64* xidx, u32 nq, std::vector<std::vector<u64>>* init_open,
                       std::vector<std::vector<u64>>* step_open) {
  qpzk_ctx* c = F->c;
  const CircuitDesc& d = F->q->desc;
  const u32 h = (u32)F->q->common.cap_height, lb = d.degree_bits + d.rate_bits;
  const u64 N = 1ull << lb;
  const u32 L0 = lb - h;
  DevBuf xdev(c);
  QP(xdev.alloc((size_t)nq * 8));
  CU(cudaMemcpyAsync(xdev.p, xidx, (size_t)nq * 8, cudaMemcpyHostToDevice, c->stream));
  init_open->assign(4, {});
  step_open->assign(F->trees.size(), {});
  std::vector<std::unique_ptr<DevBuf>> keep;
  for (int o = 0; o < 4; o++) {
    u32 width = F->oracles[o]->width();
    size_t per = width + 4ull * L0;
    keep.emplace_back(new DevBuf(c));
    QP(keep.back()->alloc(per * nq * 8));
    k_gather_openings<<<nq, 128, 0, c->stream>>>(F->oracles[o]->lde, 1, N, width, F->oracles[o]->levels, lb, h, xdev.p, 0,
                                                 keep.back()->p);
    c->launches++;
    (*init_open)[o].resize(per * nq);
    CU(cudaMemcpyAsync((*init_open)[o].data(), keep.back()->p, per * nq * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  u32 sh = 0;
  for (size_t s = 0; s < F->trees.size(); s++) {
    const FriTree& t = F->trees[s];
    sh += t.arity_bits;
    u32 width = 2u << t.arity_bits, L = t.log_n - h;
    size_t per = width + 4ull * L;
    keep.emplace_back(new DevBuf(c));
    QP(keep.back()->alloc(per * nq * 8));
    k_gather_openings<<<nq, 128, 0, c->stream>>>(t.leaves, width, 1, width, t.levels, t.log_n, h, xdev.p, sh,
                                                 keep.back()->p);
    c->launches++;
    (*step_open)[s].resize(per * nq);
    CU(cudaMemcpyAsync((*step_open)[s].data(), keep.back()->p, per * nq * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaGetLastError());
  CU(ctx_wait(c));
  return QPZK_OK;
}

// ---- H8: Z and partial products on the subgroup: zs_vals [nch*(1+npp)][n] = Z_0..Z_{nch-1}, then the
// partial products of each challenge (all_wires_permutation_partial_products + the running product) ----
static int compute_zs_partial_products(qpzk_circuit* q, const u64* wires_dev, const Challenges& chal, DevBuf* zs_vals) {
  qpzk_ctx* c = q->ctx;
  const CircuitDesc& d = q->desc;
  const u32 k = d.degree_bits, nch = d.num_challenges, npp = d.num_partial_products;
  const u64 n = 1ull << k;
  RootTab tab_n;
  QP(get_root_tab(c, (int)k, false, &tab_n));
  const u32 nchunks = npp + 1, nzs = nch * (1 + npp);
  DevBuf chunk_q(c), row_prod(c);
  QP(chunk_q.alloc((size_t)nch * nchunks * n * 8));
  QP(row_prod.alloc((size_t)nch * n * 8));
  QP(zs_vals->alloc((size_t)nzs * n * 8));
  k_zs_chunk_quotients<<<dim3((unsigned)((n + 127) / 128), nch), 128, 0, c->stream>>>(
      wires_dev, q->cs_values, q->k_is_dev, d, chal, tab_n, chunk_q.p, row_prod.p);
  k_prefix_product<<<nch, 1024, 0, c->stream>>>(row_prod.p, zs_vals->p, n);
  k_partial_products<<<dim3((unsigned)((n + 127) / 128), nch), 128, 0, c->stream>>>(chunk_q.p, zs_vals->p, nch, npp, n,
                                                                                   zs_vals->p + (size_t)nch * n);
  c->launches += 3;
  CU(cudaGetLastError());
  return QPZK_OK;
}

// ---- H9: compute_quotient_polys: vanishing(x)/Z_H(x) on the quotient coset from the three committed
// oracles, coset IFFT, coefficients [nch][qdf*n] (= nch*qdf chunks of n) ----
static int compute_quotient_chunks(qpzk_circuit* q, const qpzk_batch* wires_b, const qpzk_batch* zs_b, const u64* pi_hash,
                                   const Challenges& chal, DevBuf* qcoeffs) {
  qpzk_ctx* c = q->ctx;
  const CommonHost& cm = q->common;
  const CircuitDesc& d = q->desc;
  const u32 k = d.degree_bits, r = d.rate_bits, nch = d.num_challenges, qdb = d.quotient_degree_bits;
  const u64 n = 1ull << k, N = n << r;
  const u32 qlb = k + qdb;
  const u64 qlde = 1ull << qlb;
  std::vector<u64> zh(1u << qdb), zh_inv(1u << qdb);
  {
    u64 gn = glh::pow(GL_GEN, n), wq = glh::root_of_unity(qdb);
    for (u32 i = 0; i < (1u << qdb); i++) {
      zh[i] = glh::sub(glh::mul(gn, glh::pow(wq, i)), 1);
      zh_inv[i] = glh::inv(zh[i]);
    }
  }
  DevBuf small(c), qvals(c);
  std::vector<u64> apw(2 * QPZK_APW_STRIDE, 0);   // alpha_c^t for the reduction of the constraint terms
  for (u32 ci = 0; ci < nch; ci++) {
    u64 pwr = 1;
    for (u32 t = 0; t < QPZK_APW_STRIDE; t++) {
      apw[ci * QPZK_APW_STRIDE + t] = pwr;
      pwr = glh::mul(pwr, chal.alpha[ci]);
    }
  }
  QP(small.alloc((4 + 2 * (1u << qdb) + 2 * QPZK_APW_STRIDE) * 8));
  CU(cudaMemcpyAsync(small.p + 4 + 2 * (1u << qdb), apw.data(), apw.size() * 8, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(small.p, pi_hash, 32, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(small.p + 4, zh.data(), zh.size() * 8, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(small.p + 4 + zh.size(), zh_inv.data(), zh.size() * 8, cudaMemcpyHostToDevice, c->stream));
  QP(qvals.alloc((size_t)nch * qlde * 8));
  QP(qcoeffs->alloc((size_t)nch * qlde * 8));
  RootTab tab_q;
  QP(get_root_tab(c, (int)qlb, false, &tab_q));
  if (cm.recursion)
    k_quotient<true><<<(unsigned)((qlde + 127) / 128), 128, 0, c->stream>>>(
        q->cs_batch->lde, wires_b->lde, zs_b->lde, N, N, N, r - qdb, q->k_is_dev, d, chal, small.p, small.p + 4,
        small.p + 4 + zh.size(), small.p + 4 + 2 * zh.size(), q->l0_den_inv_dev, tab_q, qvals.p);
  else
    k_quotient<false><<<(unsigned)((qlde + 127) / 128), 128, 0, c->stream>>>(
        q->cs_batch->lde, wires_b->lde, zs_b->lde, N, N, N, r - qdb, q->k_is_dev, d, chal, small.p, small.p + 4,
        small.p + 4 + zh.size(), small.p + 4 + 2 * zh.size(), q->l0_den_inv_dev, tab_q, qvals.p);
  c->launches++;
  CU(cudaGetLastError());
  // coset IFFT: values on g*<w> -> coefficients; then split into qdf chunks of n (contiguous already)
  QP(launch_ifft(c, qvals.p, qlde, qcoeffs->p, qlde, nch, (int)qlb));
  RootTab tab_ginv;
  QP(get_pow_tab(c, glh::inv(GL_GEN), (int)qlb, &tab_ginv));
  k_scale_by_powers<<<dim3((unsigned)((qlde + 255) / 256), nch), 256, 0, c->stream>>>(qcoeffs->p, qlde, tab_ginv);
  c->launches++;
  CU(cudaGetLastError());
  // the small staging vectors above are pageable host memory: the copies must have left before they die
  CU(ctx_wait(c));
  return QPZK_OK;
}

static int prove_impl(qpzk_circuit* q, const u64* wires_host, const u64* pis, u32 npi, const u64* salt_w,
                      const u64* salt_z, const u64* salt_q, u32 flags, std::vector<uint8_t>* proof) {
  qpzk_ctx* c = q->ctx;
  const CommonHost& cm = q->common;
  const CircuitDesc& d = q->desc;
  const u32 k = d.degree_bits, r = d.rate_bits, h = (u32)cm.cap_height, nch = d.num_challenges;
  const u32 npp = d.num_partial_products, nw = d.num_wires, qdf = d.qdf, qdb = d.quotient_degree_bits;
  const u64 n = 1ull << k, N = n << r;
  const u32 lb = k + r;
  const u32 salt_cols = cm.hiding ? QPZK_SALT_SIZE : 0;
  if (cm.hiding && !(salt_w && salt_z && salt_q)) return fail(QPZK_ERR_BAD_ARG, "hiding circuit needs salts");
  if (npi != cm.num_public_inputs) return fail(QPZK_ERR_BAD_ARG, "public input count mismatch");
  const bool want_trace = flags & 1;
  const bool on_device = flags & 2;  // wires / salts are device pointers (HBM-resident witness)
  cudaEvent_t evs[2];
  CU(cudaEventCreate(&evs[0]));
  CU(cudaEventCreate(&evs[1]));
  int stage = 0;
  memset(q->stage_ms, 0, sizeof q->stage_ms);
  auto tic = [&]() { cudaEventRecord(evs[0], c->stream); };
  auto toc = [&]() {
    cudaEventRecord(evs[1], c->stream);
    cudaEventSynchronize(evs[1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, evs[0], evs[1]);
    if (stage < 16) q->stage_ms[stage++] = ms;
  };

  u64 pi_hash[4];
  host_hash_no_pad(pis, npi, pi_hash);
  HostChallenger ch;
  ch.observe_n(q->digest, 4);
  ch.observe_n(pi_hash, 4);

  // ---- (2) commit wires ----
  tic();
  DevBuf wires_up(c);
  const u64* wires_dev = wires_host;
  if (!on_device) {
    QP(wires_up.alloc((size_t)nw * n * 8));
    CU(cudaMemcpyAsync(wires_up.p, wires_host, (size_t)nw * n * 8, cudaMemcpyHostToDevice, c->stream));
    wires_dev = wires_up.p;
  }
  qpzk_batch* wires_b = nullptr;
  QP(commit_impl(c, wires_dev, false, false, nw, k, r, h, cm.hiding ? salt_w : nullptr, !on_device, salt_cols, &wires_b));
  std::unique_ptr<qpzk_batch, void (*)(qpzk_batch*)> wires_guard(wires_b, qpzk_batch_free);
  std::vector<u64> cap(4ull << h);
  QP(qpzk_batch_cap(wires_b, cap.data()));
  std::vector<u64> wires_cap = cap;
  ch.observe_n(cap.data(), cap.size());
  toc();  // stage 0: wires commit
  Challenges chal;
  memset(&chal, 0, sizeof chal);
  for (u32 i = 0; i < nch; i++) chal.beta[i] = ch.get();
  for (u32 i = 0; i < nch; i++) chal.gamma[i] = ch.get();

  // ---- (4,5) Z + partial products, commit ----
  tic();
  const u32 nzs = nch * (1 + npp);
  DevBuf zs_vals(c);
  QP(compute_zs_partial_products(q, wires_dev, chal, &zs_vals));
  qpzk_batch* zs_b = nullptr;
  QP(commit_impl(c, zs_vals.p, false, false, nzs, k, r, h, cm.hiding ? salt_z : nullptr, !on_device, salt_cols, &zs_b));
  std::unique_ptr<qpzk_batch, void (*)(qpzk_batch*)> zs_guard(zs_b, qpzk_batch_free);
  QP(qpzk_batch_cap(zs_b, cap.data()));
  std::vector<u64> zs_cap = cap;
  ch.observe_n(cap.data(), cap.size());
  toc();  // stage 1: Z/pp + commit
  for (u32 i = 0; i < nch; i++) chal.alpha[i] = ch.get();
  if (want_trace) {
    q->tr_zs_pp.resize((size_t)nzs * n);
    CU(cudaMemcpyAsync(q->tr_zs_pp.data(), zs_vals.p, (size_t)nzs * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
  }

  // ---- (6,7) quotient ----
  tic();
  const u32 qlb = k + qdb;
  const u64 qlde = 1ull << qlb;
  DevBuf qcoeffs(c);
  QP(compute_quotient_chunks(q, wires_b, zs_b, pi_hash, chal, &qcoeffs));
  qpzk_batch* q_b = nullptr;
  QP(commit_impl(c, qcoeffs.p, false, true, nch * qdf, k, r, h, cm.hiding ? salt_q : nullptr, !on_device, salt_cols, &q_b));
  std::unique_ptr<qpzk_batch, void (*)(qpzk_batch*)> q_guard(q_b, qpzk_batch_free);
  QP(qpzk_batch_cap(q_b, cap.data()));
  std::vector<u64> q_cap = cap;
  ch.observe_n(cap.data(), cap.size());
  toc();  // stage 2: quotient + commit
  if (want_trace) {
    q->tr_quotient.resize((size_t)nch * qlde);
    CU(cudaMemcpyAsync(q->tr_quotient.data(), qcoeffs.p, (size_t)nch * qlde * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
  }
  u64 zeta[2] = {ch.get(), 0};
  zeta[1] = ch.get();

  // ---- (8) openings ----
  tic();
  qpzk_batch* oracles[4] = {q->cs_batch, wires_b, zs_b, q_b};
  u32 total_polys = 0;
  for (auto* b : oracles) total_polys += b->ncols;
  u64 wn = glh::root_of_unity(k);
  u64 zeta_next[2] = {glh::mul(zeta[0], wn), glh::mul(zeta[1], wn)};
  DevBuf zpow(c), zpow_next(c), open_dev(c);
  QP(zpow.alloc(n * 16));
  QP(zpow_next.alloc(n * 16));
  QP(open_dev.alloc((size_t)(total_polys + nch) * 16));
  k_ext_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(gl2{zeta[0], zeta[1]}, n, zpow.p);
  k_ext_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(gl2{zeta_next[0], zeta_next[1]}, n, zpow_next.p);
  c->launches += 2;
  {
    u32 off = 0;
    for (auto* b : oracles) {
      k_eval_at_ext<<<b->ncols, 256, 0, c->stream>>>(b->coeffs, n, zpow.p, open_dev.p + 2ull * off);
      off += b->ncols;
      c->launches++;
    }
    k_eval_at_ext<<<nch, 256, 0, c->stream>>>(zs_b->coeffs, n, zpow_next.p, open_dev.p + 2ull * off);
    c->launches++;
  }
  CU(cudaGetLastError());
  std::vector<u64> opens((size_t)(total_polys + nch) * 2);
  CU(cudaMemcpyAsync(opens.data(), open_dev.p, opens.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  // observe order: constants, sigmas, wires, zs, partial_products, quotient (= oracle order) then zs_next
  ch.observe_n(opens.data(), opens.size());
  toc();  // stage 3: openings
  u64 alpha[2] = {ch.get(), 0};
  alpha[1] = ch.get();

  // ---- (9) FRI: batch combine ----
  tic();
  qpzk_fri* F = nullptr;
  QP(fri_begin(q, oracles, zeta, alpha, &F));
  std::unique_ptr<qpzk_fri> fri_guard(F);
  if (want_trace) {
    std::vector<u64> soa(2 * n);
    CU(cudaMemcpyAsync(soa.data(), F->fpoly, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    q->tr_final_poly.resize(2 * n);
    for (u64 m = 0; m < n; m++) {
      q->tr_final_poly[2 * m] = soa[m];
      q->tr_final_poly[2 * m + 1] = soa[n + m];
    }
  }
  toc();  // stage 4: FRI combine

  // ---- (9) FRI: commit phase ----
  tic();
  std::vector<std::vector<u64>> fri_caps;
  std::vector<u64> fri_betas;
  for (size_t round = 0; round < cm.arities.size(); round++) {
    std::vector<u64> fc(4ull << h);
    QP(fri_commit_round(F, fc.data()));
    ch.observe_n(fc.data(), fc.size());
    fri_caps.push_back(fc);
    u64 b0 = ch.get(), b1 = ch.get();
    fri_betas.push_back(b0);
    fri_betas.push_back(b1);
    QP(fri_fold(F, b0, b1));
  }
  std::vector<u64> final_poly;
  QP(fri_final_poly(F, &final_poly));
  ch.observe_n(final_poly.data(), final_poly.size());
  toc();  // stage 5: FRI commit phase

  // ---- (9) proof of work ----
  tic();
  u64 pow_witness = 0;
  {
    PowState ps;
    memcpy(ps.s, ch.state, sizeof ps.s);
    u32 pos = (u32)ch.in.size();
    for (u32 i = 0; i < pos; i++) ps.s[i] = ch.in[i];
    QP(grind_pow(c, ps, pos, cm.pow_bits, &pow_witness));
  }
  ch.observe(pow_witness);
  u64 pow_resp = ch.get();
  if ((pow_resp >> (64 - cm.pow_bits)) != 0 && cm.pow_bits) return fail(QPZK_ERR_CUDA, "pow response mismatch");
  toc();  // stage 6: PoW

  // ---- (9) query rounds ----
  tic();
  const u32 nq = (u32)cm.num_queries;
  std::vector<u64> xidx(nq);
  for (u32 i = 0; i < nq; i++) xidx[i] = ch.get() & (N - 1);
  std::vector<std::vector<u64>> init_open, step_open;
  const u32 L0 = lb - h;
  QP(fri_queries(F, xidx.data(), nq, &init_open, &step_open));
  toc();  // stage 7: queries

  // ---- (10) ProofWithPublicInputs::to_bytes ----
  ByteWriter w;
  w.felts(wires_cap.data(), wires_cap.size());
  w.felts(zs_cap.data(), zs_cap.size());
  w.felts(q_cap.data(), q_cap.size());
  {
    // openings in serialised order: constants, sigmas, wires, zs, zs_next, partial_products, quotient
    const u64* o = opens.data();
    size_t n_cs = (size_t)(cm.num_constants + cm.num_routed), off_w = n_cs, off_z = off_w + nw;
    size_t off_pp = off_z + nch, off_q = off_z + nzs, off_next = total_polys;
    w.felts(o, 2 * n_cs);
    w.felts(o + 2 * off_w, 2ull * nw);
    w.felts(o + 2 * off_z, 2ull * nch);
    w.felts(o + 2 * off_next, 2ull * nch);
    w.felts(o + 2 * off_pp, 2ull * nch * npp);
    w.felts(o + 2 * off_q, 2ull * nch * qdf);
  }
  for (auto& fc : fri_caps) w.felts(fc.data(), fc.size());
  for (u32 qi = 0; qi < nq; qi++) {
    for (int o = 0; o < 4; o++) {
      u32 width = oracles[o]->width();
      size_t per = width + 4ull * L0;
      const u64* p = init_open[o].data() + per * qi;
      w.felts(p, width);
      w.u(L0, 1);
      w.felts(p + width, 4ull * L0);
    }
    for (size_t s = 0; s < step_open.size(); s++) {
      u32 width = 2u << cm.arities[s];
      u32 L = (u32)((step_open[s].size() / nq - width) / 4);
      const u64* p = step_open[s].data() + (size_t)(width + 4ull * L) * qi;
      w.felts(p, width);
      w.u(L, 1);
      w.felts(p + width, 4ull * L);
    }
  }
  w.felts(final_poly.data(), final_poly.size());
  w.u(pow_witness, 8);
  w.u(npi, 8);
  for (u32 i = 0; i < npi; i++) w.u(pis[i] >= GL_P ? pis[i] - GL_P : pis[i], 8);
  *proof = std::move(w.b);
  if (want_trace) {
    q->tr_challenges.clear();
    for (u32 i = 0; i < nch; i++) q->tr_challenges.push_back(chal.beta[i]);
    for (u32 i = 0; i < nch; i++) q->tr_challenges.push_back(chal.gamma[i]);
    for (u32 i = 0; i < nch; i++) q->tr_challenges.push_back(chal.alpha[i]);
    q->tr_challenges.push_back(zeta[0]); q->tr_challenges.push_back(zeta[1]);
    q->tr_challenges.push_back(alpha[0]); q->tr_challenges.push_back(alpha[1]);
    for (u64 b : fri_betas) q->tr_challenges.push_back(b);
  }
  cudaEventDestroy(evs[0]);
  cudaEventDestroy(evs[1]);
  return QPZK_OK;
}

}  // namespace qpzk

extern "C" {

static int challenges_from(const qpzk_circuit* q, const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas,
                           Challenges* chal) {
  memset(chal, 0, sizeof *chal);
  for (u32 i = 0; i < q->desc.num_challenges; i++) {
    if (betas) chal->beta[i] = betas[i];
    if (gammas) chal->gamma[i] = gammas[i];
    if (alphas) chal->alpha[i] = alphas[i];
  }
  return QPZK_OK;
}

int qpzk_zs_partial_products(qpzk_circuit* q, const uint64_t* wires, const uint64_t* betas, const uint64_t* gammas,
                             uint64_t* out) {
  if (!q || !wires || !betas || !gammas || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_ctx* c = q->ctx;
  CU(cudaSetDevice(c->device));
  const CircuitDesc& d = q->desc;
  const u64 n = 1ull << d.degree_bits;
  Challenges chal;
  challenges_from(q, betas, gammas, nullptr, &chal);
  DevBuf wd(c), zs(c);
  QP(wd.alloc((size_t)d.num_wires * n * 8));
  CU(cudaMemcpyAsync(wd.p, wires, (size_t)d.num_wires * n * 8, cudaMemcpyHostToDevice, c->stream));
  QP(compute_zs_partial_products(q, wd.p, chal, &zs));
  const size_t cnt = (size_t)d.num_challenges * (1 + d.num_partial_products) * n;
  CU(cudaMemcpyAsync(out, zs.p, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_quotient(qpzk_circuit* q, const qpzk_batch* wires_batch, const qpzk_batch* zs_batch, const uint64_t* pi_hash,
                  const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, uint64_t* out_chunks) {
  if (!q || !wires_batch || !zs_batch || !pi_hash || !betas || !gammas || !alphas || !out_chunks)
    return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_ctx* c = q->ctx;
  const CircuitDesc& d = q->desc;
  if (wires_batch->ctx != c || zs_batch->ctx != c) return fail(QPZK_ERR_BAD_ARG, "batches belong to another context");
  if (wires_batch->degree_bits != d.degree_bits || zs_batch->degree_bits != d.degree_bits ||
      wires_batch->rate_bits != d.rate_bits || zs_batch->rate_bits != d.rate_bits || wires_batch->ncols != d.num_wires ||
      zs_batch->ncols != d.num_challenges * (1 + d.num_partial_products))
    return fail(QPZK_ERR_BAD_ARG, "batch shape does not match the circuit");
  CU(cudaSetDevice(c->device));
  Challenges chal;
  challenges_from(q, betas, gammas, alphas, &chal);
  DevBuf qc(c);
  QP(compute_quotient_chunks(q, wires_batch, zs_batch, pi_hash, chal, &qc));
  const size_t cnt = ((size_t)d.num_challenges << (d.degree_bits + d.quotient_degree_bits));
  CU(cudaMemcpyAsync(out_chunks, qc.p, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

int qpzk_fri_begin(qpzk_circuit* q, const qpzk_batch* wires_batch, const qpzk_batch* zs_batch,
                   const qpzk_batch* quotient_batch, const uint64_t* zeta, const uint64_t* alpha, qpzk_fri** out) {
  if (!q || !wires_batch || !zs_batch || !quotient_batch || !zeta || !alpha || !out)
    return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_ctx* c = q->ctx;
  const CircuitDesc& d = q->desc;
  qpzk_batch* oracles[4] = {q->cs_batch, const_cast<qpzk_batch*>(wires_batch), const_cast<qpzk_batch*>(zs_batch),
                            const_cast<qpzk_batch*>(quotient_batch)};
  const u32 want[4] = {d.num_constants + d.num_routed, d.num_wires, d.num_challenges * (1 + d.num_partial_products),
                       d.num_challenges * d.qdf};
  for (int o = 0; o < 4; o++)
    if (oracles[o]->ctx != c || oracles[o]->degree_bits != d.degree_bits || oracles[o]->rate_bits != d.rate_bits ||
        oracles[o]->ncols != want[o] || oracles[o]->cap_height != q->common.cap_height)
      return fail(QPZK_ERR_BAD_ARG, "oracle shape does not match the circuit");
  CU(cudaSetDevice(c->device));
  return fri_begin(q, oracles, zeta, alpha, out);
}
uint32_t qpzk_fri_num_rounds(const qpzk_fri* f) { return f ? (uint32_t)f->q->common.arities.size() : 0; }
int qpzk_fri_commit_round(qpzk_fri* f, uint64_t* cap_out) {
  if (!f || !cap_out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(f->c->device));
  return fri_commit_round(f, cap_out);
}
int qpzk_fri_fold(qpzk_fri* f, const uint64_t* beta) {
  if (!f || !beta) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(f->c->device));
  return fri_fold(f, beta[0], beta[1]);
}
int qpzk_fri_final_poly(qpzk_fri* f, uint64_t* out, size_t cap_words, size_t* len_words) {
  if (!f || !len_words) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(f->c->device));
  std::vector<u64> v;
  QP(fri_final_poly(f, &v));
  *len_words = v.size();
  if (out) {
    if (v.size() > cap_words) return fail(QPZK_ERR_BAD_ARG, "buffer too small");
    memcpy(out, v.data(), v.size() * 8);
  }
  return QPZK_OK;
}
int qpzk_fri_query(qpzk_fri* f, uint64_t x_index, uint64_t* out, size_t cap_words, size_t* len_words) {
  if (!f || !len_words) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  const CircuitDesc& d = f->q->desc;
  if (x_index >> (d.degree_bits + d.rate_bits)) return fail(QPZK_ERR_BAD_ARG, "x_index out of range");
  if (f->round != f->q->common.arities.size()) return fail(QPZK_ERR_BAD_ARG, "FRI rounds not finished");
  CU(cudaSetDevice(f->c->device));
  std::vector<std::vector<u64>> init_open, step_open;
  QP(fri_queries(f, &x_index, 1, &init_open, &step_open));
  size_t total = 0;
  for (auto& v : init_open) total += v.size();
  for (auto& v : step_open) total += v.size();
  *len_words = total;
  if (out) {
    if (total > cap_words) return fail(QPZK_ERR_BAD_ARG, "buffer too small");
    size_t off = 0;
    for (auto& v : init_open) { memcpy(out + off, v.data(), v.size() * 8); off += v.size(); }
    for (auto& v : step_open) { memcpy(out + off, v.data(), v.size() * 8); off += v.size(); }
  }
  return QPZK_OK;
}
void qpzk_fri_free(qpzk_fri* f) { delete f; }

int qpzk_prove(qpzk_circuit* q, const uint64_t* wires, const uint64_t* public_inputs, uint32_t num_public_inputs,
               const uint64_t* salts_wires, const uint64_t* salts_zs, const uint64_t* salts_quotient, uint32_t flags,
               uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  if (!q || !wires || (!public_inputs && num_public_inputs) || !proof_len)
    return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  CU(cudaSetDevice(q->ctx->device));
  std::vector<uint8_t> bytes;
  int rc = prove_impl(q, wires, public_inputs, num_public_inputs, salts_wires, salts_zs, salts_quotient, flags, &bytes);
  if (rc != QPZK_OK) {
    ctx_wait(q->ctx);
    return rc;
  }
  *proof_len = bytes.size();
  if (proof_out) {
    if (bytes.size() > proof_cap) return fail(QPZK_ERR_BAD_ARG, "proof buffer too small");
    memcpy(proof_out, bytes.data(), bytes.size());
  }
  return QPZK_OK;
}

}  // extern "C"
