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

// Everything the kernels index with comes from these bytes, so every range is checked here, once, before a
// single launch: the CircuitDesc arrays are 16 entries, the alpha-power table QPZK_APW_STRIDE, a gate reads
// the wires and constants its layout says, shifts by degree_bits must stay below 64, and so on.
static bool validate_common(const CommonHost& c, std::string* err) {
#define QPZK_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) {              \
      *err = msg;               \
      return false;             \
    }                           \
  } while (0)
  QPZK_REQUIRE(c.degree_bits <= 30 && c.rate_bits <= 30 && c.degree_bits + c.rate_bits <= 30, "degree_bits + rate_bits above 30");
  QPZK_REQUIRE(c.cap_height <= c.degree_bits + c.rate_bits && c.cap_height <= 16, "cap_height out of range");
  QPZK_REQUIRE(c.num_challenges >= 1 && c.num_challenges <= 2, "num_challenges must be 1 or 2");
  QPZK_REQUIRE(c.num_wires >= 1 && c.num_wires <= 4096 && c.num_routed <= c.num_wires, "bad wire counts");
  QPZK_REQUIRE(c.num_constants <= 4096 && c.num_public_inputs <= (1u << 20), "bad constant / public input counts");
  QPZK_REQUIRE(c.num_queries >= 1 && c.num_queries <= QPZK_MAX_QUERIES, "num_queries out of range");
  QPZK_REQUIRE(c.pow_bits <= 40, "proof-of-work bits above 40");
  QPZK_REQUIRE(c.arities.size() <= QPZK_MAX_FRI_ROUNDS, "too many FRI rounds");
  {
    u64 kcur = c.degree_bits;
    for (u64 ab : c.arities) {
      QPZK_REQUIRE(ab >= 1 && ab <= 8 && ab <= kcur, "bad FRI arity");
      QPZK_REQUIRE(kcur + c.rate_bits - ab >= c.cap_height, "FRI tree smaller than the cap");
      kcur -= ab;
    }
  }
  u32 qdb = 0;
  while ((1ull << qdb) < c.qdf && qdb < 32) qdb++;
  QPZK_REQUIRE(c.qdf >= 1 && (1ull << qdb) == c.qdf && qdb <= c.rate_bits, "quotient_degree_factor must be a power of two within the rate");
  QPZK_REQUIRE(c.num_partial_products + 1 <= 12, "more than 11 partial products");
  QPZK_REQUIRE((c.num_partial_products + 1) * c.qdf >= c.num_routed, "partial-product chunks do not cover the routed wires");
  QPZK_REQUIRE(c.k_is.size() == c.num_routed, "k_is / num_routed_wires mismatch");
  const size_t ng = c.gates.size(), ns = c.groups.size();
  QPZK_REQUIRE(ng >= 1 && ng <= QPZK_MAX_GATES && ns >= 1 && ns <= QPZK_MAX_GATES, "bad gate / selector group counts");
  QPZK_REQUIRE(c.selector_indices.size() == ng && c.num_constants >= ns, "selector data mismatch");
  for (size_t s = 0; s < ns; s++)
    QPZK_REQUIRE(c.groups[s].first <= c.groups[s].second && c.groups[s].second <= ng, "selector group out of range");
  u64 max_constraints = 0;
  for (size_t g = 0; g < ng; g++) {
    const u64 si = c.selector_indices[g];
    QPZK_REQUIRE(si < ns && c.groups[si].first <= g && g < c.groups[si].second, "gate outside its selector group");
    const u64 p1 = c.gates[g].second, p2 = c.gate_p2[g], p3 = c.gate_p3[g];
    u64 wires = 0, consts = 0, cons = 0;  // what the evaluator reads / emits
    switch (c.gates[g].first) {
      case G_NOOP: break;
      case G_CONSTANT: wires = p1; consts = p1; cons = p1; break;
      case G_PUBLIC_INPUT: wires = 4; cons = 4; break;
      case G_BASE_SUM_2: wires = 1 + p1; cons = 1 + p1; break;
      case G_ARITHMETIC: wires = 4 * p1; consts = 2; cons = p1; break;
      case G_POSEIDON: wires = 135; cons = 123; break;
      case G_ARITHMETIC_EXT: wires = 8 * p1; consts = 2; cons = 2 * p1; break;
      case G_MUL_EXT: wires = 6 * p1; consts = 1; cons = 2 * p1; break;
      case G_POSEIDON_MDS: wires = 48; cons = 24; break;
      case G_RANDOM_ACCESS: {
        const u64 vec = 1ull << p1;
        wires = (2 + vec) * p2 + p3 + p2 * p1;
        consts = p3;
        cons = p2 * (p1 + 2) + p3;
        break;
      }
      case G_REDUCING: QPZK_REQUIRE(p1 >= 1, "ReducingGate without coefficients"); wires = 6 + p1 + 2 * (p1 - 1); cons = 2 * p1; break;
      case G_REDUCING_EXT: QPZK_REQUIRE(p1 >= 1, "ReducingExtensionGate without coefficients"); wires = 6 + 2 * p1 + 2 * (p1 - 1); cons = 2 * p1; break;
      case G_EXPONENTIATION: wires = 2 + 2 * p1; cons = p1 + 1; break;
      case G_COSET_INTERPOLATION: {
        const u64 np = 1ull << p1, nint = (np - 2) / (p2 - 1);
        wires = 1 + 2 * np + 2 + 2 + 4 * nint + 2;
        cons = 2 + 4 * nint + 2;
        break;
      }
      default: *err = "unsupported gate"; return false;
    }
    QPZK_REQUIRE(p1 <= 4096 && wires <= c.num_wires, "gate needs more wires than the circuit has");
    QPZK_REQUIRE(ns + consts <= c.num_constants, "gate needs more constants than the circuit has");
    if (cons > max_constraints) max_constraints = cons;
  }
  QPZK_REQUIRE(max_constraints <= c.num_gate_constraints, "num_gate_constraints smaller than a gate's constraint count");
  QPZK_REQUIRE(c.num_challenges * (1 + c.num_partial_products + 1) + c.num_gate_constraints <= QPZK_APW_STRIDE,
               "too many constraint terms for the alpha-power table");
#undef QPZK_REQUIRE
  return true;
}

struct ByteWriter {
  std::vector<uint8_t> b;
  void u(u64 v, size_t bytes) { for (size_t i = 0; i < bytes; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void felts(const u64* x, size_t n) {  // little-endian host (x86-64 / aarch64): the words are the bytes
    const uint8_t* p = reinterpret_cast<const uint8_t*>(x);
    b.insert(b.end(), p, p + 8 * n);
  }
};

}  // namespace qpzk

// Where the pieces of a proof land in the per-circuit arena (u64 word offsets). The arena is one device
// buffer, mirrored by one pinned host buffer and copied back ONCE per proof; word 0 holds the transcript
// (challenges, PoW witness, query indices).
struct ArenaLayout {
  size_t caps = 0;        // [3][capw]   wires, zs|partial products, quotient
  size_t opens = 0;       // [(total_polys + nch)][2]  oracle order, then Z(g zeta)
  size_t fri_caps = 0;    // [rounds][capw]
  size_t final_poly = 0;  // [m][2]
  size_t init_open[4] = {0, 0, 0, 0};   // [nq][width_o + 4 L0]
  size_t step_open[QPZK_MAX_FRI_ROUNDS] = {0};  // [nq][2 * 2^arity + 4 L_s]
  size_t total = 0;
  u32 capw = 0, total_polys = 0, final_len = 0, L0 = 0;
  u32 width[4] = {0, 0, 0, 0};
  u32 step_width[QPZK_MAX_FRI_ROUNDS] = {0}, step_L[QPZK_MAX_FRI_ROUNDS] = {0};
};

struct qpzk_circuit {
  qpzk_ctx* ctx = nullptr;
  CommonHost common;
  CircuitDesc desc;
  u64 digest[4] = {0, 0, 0, 0};
  u64* k_is_dev = nullptr;
  u64* coset_aux_dev = nullptr;
  u64* l0_den_inv_dev = nullptr;   // [2^(degree_bits + qdb)], see k_build_l0_den_inv
  u64* zh_dev = nullptr;           // Z_H on the quotient coset: [2^qdb] values, then their inverses
  u64* cs_values = nullptr;        // [num_constants + num_routed][n] values on the subgroup (for Z)
  qpzk_batch* cs_batch = nullptr;
  std::vector<u64> cs_cap;
  // one proof at a time per circuit handle: the arena and the transcript in it belong to the proof in flight
  ArenaLayout lay;
  u64* arena_dev = nullptr;
  u64* arena_host = nullptr;       // pinned
  cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done = nullptr;
  bool in_flight = false;
  std::vector<u64> pis;            // canonical public inputs of the proof in flight
  // debug trace of the last proof
  bool want_trace = false;
  std::vector<u64> tr_challenges, tr_zs_pp, tr_quotient, tr_final_poly;
  float stage_ms[16] = {0};
  std::mutex mu;                   // serialises begin / end on this handle
  TranscriptDev* transcript() const { return reinterpret_cast<TranscriptDev*>(arena_dev); }
  ~qpzk_circuit() {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (in_flight) ctx_wait(ctx);
    dev_free(ctx, k_is_dev);
    dev_free(ctx, coset_aux_dev);
    dev_free(ctx, l0_den_inv_dev);
    dev_free(ctx, zh_dev);
    dev_free(ctx, cs_values);
    dev_free(ctx, arena_dev);
    if (arena_host) cudaFreeHost(arena_host);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    if (done) cudaEventDestroy(done);
    delete cs_batch;
  }
};

namespace qpzk {

static const size_t kTranscriptWords = (sizeof(TranscriptDev) + 7) / 8;

static void layout_arena(const CommonHost& cm, const CircuitDesc& d, ArenaLayout* L) {
  const u32 h = (u32)cm.cap_height, nq = (u32)cm.num_queries, salt = cm.hiding ? QPZK_SALT_SIZE : 0;
  L->capw = 4u << h;
  L->width[0] = d.num_constants + d.num_routed;
  L->width[1] = d.num_wires + salt;
  L->width[2] = d.num_challenges * (1 + d.num_partial_products) + salt;
  L->width[3] = d.num_challenges * d.qdf + salt;
  L->total_polys = L->width[0] + (L->width[1] - salt) + (L->width[2] - salt) + (L->width[3] - salt);
  L->L0 = d.degree_bits + d.rate_bits - h;
  size_t off = (kTranscriptWords + 3) & ~(size_t)3;
  L->caps = off; off += 3 * (size_t)L->capw;
  L->opens = off; off += 2 * (size_t)(L->total_polys + d.num_challenges);
  L->fri_caps = off; off += cm.arities.size() * (size_t)L->capw;
  u32 kcur = d.degree_bits;
  for (size_t s = 0; s < cm.arities.size(); s++) {
    const u32 ab = (u32)cm.arities[s];
    L->step_width[s] = 2u << ab;
    L->step_L[s] = kcur + d.rate_bits - ab - h;
    kcur -= ab;
  }
  L->final_len = 1u << kcur;
  L->final_poly = off; off += 2 * (size_t)L->final_len;
  for (int o = 0; o < 4; o++) {
    L->init_open[o] = off;
    off += (size_t)nq * (L->width[o] + 4 * (size_t)L->L0);
  }
  for (size_t s = 0; s < cm.arities.size(); s++) {
    L->step_open[s] = off;
    off += (size_t)nq * (L->step_width[s] + 4 * (size_t)L->step_L[s]);
  }
  L->total = off;
}

// fri_proof_of_work: the smallest w such that the permutation of the sponge state with w written at
// `pos` has >= min_lz leading zero bits in output word 7. Candidate windows grow from the expected
// witness size (2^min_lz) upwards: a window much larger than that only burns permutations behind
// the witness before the early exit can see it. (Host-driven form, behind the qpzk_fri_pow hook; the proof
// pipeline grinds from the device-resident transcript with k_pow_grind_dev.)
static int grind_pow(qpzk_ctx* c, const PowState& ps, u32 pos, u32 min_lz, u64* witness) {
  DevBuf best(c);
  QP(best.alloc(8));
  CU(cudaMemsetAsync(best.p, 0xff, 8, c->stream));
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

// OpeningSet::new for `npolys` coefficient columns at the point whose powers are pw: polynomials longer than 2^12
// coefficients are split over several CTAs (k_eval_at_ext / k_eval_reduce): a batch of 2 to 135 polynomials leaves most
// SMs idle, and what one CTA takes is the latency of its chain of loads
static int launch_eval_at_ext(qpzk_ctx* c, const u64* coeffs, u32 npolys, u64 n, const u64* pw, u64* out) {
  if (!npolys) return QPZK_OK;
  const u64 chunk = n > (1ull << 15) ? (1ull << 13) : (n > (1ull << 12) ? (1ull << 12) : n);
  const u32 chunks = (u32)((n + chunk - 1) / chunk);
  if (chunks == 1) {
    k_eval_at_ext<<<dim3(npolys, 1), 256, 0, c->stream>>>(coeffs, n, chunk, pw, out, nullptr);
    c->launches++;
  } else {
    DevBuf part(c);
    QP(part.alloc((size_t)npolys * chunks * 16));
    k_eval_at_ext<<<dim3(npolys, chunks), 256, 0, c->stream>>>(coeffs, n, chunk, pw, out, part.p);
    k_eval_reduce<<<(npolys + 127) / 128, 128, 0, c->stream>>>(part.p, npolys, chunks, out);
    c->launches += 2;
  }
  CU(cudaGetLastError());
  return QPZK_OK;
}
// divide_by_linear: one CTA up to 2^15 coefficients, the three-phase grid-wide form above that
static int launch_divide_by_linear(qpzk_ctx* c, const u64* p, u64* q, u64 n, const u64* z_dev) {
  if (n <= (1ull << 15)) {
    k_divide_by_linear<<<1, 1024, 0, c->stream>>>(p, q, n, z_dev);
    c->launches++;
    CU(cudaGetLastError());
    return QPZK_OK;
  }
  const u64 per = 16, T = (n + per - 1) / per;
  DevBuf maps(c), carry(c);
  QP(maps.alloc(T * sizeof(DivMap)));
  QP(carry.alloc(T * 16));
  k_divlin_local<<<(unsigned)((T + 255) / 256), 256, 0, c->stream>>>(p, n, per, z_dev, reinterpret_cast<DivMap*>(maps.p));
  k_divlin_scan<<<1, 1024, 0, c->stream>>>(reinterpret_cast<const DivMap*>(maps.p), T, carry.p);
  k_divlin_apply<<<(unsigned)((T + 255) / 256), 256, 0, c->stream>>>(p, q, n, per, z_dev, carry.p);
  c->launches += 3;
  CU(cudaGetLastError());
  return QPZK_OK;
}

static u32 pow_grid_blocks(u32 min_lz) {  // candidates per sweep of k_pow_grind_dev ~ the expected witness
  const u32 lg = min_lz < 12 ? 12 : (min_lz > 16 ? 16 : min_lz);
  return (1u << lg) / 128;
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
  DevBuf zpow(c), res(c), pt(c);
  QP(zpow.alloc(n * 16));
  QP(res.alloc((size_t)b->ncols * 16));
  QP(pt.alloc(16));
  Words16 w;
  w.w[0] = point[0];
  w.w[1] = point[1];
  k_set_words<<<1, 32, 0, c->stream>>>(pt.p, w, 2);
  k_ext_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(pt.p, n, zpow.p);
  c->launches += 2;
  CU(cudaGetLastError());
  QP(launch_eval_at_ext(c, b->coeffs, b->ncols, n, zpow.p, res.p));
  CU(cudaMemcpyAsync(out, res.p, (size_t)b->ncols * 16, cudaMemcpyDeviceToHost, c->stream));
  CU(ctx_wait(c));
  return QPZK_OK;
}

// out[c][i] = n * in[c][(n - i) mod n]: an inverse transform read backwards is the forward transform
// (values on the subgroup from coefficients, natural order both sides).
__global__ void k_forward_from_inverse(const u64* __restrict__ in, u64* __restrict__ out, u32 k, u32 ncols, u64 n_mod_p) {
  const u64 n = (u64)1 << k;
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (u32 c = blockIdx.y; c < ncols; c += gridDim.y)
    out[(u64)c * n + i] = gl_canon(gl_mul(in[(u64)c * n + ((n - i) & (n - 1))], n_mod_p));
}

// constants_sigmas != NULL: the value columns, committed here (CircuitBuilder::build).
// Otherwise commitment_bytes: the serialized constants_sigmas_commitment of a prover restored from files
// (`read_polynomial_batch`, qpzk_batch_from_bytes) - nothing is recomputed except the value columns the
// permutation argument reads, one forward transform of the coefficients.
static int circuit_create_impl(qpzk_ctx* c, const uint8_t* common_bytes, size_t common_len, const uint64_t* digest4,
                               const uint64_t* constants_sigmas, size_t constants_sigmas_words,
                               const uint8_t* commitment_bytes, uint64_t commitment_len, uint32_t import_flags,
                               qpzk_circuit** out) {
  return guarded([&]() -> int {
    if (!c || !common_bytes || !digest4 || (!constants_sigmas && !commitment_bytes) || !out)
      return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    std::unique_ptr<qpzk_circuit> q(new qpzk_circuit());
    q->ctx = c;
    std::string err;
    if (!parse_common_host(common_bytes, common_len, &q->common, &err)) return fail(QPZK_ERR_UNSUPPORTED, err);
    if (!validate_common(q->common, &err)) return fail(QPZK_ERR_UNSUPPORTED, "unsupported circuit configuration: " + err);
    const CommonHost& cm = q->common;
    u32 qdb = 0;
    while ((1ull << qdb) < cm.qdf) qdb++;
    const u64 n = 1ull << cm.degree_bits;
    const u32 ncs = (u32)(cm.num_constants + cm.num_routed);
    if (constants_sigmas && constants_sigmas_words != (size_t)ncs * n)
      return fail(QPZK_ERR_BAD_ARG, "constants_sigmas must hold (num_constants + num_routed_wires) * 2^degree_bits words");
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
    // small per-circuit tables: k_is, the coset-interpolation points and weights, Z_H on the quotient coset.
    // They are staged through host vectors that live until the wait below.
    QP(dev_alloc(c, (cm.k_is.size() ? cm.k_is.size() : 1) * 8, &q->k_is_dev));
    CU(cudaMemcpyAsync(q->k_is_dev, cm.k_is.data(), cm.k_is.size() * 8, cudaMemcpyHostToDevice, c->stream));
    std::vector<u64> aux;
    if (!cm.coset_weights.empty()) {  // subgroup points, then the barycentric weights from the common data
      const size_t np = cm.coset_weights.size();
      u32 bits = 0;
      while ((1ull << bits) < np) bits++;
      aux.resize(2 * np);
      u64 g = glh::root_of_unity(bits);
      aux[0] = 1;
      for (size_t i = 1; i < np; i++) aux[i] = glh::mul(aux[i - 1], g);
      for (size_t i = 0; i < np; i++) aux[np + i] = cm.coset_weights[i];
      QP(dev_alloc(c, aux.size() * 8, &q->coset_aux_dev));
      CU(cudaMemcpyAsync(q->coset_aux_dev, aux.data(), aux.size() * 8, cudaMemcpyHostToDevice, c->stream));
      d.coset_aux = q->coset_aux_dev;
    }
    std::vector<u64> zh(2u << qdb);
    {
      u64 gn = glh::pow(GL_GEN, n), wq = glh::root_of_unity(qdb);
      for (u32 i = 0; i < (1u << qdb); i++) {
        zh[i] = glh::sub(glh::mul(gn, glh::pow(wq, i)), 1);
        zh[(1u << qdb) + i] = glh::inv(zh[i]);
      }
    }
    QP(dev_alloc(c, zh.size() * 8, &q->zh_dev));
    CU(cudaMemcpyAsync(q->zh_dev, zh.data(), zh.size() * 8, cudaMemcpyHostToDevice, c->stream));
    QP(dev_alloc(c, (size_t)ncs * n * 8, &q->cs_values));
    if (constants_sigmas) {
      QP(h2d_copy(c, q->cs_values, constants_sigmas, (size_t)ncs * n * 8));
      // build(): PolynomialBatch::from_values(constants | sigmas), never blinded
      QP(commit_impl(c, q->cs_values, false, false, ncs, (u32)cm.degree_bits, (u32)cm.rate_bits, (u32)cm.cap_height, nullptr,
                     false, 0, &q->cs_batch));
    } else {
      QP(qpzk_batch_from_bytes(c, commitment_bytes, commitment_len, import_flags, &q->cs_batch, nullptr));
      const qpzk_batch* b = q->cs_batch;
      if (b->ncols != ncs || b->salt_cols != 0 || b->degree_bits != cm.degree_bits || b->rate_bits != cm.rate_bits ||
          b->cap_height != cm.cap_height)
        return fail(QPZK_ERR_BAD_ARG, "the serialised constants_sigmas_commitment does not have the shape the common data states");
      DevBuf tmp(c);
      QP(tmp.alloc((size_t)ncs * n * 8));
      QP(launch_ifft(c, b->coeffs, n, tmp.p, n, ncs, (int)cm.degree_bits));  // n^-1 * sum_j c_j w^(-ij)
      k_forward_from_inverse<<<dim3((unsigned)((n + 255) / 256), ncs < 64 ? ncs : 64), 256, 0, c->stream>>>(
          tmp.p, q->cs_values, (u32)cm.degree_bits, ncs, n);
      c->launches++;
      CU(cudaGetLastError());
    }
    q->cs_cap.resize(4ull << cm.cap_height);
    QP(qpzk_batch_cap(q->cs_batch, q->cs_cap.data()));
    {
      const u32 qlb = (u32)cm.degree_bits + qdb;
      RootTab tab_q;
      QP(get_root_tab(c, (int)qlb, false, &tab_q));
      QP(dev_alloc(c, sizeof(u64) << qlb, &q->l0_den_inv_dev));
      k_build_l0_den_inv<<<(unsigned)(((1ull << qlb) + 127) / 128), 128, 0, c->stream>>>(q->l0_den_inv_dev, tab_q,
                                                                                      (u32)cm.degree_bits, qlb);
      c->launches++;
      CU(cudaGetLastError());
    }
    layout_arena(cm, d, &q->lay);
    QP(dev_alloc(c, q->lay.total * 8, &q->arena_dev));
    CU(cudaHostAlloc((void**)&q->arena_host, q->lay.total * 8, cudaHostAllocDefault));
    for (auto& e : q->ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&q->done, cudaEventBlockingSync | cudaEventDisableTiming));
    CU(ctx_wait(c));
    *out = q.release();
    return QPZK_OK;
  });
}

int qpzk_circuit_create(qpzk_ctx* c, const uint8_t* common_bytes, size_t common_len, const uint64_t* digest4,
                        const uint64_t* constants_sigmas, size_t constants_sigmas_words, qpzk_circuit** out) {
  if (!constants_sigmas) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  return circuit_create_impl(c, common_bytes, common_len, digest4, constants_sigmas, constants_sigmas_words, nullptr, 0, 0, out);
}
int qpzk_circuit_create_from_commitment(qpzk_ctx* c, const uint8_t* common_bytes, size_t common_len,
                                        const uint64_t* digest4, const uint8_t* commitment_bytes, uint64_t commitment_len,
                                        uint32_t flags, qpzk_circuit** out) {
  if (!commitment_bytes) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  return circuit_create_impl(c, common_bytes, common_len, digest4, nullptr, 0, commitment_bytes, commitment_len, flags, out);
}
/* `constants_sigmas_commitment` of this circuit in the serialized layout (what a patched
 * ProverOnlyCircuitData::to_bytes writes for that field). */
int qpzk_circuit_commitment_size(const qpzk_circuit* q, uint64_t* nbytes) {
  if (!q) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  return qpzk_batch_serialized_size(q->cs_batch, nbytes);
}
int qpzk_circuit_commitment_to_bytes(const qpzk_circuit* q, uint8_t* out, uint64_t capacity) {
  if (!q) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  return qpzk_batch_to_bytes(q->cs_batch, out, capacity);
}

int qpzk_circuit_cap(const qpzk_circuit* q, uint64_t* out, size_t cap_words) {
  if (!q || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  if (cap_words < q->cs_cap.size()) return fail(QPZK_ERR_BAD_ARG, "buffer too small for 2^cap_height digests");
  memcpy(out, q->cs_cap.data(), q->cs_cap.size() * 8);
  return QPZK_OK;
}
int qpzk_circuit_info(const qpzk_circuit* q, uint32_t* out) {
  if (!q || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  const CommonHost& cm = q->common;
  const uint32_t v[8] = {(uint32_t)cm.degree_bits, (uint32_t)cm.rate_bits, (uint32_t)cm.cap_height, (uint32_t)cm.num_wires,
                         (uint32_t)cm.num_routed, (uint32_t)cm.num_challenges, cm.hiding ? QPZK_SALT_SIZE : 0u,
                         (uint32_t)cm.num_public_inputs};
  memcpy(out, v, sizeof v);
  return QPZK_OK;
}
// VerifierOnlyCircuitData bytes (cap_height, constants_sigmas_cap, circuit_digest); returns length.
size_t qpzk_circuit_verifier_only(const qpzk_circuit* q, uint8_t* out, size_t cap) {
  if (!q) return 0;
  try {
    ByteWriter w;
    w.u(q->common.cap_height, 8);
    w.felts(q->cs_cap.data(), q->cs_cap.size());
    w.felts(q->digest, 4);
    if (out && w.b.size() <= cap) memcpy(out, w.b.data(), w.b.size());
    return w.b.size();
  } catch (...) {
    return 0;
  }
}
void qpzk_circuit_free(qpzk_circuit* q) {
  if (!q) return;
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

// ---- H11-H14: the FRI prover as a resumable object (prove_openings, fri_committed_trees,
// fri_prover_query_rounds of qp-plonky2 fri/oracle.rs, fri/prover.rs). Every challenge it needs is a
// DEVICE pointer: into the transcript when qpzk_prove drives it, into its own small scratch when the caller
// owns the transcript (qpzk_fri_* hooks): begin -> [commit_round -> (observe cap, squeeze beta) -> fold]*
// -> final_poly -> queries. ----
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
  u64* chal = nullptr;        // hooks only: zeta[2] | zeta_next[2] | alpha[2] | beta[2]
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
    dev_free(c, chal);
  }
};

namespace qpzk {

// prove_openings up to the polynomial that enters FRI: batch 0 = every polynomial of the four oracles
// at zeta, batch 1 = the Z polynomials at g*zeta; final = q0 * alpha^(len batch 1) + q1 (no multiply-by-X).
static int fri_begin(qpzk_circuit* q, const qpzk_batch* const* oracles, const u64* zeta_dev, const u64* zeta_next_dev,
                     const u64* alpha_dev, qpzk_fri* F) {
  qpzk_ctx* c = q->ctx;
  const CircuitDesc& d = q->desc;
  const u32 k = d.degree_bits, r = d.rate_bits, nch = d.num_challenges;
  const u64 n = 1ull << k, N = n << r;
  F->q = q;
  F->c = c;
  u32 total_polys = 0;
  for (int o = 0; o < 4; o++) {
    F->oracles[o] = oracles[o];
    total_polys += oracles[o]->ncols;
  }
  DevBuf apow(c), comp0(c), comp1(c), q0(c), q1(c);
  QP(apow.alloc((size_t)total_polys * 16));
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
  k_ext_powers<<<(unsigned)((total_polys + 127) / 128), 128, 0, c->stream>>>(alpha_dev, total_polys, apow.p);  // alpha^j
  k_fri_compose<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(pl0, n, apow.p, comp0.p);
  k_fri_compose<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(pl1, n, apow.p, comp1.p);
  c->launches += 3;
  QP(launch_divide_by_linear(c, comp0.p, q0.p, n, zeta_dev));
  QP(launch_divide_by_linear(c, comp1.p, q1.p, n, zeta_next_dev));
  k_ext_axpy<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(q0.p, q1.p, apow.p + 2ull * nch, n, F->fpoly);
  c->launches++;
  CU(cudaGetLastError());
  F->coeffs_cur = F->fpoly;
  F->cur_n = n;
  F->cur_k = k;
  return QPZK_OK;
}

// One commit-phase round: LDE of the current coefficients on the coset shift*<w>, leaves of 2^arity_bits
// extension evaluations, MerkleTree::new. Returns the device pointer of the cap.
static int fri_commit_round(qpzk_fri* F, const u64** cap_dev) {
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
  *cap_dev = cap_ptr(t.levels, t.log_n, h);
  return QPZK_OK;
}

// Fold in coefficient space: chunks of 2^arity_bits coefficients combined with powers of beta.
static int fri_fold(qpzk_fri* F, const u64* beta_dev) {
  qpzk_ctx* c = F->c;
  const CommonHost& cm = F->q->common;
  if (F->round >= cm.arities.size() || F->trees.size() != F->round + 1) return fail(QPZK_ERR_BAD_ARG, "FRI fold out of order");
  const u64 ab = cm.arities[F->round];
  u64* dst = F->flip ? F->fold_b : F->fold_a;
  F->flip = !F->flip;
  u64 n_out = F->cur_n >> ab;
  k_fri_fold<<<(unsigned)((n_out + 127) / 128), 128, 0, c->stream>>>(F->coeffs_cur, F->cur_n, (u32)ab, beta_dev, dst);
  c->launches++;
  CU(cudaGetLastError());
  F->coeffs_cur = dst;
  F->cur_n = n_out;
  F->cur_k -= (u32)ab;
  for (u64 e = 0; e < ab; e++) F->shift = glh::mul(F->shift, F->shift);
  F->round++;
  return QPZK_OK;
}

// fri_prover_query_rounds for nq indices at once (indices on the device). init_out[o] = nq x (salted row | path),
// step_out[s] = nq x (2^arity ext evaluations | path): device destinations.
static int fri_queries(qpzk_fri* F, const u64* xidx_dev, u32 nq, u64* const* init_out, u64* const* step_out) {
  qpzk_ctx* c = F->c;
  const CircuitDesc& d = F->q->desc;
  const u32 h = (u32)F->q->common.cap_height, lb = d.degree_bits + d.rate_bits;
  for (int o = 0; o < 4; o++) {
    const qpzk_batch* b = F->oracles[o];
    k_gather_openings<<<nq, 128, 0, c->stream>>>(b->lde, 1, b->lde_stride, b->width(), b->levels, lb, h, xidx_dev, 0, b->leaf0,
                                                 b->leaf1, init_out[o]);
    c->launches++;
  }
  u32 sh = 0;
  for (size_t s = 0; s < F->trees.size(); s++) {
    const FriTree& t = F->trees[s];
    sh += t.arity_bits;
    const u32 width = 2u << t.arity_bits;
    k_gather_openings<<<nq, 128, 0, c->stream>>>(t.leaves, width, 1, width, t.levels, t.log_n, h, xidx_dev, sh, 0,
                                                 1ull << t.log_n, step_out[s]);
    c->launches++;
  }
  CU(cudaGetLastError());
  return QPZK_OK;
}

// ---- H8: Z and partial products on the subgroup: zs_vals [nch*(1+npp)][n] = Z_0..Z_{nch-1}, then the
// partial products of each challenge (all_wires_permutation_partial_products + the running product) ----
static int compute_zs_partial_products(qpzk_circuit* q, const u64* wires_dev, const Challenges* chal_dev, DevBuf* zs_vals) {
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
      wires_dev, q->cs_values, q->k_is_dev, d, chal_dev, tab_n, chunk_q.p, row_prod.p);
  k_prefix_product<<<nch, 1024, 0, c->stream>>>(row_prod.p, zs_vals->p, n);
  k_partial_products<<<dim3((unsigned)((n + 127) / 128), nch), 128, 0, c->stream>>>(chunk_q.p, zs_vals->p, nch, npp, n,
                                                                                   zs_vals->p + (size_t)nch * n);
  c->launches += 3;
  CU(cudaGetLastError());
  return QPZK_OK;
}

// ---- H9: compute_quotient_polys: vanishing(x)/Z_H(x) on the quotient coset from the three committed
// oracles -> values [nch][2^(k+qdb)]; then coset IFFT -> coefficients [nch][qdf*n] (= nch*qdf chunks of n).
// apw_dev: alpha powers [nch][QPZK_APW_STRIDE] (k_alpha_powers); chal_dev / pi_hash_dev: device pointers.
// Whole proof: the kernel writes natural order, ready for the IFFT. One rank of a sharded proof evaluates only
// the points it holds - leaf positions [leaf0, leaf1) of the batches - and writes BIT-REVERSED order, so that
// its part is one contiguous block per challenge for the all-gather; quotient_values_to_chunks undoes the
// permutation. (Sharding needs the quotient domain to be the whole LDE domain: quotient_degree_bits ==
// rate_bits, as in every standard configuration.) ----
static int compute_quotient_values(qpzk_circuit* q, const qpzk_batch* wires_b, const qpzk_batch* zs_b, const u64* pi_hash_dev,
                                   const Challenges* chal_dev, const u64* apw_dev, bool sharded, DevBuf* qvals) {
  qpzk_ctx* c = q->ctx;
  const CommonHost& cm = q->common;
  const CircuitDesc& d = q->desc;
  const u32 k = d.degree_bits, r = d.rate_bits, nch = d.num_challenges, qdb = d.quotient_degree_bits;
  const u32 qlb = k + qdb;
  const u64 qlde = 1ull << qlb;
  u64 q0 = 0, q1 = qlde;
  if (sharded) {
    if (qdb != r) return fail(QPZK_ERR_UNSUPPORTED, "a sharded proof needs quotient_degree_factor == 2^rate_bits");
    q0 = wires_b->leaf0;
    q1 = wires_b->leaf1;
    if (zs_b->leaf0 != q0 || zs_b->leaf1 != q1) return fail(QPZK_ERR_BAD_ARG, "oracle shards differ");
  }
  QP(qvals->alloc((size_t)nch * qlde * 8));
  RootTab tab_q;
  QP(get_root_tab(c, (int)qlb, false, &tab_q));
  const u64* zh = q->zh_dev;
  const u64* zh_inv = q->zh_dev + (1u << qdb);
  const unsigned grid = (unsigned)((q1 - q0 + 127) / 128);
  if (cm.recursion)
    k_quotient<true><<<grid, 128, 0, c->stream>>>(
        q->cs_batch->lde, wires_b->lde, zs_b->lde, q->cs_batch->lde_stride, wires_b->lde_stride, zs_b->lde_stride, r - qdb,
        q->k_is_dev, d, chal_dev, pi_hash_dev, zh, zh_inv, apw_dev, q->l0_den_inv_dev, tab_q, q0, q1, sharded ? 1u : 0u, qvals->p);
  else
    k_quotient<false><<<grid, 128, 0, c->stream>>>(
        q->cs_batch->lde, wires_b->lde, zs_b->lde, q->cs_batch->lde_stride, wires_b->lde_stride, zs_b->lde_stride, r - qdb,
        q->k_is_dev, d, chal_dev, pi_hash_dev, zh, zh_inv, apw_dev, q->l0_den_inv_dev, tab_q, q0, q1, sharded ? 1u : 0u, qvals->p);
  c->launches++;
  CU(cudaGetLastError());
  return QPZK_OK;
}
static int quotient_values_to_chunks(qpzk_circuit* q, DevBuf* qvals, bool bitrev_in, DevBuf* qcoeffs) {
  qpzk_ctx* c = q->ctx;
  const CircuitDesc& d = q->desc;
  const u32 nch = d.num_challenges, qlb = d.degree_bits + d.quotient_degree_bits;
  const u64 qlde = 1ull << qlb;
  QP(qcoeffs->alloc((size_t)nch * qlde * 8));
  const u64* vals = qvals->p;
  DevBuf nat(c);
  if (bitrev_in) {  // gathered shards arrive in leaf (bit-reversed) order
    QP(nat.alloc((size_t)nch * qlde * 8));
    k_bitrev_rows<<<dim3((unsigned)((qlde + 255) / 256), nch), 256, 0, c->stream>>>(qvals->p, nat.p, qlde, 0, qlde, (int)qlb, nch);
    c->launches++;
    vals = nat.p;
  }
  // coset IFFT: values on g*<w> -> coefficients; then split into qdf chunks of n (contiguous already)
  QP(launch_ifft(c, vals, qlde, qcoeffs->p, qlde, nch, (int)qlb));
  RootTab tab_ginv;
  QP(get_pow_tab(c, glh::inv(GL_GEN), (int)qlb, &tab_ginv));
  k_scale_by_powers<<<dim3((unsigned)((qlde + 255) / 256), nch), 256, 0, c->stream>>>(qcoeffs->p, qlde, tab_ginv);
  c->launches++;
  CU(cudaGetLastError());
  return QPZK_OK;
}
static int compute_quotient_chunks(qpzk_circuit* q, const qpzk_batch* wires_b, const qpzk_batch* zs_b, const u64* pi_hash_dev,
                                   const Challenges* chal_dev, const u64* apw_dev, DevBuf* qcoeffs) {
  DevBuf qvals(q->ctx);
  QP(compute_quotient_values(q, wires_b, zs_b, pi_hash_dev, chal_dev, apw_dev, false, &qvals));
  return quotient_values_to_chunks(q, &qvals, false, qcoeffs);
}

static inline void tr_step(qpzk_ctx* c, TranscriptDev* T, const u64* src, u32 n, u32 src_mode, u64 m, u64* copy, u32 nsq1,
                           u32 dst1, u32 nsq2, u32 dst2, u32 post, u64 aux) {
  k_transcript_step<<<1, 16, 0, c->stream>>>(T, src, n, src_mode, m, copy, nsq1, dst1, nsq2, dst2, post, aux);
  c->launches++;
}

}  // namespace qpzk

// One proof being enqueued, phase by phase: the order of operations of `prove()` with NO host
// synchronisation; everything lands in the circuit's arena and prove_finish waits once and serialises.
//
// A whole proof runs the five phases back to back. One rank of a multi-GPU proof runs the same
// phases on its range of cap subtrees [sub_begin, sub_end) - whole LDE cosets - and the caller exchanges, on
// the context's stream between phases,
//   after wires / zs / quotient_commit : all-gather of the 2^cap_height subtree roots, in place on cap_dev(i)
//   after quotient_eval                : all-gather of the quotient values, in place on qvals (bit-reversed
//                                        order: a rank's leaves are one contiguous block per challenge)
//   after fri                          : sum (all-reduce) of the opened rows of the three sharded oracles - a
//                                        rank writes zeros for leaves it does not hold.
// IFFTs, Z / partial products, openings and FRI are replicated: every rank holds every coefficient, runs the
// identical transcript on its own device and ends with the identical arena.
struct qpzk_sprove {
  qpzk_circuit* q = nullptr;
  u32 sub_begin = 0, sub_end = 0, flags = 0;
  bool sharded = false;
  int phase = 0;
  const u64 *salt_w = nullptr, *salt_z = nullptr, *salt_q = nullptr;
  const u64* wires_dev = nullptr;
  DevBuf wires_up, zs_vals, apw, qvals, qcoeffs, salt_buf;
  bool seeded = false;   // QPZK_PROVE_SEEDED_SALTS: the salts of every blinded oracle are drawn on the device
  SaltSeed seed;
  std::unique_ptr<qpzk_batch> wires_b, zs_b, q_b;
  qpzk_fri F;
  explicit qpzk_sprove(qpzk_circuit* q_)
      : q(q_), wires_up(q_->ctx), zs_vals(q_->ctx), apw(q_->ctx), qvals(q_->ctx), qcoeffs(q_->ctx), salt_buf(q_->ctx) {}

  void set_seed(const u64* seed_words) {  // flags bit 2: salts_wires is a 32-byte seed on the HOST
    seeded = (flags & 4) && q->common.hiding;
    if (seeded) memcpy(seed.key, seed_words, sizeof seed.key);
  }
  int commit(const u64* in, bool is_coeffs, u32 ncols, const u64* salts, u32 oracle, std::unique_ptr<qpzk_batch>* out) {
    const CommonHost& cm = q->common;
    qpzk_ctx* c = q->ctx;
    const u32 salt_cols = cm.hiding ? QPZK_SALT_SIZE : 0;
    bool salts_host = !(flags & 2);
    if (cm.hiding && seeded) {
      const u64 count = (u64)QPZK_SALT_SIZE << (q->desc.degree_bits + q->desc.rate_bits);
      if (!salt_buf.p) QP(salt_buf.alloc(count * 8));
      k_salts_chacha8<<<(unsigned)((count / 8 + 127) / 128), 128, 0, c->stream>>>(seed, oracle, count, salt_buf.p);
      c->launches++;
      CU(cudaGetLastError());
      salts = salt_buf.p;
      salts_host = false;
    }
    qpzk_batch* raw = nullptr;
    QP(commit_impl(c, in, false, is_coeffs, ncols, q->desc.degree_bits, q->desc.rate_bits, (u32)cm.cap_height,
                   cm.hiding ? salts : nullptr, salts_host, salt_cols, &raw, sub_begin, sub_end, false));
    out->reset(raw);
    return QPZK_OK;
  }

  // ---- transcript start + (2) commit wires ----
  int phase_wires(const u64* wires_in, const u64* pis, u32 npi) {
    qpzk_ctx* c = q->ctx;
    const CircuitDesc& d = q->desc;
    const u64 n = 1ull << d.degree_bits;
    TranscriptDev* T = q->transcript();
    // the transcript up to the first commitment is host work on inputs only: circuit digest | H(public inputs)
    TranscriptInit init;
    {
      host_hash_no_pad(pis, npi, init.pi_hash);
      HostChallenger ch;
      ch.observe_n(q->digest, 4);
      ch.observe_n(init.pi_hash, 4);   // the eighth observation runs the duplex
      memcpy(init.state, ch.state, sizeof init.state);
    }
    q->pis.assign(npi, 0);
    for (u32 i = 0; i < npi; i++) q->pis[i] = pis[i] >= GL_P ? pis[i] - GL_P : pis[i];
    q->want_trace = flags & 1;
    memset(q->stage_ms, 0, sizeof q->stage_ms);
    k_transcript_init<<<1, 32, 0, c->stream>>>(T, init);
    c->launches++;
    CU(cudaEventRecord(q->ev[0], c->stream));
    wires_dev = wires_in;
    if (!(flags & 2)) {
      QP(wires_up.alloc((size_t)d.num_wires * n * 8));
      QP(h2d_copy(c, wires_up.p, wires_in, (size_t)d.num_wires * n * 8));
      wires_dev = wires_up.p;
    }
    QP(commit(wires_dev, false, d.num_wires, salt_w, 0, &wires_b));
    phase = 1;
    return QPZK_OK;
  }

  // ---- observe the wires cap; (4,5) Z + partial products, commit ----
  int phase_zs() {
    qpzk_ctx* c = q->ctx;
    const CircuitDesc& d = q->desc;
    const ArenaLayout& L = q->lay;
    const u32 k = d.degree_bits, r = d.rate_bits, h = (u32)q->common.cap_height, nch = d.num_challenges;
    const u64 n = 1ull << k;
    TranscriptDev* T = q->transcript();
    u64* A = q->arena_dev;
    tr_step(c, T, cap_ptr(wires_b->levels, k + r, h), L.capw, 0, 0, A + L.caps, nch, QPZK_TR_OFF(ch.beta), nch,
            QPZK_TR_OFF(ch.gamma), 0, 0);
    CU(cudaEventRecord(q->ev[1], c->stream));  // stage 0: wires commit
    const u32 nzs = nch * (1 + d.num_partial_products);
    QP(compute_zs_partial_products(q, wires_dev, &T->ch, &zs_vals));
    QP(commit(zs_vals.p, false, nzs, salt_z, 1, &zs_b));
    if (q->want_trace) {
      q->tr_zs_pp.resize((size_t)nzs * n);
      CU(cudaMemcpyAsync(q->tr_zs_pp.data(), zs_vals.p, (size_t)nzs * n * 8, cudaMemcpyDeviceToHost, c->stream));
      CU(ctx_wait(c));
    }
    phase = 2;
    return QPZK_OK;
  }

  // ---- observe the zs cap; (6) quotient values on this rank's points ----
  int phase_quotient_eval() {
    qpzk_ctx* c = q->ctx;
    const CircuitDesc& d = q->desc;
    const ArenaLayout& L = q->lay;
    const u32 k = d.degree_bits, r = d.rate_bits, h = (u32)q->common.cap_height, nch = d.num_challenges;
    TranscriptDev* T = q->transcript();
    u64* A = q->arena_dev;
    tr_step(c, T, cap_ptr(zs_b->levels, k + r, h), L.capw, 0, 0, A + L.caps + L.capw, nch, QPZK_TR_OFF(ch.alpha), 0, 0, 0, 0);
    CU(cudaEventRecord(q->ev[2], c->stream));  // stage 1: Z/pp + commit
    QP(apw.alloc((size_t)2 * QPZK_APW_STRIDE * 8));
    k_alpha_powers<<<nch, 256, 0, c->stream>>>(T, QPZK_APW_STRIDE, apw.p);
    c->launches++;
    QP(compute_quotient_values(q, wires_b.get(), zs_b.get(), T->pi_hash, &T->ch, apw.p, sharded, &qvals));
    phase = 3;
    return QPZK_OK;
  }

  // ---- (6,7) quotient coefficients and their commitment ----
  int phase_quotient_commit() {
    qpzk_ctx* c = q->ctx;
    const CircuitDesc& d = q->desc;
    const u32 nch = d.num_challenges, qlb = d.degree_bits + d.quotient_degree_bits;
    const u64 qlde = 1ull << qlb;
    QP(quotient_values_to_chunks(q, &qvals, sharded, &qcoeffs));
    QP(commit(qcoeffs.p, true, nch * d.qdf, salt_q, 2, &q_b));
    if (q->want_trace) {
      q->tr_quotient.resize((size_t)nch * qlde);
      CU(cudaMemcpyAsync(q->tr_quotient.data(), qcoeffs.p, (size_t)nch * qlde * 8, cudaMemcpyDeviceToHost, c->stream));
      CU(ctx_wait(c));
    }
    phase = 4;
    return QPZK_OK;
  }

  // ---- observe the quotient cap; (8) openings; (9) FRI: combine, commit phase, proof of work, queries ----
  int phase_fri() {
    qpzk_ctx* c = q->ctx;
    const CommonHost& cm = q->common;
    const CircuitDesc& d = q->desc;
    const ArenaLayout& L = q->lay;
    const u32 k = d.degree_bits, r = d.rate_bits, h = (u32)cm.cap_height, nch = d.num_challenges;
    const u64 n = 1ull << k, N = n << r;
    TranscriptDev* T = q->transcript();
    u64* A = q->arena_dev;
    tr_step(c, T, cap_ptr(q_b->levels, k + r, h), L.capw, 0, 0, A + L.caps + 2 * (size_t)L.capw, 2, QPZK_TR_OFF(zeta), 0, 0, 1,
            glh::root_of_unity(k));
    CU(cudaEventRecord(q->ev[3], c->stream));  // stage 2: quotient + commit

    // openings: observe order constants, sigmas, wires, zs, partial_products, quotient (= oracle order), then zs_next
    const qpzk_batch* oracles[4] = {q->cs_batch, wires_b.get(), zs_b.get(), q_b.get()};
    const u32 total_polys = L.total_polys;
    {
      DevBuf zpow(c), zpow_next(c);
      QP(zpow.alloc(n * 16));
      QP(zpow_next.alloc(n * 16));
      u64* open_dev = A + L.opens;
      k_ext_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(T->zeta, n, zpow.p);
      k_ext_powers<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(T->zeta_next, n, zpow_next.p);
      c->launches += 2;
      u32 off = 0;
      for (auto* b : oracles) {
        QP(launch_eval_at_ext(c, b->coeffs, b->ncols, n, zpow.p, open_dev + 2ull * off));
        off += b->ncols;
      }
      QP(launch_eval_at_ext(c, zs_b->coeffs, nch, n, zpow_next.p, open_dev + 2ull * off));
      tr_step(c, T, open_dev, 2 * (total_polys + nch), 0, 0, nullptr, 2, QPZK_TR_OFF(fri_alpha), 0, 0, 0, 0);
    }
    CU(cudaEventRecord(q->ev[4], c->stream));  // stage 3: openings

    QP(fri_begin(q, oracles, T->zeta, T->zeta_next, T->fri_alpha, &F));
    if (q->want_trace) {
      std::vector<u64> soa(2 * n);
      CU(cudaMemcpyAsync(soa.data(), F.fpoly, n * 16, cudaMemcpyDeviceToHost, c->stream));
      CU(ctx_wait(c));
      q->tr_final_poly.resize(2 * n);
      for (u64 m = 0; m < n; m++) {
        q->tr_final_poly[2 * m] = soa[m];
        q->tr_final_poly[2 * m + 1] = soa[n + m];
      }
    }
    CU(cudaEventRecord(q->ev[5], c->stream));  // stage 4: FRI combine

    for (size_t round = 0; round < cm.arities.size(); round++) {
      const u64* cap_dev = nullptr;
      QP(fri_commit_round(&F, &cap_dev));
      tr_step(c, T, cap_dev, L.capw, 0, 0, A + L.fri_caps + round * (size_t)L.capw, 2, QPZK_TR_OFF(fri_beta) + 2 * (u32)round, 0,
              0, 0, 0);
      QP(fri_fold(&F, T->fri_beta[round]));
    }
    // the polynomial left after the last fold, observed (and stored) as interleaved extension coefficients
    tr_step(c, T, F.coeffs_cur, 2 * (u32)F.cur_n, 1, F.cur_n, A + L.final_poly, 0, 0, 0, 0, 0, 0);
    CU(cudaEventRecord(q->ev[6], c->stream));  // stage 5: FRI commit phase

    // proof of work, then the response and the query indices
    k_pow_grind_dev<<<pow_grid_blocks(cm.pow_bits), 128, 0, c->stream>>>(T, cm.pow_bits);
    c->launches++;
    const u32 nq = (u32)cm.num_queries;
    tr_step(c, T, &T->pow_witness, 1, 0, 0, nullptr, 1 + nq, QPZK_TR_OFF(pow_resp), 0, 0, 2, N - 1);
    CU(cudaEventRecord(q->ev[7], c->stream));  // stage 6: PoW

    u64* init_out[4];
    u64* step_out[QPZK_MAX_FRI_ROUNDS];
    for (int o = 0; o < 4; o++) init_out[o] = A + L.init_open[o];
    for (size_t s = 0; s < cm.arities.size(); s++) step_out[s] = A + L.step_open[s];
    QP(fri_queries(&F, T->xidx, nq, init_out, step_out));
    CU(cudaEventRecord(q->ev[8], c->stream));  // stage 7: queries
    phase = 5;
    return QPZK_OK;
  }

  // ---- the arena goes back to the host ----
  int phase_download() {
    qpzk_ctx* c = q->ctx;
    CU(cudaMemcpyAsync(q->arena_host, q->arena_dev, q->lay.total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(q->done, c->stream));
    phase = 6;
    return QPZK_OK;
  }
};

namespace qpzk {

static int prove_enqueue(qpzk_circuit* q, const u64* wires_in, const u64* pis, u32 npi, const u64* salt_w,
                         const u64* salt_z, const u64* salt_q, u32 flags) {
  qpzk_sprove run(q);
  run.flags = flags;
  run.salt_w = salt_w;
  run.salt_z = salt_z;
  run.salt_q = salt_q;
  run.set_seed(salt_w);
  QP(run.phase_wires(wires_in, pis, npi));
  QP(run.phase_zs());
  QP(run.phase_quotient_eval());
  QP(run.phase_quotient_commit());
  QP(run.phase_fri());
  QP(run.phase_download());
  // the FRI state, the three batches and every scratch buffer are released here in stream order
  return QPZK_OK;
}

// Wait for the proof in flight (the way the context waits), then ProofWithPublicInputs::to_bytes from the
// pinned arena.
static int prove_finish(qpzk_circuit* q, std::vector<uint8_t>* proof) {
  qpzk_ctx* c = q->ctx;
  const CommonHost& cm = q->common;
  const CircuitDesc& d = q->desc;
  const ArenaLayout& L = q->lay;
  cudaError_t e;
  if (c->blocking_sync)
    e = cudaEventSynchronize(q->done);
  else if (c->yield_sync) {
    while ((e = cudaEventQuery(q->done)) == cudaErrorNotReady) sched_yield();
  } else {
    while ((e = cudaEventQuery(q->done)) == cudaErrorNotReady) {
    }
  }
  CU(e);
  for (int i = 0; i < 8; i++) cudaEventElapsedTime(&q->stage_ms[i], q->ev[i], q->ev[i + 1]);
  const u64* H = q->arena_host;
  const TranscriptDev* T = reinterpret_cast<const TranscriptDev*>(H);
  const u32 nch = d.num_challenges, npp = d.num_partial_products, nw = d.num_wires, qdf = d.qdf;
  const u32 nzs = nch * (1 + npp), nq = (u32)cm.num_queries;
  if (T->pow_witness == ~0ull) return fail(QPZK_ERR_CUDA, "proof of work failed");
  if (cm.pow_bits && (T->pow_resp >> (64 - cm.pow_bits)) != 0) return fail(QPZK_ERR_CUDA, "pow response mismatch");
  ByteWriter w;
  w.b.reserve(L.total * 8 + 64 * nq + 64);
  w.felts(H + L.caps, 3 * (size_t)L.capw);
  {
    // openings in serialised order: constants, sigmas, wires, zs, zs_next, partial_products, quotient
    const u64* o = H + L.opens;
    size_t n_cs = (size_t)(cm.num_constants + cm.num_routed), off_w = n_cs, off_z = off_w + nw;
    size_t off_pp = off_z + nch, off_q = off_z + nzs, off_next = L.total_polys;
    w.felts(o, 2 * n_cs);
    w.felts(o + 2 * off_w, 2ull * nw);
    w.felts(o + 2 * off_z, 2ull * nch);
    w.felts(o + 2 * off_next, 2ull * nch);
    w.felts(o + 2 * off_pp, 2ull * nch * npp);
    w.felts(o + 2 * off_q, 2ull * nch * qdf);
  }
  w.felts(H + L.fri_caps, cm.arities.size() * (size_t)L.capw);
  for (u32 qi = 0; qi < nq; qi++) {
    for (int o = 0; o < 4; o++) {
      const u32 width = L.width[o];
      const size_t per = width + 4ull * L.L0;
      const u64* p = H + L.init_open[o] + per * qi;
      w.felts(p, width);
      w.u(L.L0, 1);
      w.felts(p + width, 4ull * L.L0);
    }
    for (size_t s = 0; s < cm.arities.size(); s++) {
      const u32 width = L.step_width[s], Ls = L.step_L[s];
      const u64* p = H + L.step_open[s] + (size_t)(width + 4ull * Ls) * qi;
      w.felts(p, width);
      w.u(Ls, 1);
      w.felts(p + width, 4ull * Ls);
    }
  }
  w.felts(H + L.final_poly, 2 * (size_t)L.final_len);
  w.u(T->pow_witness, 8);
  w.u(q->pis.size(), 8);
  w.felts(q->pis.data(), q->pis.size());
  *proof = std::move(w.b);
  if (q->want_trace) {
    q->tr_challenges.clear();
    for (u32 i = 0; i < nch; i++) q->tr_challenges.push_back(T->ch.beta[i]);
    for (u32 i = 0; i < nch; i++) q->tr_challenges.push_back(T->ch.gamma[i]);
    for (u32 i = 0; i < nch; i++) q->tr_challenges.push_back(T->ch.alpha[i]);
    q->tr_challenges.push_back(T->zeta[0]); q->tr_challenges.push_back(T->zeta[1]);
    q->tr_challenges.push_back(T->fri_alpha[0]); q->tr_challenges.push_back(T->fri_alpha[1]);
    for (size_t s = 0; s < cm.arities.size(); s++) {
      q->tr_challenges.push_back(T->fri_beta[s][0]);
      q->tr_challenges.push_back(T->fri_beta[s][1]);
    }
  }
  return QPZK_OK;
}

}  // namespace qpzk

extern "C" {

// Caller-provided challenges for the per-stage hooks: written into a small device buffer by a kernel.
static int upload_challenges(qpzk_circuit* q, const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas,
                             const uint64_t* pi_hash, DevBuf* buf /* Challenges | pi_hash[4] */) {
  qpzk_ctx* c = q->ctx;
  QP(buf->alloc(sizeof(Challenges) + 32));
  const u32 nch = q->desc.num_challenges;
  Words16 w;
  memset(&w, 0, sizeof w);
  for (u32 i = 0; i < nch; i++) {
    if (betas) w.w[i] = betas[i];
    if (gammas) w.w[QPZK_MAX_CHALLENGES + i] = gammas[i];
    if (alphas) w.w[2 * QPZK_MAX_CHALLENGES + i] = alphas[i];
  }
  for (u32 i = 0; i < 4; i++)
    if (pi_hash) w.w[3 * QPZK_MAX_CHALLENGES + i] = pi_hash[i];
  static_assert(sizeof(Challenges) == 3 * QPZK_MAX_CHALLENGES * 8 && 3 * QPZK_MAX_CHALLENGES + 4 <= 16, "Words16 layout");
  k_set_words<<<1, 32, 0, c->stream>>>(buf->p, w, 3 * QPZK_MAX_CHALLENGES + 4);
  c->launches++;
  CU(cudaGetLastError());
  return QPZK_OK;
}

int qpzk_zs_partial_products(qpzk_circuit* q, const uint64_t* wires, const uint64_t* betas, const uint64_t* gammas,
                             uint64_t* out) {
  return guarded([&]() -> int {
    if (!q || !wires || !betas || !gammas || !out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    qpzk_ctx* c = q->ctx;
    CU(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> lk(q->mu);
    const CircuitDesc& d = q->desc;
    const u64 n = 1ull << d.degree_bits;
    DevBuf wd(c), zs(c), chal(c);
    QP(upload_challenges(q, betas, gammas, nullptr, nullptr, &chal));
    QP(wd.alloc((size_t)d.num_wires * n * 8));
    CU(cudaMemcpyAsync(wd.p, wires, (size_t)d.num_wires * n * 8, cudaMemcpyHostToDevice, c->stream));
    QP(compute_zs_partial_products(q, wd.p, reinterpret_cast<const Challenges*>(chal.p), &zs));
    const size_t cnt = (size_t)d.num_challenges * (1 + d.num_partial_products) * n;
    CU(cudaMemcpyAsync(out, zs.p, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    return QPZK_OK;
  });
}

int qpzk_quotient(qpzk_circuit* q, const qpzk_batch* wires_batch, const qpzk_batch* zs_batch, const uint64_t* pi_hash,
                  const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, uint64_t* out_chunks) {
  return guarded([&]() -> int {
    if (!q || !wires_batch || !zs_batch || !pi_hash || !betas || !gammas || !alphas || !out_chunks)
      return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    qpzk_ctx* c = q->ctx;
    const CircuitDesc& d = q->desc;
    if (wires_batch->ctx != c || zs_batch->ctx != c) return fail(QPZK_ERR_BAD_ARG, "batches belong to another context");
    if (wires_batch->degree_bits != d.degree_bits || zs_batch->degree_bits != d.degree_bits ||
        wires_batch->rate_bits != d.rate_bits || zs_batch->rate_bits != d.rate_bits || wires_batch->ncols != d.num_wires ||
        zs_batch->ncols != d.num_challenges * (1 + d.num_partial_products) || wires_batch->sharded() || zs_batch->sharded())
      return fail(QPZK_ERR_BAD_ARG, "batch shape does not match the circuit");
    CU(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> lk(q->mu);
    DevBuf chal(c), apw(c), qc(c);
    QP(upload_challenges(q, betas, gammas, alphas, pi_hash, &chal));
    QP(apw.alloc((size_t)2 * QPZK_APW_STRIDE * 8));
    {
      // alpha powers from the uploaded challenges: the Challenges block sits where k_alpha_powers expects
      // TranscriptDev::ch, so hand it a pointer shifted back by that offset
      const TranscriptDev* fake = reinterpret_cast<const TranscriptDev*>(reinterpret_cast<const char*>(chal.p) -
                                                                         offsetof(TranscriptDev, ch));
      k_alpha_powers<<<d.num_challenges, 256, 0, c->stream>>>(fake, QPZK_APW_STRIDE, apw.p);
      c->launches++;
    }
    const Challenges* ch = reinterpret_cast<const Challenges*>(chal.p);
    QP(compute_quotient_chunks(q, wires_batch, zs_batch, chal.p + 3 * QPZK_MAX_CHALLENGES, ch, apw.p, &qc));
    const size_t cnt = ((size_t)d.num_challenges << (d.degree_bits + d.quotient_degree_bits));
    CU(cudaMemcpyAsync(out_chunks, qc.p, cnt * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    return QPZK_OK;
  });
}

int qpzk_fri_begin(qpzk_circuit* q, const qpzk_batch* wires_batch, const qpzk_batch* zs_batch,
                   const qpzk_batch* quotient_batch, const uint64_t* zeta, const uint64_t* alpha, qpzk_fri** out) {
  return guarded([&]() -> int {
    if (!q || !wires_batch || !zs_batch || !quotient_batch || !zeta || !alpha || !out)
      return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    qpzk_ctx* c = q->ctx;
    const CircuitDesc& d = q->desc;
    const qpzk_batch* oracles[4] = {q->cs_batch, wires_batch, zs_batch, quotient_batch};
    const u32 want[4] = {d.num_constants + d.num_routed, d.num_wires, d.num_challenges * (1 + d.num_partial_products),
                         d.num_challenges * d.qdf};
    for (int o = 0; o < 4; o++)
      if (oracles[o]->ctx != c || oracles[o]->degree_bits != d.degree_bits || oracles[o]->rate_bits != d.rate_bits ||
          oracles[o]->ncols != want[o] || oracles[o]->cap_height != q->common.cap_height || oracles[o]->sharded())
        return fail(QPZK_ERR_BAD_ARG, "oracle shape does not match the circuit");
    CU(cudaSetDevice(c->device));
    std::unique_ptr<qpzk_fri> F(new qpzk_fri());
    F->c = c;
    QP(dev_alloc(c, 8 * 8, &F->chal));
    const u64 wn = glh::root_of_unity(d.degree_bits);
    Words16 w;
    memset(&w, 0, sizeof w);
    w.w[0] = zeta[0]; w.w[1] = zeta[1];
    w.w[2] = glh::mul(zeta[0] % GL_P, wn); w.w[3] = glh::mul(zeta[1] % GL_P, wn);
    w.w[4] = alpha[0]; w.w[5] = alpha[1];
    k_set_words<<<1, 32, 0, c->stream>>>(F->chal, w, 6);
    c->launches++;
    QP(fri_begin(q, oracles, F->chal, F->chal + 2, F->chal + 4, F.get()));
    *out = F.release();
    return QPZK_OK;
  });
}
uint32_t qpzk_fri_num_rounds(const qpzk_fri* f) { return f ? (uint32_t)f->q->common.arities.size() : 0; }
int qpzk_fri_commit_round(qpzk_fri* f, uint64_t* cap_out) {
  return guarded([&]() -> int {
    if (!f || !cap_out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    qpzk_ctx* c = f->c;
    CU(cudaSetDevice(c->device));
    const u64* cap_dev = nullptr;
    QP(fri_commit_round(f, &cap_dev));
    CU(cudaMemcpyAsync(cap_out, cap_dev, (size_t)32 << f->q->common.cap_height, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    return QPZK_OK;
  });
}
int qpzk_fri_fold(qpzk_fri* f, const uint64_t* beta) {
  return guarded([&]() -> int {
    if (!f || !beta) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    qpzk_ctx* c = f->c;
    CU(cudaSetDevice(c->device));
    Words16 w;
    memset(&w, 0, sizeof w);
    w.w[0] = beta[0]; w.w[1] = beta[1];
    k_set_words<<<1, 32, 0, c->stream>>>(f->chal + 6, w, 2);
    c->launches++;
    return fri_fold(f, f->chal + 6);
  });
}
int qpzk_fri_final_poly(qpzk_fri* f, uint64_t* out, size_t cap_words, size_t* len_words) {
  return guarded([&]() -> int {
    if (!f || !len_words) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    qpzk_ctx* c = f->c;
    CU(cudaSetDevice(c->device));
    if (f->round != f->q->common.arities.size()) return fail(QPZK_ERR_BAD_ARG, "FRI rounds not finished");
    const u64 m = f->cur_n;
    *len_words = 2 * m;
    if (!out) return QPZK_OK;
    if (2 * m > cap_words) return fail(QPZK_ERR_BAD_ARG, "buffer too small");
    std::vector<u64> soa(2 * m);
    CU(cudaMemcpyAsync(soa.data(), f->coeffs_cur, 2 * m * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    for (u64 i = 0; i < m; i++) {
      out[2 * i] = soa[i];
      out[2 * i + 1] = soa[m + i];
    }
    return QPZK_OK;
  });
}
int qpzk_fri_query(qpzk_fri* f, uint64_t x_index, uint64_t* out, size_t cap_words, size_t* len_words) {
  return guarded([&]() -> int {
    if (!f || !len_words) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    const CircuitDesc& d = f->q->desc;
    const ArenaLayout& L = f->q->lay;
    qpzk_ctx* c = f->c;
    if (x_index >> (d.degree_bits + d.rate_bits)) return fail(QPZK_ERR_BAD_ARG, "x_index out of range");
    if (f->round != f->q->common.arities.size()) return fail(QPZK_ERR_BAD_ARG, "FRI rounds not finished");
    CU(cudaSetDevice(c->device));
    size_t total = 0, off[4 + QPZK_MAX_FRI_ROUNDS];
    for (int o = 0; o < 4; o++) { off[o] = total; total += f->oracles[o]->width() + 4ull * L.L0; }
    for (size_t s = 0; s < f->trees.size(); s++) { off[4 + s] = total; total += L.step_width[s] + 4ull * L.step_L[s]; }
    *len_words = total;
    if (!out) return QPZK_OK;
    if (total > cap_words) return fail(QPZK_ERR_BAD_ARG, "buffer too small");
    DevBuf buf(c), xi(c);
    QP(buf.alloc(total * 8));
    QP(xi.alloc(8));
    Words16 w;
    memset(&w, 0, sizeof w);
    w.w[0] = x_index;
    k_set_words<<<1, 32, 0, c->stream>>>(xi.p, w, 1);
    c->launches++;
    u64* init_out[4];
    u64* step_out[QPZK_MAX_FRI_ROUNDS];
    for (int o = 0; o < 4; o++) init_out[o] = buf.p + off[o];
    for (size_t s = 0; s < f->trees.size(); s++) step_out[s] = buf.p + off[4 + s];
    QP(fri_queries(f, xi.p, 1, init_out, step_out));
    CU(cudaMemcpyAsync(out, buf.p, total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(ctx_wait(c));
    return QPZK_OK;
  });
}
void qpzk_fri_free(qpzk_fri* f) { delete f; }

static int check_prove_args(qpzk_circuit* q, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
                            uint32_t npi, const uint64_t* sw, const uint64_t* sz, const uint64_t* sq, size_t salt_words,
                            uint32_t flags = 0) {
  if (!q || !wires || (!public_inputs && npi)) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  const CommonHost& cm = q->common;
  if (npi != cm.num_public_inputs) return fail(QPZK_ERR_BAD_ARG, "public input count mismatch");
  if (wires_words != ((size_t)cm.num_wires << cm.degree_bits))
    return fail(QPZK_ERR_BAD_ARG, "wires must hold num_wires * 2^degree_bits words");
  if (cm.hiding && (flags & 4)) {
    if (!sw || salt_words != 4) return fail(QPZK_ERR_BAD_ARG, "seeded salts: salts_wires must point to a 4-word seed (salt_words = 4)");
  } else if (cm.hiding) {
    if (!(sw && sz && sq)) return fail(QPZK_ERR_BAD_ARG, "hiding circuit needs salts");
    if (salt_words != ((size_t)QPZK_SALT_SIZE << (cm.degree_bits + cm.rate_bits)))
      return fail(QPZK_ERR_BAD_ARG, "each salt array must hold 4 * 2^(degree_bits + rate_bits) words");
  }
  return QPZK_OK;
}

int qpzk_prove_begin(qpzk_circuit* q, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
                     uint32_t num_public_inputs, const uint64_t* salts_wires, const uint64_t* salts_zs,
                     const uint64_t* salts_quotient, size_t salt_words, uint32_t flags) {
  return guarded([&]() -> int {
    QP(check_prove_args(q, wires, wires_words, public_inputs, num_public_inputs, salts_wires, salts_zs, salts_quotient, salt_words, flags));
    CU(cudaSetDevice(q->ctx->device));
    std::lock_guard<std::mutex> lk(q->mu);
    if (q->in_flight) return fail(QPZK_ERR_BAD_ARG, "a proof is already in flight on this circuit handle: call qpzk_prove_end first");
    int rc = prove_enqueue(q, wires, public_inputs, num_public_inputs, salts_wires, salts_zs, salts_quotient, flags);
    if (rc != QPZK_OK) {
      ctx_wait(q->ctx);  // drain what was enqueued before the handles it used go away
      return rc;
    }
    q->in_flight = true;
    return QPZK_OK;
  });
}

int qpzk_prove_end(qpzk_circuit* q, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  return guarded([&]() -> int {
    if (!q || !proof_len) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(q->ctx->device));
    std::lock_guard<std::mutex> lk(q->mu);
    if (!q->in_flight) return fail(QPZK_ERR_BAD_ARG, "no proof in flight on this circuit handle");
    q->in_flight = false;
    std::vector<uint8_t> bytes;
    QP(prove_finish(q, &bytes));
    *proof_len = bytes.size();
    if (proof_out) {
      if (bytes.size() > proof_cap) return fail(QPZK_ERR_BAD_ARG, "proof buffer too small");
      memcpy(proof_out, bytes.data(), bytes.size());
    }
    return QPZK_OK;
  });
}

// ---- one proof over several GPUs: every rank runs the phases on its cap subtrees, the caller exchanges
// between them (qpzk_sprove in this file; include/qpzk.h) ----
int qpzk_sprove_begin(qpzk_circuit* q, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
                      uint32_t num_public_inputs, const uint64_t* salts_wires, const uint64_t* salts_zs,
                      const uint64_t* salts_quotient, size_t salt_words, uint32_t flags, uint32_t subtree_begin,
                      uint32_t subtree_end, qpzk_sprove** out) {
  return guarded([&]() -> int {
    if (!out) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    QP(check_prove_args(q, wires, wires_words, public_inputs, num_public_inputs, salts_wires, salts_zs, salts_quotient, salt_words, flags));
    const u32 ncap = 1u << q->common.cap_height;
    if (subtree_begin >= subtree_end || subtree_end > ncap) return fail(QPZK_ERR_BAD_ARG, "bad subtree range");
    if (flags & 1) return fail(QPZK_ERR_UNSUPPORTED, "the parity trace is not kept for sharded proofs");
    CU(cudaSetDevice(q->ctx->device));
    std::lock_guard<std::mutex> lk(q->mu);
    if (q->in_flight) return fail(QPZK_ERR_BAD_ARG, "a proof is already in flight on this circuit handle");
    std::unique_ptr<qpzk_sprove> s(new qpzk_sprove(q));
    s->flags = flags;
    s->sub_begin = subtree_begin;
    s->sub_end = subtree_end;
    s->sharded = subtree_begin != 0 || subtree_end != ncap;
    s->salt_w = salts_wires;
    s->salt_z = salts_zs;
    s->salt_q = salts_quotient;
    s->set_seed(salts_wires);
    int rc = s->phase_wires(wires, public_inputs, num_public_inputs);
    if (rc != QPZK_OK) {
      ctx_wait(q->ctx);
      return rc;
    }
    q->in_flight = true;
    *out = s.release();
    return QPZK_OK;
  });
}

int qpzk_sprove_next(qpzk_sprove* s) {
  return guarded([&]() -> int {
    if (!s) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(s->q->ctx->device));
    std::lock_guard<std::mutex> lk(s->q->mu);
    switch (s->phase) {
      case 1: return s->phase_zs();
      case 2: return s->phase_quotient_eval();
      case 3: return s->phase_quotient_commit();
      case 4: return s->phase_fri();
      case 5: return s->phase_download();
      default: return fail(QPZK_ERR_BAD_ARG, "no phase left: call qpzk_sprove_end");
    }
  });
}

uint32_t qpzk_sprove_phase(const qpzk_sprove* s) { return s ? (uint32_t)s->phase : 0; }

int qpzk_sprove_exchange(const qpzk_sprove* s, uint32_t index, uint64_t** dev_ptr, uint64_t* words, uint64_t* own_begin,
                         uint64_t* own_end, uint32_t* kind) {
  if (!s || !dev_ptr || !words || !own_begin || !own_end || !kind) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  const qpzk_circuit* q = s->q;
  const CircuitDesc& d = q->desc;
  const ArenaLayout& L = q->lay;
  const u32 lb = d.degree_bits + d.rate_bits, h = (u32)q->common.cap_height;
  *kind = 0;
  *dev_ptr = nullptr;
  *words = *own_begin = *own_end = 0;
  if (!s->sharded) return QPZK_OK;
  const qpzk_batch* capb = s->phase == 1 ? s->wires_b.get() : s->phase == 2 ? s->zs_b.get() : s->phase == 4 ? s->q_b.get() : nullptr;
  if (capb) {
    if (index > 0) return QPZK_OK;
    *kind = QPZK_EXCHANGE_ALLGATHER;
    *dev_ptr = const_cast<u64*>(cap_ptr(capb->levels, lb, h));
    *words = L.capw;
    *own_begin = 4ull * s->sub_begin;
    *own_end = 4ull * s->sub_end;
  } else if (s->phase == 3) {
    if (index >= d.num_challenges) return QPZK_OK;
    const u64 qlde = 1ull << (d.degree_bits + d.quotient_degree_bits);
    *kind = QPZK_EXCHANGE_ALLGATHER;
    *dev_ptr = s->qvals.p + (size_t)index * qlde;
    *words = qlde;
    *own_begin = s->wires_b->leaf0;
    *own_end = s->wires_b->leaf1;
  } else if (s->phase == 5) {
    if (index > 0) return QPZK_OK;
    const u64 nq = q->common.num_queries;
    *kind = QPZK_EXCHANGE_SUM;
    *dev_ptr = q->arena_dev + L.init_open[1];
    *words = L.init_open[3] + nq * (L.width[3] + 4ull * L.L0) - L.init_open[1];
    *own_end = *words;
  }
  return QPZK_OK;
}

int qpzk_sprove_end(qpzk_sprove* s, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  if (!s) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  qpzk_circuit* q = s->q;
  int rc = guarded([&]() -> int {
    CU(cudaSetDevice(q->ctx->device));
    std::lock_guard<std::mutex> lk(q->mu);
    q->in_flight = false;
    if (s->phase != 6) {
      ctx_wait(q->ctx);
      return fail(QPZK_ERR_BAD_ARG, "proof abandoned before its last phase");
    }
    if (!proof_len) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
    std::vector<uint8_t> bytes;
    QP(prove_finish(q, &bytes));
    *proof_len = bytes.size();
    if (proof_out) {
      if (bytes.size() > proof_cap) return fail(QPZK_ERR_BAD_ARG, "proof buffer too small");
      memcpy(proof_out, bytes.data(), bytes.size());
    }
    return QPZK_OK;
  });
  delete s;
  return rc;
}

int qpzk_prove(qpzk_circuit* q, const uint64_t* wires, size_t wires_words, const uint64_t* public_inputs,
               uint32_t num_public_inputs, const uint64_t* salts_wires, const uint64_t* salts_zs,
               const uint64_t* salts_quotient, size_t salt_words, uint32_t flags, uint8_t* proof_out, size_t proof_cap,
               size_t* proof_len) {
  if (!proof_len) return fail(QPZK_ERR_BAD_ARG, "NULL argument");
  int rc = qpzk_prove_begin(q, wires, wires_words, public_inputs, num_public_inputs, salts_wires, salts_zs, salts_quotient,
                            salt_words, flags);
  if (rc != QPZK_OK) return rc;
  return qpzk_prove_end(q, proof_out, proof_cap, proof_len);
}

}  // extern "C"
