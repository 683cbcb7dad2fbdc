// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Fiat-Shamir `Challenger` of qp-plonky2 1.1.1 (duplex sponge over Poseidon; SURVEY.md §8(a) H15,
// App. A.6): inputs are buffered up to the rate and OVERWRITE state[0..len) before a permutation;
// outputs are state[0..8) popped from the END.
#pragma once
#include <vector>

#include "poseidon.hpp"

namespace orc {

struct Challenger {
  State state{};
  std::vector<u64> in, out;

  void duplexing() {
    for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
    in.clear();
    poseidon(state);
    out.assign(state.begin(), state.begin() + SPONGE_RATE);
  }
  void observe(u64 x) {
    out.clear();
    in.push_back(x);
    if (in.size() == (size_t)SPONGE_RATE) duplexing();
  }
  void observe_hash(const Hash& h) {
    for (int i = 0; i < 4; i++) observe(h.e[i]);
  }
  void observe_cap(const std::vector<Hash>& cap) {
    for (const Hash& h : cap) observe_hash(h);
  }
  void observe_ext(E2 x) {
    observe(x.a);
    observe(x.b);
  }
  u64 get_challenge() {
    if (!in.empty() || out.empty()) duplexing();
    u64 v = out.back();
    out.pop_back();
    return v;
  }
  E2 get_ext_challenge() {
    u64 a = get_challenge();
    u64 b = get_challenge();
    return E2{a, b};
  }
};

}  // namespace orc
