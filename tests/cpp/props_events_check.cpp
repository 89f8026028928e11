// Checks that props::apply_arc_events (order-independent, from the OR of per-arc events) equals the sequential replay
// of props::on_add_tr (mutate_properties.rs:43-100) for random arc sequences and random starting property words.
#include <cstdio>
#include <random>
#include <vector>

#include "../../rustfst_b200/csrc/fst_types.h"

using namespace b200;

int main() {
  std::mt19937_64 rng(12345);
  const float weights[] = {0.0f, 0.0005f, 1.0f, 2.5f, std::numeric_limits<float>::infinity()};
  long checked = 0;
  for (int iter = 0; iter < 200000; iter++) {
    // a plausible starting word: null properties after some of the mutations that precede add_tr in practice
    uint64_t p0 = props::kNull;
    if (rng() & 1) p0 = props::on_add_state(p0);
    if (rng() & 1) { float w = weights[rng() % 5]; p0 = props::on_set_final(p0, nullptr, &w); }
    if (rng() & 1) p0 = props::on_set_start(p0);
    if ((rng() & 7) == 0) p0 = rng() & props::kTrinary;  // and sometimes an arbitrary word
    const int n_states = 1 + (int)(rng() % 4);
    uint64_t seq = p0;
    uint32_t ev = 0;
    bool any = false;
    for (int s = 0; s < n_states; s++) {
      const int na = (int)(rng() % 4);
      Tr prev{};
      bool has_prev = false;
      for (int k = 0; k < na; k++) {
        Tr tr;
        tr.ilabel = (Label)(rng() % 3);
        tr.olabel = (rng() & 1) ? tr.ilabel : (Label)(rng() % 3);
        tr.weight = weights[rng() % 5];
        tr.nextstate = (StateId)(rng() % (n_states + 1));
        seq = props::on_add_tr(seq, (StateId)s, tr, has_prev ? &prev : nullptr);
        ev |= props::arc_events((StateId)s, tr, has_prev ? &prev : nullptr);
        prev = tr; has_prev = true; any = true;
      }
    }
    const uint64_t par = props::apply_arc_events(p0, ev, any);
    if (par != seq) {
      std::printf("MISMATCH iter %d: p0 %llx seq %llx par %llx ev %x\n", iter, (unsigned long long)p0,
                  (unsigned long long)seq, (unsigned long long)par, ev);
      return 1;
    }
    checked++;
  }
  std::printf("ok %ld\n", checked);
  return 0;
}
