// oracle/fixtures/fst_shim.h — TEST INFRASTRUCTURE ONLY.
//
// A ~100-line stand-in for the slice of the OpenFst C++ API that the reference's fixture builders
// (rustfst-tests-data/fst_NNN/fst_NNN.h, compiled IN PLACE from /root/reference, never copied) use:
// VectorFst<Arc>::{AddState,SetStart,SetFinal,AddArc,EmplaceArc,Read}, ConstFst<Arc>::Read,
// StdArc / ArcTpl<TropicalWeight>, TropicalWeight{One,Zero}.  It lets gen_fixtures.cpp materialise the
// fst_000..fst_020 inputs (get_fst / get_fst_compose) as OpenFst binary files under tests/golden/.
#pragma once
#include <string>
#include "../oracle.hpp"

namespace fst {

template <class T>
class TropicalWeightTpl {
 public:
  TropicalWeightTpl() : v_(0) {}
  TropicalWeightTpl(T f) : v_(f) {}  // NOLINT implicit like OpenFst
  static TropicalWeightTpl One() { return TropicalWeightTpl(0); }
  static TropicalWeightTpl Zero() { return TropicalWeightTpl(std::numeric_limits<T>::infinity()); }
  T Value() const { return v_; }
 private:
  T v_;
};
using TropicalWeight = TropicalWeightTpl<float>;

template <class W>
struct ArcTpl {
  using Weight = W;
  using Label = int;
  using StateId = int;
  ArcTpl() {}
  ArcTpl(int il, int ol, W w, int ns) : ilabel(il), olabel(ol), weight(w), nextstate(ns) {}
  int ilabel = 0, olabel = 0;
  W weight;
  int nextstate = 0;
};
using StdArc = ArcTpl<TropicalWeight>;

template <class Arc>
class ConstFst;

template <class Arc>
class VectorFst {
 public:
  using Weight = typename Arc::Weight;
  VectorFst() {}
  VectorFst(const VectorFst&) = default;
  explicit VectorFst(const ConstFst<Arc>& c);
  int AddState() { return (int)impl.add_state(); }
  void SetStart(int s) { impl.set_start((oracle::StateId)s); }
  void SetFinal(int s, Weight w) { impl.set_final((oracle::StateId)s, (float)w.Value()); }
  void AddArc(int s, const Arc& a) {
    impl.add_tr((oracle::StateId)s, oracle::Tr{(oracle::Label)a.ilabel, (oracle::Label)a.olabel,
                                              (float)a.weight.Value(), (oracle::StateId)a.nextstate});
  }
  void EmplaceArc(int s, int il, int ol, Weight w, int ns) { AddArc(s, Arc(il, ol, w, ns)); }
  static VectorFst* Read(const std::string& path) {
    auto bytes = oracle::read_file(path);
    auto* f = new VectorFst();
    f->impl = oracle::fst_from_bytes(bytes.data(), bytes.size(), false);
    return f;
  }
  oracle::Fst impl;
};

template <class Arc>
class ConstFst {
 public:
  static ConstFst* Read(const std::string& path) {
    auto bytes = oracle::read_file(path);
    auto* f = new ConstFst();
    f->impl = oracle::fst_from_bytes(bytes.data(), bytes.size(), true);
    return f;
  }
  oracle::Fst impl;
};

template <class Arc>
VectorFst<Arc>::VectorFst(const ConstFst<Arc>& c) : impl(c.impl) {}

}  // namespace fst
