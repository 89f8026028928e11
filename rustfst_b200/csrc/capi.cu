// capi.cu — the extern "C" surface declared in include/rustfst_b200.h.
// Error convention and handle ownership follow rustfst-ffi/src/lib.rs:29-85 and rustfst-ffi/src/fst/mod.rs.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "../../include/rustfst_b200.h"
#include "algos.h"

using namespace b200;

// ---- opaque handle types
struct CFst {
  HostFst fst;
  bool is_const = false;  // handle made by const_fst_* (ConstFst<TropicalWeight>); everything else is a VectorFst
};
struct CTrs { std::shared_ptr<std::vector<Tr>> v; };
struct CTrsIterator { std::vector<Tr> trs; size_t index = 0; };
struct CMutTrsIterator { CFst* fst; StateId state; size_t index = 0; };
struct CStateIterator { size_t n; size_t index = 0; };
struct CMatcherConfig { uint32_t sigma_label; size_t rewrite_mode; std::vector<uint32_t> allowed; };
struct CComposeConfig { size_t filter; bool connect; bool has_m1, has_m2; CMatcherConfig m1, m2; };
struct CShortestPathConfig { float delta; size_t nshortest; bool unique; };
struct B200DeviceFst {
  Stream stream;
  DevFst d;
  B200DeviceFst() : d(stream.s) {}
};

static_assert(sizeof(CTr) == sizeof(Tr), "CTr and the device arc record must coincide");

namespace {
thread_local std::string g_last_error;
thread_local bool g_has_error = false;

template <class F>
RUSTFST_FFI_RESULT wrap(F&& f) {  // rustfst-ffi/src/lib.rs:43-55
  try {
    f();
    return RUSTFST_FFI_RESULT_OK;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    g_has_error = true;
    if (std::getenv("AMSTRAM_FFI_ERROR_STDERR")) std::fprintf(stderr, "%s\n", e.what());
    return RUSTFST_FFI_RESULT_KO;
  }
}
char* dup_cstr(const std::string& s) {
  char* p = (char*)std::malloc(s.size() + 1);
  std::memcpy(p, s.c_str(), s.size() + 1);
  return p;
}
const Tr& as_tr(const CTr* t) { return *reinterpret_cast<const Tr*>(t); }
CTr* new_ctr(const Tr& t) {
  CTr* c = new CTr;
  std::memcpy(c, &t, sizeof(Tr));
  return c;
}
template <class T>
T* nn(T* p, const char* what) {  // ffi_convert raw_borrow on a null pointer is an error, not UB
  if (!p) throw FstError(std::string("unexpected null pointer: ") + what);
  return p;
}
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// Downcasts of the opaque handle (rustfst-ffi/src/fst/mod.rs:99-111 as_fst!, algorithms/compose.rs:315-321)
template <class T>
T* vec_alg(T* p, const char* what) {  // algorithms: "Could not downcast to vector FST"
  if (nn(p, what)->is_const) throw FstError("Could not downcast to vector FST");
  return p;
}
template <class T>
T* vec_h(T* p, const char* what) {  // vec_fst_* accessors
  if (nn(p, what)->is_const) throw FstError("Could not downcast to VectorFst<TropicalWeight> FST");
  return p;
}
template <class T>
T* const_h(T* p, const char* what) {  // const_fst_* accessors
  if (!nn(p, what)->is_const) throw FstError("Could not downcast to ConstFst<TropicalWeight> FST");
  return p;
}

ComposeOptions to_options(const CComposeConfig* c) {
  ComposeOptions o;
  if (!c) return o;
  if (c->filter > 6) throw FstError("EnumConversionError");
  o.filter = (int)c->filter;
  o.connect = c->connect;
  auto conv = [](const CMatcherConfig& m, SigmaSpec& sp) {
    if (m.rewrite_mode > 2) throw FstError("EnumConversionError");
    sp.enabled = true; sp.sigma_label = m.sigma_label; sp.rewrite_mode = (int)m.rewrite_mode; sp.allowed = m.allowed;
    std::sort(sp.allowed.begin(), sp.allowed.end());
  };
  if (c->has_m1) conv(c->m1, o.sigma1);
  if (c->has_m2) conv(c->m2, o.sigma2);
  if ((c->has_m1 || c->has_m2) && o.filter == kAutoFilter)  // compose_static.rs:219-223
    throw FstError("Custom MatcherConfig not supported with AutoFilter");
  return o;
}
void fill(B200ComposeStats* out, const ComposeStats& st, float h2d, float d2h) {
  if (!out) return;
  out->states_expanded = st.states_expanded; out->arcs_iterated = st.arcs_iterated;
  out->arcs_emitted = st.arcs_emitted; out->waves = st.waves; out->states_out = st.states_out;
  out->arcs_out = st.arcs_out; out->kernel_launches = st.kernel_launches; out->emit_launches = st.emit_launches;
  out->ms_expand = st.ms_expand; out->ms_connect = st.ms_connect; out->ms_emit_kernel = st.ms_emit_kernel;
  out->ms_h2d = h2d; out->ms_d2h = d2h;
  out->ms_phase_match = st.ms_phase[0]; out->ms_phase_emit = st.ms_phase[1];
  out->ms_phase_rank = st.ms_phase[2]; out->ms_phase_resolve = st.ms_phase[3];
}
void fill(B200SsspStats* out, const SsspStats& st, int kind, float h2d) {
  if (!out) return;
  out->arcs_relaxed = st.arcs_relaxed; out->states_settled = st.states_settled; out->waves = st.waves;
  out->kernel_launches = st.kernel_launches; out->relax_launches = st.relax_launches; out->path = st.path;
  out->queue_kind = kind; out->ms_device = st.ms_device; out->ms_relax_kernel = st.ms_relax_kernel;
  out->ms_h2d = h2d; out->ms_queue_plan_host = st.plan_host_ms;
  out->ms_order_device = st.order_device_ms; out->order_on_device = st.order_on_device ? 1 : 0;
  out->sweep = st.sweep ? 1 : 0;
}

// Queue plan (AutoQueue::new, auto_queue.rs:23-99) of a machine that is resident on the device.  The discipline follows
// from the property word; the DFS order of a machine known to be ACYCLIC is computed on the device (dag_order.cu).
// `host` (may be null for device-only callers) is the same machine on the host: it is needed for machines that are not
// known to be acyclic (Tarjan SCC numbering) and as the fallback when the device path declines (too deep).
struct DevicePlan {
  QueuePlan plan;
  DevBuf<uint32_t> order;
};
void resolve_plan(DevicePlan& dp, const CsrFst* host, const DevFst& d, cudaStream_t s) {
  const QueueKind kind = queue_kind_from_props(d.props, d.has_start);
  const bool try_device = kind == kTopOrderQueue && !std::getenv("B200_HOST_DFS");
  if (kind == kStateOrderQueue || kind == kLifoQueue) { dp.plan.kind = kind; return; }
  if (try_device) {
    dp.order = DevBuf<uint32_t>(s);
    float ms = 0;
    if (dag_top_order_device(d, dp.order, &ms, nullptr, s)) {
      dp.plan.kind = kTopOrderQueue; dp.plan.d_order = dp.order.p; dp.plan.device_ms = ms;
      return;
    }
  }
  if (!host) throw FstError("shortest path on a device-resident FST that is not known to be top-sorted, acyclic or "
                            "unweighted needs the host copy (plan_from) for the DFS order");
  dp.plan = build_queue_plan(*host);
}

CFst* compose_host(const CFst* a, const CFst* b, const CComposeConfig* cfg, B200ComposeStats* stats) {
  ComposeOptions opt = to_options(cfg);
  const CsrFst& ha = vec_alg(a, "fst_1")->fst.checked();
  const CsrFst& hb = vec_alg(b, "fst_2")->fst.checked();
  Stream st;
  double t0 = now_ms();
  DevFst da = upload(ha, st.s);
  DevFst db = upload(hb, st.s);
  double t1 = now_ms();
  ComposeStats cs;
  DevFst dr = compose_device(da, db, opt, &cs, st.s);
  double t2 = now_ms();
  CsrFst hr = download(dr, st.s);
  double t3 = now_ms();
  fill(stats, cs, (float)(t1 - t0), (float)(t3 - t2));
  return new CFst{HostFst(std::move(hr))};
}

// shortest_path_with_config (shortest_path.rs:107-171) on a device-resident machine: nshortest == 1 -> single
// shortest path; > 1 -> distances + reversed machine (determinized on demand when unique) + n-best search.
void check_sp_config(const CShortestPathConfig*) {}  // every combination is served (unique: nshortest.cu)
CsrFst shortest_path_dispatch(const DevFst& d, const QueuePlan& plan, const CShortestPathConfig* cfg,
                              B200SsspStats* stats, float ms_h2d, cudaStream_t s, bool force_serial) {
  const size_t nshortest = cfg ? cfg->nshortest : 1;  // shortest_path.rs:30-38 defaults
  if (nshortest != 1) {                                // shortest_path.rs:135-170
    NShortestStats ns;
    CsrFst r = n_shortest_paths_device(d, plan, nshortest, cfg->delta, &ns, s, force_serial, cfg->unique);
    ns.distance.ms_device = ns.ms_total;
    fill(stats, ns.distance, (int)plan.kind, ms_h2d);
    return r;
  }
  SsspStats ss;
  CsrFst r = shortest_path_device(d, plan, &ss, s, force_serial);
  fill(stats, ss, (int)plan.kind, ms_h2d);
  return r;
}

CFst* shortest_path_host(const CFst* in, const CShortestPathConfig* cfg, B200SsspStats* stats, bool force_serial) {
  if (cfg && cfg->nshortest == 0) {  // shortest_path.rs:120-122
    if (stats) std::memset(stats, 0, sizeof(*stats));
    return new CFst{};
  }
  check_sp_config(cfg);
  const CsrFst& h = vec_alg(in, "fst")->fst.checked();
  Stream st;
  double t0 = now_ms();
  DevFst d = upload(h, st.s);
  double t1 = now_ms();
  DevicePlan dp;
  resolve_plan(dp, &h, d, st.s);
  return new CFst{HostFst(shortest_path_dispatch(d, dp.plan, cfg, stats, (float)(t1 - t0), st.s, force_serial))};
}
}  // namespace

extern "C" {

const char* b200_version(void) { return "rustfst_b200 0.1.0 (sm_100a)"; }

RUSTFST_FFI_RESULT rustfst_ffi_get_last_error(char** error) {
  return wrap([&] {
    std::string msg = g_has_error ? g_last_error : std::string("No error message");
    g_has_error = false;
    g_last_error.clear();
    *error = dup_cstr(msg);
  });
}
RUSTFST_FFI_RESULT rustfst_destroy_string(char* string) {
  return wrap([&] { std::free(string); });
}

// ---------------------------------------------------------------- hot path
RUSTFST_FFI_RESULT fst_compose(const CFst* a, const CFst* b, const CFst** out) {
  return wrap([&] { *out = compose_host(a, b, nullptr, nullptr); });
}
RUSTFST_FFI_RESULT fst_compose_with_config(const CFst* a, const CFst* b, const CComposeConfig* cfg, const CFst** out) {
  return wrap([&] { *out = compose_host(a, b, nn(cfg, "config"), nullptr); });
}
RUSTFST_FFI_RESULT b200_compose_with_stats(const CFst* a, const CFst* b, const CComposeConfig* cfg, const CFst** out,
                                           B200ComposeStats* stats) {
  return wrap([&] { *out = compose_host(a, b, cfg, stats); });
}
RUSTFST_FFI_RESULT fst_compose_config_new(size_t filter, bool connect, const CMatcherConfig* m1,
                                          const CMatcherConfig* m2, const CComposeConfig** config) {
  return wrap([&] {
    auto* c = new CComposeConfig{filter, connect, m1 != nullptr, m2 != nullptr, {}, {}};
    if (m1) c->m1 = *m1;
    if (m2) c->m2 = *m2;
    *config = c;
  });
}
RUSTFST_FFI_RESULT fst_compose_config_destroy(CComposeConfig* p) { return wrap([&] { delete p; }); }
RUSTFST_FFI_RESULT fst_matcher_config_new(size_t sigma_label, size_t rewrite_mode, CIntArray allowed,
                                          const CMatcherConfig** config) {
  return wrap([&] {
    auto* c = new CMatcherConfig{(uint32_t)sigma_label, rewrite_mode, {}};
    if (allowed.data && allowed.size) c->allowed.assign(allowed.data, allowed.data + allowed.size);
    *config = c;
  });
}
RUSTFST_FFI_RESULT fst_matcher_config_destroy(CMatcherConfig* p) { return wrap([&] { delete p; }); }

RUSTFST_FFI_RESULT fst_shortest_path(const CFst* in, const CFst** out) {
  return wrap([&] { *out = shortest_path_host(in, nullptr, nullptr, false); });
}
RUSTFST_FFI_RESULT fst_shortest_path_with_config(const CFst* in, const CShortestPathConfig* cfg, const CFst** out) {
  return wrap([&] { *out = shortest_path_host(in, nn(cfg, "config"), nullptr, false); });
}
RUSTFST_FFI_RESULT b200_shortest_path_with_stats(const CFst* in, const CShortestPathConfig* cfg, const CFst** out,
                                                 B200SsspStats* stats, bool force_serial) {
  return wrap([&] { *out = shortest_path_host(in, cfg, stats, force_serial); });
}
RUSTFST_FFI_RESULT fst_shortest_path_config_new(float delta, size_t nshortest, bool unique,
                                                const CShortestPathConfig** ptr) {
  return wrap([&] { *ptr = new CShortestPathConfig{delta, nshortest, unique}; });
}
RUSTFST_FFI_RESULT b200_shortest_path_config_destroy(CShortestPathConfig* p) { return wrap([&] { delete p; }); }

RUSTFST_FFI_RESULT fst_connect(CFst* ptr) {
  return wrap([&] {
    const CsrFst& h = vec_alg(ptr, "fst")->fst.checked();
    Stream st;
    DevFst d = upload(h, st.s);
    DevFst r = connect_device(d, false, nullptr, st.s);
    ptr->fst.replace(download(r, st.s));
  });
}
RUSTFST_FFI_RESULT fst_reverse(const CFst* ptr, const CFst** res_ptr) {
  return wrap([&] {
    const CsrFst& h = vec_alg(ptr, "fst")->fst.checked();
    Stream st;
    DevFst d = upload(h, st.s);
    *res_ptr = new CFst{HostFst(reverse_fst_device(d, st.s))};
  });
}
RUSTFST_FFI_RESULT fst_isomorphic(const CFst* fst, const CFst* other_fst, size_t* is_isomorphic) {
  return wrap([&] {
    const CsrFst& ha = vec_alg(fst, "fst")->fst.checked();
    const CsrFst& hb = vec_alg(other_fst, "other_fst")->fst.checked();
    // Large machines (a composed lattice being verified) are paired on the device; the sequential restatement handles
    // small ones, machines with Some(+inf) final weights, boxes without a GPU, and the cases the device leaves undecided
    // (the reference's non-determinism error depends on its visiting order).  B200_ISO_DEVICE=1 forces the device path.
    int ndev = 0;
    const bool big = ha.arcs.size() + hb.arcs.size() >= (1u << 16) || std::getenv("B200_ISO_DEVICE") != nullptr;
    if (big && ha.inf_finals.empty() && hb.inf_finals.empty() && cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
      Stream st;
      DevFst da = upload(ha, st.s), db = upload(hb, st.s);
      const int r = isomorphic_device(da, db, kDelta, st.s);
      if (r >= 0) { *is_isomorphic = (size_t)r; return; }
    }
    cudaGetLastError();
    *is_isomorphic = isomorphic(ha, hb) ? 1 : 0;
  });
}
RUSTFST_FFI_RESULT b200_device_isomorphic(const B200DeviceFst* a, const B200DeviceFst* b, int32_t* result) {
  return wrap([&] { *result = isomorphic_device(nn(a, "fst_1")->d, nn(b, "fst_2")->d, kDelta, a->stream.s); });
}
RUSTFST_FFI_RESULT fst_top_sort(CFst* ptr) {
  return wrap([&] {  // top_sort.rs:75-95: the DFS is the reference's sequential one, the renumbering one pass over the CSR
    HostFst& f = vec_alg(ptr, "fst")->fst;
    std::vector<uint32_t> order;
    if (top_order(f.checked(), order)) {
      f.state_sort(order);
      f.or_properties(props::kAcyclic | props::kInitialAcyclic | props::kTopSorted);
    } else {
      f.or_properties(props::kCyclic | props::kNotTopSorted);
    }
  });
}
RUSTFST_FFI_RESULT fst_tr_sort(CFst* ptr, bool ilabel_comp) {
  return wrap([&] {
    vec_alg(ptr, "fst");
    // Large machines are sorted on the device (one radix sort + gather); small ones, and any machine on a box
    // without a GPU, by the host container (same stable order either way; tr_sort is not part of the hot path).
    int ndev = 0;
    const CsrFst& h = ptr->fst.checked();
    if (h.arcs.size() >= (1u << 16) && h.inf_finals.empty() && cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
      Stream st;
      DevFst d = upload(h, st.s);
      tr_sort_device(d, ilabel_comp, st.s);
      ptr->fst.replace(download(d, st.s));
      return;
    }
    cudaGetLastError();
    ptr->fst.tr_sort(ilabel_comp);
  });
}

// ---------------------------------------------------------------- Fst accessors
RUSTFST_FFI_RESULT fst_start(const CFst* fst, CStateId* state) {
  return wrap([&] { StateId s; if (nn(fst, "fst")->fst.start(&s)) *state = s; });
}
RUSTFST_FFI_RESULT fst_final_weight(const CFst* fst, CStateId s, float* w) {
  return wrap([&] { float v; if (nn(fst, "fst")->fst.final_weight(s, &v)) *w = v; });
}
RUSTFST_FFI_RESULT fst_num_trs(const CFst* fst, CStateId s, size_t* n) {
  return wrap([&] { *n = nn(fst, "fst")->fst.num_trs(s); });
}
RUSTFST_FFI_RESULT fst_get_trs(const CFst* fst, CStateId s, const CTrs** trs) {
  return wrap([&] { *trs = new CTrs{std::make_shared<std::vector<Tr>>(nn(fst, "fst")->fst.get_trs(s))}; });
}
RUSTFST_FFI_RESULT fst_is_final(const CFst* fst, CStateId s, size_t* is_final) {
  return wrap([&] { float v; *is_final = nn(fst, "fst")->fst.final_weight(s, &v) ? 1 : 0; });
}
RUSTFST_FFI_RESULT fst_is_start(const CFst* fst, CStateId s, size_t* is_start) {
  return wrap([&] { StateId st; *is_start = (nn(fst, "fst")->fst.start(&st) && st == s) ? 1 : 0; });
}
// Symbol tables are outside the hot path: handles never carry one, so the optional out slot is left untouched.
RUSTFST_FFI_RESULT fst_input_symbols(const CFst* fst, const CSymbolTable**) { return wrap([&] { nn(fst, "fst"); }); }
RUSTFST_FFI_RESULT fst_output_symbols(const CFst* fst, const CSymbolTable**) { return wrap([&] { nn(fst, "fst"); }); }
RUSTFST_FFI_RESULT fst_weight_one(float* w) { return wrap([&] { *w = 0.0f; }); }
RUSTFST_FFI_RESULT fst_weight_zero(float* w) { return wrap([&] { *w = w_zero(); }); }
RUSTFST_FFI_RESULT fst_destroy(CFst* p) { return wrap([&] { delete p; }); }

// ---------------------------------------------------------------- VectorFst
RUSTFST_FFI_RESULT vec_fst_new(const CFst** ptr) { return wrap([&] { *ptr = new CFst{}; }); }
RUSTFST_FFI_RESULT vec_fst_set_start(CFst* f, CStateId s) { return wrap([&] { vec_h(f, "fst")->fst.set_start(s); }); }
RUSTFST_FFI_RESULT vec_fst_set_final(CFst* f, CStateId s, float w) {
  return wrap([&] { vec_h(f, "fst")->fst.set_final(s, w); });
}
RUSTFST_FFI_RESULT vec_fst_add_state(CFst* f, CStateId* s) { return wrap([&] { *s = vec_h(f, "fst")->fst.add_state(); }); }
RUSTFST_FFI_RESULT vec_fst_delete_states(CFst* f) { return wrap([&] { vec_h(f, "fst")->fst.del_all_states(); }); }
RUSTFST_FFI_RESULT vec_fst_add_tr(CFst* f, CStateId s, const CTr* tr) {
  return wrap([&] { vec_h(f, "fst")->fst.add_tr(s, as_tr(nn(tr, "tr"))); });
}
RUSTFST_FFI_RESULT vec_fst_del_final_weight(CFst* f, CStateId s) {
  return wrap([&] { vec_h(f, "fst")->fst.delete_final_weight(s); });
}
RUSTFST_FFI_RESULT vec_fst_from_path(const CFst** ptr, const char* path) {
  return wrap([&] {
    auto bytes = io::read_file(nn(path, "path"));
    *ptr = new CFst{HostFst(io::parse_vector_fst(bytes.data(), bytes.size()))};
  });
}
RUSTFST_FFI_RESULT vec_fst_write_file(const CFst* f, const char* path) {
  return wrap([&] {
    const CsrFst& c = vec_h(f, "fst")->fst.freeze();
    const size_t size = io::vector_fst_bytes(c);
    std::unique_ptr<uint8_t, decltype(&std::free)> buf((uint8_t*)std::malloc(size ? size : 1), &std::free);
    if (!buf) throw std::bad_alloc();
    io::store_vector_fst_into(c, buf.get());
    io::write_file(nn(path, "path"), buf.get(), size);
  });
}
RUSTFST_FFI_RESULT vec_fst_num_states(const CFst* f, size_t* n) {
  return wrap([&] { *n = vec_h(f, "fst")->fst.num_states(); });
}
RUSTFST_FFI_RESULT vec_fst_equals(const CFst* a, const CFst* b, size_t* eq) {
  return wrap([&] { *eq = vec_h(a, "fst")->fst.equals(vec_h(b, "other_fst")->fst) ? 1 : 0; });
}
RUSTFST_FFI_RESULT vec_fst_copy(const CFst* f, const CFst** clone) {
  return wrap([&] { *clone = new CFst{HostFst(vec_h(f, "fst")->fst)}; });
}
RUSTFST_FFI_RESULT vec_fst_display(const CFst* f, const char** s) {
  return wrap([&] { *s = dup_cstr(vec_h(f, "fst")->fst.display()); });
}
RUSTFST_FFI_RESULT vec_fst_to_bytes(const CFst* f, const CArrayU8** out) {
  return wrap([&] {
    const CsrFst& c = vec_h(f, "fst")->fst.freeze();
    const size_t size = io::vector_fst_bytes(c);
    uint8_t* p = (uint8_t*)std::malloc(size ? size : 1);
    if (!p) throw std::bad_alloc();
    io::store_vector_fst_into(c, p);
    *out = new CArrayU8{p, size};
  });
}
RUSTFST_FFI_RESULT b200_bytes_destroy(CArrayU8* b) {
  return wrap([&] { if (b) { std::free((void*)b->data_ptr); delete b; } });
}
RUSTFST_FFI_RESULT vec_fst_from_bytes(const CArrayU8* bytes, const CFst** ptr) {
  return wrap([&] {
    nn(bytes, "bytes");
    *ptr = new CFst{HostFst(io::parse_vector_fst(bytes->data_ptr, bytes->size))};
  });
}

// ---------------------------------------------------------------- ConstFst (rustfst-ffi/src/fst/const_fst.rs)
RUSTFST_FFI_RESULT const_fst_from_path(const CFst** ptr, const char* path) {
  return wrap([&] {
    auto bytes = io::read_file(nn(path, "path"));
    *ptr = new CFst{HostFst(io::parse_const_fst(bytes.data(), bytes.size())), true};
  });
}
RUSTFST_FFI_RESULT const_fst_write_file(const CFst* f, const char* path) {
  return wrap([&] { io::write_file(nn(path, "path"), io::store_const_fst(const_h(f, "fst")->fst.freeze())); });
}
RUSTFST_FFI_RESULT const_fst_equals(const CFst* a, const CFst* b, size_t* eq) {
  return wrap([&] { *eq = const_h(a, "fst")->fst.equals(const_h(b, "other_fst")->fst) ? 1 : 0; });
}
RUSTFST_FFI_RESULT const_fst_copy(const CFst* f, const CFst** clone) {
  return wrap([&] { *clone = new CFst{HostFst(const_h(f, "fst")->fst), true}; });
}
RUSTFST_FFI_RESULT const_fst_from_vec_fst(const CFst* vec_fst, const CFst** const_fst) {
  return wrap([&] {  // ConstFst::from(vec_fst.clone()): all properties are computed once and frozen (converters.rs:7-37)
    auto c = std::make_unique<CFst>(CFst{HostFst(vec_h(vec_fst, "vec_fst")->fst), true});
    c->fst.compute_and_update_properties_all();
    *const_fst = c.release();
  });
}
RUSTFST_FFI_RESULT const_fst_display(const CFst* f, const char** s) {
  return wrap([&] { *s = dup_cstr(const_h(f, "fst")->fst.display()); });
}

// ---------------------------------------------------------------- Tr
RUSTFST_FFI_RESULT tr_new(CLabel il, CLabel ol, float w, CStateId ns, const CTr** out) {
  return wrap([&] { *out = new CTr{il, ol, w, ns}; });
}
RUSTFST_FFI_RESULT tr_ilabel(const CTr* t, CLabel* v) { return wrap([&] { *v = nn(t, "tr")->ilabel; }); }
RUSTFST_FFI_RESULT tr_set_ilabel(CTr* t, size_t v) { return wrap([&] { nn(t, "tr")->ilabel = (CLabel)v; }); }
RUSTFST_FFI_RESULT tr_olabel(const CTr* t, CLabel* v) { return wrap([&] { *v = nn(t, "tr")->olabel; }); }
RUSTFST_FFI_RESULT tr_set_olabel(CTr* t, size_t v) { return wrap([&] { nn(t, "tr")->olabel = (CLabel)v; }); }
RUSTFST_FFI_RESULT tr_weight(const CTr* t, float* v) { return wrap([&] { *v = nn(t, "tr")->weight; }); }
RUSTFST_FFI_RESULT tr_set_weight(CTr* t, float v) { return wrap([&] { nn(t, "tr")->weight = v; }); }
RUSTFST_FFI_RESULT tr_next_state(const CTr* t, CStateId* v) { return wrap([&] { *v = nn(t, "tr")->nextstate; }); }
RUSTFST_FFI_RESULT tr_set_next_state(CTr* t, size_t v) { return wrap([&] { nn(t, "tr")->nextstate = (CStateId)v; }); }
RUSTFST_FFI_RESULT tr_delete(CTr* t) { return wrap([&] { delete t; }); }

// ---------------------------------------------------------------- Trs
RUSTFST_FFI_RESULT trs_vec_new(const CTrs** out) {
  return wrap([&] { *out = new CTrs{std::make_shared<std::vector<Tr>>()}; });
}
RUSTFST_FFI_RESULT trs_vec_remove(CTrs* trs, size_t index, const CTr** removed) {
  return wrap([&] {
    auto& v = *nn(trs, "trs")->v;
    if (index >= v.size()) throw FstError("removal index (is " + std::to_string(index) + ") should be < len (is " +
                                          std::to_string(v.size()) + ")");
    // TrsVec::remove goes through Arc::make_mut: copy-on-write when shared (rustfst/src/trs.rs)
    if (trs->v.use_count() > 1) trs->v = std::make_shared<std::vector<Tr>>(*trs->v);
    Tr t = (*trs->v)[index];
    trs->v->erase(trs->v->begin() + (long)index);
    *removed = new_ctr(t);
  });
}
RUSTFST_FFI_RESULT trs_vec_push(CTrs* trs, const CTr* tr) {
  return wrap([&] {
    nn(trs, "trs");
    if (trs->v.use_count() > 1) trs->v = std::make_shared<std::vector<Tr>>(*trs->v);
    trs->v->push_back(as_tr(nn(tr, "tr")));
  });
}
RUSTFST_FFI_RESULT trs_vec_shallow_clone(const CTrs* trs, const CTrs** out) {
  return wrap([&] { *out = new CTrs{nn(trs, "trs")->v}; });
}
RUSTFST_FFI_RESULT trs_vec_len(const CTrs* trs, size_t* n) { return wrap([&] { *n = nn(trs, "trs")->v->size(); }); }
RUSTFST_FFI_RESULT trs_vec_display(const CTrs* trs, const char** out) {
  return wrap([&] {
    // format!("{:?}", TrsVec<TropicalWeight>) (rustfst-ffi/src/trs.rs:100): the derived Debug of TrsVec(Arc<Vec<Tr>>),
    // Tr, TropicalWeight { value: OrderedFloat<f32> } — floats in Rust's shortest round-trip form
    auto fmt_f32 = [](float w) {
      if (std::isinf(w)) return std::string(w > 0 ? "inf" : "-inf");
      if (std::isnan(w)) return std::string("NaN");
      char buf[64];
      auto r = std::to_chars(buf, buf + sizeof(buf), w);  // shortest representation that round-trips
      std::string t(buf, r.ptr);
      if (t.find('e') != std::string::npos) {  // Rust's {:?} switches to exponent form at other thresholds; fixed is exact here
        r = std::to_chars(buf, buf + sizeof(buf), w, std::chars_format::fixed);
        t.assign(buf, r.ptr);
      }
      if (t.find('.') == std::string::npos) t += ".0";
      return t;
    };
    std::string s = "TrsVec([";
    bool first = true;
    for (const Tr& t : *nn(trs, "trs")->v) {
      if (!first) s += ", ";
      first = false;
      s += "Tr { ilabel: " + std::to_string(t.ilabel) + ", olabel: " + std::to_string(t.olabel) +
           ", weight: TropicalWeight { value: OrderedFloat(" + fmt_f32(t.weight) + ") }, nextstate: " +
           std::to_string(t.nextstate) + " }";
    }
    s += "])";
    *out = dup_cstr(s);
  });
}
RUSTFST_FFI_RESULT trs_vec_delete(CTrs* p) { return wrap([&] { delete p; }); }

// ---------------------------------------------------------------- iterators
RUSTFST_FFI_RESULT trs_iterator_new(CFst* fst, CStateId s, const CTrsIterator** out) {
  return wrap([&] {
    nn(fst, "fst");
    // an unknown state leaves the caller's slot untouched and still returns OK (iterators.rs:49-62)
    if (s < fst->fst.num_states()) *out = new CTrsIterator{fst->fst.get_trs(s), 0};
  });
}
RUSTFST_FFI_RESULT trs_iterator_next(CTrsIterator* it, const CTr** out) {
  return wrap([&] {
    nn(it, "iter");
    if (it->index < it->trs.size()) *out = new_ctr(it->trs[it->index]);
    it->index++;
  });
}
RUSTFST_FFI_RESULT trs_iterator_done(const CTrsIterator* it, size_t* done) {
  return wrap([&] { *done = nn(it, "iter")->trs.size() == it->index ? 1 : 0; });
}
RUSTFST_FFI_RESULT trs_iterator_reset(CTrsIterator* it) { return wrap([&] { nn(it, "iter")->index = 0; }); }
RUSTFST_FFI_RESULT trs_iterator_destroy(CTrsIterator* it) { return wrap([&] { delete it; }); }

RUSTFST_FFI_RESULT mut_trs_iterator_new(CFst* fst, CStateId s, const CMutTrsIterator** out) {
  return wrap([&] {
    nn(fst, "fst")->fst.num_trs(s);  // validates the state
    *out = new CMutTrsIterator{fst, s, 0};
  });
}
RUSTFST_FFI_RESULT mut_trs_iterator_next(CMutTrsIterator* it) { return wrap([&] { nn(it, "iter")->index++; }); }
RUSTFST_FFI_RESULT mut_trs_iterator_value(CMutTrsIterator* it, const CTr** out) {
  return wrap([&] {
    nn(it, "iter");
    auto trs = it->fst->fst.get_trs(it->state);
    if (it->index < trs.size()) *out = new_ctr(trs[it->index]);
  });
}
RUSTFST_FFI_RESULT mut_trs_iterator_set_value(CMutTrsIterator* it, const CTr* tr) {
  return wrap([&] { nn(it, "iter")->fst->fst.set_tr(it->state, it->index, as_tr(nn(tr, "tr"))); });
}
RUSTFST_FFI_RESULT mut_trs_iterator_done(const CMutTrsIterator* it, size_t* done) {
  return wrap([&] { *done = nn(it, "iter")->index >= it->fst->fst.num_trs(it->state) ? 1 : 0; });
}
RUSTFST_FFI_RESULT mut_trs_iterator_reset(CMutTrsIterator* it) { return wrap([&] { nn(it, "iter")->index = 0; }); }
RUSTFST_FFI_RESULT mut_trs_iterator_destroy(CMutTrsIterator* it) { return wrap([&] { delete it; }); }

RUSTFST_FFI_RESULT state_iterator_new(CFst* fst, const CStateIterator** out) {
  return wrap([&] { *out = new CStateIterator{nn(fst, "fst")->fst.num_states(), 0}; });
}
RUSTFST_FFI_RESULT state_iterator_next(CStateIterator* it, CStateId* state) {
  return wrap([&] {
    nn(it, "iter");
    if (it->index < it->n) *state = (CStateId)it->index;
    it->index++;
  });
}
RUSTFST_FFI_RESULT state_iterator_done(CStateIterator* it, size_t* done) {
  return wrap([&] { *done = nn(it, "iter")->index >= it->n ? 1 : 0; });
}
RUSTFST_FFI_RESULT state_iterator_destroy(CStateIterator* it) { return wrap([&] { delete it; }); }

// ---------------------------------------------------------------- b200_ additions
RUSTFST_FFI_RESULT b200_fst_properties(const CFst* f, uint64_t* p) {
  return wrap([&] { *p = nn(f, "fst")->fst.properties(); });
}
RUSTFST_FFI_RESULT b200_fst_set_properties(CFst* f, uint64_t p) {
  return wrap([&] { nn(f, "fst")->fst.set_properties(p); });
}
RUSTFST_FFI_RESULT b200_fst_from_csr(uint64_t n, const uint32_t* offsets, const CTr* arcs, const float* finals,
                                     int64_t start, uint64_t props_word, const CFst** out) {
  return wrap([&] {
    if (n >= 0x7FFFFFFFull) throw FstError("too many states");
    for (uint64_t s = 0; s < n; s++)
      if (offsets[s + 1] < offsets[s]) throw FstError("b200_fst_from_csr: offsets must be non-decreasing");
    CsrFst c;
    c.offsets.assign(offsets, offsets + n + 1);
    size_t a = c.offsets[n];
    c.arcs.resize(a);
    if (a) std::memcpy(c.arcs.data(), arcs, a * sizeof(Tr));
    c.finals.assign(finals, finals + n);
    c.has_start = start >= 0;
    c.start = (StateId)start;
    if (c.has_start && (uint64_t)start >= n) throw FstError("The state " + std::to_string(start) + " doesn't exist");
    c.props = props_word & props::kTrinary;
    validate_state_ids(c, "b200_fst_from_csr: a transition points to a state that does not exist");
    *out = new CFst{HostFst(std::move(c))};
  });
}
RUSTFST_FFI_RESULT b200_fst_compute_properties(CFst* f, uint64_t* p) {
  return wrap([&] {
    const uint64_t v = nn(f, "fst")->fst.compute_and_update_properties_all();
    if (p) *p = v;
  });
}
RUSTFST_FFI_RESULT b200_fst_num_states(const CFst* f, uint64_t* n) {
  return wrap([&] { *n = nn(f, "fst")->fst.num_states(); });
}
RUSTFST_FFI_RESULT b200_fst_num_trs_total(const CFst* f, uint64_t* n) {
  return wrap([&] { *n = nn(f, "fst")->fst.freeze().arcs.size(); });
}
RUSTFST_FFI_RESULT b200_fst_to_csr(const CFst* f, uint32_t* offsets, CTr* arcs, float* finals, int64_t* start) {
  return wrap([&] {
    const CsrFst& c = nn(f, "fst")->fst.freeze();
    if (offsets) std::memcpy(offsets, c.offsets.data(), c.offsets.size() * 4);
    if (arcs && !c.arcs.empty()) std::memcpy(arcs, c.arcs.data(), c.arcs.size() * sizeof(Tr));
    if (finals && !c.finals.empty()) std::memcpy(finals, c.finals.data(), c.finals.size() * 4);
    if (start) *start = c.has_start ? (int64_t)c.start : -1;
  });
}

RUSTFST_FFI_RESULT b200_device_fst_upload(const CFst* f, const B200DeviceFst** out) {
  return wrap([&] {
    const CsrFst& h = nn(f, "fst")->fst.checked();
    auto d = std::make_unique<B200DeviceFst>();
    d->d = upload(h, d->stream.s);
    *out = d.release();
  });
}
RUSTFST_FFI_RESULT b200_device_fst_download(const B200DeviceFst* d, const CFst** out) {
  return wrap([&] { *out = new CFst{HostFst(download(nn(d, "dfst")->d, d->stream.s))}; });
}
RUSTFST_FFI_RESULT b200_device_fst_info(const B200DeviceFst* d, uint64_t* n, uint64_t* a, uint64_t* p) {
  return wrap([&] {
    nn(d, "dfst");
    if (n) *n = d->d.num_states;
    if (a) *a = d->d.num_arcs;
    if (p) *p = d->d.props;
  });
}
RUSTFST_FFI_RESULT b200_device_fst_destroy(B200DeviceFst* d) {
  return wrap([&] {
    if (!d) return;
    cudaStreamSynchronize(d->stream.s);
    delete d;
  });
}
RUSTFST_FFI_RESULT b200_device_compose(const B200DeviceFst* a, const B200DeviceFst* b, const CComposeConfig* cfg,
                                       const B200DeviceFst** out, B200ComposeStats* stats) {
  return wrap([&] {
    ComposeOptions opt = to_options(cfg);
    auto r = std::make_unique<B200DeviceFst>();
    ComposeStats cs;
    r->d = compose_device(nn(a, "fst_1")->d, nn(b, "fst_2")->d, opt, &cs, r->stream.s);
    fill(stats, cs, 0.0f, 0.0f);
    *out = r.release();
  });
}
RUSTFST_FFI_RESULT b200_device_shortest_path(const B200DeviceFst* d, const CFst* plan_from, const CFst** out,
                                             B200SsspStats* stats, bool force_serial) {
  return wrap([&] {
    DevicePlan dp;
    resolve_plan(dp, plan_from ? &plan_from->fst.checked() : nullptr, nn(d, "dfst")->d, d->stream.s);
    SsspStats ss;
    CsrFst r = shortest_path_device(d->d, dp.plan, &ss, d->stream.s, force_serial);
    fill(stats, ss, (int)dp.plan.kind, 0.0f);
    *out = new CFst{HostFst(std::move(r))};
  });
}
RUSTFST_FFI_RESULT b200_device_shortest_path_with_config(const B200DeviceFst* d, const CFst* plan_from,
                                                         const CShortestPathConfig* cfg, const CFst** out,
                                                         B200SsspStats* stats, bool force_serial) {
  return wrap([&] {
    if (cfg && cfg->nshortest == 0) {
      if (stats) std::memset(stats, 0, sizeof(*stats));
      *out = new CFst{};
      return;
    }
    check_sp_config(cfg);
    DevicePlan dp;
    resolve_plan(dp, plan_from ? &plan_from->fst.checked() : nullptr, nn(d, "dfst")->d, d->stream.s);
    *out = new CFst{HostFst(shortest_path_dispatch(d->d, dp.plan, cfg, stats, 0.0f, d->stream.s, force_serial))};
  });
}
// Core of the batched mode: acceptors[i] o T for i in [0, n) -> one PackedBatch.  T comes either as a host handle
// (uploaded here) or as a device-resident handle (the shared transducer of a long-running service stays in HBM).
static void compose_batch_core(const CFst* const* acceptors, size_t n, const CFst* transducer,
                               const B200DeviceFst* dev_transducer, const CComposeConfig* cfg, PackedBatch& out,
                               B200ComposeStats* total) {
  ComposeOptions opt = to_options(cfg);
  B200ComposeStats acc;
  std::memset(&acc, 0, sizeof(acc));
  Stream st;
  double t0 = now_ms();
  DevFst dt_local(st.s);
  const DevFst* dtp = nullptr;
  if (dev_transducer) dtp = &dev_transducer->d;
  else {
    const CsrFst& ht = vec_alg(transducer, "transducer")->fst.checked();
    dt_local = upload(ht, st.s);
    dtp = &dt_local;
  }
  const DevFst& dt = *dtp;
  acc.ms_h2d += (float)(now_ms() - t0);

  // ---- one BFS for the whole batch: the acceptors become one FST with disjoint state ranges and n start tuples.
  // Every acceptor must lead to the same match side (same sortedness bits); otherwise fall back to a loop.
  std::vector<const CsrFst*> hs(n);
  bool uniform = n > 0 && dt.has_start;
  uint64_t and_props = ~0ull, or_props = 0;
  size_t sum_states = 0, sum_arcs = 0;
  for (size_t i = 0; i < n; i++) {
    hs[i] = &vec_alg(acceptors[i], "acceptor")->fst.checked();
    if (!hs[i]->inf_finals.empty()) uniform = false;
    if (!hs[i]->has_start) uniform = false;
    and_props &= hs[i]->props; or_props |= hs[i]->props;
    sum_states += hs[i]->num_states(); sum_arcs += hs[i]->arcs.size();
  }
  const uint64_t steer = props::kOLabelSorted | props::kNotOLabelSorted | props::kAcceptor | props::kNotAcceptor |
                         props::kNoEpsilons | props::kNoIEpsilons | props::kNoOEpsilons | props::kAcyclic |
                         props::kInitialAcyclic | props::kIDeterministic | props::kODeterministic;
  if ((and_props & steer) != (or_props & steer)) uniform = false;  // property bits that steer compose must agree
  if (sum_states >= 0x7FFFFFF0ull || sum_arcs >= 0xFFFFFFF0ull) uniform = false;

  bool done = false;
  if (uniform) {
    const double tu0 = now_ms();
    CsrFst u;
    u.offsets.resize(sum_states + 1);
    u.arcs.resize(sum_arcs);
    u.finals.resize(sum_states);
    std::vector<uint32_t> starts(n), base_state(n + 1), base_arc(n + 1);
    {
      size_t so = 0, ao = 0;
      for (size_t i = 0; i < n; i++) {
        base_state[i] = (uint32_t)so; base_arc[i] = (uint32_t)ao;
        starts[i] = (uint32_t)(so + hs[i]->start);
        so += hs[i]->num_states(); ao += hs[i]->arcs.size();
      }
      base_state[n] = (uint32_t)so; base_arc[n] = (uint32_t)ao;
    }
    // The union is built in page-locked memory chunk by chunk and every chunk goes to the device as soon as it is
    // complete: host threads fill chunk k + 1 while the copy engine moves chunk k (the upload is PCIe-bound).
    u.offsets[sum_states] = base_arc[n];
    u.has_start = true; u.start = starts[0];
    u.props = and_props & props::kTrinary;
    DevFst du(st.s);
    du.offsets.reserve_discard(sum_states + 1);
    du.arcs.reserve_discard(sum_arcs ? sum_arcs : 1);
    du.finals.reserve_discard(sum_states ? sum_states : 1);
    du.num_states = (uint32_t)sum_states; du.num_arcs = (uint32_t)sum_arcs;
    du.has_start = true; du.start = u.start; du.props = u.props;
    DevBuf<uint32_t> d_starts(st.s, n), d_s1(st.s), d_map(st.s);
    B200_CUDA(cudaMemcpyAsync(d_starts.p, starts.data(), n * 4, cudaMemcpyHostToDevice, st.s));
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = n < 64 ? 1 : std::min<size_t>(hw ? hw : 1, 8);
    const size_t n_chunks = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, n), sum_arcs / (128 * 1024)));
    std::vector<size_t> bound(n_chunks + 1);
    for (size_t k = 0; k <= n_chunks; k++) bound[k] = n * k / n_chunks;
    std::vector<std::atomic<uint32_t>> chunk_done(n_chunks);
    for (auto& d : chunk_done) d.store(0);
    auto build = [&](size_t lo_i, size_t hi_i) {
      for (size_t i = lo_i; i < hi_i; i++) {
        const CsrFst& h = *hs[i];
        const size_t so = base_state[i], ao = base_arc[i], ns = h.num_states(), na = h.arcs.size();
#if defined(__SSE2__)
        // Streaming stores: the copy engine reads these buffers next, and lines that sit dirty in a CPU cache are
        // served to it several times slower than lines that went straight to memory.
        uint32_t* const uo = u.offsets.data() + so;
        int* const uf = reinterpret_cast<int*>(u.finals.data() + so);
        const int* const hf = reinterpret_cast<const int*>(h.finals.data());
        for (size_t s = 0; s < ns; s++) {
          _mm_stream_si32(reinterpret_cast<int*>(uo + s), (int)(uint32_t)(ao + h.offsets[s]));
          _mm_stream_si32(uf + s, hf[s]);
        }
        __m128i* const ua = reinterpret_cast<__m128i*>(u.arcs.data() + ao);  // 16-byte records in a 16-byte aligned pool block
        const __m128i add = _mm_set_epi32((int)(uint32_t)so, 0, 0, 0);           // nextstate is the last word of a record
        for (size_t k = 0; k < na; k++)
          _mm_stream_si128(ua + k, _mm_add_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(&h.arcs[k])), add));
#else
        for (size_t s = 0; s < ns; s++) u.offsets[so + s] = (uint32_t)(ao + h.offsets[s]);
        std::memcpy(u.finals.data() + so, h.finals.data(), ns * 4);
        for (size_t k = 0; k < na; k++) { Tr t = h.arcs[k]; t.nextstate += (uint32_t)so; u.arcs[ao + k] = t; }
#endif
      }
#if defined(__SSE2__)
      _mm_sfence();  // streaming stores are weakly ordered: make them globally visible before the chunk is announced
#endif
    };
    auto worker = [&](size_t t) {
      for (size_t k = 0; k < n_chunks; k++) {
        const size_t len = bound[k + 1] - bound[k], share = (len + nt - 1) / nt;
        const size_t lo_i = std::min(bound[k + 1], bound[k] + t * share), hi_i = std::min(bound[k + 1], lo_i + share);
        build(lo_i, hi_i);
        chunk_done[k].fetch_add(1, std::memory_order_release);
      }
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; t++) th.emplace_back(worker, t);
    double t_union = 0;
    try {
      for (size_t k = 0; k < n_chunks; k++) {
        {  // the calling thread takes share 0 of every chunk
          const size_t len = bound[k + 1] - bound[k], share = (len + nt - 1) / nt;
          build(bound[k], std::min(bound[k + 1], bound[k] + share));
          chunk_done[k].fetch_add(1, std::memory_order_release);
        }
        while (chunk_done[k].load(std::memory_order_acquire) < nt) std::this_thread::yield();
        const size_t s0 = base_state[bound[k]], s1 = base_state[bound[k + 1]], a0 = base_arc[bound[k]], a1 = base_arc[bound[k + 1]];
        const size_t n_off = s1 - s0 + (k + 1 == n_chunks ? 1 : 0);  // the last chunk carries the closing offset
        if (n_off) B200_CUDA(cudaMemcpyAsync(du.offsets.p + s0, u.offsets.data() + s0, n_off * 4, cudaMemcpyHostToDevice, st.s));
        if (s1 > s0) B200_CUDA(cudaMemcpyAsync(du.finals.p + s0, u.finals.data() + s0, (s1 - s0) * 4, cudaMemcpyHostToDevice, st.s));
        if (a1 > a0) B200_CUDA(cudaMemcpyAsync(du.arcs.p + a0, u.arcs.data() + a0, (a1 - a0) * sizeof(Tr), cudaMemcpyHostToDevice, st.s));
      }
      t_union = now_ms() - tu0;
    } catch (...) {
      for (auto& t : th) t.join();
      throw;
    }
    for (auto& t : th) t.join();
    B200_CUDA(cudaStreamSynchronize(st.s));
    acc.ms_h2d += (float)(now_ms() - tu0);
    BatchStarts bs;
    bs.d_starts1 = d_starts.p; bs.n = (uint32_t)n; bs.out_s1 = &d_s1; bs.out_start_map = &d_map;
    ComposeStats cs;
    DevFst dr(st.s);
    if (compose_device_persistent(du, dt, opt, &cs, st.s, &dr, &bs)) {
      // ---- split the union result per acceptor on the device and bring it back as one block
      t0 = now_ms();
      uint64_t launches = 0;
      split_batch_device(dr, d_s1.p, d_map.p, base_state, out, &launches, st.s);
      cs.kernel_launches += launches;
      out.props.resize(n);
      for (size_t i = 0; i < n; i++) {
        uint64_t p = props::of_compose(hs[i]->props, dt.props);
        if (opt.connect) p = props::after_connect(p);
        out.props[i] = p & props::kTrinary;
      }
      acc.ms_d2h += (float)(now_ms() - t0);
      fill(&acc, cs, acc.ms_h2d, acc.ms_d2h);
      if (std::getenv("B200_BATCH_TRACE"))
        std::fprintf(stderr, "[batch] n=%zu union build (overlapped with the upload) %.2f ms, union + h2d %.2f, expand %.2f, connect %.2f, split + d2h %.2f ms\n",
                     n, t_union, acc.ms_h2d, cs.ms_expand, cs.ms_connect, acc.ms_d2h);
      done = true;
    }
  }
  if (!done) {  // heterogeneous batch: one composition at a time, packed on the host
    out = PackedBatch();
    for (size_t i = 0; i < n; i++) {
      DevFst da = upload(vec_alg(acceptors[i], "acceptor")->fst.checked(), st.s);
      ComposeStats cs;
      DevFst dr = compose_device(da, dt, opt, &cs, st.s);
      out.append(download(dr, st.s));
      acc.states_expanded += cs.states_expanded; acc.arcs_iterated += cs.arcs_iterated;
      acc.arcs_emitted += cs.arcs_emitted; acc.waves += cs.waves; acc.states_out += cs.states_out;
      acc.arcs_out += cs.arcs_out; acc.kernel_launches += cs.kernel_launches; acc.emit_launches += cs.emit_launches;
      acc.ms_expand += cs.ms_expand; acc.ms_connect += cs.ms_connect; acc.ms_emit_kernel += cs.ms_emit_kernel;
    }
    if (n == 0) { out.state_off.assign(1, 0); out.arc_off.assign(1, 0); out.offsets.assign(1, 0); }
  }
  if (total) *total = acc;
}

struct B200PackedBatch { PackedBatch b; };

RUSTFST_FFI_RESULT b200_compose_batch(const CFst* const* acceptors, size_t n, const CFst* transducer,
                                      const CComposeConfig* cfg, const CFst** results, B200ComposeStats* total) {
  return wrap([&] {
    for (size_t i = 0; i < n; i++) results[i] = nullptr;
    PackedBatch pb;
    compose_batch_core(acceptors, n, nn(transducer, "transducer"), nullptr, cfg, pb, total);
    // the handles are only released to the caller once every one of them exists
    std::vector<std::unique_ptr<CFst>> made(n);
    for (size_t i = 0; i < n; i++) made[i].reset(new CFst{HostFst(pb.result(i))});
    for (size_t i = 0; i < n; i++) results[i] = made[i].release();
  });
}
RUSTFST_FFI_RESULT b200_compose_batch_packed(const CFst* const* acceptors, size_t n, const CFst* transducer,
                                             const B200DeviceFst* dev_transducer, const CComposeConfig* cfg,
                                             const B200PackedBatch** out, B200ComposeStats* total) {
  return wrap([&] {
    if (!transducer && !dev_transducer) throw FstError("unexpected null pointer: transducer");
    auto r = std::make_unique<B200PackedBatch>();
    compose_batch_core(acceptors, n, transducer, dev_transducer, cfg, r->b, total);
    *out = r.release();
  });
}
RUSTFST_FFI_RESULT b200_packed_batch_info(const B200PackedBatch* b, uint64_t* n, uint64_t* states, uint64_t* trs,
                                          uint64_t* bytes) {
  return wrap([&] {
    nn(b, "batch");
    if (n) *n = b->b.n;
    if (states) *states = b->b.finals.size();
    if (trs) *trs = b->b.arcs.size();
    if (bytes) *bytes = b->b.byte_size();
  });
}
RUSTFST_FFI_RESULT b200_packed_batch_get(const B200PackedBatch* b, size_t i, const CFst** out) {
  return wrap([&] { *out = new CFst{HostFst(nn(b, "batch")->b.result(i))}; });
}
RUSTFST_FFI_RESULT b200_packed_batch_serialize(const B200PackedBatch* b, uint8_t* dst, size_t capacity) {
  return wrap([&] {
    if (capacity < nn(b, "batch")->b.byte_size()) throw FstError("packed batch: destination too small");
    b->b.serialize(nn(dst, "dst"));
  });
}
RUSTFST_FFI_RESULT b200_packed_batch_deserialize(const uint8_t* src, size_t len, const B200PackedBatch** out) {
  return wrap([&] {
    auto r = std::make_unique<B200PackedBatch>();
    r->b = PackedBatch::deserialize(nn(src, "src"), len);
    *out = r.release();
  });
}
RUSTFST_FFI_RESULT b200_packed_batch_destroy(B200PackedBatch* b) {
  return wrap([&] { delete b; });
}
RUSTFST_FFI_RESULT b200_shortest_path_queue_plan(const CFst* fst, int32_t* kind, uint32_t* order_or_scc,
                                                 uint8_t* scc_is_fifo, uint32_t* n_scc) {
  return wrap([&] {
    QueuePlan plan = build_queue_plan(nn(fst, "fst")->fst.checked());
    if (kind) *kind = (int32_t)plan.kind;
    if (n_scc) *n_scc = (uint32_t)plan.scc_is_fifo.size();
    if (order_or_scc) {
      const std::vector<uint32_t>& v = plan.kind == kTopOrderQueue ? plan.order : plan.scc;
      if (!v.empty()) std::memcpy(order_or_scc, v.data(), v.size() * 4);
    }
    if (scc_is_fifo && !plan.scc_is_fifo.empty()) std::memcpy(scc_is_fifo, plan.scc_is_fifo.data(), plan.scc_is_fifo.size());
  });
}
RUSTFST_FFI_RESULT b200_dag_top_order_device(const CFst* fst, uint32_t* order, int32_t* ok, float* ms_device) {
  return wrap([&] {
    const CsrFst& h = nn(fst, "fst")->fst.checked();
    Stream st;
    DevFst d = upload(h, st.s);
    DevBuf<uint32_t> d_order(st.s);
    float ms = 0;
    const bool done = dag_top_order_device(d, d_order, &ms, nullptr, st.s);
    if (ok) *ok = done ? 1 : 0;
    if (ms_device) *ms_device = ms;
    if (done && order && h.num_states()) {
      B200_CUDA(cudaMemcpyAsync(order, d_order.p, h.num_states() * 4, cudaMemcpyDeviceToHost, st.s));
      st.sync();
    }
  });
}
RUSTFST_FFI_RESULT b200_set_device(int device) {
  return wrap([&] {
    require_device();
    B200_CUDA(cudaSetDevice(device));
    configure_device_pool(device);
  });
}
RUSTFST_FFI_RESULT b200_device_count(int* count) {
  return wrap([&] {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
  });
}
RUSTFST_FFI_RESULT b200_device_synchronize(void) {
  return wrap([&] { require_device(); B200_CUDA(cudaDeviceSynchronize()); });
}

}  // extern "C"
