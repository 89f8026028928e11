/* rustfst_b200.h — C ABI of librustfst_b200.so: a B200-native drop-in for the compose / shortest-path slice of
 * rustfst-ffi (garvys-org/rustfst v1.3.1 @ 8e1391d).
 *
 * Every function below replaces the rustfst-ffi export of the same name, with the same argument meaning, ownership
 * and error convention; the reference definition is cited as file:line relative to the reference checkout.
 * Functions prefixed b200_ are additions (bulk CSR ingest/export, device residency, timing) that have no
 * counterpart in the reference.
 *
 * Conventions (rustfst-ffi/src/lib.rs:29-85):
 *   - every call returns RUSTFST_FFI_RESULT_OK (0) or RUSTFST_FFI_RESULT_KO (1);
 *   - on KO a message is kept in a thread-local slot; rustfst_ffi_get_last_error() takes it as a heap C string
 *     that the caller releases with rustfst_destroy_string(); AMSTRAM_FFI_ERROR_STDERR also prints it;
 *   - handles are opaque heap objects; inputs are borrowed, outputs are new handles owned by the caller
 *     (fst_destroy / tr_delete / ... are null-safe);
 *   - "optional" out-parameters (fst_start without a start state, fst_final_weight of a non-final state,
 *     fst_input_symbols without a table) return OK and leave the caller's slot untouched
 *     (rustfst-ffi/src/fst/mod.rs:128-155,243-279).
 *   - CLabel = CStateId = unsigned int (feature rustfst-state-label-u32), weights are float.
 *
 * compose / connect / shortest_path run on the GPU (sm_100a).  There is no CPU fallback: without a usable CUDA
 * device those calls return KO with an explanatory message.
 */
#ifndef RUSTFST_B200_H
#define RUSTFST_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum RUSTFST_FFI_RESULT { RUSTFST_FFI_RESULT_OK = 0, RUSTFST_FFI_RESULT_KO = 1 } RUSTFST_FFI_RESULT;

typedef unsigned int CLabel;   /* rustfst-ffi/src/lib.rs:19-22 */
typedef unsigned int CStateId; /* rustfst-ffi/src/lib.rs:24-27 */

typedef struct CFst CFst;                                 /* rustfst-ffi/src/fst/mod.rs:88-89 (Box<dyn BindableFst>) */
typedef struct CTrs CTrs;                                 /* rustfst-ffi/src/trs.rs:14-15 */
typedef struct CTrsIterator CTrsIterator;                 /* rustfst-ffi/src/iterators.rs:40-41 */
typedef struct CMutTrsIterator CMutTrsIterator;           /* rustfst-ffi/src/iterators.rs:133-138 */
typedef struct CStateIterator CStateIterator;             /* rustfst-ffi/src/iterators.rs:321-322 */
typedef struct CSymbolTable CSymbolTable;                 /* rustfst-ffi/src/symbol_table.rs (never produced here) */
typedef struct CComposeConfig CComposeConfig;             /* rustfst-ffi/src/algorithms/compose.rs:163-170 */
typedef struct CMatcherConfig CMatcherConfig;             /* rustfst-ffi/src/algorithms/compose.rs:120-123 */
typedef struct CShortestPathConfig CShortestPathConfig;   /* rustfst-ffi/src/algorithms/shortest_path.rs:11-17 */

/* rustfst-ffi/src/tr.rs:9-28 — repr(C), 16 bytes; also the on-device arc record */
typedef struct CTr {
  CLabel ilabel;
  CLabel olabel;
  float weight;
  CStateId nextstate;
} CTr;

/* ffi_convert::CArray<u8> as used by vec_fst_from_bytes / vec_fst_to_bytes (rustfst-ffi/src/fst/vector_fst.rs:319-354) */
typedef struct CArrayU8 {
  const uint8_t* data_ptr;
  size_t size;
} CArrayU8;

/* rustfst-ffi/src/algorithms/compose.rs:172-177 */
typedef struct CIntArray {
  const uint32_t* data;
  size_t size;
} CIntArray;

/* ---- errors: rustfst-ffi/src/lib.rs:57-85 */
RUSTFST_FFI_RESULT rustfst_ffi_get_last_error(char** error);
RUSTFST_FFI_RESULT rustfst_destroy_string(char* string);

/* ---- the hot path ------------------------------------------------------------------------------------------ */
/* rustfst-ffi/src/algorithms/compose.rs:308-334.  ComposeConfig::default(): AutoFilter (= Sequence filter over
 * sorted matchers) and connect = true.  fst_1 must be olabel-sorted and/or fst_2 ilabel-sorted according to their
 * stored property bits, else KO with the reference's "(sort?)" / "Properties are not known" message. */
RUSTFST_FFI_RESULT fst_compose(const CFst* fst_1, const CFst* fst_2, const CFst** composition_ptr);
/* rustfst-ffi/src/algorithms/compose.rs:340-372 */
RUSTFST_FFI_RESULT fst_compose_with_config(const CFst* fst_1, const CFst* fst_2, const CComposeConfig* config,
                                           const CFst** composition_ptr);
/* rustfst-ffi/src/algorithms/compose.rs:229-268; compose_filter: 0 Auto, 1 Null, 2 Trivial, 3 Sequence,
 * 4 AltSequence, 5 Match, 6 NoMatch (:20-33); matcher configs may be NULL */
RUSTFST_FFI_RESULT fst_compose_config_new(size_t compose_filter, bool connect, const CMatcherConfig* matcher1_config,
                                          const CMatcherConfig* matcher2_config, const CComposeConfig** config);
RUSTFST_FFI_RESULT fst_compose_config_destroy(CComposeConfig* ptr);         /* compose.rs:290-302 */
/* rustfst-ffi/src/algorithms/compose.rs:191-223: sigma matcher config (sigma_label, rewrite_mode 0 Auto / 1 Always /
 * 2 Never, allow-list; empty = any label).  Supported by the persistent compose back end; semantics follow
 * rustfst/src/algorithms/compose/matchers/sigma_matcher.rs (incl. REQUIRE_MATCH / REQUIRE_PRIORITY and its errors). */
RUSTFST_FFI_RESULT fst_matcher_config_new(size_t sigma_label, size_t rewrite_mode, CIntArray sigma_allowed_matches,
                                          const CMatcherConfig** config);
RUSTFST_FFI_RESULT fst_matcher_config_destroy(CMatcherConfig* ptr);         /* compose.rs:274-286 */

/* rustfst-ffi/src/algorithms/shortest_path.rs:44-57 (ShortestPathConfig::default(): nshortest = 1) */
RUSTFST_FFI_RESULT fst_shortest_path(const CFst* ptr, const CFst** res_fst);
/* rustfst-ffi/src/algorithms/shortest_path.rs:62-83.  nshortest > 1 with unique = true (shortest_path.rs:156-165)
 * needs an acceptor, like the reference ("DeterminizeFsaImpl : expected acceptor as argument"); the subsets of the
 * determinized reversed machine are built on demand and kept sorted by state (the reference's own subset order comes
 * out of a RandomState HashMap, determinize_fsa_op.rs:154-165, and is not reproducible; paths and weights are). */
RUSTFST_FFI_RESULT fst_shortest_path_with_config(const CFst* ptr, const CShortestPathConfig* config,
                                                 const CFst** res_fst);
/* rustfst-ffi/src/algorithms/shortest_path.rs:22-39 */
RUSTFST_FFI_RESULT fst_shortest_path_config_new(float delta, size_t nshortest, bool unique,
                                                const CShortestPathConfig** ptr);
RUSTFST_FFI_RESULT b200_shortest_path_config_destroy(CShortestPathConfig* ptr); /* the reference leaks these */

/* rustfst-ffi/src/algorithms/connect.rs:13-23 (in place) */
RUSTFST_FFI_RESULT fst_connect(CFst* ptr);
/* rustfst-ffi/src/algorithms/reverse.rs:14-29 (rustfst/src/algorithms/reverse.rs:33-87): new FST with a superinitial
 * state 0; built on the device (one stable radix sort + gather), the same kernel that feeds the n-best search. */
RUSTFST_FFI_RESULT fst_reverse(const CFst* ptr, const CFst** res_ptr);
/* rustfst-ffi/src/algorithms/top_sort.rs:13-21 (rustfst/src/algorithms/top_sort.rs:75-95): in place; host side (the
 * numbering is the finish order of the reference's sequential DFS).  A composed lattice that is top-sorted once takes
 * the StateOrderQueue route of fst_shortest_path afterwards (no DFS per call). */
RUSTFST_FFI_RESULT fst_top_sort(CFst* ptr);
/* rustfst-ffi/src/algorithms/isomorphic.rs:11-30 (rustfst/src/algorithms/isomorphic.rs:49-160): same machines up to
 * state numbering and arc order (weights within KDELTA); host side. */
RUSTFST_FFI_RESULT fst_isomorphic(const CFst* fst, const CFst* other_fst, size_t* is_isomorphic);
/* rustfst-ffi/src/algorithms/tr_sort.rs:14-30 (in place; host, stable) */
RUSTFST_FFI_RESULT fst_tr_sort(CFst* ptr, bool ilabel_comp);

/* ---- Fst trait accessors: rustfst-ffi/src/fst/mod.rs:127-385 */
RUSTFST_FFI_RESULT fst_start(const CFst* fst, CStateId* state);
RUSTFST_FFI_RESULT fst_final_weight(const CFst* fst, CStateId state_id, float* final_weight);
RUSTFST_FFI_RESULT fst_num_trs(const CFst* fst, CStateId state, size_t* num_trs);
RUSTFST_FFI_RESULT fst_get_trs(const CFst* fst, CStateId state, const CTrs** trs);
RUSTFST_FFI_RESULT fst_is_final(const CFst* fst, CStateId state, size_t* is_final);
RUSTFST_FFI_RESULT fst_is_start(const CFst* fst, CStateId state, size_t* is_start);
RUSTFST_FFI_RESULT fst_input_symbols(const CFst* fst, const CSymbolTable** input_symt);
RUSTFST_FFI_RESULT fst_output_symbols(const CFst* fst, const CSymbolTable** output_symt);
RUSTFST_FFI_RESULT fst_weight_one(float* weight_one);
RUSTFST_FFI_RESULT fst_weight_zero(float* weight_zero);
RUSTFST_FFI_RESULT fst_destroy(CFst* fst_ptr);

/* ---- VectorFst: rustfst-ffi/src/fst/vector_fst.rs:13-354 */
RUSTFST_FFI_RESULT vec_fst_new(const CFst** ptr);
RUSTFST_FFI_RESULT vec_fst_set_start(CFst* fst, CStateId state);
RUSTFST_FFI_RESULT vec_fst_set_final(CFst* fst, CStateId state, float weight);
RUSTFST_FFI_RESULT vec_fst_add_state(CFst* fst, CStateId* state);
RUSTFST_FFI_RESULT vec_fst_delete_states(CFst* fst);
RUSTFST_FFI_RESULT vec_fst_add_tr(CFst* fst, CStateId state, const CTr* tr);
RUSTFST_FFI_RESULT vec_fst_del_final_weight(CFst* fst, CStateId state);
RUSTFST_FFI_RESULT vec_fst_from_path(const CFst** ptr, const char* path);
RUSTFST_FFI_RESULT vec_fst_write_file(const CFst* fst, const char* path);
RUSTFST_FFI_RESULT vec_fst_num_states(const CFst* fst, size_t* num_states);
RUSTFST_FFI_RESULT vec_fst_equals(const CFst* fst, const CFst* other_fst, size_t* is_equal);
RUSTFST_FFI_RESULT vec_fst_copy(const CFst* fst_ptr, const CFst** clone_ptr);
RUSTFST_FFI_RESULT vec_fst_display(const CFst* fst_ptr, const char** s);
RUSTFST_FFI_RESULT vec_fst_to_bytes(const CFst* fst_ptr, const CArrayU8** output_bytes);
RUSTFST_FFI_RESULT vec_fst_from_bytes(const CArrayU8* bytes, const CFst** ptr);

/* ConstFst handles (rustfst-ffi/src/fst/const_fst.rs:10-155; OpenFst binary "const" format, packed and 16-byte aligned
 * versions: rustfst/src/fst_impls/const_fst/serializable_fst.rs).  A const handle answers the generic fst_* accessors;
 * algorithms and vec_fst_* refuse it with the reference's downcast errors.  const_fst_draw is not part of this build. */
RUSTFST_FFI_RESULT const_fst_from_path(const CFst** ptr, const char* path);
RUSTFST_FFI_RESULT const_fst_write_file(const CFst* fst, const char* path);
RUSTFST_FFI_RESULT const_fst_equals(const CFst* fst, const CFst* other_fst, size_t* is_equal);
RUSTFST_FFI_RESULT const_fst_copy(const CFst* fst_ptr, const CFst** clone_ptr);
RUSTFST_FFI_RESULT const_fst_display(const CFst* fst_ptr, const char** s);
/* rustfst-ffi/src/fst/const_fst.rs:157-170: copy of a VectorFst with ALL properties computed (converters.rs:7-37). */
RUSTFST_FFI_RESULT const_fst_from_vec_fst(const CFst* vec_fst_ptr, const CFst** const_fst_ptr);
RUSTFST_FFI_RESULT b200_bytes_destroy(CArrayU8* bytes); /* the reference leaks the array of vec_fst_to_bytes */

/* ---- Tr: rustfst-ffi/src/tr.rs:47-191 (setters take the VALUE in the pointer-typed parameter, as upstream) */
RUSTFST_FFI_RESULT tr_new(CLabel ilabel, CLabel olabel, float weight, CStateId nextstate, const CTr** new_struct);
RUSTFST_FFI_RESULT tr_ilabel(const CTr* tr, CLabel* ilabel);
RUSTFST_FFI_RESULT tr_set_ilabel(CTr* tr, size_t ilabel);
RUSTFST_FFI_RESULT tr_olabel(const CTr* tr, CLabel* olabel);
RUSTFST_FFI_RESULT tr_set_olabel(CTr* tr, size_t olabel);
RUSTFST_FFI_RESULT tr_weight(const CTr* tr, float* weight);
RUSTFST_FFI_RESULT tr_set_weight(CTr* tr, float weight);
RUSTFST_FFI_RESULT tr_next_state(const CTr* tr, CStateId* next_state);
RUSTFST_FFI_RESULT tr_set_next_state(CTr* tr, size_t next_state);
RUSTFST_FFI_RESULT tr_delete(CTr* tr_ptr);

/* ---- Trs: rustfst-ffi/src/trs.rs:18-121 */
RUSTFST_FFI_RESULT trs_vec_new(const CTrs** new_struct);
RUSTFST_FFI_RESULT trs_vec_remove(CTrs* trs, size_t index, const CTr** removed_tr_ptr);
RUSTFST_FFI_RESULT trs_vec_push(CTrs* trs, const CTr* new_tr);
RUSTFST_FFI_RESULT trs_vec_shallow_clone(const CTrs* trs, const CTrs** cloned_trs_ptr);
RUSTFST_FFI_RESULT trs_vec_len(const CTrs* trs, size_t* num_trs);
RUSTFST_FFI_RESULT trs_vec_display(const CTrs* trs, const char** string);
RUSTFST_FFI_RESULT trs_vec_delete(CTrs* trs_ptr);

/* ---- iterators: rustfst-ffi/src/iterators.rs:44-390 */
RUSTFST_FFI_RESULT trs_iterator_new(CFst* fst_ptr, CStateId state_id, const CTrsIterator** iter_ptr);
RUSTFST_FFI_RESULT trs_iterator_next(CTrsIterator* iter_ptr, const CTr** tr_ptr);
RUSTFST_FFI_RESULT trs_iterator_done(const CTrsIterator* iter_ptr, size_t* done);
RUSTFST_FFI_RESULT trs_iterator_reset(CTrsIterator* iter_ptr);
RUSTFST_FFI_RESULT trs_iterator_destroy(CTrsIterator* iter_ptr);
RUSTFST_FFI_RESULT mut_trs_iterator_new(CFst* fst_ptr, CStateId state_id, const CMutTrsIterator** iter_ptr);
RUSTFST_FFI_RESULT mut_trs_iterator_next(CMutTrsIterator* iter_ptr);
RUSTFST_FFI_RESULT mut_trs_iterator_value(CMutTrsIterator* iter_ptr, const CTr** tr_ptr);
RUSTFST_FFI_RESULT mut_trs_iterator_set_value(CMutTrsIterator* iter_ptr, const CTr* tr_ptr);
RUSTFST_FFI_RESULT mut_trs_iterator_done(const CMutTrsIterator* iter_ptr, size_t* done);
RUSTFST_FFI_RESULT mut_trs_iterator_reset(CMutTrsIterator* iter_ptr);
RUSTFST_FFI_RESULT mut_trs_iterator_destroy(CMutTrsIterator* iter_ptr);
RUSTFST_FFI_RESULT state_iterator_new(CFst* fst_ptr, const CStateIterator** iter_ptr);
RUSTFST_FFI_RESULT state_iterator_next(CStateIterator* iter_ptr, CStateId* state);
RUSTFST_FFI_RESULT state_iterator_done(CStateIterator* iter_ptr, size_t* done);
RUSTFST_FFI_RESULT state_iterator_destroy(CStateIterator* iter_ptr);

/* ================================================================================================================
 * b200_ additions
 * ============================================================================================================== */

/* 64-bit FstProperties word of a handle (rustfst/src/fst_properties/properties.rs:21-103); the reference keeps it
 * inside VectorFst but never exports it through the FFI.  b200_fst_set_properties overwrites the stored bits
 * (trusted verbatim, like a header read: serializable_fst.rs:165). */
RUSTFST_FFI_RESULT b200_fst_properties(const CFst* fst, uint64_t* props);
RUSTFST_FFI_RESULT b200_fst_set_properties(CFst* fst, uint64_t props);

/* Bulk ingest/export in the device layout (vec_fst_add_tr is one FFI call per arc: unusable at 10^7 arcs).
 * offsets: num_states + 1 entries; arcs: offsets[num_states] records; finals: +inf = not final; start < 0 = none. */
RUSTFST_FFI_RESULT b200_fst_from_csr(uint64_t num_states, const uint32_t* offsets, const CTr* arcs, const float* finals,
                                     int64_t start, uint64_t props, const CFst** out);
RUSTFST_FFI_RESULT b200_fst_num_states(const CFst* fst, uint64_t* num_states); /* any handle kind (vector or const) */
/* compute_and_update_properties_all (rustfst/src/fst_traits/mutable_fst.rs:435-446): fills in every unknown property
 * bit from the machine's content (host; e.g. the sortedness bits compose needs on a machine built from raw arrays). */
RUSTFST_FFI_RESULT b200_fst_compute_properties(CFst* fst, uint64_t* props /* may be NULL */);
RUSTFST_FFI_RESULT b200_fst_num_trs_total(const CFst* fst, uint64_t* num_trs);
/* Copies into caller buffers sized with vec_fst_num_states / b200_fst_num_trs_total (any pointer may be NULL). */
RUSTFST_FFI_RESULT b200_fst_to_csr(const CFst* fst, uint32_t* offsets, CTr* arcs, float* finals, int64_t* start);

/* Counters and device timings of one call (all times are CUDA-event milliseconds on the call's stream). */
typedef struct B200ComposeStats {
  uint64_t states_expanded, arcs_iterated, arcs_emitted, waves, states_out, arcs_out;
  uint64_t kernel_launches, emit_launches;
  float ms_expand, ms_connect, ms_emit_kernel;
  float ms_h2d, ms_d2h; /* host<->device marshalling inside the host-buffer entry points (wall clock) */
  float ms_phase_match, ms_phase_emit, ms_phase_rank, ms_phase_resolve; /* persistent kernel, %globaltimer */
} B200ComposeStats;
typedef struct B200SsspStats {
  uint64_t arcs_relaxed, states_settled, waves, kernel_launches, relax_launches;
  int32_t path;      /* 0 parallel relaxation + certificate, 1 order-faithful serial kernel, 2 order-faithful
                        parallel fold; for nshortest > 1: the path the forward-distance pass took */
  int32_t queue_kind; /* 0 StateOrder, 1 TopOrder, 2 Lifo, 3 Scc */
  float ms_device, ms_relax_kernel;
  float ms_h2d;
  double ms_queue_plan_host; /* host time spent building the queue plan (0 when the order was computed on the device) */
  float ms_order_device;     /* device time of the TopOrderQueue order of an acyclic machine (dag_order.cu), else 0 */
  int32_t order_on_device;   /* 1: that order was computed on the device, 0: host DFS or no DFS order needed */
  int32_t sweep;             /* 1: the relaxation waves went over their visit budget (a deep top-sorted DAG with skip
                                arcs) and the distances come from the in-order sweep kernel */
} B200SsspStats;
/* Same as fst_compose_with_config (config may be NULL = default) but also reports stats. */
RUSTFST_FFI_RESULT b200_compose_with_stats(const CFst* fst_1, const CFst* fst_2, const CComposeConfig* config,
                                           const CFst** composition_ptr, B200ComposeStats* stats);
RUSTFST_FFI_RESULT b200_shortest_path_with_stats(const CFst* ptr, const CShortestPathConfig* config,
                                                 const CFst** res_fst, B200SsspStats* stats, bool force_serial);

/* Device-resident FSTs: inputs already in HBM when a timed region starts, results left in HBM. */
typedef struct B200DeviceFst B200DeviceFst;
RUSTFST_FFI_RESULT b200_device_fst_upload(const CFst* fst, const B200DeviceFst** out);
RUSTFST_FFI_RESULT b200_device_fst_download(const B200DeviceFst* dfst, const CFst** out);
RUSTFST_FFI_RESULT b200_device_fst_info(const B200DeviceFst* dfst, uint64_t* num_states, uint64_t* num_trs,
                                        uint64_t* props);
RUSTFST_FFI_RESULT b200_device_fst_destroy(B200DeviceFst* dfst);
RUSTFST_FFI_RESULT b200_device_compose(const B200DeviceFst* fst_1, const B200DeviceFst* fst_2,
                                       const CComposeConfig* config, const B200DeviceFst** out,
                                       B200ComposeStats* stats);
/* isomorphic (rustfst/src/algorithms/isomorphic.rs:49-160) of two device-resident machines, as a verifier that needs
 * no download: *result = 1 isomorphic, 0 not isomorphic, -1 undecided on the device (a check failed after rows with
 * equal neighbouring arcs were visited, where the reference either returns false or raises its non-determinism error
 * depending on the visiting order; fst_isomorphic on the host copies gives the reference's answer). */
RUSTFST_FFI_RESULT b200_device_isomorphic(const B200DeviceFst* fst_1, const B200DeviceFst* fst_2, int32_t* result);

/* plan_from (may be NULL) is the host copy dfst was uploaded from.  The queue discipline is decided from the property
 * word stored with dfst; the DFS order of an acyclic machine is computed on the device.  The host copy is only needed
 * for machines that are not known to be acyclic (Tarjan SCC order) or too deep for the device path (more than 65 536
 * topological levels). */
RUSTFST_FFI_RESULT b200_device_shortest_path(const B200DeviceFst* dfst, const CFst* plan_from, const CFst** res_fst,
                                             B200SsspStats* stats, bool force_serial);

/* Same with a ShortestPathConfig (nshortest > 1: n-best over the device-resident machine; config may be NULL). */
RUSTFST_FFI_RESULT b200_device_shortest_path_with_config(const B200DeviceFst* dfst, const CFst* plan_from,
                                                         const CShortestPathConfig* config, const CFst** res_fst,
                                                         B200SsspStats* stats, bool force_serial);

/* Batched compose: acceptors[i] o transducer for i in [0, n) on the current device (transducer uploaded once). */
RUSTFST_FFI_RESULT b200_compose_batch(const CFst* const* acceptors, size_t n, const CFst* transducer,
                                      const CComposeConfig* config, const CFst** results /* n slots */,
                                      B200ComposeStats* total_stats);

/* The same as one block: the n results stay packed (no per-result handle is created), can be fetched one by one,
 * and serialise to / from one contiguous byte string — what a rank sends to rank 0 over NCCL in the sharded mode.
 * The transducer is given either as a host handle (uploaded by the call) or as a device-resident handle (stays in
 * HBM across calls); exactly one of the two may be NULL. */
typedef struct B200PackedBatch B200PackedBatch;
RUSTFST_FFI_RESULT b200_compose_batch_packed(const CFst* const* acceptors, size_t n, const CFst* transducer,
                                             const B200DeviceFst* dev_transducer, const CComposeConfig* config,
                                             const B200PackedBatch** out, B200ComposeStats* total_stats);
RUSTFST_FFI_RESULT b200_packed_batch_info(const B200PackedBatch* batch, uint64_t* n, uint64_t* num_states,
                                          uint64_t* num_trs, uint64_t* serialized_bytes);
RUSTFST_FFI_RESULT b200_packed_batch_get(const B200PackedBatch* batch, size_t i, const CFst** out);
RUSTFST_FFI_RESULT b200_packed_batch_serialize(const B200PackedBatch* batch, uint8_t* dst, size_t capacity);
RUSTFST_FFI_RESULT b200_packed_batch_deserialize(const uint8_t* src, size_t len, const B200PackedBatch** out);
RUSTFST_FFI_RESULT b200_packed_batch_destroy(B200PackedBatch* batch);

/* The queue discipline rustfst's AutoQueue would pick for `fst` from its stored property bits (host logic, no GPU):
 * kind 0 StateOrder, 1 TopOrder, 2 Lifo, 3 Scc.  order_or_scc (num_states entries, may be NULL) receives order[state]
 * for TopOrder and scc[state] for Scc; scc_is_fifo (num_states entries, may be NULL) the per-component queue type. */
RUSTFST_FFI_RESULT b200_shortest_path_queue_plan(const CFst* fst, int32_t* kind, uint32_t* order_or_scc,
                                                 uint8_t* scc_is_fifo, uint32_t* n_scc);

/* The TopOrderQueue order of an acyclic machine computed ON THE DEVICE (csrc/dag_order.cu; replaces the sequential DFS
 * of rustfst/src/algorithms/top_sort.rs:12-61 + dfs_visit.rs:97-187 for acyclic inputs): order[state] = position in
 * reverse DFS finish order.  *ok = 0 when the machine is cyclic or deeper than the device path handles (then `order`
 * is untouched and b200_shortest_path_queue_plan gives the host answer).  Tests compare the two. */
RUSTFST_FFI_RESULT b200_dag_top_order_device(const CFst* fst, uint32_t* order /* num_states */, int32_t* ok,
                                             float* ms_device);

/* Device management for one-process-per-GPU launches. */
RUSTFST_FFI_RESULT b200_set_device(int device);
RUSTFST_FFI_RESULT b200_device_count(int* count);
RUSTFST_FFI_RESULT b200_device_synchronize(void);
const char* b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RUSTFST_B200_H */
