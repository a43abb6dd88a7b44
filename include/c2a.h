/*
 * c2a.h — C ABI of the B200-native gate-graph builder / topological sorter for the
 * circom-2-arithc flattening hot path.
 *
 * The reference (namnc/circom-2-arithc, Rust) has no FFI of its own; the seams this ABI replaces are
 * ordinary Rust calls (all paths relative to the reference tree):
 *
 *   emit side   Compiler::add_signal      src/compiler.rs:139-161
 *               Compiler::add_gate        src/compiler.rs:174-209
 *               Compiler::add_connection  src/compiler.rs:213-278
 *               Compiler::add_inputs/add_outputs/get_signals   src/compiler.rs:131-137,163-171
 *   back end    topological_sort          src/topological_sort.rs:3-50  (call site src/compiler.rs:408-421)
 *               Compiler::build_circuit   src/compiler.rs:321-494
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, one caller thread per handle,
 * the CUDA device is chosen when the handle is created.  Every entry point returns a c2a_status:
 * 0 = ok, positive = a reference error (CircuitError variant, src/compiler.rs:550-576, or a place
 * where the reference panics), negative = CUDA / runtime failure (see c2a_last_error()).
 * There is NO CPU fallback: if no CUDA device is usable the back-end calls fail with C2A_ERR_CUDA.
 */
#ifndef C2A_H_
#define C2A_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2A_ABI_VERSION 1
#define C2A_NONE 0xFFFFFFFFu /* "no gate" / "no wire" marker in u32 outputs */

/* AGateType discriminants, declaration order of src/a_gate_type.rs:7-28.  The Bristol op token is the
 * variant name (strum Display), see c2a_gate_type_name(). */
typedef enum {
  C2A_AAdd = 0, C2A_ADiv, C2A_AEq, C2A_AGEq, C2A_AGt, C2A_ALEq, C2A_ALt, C2A_AMul, C2A_ANeq, C2A_ASub,
  C2A_AXor, C2A_APow, C2A_AIntDiv, C2A_AMod, C2A_AShiftL, C2A_AShiftR, C2A_ABoolOr, C2A_ABoolAnd,
  C2A_ABitOr, C2A_ABitAnd, C2A_GATE_TYPE_COUNT
} c2a_gate_type;

typedef enum {
  C2A_OK = 0,
  C2A_ERR_CYCLIC_DEPENDENCY = 1,       /* CircuitError::CyclicDependency{"detected at i=<err_index>"}  topological_sort.rs:34-38 */
  C2A_ERR_INCONSISTENCY = 2,           /* CircuitError::Inconsistency{message}                          compiler.rs:337,347,375 */
  C2A_ERR_SIGNAL_ALREADY_DECLARED = 3, /* CircuitError::SignalAlreadyDeclared                           compiler.rs:146-148 */
  C2A_ERR_CANNOT_MERGE_OUTPUT_NODES = 4,   /* compiler.rs:239-241 */
  C2A_ERR_CANNOT_MERGE_CONSTANT_NODES = 5, /* compiler.rs:243-245 */
  C2A_ERR_REFERENCE_PANIC = 6,         /* the reference panics here: add_gate on an undeclared out signal
                                          (compiler.rs:201 unwrap) or a constant whose node never got a
                                          wire (compiler.rs:473 index) */
  C2A_ERR_INVALID_ARGUMENT = 7,
  C2A_ERR_EVALUATION = 8,              /* c2a_evaluate: a gate reads a wire without a value or divides by zero (the reference's
                                          test-side executor panics there, tests/integration.rs:90-119) */
  C2A_ERR_CUDA = -1,                   /* CUDA runtime / launch failure, or no device */
  C2A_ERR_NO_MEMORY = -2
} c2a_status;

/* ArithmeticGate{op, lh_in, rh_in, out} (src/compiler.rs:85-90): 16 bytes, fields are NODE ids. */
typedef struct { uint32_t op, lh, rh, out; } c2a_gate;

/* One record of the emission event stream (what src/process.rs calls on the Compiler, in order). */
typedef enum { C2A_EV_SIGNAL = 0, C2A_EV_SIGNAL_CONST = 1, C2A_EV_GATE = 2, C2A_EV_CONNECT = 3 } c2a_event_kind;
typedef struct {
  uint32_t kind; /* c2a_event_kind | (c2a_gate_type << 8) for C2A_EV_GATE */
  uint32_t a;    /* SIGNAL: id        GATE: lhs signal   CONNECT: a */
  uint32_t b;    /* SIGNAL_CONST: value GATE: rhs signal CONNECT: b */
  uint32_t c;    /* GATE: out signal */
} c2a_event;

typedef struct c2a_handle c2a_handle;     /* device context: stream, scratch pool */
typedef struct c2a_compiler c2a_compiler; /* host emitter: the Compiler restated over a union-find */

/* ---- library ---- */
int c2a_abi_version(void);
const char* c2a_gate_type_name(uint32_t op);          /* "AAdd" ... "ABitAnd"; NULL if out of range */
int c2a_gate_type_from_name(const char* name);        /* -1 if unknown (strum EnumString) */
const char* c2a_status_string(int status);            /* thiserror Display strings of compiler.rs:550-576 */

/* ---- device context ---- */
int c2a_create(int device, c2a_handle** out);         /* C2A_ERR_CUDA when no usable device */
void c2a_destroy(c2a_handle* h);
const char* c2a_last_error(const c2a_handle* h);      /* text of the last non-zero status on this handle */
int c2a_device_count(void);                           /* 0 when CUDA is unavailable */
uint64_t c2a_kernel_launches(const c2a_handle* h);    /* number of kernels this handle has launched so far */
double c2a_last_kernel_ms(const c2a_handle* h, const char* name); /* CUDA-event time of the named phase in the last
                                                         call ("producer","deps","relax","trees","wire_first",
                                                         "wire_scan","gather","kahn","total"); <0 if unknown */
const char* c2a_last_phases(c2a_handle* h);           /* "name=ms,..." for every phase of the last call */
void c2a_set_timing(c2a_handle* h, int on);           /* phase events on/off (default on) */
void c2a_set_timing_only(c2a_handle* h, const char* phase); /* record events around this one phase only (NULL = every phase).
                                                         Each event pair costs ~4 us of stream time; ~35 phases per step add up
                                                         to 0.3 ms at 10 M gates, so a throughput run times one kernel at a time. */
void* c2a_stream(c2a_handle* h);                      /* the cudaStream_t every kernel of this handle runs on */

/* ---- back end (device).  Host-pointer forms copy H2D/D2H inside the call. ---- */

/* topological_sort as called from build_circuit: deps(g) = [producer[lh], producer[rh]] where
 * producer[node] = LAST gate whose out is node (src/compiler.rs:401-421).  order_out[G] is the exact DFS
 * post-order of src/topological_sort.rs.  On C2A_ERR_CYCLIC_DEPENDENCY *err_index = the i of the message. */
int c2a_topo_sort(c2a_handle*, const c2a_gate* gates, uint64_t G, uint32_t node_bound,
                  uint32_t* order_out, uint64_t* err_index);

/* Generic get_deps form (src/topological_sort.rs:3-6): item i depends on dep_idx[dep_off[i] .. dep_off[i+1]),
 * visited in that order.  At most 2 deps per item are accelerated on the device (the only shape the
 * reference ever passes); rows longer than 2 return C2A_ERR_INVALID_ARGUMENT. */
int c2a_topo_sort_deps(c2a_handle*, uint64_t n, const uint64_t* dep_off, const uint32_t* dep_idx,
                       uint32_t* order_out, uint64_t* err_index);

/* Compiler::build_circuit minus the string maps (src/compiler.rs:388-464): inputs take wires 0..n_in in
 * the order given, intermediate nodes first-seen over the sorted gates [lh, rh, out] skipping output
 * nodes, outputs last in the order given.  wire_of_node[node_bound] gets C2A_NONE for nodes without a wire.
 * Any of order_out / wire_of_node / new_gates may be NULL when not wanted. */
int c2a_build_circuit(c2a_handle*, const c2a_gate* gates, uint64_t G, uint32_t node_bound,
                      const uint32_t* input_nodes, uint32_t n_in,
                      const uint32_t* output_nodes, uint32_t n_out,
                      uint32_t* order_out, uint32_t* wire_of_node, c2a_gate* new_gates,
                      uint32_t* wire_count, uint64_t* err_index);

/* Same, all pointers are DEVICE pointers on the handle's device (no copies; asynchronous on the handle's
 * stream until the final status read).  Used by bench.py for the HBM-resident number. */
int c2a_build_circuit_device(c2a_handle*, const c2a_gate* d_gates, uint64_t G, uint32_t node_bound,
                             const uint32_t* d_input_nodes, uint32_t n_in,
                             const uint32_t* d_output_nodes, uint32_t n_out,
                             uint32_t* d_order_out, uint32_t* d_wire_of_node, c2a_gate* d_new_gates,
                             uint32_t* wire_count, uint64_t* err_index);

/* Multi-GPU reconciliation: a rank that sorted one independent component subtree rebases its
 * LOCAL wire ids (inputs [0,n_in), intermediates [n_in,n_in+n_mid), outputs after) and gate indices to the global
 * numbering once the per-rank counts are known (one NCCL all-gather by the caller).  d_order may be NULL. */
int c2a_rebase_wires_device(c2a_handle*, c2a_gate* d_new_gates, uint32_t* d_order, uint64_t G, uint32_t n_in, uint32_t n_mid,
                            uint32_t off_in, uint32_t off_mid, uint32_t off_out, uint32_t gate_base);

/* Same rebase with the offsets computed ON THE DEVICE from the all-gathered per-rank counts: d_counts[world][4] (u64, rank-major:
 * n_in, n_mid, n_out, G) as NCCL left them.  No host read of the counts, and the call only enqueues work on the handle's
 * stream (it returns before the kernel ran; the next synchronising call on the handle orders after it). */
int c2a_rebase_wires_gathered_device(c2a_handle*, c2a_gate* d_new_gates, uint32_t* d_order, uint64_t G, const uint64_t* d_counts,
                                     uint32_t rank, uint32_t world);

/* the gathered-counts rebase for any array of LOCAL wire ids (a wire map, a list of named wires); C2A_NONE entries stay; enqueue-only */
int c2a_rebase_wire_ids_gathered_device(c2a_handle*, uint32_t* d_wire_ids, uint64_t n, const uint64_t* d_counts, uint32_t rank, uint32_t world);

/* same mapping applied to a node-indexed wire map (entries equal to C2A_NONE are left alone) */
int c2a_rebase_wire_map_device(c2a_handle*, uint32_t* d_wire_of_node, uint64_t n, uint32_t n_in, uint32_t n_mid,
                               uint32_t off_in, uint32_t off_mid, uint32_t off_out);

/* Where does ONE gate vector split into `world` independent component subtrees?  bounds_out[world + 1] receives
 * contiguous gate ranges [bounds[k], bounds[k+1]) balanced by gate count such that no dependency edge (src/compiler.rs:408-421) and
 * no non-I/O node crosses a bound: building every range on its own GPU with the SAME input / output node lists and shifting the wire
 * ids by the other ranks' intermediate counts (c2a_emitted_gather_device, c2a_rebase_*) reproduces the single-GPU circuit bit for
 * bit.  *n_shards = world, or 1 when the DAG has fewer independent subtrees (one SHA-256 / Keccak instance, one long chain):
 * replicas only.  d_gates: DEVICE pointer (it may be the resident emitted circuit); the I/O node lists are host pointers.  All on
 * the device (first / last use per node, dependency spans, one running maximum) - the numpy planner of the Python mirror walks
 * the gate vector on the host. */
int c2a_plan_shards_device(c2a_handle*, const c2a_gate* d_gates, uint64_t G, uint32_t node_bound, const uint32_t* input_nodes, uint32_t n_in,
                           const uint32_t* output_nodes, uint32_t n_out, uint32_t world, uint64_t* bounds_out, uint32_t* n_shards);

/* Level-synchronous Kahn frontier over the same dependency relation (not the reference order; used for the
 * layer-wise sweeps and the evaluator).  level_order[G] is level-major; level_off[*n_levels+1] delimits levels
 * (caller provides capacity level_cap+1; more levels than level_cap -> C2A_ERR_INVALID_ARGUMENT).
 * Leftover gates => C2A_ERR_CYCLIC_DEPENDENCY (err_index = smallest gate index on/behind a cycle). */
int c2a_topo_levels(c2a_handle*, const c2a_gate* gates, uint64_t G, uint32_t node_bound,
                    uint32_t* level_order, uint32_t* level_off, uint32_t level_cap, uint32_t* n_levels,
                    uint64_t* err_index);
int c2a_topo_levels_device(c2a_handle*, const c2a_gate* d_gates, uint64_t G, uint32_t node_bound,
                           uint32_t* d_level_order, uint32_t* d_level_off, uint32_t level_cap,
                           uint32_t* n_levels, uint64_t* err_index);

/* Opt-in layer-wise sweeps (NOT in the reference; never alter build_circuit output).
 * const_mask[g]=1 when both operands are constant or constant-derived, const_value[g] its u32 value under
 * execute_op semantics (src/process.rs:649-750; gates that would error are left non-constant);
 * dead_mask[g]=1 when gate g does not reach any node in output_nodes. */
int c2a_sweep_masks(c2a_handle*, const c2a_gate* gates, uint64_t G, uint32_t node_bound,
                    const uint32_t* const_nodes, const uint32_t* const_values, uint32_t n_const,
                    const uint32_t* output_nodes, uint32_t n_out,
                    uint8_t* const_mask, uint32_t* const_value, uint8_t* dead_mask, uint64_t* err_index);

/* u32 circuit evaluator (restates the reference's test-side simulator, tests/integration.rs:90-119, release-build
 * semantics: add/sub/mul/pow wrap, shifts by >= 32 give 0, division by zero fails).  gates[G] are in WIRE ids and in an
 * executable order (what build_circuit returns); values[wire_count] / has[wire_count] carry the initial assignment in
 * (inputs + constants) and every wire's value out.  The circuit must be single-assignment (else INVALID_ARGUMENT).
 * Evaluated level by level on the device; on C2A_ERR_EVALUATION *err_index = the position of the first gate the
 * straight-line executor would fail at, and the contents of values/has are unspecified. */
int c2a_evaluate(c2a_handle*, const c2a_gate* gates, uint64_t G, uint32_t wire_count, uint32_t* values, uint8_t* has, uint64_t* err_index);

/* ---- emit side (device).  The whole event stream is replayed on the GPU: add_signal / add_gate / add_connection
 * (src/compiler.rs:139-278) with the reference's node-id allocation reproduced exactly (effective connections =
 * minimum spanning forest of the connection graph under event order; see csrc/c2a_emit.cuh).  The node-id gate vector
 * and the signal -> node map stay RESIDENT on the handle and feed c2a_emitted_build_circuit without a host round trip.
 * Streams on which the reference would error, or that use the "signal in no node => node 0" quirk (:183), are replayed
 * by the exact host emitter below (info.path says which ran); its status / *err_event are returned unchanged. ---- */
#define C2A_EMIT_PATH_DEVICE 1u
#define C2A_EMIT_PATH_HOST 2u
typedef struct {
  uint64_t n_events, n_signals, n_gates, n_connections;
  uint64_t n_effective;    /* connections that merged two nodes (each consumes a node id, compiler.rs:257) */
  uint32_t node_count;     /* the reference's node_count: node ids are 1..node_count */
  uint32_t signal_bound;   /* 1 + largest declared signal id (0 when ids are too sparse for a dense map) */
  uint32_t path;           /* C2A_EMIT_PATH_DEVICE / C2A_EMIT_PATH_HOST */
  uint32_t rounds;         /* Boruvka rounds (device path) */
  uint32_t decline_flags;  /* non-zero: why the device path handed the stream to the host emitter */
  uint32_t reserved;
} c2a_emit_info;
/* ev: host pointer (pinned memory makes the copy asynchronous). */
int c2a_emit_events_device(c2a_handle*, const c2a_event* ev, uint64_t n, c2a_emit_info* info, uint64_t* err_event);
/* same, the events are already resident on the handle's device (DEVICE pointer; used for the HBM-resident bench number) */
int c2a_emit_events_resident(c2a_handle*, const c2a_event* d_ev, uint64_t n, c2a_emit_info* info, uint64_t* err_event);
/* copies of the resident result: gates_out[n_gates] (node ids, emission order = Compiler.gates),
 * node_of_signal_out[signal_bound] (0 = never declared).  Either may be NULL. */
int c2a_emitted_fetch(c2a_handle*, c2a_gate* gates_out, uint32_t* node_of_signal_out);
/* c2a_build_circuit on the resident emitted circuit; inputs/outputs are listed as SIGNAL ids (node_bound is
 * node_count + 1, so wire_of_node needs node_count + 1 entries). */
int c2a_emitted_build_circuit(c2a_handle*, const uint32_t* input_signals, uint32_t n_in, const uint32_t* output_signals, uint32_t n_out,
                              uint32_t* order_out, uint32_t* wire_of_node, c2a_gate* new_gates, uint32_t* wire_count, uint64_t* err_index);
/* same with DEVICE pointers for the three result arrays (the I/O signal lists stay host pointers) */
int c2a_emitted_build_circuit_device(c2a_handle*, const uint32_t* input_signals, uint32_t n_in, const uint32_t* output_signals, uint32_t n_out,
                                     uint32_t* d_order_out, uint32_t* d_wire_of_node, c2a_gate* d_new_gates, uint32_t* wire_count,
                                     uint64_t* err_index);

/* One shard of a sharded build of the resident circuit: c2a_emitted_build_circuit_device restricted to the gates
 * [gate_lo, gate_hi) - a range c2a_plan_shards_device returned, so no dependency edge and no non-I/O node leaves it.  The I/O lists
 * are the GLOBAL ones (every rank passes the same); order entries are local to the range (0 .. gate_hi - gate_lo), wire ids local to
 * the shard: c2a_rebase_wires_device / c2a_rebase_wire_map_device shift them once the ranks have exchanged their counts. */
int c2a_emitted_build_range_device(c2a_handle*, uint64_t gate_lo, uint64_t gate_hi, const uint32_t* input_signals, uint32_t n_in,
                                   const uint32_t* output_signals, uint32_t n_out, uint32_t* d_order_out, uint32_t* d_wire_of_node, c2a_gate* d_new_gates,
                                   uint32_t* wire_count, uint64_t* err_index);
/* DEVICE pointer to the resident node-id gate vector (n_gates records; NULL when nothing is resident).  Valid until the next call
 * on the handle that emits, or that grows the handle's scratch (any build): take it again after such a call. */
const c2a_gate* c2a_emitted_gates_device(c2a_handle*);

/* Sharded builds, second half: after c2a_emitted_build_circuit_device(..., d_new_gates = NULL, ...) numbered the
 * rank's own circuit and the ranks all-gathered their (n_in, n_mid, n_out, G) counts, this gathers the renumbered gates and
 * applies the global offsets on the fly - no separate rebase pass over the gates.  d_order is the order array of that build
 * (its entries are shifted to global gate indices; may be NULL); d_counts as for c2a_rebase_wires_gathered_device, or NULL for a
 * plain gather (world = 1).  Enqueue-only. */
int c2a_emitted_gather_device(c2a_handle*, uint32_t* d_order, c2a_gate* d_new_gates, const uint64_t* d_counts, uint32_t rank, uint32_t world);

/* Wire ids of selected signals after c2a_emitted_build_circuit[_device] on this handle (what Compiler::build_circuit looks up for
 * its input / output / constant name maps, src/compiler.rs:323-383, 466-493): wires_out[i] = wire of the node holding
 * signals[i], C2A_NONE when the signal was never declared or its node has no wire.  With this a caller that wants the
 * reference's result (gates + named wires) can pass NULL for order_out and wire_of_node and skip their device->host copies. */
int c2a_emitted_signal_wires(c2a_handle*, const uint32_t* signals, uint64_t n, uint32_t* wires_out);
/* node ids of selected signals of the resident emitted circuit (0 = never declared): the input / output <=> node pairing of
 * Compiler::build_circuit (src/compiler.rs:327-383) without copying the whole signal -> node map.  Needs only the emit. */
int c2a_emitted_signal_nodes(c2a_handle*, const uint32_t* signals, uint64_t n, uint32_t* nodes_out);
/* same with DEVICE pointers for both lists; enqueue-only (dense signal ids only: the signal -> node map must be resident) */
int c2a_emitted_signal_wires_device(c2a_handle*, const uint32_t* d_signals, uint64_t n, uint32_t* d_wires_out);

/* ---- packed event stream.  The same emission calls at ~6 bytes per event instead of 16: what the walker hands to the
 * device emitter when the stream has to cross PCIe (c2a_program_packed) and what the device reads from HBM.
 *   kinds[n_events]  one byte per event: c2a_event_kind | (c2a_gate_type << 2)
 *   words[n_words]   payload in event order:  SIGNAL / SIGNAL_CONST: the signal id - omitted when C2A_PACKED_DENSE_IDS is set
 *                    (the ids are 0, 1, 2, ... in declaration order, which is what Runtime::gen_signal produces,
 *                    src/runtime.rs:120-125);  GATE: lhs, rhs, out signal;  CONNECT: a, b.
 * Constant VALUES are not part of it: they never influence node ids, gates or errors (src/compiler.rs:139-278) and stay with
 * the host-side name / constant maps.  n_words must equal 3*gates + 2*connections (+ signals without DENSE_IDS). ---- */
#define C2A_PACKED_DENSE_IDS 1u
/* With dense ids the two operands the walker always derives from the signal it declared last need no payload word either
 * (src/process.rs:470-475 creates the temporary right before its gate; :241-273 connects that temporary to the assigned signal):
 * the op of a GATE then lives in bits 2..6 of its kind byte and bit 7 set means "out = the signal declared last" (payload: lhs,
 * rhs); bit 7 set on a CONNECT means "a = the signal declared last" (payload: b).  Events without bit 7 carry all their words.
 * n_words = 3*gates + 2*connections - flagged events.  4.0 B/event instead of 6.0 on a walker stream: it is PCIe that bounds
 * the host-to-device form of the emitter. */
#define C2A_PACKED_IMPLICIT_OPERANDS 2u
typedef struct {
  const uint8_t* kinds;
  const uint32_t* words;
  uint64_t n_events, n_words;
  uint32_t flags, reserved;
} c2a_packed_events;
/* AoS -> packed.  Returns n_words and *flags_out; writes kinds_out[n] / words_out[n_words] when they are non-NULL
 * (call once with NULL buffers to size them). */
uint64_t c2a_pack_events(const c2a_event* ev, uint64_t n, uint8_t* kinds_out, uint32_t* words_out, uint32_t* flags_out);
/* same; allow_flags says which compactions may be used (C2A_PACKED_DENSE_IDS is applied whenever it holds; add
 * C2A_PACKED_IMPLICIT_OPERANDS for the 4 B/event form).  *flags_out = the flags the stream actually carries. */
uint64_t c2a_pack_events_ex(const c2a_event* ev, uint64_t n, uint32_t allow_flags, uint8_t* kinds_out, uint32_t* words_out, uint32_t* flags_out);
/* packed -> AoS (constant values read back as 0); C2A_ERR_INVALID_ARGUMENT when n_words does not match the kinds */
int c2a_unpack_events(const c2a_packed_events* pk, c2a_event* ev_out /* n_events */);
/* c2a_emit_events_device / _resident on a packed stream (kinds / words are host, resp. DEVICE pointers; the struct itself is
 * always in host memory).  Same results, same c2a_emit_info, same error behaviour (declined streams are unpacked and replayed
 * by the host emitter). */
int c2a_emit_packed_device(c2a_handle*, const c2a_packed_events* pk, c2a_emit_info* info, uint64_t* err_event);
int c2a_emit_packed_resident(c2a_handle*, const c2a_packed_events* d_pk, c2a_emit_info* info, uint64_t* err_event);

/* ---- emit + build in ONE call: what src/program.rs::compile's add_* replay followed by Compiler::build_circuit
 * (src/compiler.rs:139-278, then :321-464) amounts to for a caller that holds the whole recording.  Exactly
 * c2a_emit_packed_device / _resident followed by c2a_emitted_build_circuit / _device - same c2a_emit_info, same arrays, same
 * statuses, the emitted circuit stays resident for c2a_emitted_signal_wires() etc. - but with ONE synchronisation, and for
 * circuits up to ~1 M gates (dense ids, <= 4 M events) the whole pipeline runs inside one cooperative kernel (csrc/c2a_fused.cuh:
 * grid barriers instead of ~55 kernel launches), which is what makes the small BASELINE configs (Poseidon / SHA-256 / Keccak
 * shaped, 1 K - 400 K gates) launch-latency free.  The result arrays are sized by the caller: gates_cap entries for order_out /
 * new_gates, wire_cap entries for wire_of_node (node_count + 1 are needed; n_gates <= n_words / 3 and
 * node_count < n_events always hold).  Too small -> C2A_ERR_INVALID_ARGUMENT with *info filled, circuit still resident. ---- */
typedef struct {
  const uint32_t* input_signals;   /* host pointers: main-template input / output SIGNAL ids, in wire order */
  const uint32_t* output_signals;
  uint32_t n_in, n_out;
  uint32_t* order_out;             /* [gates_cap]  may be NULL */
  uint32_t* wire_of_node;          /* [wire_cap]   may be NULL */
  c2a_gate* new_gates;             /* [gates_cap]  may be NULL */
  uint64_t gates_cap;
  uint32_t wire_cap, reserved;
} c2a_compile_io;
/* kinds / words and the three result arrays are HOST pointers (copies inside the call) */
int c2a_compile_packed(c2a_handle*, const c2a_packed_events* pk, const c2a_compile_io* io, c2a_emit_info* info, uint32_t* wire_count,
                       uint64_t* err_event, uint64_t* err_index);
/* kinds / words and the three result arrays are DEVICE pointers (the structs themselves and the I/O lists are in host memory) */
int c2a_compile_packed_resident(c2a_handle*, const c2a_packed_events* d_pk, const c2a_compile_io* io, c2a_emit_info* info, uint32_t* wire_count,
                                uint64_t* err_event, uint64_t* err_index);
/* tuning / testing: largest event count the single-kernel path is used for (0 = never; default and maximum 4 Mi events) and the
 * events per CTA that size its grid (0 = keep).  Process-wide. */
void c2a_set_fused_limits(uint64_t max_events, uint32_t events_per_cta);

/* ---- emit side (host).  Node ids, gate vector and error behaviour identical to the reference Compiler. ---- */
c2a_compiler* c2a_compiler_new(void);
void c2a_compiler_free(c2a_compiler*);
const char* c2a_compiler_last_error(const c2a_compiler*);
/* name may be NULL (an unnamed temporary); has_value != 0 makes it a constant signal */
int c2a_add_signal(c2a_compiler*, uint32_t id, const char* name, int has_value, uint32_t value);
int c2a_add_gate(c2a_compiler*, uint32_t op, uint32_t lhs_signal, uint32_t rhs_signal, uint32_t out_signal);
int c2a_add_connection(c2a_compiler*, uint32_t a, uint32_t b);
/* bulk replay of an event stream; stops at the first error and stores its index in *err_event */
int c2a_emit_events(c2a_compiler*, const c2a_event* ev, uint64_t n, uint64_t* err_event);
/* I/O tagging: explicit, or the reference's prefix match over signal names (src/program.rs:57-66) */
int c2a_add_input(c2a_compiler*, uint32_t signal_id, const char* name);
int c2a_add_output(c2a_compiler*, uint32_t signal_id, const char* name);
int c2a_tag_inputs_by_prefix(c2a_compiler*, const char* prefix);
int c2a_tag_outputs_by_prefix(c2a_compiler*, const char* prefix);
/* resolve gates to node ids (what the reference keeps up to date by rewriting every gate per connection) */
uint64_t c2a_num_gates(const c2a_compiler*);
uint32_t c2a_node_count(const c2a_compiler*);        /* the reference's node_count; node ids are 1..node_count */
uint64_t c2a_num_signals(const c2a_compiler*);
int c2a_get_gates(c2a_compiler*, c2a_gate* out /* num_gates */);
int c2a_signal_node(c2a_compiler*, uint32_t signal_id, uint32_t* node_id); /* 0 when the signal is unknown */
int c2a_signal_nodes(c2a_compiler*, const uint32_t* signal_ids, uint64_t n, uint32_t* node_ids); /* bulk form */
/* name of a declared signal ("random_<id>" for unnamed temporaries, "const_signal_<v>" for bulk constants);
 * returns the length, or -1 when the signal is unknown; copies at most cap-1 bytes + NUL into buf */
int64_t c2a_signal_name(c2a_compiler*, uint32_t signal_id, char* buf, uint64_t cap);
int c2a_set_signal_name(c2a_compiler*, uint32_t signal_id, const char* name);
/* Compiler::get_signals (src/compiler.rs:163-171): ids of the signals whose name starts with prefix, ascending;
 * returns the count (writes at most cap ids) */
uint64_t c2a_get_signals_by_prefix(c2a_compiler*, const char* prefix, uint32_t* ids_out, uint64_t cap);
/* live nodes: ids ascending; flags bit0=is_const bit1=is_out; signals of node i are
 * sig[sig_off[i] .. sig_off[i+1]) in the reference's merge order. Pass NULL to size. */
uint64_t c2a_num_nodes(c2a_compiler*);
int c2a_get_nodes(c2a_compiler*, uint32_t* node_ids, uint8_t* flags, uint64_t* sig_off, uint32_t* sig);

/* Compiler::build_circuit (src/compiler.rs:321-494): host name maps + device sort/renumber.
 * Results are held by the compiler object until the next build/free. Inputs/outputs are numbered in
 * ascending signal-id order (the reference's order is HashMap iteration order, i.e. unspecified). */
int c2a_compiler_build_circuit(c2a_compiler*, c2a_handle*);
uint64_t c2a_circuit_wire_count(const c2a_compiler*);
const uint32_t* c2a_circuit_order(const c2a_compiler*);       /* num_gates */
const c2a_gate* c2a_circuit_gates(const c2a_compiler*);       /* num_gates, wire ids */
/* circuit.info as JSON: {"input_name_to_wire_index":{},"constants":{name:{"value":"..","wire_index":n}},
 * "output_name_to_wire_index":{}} with keys sorted */
const char* c2a_circuit_info_json(const c2a_compiler*);
/* Compiler::generate_circuit_report (src/compiler.rs:287-319, 503-531) as compact JSON: {"inputs":[{"id","names","value"}..],
 * "outputs":[..],"value_type":..}; inputs = nodes no gate writes, outputs = written nodes no gate reads, ascending node id.
 * The string is owned by the compiler object (valid until the next call / free). */
const char* c2a_circuit_report_json(c2a_compiler*, const char* value_type);
/* circuit.txt body (src/main.rs:34-36 calls bristol-circuit's write_bristol; crate un-vendored, text layout PARITY UNPINNED): one
 * line "2 1 <in0> <in1> <out> <Op>\n" per gate, Op = the AGateType Display token.  Returns the byte count; writes when it fits. */
uint64_t c2a_bristol_gate_lines(const c2a_gate* gates, uint64_t G, char* out, uint64_t cap);

int c2a_signal_value(c2a_compiler*, uint32_t signal_id, int* has_value, uint32_t* value); /* Signal.value (src/compiler.rs:17-27) */

/* ---- front end (host): src/program.rs::compile (:18-74) = parse a .circom program (the subset the reference accepts,
 * README.md:16-40) and walk its main template (src/process.rs, src/runtime.rs), issuing add_signal / add_gate /
 * add_connection in the reference's order.  The calls are recorded as a c2a_event stream (for c2a_emit_events_device)
 * and, when `into` is non-NULL, also applied to a host emitter together with the signal names and the prefix-match
 * input/output tagging of src/program.rs:57-66.  Status: 0, or a c2a_program_status whose text (the thiserror Display
 * string of ProgramError, src/program.rs:77-117) is returned by c2a_program_error(). ---- */
typedef enum {
  C2A_PROG_PARSING_ERROR = 101, C2A_PROG_RUNTIME_ERROR = 102, C2A_PROG_CIRCUIT_ERROR = 103, C2A_PROG_EMPTY_DATA_ITEM = 104,
  C2A_PROG_EXPRESSION_NOT_IMPLEMENTED = 105, C2A_PROG_STATEMENT_NOT_IMPLEMENTED = 106, C2A_PROG_INVALID_DATA_TYPE = 107,
  C2A_PROG_MAIN_NOT_A_CALL = 108, C2A_PROG_OPERATION_ERROR = 109, C2A_PROG_OPERATION_NOT_SUPPORTED = 110,
  C2A_PROG_SIGNAL_SUBSTITUTION_NOT_IMPLEMENTED = 111, C2A_PROG_UNDEFINED_CALLABLE = 112, C2A_PROG_CALL_ERROR = 113
} c2a_program_status;
typedef struct c2a_program c2a_program;
c2a_program* c2a_program_new(void);
void c2a_program_free(c2a_program*);
int c2a_program_compile_file(c2a_program*, const char* path, c2a_compiler* into);
int c2a_program_compile_source(c2a_program*, const char* source, const char* include_dir, c2a_compiler* into);
const char* c2a_program_error(const c2a_program*);
uint64_t c2a_program_num_events(const c2a_program*);
const c2a_event* c2a_program_events(const c2a_program*);
uint64_t c2a_program_num_signals(const c2a_program*);
const char* c2a_program_signal_name(const c2a_program*, uint32_t signal_id);
/* many names at once, '\n'-separated; returns the number of bytes they take and writes them when out != NULL and cap suffices */
uint64_t c2a_program_signal_names(const c2a_program*, const uint32_t* signal_ids, uint64_t n, char* out, uint64_t cap);
uint32_t c2a_program_num_inputs(const c2a_program*);   /* signals tagged as circuit inputs (ascending ids) */
uint32_t c2a_program_num_outputs(const c2a_program*);
const uint32_t* c2a_program_inputs(const c2a_program*);
const uint32_t* c2a_program_outputs(const c2a_program*);
/* the recorded calls as a packed stream (arrays owned by the program object, valid until it is freed or recompiled).  This IS
 * the walker's recording - nothing is converted; c2a_program_events() writes the 16-byte records on first request. */
int c2a_program_packed(c2a_program*, c2a_packed_events* out);
/* The recording in COMPRESSED form.  From its second instance on, a (template, arguments) pair is not interpreted again by the
 * walker: its calls are those of the first instance with every signal id shifted (a call runs in an empty context,
 * src/runtime.rs:75-77).  Without a host emitter attached the walker does not even copy them: the packed arrays grow by the
 * instance's length, the range stays unwritten, and a c2a_replay record says where it comes from:
 *     kinds[k_dst .. k_dst + k_len) = kinds[k_src ..],      words[w_dst .. w_dst + w_len) = words[w_src ..] + delta
 * (sources lie before their destination; records ascend by destination and are disjoint; a source may contain destinations
 * of earlier records: gen = 1 + the largest gen inside the source range, so all records of one gen are independent once the
 * smaller gens are done).  c2a_emit_compressed_device() ships only the written ranges and the records and expands them on the
 * GPU; c2a_program_packed() / c2a_program_events() carry the records out on the host first. */
typedef struct { uint64_t k_dst, k_src, k_len, w_dst, w_src, w_len; uint32_t delta, gen; } c2a_replay;
typedef struct {
  const uint8_t* kinds;      /* n_events bytes; ranges that are the destination of a replay record may be unwritten */
  const uint32_t* words;     /* n_words words; likewise */
  uint64_t n_events, n_words;
  const c2a_replay* replays;
  uint64_t n_replays;
  uint32_t max_gen;
  uint32_t flags;            /* C2A_PACKED_DENSE_IDS (always set by the walker) */
} c2a_compressed_events;
int c2a_program_compressed(c2a_program*, c2a_compressed_events* out);
/* expand on the device, then exactly c2a_emit_packed_resident on the expanded stream (same results, same errors) */
int c2a_emit_compressed_device(c2a_handle*, const c2a_compressed_events*, c2a_emit_info* info, uint64_t* err_event);
/* the constant signals (add_signal with a value, src/process.rs:558-579) in declaration order, and their values: what the
 * `constants` map of CircuitInfo is built from (src/compiler.rs:466-493) without reading the event records */
uint64_t c2a_program_num_constants(const c2a_program*);
const uint32_t* c2a_program_constant_signals(const c2a_program*);
const uint32_t* c2a_program_constant_values(const c2a_program*);

#ifdef __cplusplus
}
#endif
#endif /* C2A_H_ */
