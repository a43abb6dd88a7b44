// c2a_host.cpp — host half of the C ABI: the reference's `Compiler` (src/compiler.rs:107-284) restated over
// a union-find so that emission is O(alpha) per call instead of the reference's O(signals) scans
// (add_gate :185-195, add_connection :219-226) and O(gates) rewrite per connection (:260-270).
//
// What must be bit-identical to the reference, and how it is kept:
//   * node ids: one counter, pre-incremented (:497-500); +1 per add_signal (:157), +1 per EFFECTIVE merge
//     (:257), nothing when both signals already share a node (:235-237).  Each union-find root carries the
//     id of the node it currently represents.
//   * merged node = a's signals followed by b's (:254-255), flags OR-ed (:251-252): roots carry an ordered
//     singly linked list (head/tail) and the two flags.
//   * gates hold NODE ids that the reference keeps current by rewriting every gate on every merge.  Here
//     gates hold union-find ELEMENTS and are resolved once, on demand (c2a_get_gates): same result.
//   * a signal id that is in no node resolves to node id 0 (:183).  Gates that captured "node 0" ARE
//     rewritten by a later add_connection involving an unknown signal (:260-270 compares against id 0), so
//     "node 0" is modelled as a virtual union-find element that is replaced by a fresh one once merged.
//   * errors: SignalAlreadyDeclared (:146), CannotMergeOutputNodes / CannotMergeConstantNodes (:239-245,
//     evaluated at merge time), add_gate on an undeclared out signal panics upstream (:201) -> C2A_ERR_REFERENCE_PANIC.
//
// build_circuit (:321-494): the string maps are built here, the sort / wire numbering / gather run on the
// device through c2a_build_circuit().  There is no CPU implementation of that part in this library.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/c2a.h"

namespace {

const char* kGateNames[C2A_GATE_TYPE_COUNT] = {"AAdd", "ADiv", "AEq", "AGEq", "AGt", "ALEq", "ALt", "AMul", "ANeq", "ASub",
                                               "AXor", "APow", "AIntDiv", "AMod", "AShiftL", "AShiftR", "ABoolOr", "ABoolAnd", "ABitOr", "ABitAnd"};

constexpr uint32_t kNoElem = 0xFFFFFFFFu;
constexpr uint8_t kConst = 1, kOut = 2;

struct ElemGate {
  uint32_t op, l, r, o;  // union-find elements
};

}  // namespace

struct Elem {  // one union-find element = one declared signal (or a virtual "node 0" stand-in)
  uint32_t parent;
  uint32_t node_id;   // root: id of the node this class currently is (0 for an unmerged virtual element)
  uint32_t head, tail;  // root: first/last element of the ordered signal list (kNoElem when empty)
  uint32_t next;      // next element in its node's signal list
  uint32_t sig_id;    // kNoElem for virtual elements
  uint32_t value;
  uint32_t name_off;  // kNameNone: unnamed temp ("random_<id>"), kNameConst: "const_signal_<value>", else offset in name_pool
  uint8_t rank, flags, has_value;  // root: rank, kConst | kOut
};
constexpr uint32_t kNameNone = 0xFFFFFFFFu, kNameConst = 0xFFFFFFFEu;

struct c2a_compiler {
  std::vector<Elem> el;  // array-of-structs: one cache line touch per element (2.7x faster emit than per-field vectors)
  std::string name_pool;
  // --- signal id -> element
  std::vector<uint32_t> dense;                    // ids < dense.size()
  std::unordered_map<uint32_t, uint32_t> sparse;  // the rest
  uint64_t n_signals = 0;
  uint32_t zero_elem = kNoElem;  // current "node 0" stand-in
  uint32_t node_count = 0;
  std::vector<ElemGate> gates;
  std::map<uint32_t, std::string> inputs, outputs;  // signal id -> name (ascending id = the deterministic stand-in for HashMap order)
  std::vector<uint32_t> const_signals;              // ids with a value, in declaration order
  std::string err;
  // --- build results
  std::vector<uint32_t> order;
  std::vector<c2a_gate> new_gates;
  uint64_t wire_count = 0;
  std::string info_json;
  std::string report_json;

  uint32_t new_elem(uint32_t sid) {
    uint32_t e = (uint32_t)el.size();
    el.push_back(Elem{e, 0, kNoElem, kNoElem, kNoElem, sid, 0, kNameNone, 0, 0, 0});
    return e;
  }
  uint32_t elem_of(uint32_t sid) const {
    if (sid < dense.size()) return dense[sid];
    if (sparse.empty()) return kNoElem;
    auto it = sparse.find(sid);
    return it == sparse.end() ? kNoElem : it->second;
  }
  void bind(uint32_t sid, uint32_t e) {
    // ids are sequential in practice (src/runtime.rs:120-125): keep them in a flat table, spill far-away ids
    if (sid < dense.size()) { dense[sid] = e; return; }
    if (sid == dense.size()) { dense.push_back(e); return; }
    if (sid <= dense.size() + (1u << 20) + dense.size() / 2) {
      dense.resize((size_t)sid + 1, kNoElem);
      dense[sid] = e;
    } else sparse[sid] = e;
  }
  uint32_t find(uint32_t e) {
    Elem* a = el.data();
    uint32_t r = e;
    while (a[r].parent != r) r = a[r].parent;
    while (a[e].parent != r) { uint32_t n = a[e].parent; a[e].parent = r; e = n; }
    return r;
  }
  uint32_t zero() {  // element standing for "node id 0" (:183)
    if (zero_elem == kNoElem) zero_elem = new_elem(kNoElem);
    return zero_elem;
  }
  std::string name_of(uint32_t e) const {
    if (el[e].name_off == kNameConst) return "const_signal_" + std::to_string(el[e].value);
    if (el[e].name_off == kNameNone) return "random_" + std::to_string(el[e].sig_id);
    return std::string(name_pool.c_str() + el[e].name_off);
  }
  void set_name(uint32_t e, const char* name) {
    el[e].name_off = (uint32_t)name_pool.size();
    name_pool.append(name);
    name_pool.push_back('\0');
  }

  int add_signal(uint32_t id, const char* name, int has_value, uint32_t value, int synth_const_name) {
    if (elem_of(id) != kNoElem) return C2A_ERR_SIGNAL_ALREADY_DECLARED;  // :146-148
    uint32_t e = new_elem(id);
    bind(id, e);
    Elem& x = el[e];
    if (name) set_name(e, name);
    else if (synth_const_name) x.name_off = kNameConst;
    x.has_value = has_value ? 1 : 0;
    x.value = value;
    if (has_value) const_signals.push_back(id);
    x.node_id = ++node_count;           // :157
    x.flags = has_value ? kConst : 0;   // :155
    x.head = x.tail = e;
    ++n_signals;
    return C2A_OK;
  }

  int add_gate(uint32_t op, uint32_t lhs, uint32_t rhs, uint32_t out) {
    if (op >= C2A_GATE_TYPE_COUNT) { err = "unsupported gate type: " + std::to_string(op); return C2A_ERR_INVALID_ARGUMENT; }
    uint32_t eo = elem_of(out);
    if (eo == kNoElem) { err = "add_gate: output signal " + std::to_string(out) + " is in no node (the reference panics at src/compiler.rs:201)"; return C2A_ERR_REFERENCE_PANIC; }
    uint32_t e_l = elem_of(lhs), e_r = elem_of(rhs);
    if (e_l == kNoElem) e_l = zero();
    if (e_r == kNoElem) e_r = zero();
    el[find(eo)].flags |= kOut;  // :201
    gates.push_back({op, e_l, e_r, eo});
    return C2A_OK;
  }

  int add_connection(uint32_t a, uint32_t b) {
    uint32_t ea = elem_of(a), eb = elem_of(b);
    if (ea == kNoElem && eb == kNoElem) return C2A_OK;  // both resolve to node 0 (:235-237)
    bool za = ea == kNoElem, zb = eb == kNoElem;
    if (za) ea = zero();
    if (zb) eb = zero();
    uint32_t ra = find(ea), rb = find(eb);
    if (ra == rb) return C2A_OK;  // :235-237
    Elem* x = el.data();
    if ((x[ra].flags & kOut) && (x[rb].flags & kOut)) return C2A_ERR_CANNOT_MERGE_OUTPUT_NODES;        // :239-241
    if ((x[ra].flags & kConst) && (x[rb].flags & kConst)) return C2A_ERR_CANNOT_MERGE_CONSTANT_NODES;  // :243-245
    // union by rank; the surviving root takes the merged node's data
    uint32_t root = ra, child = rb;
    if (x[ra].rank < x[rb].rank) { root = rb; child = ra; }
    else if (x[ra].rank == x[rb].rank) x[ra].rank++;
    uint8_t f = x[ra].flags | x[rb].flags;  // :251-252
    uint32_t h, t;                          // a's signals then b's (:254-255)
    if (x[ra].head == kNoElem) { h = x[rb].head; t = x[rb].tail; }
    else if (x[rb].head == kNoElem) { h = x[ra].head; t = x[ra].tail; }
    else { x[x[ra].tail].next = x[rb].head; h = x[ra].head; t = x[rb].tail; }
    x[child].parent = root;
    x[root].flags = f;
    x[root].head = h;
    x[root].tail = t;
    x[root].node_id = ++node_count;  // :257
    if (za || zb) zero_elem = kNoElem;  // gates that captured node 0 now follow the merged node (:260-270); later unknowns see a fresh 0
    return C2A_OK;
  }
};

extern "C" {

const char* c2a_gate_type_name(uint32_t op) { return op < C2A_GATE_TYPE_COUNT ? kGateNames[op] : nullptr; }
int c2a_gate_type_from_name(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < C2A_GATE_TYPE_COUNT; ++i)
    if (!strcmp(name, kGateNames[i])) return i;
  return -1;
}
const char* c2a_status_string(int st) {
  switch (st) {
    case C2A_OK: return "ok";
    case C2A_ERR_CYCLIC_DEPENDENCY: return "Cyclic dependency";               // compiler.rs:570 (+ ": {message}")
    case C2A_ERR_INCONSISTENCY: return "Inconsistency";                        // :572
    case C2A_ERR_SIGNAL_ALREADY_DECLARED: return "Signal already declared";   // :564
    case C2A_ERR_CANNOT_MERGE_OUTPUT_NODES: return "Cannot merge output nodes";      // :554
    case C2A_ERR_CANNOT_MERGE_CONSTANT_NODES: return "Cannot merge constant nodes";  // :552
    case C2A_ERR_REFERENCE_PANIC: return "reference panics here";
    case C2A_ERR_INVALID_ARGUMENT: return "invalid argument";
    case C2A_ERR_EVALUATION: return "evaluation failed";
    case C2A_ERR_CUDA: return "CUDA error";
    case C2A_ERR_NO_MEMORY: return "out of device memory";
  }
  return "unknown status";
}

c2a_compiler* c2a_compiler_new(void) { return new c2a_compiler(); }
void c2a_compiler_free(c2a_compiler* c) { delete c; }
const char* c2a_compiler_last_error(const c2a_compiler* c) { return c ? c->err.c_str() : "null compiler"; }

int c2a_add_signal(c2a_compiler* c, uint32_t id, const char* name, int has_value, uint32_t value) { return c->add_signal(id, name, has_value, value, 0); }
int c2a_add_gate(c2a_compiler* c, uint32_t op, uint32_t l, uint32_t r, uint32_t o) { return c->add_gate(op, l, r, o); }
int c2a_add_connection(c2a_compiler* c, uint32_t a, uint32_t b) { return c->add_connection(a, b); }

int c2a_emit_events(c2a_compiler* c, const c2a_event* ev, uint64_t n, uint64_t* err_event) {
  {  // size the tables once: one streaming pass over the kinds
    uint64_t ns = 0, ng = 0;
    for (uint64_t i = 0; i < n; ++i) { uint32_t k = ev[i].kind & 0xFF; ns += k <= C2A_EV_SIGNAL_CONST; ng += k == C2A_EV_GATE; }
    c->el.reserve(c->el.size() + ns + 1);
    c->dense.reserve(c->dense.size() + ns + 1);
    c->gates.reserve(c->gates.size() + ng);
  }
  for (uint64_t i = 0; i < n; ++i) {
    int st;
    switch (ev[i].kind & 0xFF) {
      case C2A_EV_SIGNAL: st = c->add_signal(ev[i].a, nullptr, 0, 0, 0); break;
      case C2A_EV_SIGNAL_CONST: st = c->add_signal(ev[i].a, nullptr, 1, ev[i].b, 1); break;
      case C2A_EV_GATE: st = c->add_gate(ev[i].kind >> 8, ev[i].a, ev[i].b, ev[i].c); break;
      case C2A_EV_CONNECT: st = c->add_connection(ev[i].a, ev[i].b); break;
      default: st = C2A_ERR_INVALID_ARGUMENT;
    }
    if (st) { if (err_event) *err_event = i; return st; }
  }
  return C2A_OK;
}

// ---- packed event stream (include/c2a.h): kinds byte + payload words -----------------------------------------------
uint64_t c2a_pack_events_ex(const c2a_event* ev, uint64_t n, uint32_t allow_flags, uint8_t* kinds_out, uint32_t* words_out, uint32_t* flags_out) {
  // dense = the declared ids are exactly 0, 1, 2, ... in declaration order (Runtime::gen_signal, src/runtime.rs:120-125)
  bool dense = true;
  uint64_t ns = 0, ng = 0, nc = 0, ni = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t k = ev[i].kind & 0xFF;
    if (k <= C2A_EV_SIGNAL_CONST) { if (ev[i].a != ns) dense = false; ++ns; }
    else if (k == C2A_EV_GATE) ++ng;
    else ++nc;  // CONNECT, or an invalid kind (kept as kind 3 with op bits set below so that the device flags it)
  }
  // implicit operands (dense ids only): a gate whose out signal / a connection whose first signal is the signal declared last
  // (what the walker always emits: src/process.rs:470-475 creates the temporary right before the gate, :241-273 connects it)
  const bool implicit = dense && (allow_flags & C2A_PACKED_IMPLICIT_OPERANDS);
  if (implicit) {
    uint64_t s_before = 0;
    for (uint64_t i = 0; i < n; ++i) {
      uint32_t k = ev[i].kind & 0xFF;
      if (k <= C2A_EV_SIGNAL_CONST) ++s_before;
      else if (k == C2A_EV_GATE) { if (s_before && ev[i].c == s_before - 1 && (ev[i].kind >> 8) < 31) ++ni; }
      else if (k == C2A_EV_CONNECT) { if (s_before && ev[i].a == s_before - 1) ++ni; }
    }
  }
  const uint64_t n_words = 3 * ng + 2 * nc - ni + (dense ? 0 : ns);
  if (flags_out) *flags_out = (dense ? C2A_PACKED_DENSE_IDS : 0u) | (implicit ? C2A_PACKED_IMPLICIT_OPERANDS : 0u);
  if (!kinds_out || !words_out) return n_words;
  uint64_t w = 0, s_before = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t k = ev[i].kind & 0xFF;
    if (k <= C2A_EV_SIGNAL_CONST) {
      kinds_out[i] = (uint8_t)k;
      if (!dense) words_out[w++] = ev[i].a;
      ++s_before;
    } else if (k == C2A_EV_GATE) {
      uint32_t op = ev[i].kind >> 8;
      if (implicit) {  // op in bits 2..6, bit 7 = "out is the signal declared last"
        const bool im = s_before && ev[i].c == s_before - 1 && op < 31;
        kinds_out[i] = (uint8_t)(C2A_EV_GATE | (op < 31 ? op << 2 : 31u << 2) | (im ? 0x80u : 0u));
        words_out[w++] = ev[i].a; words_out[w++] = ev[i].b;
        if (!im) words_out[w++] = ev[i].c;
      } else {
        kinds_out[i] = (uint8_t)(C2A_EV_GATE | (op < 63 ? op << 2 : 63u << 2));  // an out-of-range op stays out of range
        words_out[w++] = ev[i].a; words_out[w++] = ev[i].b; words_out[w++] = ev[i].c;
      }
    } else {
      const bool valid = k == C2A_EV_CONNECT;
      const bool im = implicit && valid && s_before && ev[i].a == s_before - 1;
      kinds_out[i] = (uint8_t)(C2A_EV_CONNECT | (valid ? 0u : 1u << 2) | (im ? 0x80u : 0u));  // op bits on a non-gate = invalid kind
      if (!im) words_out[w++] = ev[i].a;
      words_out[w++] = ev[i].b;
    }
  }
  return n_words;
}

uint64_t c2a_pack_events(const c2a_event* ev, uint64_t n, uint8_t* kinds_out, uint32_t* words_out, uint32_t* flags_out) {
  return c2a_pack_events_ex(ev, n, C2A_PACKED_DENSE_IDS, kinds_out, words_out, flags_out);
}

int c2a_unpack_events(const c2a_packed_events* pk, c2a_event* out) {
  if (!pk || (pk->n_events && (!pk->kinds || !out))) return C2A_ERR_INVALID_ARGUMENT;
  const bool dense = pk->flags & C2A_PACKED_DENSE_IDS;
  const bool implicit = dense && (pk->flags & C2A_PACKED_IMPLICIT_OPERANDS);
  uint64_t w = 0, ns = 0;
  for (uint64_t i = 0; i < pk->n_events; ++i) {
    uint32_t kb = pk->kinds[i], k = kb & 3u, op = implicit ? (kb >> 2) & 31u : kb >> 2;
    const bool im = implicit && (kb & 0x80u);
    uint32_t need = k <= C2A_EV_SIGNAL_CONST ? (dense ? 0u : 1u) : (k == C2A_EV_GATE ? 3u : 2u) - (im ? 1u : 0u);
    if (w + need > pk->n_words || (need && !pk->words)) return C2A_ERR_INVALID_ARGUMENT;
    const uint32_t last = (uint32_t)ns - 1u;  // the signal declared last (0xFFFFFFFF when there is none: an undeclared reference)
    if (k <= C2A_EV_SIGNAL_CONST) {
      out[i] = c2a_event{(op || im) ? 0xFFu : k, dense ? (uint32_t)ns : pk->words[w], 0, 0};
      ++ns;
    } else if (k == C2A_EV_GATE) {
      out[i] = c2a_event{(uint32_t)C2A_EV_GATE | (op << 8), pk->words[w], pk->words[w + 1], im ? last : pk->words[w + 2]};
    } else {
      out[i] = im ? c2a_event{op ? 0xFFu : (uint32_t)C2A_EV_CONNECT, last, pk->words[w], 0}
                  : c2a_event{op ? 0xFFu : (uint32_t)C2A_EV_CONNECT, pk->words[w], pk->words[w + 1], 0};
    }
    w += need;
  }
  return w == pk->n_words ? C2A_OK : C2A_ERR_INVALID_ARGUMENT;
}

int c2a_set_signal_name(c2a_compiler* c, uint32_t id, const char* name) {
  uint32_t e = c->elem_of(id);
  if (e == kNoElem || !name) return C2A_ERR_INVALID_ARGUMENT;
  c->set_name(e, name);
  return C2A_OK;
}

int64_t c2a_signal_name(c2a_compiler* c, uint32_t id, char* buf, uint64_t cap) {
  uint32_t e = c->elem_of(id);
  if (e == kNoElem) return -1;
  std::string nm = c->name_of(e);
  if (buf && cap) { size_t k = std::min<size_t>(nm.size(), cap - 1); memcpy(buf, nm.data(), k); buf[k] = 0; }
  return (int64_t)nm.size();
}

int c2a_signal_value(c2a_compiler* c, uint32_t id, int* has_value, uint32_t* value) {
  uint32_t e = c->elem_of(id);
  if (e == kNoElem) return C2A_ERR_INVALID_ARGUMENT;
  if (has_value) *has_value = c->el[e].has_value;
  if (value) *value = c->el[e].value;
  return C2A_OK;
}

uint64_t c2a_get_signals_by_prefix(c2a_compiler* c, const char* prefix, uint32_t* ids_out, uint64_t cap) {
  std::vector<uint32_t> ids;
  size_t n = strlen(prefix);
  for (uint32_t e = 0; e < c->el.size(); ++e)
    if (c->el[e].sig_id != kNoElem && c->name_of(e).compare(0, n, prefix) == 0) ids.push_back(c->el[e].sig_id);
  std::sort(ids.begin(), ids.end());
  for (size_t i = 0; i < ids.size() && i < cap; ++i) ids_out[i] = ids[i];
  return ids.size();
}

int c2a_add_input(c2a_compiler* c, uint32_t id, const char* name) { c->inputs[id] = name ? name : ""; return C2A_OK; }
int c2a_add_output(c2a_compiler* c, uint32_t id, const char* name) { c->outputs[id] = name ? name : ""; return C2A_OK; }

// src/compiler.rs:163-171 + src/program.rs:57-66: every signal whose name starts with the prefix
static void tag_prefix(c2a_compiler* c, const char* prefix, bool input) {
  size_t n = strlen(prefix);
  for (uint32_t e = 0; e < c->el.size(); ++e) {
    if (c->el[e].sig_id == kNoElem) continue;
    std::string nm = c->name_of(e);
    if (nm.compare(0, n, prefix) == 0) (input ? c->inputs : c->outputs)[c->el[e].sig_id] = nm;
  }
}
int c2a_tag_inputs_by_prefix(c2a_compiler* c, const char* p) { if (!p) return C2A_ERR_INVALID_ARGUMENT; tag_prefix(c, p, true); return C2A_OK; }
int c2a_tag_outputs_by_prefix(c2a_compiler* c, const char* p) { if (!p) return C2A_ERR_INVALID_ARGUMENT; tag_prefix(c, p, false); return C2A_OK; }

uint64_t c2a_num_gates(const c2a_compiler* c) { return c->gates.size(); }
uint32_t c2a_node_count(const c2a_compiler* c) { return c->node_count; }
uint64_t c2a_num_signals(const c2a_compiler* c) { return c->n_signals; }

int c2a_get_gates(c2a_compiler* c, c2a_gate* out) {
  const size_t n = c->gates.size();
  for (size_t i = 0; i < n; ++i) {
    const ElemGate& g = c->gates[i];
    out[i].op = g.op;
    out[i].lh = c->el[c->find(g.l)].node_id;
    out[i].rh = c->el[c->find(g.r)].node_id;
    out[i].out = c->el[c->find(g.o)].node_id;
  }
  return C2A_OK;
}

int c2a_signal_node(c2a_compiler* c, uint32_t sid, uint32_t* node) {
  uint32_t e = c->elem_of(sid);
  *node = e == kNoElem ? 0 : c->el[c->find(e)].node_id;
  return C2A_OK;
}

int c2a_signal_nodes(c2a_compiler* c, const uint32_t* sids, uint64_t n, uint32_t* nodes) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t e = c->elem_of(sids[i]);
    nodes[i] = e == kNoElem ? 0 : c->el[c->find(e)].node_id;
  }
  return C2A_OK;
}

static void live_roots(c2a_compiler* c, std::vector<std::pair<uint32_t, uint32_t>>* out) {  // (node id, root)
  for (uint32_t e = 0; e < c->el.size(); ++e)
    if (c->el[e].parent == e && c->el[e].node_id != 0) out->push_back({c->el[e].node_id, e});
  std::sort(out->begin(), out->end());
}
uint64_t c2a_num_nodes(c2a_compiler* c) {
  std::vector<std::pair<uint32_t, uint32_t>> v;
  live_roots(c, &v);
  return v.size();
}
int c2a_get_nodes(c2a_compiler* c, uint32_t* ids, uint8_t* flags, uint64_t* sig_off, uint32_t* sig) {
  std::vector<std::pair<uint32_t, uint32_t>> v;
  live_roots(c, &v);
  uint64_t off = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    uint32_t r = v[i].second;
    if (ids) ids[i] = v[i].first;
    if (flags) flags[i] = c->el[r].flags;
    if (sig_off) sig_off[i] = off;
    for (uint32_t e = c->el[r].head; e != kNoElem; e = c->el[e].next) {
      if (sig) sig[off] = c->el[e].sig_id;
      ++off;
    }
  }
  if (sig_off) sig_off[v.size()] = off;
  return C2A_OK;
}

static std::string jesc(const std::string& s) {
  std::string o;
  for (char ch : s) { if (ch == '"' || ch == '\\') o += '\\'; o += ch; }
  return o;
}

// Compiler::generate_circuit_report, src/compiler.rs:287-319 + get_node_report :503-531.  inputs = nodes no gate writes,
// outputs = written nodes no gate reads, both ascending by node id; per node: the names of its signals in merge order without
// the temporaries ("random_"), and the value of its last constant signal.  One pass over the gates and one over the nodes.
const char* c2a_circuit_report_json(c2a_compiler* c, const char* value_type) {
  if (!c) return nullptr;
  std::vector<std::pair<uint32_t, uint32_t>> v;
  live_roots(c, &v);
  std::vector<uint8_t> consumed((size_t)c->node_count + 1, 0);
  for (const ElemGate& g : c->gates) {
    consumed[c->el[c->find(g.l)].node_id] = 1;
    consumed[c->el[c->find(g.r)].node_id] = 1;
  }
  std::string& o = c->report_json;
  o.clear();
  auto node = [&](uint32_t id, uint32_t r, bool first) {
    if (!first) o += ",";
    o += "{\"id\":" + std::to_string(id) + ",\"names\":[";
    bool first_name = true, has = false;
    uint32_t value = 0;
    for (uint32_t e = c->el[r].head; e != kNoElem; e = c->el[e].next) {
      if (c->el[e].name_off != kNameNone) {
        std::string nm = c->name_of(e);
        if (nm.find("random_") == std::string::npos) {
          if (!first_name) o += ",";
          o += "\"" + jesc(nm) + "\"";
          first_name = false;
        }
      }
      if (c->el[e].has_value) { has = true; value = c->el[e].value; }
    }
    o += "],\"value\":" + (has ? std::to_string(value) : std::string("null")) + "}";
  };
  o += "{\"inputs\":[";
  bool first = true;
  for (auto& kv : v)
    if (!(c->el[kv.second].flags & kOut)) { node(kv.first, kv.second, first); first = false; }
  o += "],\"outputs\":[";
  first = true;
  for (auto& kv : v)
    if ((c->el[kv.second].flags & kOut) && !consumed[kv.first]) { node(kv.first, kv.second, first); first = false; }
  o += "],\"value_type\":\"" + jesc(value_type ? value_type : "sint") + "\"}";
  return o.c_str();
}

// The gate lines of bristol-circuit's write_bristol ("2 1 <in0> <in1> <out> <Op>\n" per gate; crate un-vendored, layout PARITY
// UNPINNED - see include/c2a.h).  Returns the number of bytes the lines take; writes them when out != NULL and cap suffices.
uint64_t c2a_bristol_gate_lines(const c2a_gate* gates, uint64_t G, char* out, uint64_t cap) {
  auto put_u32 = [](char* p, uint32_t v) -> char* {
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
  };
  auto digits = [](uint32_t v) -> uint64_t { uint64_t d = 1; while (v >= 10) { v /= 10; ++d; } return d; };
  size_t name_len[C2A_GATE_TYPE_COUNT];
  for (int i = 0; i < C2A_GATE_TYPE_COUNT; ++i) name_len[i] = strlen(kGateNames[i]);
  uint64_t need = 0;
  for (uint64_t i = 0; i < G; ++i) {
    if (gates[i].op >= C2A_GATE_TYPE_COUNT) return 0;
    need += 4 + digits(gates[i].lh) + 1 + digits(gates[i].rh) + 1 + digits(gates[i].out) + 1 + name_len[gates[i].op] + 1;
  }
  if (!out || cap < need) return need;
  char* p = out;
  for (uint64_t i = 0; i < G; ++i) {
    const c2a_gate& g = gates[i];
    memcpy(p, "2 1 ", 4); p += 4;
    p = put_u32(p, g.lh); *p++ = ' ';
    p = put_u32(p, g.rh); *p++ = ' ';
    p = put_u32(p, g.out); *p++ = ' ';
    memcpy(p, kGateNames[g.op], name_len[g.op]); p += name_len[g.op];
    *p++ = '\n';
  }
  return need;
}

// Compiler::build_circuit, src/compiler.rs:321-494
int c2a_compiler_build_circuit(c2a_compiler* c, c2a_handle* h) {
  if (!c) return C2A_ERR_INVALID_ARGUMENT;
  if (!h) { c->err = "no device handle: the back end has no CPU implementation"; return C2A_ERR_CUDA; }
  // :327-361  IO <=> node, walked in ascending signal id
  std::vector<std::pair<std::string, uint32_t>> input_to_node, output_to_node;
  std::set<std::string> in_names, out_names;
  {
    auto ii = c->inputs.begin(), oi = c->outputs.begin();
    while (ii != c->inputs.end() || oi != c->outputs.end()) {
      uint32_t sid;
      if (oi == c->outputs.end() || (ii != c->inputs.end() && ii->first <= oi->first)) sid = ii->first; else sid = oi->first;
      uint32_t e = c->elem_of(sid);
      bool is_in = ii != c->inputs.end() && ii->first == sid, is_out = oi != c->outputs.end() && oi->first == sid;
      if (e != kNoElem) {  // signals that are in no node are never reached by the reference's node walk
        uint32_t node = c->el[c->find(e)].node_id;
        if (is_in) {
          if (!in_names.insert(ii->second).second) { c->err = "Duplicate input " + ii->second; return C2A_ERR_INCONSISTENCY; }  // :337-341
          input_to_node.push_back({ii->second, node});
        }
        if (is_out) {
          if (!out_names.insert(oi->second).second) { c->err = "Duplicate output " + oi->second; return C2A_ERR_INCONSISTENCY; }  // :347-351
          output_to_node.push_back({oi->second, node});
        }
      }
      if (is_in) ++ii;
      if (is_out) ++oi;
    }
  }
  {  // :363-383
    std::map<uint32_t, std::string> node_to_input;
    for (auto& p : input_to_node) node_to_input[p.second] = p.first;
    for (auto& p : output_to_node) {
      auto f = node_to_input.find(p.second);
      if (f != node_to_input.end()) {
        c->err = "Node " + std::to_string(p.second) + " used for both input " + f->second + " and output " + p.first;
        return C2A_ERR_INCONSISTENCY;
      }
    }
  }
  std::vector<uint32_t> in_nodes, out_nodes;
  for (auto& p : input_to_node) in_nodes.push_back(p.second);
  for (auto& p : output_to_node) out_nodes.push_back(p.second);

  // :385-464 on the device
  const uint64_t G = c->gates.size();
  const uint32_t node_bound = c->node_count + 1;
  std::vector<c2a_gate> gates(G);
  c2a_get_gates(c, gates.data());
  c->order.assign(G, 0);
  c->new_gates.assign(G, c2a_gate{0, 0, 0, 0});
  std::vector<uint32_t> wire(node_bound, C2A_NONE);
  uint32_t wire_count = 0;
  uint64_t err_index = 0;
  int st = c2a_build_circuit(h, gates.data(), G, node_bound, in_nodes.data(), (uint32_t)in_nodes.size(), out_nodes.data(),
                             (uint32_t)out_nodes.size(), c->order.data(), wire.data(), c->new_gates.data(), &wire_count, &err_index);
  if (st == C2A_ERR_CYCLIC_DEPENDENCY) { c->err = "detected at i=" + std::to_string(err_index); return st; }
  if (st != C2A_OK) { c->err = c2a_last_error(h); return st; }
  c->wire_count = wire_count;

  // :466-493
  std::string j = "{\"input_name_to_wire_index\":{";
  {
    std::map<std::string, uint32_t> m;
    for (auto& p : input_to_node) m[p.first] = wire[p.second];
    bool first = true;
    for (auto& kv : m) { j += first ? "" : ","; j += "\"" + jesc(kv.first) + "\":" + std::to_string(kv.second); first = false; }
  }
  j += "},\"constants\":{";
  {
    std::map<std::string, std::pair<uint32_t, std::string>> consts;  // key "<name>_<signal_id>" (:356)
    std::vector<uint32_t> ids = c->const_signals;
    std::sort(ids.begin(), ids.end());
    for (uint32_t sid : ids) {
      uint32_t e = c->elem_of(sid);
      consts[c->name_of(e) + "_" + std::to_string(sid)] = {c->el[c->find(e)].node_id, std::to_string(c->el[e].value)};
    }
    bool first = true;
    for (auto& kv : consts) {
      uint32_t w = wire[kv.second.first];
      if (w == C2A_NONE) { c->err = "constant " + kv.first + " has no wire (the reference panics at src/compiler.rs:473)"; return C2A_ERR_REFERENCE_PANIC; }
      j += first ? "" : ",";
      j += "\"" + jesc(kv.first) + "\":{\"value\":\"" + kv.second.second + "\",\"wire_index\":" + std::to_string(w) + "}";
      first = false;
    }
  }
  j += "},\"output_name_to_wire_index\":{";
  {
    std::map<std::string, uint32_t> m;
    for (auto& p : output_to_node) m[p.first] = wire[p.second];
    bool first = true;
    for (auto& kv : m) { j += first ? "" : ","; j += "\"" + jesc(kv.first) + "\":" + std::to_string(kv.second); first = false; }
  }
  j += "}}";
  c->info_json = j;
  return C2A_OK;
}

uint64_t c2a_circuit_wire_count(const c2a_compiler* c) { return c->wire_count; }
const uint32_t* c2a_circuit_order(const c2a_compiler* c) { return c->order.data(); }
const c2a_gate* c2a_circuit_gates(const c2a_compiler* c) { return c->new_gates.data(); }
const char* c2a_circuit_info_json(const c2a_compiler* c) { return c->info_json.c_str(); }

}  // extern "C"
