// =====================================================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the circom-2-arithc flattening hot path, written to be *faithful* to the reference's
// data-structure shapes (hash maps keyed by node id, linear scans, one dependency vector per gate, DFS with
// roots ascending), not fast.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (circom-2-arithc_b200/) never does.
//
// Parity pin status: the Rust reference cannot be built in this environment (no cargo/rustc, un-vendored
// git dependencies), so this restatement is pinned by the reference's OWN tests restated in
// tests/test_oracle_goldens.py: src/compiler.rs:584-795 (node ids, merge semantics, merge errors),
// tests/integration.rs:279-441 (exact constants / output maps, functional simulations),
// src/process.rs:772-822 (execute_op KATs).  src/topological_sort.rs has no test upstream; its order is
// pinned by restating the 50-line function verbatim (recursion made explicit) and by the simulations.
//
// Every function cites the reference lines it follows (paths relative to the reference tree).
// =====================================================================================================
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <optional>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace orc {

// status codes: numerically identical to include/c2a.h (compared by the tests)
enum : int {
  OK = 0,
  E_CYCLIC = 1,
  E_INCONSISTENCY = 2,
  E_SIGNAL_ALREADY_DECLARED = 3,
  E_MERGE_OUT = 4,
  E_MERGE_CONST = 5,
  E_PANIC = 6,
  E_INVALID = 7,
};

static const char* kGateNames[20] = {  // src/a_gate_type.rs:7-28, declaration order
    "AAdd", "ADiv",    "AEq",  "AGEq",    "AGt",     "ALEq",    "ALt",      "AMul",   "ANeq",   "ASub",
    "AXor", "APow",    "AIntDiv", "AMod", "AShiftL", "AShiftR", "ABoolOr",  "ABoolAnd", "ABitOr", "ABitAnd"};

struct Signal {  // src/compiler.rs:17-27
  std::string name;
  std::optional<uint32_t> value;
};

struct Node {  // src/compiler.rs:32-81
  bool is_const = false;
  bool is_out = false;
  std::vector<uint32_t> signals;
  bool contains_signal(uint32_t id) const {  // :63-65  (Vec::contains, linear)
    return std::find(signals.begin(), signals.end(), id) != signals.end();
  }
};

struct Gate {  // src/compiler.rs:85-90
  uint32_t op, lh_in, rh_in, out;
};

// ---- src/topological_sort.rs:3-50 ---------------------------------------------------------------
// Same traversal, recursion replaced by an explicit frame stack (the reference recurses and would overflow
// the thread stack on ~1e5-deep chains; order of visits and of `sorted.push` is unchanged).
// Returns OK or E_CYCLIC with *cycle_at = the `i` of "detected at i={}" (:34-38).
static int topological_sort(size_t len, const std::function<std::vector<size_t>(size_t)>& get_deps,
                            std::vector<size_t>& sorted, size_t* cycle_at) {
  sorted.clear();
  sorted.reserve(len);                    // :7
  std::vector<char> visiting(len, 0);     // :8
  std::vector<char> visited(len, 0);      // :9
  struct Frame {
    size_t i;
    std::vector<size_t> deps;
    size_t next;
  };
  std::vector<Frame> stack;
  for (size_t root = 0; root < len; ++root) {  // :11-13
    // topological_sort_visit(root, ...)
    if (visited[root]) continue;               // :30-32
    if (visiting[root]) {                      // :34-38 (cannot trigger at the root, kept for symmetry)
      *cycle_at = root;
      return E_CYCLIC;
    }
    visiting[root] = 1;                        // :40
    stack.push_back({root, get_deps(root), 0});
    while (!stack.empty()) {
      Frame& f = stack.back();
      if (f.next < f.deps.size()) {            // :42-44  for j in get_deps(i) { visit(j) }
        size_t j = f.deps[f.next++];
        if (visited[j]) continue;              // :30-32
        if (visiting[j]) {                     // :34-38
          *cycle_at = j;
          return E_CYCLIC;
        }
        visiting[j] = 1;                       // :40
        stack.push_back({j, get_deps(j), 0});  // get_deps is invoked exactly once per item, on entry
      } else {
        sorted.push_back(f.i);                 // :46
        visited[f.i] = 1;                      // :47
        stack.pop_back();
      }
    }
  }
  if (sorted.size() != len) return E_INVALID;  // :15-18 assert
  return OK;
}

// ---- back-end core: src/compiler.rs:385-464 on plain arrays -------------------------------------------
struct BackendResult {
  std::vector<size_t> order;
  std::unordered_map<uint32_t, uint32_t> node_id_to_wire_id;
  std::vector<Gate> new_gates;  // op kept numeric; wire ids
  uint32_t wire_count = 0;
  size_t cycle_at = 0;
};

static int backend_core(const std::vector<Gate>& gates, const std::vector<uint32_t>& input_nodes,
                        const std::vector<uint32_t>& output_nodes, BackendResult& r) {
  auto& node_id_to_wire_id = r.node_id_to_wire_id;  // :388
  uint32_t next_wire_id = 0;                        // :389
  for (uint32_t node_id : input_nodes) {            // :392-395 (iteration order supplied by the caller)
    node_id_to_wire_id[node_id] = next_wire_id;     // HashMap::insert overwrites
    next_wire_id += 1;
  }
  std::unordered_map<uint32_t, size_t> node_id_to_required_gate;  // :401
  for (size_t gate_id = 0; gate_id < gates.size(); ++gate_id)     // :403-406 (insert overwrites: last wins)
    node_id_to_required_gate[gates[gate_id].out] = gate_id;

  int st = topological_sort(  // :408-421
      gates.size(),
      [&](size_t gate_id) {
        const Gate& gate = gates[gate_id];
        std::vector<size_t> deps;
        auto l = node_id_to_required_gate.find(gate.lh_in);
        if (l != node_id_to_required_gate.end()) deps.push_back(l->second);
        auto rr = node_id_to_required_gate.find(gate.rh_in);
        if (rr != node_id_to_required_gate.end()) deps.push_back(rr->second);
        return deps;
      },
      r.order, &r.cycle_at);
  if (st != OK) return st;

  std::unordered_set<uint32_t> output_node_ids(output_nodes.begin(), output_nodes.end());  // :423
  for (size_t gate_id : r.order) {  // :427-443
    const Gate& gate = gates[gate_id];
    const uint32_t ids[3] = {gate.lh_in, gate.rh_in, gate.out};
    for (uint32_t node_id : ids) {
      if (output_node_ids.count(node_id)) continue;     // :431-434
      if (node_id_to_wire_id.count(node_id)) continue;  // :436-438
      node_id_to_wire_id[node_id] = next_wire_id;       // :440-441
      next_wire_id += 1;
    }
  }
  for (uint32_t node_id : output_nodes) {  // :446-449
    node_id_to_wire_id[node_id] = next_wire_id;
    next_wire_id += 1;
  }
  r.new_gates.clear();
  r.new_gates.reserve(gates.size());
  for (size_t gate_id : r.order) {  // :452-464
    const Gate& gate = gates[gate_id];
    r.new_gates.push_back({gate.op, node_id_to_wire_id.at(gate.lh_in), node_id_to_wire_id.at(gate.rh_in),
                           node_id_to_wire_id.at(gate.out)});
  }
  r.wire_count = next_wire_id;  // :479
  return OK;
}

// ---- src/compiler.rs:107-284 ----------------------------------------------------------------------------
struct Compiler {
  uint32_t node_count = 0;
  std::map<uint32_t, std::string> inputs;   // signal id -> name   (HashMap upstream; ordered here, see build)
  std::map<uint32_t, std::string> outputs;
  std::unordered_map<uint32_t, Signal> signals;
  std::unordered_map<uint32_t, Node> nodes;
  std::vector<Gate> gates;
  std::string last_error;

  // build results
  BackendResult built;
  std::string info_json;
  std::vector<uint32_t> built_input_nodes, built_output_nodes;

  uint32_t get_node_id() {  // :497-500
    node_count += 1;
    return node_count;
  }

  int add_signal(uint32_t id, const std::string& name, std::optional<uint32_t> value) {  // :139-161
    if (signals.count(id)) return E_SIGNAL_ALREADY_DECLARED;                             // :146-148
    signals[id] = Signal{name, value};                                                   // :151-152
    Node node;                                                                           // :155
    node.signals = {id};
    node.is_const = value.has_value();
    node.is_out = false;
    uint32_t node_id = get_node_id();  // :157
    nodes[node_id] = std::move(node);  // :158
    return OK;
  }

  int add_gate(uint32_t op, uint32_t lhs, uint32_t rhs, uint32_t out) {  // :174-209
    uint32_t ids[3] = {0, 0, 0};                                         // :183
    for (auto& kv : nodes) {                                             // :185-195  scan every node
      if (kv.second.contains_signal(lhs)) ids[0] = kv.first;
      if (kv.second.contains_signal(rhs)) ids[1] = kv.first;
      if (kv.second.contains_signal(out)) ids[2] = kv.first;
    }
    auto it = nodes.find(ids[2]);
    if (it == nodes.end()) return E_PANIC;  // :201 `.unwrap()` on None
    it->second.is_out = true;               // :201
    gates.push_back({op, ids[0], ids[1], ids[2]});  // :204-206
    return OK;
  }

  int add_connection(uint32_t a, uint32_t b) {  // :213-278
    static const Node empty;                    // :215 `let n = Node::new();`
    uint32_t node_a_id = 0, node_b_id = 0;      // :217
    const Node* node_a = &empty;
    const Node* node_b = &empty;
    for (auto& kv : nodes) {  // :219-226
      if (kv.second.contains_signal(a)) { node_a_id = kv.first; node_a = &kv.second; }
      if (kv.second.contains_signal(b)) { node_b_id = kv.first; node_b = &kv.second; }
    }
    if (node_a_id == node_b_id) return OK;                         // :235-237
    if (node_a->is_out && node_b->is_out) return E_MERGE_OUT;      // :239-241
    if (node_a->is_const && node_b->is_const) return E_MERGE_CONST;  // :243-245
    Node merged;                                                   // :248
    merged.is_out = node_a->is_out || node_b->is_out;              // :251
    merged.is_const = node_a->is_const || node_b->is_const;        // :252
    merged.signals = node_a->signals;                              // :254
    merged.signals.insert(merged.signals.end(), node_b->signals.begin(), node_b->signals.end());  // :255
    uint32_t merged_id = get_node_id();                            // :257
    for (Gate& g : gates) {                                        // :260-270  rewrite every gate
      if (g.lh_in == node_a_id || g.lh_in == node_b_id) g.lh_in = merged_id;
      if (g.rh_in == node_a_id || g.rh_in == node_b_id) g.rh_in = merged_id;
      if (g.out == node_a_id || g.out == node_b_id) g.out = merged_id;
    }
    nodes.erase(node_a_id);  // :273-275
    nodes.erase(node_b_id);
    nodes[merged_id] = std::move(merged);
    return OK;
  }

  // src/compiler.rs:163-171 + src/program.rs:57-66: tag every signal whose NAME starts with `prefix`
  void tag_by_prefix(const std::string& prefix, bool input) {
    for (auto& kv : signals)
      if (kv.second.name.compare(0, prefix.size(), prefix) == 0) (input ? inputs : outputs)[kv.first] = kv.second.name;
  }

  static std::string jesc(const std::string& s) {
    std::string o;
    for (char c : s) {
      if (c == '"' || c == '\\') { o += '\\'; o += c; }
      else o += c;
    }
    return o;
  }

  // src/compiler.rs:321-494.  HashMap iteration order is unspecified upstream; the deterministic stand-in
  // used here (and by the product) is ascending signal id for inputs/outputs and ascending node id for the
  // node walk.  Everything else is exactly the reference's sequence.
  int build_circuit() {
    std::vector<std::pair<std::string, uint32_t>> input_to_node_id, output_to_node_id;  // insertion-ordered
    std::set<std::string> in_names, out_names;
    std::map<std::string, std::pair<uint32_t, std::string>> constant_to_node_id_and_value;
    // signal id -> node id (the reference walks nodes -> signals; same relation)
    std::vector<uint32_t> node_ids;
    for (auto& kv : nodes) node_ids.push_back(kv.first);
    std::sort(node_ids.begin(), node_ids.end());
    std::map<uint32_t, uint32_t> sig_node;
    for (uint32_t nid : node_ids)
      for (uint32_t sid : nodes[nid].signals) sig_node[sid] = nid;
    for (auto& kv : sig_node) {  // :327-361, ascending signal id
      uint32_t signal_id = kv.first, node_id = kv.second;
      auto in = inputs.find(signal_id);
      if (in != inputs.end()) {
        if (!in_names.insert(in->second).second) {  // :337-341
          last_error = "Duplicate input " + in->second;
          return E_INCONSISTENCY;
        }
        input_to_node_id.push_back({in->second, node_id});
      }
      auto out = outputs.find(signal_id);
      if (out != outputs.end()) {
        if (!out_names.insert(out->second).second) {  // :347-351
          last_error = "Duplicate output " + out->second;
          return E_INCONSISTENCY;
        }
        output_to_node_id.push_back({out->second, node_id});
      }
      const Signal& s = signals[signal_id];
      if (s.value.has_value())  // :354-359  key "<name>_<signal_id>"
        constant_to_node_id_and_value[s.name + "_" + std::to_string(signal_id)] = {node_id, std::to_string(*s.value)};
    }
    {  // :363-383
      std::map<uint32_t, std::string> node_id_to_input_name;
      for (auto& p : input_to_node_id) node_id_to_input_name[p.second] = p.first;
      for (auto& p : output_to_node_id) {
        auto f = node_id_to_input_name.find(p.second);
        if (f != node_id_to_input_name.end()) {
          last_error = "Node " + std::to_string(p.second) + " used for both input " + f->second + " and output " + p.first;
          return E_INCONSISTENCY;
        }
      }
    }
    built_input_nodes.clear();
    built_output_nodes.clear();
    for (auto& p : input_to_node_id) built_input_nodes.push_back(p.second);
    for (auto& p : output_to_node_id) built_output_nodes.push_back(p.second);
    built = BackendResult();
    int st = backend_core(gates, built_input_nodes, built_output_nodes, built);  // :385-464
    if (st == E_CYCLIC) last_error = "detected at i=" + std::to_string(built.cycle_at);
    if (st != OK) return st;
    // :466-493
    std::string j = "{\"input_name_to_wire_index\":{";
    {
      std::map<std::string, uint32_t> m;
      for (auto& p : input_to_node_id) m[p.first] = built.node_id_to_wire_id.at(p.second);
      bool first = true;
      for (auto& kv : m) { j += (first ? "" : ","); j += "\"" + jesc(kv.first) + "\":" + std::to_string(kv.second); first = false; }
    }
    j += "},\"constants\":{";
    {
      bool first = true;
      for (auto& kv : constant_to_node_id_and_value) {
        auto w = built.node_id_to_wire_id.find(kv.second.first);
        if (w == built.node_id_to_wire_id.end()) {  // :473 index on a missing key panics
          last_error = "constant " + kv.first + " has no wire";
          return E_PANIC;
        }
        j += (first ? "" : ",");
        j += "\"" + jesc(kv.first) + "\":{\"value\":\"" + kv.second.second + "\",\"wire_index\":" + std::to_string(w->second) + "}";
        first = false;
      }
    }
    j += "},\"output_name_to_wire_index\":{";
    {
      std::map<std::string, uint32_t> m;
      for (auto& p : output_to_node_id) m[p.first] = built.node_id_to_wire_id.at(p.second);
      bool first = true;
      for (auto& kv : m) { j += (first ? "" : ","); j += "\"" + jesc(kv.first) + "\":" + std::to_string(kv.second); first = false; }
    }
    j += "}}";
    info_json = j;
    return OK;
  }

  // src/compiler.rs:287-319, 503-531: report.json content
  std::string report_json(const std::string& value_type) {
    std::vector<uint32_t> input_nodes, output_nodes;
    for (auto& kv : nodes) (kv.second.is_out ? output_nodes : input_nodes).push_back(kv.first);  // :291-297
    output_nodes.erase(std::remove_if(output_nodes.begin(), output_nodes.end(),
                                      [&](uint32_t id) {  // :300-304
                                        for (auto& g : gates)
                                          if (g.lh_in == id || g.rh_in == id) return true;
                                        return false;
                                      }),
                       output_nodes.end());
    std::sort(input_nodes.begin(), input_nodes.end());   // :307-308
    std::sort(output_nodes.begin(), output_nodes.end());
    auto reports = [&](const std::vector<uint32_t>& ids) {  // :503-531
      std::string o = "[";
      bool first = true;
      for (uint32_t id : ids) {
        const Node& n = nodes[id];
        std::string names;
        std::optional<uint32_t> value;
        bool nf = true;
        for (uint32_t sid : n.signals) {
          const Signal& s = signals[sid];
          if (s.name.find("random_") == std::string::npos) { names += (nf ? "" : ","); names += "\"" + jesc(s.name) + "\""; nf = false; }  // :519
          if (s.value.has_value()) value = s.value;  // :522-524
        }
        o += (first ? "" : ",");
        o += "{\"id\":" + std::to_string(id) + ",\"names\":[" + names + "],\"value\":" + (value ? std::to_string(*value) : std::string("null")) + "}";
        first = false;
      }
      return o + "]";
    };
    return "{\"inputs\":" + reports(input_nodes) + ",\"outputs\":" + reports(output_nodes) + ",\"value_type\":\"" + value_type + "\"}";
  }
};

// ---- src/process.rs:649-750 execute_op (u32; Rust debug-build overflow panics are reported as errors) ----
// returns 0 ok, 1 OperationError (message in *msg), 2 arithmetic overflow (Rust panics in debug, wraps in release)
static int execute_op(uint32_t lhs, uint32_t rhs, uint32_t op, uint32_t* res, const char** msg) {
  *msg = "";
  switch (op) {
    case 7: { uint64_t v = (uint64_t)lhs * rhs; *res = (uint32_t)v; return v >> 32 ? 2 : 0; }  // Mul
    case 1: if (rhs == 0) { *msg = "Division by zero"; return 1; } *res = lhs / rhs; return 0;  // Div
    case 0: { uint64_t v = (uint64_t)lhs + rhs; *res = (uint32_t)v; return v >> 32 ? 2 : 0; }  // Add
    case 9: if (lhs < rhs) { *msg = "Subtraction underflow"; return 1; } *res = lhs - rhs; return 0;  // Sub
    case 11: {  // Pow (u32::pow: square-and-multiply; overflow panics in debug builds, wraps in release)
      uint64_t base = lhs, acc = 1; uint32_t e = rhs; bool ovf = false;
      while (e) {
        if (e & 1) { acc *= base; if (acc >> 32) { ovf = true; acc &= 0xFFFFFFFFull; } }
        e >>= 1;
        if (e) { base *= base; if (base >> 32) { ovf = true; base &= 0xFFFFFFFFull; } }
      }
      *res = (uint32_t)acc; return ovf ? 2 : 0;
    }
    case 12: if (rhs == 0) { *msg = "Integer division by zero"; return 1; } *res = lhs / rhs; return 0;  // IntDiv
    case 13: if (rhs == 0) { *msg = "Modulo by zero"; return 1; } *res = lhs % rhs; return 0;            // Mod
    case 14: *res = rhs < 32 ? lhs << rhs : 0; return rhs < 32 ? 0 : 2;  // ShiftL
    case 15: *res = rhs < 32 ? lhs >> rhs : 0; return rhs < 32 ? 0 : 2;  // ShiftR
    case 5: *res = lhs <= rhs; return 0;
    case 3: *res = lhs >= rhs; return 0;
    case 6: *res = lhs < rhs; return 0;
    case 4: *res = lhs > rhs; return 0;
    case 2: *res = lhs == rhs; return 0;
    case 8: *res = lhs != rhs; return 0;
    case 16: *res = (lhs != 0 || rhs != 0); return 0;
    case 17: *res = (lhs != 0 && rhs != 0); return 0;
    case 18: *res = lhs | rhs; return 0;
    case 19: *res = lhs & rhs; return 0;
    case 10: *res = lhs ^ rhs; return 0;
  }
  *msg = "bad opcode";
  return 1;
}

}  // namespace orc

// =====================================================================================================
// C ABI for ctypes (tests, bench cpu_baseline).  Names are orc_* so they can never be mistaken for c2a_*.
// =====================================================================================================
extern "C" {

struct orc_event { uint32_t kind, a, b, c; };  // same layout as c2a_event

void* orc_new() { return new orc::Compiler(); }
void orc_free(void* h) { delete (orc::Compiler*)h; }
const char* orc_last_error(void* h) { return ((orc::Compiler*)h)->last_error.c_str(); }
const char* orc_gate_name(uint32_t op) { return op < 20 ? orc::kGateNames[op] : nullptr; }

int orc_add_signal(void* h, uint32_t id, const char* name, int has_value, uint32_t value) {
  return ((orc::Compiler*)h)->add_signal(id, name ? name : ("random_" + std::to_string(id)), has_value ? std::optional<uint32_t>(value) : std::nullopt);
}
int orc_add_gate(void* h, uint32_t op, uint32_t l, uint32_t r, uint32_t o) { return ((orc::Compiler*)h)->add_gate(op, l, r, o); }
int orc_add_connection(void* h, uint32_t a, uint32_t b) { return ((orc::Compiler*)h)->add_connection(a, b); }
int orc_emit_events(void* h, const orc_event* ev, uint64_t n, uint64_t* err_event) {
  auto* c = (orc::Compiler*)h;
  for (uint64_t i = 0; i < n; ++i) {
    int st = 0;
    switch (ev[i].kind & 0xFF) {
      case 0: st = c->add_signal(ev[i].a, "random_" + std::to_string(ev[i].a), std::nullopt); break;
      case 1: st = c->add_signal(ev[i].a, "const_signal_" + std::to_string(ev[i].b), ev[i].b); break;
      case 2: st = c->add_gate(ev[i].kind >> 8, ev[i].a, ev[i].b, ev[i].c); break;
      case 3: st = c->add_connection(ev[i].a, ev[i].b); break;
      default: st = orc::E_INVALID;
    }
    if (st) { if (err_event) *err_event = i; return st; }
  }
  return 0;
}
// rename a signal after a bulk replay (bulk events carry no names)
int orc_set_signal_name(void* h, uint32_t id, const char* name) {
  auto* c = (orc::Compiler*)h;
  auto it = c->signals.find(id);
  if (it == c->signals.end()) return orc::E_INVALID;
  it->second.name = name;
  return 0;
}
void orc_add_input(void* h, uint32_t id, const char* name) { ((orc::Compiler*)h)->inputs[id] = name; }
void orc_add_output(void* h, uint32_t id, const char* name) { ((orc::Compiler*)h)->outputs[id] = name; }
void orc_tag_inputs_by_prefix(void* h, const char* p) { ((orc::Compiler*)h)->tag_by_prefix(p, true); }
void orc_tag_outputs_by_prefix(void* h, const char* p) { ((orc::Compiler*)h)->tag_by_prefix(p, false); }
uint64_t orc_num_gates(void* h) { return ((orc::Compiler*)h)->gates.size(); }
uint32_t orc_node_count(void* h) { return ((orc::Compiler*)h)->node_count; }
uint64_t orc_num_signals(void* h) { return ((orc::Compiler*)h)->signals.size(); }
void orc_get_gates(void* h, uint32_t* out) {
  auto* c = (orc::Compiler*)h;
  for (size_t i = 0; i < c->gates.size(); ++i) { out[4 * i] = c->gates[i].op; out[4 * i + 1] = c->gates[i].lh_in; out[4 * i + 2] = c->gates[i].rh_in; out[4 * i + 3] = c->gates[i].out; }
}
uint64_t orc_num_nodes(void* h) { return ((orc::Compiler*)h)->nodes.size(); }
// ids ascending; flags bit0 const bit1 out; sig_off has num_nodes+1 entries
void orc_get_nodes(void* h, uint32_t* ids, uint8_t* flags, uint64_t* sig_off, uint32_t* sig) {
  auto* c = (orc::Compiler*)h;
  std::vector<uint32_t> v;
  for (auto& kv : c->nodes) v.push_back(kv.first);
  std::sort(v.begin(), v.end());
  uint64_t off = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    const orc::Node& n = c->nodes[v[i]];
    ids[i] = v[i];
    flags[i] = (n.is_const ? 1 : 0) | (n.is_out ? 2 : 0);
    sig_off[i] = off;
    if (sig) for (uint32_t s : n.signals) sig[off++] = s; else off += n.signals.size();
  }
  sig_off[v.size()] = off;
}
int orc_signal_node(void* h, uint32_t sid, uint32_t* node) {
  auto* c = (orc::Compiler*)h;
  *node = 0;
  for (auto& kv : c->nodes) if (kv.second.contains_signal(sid)) *node = kv.first;
  return 0;
}
int orc_build_circuit(void* h) { return ((orc::Compiler*)h)->build_circuit(); }
uint32_t orc_circuit_wire_count(void* h) { return ((orc::Compiler*)h)->built.wire_count; }
uint64_t orc_circuit_cycle_at(void* h) { return ((orc::Compiler*)h)->built.cycle_at; }
void orc_circuit_order(void* h, uint32_t* out) { auto& o = ((orc::Compiler*)h)->built.order; for (size_t i = 0; i < o.size(); ++i) out[i] = (uint32_t)o[i]; }
void orc_circuit_gates(void* h, uint32_t* out) {
  auto& g = ((orc::Compiler*)h)->built.new_gates;
  for (size_t i = 0; i < g.size(); ++i) { out[4 * i] = g[i].op; out[4 * i + 1] = g[i].lh_in; out[4 * i + 2] = g[i].rh_in; out[4 * i + 3] = g[i].out; }
}
uint32_t orc_circuit_n_inputs(void* h) { return (uint32_t)((orc::Compiler*)h)->built_input_nodes.size(); }
uint32_t orc_circuit_n_outputs(void* h) { return (uint32_t)((orc::Compiler*)h)->built_output_nodes.size(); }
void orc_circuit_io_nodes(void* h, uint32_t* in, uint32_t* out) {
  auto* c = (orc::Compiler*)h;
  std::copy(c->built_input_nodes.begin(), c->built_input_nodes.end(), in);
  std::copy(c->built_output_nodes.begin(), c->built_output_nodes.end(), out);
}
const char* orc_circuit_info_json(void* h) { return ((orc::Compiler*)h)->info_json.c_str(); }
static std::string g_report;
const char* orc_report_json(void* h, const char* value_type) { g_report = ((orc::Compiler*)h)->report_json(value_type); return g_report.c_str(); }

// generic get_deps form of src/topological_sort.rs:3-6
int orc_topological_sort(uint64_t n, const uint64_t* dep_off, const uint32_t* dep_idx, uint32_t* order_out, uint64_t* err_index) {
  std::vector<size_t> sorted;
  size_t cyc = 0;
  int st = orc::topological_sort(
      n, [&](size_t i) { std::vector<size_t> d; for (uint64_t k = dep_off[i]; k < dep_off[i + 1]; ++k) d.push_back(dep_idx[k]); return d; }, sorted, &cyc);
  if (st == orc::E_CYCLIC && err_index) *err_index = cyc;
  if (st) return st;
  for (size_t i = 0; i < sorted.size(); ++i) order_out[i] = (uint32_t)sorted[i];
  return 0;
}

// src/compiler.rs:385-464 on raw arrays (node ids). wire_of_node[node_bound] gets 0xFFFFFFFF where unset.
int orc_backend_raw(const uint32_t* gates, uint64_t G, uint32_t node_bound, const uint32_t* input_nodes, uint32_t n_in,
                    const uint32_t* output_nodes, uint32_t n_out, uint32_t* order_out, uint32_t* wire_of_node,
                    uint32_t* new_gates, uint32_t* wire_count, uint64_t* err_index) {
  std::vector<orc::Gate> g(G);
  for (uint64_t i = 0; i < G; ++i) g[i] = {gates[4 * i], gates[4 * i + 1], gates[4 * i + 2], gates[4 * i + 3]};
  std::vector<uint32_t> in(input_nodes, input_nodes + n_in), out(output_nodes, output_nodes + n_out);
  orc::BackendResult r;
  int st = orc::backend_core(g, in, out, r);
  if (st == orc::E_CYCLIC && err_index) *err_index = r.cycle_at;
  if (st) return st;
  if (order_out) for (uint64_t i = 0; i < G; ++i) order_out[i] = (uint32_t)r.order[i];
  if (wire_of_node) {
    for (uint32_t i = 0; i < node_bound; ++i) wire_of_node[i] = 0xFFFFFFFFu;
    for (auto& kv : r.node_id_to_wire_id) if (kv.first < node_bound) wire_of_node[kv.first] = kv.second;
  }
  if (new_gates) for (uint64_t i = 0; i < G; ++i) { new_gates[4 * i] = r.new_gates[i].op; new_gates[4 * i + 1] = r.new_gates[i].lh_in; new_gates[4 * i + 2] = r.new_gates[i].rh_in; new_gates[4 * i + 3] = r.new_gates[i].out; }
  if (wire_count) *wire_count = r.wire_count;
  return 0;
}

// Timing entry for the CPU baseline: runs backend_core `reps` times on pre-built gates, returns best seconds.
double orc_backend_time(const uint32_t* gates, uint64_t G, const uint32_t* input_nodes, uint32_t n_in,
                        const uint32_t* output_nodes, uint32_t n_out, int reps, int* status);

int orc_execute_op(uint32_t lhs, uint32_t rhs, uint32_t op, uint32_t* res, const char** msg) { return orc::execute_op(lhs, rhs, op, res, msg); }

// Straight-line u32 evaluator restating tests/integration.rs:90-119 (sim-circuit is un-vendored: it executes
// the gates in the order given, reading input wires that must already hold a value).
// wires[wire_count] in/out; has[wire_count] marks wires with a value. Returns 0, or 1 + gate index of the
// first gate that reads an unset wire / hits a Rust panic (÷0, overflow in debug builds is NOT modelled: wraps).
int64_t orc_simulate(const uint32_t* gates, uint64_t G, uint32_t wire_count, uint32_t* wires, uint8_t* has) {
  for (uint64_t i = 0; i < G; ++i) {
    uint32_t op = gates[4 * i], a = gates[4 * i + 1], b = gates[4 * i + 2], o = gates[4 * i + 3];
    if (a >= wire_count || b >= wire_count || o >= wire_count || !has[a] || !has[b]) return 1 + (int64_t)i;
    uint32_t x = wires[a], y = wires[b], r = 0;
    switch (op) {
      case 0: r = x + y; break;
      case 1: case 12: if (!y) return 1 + (int64_t)i; r = x / y; break;
      case 2: r = x == y; break;
      case 3: r = x >= y; break;
      case 4: r = x > y; break;
      case 5: r = x <= y; break;
      case 6: r = x < y; break;
      case 7: r = x * y; break;
      case 8: r = x != y; break;
      case 9: r = x - y; break;
      case 10: r = x ^ y; break;
      case 11: { uint32_t acc = 1, base = x, e = y; while (e) { if (e & 1) acc *= base; e >>= 1; base *= base; } r = acc; break; }
      case 13: if (!y) return 1 + (int64_t)i; r = x % y; break;
      case 14: r = y < 32 ? x << y : 0; break;
      case 15: r = y < 32 ? x >> y : 0; break;
      case 16: r = (x != 0 || y != 0); break;
      case 17: r = (x != 0 && y != 0); break;
      case 18: r = x | y; break;
      case 19: r = x & y; break;
      default: return 1 + (int64_t)i;
    }
    wires[o] = r;
    has[o] = 1;
  }
  return 0;
}

}  // extern "C"

#include <chrono>
extern "C" double orc_backend_time(const uint32_t* gates, uint64_t G, const uint32_t* input_nodes, uint32_t n_in,
                                   const uint32_t* output_nodes, uint32_t n_out, int reps, int* status) {
  std::vector<orc::Gate> g(G);
  for (uint64_t i = 0; i < G; ++i) g[i] = {gates[4 * i], gates[4 * i + 1], gates[4 * i + 2], gates[4 * i + 3]};
  std::vector<uint32_t> in(input_nodes, input_nodes + n_in), out(output_nodes, output_nodes + n_out);
  double best = 1e300;
  int st = 0;
  for (int r = 0; r < reps; ++r) {
    orc::BackendResult res;
    auto t0 = std::chrono::steady_clock::now();
    st = orc::backend_core(g, in, out, res);
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  if (status) *status = st;
  return best;
}
