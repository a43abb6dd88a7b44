"""ctypes wrapper of oracle/libc2a_oracle.so (TEST INFRASTRUCTURE: the CPU restatement of the reference).
Mirrors the product's Python `Compiler` surface so parity tests can drive both with the same code."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "libc2a_oracle.so")


def build():
    src = os.path.join(ODIR, "c2a_oracle.cpp")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ODIR], stdout=subprocess.DEVNULL)


build()
lib = C.CDLL(SO)
vp, cp, u32, u64, i32 = C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int
u32p, u64p = C.POINTER(u32), C.POINTER(u64)
_S = {
    "orc_new": (vp, []), "orc_free": (None, [vp]), "orc_last_error": (cp, [vp]), "orc_gate_name": (cp, [u32]),
    "orc_add_signal": (i32, [vp, u32, cp, i32, u32]), "orc_add_gate": (i32, [vp, u32, u32, u32, u32]),
    "orc_add_connection": (i32, [vp, u32, u32]), "orc_emit_events": (i32, [vp, vp, u64, u64p]),
    "orc_set_signal_name": (i32, [vp, u32, cp]), "orc_add_input": (None, [vp, u32, cp]), "orc_add_output": (None, [vp, u32, cp]),
    "orc_tag_inputs_by_prefix": (None, [vp, cp]), "orc_tag_outputs_by_prefix": (None, [vp, cp]),
    "orc_num_gates": (u64, [vp]), "orc_node_count": (u32, [vp]), "orc_num_signals": (u64, [vp]), "orc_get_gates": (None, [vp, vp]),
    "orc_num_nodes": (u64, [vp]), "orc_get_nodes": (None, [vp, vp, vp, vp, vp]), "orc_signal_node": (i32, [vp, u32, u32p]),
    "orc_build_circuit": (i32, [vp]), "orc_circuit_wire_count": (u32, [vp]), "orc_circuit_cycle_at": (u64, [vp]),
    "orc_circuit_order": (None, [vp, vp]), "orc_circuit_gates": (None, [vp, vp]), "orc_circuit_n_inputs": (u32, [vp]),
    "orc_circuit_n_outputs": (u32, [vp]), "orc_circuit_io_nodes": (None, [vp, vp, vp]), "orc_circuit_info_json": (cp, [vp]),
    "orc_report_json": (cp, [vp, cp]),
    "orc_topological_sort": (i32, [u64, vp, vp, vp, u64p]),
    "orc_backend_raw": (i32, [vp, u64, u32, vp, u32, vp, u32, vp, vp, vp, u32p, u64p]),
    "orc_backend_time": (C.c_double, [vp, u64, vp, u32, vp, u32, i32, C.POINTER(i32)]),
    "orc_execute_op": (i32, [u32, u32, u32, u32p, C.POINTER(cp)]),
    "orc_simulate": (C.c_int64, [vp, u64, u32, vp, vp]),
}
for _n, (_r, _a) in _S.items():
    _f = getattr(lib, _n)
    _f.restype, _f.argtypes = _r, _a


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


class OracleError(Exception):
    def __init__(self, status, message=""):
        self.status = status
        self.message = message
        super().__init__(f"status {status}: {message}")


class OracleCompiler:
    def __init__(self):
        self._c = lib.orc_new()

    def __del__(self):
        if lib is not None and self._c:
            lib.orc_free(self._c)
            self._c = None

    def _chk(self, st):
        if st:
            raise OracleError(st, lib.orc_last_error(self._c).decode())

    def add_signal(self, id, name, value=None):
        self._chk(lib.orc_add_signal(self._c, id, None if name is None else name.encode(), value is not None, value or 0))

    def add_gate(self, op, l, r, o):
        self._chk(lib.orc_add_gate(self._c, int(op), l, r, o))

    def add_connection(self, a, b):
        self._chk(lib.orc_add_connection(self._c, a, b))

    def emit_events(self, ev):
        ev = np.ascontiguousarray(ev, dtype=np.uint32)
        bad = u64(0)
        st = lib.orc_emit_events(self._c, _p(ev), ev.shape[0], C.byref(bad))
        if st:
            raise OracleError(st, f"event {bad.value}")

    def set_signal_name(self, id, name):
        lib.orc_set_signal_name(self._c, id, name.encode())

    def add_inputs(self, d):
        for k, v in d.items():
            lib.orc_add_input(self._c, k, v.encode())

    def add_outputs(self, d):
        for k, v in d.items():
            lib.orc_add_output(self._c, k, v.encode())

    def tag_inputs_by_prefix(self, p):
        lib.orc_tag_inputs_by_prefix(self._c, p.encode())

    def tag_outputs_by_prefix(self, p):
        lib.orc_tag_outputs_by_prefix(self._c, p.encode())

    @property
    def node_count(self):
        return lib.orc_node_count(self._c)

    @property
    def num_signals(self):
        return lib.orc_num_signals(self._c)

    def signal_node(self, sid):
        n = u32(0)
        lib.orc_signal_node(self._c, sid, C.byref(n))
        return n.value

    def gate_array(self):
        G = lib.orc_num_gates(self._c)
        g = np.empty((G, 4), dtype=np.uint32)
        lib.orc_get_gates(self._c, _p(g))
        return g

    def nodes(self):
        n = lib.orc_num_nodes(self._c)
        ids = np.empty(n, dtype=np.uint32)
        flags = np.empty(n, dtype=np.uint8)
        off = np.empty(n + 1, dtype=np.uint64)
        lib.orc_get_nodes(self._c, _p(ids), _p(flags), _p(off), None)
        sig = np.empty(max(int(off[n]), 1), dtype=np.uint32)
        lib.orc_get_nodes(self._c, _p(ids), _p(flags), _p(off), _p(sig))
        return {int(ids[i]): {"is_const": bool(flags[i] & 1), "is_out": bool(flags[i] & 2),
                              "signals": [int(s) for s in sig[int(off[i]):int(off[i + 1])]]} for i in range(n)}

    def build_circuit(self):
        st = lib.orc_build_circuit(self._c)
        if st:
            raise OracleError(st, lib.orc_last_error(self._c).decode())
        G = lib.orc_num_gates(self._c)
        order = np.empty(G, dtype=np.uint32)
        ga = np.empty((G, 4), dtype=np.uint32)
        lib.orc_circuit_order(self._c, _p(order))
        lib.orc_circuit_gates(self._c, _p(ga))
        return {"wire_count": lib.orc_circuit_wire_count(self._c), "order": order, "gates": ga,
                "info": json.loads(lib.orc_circuit_info_json(self._c).decode())}

    def report(self, value_type="sint"):
        return json.loads(lib.orc_report_json(self._c, value_type.encode()).decode())


def topological_sort(dep_lists):
    n = len(dep_lists)
    off = np.zeros(n + 1, dtype=np.uint64)
    idx = []
    for i, d in enumerate(dep_lists):
        idx.extend(d)
        off[i + 1] = len(idx)
    idx = np.asarray(idx, dtype=np.uint32)
    order = np.empty(n, dtype=np.uint32)
    err = u64(0)
    st = lib.orc_topological_sort(n, _p(off), _p(idx), _p(order), C.byref(err))
    if st:
        raise OracleError(st, f"detected at i={err.value}")
    return order


def backend_raw(gates, node_bound, input_nodes, output_nodes):
    """-> (status, err_index, order, wire_of_node, new_gates, wire_count)"""
    g = np.ascontiguousarray(gates, dtype=np.uint32).reshape(-1, 4)
    G = g.shape[0]
    inn = np.ascontiguousarray(input_nodes, dtype=np.uint32)
    outn = np.ascontiguousarray(output_nodes, dtype=np.uint32)
    order = np.empty(G, dtype=np.uint32)
    wire = np.empty(node_bound, dtype=np.uint32)
    ng = np.empty((G, 4), dtype=np.uint32)
    wc = u32(0)
    err = u64(0)
    st = lib.orc_backend_raw(_p(g), G, node_bound, _p(inn), inn.shape[0], _p(outn), outn.shape[0], _p(order), _p(wire), _p(ng), C.byref(wc), C.byref(err))
    return st, err.value, order, wire, ng, wc.value


def backend_time(gates, input_nodes, output_nodes, reps=3):
    g = np.ascontiguousarray(gates, dtype=np.uint32).reshape(-1, 4)
    inn = np.ascontiguousarray(input_nodes, dtype=np.uint32)
    outn = np.ascontiguousarray(output_nodes, dtype=np.uint32)
    st = i32(0)
    t = lib.orc_backend_time(_p(g), g.shape[0], _p(inn), inn.shape[0], _p(outn), outn.shape[0], reps, C.byref(st))
    return t, st.value


def execute_op(lhs, rhs, op):
    res = u32(0)
    msg = cp()
    st = lib.orc_execute_op(lhs, rhs, int(op), C.byref(res), C.byref(msg))
    return st, res.value, (msg.value or b"").decode()


def simulate(gates, wire_count, values: dict):
    """values: wire -> u32.  Returns the dict of all wires with a value after running the gates in order."""
    g = np.ascontiguousarray(gates, dtype=np.uint32).reshape(-1, 4)
    wires = np.zeros(wire_count, dtype=np.uint32)
    has = np.zeros(wire_count, dtype=np.uint8)
    for k, v in values.items():
        wires[k] = v
        has[k] = 1
    st = lib.orc_simulate(_p(g), g.shape[0], wire_count, _p(wires), _p(has))
    if st:
        raise OracleError(int(st), f"simulation failed at gate {st - 1}")
    return {i: int(wires[i]) for i in range(wire_count) if has[i]}
