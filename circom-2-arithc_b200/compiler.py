"""Host-side mirror of the reference's circuit IR interface for the flattening hot path.

Same names, argument meaning and error behaviour as the reference so that tests read like its own:
  AGateType                      src/a_gate_type.rs:7-28
  Compiler.add_signal            src/compiler.rs:139-161
  Compiler.add_gate              src/compiler.rs:174-209
  Compiler.add_connection        src/compiler.rs:213-278
  Compiler.add_inputs/outputs    src/compiler.rs:131-137
  Compiler.get_signals           src/compiler.rs:163-171
  Compiler.build_circuit         src/compiler.rs:321-494   -> BristolCircuit (bristol-circuit crate shape)
  topological_sort               src/topological_sort.rs:3-6
Everything is executed by libc2a.so (C ABI, include/c2a.h); the sort / wire numbering / gather run on the GPU.
"""
import ctypes as C
import enum
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from ._lib import lib, C2AError, CircuitError, Status, EmitInfo, PackedEvents, CompileIO

NONE = 0xFFFFFFFF
EVENT_DTYPE = np.dtype([("kind", "<u4"), ("a", "<u4"), ("b", "<u4"), ("c", "<u4")])
GATE_DTYPE = np.dtype([("op", "<u4"), ("lh", "<u4"), ("rh", "<u4"), ("out", "<u4")])
from .gate_types import AGateType, EV_SIGNAL, EV_SIGNAL_CONST, EV_GATE, EV_CONNECT  # noqa: E402,F401


_CIRCUIT_STATUSES = {Status.CYCLIC_DEPENDENCY, Status.INCONSISTENCY, Status.SIGNAL_ALREADY_DECLARED,
                     Status.CANNOT_MERGE_OUTPUT_NODES, Status.CANNOT_MERGE_CONSTANT_NODES, Status.REFERENCE_PANIC}


def _raise(status: int, message: str):
    if status == 0:
        return
    try:
        st = Status(status)
    except ValueError:
        raise C2AError(status, message)
    if st in _CIRCUIT_STATUSES:
        raise CircuitError(st, message)
    raise C2AError(st, message)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


PACKED_DENSE_IDS, PACKED_IMPLICIT_OPERANDS = 1, 2


def pack_events(events: np.ndarray, implicit: bool = False):
    """c2a_pack_events[_ex]: AoS events -> (kinds u8[n], words u32[n_words], flags).  implicit=True also drops the operand the
    walker derives from the signal it declared last (C2A_PACKED_IMPLICIT_OPERANDS: 4 B/event instead of 6 on a walker stream)."""
    ev = np.ascontiguousarray(events)
    n = ev.shape[0]
    flags = C.c_uint32(0)
    allow = PACKED_DENSE_IDS | (PACKED_IMPLICIT_OPERANDS if implicit else 0)
    nw = int(lib.c2a_pack_events_ex(_ptr(ev), n, allow, None, None, C.byref(flags)))
    kinds = np.empty(n, dtype=np.uint8)
    words = np.empty(nw, dtype=np.uint32)
    lib.c2a_pack_events_ex(_ptr(ev), n, allow, _ptr(kinds), _ptr(words), C.byref(flags))
    return kinds, words, int(flags.value)


def unpack_events(kinds: np.ndarray, words: np.ndarray, flags: int) -> np.ndarray:
    """c2a_unpack_events -> (n, 4) u32 AoS events (constant values read back as 0)."""
    kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
    words = np.ascontiguousarray(words, dtype=np.uint32)
    pk = PackedEvents(_ptr(kinds), _ptr(words), kinds.shape[0], words.shape[0], flags, 0)
    out = np.empty((kinds.shape[0], 4), dtype=np.uint32)
    _raise(lib.c2a_unpack_events(C.byref(pk), _ptr(out)), "n_words does not match the kinds")
    return out


class DeviceContext:
    """Owns a c2a_handle (one CUDA stream + scratch slab on one GPU).  Raises when no CUDA device is usable:
    the back end has no CPU implementation."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        st = lib.c2a_create(int(device), C.byref(h))
        if st != 0:
            raise C2AError(st, f"c2a_create(device={device}) failed: no usable CUDA device; "
                               f"the topological sort / build_circuit back end has no CPU fallback")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib.c2a_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def last_error(self) -> str:
        return lib.c2a_last_error(self._h).decode()

    def kernel_launches(self) -> int:
        return int(lib.c2a_kernel_launches(self._h))

    def phases(self) -> Dict[str, float]:
        s = lib.c2a_last_phases(self._h).decode()
        return {k: float(v) for k, v in (kv.split("=") for kv in s.split(",") if kv)}

    # ---- raw back-end calls on numpy arrays (host buffers; H2D/D2H inside the call) ----
    def topo_sort(self, gates: np.ndarray, node_bound: int) -> np.ndarray:
        gates = _as_gates(gates)
        G = gates.shape[0]
        order = np.empty(G, dtype=np.uint32)
        err = C.c_uint64(0)
        st = lib.c2a_topo_sort(self._h, _ptr(gates), G, node_bound, _ptr(order), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        _raise(st, self.last_error())
        return order

    def topo_sort_deps(self, dep_off: np.ndarray, dep_idx: np.ndarray) -> np.ndarray:
        dep_off = np.ascontiguousarray(dep_off, dtype=np.uint64)
        dep_idx = np.ascontiguousarray(dep_idx, dtype=np.uint32)
        n = dep_off.shape[0] - 1
        order = np.empty(n, dtype=np.uint32)
        err = C.c_uint64(0)
        st = lib.c2a_topo_sort_deps(self._h, n, _ptr(dep_off), _ptr(dep_idx), _ptr(order), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        _raise(st, self.last_error())
        return order

    def build_circuit(self, gates: np.ndarray, node_bound: int, input_nodes, output_nodes,
                      want_order=True, want_wires=True, want_gates=True):
        """-> (order[G], wire_of_node[node_bound], new_gates[G,4], wire_count)"""
        gates = _as_gates(gates)
        G = gates.shape[0]
        inn = np.ascontiguousarray(input_nodes, dtype=np.uint32)
        outn = np.ascontiguousarray(output_nodes, dtype=np.uint32)
        order = np.empty(G, dtype=np.uint32) if want_order else None
        wire = np.empty(node_bound, dtype=np.uint32) if want_wires else None
        ng = np.empty((G, 4), dtype=np.uint32) if want_gates else None
        wc = C.c_uint32(0)
        err = C.c_uint64(0)
        st = lib.c2a_build_circuit(self._h, _ptr(gates), G, node_bound, _ptr(inn), inn.shape[0], _ptr(outn), outn.shape[0],
                                   _ptr(order), _ptr(wire), _ptr(ng), C.byref(wc), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        _raise(st, self.last_error())
        return order, wire, ng, wc.value

    def topo_levels(self, gates: np.ndarray, node_bound: int, level_cap: Optional[int] = None):
        """Kahn levels -> (level_order[G], level_off[n_levels+1])."""
        gates = _as_gates(gates)
        G = gates.shape[0]
        cap = int(level_cap if level_cap is not None else G + 1)
        lo = np.empty(G, dtype=np.uint32)
        off = np.empty(cap + 1, dtype=np.uint32)
        nl = C.c_uint32(0)
        err = C.c_uint64(0)
        st = lib.c2a_topo_levels(self._h, _ptr(gates), G, node_bound, _ptr(lo), _ptr(off), cap, C.byref(nl), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        _raise(st, self.last_error())
        return lo, off[: nl.value + 1].copy()

    def sweep_masks(self, gates: np.ndarray, node_bound: int, const_nodes, const_values, output_nodes):
        gates = _as_gates(gates)
        G = gates.shape[0]
        cn = np.ascontiguousarray(const_nodes, dtype=np.uint32)
        cv = np.ascontiguousarray(const_values, dtype=np.uint32)
        on = np.ascontiguousarray(output_nodes, dtype=np.uint32)
        cm = np.empty(G, dtype=np.uint8)
        cval = np.empty(G, dtype=np.uint32)
        dm = np.empty(G, dtype=np.uint8)
        err = C.c_uint64(0)
        st = lib.c2a_sweep_masks(self._h, _ptr(gates), G, node_bound, _ptr(cn), _ptr(cv), cn.shape[0], _ptr(on), on.shape[0],
                                 _ptr(cm), _ptr(cval), _ptr(dm), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        _raise(st, self.last_error())
        return cm, cval, dm


    def evaluate(self, gates: np.ndarray, wire_count: int, values: Dict[int, int]) -> Dict[int, int]:
        """c2a_evaluate: run a built circuit (wire-id gates in executable order) on u32 values, level-parallel on the GPU.
        values: wire -> u32 for inputs and constants.  Returns wire -> value for every wire that holds one."""
        g = _as_gates(gates)
        vals = np.zeros(max(wire_count, 1), dtype=np.uint32)
        has = np.zeros(max(wire_count, 1), dtype=np.uint8)
        for k, v in values.items():
            vals[k] = v & 0xFFFFFFFF
            has[k] = 1
        err = C.c_uint64(0)
        st = lib.c2a_evaluate(self._h, _ptr(g), g.shape[0], wire_count, _ptr(vals), _ptr(has), C.byref(err))
        if st == Status.EVALUATION:
            e = C2AError(st, self.last_error())
            e.err_index = err.value
            raise e
        _raise(st, self.last_error())
        return {int(i): int(vals[i]) for i in np.nonzero(has[:wire_count])[0]}

    # ---- device emitter: the event stream is replayed on the GPU and the result stays resident ----
    def emit_events(self, events: np.ndarray) -> dict:
        """c2a_emit_events_device: add_signal / add_gate / add_connection (src/compiler.rs:139-278) for a whole event
        stream on the GPU.  Returns the c2a_emit_info fields; raises CircuitError exactly where the reference errors."""
        ev = np.ascontiguousarray(events)
        assert ev.dtype == EVENT_DTYPE or (ev.dtype == np.uint32 and ev.ndim == 2 and ev.shape[1] == 4)
        info = EmitInfo()
        bad = C.c_uint64(0)
        st = lib.c2a_emit_events_device(self._h, _ptr(ev), ev.shape[0], C.byref(info), C.byref(bad))
        if st != 0:
            e = None
            try:
                _raise(st, f"event {bad.value}: {self.last_error()}")
            except (CircuitError, C2AError) as ex:
                ex.err_event = bad.value
                e = ex
            raise e
        self._emit_info = {k: int(getattr(info, k)) for k, _ in EmitInfo._fields_ if k != "reserved"}
        return dict(self._emit_info)

    def emit_compressed(self, cx) -> dict:
        """c2a_emit_compressed_device: a walker recording whose replayed instances are expanded on the GPU (cx: CompressedEvents
        from c2a_program_compressed; the program object must stay alive during the call)."""
        info = EmitInfo()
        bad = C.c_uint64(0)
        st = lib.c2a_emit_compressed_device(self._h, C.byref(cx), C.byref(info), C.byref(bad))
        if st != 0:
            e = None
            try:
                _raise(st, f"event {bad.value}: {self.last_error()}")
            except (CircuitError, C2AError) as ex:
                ex.err_event = bad.value
                e = ex
            raise e
        self._emit_info = {k: int(getattr(info, k)) for k, _ in EmitInfo._fields_ if k != "reserved"}
        return dict(self._emit_info)

    def emit_packed(self, kinds: np.ndarray, words: np.ndarray, flags: int) -> dict:
        """c2a_emit_packed_device: the same replay from a packed stream (pack_events() / Program.packed())."""
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        pk = PackedEvents(_ptr(kinds), _ptr(words), kinds.shape[0], words.shape[0], flags, 0)
        info = EmitInfo()
        bad = C.c_uint64(0)
        st = lib.c2a_emit_packed_device(self._h, C.byref(pk), C.byref(info), C.byref(bad))
        if st != 0:
            e = None
            try:
                _raise(st, f"event {bad.value}: {self.last_error()}")
            except (CircuitError, C2AError) as ex:
                ex.err_event = bad.value
                e = ex
            raise e
        self._emit_info = {k: int(getattr(info, k)) for k, _ in EmitInfo._fields_ if k != "reserved"}
        return dict(self._emit_info)

    def compile_packed(self, kinds: np.ndarray, words: np.ndarray, flags: int, input_signals, output_signals,
                       want_order=True, want_wires=True, want_gates=True):
        """c2a_compile_packed: emit + build in one call (one synchronisation; circuits up to ~1 M gates run inside one cooperative
        kernel).  -> (info, order, wire_of_node[node_count+1], new_gates, wire_count); same results and errors as emit_packed()
        followed by emitted_build_circuit()."""
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        ins = np.ascontiguousarray(input_signals, dtype=np.uint32)
        outs = np.ascontiguousarray(output_signals, dtype=np.uint32)
        n = kinds.shape[0]
        n_g = int(np.count_nonzero((kinds & 3) == 2))
        n_c = int(np.count_nonzero((kinds & 3) == 3))
        gcap, wcap = n_g, n - n_g + 1   # node_count + 1 <= signals + connections + 1
        order = np.empty(gcap, dtype=np.uint32) if want_order else None
        wire = np.empty(wcap, dtype=np.uint32) if want_wires else None
        ng = np.empty((gcap, 4), dtype=np.uint32) if want_gates else None
        pk = PackedEvents(_ptr(kinds), _ptr(words), n, words.shape[0], flags, 0)
        io = CompileIO(_ptr(ins), _ptr(outs), ins.shape[0], outs.shape[0], _ptr(order), _ptr(wire), _ptr(ng), gcap, wcap, 0)
        info = EmitInfo()
        wc, bad, err = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        st = lib.c2a_compile_packed(self._h, C.byref(pk), C.byref(io), C.byref(info), C.byref(wc), C.byref(bad), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        if st != 0:
            e = None
            try:
                _raise(st, f"event {bad.value}: {self.last_error()}")
            except (CircuitError, C2AError) as ex:
                ex.err_event = bad.value
                e = ex
            raise e
        self._emit_info = {k: int(getattr(info, k)) for k, _ in EmitInfo._fields_ if k != "reserved"}
        nb = self._emit_info["node_count"] + 1
        return dict(self._emit_info), order, (wire[:nb] if wire is not None else None), ng, wc.value

    def emitted_signal_wires(self, signals) -> np.ndarray:
        """c2a_emitted_signal_wires: wire ids of the given signals after emitted_build_circuit (0xFFFFFFFF = none)."""
        sig = np.ascontiguousarray(signals, dtype=np.uint32)
        out = np.empty(sig.shape[0], dtype=np.uint32)
        _raise(lib.c2a_emitted_signal_wires(self._h, _ptr(sig), sig.shape[0], _ptr(out)), self.last_error())
        return out

    def emitted_signal_nodes(self, signals) -> np.ndarray:
        """c2a_emitted_signal_nodes: node ids of the given signals of the resident emitted circuit (0 = never declared)."""
        sig = np.ascontiguousarray(signals, dtype=np.uint32)
        out = np.empty(sig.shape[0], dtype=np.uint32)
        _raise(lib.c2a_emitted_signal_nodes(self._h, _ptr(sig), sig.shape[0], _ptr(out)), self.last_error())
        return out

    def emitted_fetch(self, want_gates=True, want_nodes=True):
        """-> (gates (G,4) u32 node ids in emission order, node_of_signal[signal_bound])"""
        info = self._emit_info
        g = np.empty((info["n_gates"], 4), dtype=np.uint32) if want_gates else None
        nos = np.empty(info["signal_bound"], dtype=np.uint32) if want_nodes else None
        _raise(lib.c2a_emitted_fetch(self._h, _ptr(g), _ptr(nos)), self.last_error())
        return g, nos

    def emitted_build_circuit(self, input_signals, output_signals, want_order=True, want_wires=True, want_gates=True):
        """build_circuit on the resident emitted circuit -> (order, wire_of_node[node_count+1], new_gates, wire_count)"""
        info = self._emit_info
        G, nb = info["n_gates"], info["node_count"] + 1
        ins = np.ascontiguousarray(input_signals, dtype=np.uint32)
        outs = np.ascontiguousarray(output_signals, dtype=np.uint32)
        order = np.empty(G, dtype=np.uint32) if want_order else None
        wire = np.empty(nb, dtype=np.uint32) if want_wires else None
        ng = np.empty((G, 4), dtype=np.uint32) if want_gates else None
        wc = C.c_uint32(0)
        err = C.c_uint64(0)
        st = lib.c2a_emitted_build_circuit(self._h, _ptr(ins), ins.shape[0], _ptr(outs), outs.shape[0], _ptr(order), _ptr(wire), _ptr(ng),
                                           C.byref(wc), C.byref(err))
        if st == Status.CYCLIC_DEPENDENCY:
            raise CircuitError(st, f"detected at i={err.value}")
        _raise(st, self.last_error())
        return order, wire, ng, wc.value


def _as_gates(g) -> np.ndarray:
    a = np.asarray(g)
    if a.dtype == GATE_DTYPE:
        a = a.view(np.uint32).reshape(-1, 4)
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if a.ndim != 2 or a.shape[1] != 4:
        a = a.reshape(-1, 4)
    return a


_default_ctx: Dict[int, DeviceContext] = {}


def default_context(device: int = 0) -> DeviceContext:
    if device not in _default_ctx:
        _default_ctx[device] = DeviceContext(device)
    return _default_ctx[device]


def topological_sort(length: int, get_deps, device: int = 0) -> List[int]:
    """src/topological_sort.rs:3-6 — `get_deps(i)` returns the dependencies of item i (at most 2, the shape the
    reference passes).  Runs on the GPU; raises CircuitError('Cyclic dependency: detected at i=..')."""
    off = np.zeros(length + 1, dtype=np.uint64)
    idx: List[int] = []
    for i in range(length):
        d = list(get_deps(i))
        idx.extend(d)
        off[i + 1] = len(idx)
    return default_context(device).topo_sort_deps(off, np.asarray(idx, dtype=np.uint32)).tolist()


@dataclass
class Gate:  # bristol_circuit::Gate as filled at src/compiler.rs:456-463
    inputs: List[int]
    outputs: List[int]
    op: str


@dataclass
class ConstantInfo:
    value: str
    wire_index: int


@dataclass
class CircuitInfo:
    input_name_to_wire_index: Dict[str, int] = field(default_factory=dict)
    constants: Dict[str, ConstantInfo] = field(default_factory=dict)
    output_name_to_wire_index: Dict[str, int] = field(default_factory=dict)


@dataclass
class BristolCircuit:  # src/compiler.rs:478-493
    wire_count: int
    info: CircuitInfo
    gate_array: np.ndarray  # (G,4) u32: op, in0 wire, in1 wire, out wire — the device result, sorted order
    order: np.ndarray       # (G,) u32: sorted_gate_ids (src/compiler.rs:408)
    io_widths: Optional[list] = None

    @property
    def gates(self) -> List[Gate]:
        return [Gate([int(g[1]), int(g[2])], [int(g[3])], AGateType(int(g[0])).name) for g in self.gate_array]


class Compiler:
    """The reference's `Compiler` (src/compiler.rs:107-284) over the native union-find emitter."""

    def __init__(self, device: int = 0, context: Optional[DeviceContext] = None):
        self._c = lib.c2a_compiler_new()
        self._device = device
        self._ctx = context
        self.value_type = "sint"

    def __del__(self):
        try:
            if self._c:
                lib.c2a_compiler_free(self._c)
                self._c = None
        except Exception:
            pass

    def _err(self):
        return lib.c2a_compiler_last_error(self._c).decode()

    # ---- emit side ----
    def add_signal(self, id: int, name: Optional[str], value: Optional[int] = None):
        _raise(lib.c2a_add_signal(self._c, id, None if name is None else name.encode(), value is not None, value or 0), self._err())

    def add_gate(self, gate_type, lhs_signal_id: int, rhs_signal_id: int, output_signal_id: int):
        _raise(lib.c2a_add_gate(self._c, int(gate_type), lhs_signal_id, rhs_signal_id, output_signal_id), self._err())

    def add_connection(self, a: int, b: int):
        _raise(lib.c2a_add_connection(self._c, a, b), self._err())

    def emit_events(self, events: np.ndarray):
        ev = np.ascontiguousarray(events)
        assert ev.dtype == EVENT_DTYPE or (ev.dtype == np.uint32 and ev.shape[-1] == 4)
        n = ev.shape[0]
        bad = C.c_uint64(0)
        st = lib.c2a_emit_events(self._c, _ptr(ev), n, C.byref(bad))
        _raise(st, f"event {bad.value}: {self._err()}")

    def add_inputs(self, inputs: Dict[int, str]):
        for k, v in inputs.items():
            lib.c2a_add_input(self._c, k, v.encode())

    def add_outputs(self, outputs: Dict[int, str]):
        for k, v in outputs.items():
            lib.c2a_add_output(self._c, k, v.encode())

    def get_signals(self, filter: str) -> Dict[int, str]:
        n = lib.c2a_get_signals_by_prefix(self._c, filter.encode(), None, 0)
        ids = np.empty(max(int(n), 1), dtype=np.uint32)
        lib.c2a_get_signals_by_prefix(self._c, filter.encode(), _ptr(ids), n)
        return {int(i): self.signal_name(int(i)) for i in ids[:n]}

    def tag_inputs_by_prefix(self, prefix: str):   # src/program.rs:57-60
        lib.c2a_tag_inputs_by_prefix(self._c, prefix.encode())

    def tag_outputs_by_prefix(self, prefix: str):  # src/program.rs:62-66
        lib.c2a_tag_outputs_by_prefix(self._c, prefix.encode())

    def update_type(self, value_type: str):
        self.value_type = value_type

    def set_signal_name(self, id: int, name: str):
        _raise(lib.c2a_set_signal_name(self._c, id, name.encode()), "unknown signal")

    def signal_value(self, id: int) -> Optional[int]:
        has, val = C.c_int(0), C.c_uint32(0)
        if lib.c2a_signal_value(self._c, id, C.byref(has), C.byref(val)) != 0 or not has.value:
            return None
        return int(val.value)

    def generate_circuit_report(self) -> dict:  # src/compiler.rs:287-319
        from .program import generate_circuit_report
        return generate_circuit_report(self)

    def signal_name(self, id: int) -> Optional[str]:
        buf = C.create_string_buffer(512)
        n = lib.c2a_signal_name(self._c, id, buf, 512)
        return None if n < 0 else buf.value.decode()

    # ---- inspection (the reference's fields are private; its unit tests peek at them) ----
    @property
    def node_count(self) -> int:
        return int(lib.c2a_node_count(self._c))

    @property
    def num_signals(self) -> int:
        return int(lib.c2a_num_signals(self._c))

    def signal_node(self, id: int) -> int:
        n = C.c_uint32(0)
        lib.c2a_signal_node(self._c, id, C.byref(n))
        return n.value

    def signal_nodes(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty(ids.shape[0], dtype=np.uint32)
        lib.c2a_signal_nodes(self._c, _ptr(ids), ids.shape[0], _ptr(out))
        return out

    def gate_array(self) -> np.ndarray:
        """(G,4) u32 {op, lh_in, rh_in, out} in NODE ids — `Compiler.gates` of the reference."""
        G = int(lib.c2a_num_gates(self._c))
        g = np.empty((G, 4), dtype=np.uint32)
        lib.c2a_get_gates(self._c, _ptr(g))
        return g

    def nodes(self) -> Dict[int, dict]:
        n = int(lib.c2a_num_nodes(self._c))
        ids = np.empty(n, dtype=np.uint32)
        flags = np.empty(n, dtype=np.uint8)
        off = np.empty(n + 1, dtype=np.uint64)
        lib.c2a_get_nodes(self._c, _ptr(ids), _ptr(flags), _ptr(off), None)
        sig = np.empty(max(int(off[n]), 1), dtype=np.uint32)
        lib.c2a_get_nodes(self._c, _ptr(ids), _ptr(flags), _ptr(off), _ptr(sig))
        return {int(ids[i]): {"is_const": bool(flags[i] & 1), "is_out": bool(flags[i] & 2),
                              "signals": [int(s) for s in sig[int(off[i]):int(off[i + 1])]]} for i in range(n)}

    # ---- back end ----
    def build_circuit(self) -> BristolCircuit:
        ctx = self._ctx or default_context(self._device)
        st = lib.c2a_compiler_build_circuit(self._c, ctx.handle)
        if st != 0:
            msg = self._err()
            if st == Status.CUDA or st == Status.NO_MEMORY:
                raise C2AError(st, msg)
            _raise(st, msg)
        G = int(lib.c2a_num_gates(self._c))
        info = json.loads(lib.c2a_circuit_info_json(self._c).decode())
        ci = CircuitInfo(
            input_name_to_wire_index=dict(info["input_name_to_wire_index"]),
            constants={k: ConstantInfo(v["value"], v["wire_index"]) for k, v in info["constants"].items()},
            output_name_to_wire_index=dict(info["output_name_to_wire_index"]))
        if G:
            ga = np.ctypeslib.as_array(C.cast(lib.c2a_circuit_gates(self._c), C.POINTER(C.c_uint32)), shape=(G, 4)).copy()
            order = np.ctypeslib.as_array(C.cast(lib.c2a_circuit_order(self._c), C.POINTER(C.c_uint32)), shape=(G,)).copy()
        else:
            ga = np.zeros((0, 4), dtype=np.uint32)
            order = np.zeros((0,), dtype=np.uint32)
        return BristolCircuit(wire_count=int(lib.c2a_circuit_wire_count(self._c)), info=ci, gate_array=ga, order=order)
