"""ctypes binding of include/c2a.h.  The library is REQUIRED: nothing here falls back to Python or to the
oracle — if libc2a.so is missing the import raises, and if no CUDA device is usable the back-end calls raise."""
import ctypes as C
import enum
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libc2a.so")

u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
vp, cp = C.c_void_p, C.c_char_p
u32, u64, i32, i64 = C.c_uint32, C.c_uint64, C.c_int, C.c_int64


class Status(enum.IntEnum):
    OK = 0
    CYCLIC_DEPENDENCY = 1
    INCONSISTENCY = 2
    SIGNAL_ALREADY_DECLARED = 3
    CANNOT_MERGE_OUTPUT_NODES = 4
    CANNOT_MERGE_CONSTANT_NODES = 5
    REFERENCE_PANIC = 6
    INVALID_ARGUMENT = 7
    EVALUATION = 8
    CUDA = -1
    NO_MEMORY = -2


class C2AError(RuntimeError):
    """Runtime failure of the native layer (CUDA, memory, bad arguments)."""

    def __init__(self, status, message=""):
        self.status = Status(status) if status in Status._value2member_map_ else status
        super().__init__(f"{self.status!r}: {message}" if message else repr(self.status))


class CircuitError(Exception):
    """Mirror of the reference's CircuitError (src/compiler.rs:550-576); str() is its thiserror Display text."""

    def __init__(self, status, message=""):
        self.status = Status(status)
        self.message = message
        text = {
            Status.CYCLIC_DEPENDENCY: f"Cyclic dependency: {message}",
            Status.INCONSISTENCY: f"Inconsistency: {message}",
            Status.SIGNAL_ALREADY_DECLARED: "Signal already declared",
            Status.CANNOT_MERGE_OUTPUT_NODES: "Cannot merge output nodes",
            Status.CANNOT_MERGE_CONSTANT_NODES: "Cannot merge constant nodes",
            Status.REFERENCE_PANIC: f"reference panics: {message}",
        }.get(self.status, f"{self.status!r}: {message}")
        super().__init__(text)


class EmitInfo(C.Structure):  # c2a_emit_info
    _fields_ = [("n_events", u64), ("n_signals", u64), ("n_gates", u64), ("n_connections", u64), ("n_effective", u64),
                ("node_count", u32), ("signal_bound", u32), ("path", u32), ("rounds", u32), ("decline_flags", u32), ("reserved", u32)]


class PackedEvents(C.Structure):  # c2a_packed_events
    _fields_ = [("kinds", vp), ("words", vp), ("n_events", u64), ("n_words", u64), ("flags", u32), ("reserved", u32)]


class Replay(C.Structure):  # c2a_replay
    _fields_ = [("k_dst", u64), ("k_src", u64), ("k_len", u64), ("w_dst", u64), ("w_src", u64), ("w_len", u64), ("delta", u32), ("gen", u32)]


class CompressedEvents(C.Structure):  # c2a_compressed_events
    _fields_ = [("kinds", vp), ("words", vp), ("n_events", u64), ("n_words", u64), ("replays", vp), ("n_replays", u64), ("max_gen", u32), ("flags", u32)]


class CompileIO(C.Structure):  # c2a_compile_io
    _fields_ = [("input_signals", vp), ("output_signals", vp), ("n_in", u32), ("n_out", u32), ("order_out", vp), ("wire_of_node", vp), ("new_gates", vp),
                ("gates_cap", u64), ("wire_cap", u32), ("reserved", u32)]


load_error = None
try:
    lib = C.CDLL(LIB_PATH)
except OSError as e:  # fail loudly: there is no fallback implementation
    raise ImportError(
        f"libc2a.so not found at {LIB_PATH} ({e}). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        f"or `make -C circom-2-arithc_b200/csrc`.") from e

_SIGS = {
    "c2a_abi_version": (i32, []),
    "c2a_gate_type_name": (cp, [u32]),
    "c2a_gate_type_from_name": (i32, [cp]),
    "c2a_status_string": (cp, [i32]),
    "c2a_create": (i32, [i32, C.POINTER(vp)]),
    "c2a_destroy": (None, [vp]),
    "c2a_last_error": (cp, [vp]),
    "c2a_device_count": (i32, []),
    "c2a_kernel_launches": (u64, [vp]),
    "c2a_last_kernel_ms": (C.c_double, [vp, cp]),
    "c2a_last_phases": (cp, [vp]),
    "c2a_set_timing": (None, [vp, i32]),
    "c2a_set_timing_only": (None, [vp, cp]),
    "c2a_stream": (vp, [vp]),
    "c2a_topo_sort": (i32, [vp, vp, u64, u32, vp, u64p]),
    "c2a_topo_sort_deps": (i32, [vp, u64, vp, vp, vp, u64p]),
    "c2a_build_circuit": (i32, [vp, vp, u64, u32, vp, u32, vp, u32, vp, vp, vp, u32p, u64p]),
    "c2a_build_circuit_device": (i32, [vp, vp, u64, u32, vp, u32, vp, u32, vp, vp, vp, u32p, u64p]),
    "c2a_rebase_wires_device": (i32, [vp, vp, vp, u64, u32, u32, u32, u32, u32, u32]),
    "c2a_rebase_wire_map_device": (i32, [vp, vp, u64, u32, u32, u32, u32, u32]),
    "c2a_rebase_wires_gathered_device": (i32, [vp, vp, vp, u64, vp, u32, u32]),
    "c2a_plan_shards_device": (i32, [vp, vp, u64, u32, vp, u32, vp, u32, u32, u64p, u32p]),
    "c2a_topo_levels": (i32, [vp, vp, u64, u32, vp, vp, u32, u32p, u64p]),
    "c2a_topo_levels_device": (i32, [vp, vp, u64, u32, vp, vp, u32, u32p, u64p]),
    "c2a_sweep_masks": (i32, [vp, vp, u64, u32, vp, vp, u32, vp, u32, vp, vp, vp, u64p]),
    "c2a_evaluate": (i32, [vp, vp, u64, u32, vp, vp, u64p]),
    "c2a_emit_events_device": (i32, [vp, vp, u64, vp, u64p]),
    "c2a_emit_events_resident": (i32, [vp, vp, u64, vp, u64p]),
    "c2a_emitted_gather_device": (i32, [vp, vp, vp, vp, u32, u32]),
    "c2a_emitted_signal_wires": (i32, [vp, vp, u64, vp]),
    "c2a_emitted_signal_nodes": (i32, [vp, vp, u64, vp]),
    "c2a_emitted_signal_wires_device": (i32, [vp, vp, u64, vp]),
    "c2a_rebase_wire_ids_gathered_device": (i32, [vp, vp, u64, vp, u32, u32]),
    "c2a_pack_events": (u64, [vp, u64, vp, vp, u32p]),
    "c2a_pack_events_ex": (u64, [vp, u64, u32, vp, vp, u32p]),
    "c2a_unpack_events": (i32, [vp, vp]),
    "c2a_emit_packed_device": (i32, [vp, vp, vp, u64p]),
    "c2a_emit_packed_resident": (i32, [vp, vp, vp, u64p]),
    "c2a_compile_packed": (i32, [vp, vp, vp, vp, u32p, u64p, u64p]),
    "c2a_compile_packed_resident": (i32, [vp, vp, vp, vp, u32p, u64p, u64p]),
    "c2a_set_fused_limits": (None, [u64, u32]),
    "c2a_program_packed": (i32, [vp, vp]),
    "c2a_program_compressed": (i32, [vp, vp]),
    "c2a_emit_compressed_device": (i32, [vp, vp, vp, u64p]),
    "c2a_emitted_fetch": (i32, [vp, vp, vp]),
    "c2a_emitted_build_circuit_device": (i32, [vp, vp, u32, vp, u32, vp, vp, vp, u32p, u64p]),
    "c2a_emitted_build_circuit": (i32, [vp, vp, u32, vp, u32, vp, vp, vp, u32p, u64p]),
    "c2a_emitted_build_range_device": (i32, [vp, u64, u64, vp, u32, vp, u32, vp, vp, vp, u32p, u64p]),
    "c2a_emitted_gates_device": (vp, [vp]),
    "c2a_compiler_new": (vp, []),
    "c2a_compiler_free": (None, [vp]),
    "c2a_compiler_last_error": (cp, [vp]),
    "c2a_add_signal": (i32, [vp, u32, cp, i32, u32]),
    "c2a_add_gate": (i32, [vp, u32, u32, u32, u32]),
    "c2a_add_connection": (i32, [vp, u32, u32]),
    "c2a_emit_events": (i32, [vp, vp, u64, u64p]),
    "c2a_add_input": (i32, [vp, u32, cp]),
    "c2a_add_output": (i32, [vp, u32, cp]),
    "c2a_tag_inputs_by_prefix": (i32, [vp, cp]),
    "c2a_tag_outputs_by_prefix": (i32, [vp, cp]),
    "c2a_num_gates": (u64, [vp]),
    "c2a_node_count": (u32, [vp]),
    "c2a_num_signals": (u64, [vp]),
    "c2a_get_gates": (i32, [vp, vp]),
    "c2a_signal_node": (i32, [vp, u32, u32p]),
    "c2a_signal_nodes": (i32, [vp, vp, u64, vp]),
    "c2a_signal_name": (i64, [vp, u32, C.c_char_p, u64]),
    "c2a_set_signal_name": (i32, [vp, u32, cp]),
    "c2a_get_signals_by_prefix": (u64, [vp, cp, vp, u64]),
    "c2a_num_nodes": (u64, [vp]),
    "c2a_get_nodes": (i32, [vp, vp, vp, vp, vp]),
    "c2a_signal_value": (i32, [vp, u32, C.POINTER(i32), u32p]),
    "c2a_program_new": (vp, []),
    "c2a_program_free": (None, [vp]),
    "c2a_program_compile_file": (i32, [vp, cp, vp]),
    "c2a_program_compile_source": (i32, [vp, cp, cp, vp]),
    "c2a_program_error": (cp, [vp]),
    "c2a_program_num_events": (u64, [vp]),
    "c2a_program_events": (vp, [vp]),
    "c2a_program_num_signals": (u64, [vp]),
    "c2a_program_num_constants": (u64, [vp]),
    "c2a_program_constant_signals": (vp, [vp]),
    "c2a_program_constant_values": (vp, [vp]),
    "c2a_program_signal_name": (cp, [vp, u32]),
    "c2a_program_signal_names": (u64, [vp, vp, u64, vp, u64]),
    "c2a_program_num_inputs": (u32, [vp]),
    "c2a_program_num_outputs": (u32, [vp]),
    "c2a_program_inputs": (vp, [vp]),
    "c2a_program_outputs": (vp, [vp]),
    "c2a_compiler_build_circuit": (i32, [vp, vp]),
    "c2a_circuit_wire_count": (u64, [vp]),
    "c2a_circuit_order": (vp, [vp]),
    "c2a_circuit_gates": (vp, [vp]),
    "c2a_circuit_info_json": (cp, [vp]),
    "c2a_circuit_report_json": (cp, [vp, cp]),
    "c2a_bristol_gate_lines": (u64, [vp, u64, vp, u64]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not export what include/c2a.h declares
    _f.restype = _res
    _f.argtypes = _args

EXPORTED = sorted(_SIGS)


def have_device() -> bool:
    return lib.c2a_device_count() > 0
