"""Mirror of the reference's program layer (src/program.rs, src/main.rs, src/cli.rs) on top of the native front end
(csrc/c2a_front.cpp) and the device back end:

    compile(Args | path)            src/program.rs:18-74  -> Compiler (signals named, inputs/outputs tagged, event stream kept)
    Compiler.generate_circuit_report  src/compiler.rs:287-319, 503-531
    write_outputs / main()          src/main.rs:15-50     circuit.txt, circuit_info.json, report.json

Third-party output formats whose source is not in the reference tree (bristol-circuit @ 2a8b001 `write_bristol`) are
restated from their published layout and are PARITY UNPINNED.  `--boolify-width` (boolify @ 6376405,
un-vendored, untested upstream) is accepted and rejected with a clear message.
"""
import argparse
import ctypes as C
import json
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._lib import lib, CircuitError, CompressedEvents, PackedEvents, Status
from .compiler import AGateType, BristolCircuit, CircuitInfo, Compiler, ConstantInfo, DeviceContext, EVENT_DTYPE, default_context


class ProgramError(Exception):
    """src/program.rs:77-117; str() is the thiserror Display text (e.g. 'Runtime error: Index out of bounds')."""

    def __init__(self, status: int, text: str):
        self.status = status
        super().__init__(text)


@dataclass
class Args:  # src/cli.rs:19-53 (same names and defaults)
    input: str = "./input/circuit.circom"
    output: str = "./output/"
    value_type: str = "sint"
    boolify_width: Optional[int] = None


def build_output(output_path: str, filename: str, ext: str) -> str:  # src/cli.rs:72-76
    return os.path.join(output_path, f"{filename}.{ext}")


class DeviceCompiler:
    """compile(..., emitter="device"): the walk only RECORDS the add_signal / add_gate / add_connection calls (no host union-find);
    build_circuit() replays them on the GPU as a packed stream (c2a_emit_packed_device), numbers and gathers the gates
    (c2a_emitted_build_circuit) and looks up the wires of the named signals (c2a_emitted_signal_nodes / _wires) - the same
    BristolCircuit, the same CircuitError conditions as Compiler.build_circuit (src/compiler.rs:321-494)."""

    def __init__(self, prog, device: int, context: Optional[DeviceContext]):
        self._prog, self._device, self._ctx = prog, device, context
        self.value_type = "sint"
        n = int(lib.c2a_program_num_events(prog))
        # The recorded calls and their packed form stay where the walker wrote them (hundreds of MB at 10 M gates): the arrays
        # below are VIEWS of the program's memory, alive as long as this object; `events` hands out a copy on first use.
        view = lambda p, k, ct, dt: np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(k,)) if k else np.zeros(0, dt)
        self._n_events = n
        self._events_copy = None
        nc = int(lib.c2a_program_num_constants(prog))
        self._const_signals = view(lib.c2a_program_constant_signals(prog), nc, C.c_uint32, np.uint32)
        self._const_values = view(lib.c2a_program_constant_values(prog), nc, C.c_uint32, np.uint32)
        ni, no = int(lib.c2a_program_num_inputs(prog)), int(lib.c2a_program_num_outputs(prog))
        self.input_signals = view(lib.c2a_program_inputs(prog), ni, C.c_uint32, np.uint32).copy()
        self.output_signals = view(lib.c2a_program_outputs(prog), no, C.c_uint32, np.uint32).copy()
        self._packed = None  # (kinds, words, flags) views of the fully expanded packed stream, made on first use
        self._view = view

    def compressed(self) -> CompressedEvents:
        """the recording as the walker keeps it: replayed instances are records, not copies (c2a_program_compressed)"""
        cx = CompressedEvents()
        lib.c2a_program_compressed(self._prog, C.byref(cx))
        return cx

    def _packed_views(self):
        if self._packed is None:  # c2a_program_packed carries the replay records out on the host
            pk = PackedEvents()
            lib.c2a_program_packed(self._prog, C.byref(pk))
            self._packed = (self._view(pk.kinds, int(pk.n_events), C.c_uint8, np.uint8), self._view(pk.words, int(pk.n_words), C.c_uint32, np.uint32), int(pk.flags))
        return self._packed

    _kinds = property(lambda self: self._packed_views()[0])
    _words = property(lambda self: self._packed_views()[1])
    _flags = property(lambda self: self._packed_views()[2])

    @property
    def events(self) -> np.ndarray:
        """the recorded add_signal / add_gate / add_connection calls as c2a_event rows (a private copy)"""
        if self._events_copy is None:  # the walker records the packed form; the 16-byte records are written on this request
            n = self._n_events
            self._events_copy = (np.ctypeslib.as_array(C.cast(lib.c2a_program_events(self._prog), C.POINTER(C.c_uint32)), shape=(4 * n,)).reshape(n, 4).copy()
                                 if n else np.zeros((0, 4), np.uint32))
        return self._events_copy

    def __del__(self):
        try:
            if self._prog:
                lib.c2a_program_free(self._prog)
                self._prog = None
        except Exception:
            pass

    def update_type(self, value_type: str):
        self.value_type = value_type

    def signal_name(self, sid: int) -> str:
        p = lib.c2a_program_signal_name(self._prog, int(sid))
        return p.decode() if p is not None else ""

    def signal_names(self, sids) -> list:
        """names of many signals with one call (c2a_program_signal_names)"""
        ids = np.ascontiguousarray(sids, dtype=np.uint32)
        if ids.shape[0] == 0:
            return []
        need = int(lib.c2a_program_signal_names(self._prog, ids.ctypes.data_as(C.c_void_p), ids.shape[0], None, 0))
        buf = C.create_string_buffer(need + 1)
        lib.c2a_program_signal_names(self._prog, ids.ctypes.data_as(C.c_void_p), ids.shape[0], buf, need)
        return buf.raw[:need - 1].decode().split("\n")

    def build_circuit(self) -> BristolCircuit:
        ctx = self._ctx or default_context(self._device)
        info = ctx.emit_compressed(self.compressed())                        # raises the reference's CircuitError on a bad stream
        ins, outs = self.input_signals, self.output_signals
        in_names, out_names = self.signal_names(ins), self.signal_names(outs)
        # src/compiler.rs:327-383: input / output <=> node, walked in ascending signal id
        seen_in, seen_out = set(), set()
        nodes = ctx.emitted_signal_nodes(np.concatenate([ins, outs]))
        in_nodes, out_nodes = nodes[:len(ins)], nodes[len(ins):]
        merged = sorted([(int(s), 0, i) for i, s in enumerate(ins)] + [(int(s), 1, i) for i, s in enumerate(outs)])
        for _sid, kind, i in merged:
            name, seen = (in_names[i], seen_in) if kind == 0 else (out_names[i], seen_out)
            if name in seen:                                                 # :337-341, :347-351
                raise CircuitError(Status.INCONSISTENCY, f"Duplicate {'input' if kind == 0 else 'output'} {name}")
            seen.add(name)
        node_to_input = {int(nd): nm for nd, nm in zip(in_nodes, in_names)}
        for nd, nm in zip(out_nodes, out_names):                             # :363-383
            if int(nd) in node_to_input:
                raise CircuitError(Status.INCONSISTENCY, f"Node {int(nd)} used for both input {node_to_input[int(nd)]} and output {nm}")
        order, _wire, gates, wire_count = ctx.emitted_build_circuit(ins, outs, want_wires=False)
        consts = np.zeros((self._const_signals.shape[0], 4), dtype=np.uint32)   # (kind, signal id, value, 0) rows, ascending ids
        consts[:, 0], consts[:, 1], consts[:, 2] = 1, self._const_signals, self._const_values
        named = ctx.emitted_signal_wires(np.concatenate([ins, outs, consts[:, 1]]).astype(np.uint32))
        w_in, w_out, w_c = named[:len(ins)], named[len(ins):len(ins) + len(outs)], named[len(ins) + len(outs):]
        ci = CircuitInfo()
        ci.input_name_to_wire_index = {nm: int(w) for nm, w in sorted(zip(in_names, w_in.tolist()))}
        const_keys = [f"{nm}_{sid}" for nm, sid in zip(self.signal_names(consts[:, 1]), consts[:, 1].tolist())]   # :356
        for key, val, w in sorted(zip(const_keys, consts[:, 2].tolist(), w_c.tolist())):
            if w == 0xFFFFFFFF:
                raise CircuitError(Status.REFERENCE_PANIC, f"constant {key} has no wire (the reference panics at src/compiler.rs:473)")
            ci.constants[key] = ConstantInfo(str(int(val)), int(w))
        ci.output_name_to_wire_index = {nm: int(w) for nm, w in sorted(zip(out_names, w_out.tolist()))}
        self.emit_info = info
        return BristolCircuit(wire_count=int(wire_count), info=ci, gate_array=gates, order=order)


def compile(args, *, source: Optional[str] = None, include_dir: str = ".", device: int = 0, context: Optional[DeviceContext] = None,
            emitter: str = "host"):
    """src/program.rs::compile.  `args` is an Args or a path; `source=` compiles a string instead of a file.
    emitter="host": the walk drives the native union-find emitter (a Compiler with names, nodes and the report);
    emitter="device": the walk only records its calls and build_circuit() replays them on the GPU (a DeviceCompiler)."""
    if not isinstance(args, Args):
        args = Args(input=str(args)) if args is not None else Args()
    if emitter == "device":
        prog = lib.c2a_program_new()
        st = (lib.c2a_program_compile_source(prog, source.encode(), include_dir.encode(), None) if source is not None
              else lib.c2a_program_compile_file(prog, os.fsencode(args.input), None))
        if st != 0:
            text = lib.c2a_program_error(prog).decode()
            lib.c2a_program_free(prog)
            raise ProgramError(st, text)
        dc = DeviceCompiler(prog, device, context)
        dc.update_type(args.value_type)
        return dc
    comp = Compiler(device=device, context=context)
    prog = lib.c2a_program_new()
    try:
        if source is not None:
            st = lib.c2a_program_compile_source(prog, source.encode(), include_dir.encode(), comp._c)
        else:
            st = lib.c2a_program_compile_file(prog, os.fsencode(args.input), comp._c)
        if st != 0:
            raise ProgramError(st, lib.c2a_program_error(prog).decode())
        n = int(lib.c2a_program_num_events(prog))
        ev = np.zeros((n, 4), dtype=np.uint32)
        if n:
            C.memmove(ev.ctypes.data, lib.c2a_program_events(prog), 16 * n)
        comp.events = ev
        ni, no = int(lib.c2a_program_num_inputs(prog)), int(lib.c2a_program_num_outputs(prog))
        comp.input_signals = np.ctypeslib.as_array(C.cast(lib.c2a_program_inputs(prog), C.POINTER(C.c_uint32)), shape=(ni,)).copy() if ni else np.zeros(0, np.uint32)
        comp.output_signals = np.ctypeslib.as_array(C.cast(lib.c2a_program_outputs(prog), C.POINTER(C.c_uint32)), shape=(no,)).copy() if no else np.zeros(0, np.uint32)
    finally:
        lib.c2a_program_free(prog)
    comp.update_type(args.value_type)  # src/program.rs:71
    return comp


def generate_circuit_report(comp: Compiler) -> dict:
    """src/compiler.rs:287-319 + :503-531: inputs = nodes that no gate writes, outputs = written nodes that no gate reads,
    both ascending by node id; names skip the temporaries ('random_'), value = the last constant among the node's signals.
    Built natively (c2a_circuit_report_json: one pass over the gates, one over the nodes)."""
    text = lib.c2a_circuit_report_json(comp._c, comp.value_type.encode())
    return json.loads(text.decode())


def _generate_circuit_report_py(comp: Compiler) -> dict:
    """the same report assembled in Python from nodes() / gate_array() (kept as the cross-check of the native one)"""
    nodes = comp.nodes()
    gates = comp.gate_array()
    consumed = set(gates[:, 1].tolist()) | set(gates[:, 2].tolist()) if gates.shape[0] else set()
    ins = sorted(i for i, n in nodes.items() if not n["is_out"])
    outs = sorted(i for i, n in nodes.items() if n["is_out"] and i not in consumed)

    def rep(i):
        names, value = [], None
        for sid in nodes[i]["signals"]:
            name = comp.signal_name(sid)
            if "random_" not in name:
                names.append(name)
            v = comp.signal_value(sid)
            if v is not None:
                value = v
        return {"id": i, "names": names, "value": value}
    return {"inputs": [rep(i) for i in ins], "outputs": [rep(i) for i in outs], "value_type": comp.value_type}


def bristol_text(circ: BristolCircuit) -> str:
    """bristol-circuit `write_bristol` (un-vendored; PARITY UNPINNED): '<gates> <wires>', '<n_in> 1 ...', '<n_out> 1 ...',
    blank line, then one line per gate '2 1 <in0> <in1> <out> <Op>' with the strum Display op token."""
    g = circ.gate_array
    n_in = len(circ.info.input_name_to_wire_index) + len(circ.info.constants)
    n_out = len(circ.info.output_name_to_wire_index)
    lines = [f"{g.shape[0]} {circ.wire_count}", " ".join([str(n_in)] + ["1"] * n_in), " ".join([str(n_out)] + ["1"] * n_out), ""]
    g = np.ascontiguousarray(g, dtype=np.uint32)
    need = int(lib.c2a_bristol_gate_lines(g.ctypes.data_as(C.c_void_p), g.shape[0], None, 0))   # formatted natively: 30 MB per 1 M gates
    buf = C.create_string_buffer(need + 1)
    lib.c2a_bristol_gate_lines(g.ctypes.data_as(C.c_void_p), g.shape[0], buf, need)
    return "\n".join(lines) + "\n" + buf.raw[:need].decode()


def circuit_info_dict(circ: BristolCircuit) -> dict:
    return {"input_name_to_wire_index": dict(circ.info.input_name_to_wire_index),
            "constants": {k: {"value": v.value, "wire_index": v.wire_index} for k, v in circ.info.constants.items()},
            "output_name_to_wire_index": dict(circ.info.output_name_to_wire_index)}


def write_outputs(comp: Compiler, output_dir: str) -> BristolCircuit:
    """src/main.rs:21-47 — report first, then build_circuit (GPU), then the three files."""
    report = generate_circuit_report(comp)
    try:
        os.makedirs(output_dir, exist_ok=True)
    except OSError:
        raise ProgramError(0, "Output directory creation error")
    circ = comp.build_circuit()
    with open(build_output(output_dir, "circuit", "txt"), "w") as f:
        f.write(bristol_text(circ))
    with open(build_output(output_dir, "circuit_info", "json"), "w") as f:
        f.write(json.dumps(circuit_info_dict(circ), indent=2))
    with open(build_output(output_dir, "report", "json"), "w") as f:
        f.write(json.dumps(report, indent=2))
    return circ


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="circom-2-arithc", description="Arithmetic Circuits Compiler (B200-native back end)")
    ap.add_argument("-i", "--input", default="./input/circuit.circom", help="Path to the input file")
    ap.add_argument("-o", "--output", default="./output/", help="Path to the directory where the output will be written")
    ap.add_argument("-v", "--value-type", default="sint", choices=["sint", "sfloat"], help="Type that'll be used for values in MPC backend")
    ap.add_argument("--boolify-width", type=int, default=None, help="Optional: Convert to a boolean circuit by using integers with this number of bits")
    ap.add_argument("--device", type=int, default=0, help="CUDA device index")
    a = ap.parse_args(argv)
    if a.boolify_width is not None:
        raise SystemExit("--boolify-width: the boolify crate (voltrevo/boolify @ 6376405) is not part of the reference tree; not supported")
    try:
        comp = compile(Args(a.input, a.output, a.value_type, a.boolify_width), device=a.device)
        write_outputs(comp, a.output)
    except ProgramError as e:
        raise SystemExit(f"Error: {e}")
    return 0
