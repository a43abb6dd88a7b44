"""circom-2-arithc_b200 — B200-native gate-graph builder / topological sorter.

Host-side mirror of the reference's operator interface for the flattening hot path
(reference: src/compiler.rs `Compiler`, src/a_gate_type.rs `AGateType`, src/topological_sort.rs) on top of the
C ABI in include/c2a.h (libc2a.so, built from csrc/).  The directory name carries a hyphen, so import it
through the repo-root helper:  ``from c2a_loader import c2a``.
"""
from ._lib import lib, load_error, C2AError, CircuitError, Status, have_device  # noqa: F401
from .compiler import AGateType, Compiler, BristolCircuit, Gate, DeviceContext, topological_sort, pack_events, unpack_events  # noqa: F401
from . import workloads  # noqa: F401
from . import sharding  # noqa: F401
from .program import compile, Args, ProgramError, DeviceCompiler, generate_circuit_report, write_outputs, bristol_text, main as cli_main  # noqa: F401,E402
