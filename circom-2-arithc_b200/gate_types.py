"""AGateType (src/a_gate_type.rs:7-28, discriminants in declaration order) and the c2a_event kinds (include/c2a.h).
Pure Python, no native import: bench.py's reference arm generates its workload from workloads.py through this module without
mapping libc2a.so into the process."""
import enum

EV_SIGNAL, EV_SIGNAL_CONST, EV_GATE, EV_CONNECT = 0, 1, 2, 3


class AGateType(enum.IntEnum):
    AAdd = 0
    ADiv = 1
    AEq = 2
    AGEq = 3
    AGt = 4
    ALEq = 5
    ALt = 6
    AMul = 7
    ANeq = 8
    ASub = 9
    AXor = 10
    APow = 11
    AIntDiv = 12
    AMod = 13
    AShiftL = 14
    AShiftR = 15
    ABoolOr = 16
    ABoolAnd = 17
    ABitOr = 18
    ABitAnd = 19

    def __str__(self):  # strum Display: the Bristol op token
        return self.name


