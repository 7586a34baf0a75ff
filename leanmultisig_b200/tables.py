"""Static description of the three lean_vm tables as far as Logup and the AIR sessions need it
(crates/lean_vm/src/tables/{table_trait.rs:19-56, execution/mod.rs:27-58, extension_op/mod.rs:91-124,
poseidon_16/mod.rs:140-182}, table ordering: tables/table_enum.rs:7-16)."""
from __future__ import annotations

from dataclasses import dataclass, field

LOGUP_MEMORY_DOMAINSEP, LOGUP_PRECOMPILE_DOMAINSEP, LOGUP_BYTECODE_DOMAINSEP = 0, 1, 2  # core/constants.rs:4-6
N_RUNTIME_COLUMNS, N_INSTRUCTION_COLUMNS = 8, 12                                          # execution/air.rs:4-6
COL_PC = 0


@dataclass(frozen=True)
class Lookup:
    index: int           # column holding the memory address
    values: tuple        # columns holding memory[address + i]


@dataclass(frozen=True)
class Bus:
    pull: bool           # BusDirection::Pull (numerator -selector) or Push (+selector)
    selector: int
    data: tuple          # column indices (BusData::Column)


@dataclass(frozen=True)
class Table:
    name: str
    air_id: int          # table_id of lm_air_new
    order: int           # position in the Table enum (ties in the height sort keep this order)
    n_columns: int       # AIR columns
    n_columns_total: int  # + virtual columns kept for Logup
    bus: Bus
    lookups: tuple
    is_execution: bool = False


EXECUTION = Table("execution", 0, 0, 20, 24, Bus(False, 20, (19, 21, 22, 23)),
                  (Lookup(2, (5,)), Lookup(3, (6,)), Lookup(4, (7,))), True)
EXTENSION_OP = Table("extension_op", 1, 1, 29, 31, Bus(True, 29, (30, 6, 7, 13)),
                     (Lookup(6, tuple(range(14, 19))), Lookup(7, tuple(range(19, 24))), Lookup(13, tuple(range(24, 29)))))
POSEIDON16 = Table("poseidon16", 2, 2, 109, 111, Bus(True, 0, (110, 109, 1, 2)),
                   (Lookup(6, tuple(range(9, 13))), Lookup(7, tuple(range(13, 17))), Lookup(1, tuple(range(17, 25))),
                    Lookup(2, tuple(range(93, 109)))))
ALL_TABLES = (EXECUTION, EXTENSION_OP, POSEIDON16)


@dataclass
class TableTrace:
    """lean_vm::TableTrace: columns[c] = numpy uint32 (Montgomery) of 2^log_n_rows entries"""
    columns: list
    log_n_rows: int
    non_padded_n_rows: int = 0


def sort_tables_by_height(log_heights: dict) -> list:
    """table_trait.rs:66-70: stable sort, tallest first"""
    return sorted(log_heights.items(), key=lambda kv: (-kv[1], kv[0].order))


def offset_for_table(table: Table, log_n_rows: int) -> int:
    return (sum(len(l.values) for l in table.lookups) + 1) << log_n_rows


def compute_total_active_len(log_memory: int, log_bytecode: int, tables_sorted: list) -> int:
    """logup.rs:500-518"""
    max_table_height = 1 << tables_sorted[0][1]
    log_n_cycles = next(h for t, h in tables_sorted if t.is_execution)
    return ((1 << log_memory) + max(1 << log_bytecode, max_table_height) + (1 << log_n_cycles)
            + sum(offset_for_table(t, h) for t, h in tables_sorted))
