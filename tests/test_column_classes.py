"""The per-circuit column class tables (era_zkevm_circuits_b200/column_classes.py) against valid ORACLE traces of every circuit,
evaluated here with numpy: a mis-classified column would reject a valid trace."""
import numpy as np

from era_zkevm_circuits_b200 import abi, column_classes as CC
from trace_zoo import oracle_traces

BOUND = np.array([abi.GL_P, 2, 1 << 8, 1 << 16, 1 << 32], dtype=np.uint64)


def violations(cls, trace):
    return np.argwhere(trace >= BOUND[cls][:, None])


def test_every_table_accepts_valid_traces_and_is_tight(orc):
    zoo = oracle_traces(orc)
    assert set(zoo) == set(CC.TABLES)
    for name, trace in zoo.items():
        cls = CC.column_classes(name)
        bad = violations(cls, trace)
        assert bad.size == 0, (name, bad[:5].tolist(), [CC.CLASS_NAMES[cls[c]] for c, _ in bad[:5]])
        # tightness: a column classed U32 / FIELD really uses more than the next smaller class somewhere in the zoo (spot check:
        # at least a third of the U32 columns hold a value >= 2^16, i.e. the table is not all-FIELD lip service)
        wide = cls == CC.U32
        if wide.sum() > 8:
            assert (trace[wide].max(axis=1) >= (1 << 16)).mean() > 0.33, name
