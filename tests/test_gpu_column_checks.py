"""zkc_check_trace_columns (the generic allocation-check evaluator) on every circuit's trace: valid ORACLE traces pass, injected
faults of every class are found at their row / column, and on randomly corrupted traces the kernel counts exactly the rows the
numpy evaluation of the same class table counts (host and device buffers, even and odd row counts)."""
import numpy as np
import pytest

from era_zkevm_circuits_b200 import abi, column_classes as CC
from test_column_classes import violations
from trace_zoo import oracle_traces

pytestmark = pytest.mark.gpu


def test_every_circuit_trace(engine, orc):
    import torch
    rng = np.random.default_rng(11)
    for name, trace in oracle_traces(orc).items():
        cls = CC.column_classes(name)
        viol, st, col = CC.check_trace_columns(engine, name, trace)
        assert viol == 0 and st.code == 0, (name, viol, st.first_bad_row, col)
        dev = torch.from_numpy(trace.view(np.int64)).cuda()
        assert CC.check_trace_columns(engine, name, dev)[0] == 0
        odd = np.ascontiguousarray(trace[:, :trace.shape[1] - 1])          # odd row count: the 64-bit load path
        assert CC.check_trace_columns(engine, name, odd)[0] == 0
        # one fault per class present in the table
        for k in range(5):
            cols = np.flatnonzero(cls == k)
            if not len(cols):
                continue
            c, r = int(cols[len(cols) // 2]), int(rng.integers(0, trace.shape[1]))
            bad = trace.copy()
            bad[c, r] = np.uint64([abi.GL_P, 2, 1 << 8, 1 << 16, 1 << 32][k])
            viol, st, col = CC.check_trace_columns(engine, name, bad)
            assert viol == 1 and st.code == abi.ZKC_ERR_UNSATISFIED and st.first_bad_row == r and col == c and st.failed_checks == 1 << k, (name, k, c, r)
        # random corruption: the same rows as numpy
        bad = trace.copy()
        cs, rs = rng.integers(0, trace.shape[0], 200), rng.integers(0, trace.shape[1], 200)
        bad[cs, rs] = rng.integers(0, 1 << 63, 200, dtype=np.uint64) * np.uint64(2)
        want = violations(cls, bad)
        viol, st, col = CC.check_trace_columns(engine, name, torch.from_numpy(bad.view(np.int64)).cuda())
        assert viol == len(set(want[:, 1].tolist())), name
        if len(want):
            first_row = int(want[:, 1].min())
            assert st.first_bad_row == first_row and col == int(want[want[:, 1] == first_row][:, 0].min())
