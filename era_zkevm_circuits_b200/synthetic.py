"""Synthetic, *valid* circuit inputs.  Host-side input preparation only (numpy);
in production these arrays come from the out-of-circuit VM run.  Values are built by constructing a
legal execution trace, never by sampling cells independently, so every enforcement of the
reference holds on them."""
import numpy as np

from . import abi

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (vectorised; counter-based so streams are independent)"""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64((stream * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF)
        z = z + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def ram_trace(n: int, seed: int = 0xC1, n_cells: int = 1 << 10, n_nondet: int = 0,
              heap_page: int = 10):
    """C1: n MemoryQuery records over n_cells cells; returns (unsorted, sorted) record arrays.
    timestamps strictly increasing, first access of a cell is a write, later accesses 50 % reads
    (value = last write) / 50 % writes (256-bit uniform), is_ptr on 1 % of the writes;
    sorted order = (page, index, timestamp), the order ram_permutation/mod.rs:296-316 enforces.
    n_nondet extra bootloader-heap writes with timestamp 0 lead the unsorted queue (:260-290)."""
    m = n - n_nondet
    assert m >= 0
    q = np.zeros(n, dtype=abi.MEMORY_QUERY_DTYPE)
    r = splitmix64(seed, m, 0)
    n_cells = max(1, min(n_cells, max(m, 1)))
    cell = (r % np.uint64(n_cells)).astype(np.int64)
    ts = np.arange(1, m + 1, dtype=np.uint32)
    order = np.lexsort((ts, cell))  # by cell then time
    sc = cell[order]
    first = np.ones(m, dtype=bool)
    first[1:] = sc[1:] != sc[:-1]
    coin = (splitmix64(seed, m, 1) >> np.uint64(63)).astype(bool)
    is_write_sorted = first | coin[order]
    ptr_coin = (splitmix64(seed, m, 2) % np.uint64(100)) == 0
    vals = np.stack([splitmix64(seed, m, 3 + k) for k in range(4)], axis=1).view("<u4").reshape(m, 8)
    pos = np.where(is_write_sorted, np.arange(m), 0)
    last_write = np.maximum.accumulate(pos)  # first element of every cell group is a write
    src = order[last_write]  # original index of the governing write
    body = q[n_nondet:]
    body["timestamp"] = ts
    body["memory_page"] = 8 + (cell // 64) % 16 + 16 * (cell // 1024)
    body["index"] = cell % 64
    is_write = np.zeros(m, dtype=bool)
    is_write[order] = is_write_sorted
    body["rw_flag"] = is_write
    value = np.empty((m, 8), dtype=np.uint32)
    value[order] = vals[src]
    body["value"] = value
    is_ptr = np.zeros(m, dtype=bool)
    is_ptr[order] = ptr_coin[src]
    body["is_ptr"] = is_ptr
    if n_nondet:
        head = q[:n_nondet]
        head["timestamp"] = 0
        head["memory_page"] = heap_page
        head["index"] = 100000 + np.arange(n_nondet, dtype=np.uint32)
        head["rw_flag"] = 1
        head["value"] = splitmix64(seed, n_nondet * 4, 9).view("<u4").reshape(n_nondet, 8)
    srt = np.lexsort((q["timestamp"], q["index"], q["memory_page"]))
    return q, q[srt]
